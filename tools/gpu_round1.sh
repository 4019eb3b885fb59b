# round-1 evidence run (1 GPU): tests, bench, reference arm, ncu launch list, ncu full capture of two mid-size wavefronts
mkdir -p gpurun_out
python -m pytest tests -q -m gpu 2>&1 | tail -2
python bench.py --steps 3 --warmup 3 > gpurun_out/bench_r1.json 2> gpurun_out/bench_r1.err; tail -c 600 gpurun_out/bench_r1.json
python bench.py --impl reference --steps 1 --warmup 1 | tail -1 | cut -c1-300
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/b_ncu.log 2>&1
MAXDEPTH=20 timeout 600 ncu --set full --clock-control none --import-source on -k regex:expand_kernel -s 17 -c 2 -o gpurun_out/prof_expand_r1 python tools/quick.py 3 2 2 0 0 134217728 8388608 > gpurun_out/prof.log 2>&1; tail -1 gpurun_out/prof.log | cut -c1-120
python - <<'PY'
import ctypes as C, sys
sys.path.insert(0, '.')
import _pkg
pkg = _pkg.load()
lib = pkg.load_library()
out = (C.c_double * 3)()
for cap_log, n in [(30, 1 << 28)]:
    rc = lib.vsr_probe_bench(0, 1 << cap_log, n, 0.5, 3, out)
    print("probe_bench rc", rc, "capacity 2^%d keys %d dup 0.5: %.3f ms, %.0f new, %.0f probes -> %.2f G inserts/s, %.1f GB/s of 32-B sectors" % (cap_log, n, out[0], out[1], out[2], n / out[0] / 1e6, out[2] * 32 / out[0] / 1e6))
PY
