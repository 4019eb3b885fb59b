# Round 2, re-entry (1 GPU): the new block shape (one block of up to 32 warps per SM) — the whole GPU suite, the default bench
# line, the variant A/B (old shape, invariant skip), cfg3 / cfg4 bounded runs for their kernel rates.
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1; tail -6 gpurun_out/pytest_gpu.log
( time timeout 600 python bench.py --gpus 1 --steps 3 --warmup 3 ) > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -3 gpurun_out/bench_n1.err
python - <<'PY'
import json
try:
    d = json.loads([l for l in open("gpurun_out/bench_n1.json") if l.startswith("{")][-1])
    print("value %.3e" % d["value"], "ms/step %.1f" % d["ms_per_step"], "kernel_s/step %.4f" % (d["kernel_seconds"] / d["steps"]), "frac %.4f" % d["roofline"]["frac"],
          "probe frac %.3f" % d["probe_roofline"]["frac"], "ok" if d["config"]["results_match_expected"] else "RESULTS DIFFER")
    print("e2e", d["e2e"])
except Exception as e:
    print("failed:", e)
PY
bash tools/ab.sh 3 2 2 4294967296 140000000 2>&1 | tee gpurun_out/ab_round2d.txt
echo "== cfg3 / cfg4 bounded (kernel rate of the wider layouts)"
MAXDEPTH=21 QUIET=1 python tools/quick.py 3 3 3 0 0 1600000000 300000000 | head -1 | cut -c1-400 | tee gpurun_out/quick_cfg3_d21.txt
MAXDEPTH=12 QUIET=1 python tools/quick.py 5 2 2 0 0 400000000 160000000 | head -1 | cut -c1-400 | tee gpurun_out/quick_cfg4_d12.txt
