mkdir -p gpurun_out
python -m pytest tests -q -m gpu 2>&1 | tail -4
python bench.py --steps 3 --warmup 3 > gpurun_out/bench_r1.json 2> gpurun_out/bench_r1.err; tail -c 3000 gpurun_out/bench_r1.json; tail -3 gpurun_out/bench_r1.err
python bench.py --impl reference --steps 1 --warmup 1 | tail -1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/b_ncu.log 2>&1; tail -2 gpurun_out/b_ncu.log | cut -c1-300
MAXDEPTH=20 timeout 600 ncu --set full --clock-control none --import-source on -k regex:expand_kernel -s 17 -c 2 -o gpurun_out/prof_expand_r1 python tools/quick.py 3 2 2 0 0 134217728 8388608 > gpurun_out/prof.log 2>&1; tail -3 gpurun_out/prof.log | cut -c1-300
ls -la gpurun_out
