# round-1 second evidence run (1 GPU): tests, bench, one ncu full capture of a mid-size wavefront (register-scan kernel)
mkdir -p gpurun_out
timeout 150 python -m pytest tests -q -x -m gpu 2>&1 | tail -3
timeout 150 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_r1b.json 2> gpurun_out/bench_r1b.err; tail -c 1800 gpurun_out/bench_r1b.json; tail -3 gpurun_out/bench_r1b.err
MAXDEPTH=20 timeout 80 ncu --set full --clock-control none --import-source on -k regex:expand_kernel -s 17 -c 1 -o gpurun_out/prof_expand_r1f python tools/quick.py 3 2 2 0 0 134217728 8388608 > gpurun_out/prof.log 2>&1; tail -1 gpurun_out/prof.log | cut -c1-200
