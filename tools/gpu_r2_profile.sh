# Round 2 (1 GPU): ncu evidence of the BENCHED configuration (2^32-slot seen-set = 64 GiB, 140 M-state frontiers).
# 1. launch list of one bench step (share of each kernel);  2. --set full of ONE wide wavefront (expanding depth 31:
#    117.6 M states -> 120.2 M new) in application-replay mode (kernel replay would have to save/restore 64 GiB per pass).
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r2b.csv python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-e2e > gpurun_out/b_ncu.log 2>&1; tail -c 300 gpurun_out/b_ncu.log; wc -l gpurun_out/launches_r2b.csv
MAXDEPTH=32 QUIET=1 timeout 1500 ncu --set full --replay-mode application --clock-control none --import-source on -k regex:expand_kernel -s 30 -c 1 -f -o gpurun_out/prof_expand_r2b python tools/quick.py 3 2 2 0 0 4294967296 140000000 > gpurun_out/prof.log 2>&1; tail -2 gpurun_out/prof.log | cut -c1-300
ncu -i gpurun_out/prof_expand_r2b.ncu-rep --page raw --csv > gpurun_out/prof_expand_r2b_raw.csv 2>/dev/null; wc -c gpurun_out/prof_expand_r2b_raw.csv gpurun_out/prof_expand_r2b.ncu-rep
