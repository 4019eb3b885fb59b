# Round 2, re-entry call 1 (1 GPU): the whole -m gpu suite, the default bench line, then the ncu evidence of the BENCHED
# configuration (tools/gpu_r2_profile.sh).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv,noheader
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1; tail -5 gpurun_out/pytest_gpu.log
( time timeout 600 python bench.py --gpus 1 --steps 3 --warmup 3 ) > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -c 1500 gpurun_out/bench_n1.json; tail -5 gpurun_out/bench_n1.err
bash tools/gpu_r2_profile.sh
