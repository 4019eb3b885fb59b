# Round 2, GPU call 1 (1 GPU): the GPU suite with xfail marks ignored (tracebacks of the multi-rank test that
# xfailed at the end of round 1), then the A/B of the prepared kernel variants on the shipped VSR.cfg.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv,noheader
timeout 600 python -m pytest tests -q -m gpu --runxfail -x --deselect tests/test_dist_gloo.py 2>&1 | tail -5
timeout 400 python -m pytest tests/test_dist_gloo.py -q -m gpu --runxfail 2>&1 | tail -60 > gpurun_out/dist_gpu_test.log; tail -40 gpurun_out/dist_gpu_test.log
bash tools/ab.sh 2>&1 | tee gpurun_out/ab_round2.txt | tail -40
