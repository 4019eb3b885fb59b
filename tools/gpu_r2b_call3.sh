# Round 2, re-entry call 3 (1 GPU): the GPU suite with the new owner rule and the checkpoint tests, then the 8-rank
# configuration that overflowed on 8 GPUs (same -table / -frontier, eight threads sharing this GPU through the test hook).
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1; tail -15 gpurun_out/pytest_gpu.log
python -c "
import _pkg; pkg=_pkg.load(); open('gpurun_out/cfg2.cfg','w').write(pkg.cfg_text(3, ['v1','v2'], 2))"
( time VSR_B200_MULTI_ONE_DEVICE=1 timeout 300 vsr-tlaplus_b200/vsrmc -deadlock -continue -notrace -gpus 8 -table 536870912 -frontier 22000000 -config gpurun_out/cfg2.cfg ) 2>&1 | grep -v "^depth" | tail -12 | tee gpurun_out/vsrmc_8ranks_1gpu.txt
