# Round 2, re-entry (8 GPUs, charged 8x: kept short): the bench line at N=8 with the README-constants first-violation block,
# then vsrmc -gpus 8 (one process, one thread per GPU) on the shipped VSR.cfg.
mkdir -p gpurun_out
nvidia-smi topo -m 2>/dev/null | head -11; free -g | head -2
( time timeout 330 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29538 bench.py --gpus 8 --steps 2 --warmup 1 2>gpurun_out/n8.err | tail -1 > gpurun_out/bench_n8.json ) 2>&1 | grep real
tail -c 1800 gpurun_out/bench_n8.json; echo; grep -v "^$" gpurun_out/n8.err | tail -12
python -c "
import _pkg; pkg=_pkg.load(); open('gpurun_out/cfg2.cfg','w').write(pkg.cfg_text(3, ['v1','v2'], 2))"
( time timeout 120 vsr-tlaplus_b200/vsrmc -deadlock -continue -notrace -gpus 8 -table 536870912 -frontier 22000000 -config gpurun_out/cfg2.cfg ) 2>&1 | grep -v "^depth" | tail -12 | tee gpurun_out/vsrmc_n8.txt
