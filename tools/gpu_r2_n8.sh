# Round 2 (8 GPUs): bench line at N=8 (with the README-constants first-violation block), then BASELINE configs[3]
# (ReplicaCount=5) sharded over the box to depth 14.
mkdir -p gpurun_out
nvidia-smi topo -m 2>/dev/null | head -12; free -g | head -2
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29538 bench.py --gpus 8 --steps 3 --warmup 2 2>gpurun_out/n8.err | tail -1 > gpurun_out/n8.json
tail -c 2500 gpurun_out/n8.json; echo; tail -25 gpurun_out/n8.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29539 tools/hunt.py 5 2 2 --table 1073741824 --frontier 520000000 --inbox 6000000 --depth 14 --continue-past --out gpurun_out 2>gpurun_out/cfg4_n8.err | tail -1 | cut -c1-1500; tail -22 gpurun_out/cfg4_n8.err
python -c "
import _pkg; pkg=_pkg.load(); open('gpurun_out/cfg2.cfg','w').write(pkg.cfg_text(3, ['v1','v2'], 2))"
( time vsr-tlaplus_b200/vsrmc -deadlock -continue -notrace -gpus 8 -table 536870912 -frontier 22000000 -config gpurun_out/cfg2.cfg ) 2>&1 | grep -v "^depth" | tail -12
