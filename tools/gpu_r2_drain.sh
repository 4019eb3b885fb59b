# Round 2 (1 GPU): what a drained record costs, for 1 / 2 / 4 records in flight per lane (tools/drain_bench.py emulates one rank
# of a 2-GPU job on one GPU), and that the expand path did not pay for it (plain one-GPU BFS with each library).
mkdir -p gpurun_out
for v in base; do
  lib=$PWD/build/variants/libvsr_b200_$v.so
  QUIET=1 VSR_B200_LIB=$lib python tools/quick.py 3 2 2 0 0 4294967296 140000000 | head -1 | python -c "
import json,sys; d=json.loads(sys.stdin.readline()); print('$v', 'N=1 kern %.4f' % d['kern'], d['distinct'])"
  VSR_B200_LIB=$lib python tools/drain_bench.py 3 2 2 2>&1 | tail -1 | sed "s/^/$v emulated-rank /"
done | tee gpurun_out/drain_ab.txt
