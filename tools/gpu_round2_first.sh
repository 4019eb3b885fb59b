# First GPU call of the next round (1 GPU).  Before calling:  tools/variants.sh   (builds build/variants/*.so here; they travel)
#   gpurun --timeout 900 -- 'bash tools/gpu_round2_first.sh'
# 1. the whole GPU suite, including the tests written after round 1's GPU budget ran out (plug-in layout on the GPU, ranks
#    sharing one GPU): their xfail marks come off once they have passed here
# 2. the pool-overflow path under the parity tests (variant qps1: one pool entry per parent state)
# 3. A/B of the kernel experiments on the shipped VSR.cfg (complete BFS, three runs each)
mkdir -p gpurun_out
python -m pytest tests -q -m gpu -rxX 2>&1 | tail -15
VSR_B200_LIB=$PWD/build/variants/libvsr_b200_qps1.so timeout 300 python -m pytest tests/test_gpu_parity.py -q -k "3-2-2 or deterministic or full_size" 2>&1 | tail -3
bash tools/ab.sh 2>&1 | tee gpurun_out/ab_round2.txt | tail -30
