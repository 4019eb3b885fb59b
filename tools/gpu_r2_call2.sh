# Round 2, GPU call 2 (1 GPU): the whole GPU suite (multi-rank engine paths with 2/4/8 processes sharing the GPU: CUDA IPC
# inboxes + C++ level loop, and the torch.distributed-staged pump), seen-set bucket A/B, table-size A/B, a short bench line.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -25
echo "== A/B (default table 2^31 slots auto... quick.py auto-sizes)"
bash tools/ab.sh 3 2 2 2147483648 140000000 2>&1 | tee gpurun_out/ab_round2b.txt | grep libvsr
echo "== base with 2^32 slots"
for i in 1 2; do QUIET=1 VSR_B200_LIB=$PWD/build/variants/libvsr_b200_base.so python tools/quick.py 3 2 2 0 0 4294967296 140000000 | head -1 | cut -c1-400; done | tee gpurun_out/ab_table32.txt
for i in 1 2; do QUIET=1 VSR_B200_LIB=$PWD/build/variants/libvsr_b200_bucket2.so python tools/quick.py 3 2 2 0 0 4294967296 140000000 | head -1 | cut -c1-400; done | tee -a gpurun_out/ab_table32.txt
echo "== bench"
timeout 600 python bench.py --steps 2 --warmup 1 --cpu-seconds 5 > gpurun_out/bench_r2_n1_first.json 2> gpurun_out/bench_r2_n1_first.err; tail -c 1500 gpurun_out/bench_r2_n1_first.json; tail -5 gpurun_out/bench_r2_n1_first.err
