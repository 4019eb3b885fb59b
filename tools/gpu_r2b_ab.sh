# Round 2, re-entry (1 GPU): A/B of the kernel variants built by tools/variants.sh at the benched sizes (2^32-slot seen-set, 140 M-state
# frontier buffers), three complete BFS runs each; then the parity tests on the variants that change the block shape.
mkdir -p gpurun_out
bash tools/ab.sh 3 2 2 4294967296 140000000 2>&1 | tee gpurun_out/ab_round2c.txt
for v in warps32pl pl; do
  echo "== parity with $v"; VSR_B200_LIB=build/variants/libvsr_b200_$v.so timeout 600 python -m pytest tests/test_gpu_parity.py -q -k "3-2-2 or deterministic or shipped_cfg" 2>&1 | tail -3
done | tee gpurun_out/ab_round2c_parity.txt
