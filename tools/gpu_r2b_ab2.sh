# Round 2, re-entry (1 GPU): A/B of the CAS-first seen-set access against the default, then parity on the variant.
mkdir -p gpurun_out
bash tools/ab.sh 3 2 2 4294967296 140000000 2>&1 | tee gpurun_out/ab_round2e.txt
for v in casfirst; do
  echo "== parity with $v"; VSR_B200_LIB=build/variants/libvsr_b200_$v.so timeout 300 python -m pytest tests/test_gpu_parity.py -q -k "3-2-2 or deterministic or view_ties" 2>&1 | tail -3
done | tee gpurun_out/ab_round2e_parity.txt
