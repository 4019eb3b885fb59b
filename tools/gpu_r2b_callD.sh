# Round 2, re-entry (1 GPU), the part that pins a lot of host memory (this box gives a job 200 GiB: profiles/host_memory_1gpu_box.txt):
# (a) BASELINE configs[3] to depth 13 with a frontier that does NOT fit its HBM part (330 M states) and spills 266 M states into
#     pinned host memory: counts must equal the in-HBM run (profiles/cfg4/depth13_one_gpu.txt), rate beside it;
# (b) the README constants on ONE GPU to the depth-24 violation (bench.py --cfg3-one-gpu: 2 x 54 GB pinned).
mkdir -p gpurun_out
python -c "
import _pkg; pkg=_pkg.load(); open('gpurun_out/cfg4.cfg','w').write(pkg.cfg_text(5, ['v1','v2'], 2))"
echo "== cfg4 depth 13, 330 M states in HBM + 300 M in pinned host memory per frontier buffer"
( time timeout 300 vsr-tlaplus_b200/vsrmc -deadlock -notrace -depth 13 -table 1300000000 -frontier 330000000 -spill 300000000 -config gpurun_out/cfg4.cfg ) 2>&1 | tail -9 | tee gpurun_out/cfg4_depth13_spill.txt
echo "== README constants on one GPU"
( time timeout 600 python bench.py --gpus 1 --steps 1 --warmup 1 --no-cpu-baseline --cfg3-one-gpu ) > gpurun_out/bench_n1_cfg3.json 2> gpurun_out/bench_n1_cfg3.err; tail -3 gpurun_out/bench_n1_cfg3.err
python - <<'PY'
import json
try:
    d = json.loads([l for l in open("gpurun_out/bench_n1_cfg3.json") if l.startswith("{")][-1])
    c = d.get("cfg3_first_violation") or {}
    print({k: c[k] for k in c if k not in ("counterexample_actions", "golden_state_depths")})
except Exception as e:
    print("failed:", e)
PY
