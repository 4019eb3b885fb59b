# Round 2, re-entry (2 GPUs): the bench line at N=2 with the README-constants block (2 x 152 GB) and the per-part wall clock
# of the one-call API.
mkdir -p gpurun_out
( time timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 3 --warmup 2 2>gpurun_out/n2.err | tail -1 > gpurun_out/bench_n2.json ) 2>&1 | grep real
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/bench_n2.json"))
    print("value %.3e" % d["value"], "ms/step %.1f" % d["ms_per_step"], "kernel_s/step %.4f" % (d["kernel_seconds"] / d["steps"]), "launches", d["gpu_launches"],
          "ok" if d["config"]["results_match_expected"] else "RESULTS DIFFER")
    print("e2e", d["e2e"])
    c = d.get("cfg3_first_violation") or {}
    print({k: c[k] for k in c if k not in ("counterexample_actions", "golden_state_depths")})
except Exception as e:
    print("failed:", e)
PY
grep -v "^$" gpurun_out/n2.err | grep -v "OMP_NUM\|\*\*\*\*" | tail -8
