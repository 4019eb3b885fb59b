# Round 2, re-entry (1 GPU): the default bench line with the README-constants block on ONE GPU (frontier spill to pinned host
# memory), then BASELINE configs[3] (tools/gpu_r2b_cfg4.sh).
mkdir -p gpurun_out
( time timeout 900 python bench.py --gpus 1 --steps 3 --warmup 3 ) > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -3 gpurun_out/bench_n1.err
python - <<'PY'
import json
try:
    d = json.loads([l for l in open("gpurun_out/bench_n1.json") if l.startswith("{")][-1])
    print("value %.3e" % d["value"], "ms/step %.1f" % d["ms_per_step"], "kernel_s/step %.4f" % (d["kernel_seconds"] / d["steps"]), "frac %.4f" % d["roofline"]["frac"],
          "ok" if d["config"]["results_match_expected"] else "RESULTS DIFFER")
    print("e2e", d["e2e"])
    c = d.get("cfg3_first_violation") or {}
    print({k: c[k] for k in c if k not in ("counterexample_actions", "golden_state_depths")})
except Exception as e:
    print("failed:", e)
PY
bash tools/gpu_r2b_cfg4.sh
