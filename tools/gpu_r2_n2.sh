# Round 2 (2 GPUs): the fused exchange on NVLink: bench line, direct-store push A/B, and the staged NCCL baseline it replaces.
mkdir -p gpurun_out
run() { # tag, env, extra args
  env $2 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 3 --warmup 2 $3 2>gpurun_out/n2_$1.err | tail -1 > gpurun_out/n2_$1.json
  python - "$1" <<'PY'
import json, sys
tag = sys.argv[1]
try:
    d = json.load(open("gpurun_out/n2_%s.json" % tag))
    print(tag, "value %.3e" % d["value"], "ms/step %.1f" % d["ms_per_step"], "kernel_s/step %.4f" % (d["kernel_seconds"] / d["steps"]),
          "drain-only s/step %.4f" % (d["kernel_seconds_insert"] / d["steps"]), "sent0/step %d" % (d["records_sent_rank0"] / d["steps"]),
          "e2e", d["e2e"] and round(d["e2e"]["value"] / 1e9, 3), "ok" if d["config"]["results_match_expected"] else "RESULTS DIFFER")
except Exception as e:
    print(tag, "failed:", e)
    print(open("gpurun_out/n2_%s.err" % tag).read()[-1500:])
PY
}
run p2p "X=1" ""
run direct "VSR_B200_PUSH=direct" "--no-e2e"
run staged "X=1" "--exchange staged"
