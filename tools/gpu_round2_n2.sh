# Second GPU call of the next round (2 GPUs):  gpurun --gpus 2 --timeout 600 -- 'bash tools/gpu_round2_n2.sh'
# Question: at N=2 the kernel time is 84 % of N=1 although each rank expands half the frontier.
# Reads the split the bench line now carries (kernel_seconds vs kernel_seconds_insert, records_sent_rank0,
# phase_seconds_rank0_last_step) for three sizes of the sender-side duplicate filter (1/8, 1/2, 1/1 of the seen-set's slots).
mkdir -p gpurun_out
for div in 8 2 1; do
  VSR_SENT_FILTER_DIV=$div timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
      --master-port 29531 bench.py --gpus 2 --steps 2 --warmup 3 2>gpurun_out/n2_div$div.err | tail -1 > gpurun_out/n2_div$div.json
  python - "$div" <<'PY'
import json, sys
div = sys.argv[1]
try:
    d = json.load(open("gpurun_out/n2_div%s.json" % div))
    gen = d["config"]["states_generated"]
    print("div", div, "value %.3e" % d["value"], "ms/step %.1f" % d["ms_per_step"], "kernel_s/step %.3f" % (d["kernel_seconds"] / d["steps"]),
          "insert_s/step %.3f" % (d["kernel_seconds_insert"] / d["steps"]),
          "sent by rank 0 / its remote successors %.3f" % (d["records_sent_rank0"] / d["steps"] / (gen / 4.0)),
          d["phase_seconds_rank0_last_step"], "ok" if d["config"]["results_match_expected"] else "RESULTS DIFFER")
except Exception as e:
    print("div", div, "failed:", e)
PY
done
