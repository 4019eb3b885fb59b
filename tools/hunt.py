#!/usr/bin/env python
"""Sharded BFS of one configuration until the first invariant violation (BASELINE configs[4]: time-to-first
counterexample), under torchrun or alone.  Prints the level table and the verdict as JSON, writes the counterexample in
TLC `dumpTrace tlc` format, and cross-checks the reference's published 24-state behaviour
(tests/golden/state_transfer_trace.json) against the explored set: every published state must have been seen at a BFS
depth <= its position in the published trace.

  torchrun --nproc-per-node 8 tools/hunt.py 3 3 3 --table 1073741824 --frontier 200000000
"""
import argparse
import base64
import json
import os
import sys
import time
import zlib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("R", type=int)
    ap.add_argument("V", type=int)
    ap.add_argument("L", type=int)
    ap.add_argument("--table", type=int, default=0)
    ap.add_argument("--frontier", type=int, default=0)
    ap.add_argument("--inbox", type=int, default=0, help="records per inbox segment (0 = from --frontier)")
    ap.add_argument("--part", type=int, default=0, help="frontier states per step and rank (0 = from the inbox size)")
    ap.add_argument("--depth", type=int, default=0)
    ap.add_argument("--seconds", type=float, default=0)
    ap.add_argument("--continue-past", action="store_true")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out"))
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    import torch
    import torch.distributed as tdist
    import _pkg
    pkg = _pkg.load()
    from vsr_tlaplus_b200 import dist as vdist

    if world > 1:
        torch.cuda.set_device(local)
        tdist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    group = vdist.Group.from_torch() if world > 1 else None
    mc = pkg.ModelChecker.from_constants(args.R, args.V, args.L)
    t0 = time.time()
    eng = vdist.GpuEngine(mc, rank, world, device=local, table_capacity=args.table, frontier_capacity=args.frontier,
                          inbox_records=args.inbox, keep_trace=True, group=group)
    t1 = time.time()
    res = eng.run(max_depth=args.depth, max_seconds=args.seconds, stop_on_violation=not args.continue_past, part_states=args.part,
                  verbose=True)
    torch.cuda.synchronize(dev)
    t2 = time.time()

    # golden cross-check (only meaningful for the README constants R=3, V=3, L=3)
    golden = None
    gpath = os.path.join(ROOT, "tests", "golden", "state_transfer_trace.json")
    if (args.R, args.V, args.L) == (3, 3, 3) and os.path.exists(gpath):
        fx = json.load(open(gpath))
        Flat = pkg.checker.VsrFlatState
        levels = []
        for s in fx["states"]:
            flat = Flat.from_buffer_copy(zlib.decompress(base64.b64decode(s["flat_zlib_b64"])))
            packed = mc.pack(flat)  # canonical labels
            lvl, owner = eng.lookup(packed)
            levels.append(lvl if owner == rank else 0)
        t = torch.tensor(levels, dtype=torch.int64, device=dev)
        if world > 1:
            tdist.all_reduce(t, op=tdist.ReduceOp.MAX)
        golden = [int(x) for x in t.cpu().tolist()]

    if rank == 0:
        os.makedirs(args.out, exist_ok=True)
        trace = vdist.replay_trace(mc, res.trace_cands) if res.rc in (11, 12) else []
        tag = "r%dv%dl%d_g%d" % (args.R, args.V, args.L, world)
        if trace:
            with open(os.path.join(args.out, "trace_%s.txt" % tag), "w") as f:
                f.write(mc.dump_trace_tlc(trace))
        out = dict(config=[args.R, args.V, args.L], gpus=world, rc=res.rc, distinct=res.distinct, generated=res.generated,
                   depth=res.depth, complete=res.complete, queue=res.queue, violation_level=res.violation_level,
                   h2_ties=res.h2_ties, fp_collisions=res.fp_collisions, seconds_setup=t1 - t0, seconds_bfs=t2 - t1,
                   kernel_seconds_max=res.kernel_ms_max / 1e3, states_per_second=res.distinct / (t2 - t1),
                   level_sizes=res.level_sizes, level_generated=res.level_generated,
                   trace_actions=[a for a, _ in trace], golden_state_depths=golden,
                   golden_ok=(None if golden is None else all(0 < g <= i + 1 for i, g in enumerate(golden[: res.depth]))))
        print(json.dumps(out))
        with open(os.path.join(args.out, "hunt_%s.json" % tag), "w") as f:
            json.dump(out, f)
    eng.close()
    if group is not None:
        group.close()
    if world > 1:
        tdist.destroy_process_group()


if __name__ == "__main__":
    main()
