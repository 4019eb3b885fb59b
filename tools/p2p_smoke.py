#!/usr/bin/env python
"""Smallest multi-rank run of the fused exchange without torch.distributed: `world` processes (all on cuda:DEVICE unless
--spread), each a rank of one vsr_group; the expand kernel stores remote successors into the peers' inboxes (CUDA IPC).

    python tools/p2p_smoke.py 2 [R V L] [--depth D] [--spread] [--inbox N] [--part N]
    compute-sanitizer --target-processes all python tools/p2p_smoke.py 2        # memcheck of push / drain
"""
import argparse
import json
import os
import subprocess
import sys
import uuid

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def rank_main(a):
    import _pkg
    pkg = _pkg.load()
    from vsr_tlaplus_b200 import dist as vdist
    rank, world = int(os.environ["P2P_RANK"]), a.world
    g = vdist.Group(os.environ["P2P_GROUP"], rank, world, timeout_s=120)
    mc = pkg.ModelChecker.from_constants(a.R, a.V, a.L)
    eng = vdist.GpuEngine(mc, rank, world, device=(rank if a.spread else a.device), table_capacity=a.table, frontier_capacity=a.frontier,
                          inbox_records=a.inbox, group=g)
    res = eng.run(max_depth=a.depth, stop_on_violation=False, want_trace=False, part_states=a.part)
    if rank == 0:
        print(json.dumps(dict(world=world, rc=res.rc, distinct=res.distinct, generated=res.generated, depth=res.depth, complete=res.complete,
                              kernel_ms=res.kernel_ms_max, seconds=res.seconds, sent_rank0=res.exchanged_records, level_sizes=res.level_sizes)))
    eng.close()
    g.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("world", type=int)
    ap.add_argument("R", type=int, nargs="?", default=2)
    ap.add_argument("V", type=int, nargs="?", default=2)
    ap.add_argument("L", type=int, nargs="?", default=2)
    ap.add_argument("--depth", type=int, default=0)
    ap.add_argument("--device", type=int, default=0)
    ap.add_argument("--spread", action="store_true", help="rank r on cuda:r")
    ap.add_argument("--table", type=int, default=1 << 20)
    ap.add_argument("--frontier", type=int, default=1 << 18)
    ap.add_argument("--inbox", type=int, default=1 << 14)
    ap.add_argument("--part", type=int, default=0)
    a = ap.parse_args()
    if "P2P_RANK" in os.environ:
        return rank_main(a)
    env = dict(os.environ, P2P_GROUP="/vsr-smoke-" + uuid.uuid4().hex[:10])
    procs = [subprocess.Popen([sys.executable] + sys.argv, env=dict(env, P2P_RANK=str(r))) for r in range(a.world)]
    rcs = [p.wait() for p in procs]
    sys.exit(max(abs(x) for x in rcs))


if __name__ == "__main__":
    main()
