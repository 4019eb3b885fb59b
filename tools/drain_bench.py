#!/usr/bin/env python
"""One-GPU emulation of a rank of a 2-GPU job, to time the drain path without NVLink: a world-2 engine (rank 0) in staged
mode whose records for "rank 1" are copied back into its own inbox, so that half of all successors go through
push -> inbox -> drain and the whole state space is still explored on this GPU.  Prints the expand+drain and drain-only
kernel time per BFS: drain-only seconds / records drained = what a record costs the owner.

    python tools/drain_bench.py [R V L] [--table N] [--frontier N] [--inbox N]
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("R", type=int, nargs="?", default=3)
    ap.add_argument("V", type=int, nargs="?", default=2)
    ap.add_argument("L", type=int, nargs="?", default=2)
    ap.add_argument("--table", type=int, default=1 << 32)
    ap.add_argument("--frontier", type=int, default=140_000_000)
    ap.add_argument("--inbox", type=int, default=16_000_000)
    ap.add_argument("--depth", type=int, default=0)
    a = ap.parse_args()
    import torch
    import _pkg
    pkg = _pkg.load()
    from vsr_tlaplus_b200 import dist as vdist
    mc = pkg.ModelChecker.from_constants(a.R, a.V, a.L)
    eng = vdist.GpuEngine(mc, 0, 2, device=0, table_capacity=a.table, frontier_capacity=a.frontier, inbox_records=a.inbox, keep_trace=True,
                          exchange="staged")
    part = max(1024, a.inbox // 4)
    eng.reset()
    # Init may belong to "rank 1": seed by inserting it as a record
    eng.seed()
    li = eng.finish()
    if li.new_states == 0:
        import struct
        s0 = mc.init_state()
        rec = s0 + struct.pack("<QQ", mc.fingerprint(s0) or 1, (((1 << 44) - 1) << 12) | (1 << 56))
        eng.insert(torch.frombuffer(bytearray(rec), dtype=torch.uint8).cuda(), 1)
        li = eng.finish()
    t0 = time.time()
    distinct, level, ms_all, ms_drain, drained = int(li.new_states), 1, 0.0, 0.0, 0
    while True:
        n = eng.frontier_size()
        if n == 0 or (a.depth and level >= a.depth):
            break
        nparts = (n + part - 1) // part
        drain = None
        for k in range(nparts + 1):
            if k == nparts and not drain:
                break
            sent = eng.step(k * part, part if k < nparts else 0, k & 1, drain)
            if k == nparts:
                break
            cnt = sent[1]
            if cnt:
                eng.put_incoming(k & 1, 1, eng.outgoing(1, cnt), cnt)  # what NVLink would have done
            drain = [0, cnt] if cnt else None
            drained += cnt
        li = eng.finish()
        level += 1
        distinct += int(li.new_states)
        ms_all += li.ms
        ms_drain += li.ms_insert
        if li.overflow or li.error_code:
            print("overflow/error", li.overflow, li.error_code)
            break
    print(json.dumps(dict(cfg=[a.R, a.V, a.L], distinct=distinct, depth=level, kernel_s=ms_all / 1e3, drain_only_s=ms_drain / 1e3,
                          records_drained=drained, wall=time.time() - t0, part=part)))
    eng.close()


if __name__ == "__main__":
    main()
