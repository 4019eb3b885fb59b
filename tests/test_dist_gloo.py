"""N>1 control flow of the fingerprint-sharded BFS (vsr-tlaplus_b200/dist.py) on CPU: world_size 2, 4 and 8 over
gloo, with tests/host_engine.HostEngine standing in for the CUDA engine.  Checks ownership routing, the
counts+records all-to-all, termination, G-independence of the result and the cross-rank trace walk."""
import os
import sys

import pytest
import torch.distributed as tdist
import torch.multiprocessing as mp

import orc
from conftest import ROOT


def _worker(rank, world, port, cfg, q, kw):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    tdist.init_process_group("gloo", rank=rank, world_size=world)
    import _pkg
    pkg = _pkg.load()
    from vsr_tlaplus_b200 import dist as vdist
    from host_engine import HostEngine
    R, V, L, inv = cfg
    mc = pkg.ModelChecker.from_constants(R, V, L, invariants=inv)
    part = kw.pop("part_states", 0)
    kind = kw.pop("engine", "host")
    inbox = kw.pop("inbox_records", 1 << 17)
    group = None
    if kind == "gpu-p2p":
        # every rank on cuda:0, each in its own process: the kernel stores records into the peers' inboxes through CUDA IPC
        # mappings (on an NVLink box the same mappings go over NVLink); the level loop is the C++ vsr_bfs_sharded
        group = vdist.Group(kw.pop("group_name"), rank, world, timeout_s=120)
        eng = vdist.GpuEngine(mc, rank, world, device=0, table_capacity=1 << 21, frontier_capacity=1 << 19, inbox_records=inbox,
                              collect_levels=True, group=group)
        res = eng.run(part_states=part, **kw)
    elif kind == "gpu-staged":
        # the same kernels storing into a local staging buffer, records moved by torch.distributed (gloo, through host memory)
        eng = vdist.GpuEngine(mc, rank, world, device=0, table_capacity=1 << 21, frontier_capacity=1 << 19, inbox_records=inbox,
                              collect_levels=True, exchange="staged")
        res = vdist.ShardedBfs(eng, rank, world, part_states=part).run(**kw)
    else:
        eng = HostEngine(mc, rank, world)
        res = vdist.ShardedBfs(eng, rank, world, part_states=part).run(**kw)
    if hasattr(eng, "levels"):
        levels = [sorted(lv) for lv in eng.levels if lv or True]
    else:
        sb = mc.state_bytes
        raws = [eng.collected(d + 1) for d in range(res.depth)]
        levels = [sorted(raw[i:i + sb] for i in range(0, len(raw), sb)) for raw in raws]
    trace = vdist.replay_trace(mc, res.trace_cands) if (res.trace_cands or res.rc == 12) and rank == 0 else []
    q.put((rank, dict(rc=res.rc, generated=res.generated, distinct=res.distinct, depth=res.depth, complete=res.complete,
                      level_sizes=res.level_sizes, level_generated=res.level_generated, queue=res.queue,
                      violation_level=res.violation_level, sent=res.exchanged_records), levels, trace))
    eng.close()
    if group is not None:
        group.close()
    tdist.destroy_process_group()


def run_world(world, cfg, port, **kw):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, cfg, q, kw)) for r in range(world)]
    for p in procs:
        p.start()
    import queue as _queue
    import time as _time
    got = []
    deadline = _time.time() + 420  # a stuck rendezvous or collective must not hang the suite
    while len(got) < world:
        try:
            got.append(q.get(timeout=2))
        except _queue.Empty:
            dead = [p for p in procs if p.exitcode not in (None, 0)]
            if dead or _time.time() > deadline:
                for p in procs:
                    p.kill()
                raise AssertionError("a rank died (exit code %s)" % dead[0].exitcode if dead else "ranks still running after 420 s")
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    got.sort(key=lambda x: x[0])
    return got


@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_bfs_equals_oracle(world):
    cfg = (2, 2, 2, ("AcknowledgedWriteNotLost",))
    got = run_world(world, cfg, 29511 + world)
    o = orc.bfs(orc.params(2, 2, 2), workers=2)
    for rank, res, levels, _ in got:
        assert res["rc"] == 0 and res["complete"]
        assert (res["generated"], res["distinct"], res["depth"], res["queue"]) == (o.generated, o.distinct, o.depth, 0)
        assert res["level_sizes"] == o.level_sizes
        assert res["level_generated"] == o.level_generated
    # every rank agrees on the global numbers; the shards partition each level
    depth = got[0][1]["depth"]
    for d in range(depth):
        parts = [set(levels[d]) for _, _, levels, _ in got]
        union = set().union(*parts)
        assert sum(len(p) for p in parts) == len(union) == o.level_sizes[d]
    assert sum(r["sent"] for _, r, _, _ in got) > 0  # records really crossed ranks


def test_sharded_result_is_independent_of_world_size():
    cfg = (3, 1, 1, ("AcknowledgedWriteNotLost",))
    a = run_world(1, cfg, 29531, max_depth=12)
    b = run_world(2, cfg, 29532, max_depth=12)
    for d in range(a[0][1]["depth"]):
        sa = set(a[0][2][d])
        sb = set().union(*[set(lv[d]) for _, _, lv, _ in b])
        assert sa == sb
    assert a[0][1]["generated"] == b[0][1]["generated"] and a[0][1]["distinct"] == b[0][1]["distinct"]


def test_sharded_violation_and_cross_rank_trace():
    """a violation found on one rank stops all ranks at the same level; the counterexample is rebuilt by walking parent
    records across ranks and replays as a literal behaviour"""
    cfg = (3, 2, 1, ("AcknowledgedWritesExistOnMajority",))
    got = run_world(2, cfg, 29541)
    o = orc.bfs(orc.params(3, 2, 1, invariant=2), workers=8, keep_trace=False, check_assumptions=False)
    assert o.rc == 12
    for rank, res, _, _ in got:
        assert res["rc"] == 12 and res["violation_level"] == o.depth
    trace = got[0][3]
    assert len(trace) == o.depth and trace[0][0] == "Initial predicate"
    import _pkg
    pkg = _pkg.load()
    mc = pkg.ModelChecker.from_constants(3, 2, 1, symmetry=False, invariants=("AcknowledgedWritesExistOnMajority",))
    for i in range(len(trace) - 1):
        assert trace[i + 1][1] in [t for t, _, _ in mc.successors(trace[i][1])]
    assert mc.invariant(trace[-1][1]) != 0 and all(mc.invariant(s) == 0 for _, s in trace[:-1])


def test_sharded_bfs_in_sub_wavefronts():
    """wide levels pumped in several expand/exchange/insert sub-steps (bounded exchange buffers) give the same result"""
    cfg = (2, 2, 2, ("AcknowledgedWriteNotLost",))
    got = run_world(2, cfg, 29551, part_states=7)
    o = orc.bfs(orc.params(2, 2, 2), workers=2)
    for rank, res, levels, _ in got:
        assert res["rc"] == 0 and res["complete"]
        assert (res["generated"], res["distinct"], res["depth"]) == (o.generated, o.distinct, o.depth)
        assert res["level_sizes"] == o.level_sizes


def _oracle_prefix(o, got):
    """a bounded run stops after the level that reaches max_depth: the oracle reports one more (empty) expansion"""
    n = len(got)
    assert o.level_generated[:n] == got and all(x == 0 for x in o.level_generated[n:])


@pytest.mark.gpu
@pytest.mark.parametrize("kind,world,part,inbox", [("gpu-p2p", 2, 0, 1 << 17), ("gpu-p2p", 4, 0, 1 << 17), ("gpu-p2p", 8, 0, 1 << 16),
                                                   ("gpu-p2p", 2, 3000, 1 << 14), ("gpu-p2p", 8, 1000, 1 << 12),
                                                   ("gpu-staged", 2, 0, 1 << 17), ("gpu-staged", 4, 2000, 1 << 14)])
def test_ranks_sharing_one_gpu_match_oracle(kind, world, part, inbox):
    """The CUDA engine's world > 1 paths — push_records (destination-ordered staging + TMA bulk store into the owner's inbox),
    the drain of the inbox, ownership by fingerprint bits, steps with alternating inbox halves — on a box with ONE GPU:
    `world` processes all on cuda:0.  "gpu-p2p": inboxes mapped across processes with CUDA IPC, level loop in C++
    (vsr_bfs_sharded, what bench.py runs on NVLink); "gpu-staged": records moved by torch.distributed.  cfg2 to depth 13
    against the oracle: scalars, every level's size and successor count, and the shards partition every level."""
    import uuid
    cfg = (3, 2, 2, ("AcknowledgedWriteNotLost",))
    got = run_world(world, cfg, 29560 + world + (7 if part else 0) + (20 if kind == "gpu-staged" else 0), engine=kind, part_states=part,
                    inbox_records=inbox, max_depth=13, stop_on_violation=False, want_trace=False,
                    **({"group_name": "/vsr-test-" + uuid.uuid4().hex[:10]} if kind == "gpu-p2p" else {}))
    o = orc.bfs(orc.params(3, 2, 2), workers=8, max_depth=13, keep_trace=False)
    for rank, res, levels, _ in got:
        assert res["rc"] == 0
        assert (res["generated"], res["distinct"], res["depth"]) == (o.generated, o.distinct, o.depth)
        assert res["level_sizes"] == o.level_sizes
        _oracle_prefix(o, res["level_generated"])
    for d in range(got[0][1]["depth"]):
        parts = [set(levels[d]) for _, _, levels, _ in got]
        assert sum(len(p) for p in parts) == len(set().union(*parts)) == o.level_sizes[d]
    assert sum(r["sent"] for _, r, _, _ in got) > 0


@pytest.mark.gpu
def test_p2p_violation_trace_across_ranks():
    """the C++ level loop stops every rank at the violating level and walks the parent records across ranks; the candidate
    chain replays as a literal behaviour of Next that ends in the violation"""
    import uuid
    cfg = (3, 2, 1, ("AcknowledgedWritesExistOnMajority",))
    got = run_world(4, cfg, 29591, engine="gpu-p2p", group_name="/vsr-test-" + uuid.uuid4().hex[:10])
    o = orc.bfs(orc.params(3, 2, 1, invariant=2), workers=8, keep_trace=False, check_assumptions=False)
    assert o.rc == 12
    for rank, res, _, _ in got:
        assert res["rc"] == 12 and res["violation_level"] == o.depth
    trace = got[0][3]
    assert len(trace) == o.depth and trace[0][0] == "Initial predicate"
    import _pkg
    pkg = _pkg.load()
    mc = pkg.ModelChecker.from_constants(3, 2, 1, symmetry=False, invariants=("AcknowledgedWritesExistOnMajority",))
    for i in range(len(trace) - 1):
        assert trace[i + 1][1] in [t for t, _, _ in mc.successors(trace[i][1])]
    assert mc.invariant(trace[-1][1]) != 0 and all(mc.invariant(st) == 0 for _, st in trace[:-1])


@pytest.mark.gpu
def test_p2p_inbox_overflow_stops_every_rank_with_152():
    """an inbox too small for a step: the sender flags the overflow, nobody reads past a segment, and the level's all-gather
    stops all ranks with TLC's 'state space too large' status instead of hanging or dropping states silently"""
    import uuid
    cfg = (3, 2, 2, ("AcknowledgedWriteNotLost",))
    got = run_world(2, cfg, 29593, engine="gpu-p2p", inbox_records=64, part_states=100000, max_depth=14, stop_on_violation=False,
                    want_trace=False, group_name="/vsr-test-" + uuid.uuid4().hex[:10])
    for rank, res, _, _ in got:
        assert res["rc"] == 152
