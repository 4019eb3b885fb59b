"""The rank group of a multi-GPU job (csrc/vsr_group.cpp) on CPU: the shared-memory barrier and all-gather that
vsr_bfs_sharded uses between wavefront steps, across processes (POSIX shm, as under torchrun) and across threads of one
process (as under `vsrmc -gpus N`), plus the failure paths: an aborting rank releases the waiters, a missing rank times out."""
import ctypes as C
import multiprocessing as mp
import os
import struct
import sys
import threading
import time
import uuid

import pytest

from conftest import ROOT


def _proc(rank, world, name, rounds, q, mode):
    sys.path.insert(0, ROOT)
    import _pkg
    pkg = _pkg.load()
    from vsr_tlaplus_b200 import dist as vdist
    try:
        g = vdist.Group(name, rank, world, timeout_s=30)
    except Exception as ex:  # noqa: BLE001
        q.put((rank, "open failed: %r" % (ex,)))
        return
    ok = True
    try:
        if mode == "abort" and rank == 1:
            time.sleep(0.3)
            g.abort()
            q.put((rank, "aborted"))
            return
        if mode == "timeout":
            g.set_timeout(1.0)
            if rank == 1:
                time.sleep(3.0)  # never arrives in time
                q.put((rank, "late"))
                return
        for i in range(rounds):
            got = g.allgather(struct.pack("<IIQ", rank, i, (rank + 1) * 1000003 * (i + 1)))
            for r, raw in enumerate(got):
                rr, ii, v = struct.unpack("<IIQ", raw)
                ok = ok and (rr, ii, v) == (r, i, (r + 1) * 1000003 * (i + 1))
            if i % 7 == 0:
                g.barrier()
        q.put((rank, "ok" if ok else "mismatch"))
    except pkg.VsrError as ex:
        q.put((rank, "error %d" % ex.rc))
    finally:
        g.close()


def _run(world, rounds, mode="normal"):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    name = "/vsr-grp-test-" + uuid.uuid4().hex[:10]
    ps = [ctx.Process(target=_proc, args=(r, world, name, rounds, q, mode)) for r in range(world)]
    for p in ps:
        p.start()
    out = dict(q.get(timeout=120) for _ in range(world))
    for p in ps:
        p.join(timeout=30)
    assert not os.path.exists("/dev/shm" + name)  # rank 0 unlinks the name once everybody has attached
    return out


@pytest.mark.parametrize("world", [2, 4, 8])
def test_allgather_and_barrier_across_processes(world):
    assert _run(world, 300) == {r: "ok" for r in range(world)}


def test_abort_releases_waiting_ranks():
    out = _run(3, 1000000, mode="abort")
    assert out[1] == "aborted" and out[0] == "error 153" and out[2] == "error 153"


def test_missing_rank_times_out_instead_of_hanging():
    out = _run(2, 5, mode="timeout")
    assert out[0] == "error 153"


def test_threads_of_one_process(pkg):
    lib = pkg.load_library()
    world = 4
    handles = (C.c_void_p * world)()
    assert lib.vsr_group_open_local(world, handles) == 0
    errors = []

    def work(r):
        g = C.c_void_p(handles[r])
        for i in range(200):
            mine = struct.pack("<II", r, i)
            out = (C.c_uint8 * (8 * world))()
            if lib.vsr_group_allgather(g, mine, 8, out):
                errors.append((r, i, "rc"))
                return
            raw = bytes(out)
            for k in range(world):
                if struct.unpack("<II", raw[8 * k:8 * k + 8]) != (k, i):
                    errors.append((r, i, k))
        lib.vsr_group_close(g)

    ts = [threading.Thread(target=work, args=(r,)) for r in range(world)]
    for t in ts:
        t.start()
    for t in ts:
        t.join(timeout=60)
    assert not errors
