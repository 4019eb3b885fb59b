"""Multi-GPU pump of the BFS wavefront: one process per GPU, the reachable set sharded by the
high bits of the 64-bit fingerprint, one all-to-all of newly generated packed states per wavefront.

torch.distributed is plumbing only: the counts exchange, the variable-size all-to-all of records (NCCL
grouped send/recv over NVLink/NVSwitch) and three tiny all-reduces per level.  Successor generation,
fingerprints, the seen-set and the invariant live in the CUDA engine behind the C ABI
(vsr_engine_expand / vsr_engine_insert_records); this file never looks inside a record.

``ShardedBfs`` is engine-agnostic on purpose: tests drive it over gloo with a host engine built from
the C ABI's single-state functions to exercise the N>1 control flow without a GPU.
"""
from __future__ import annotations

import ctypes as C
import time
from dataclasses import dataclass, field
from typing import List, Optional, Tuple

import torch
import torch.distributed as dist

from . import checker as ck

I64_MAX = (1 << 63) - 1
GID_SHIFT = 40  # global state id = rank << 40 | local id (vsr_gpu.cuh make_gid)


class GpuEngine:
    """The C-ABI stepwise engine of one rank plus the device buffers the exchange needs."""

    def __init__(self, mc: "ck.ModelChecker", rank: int, world: int, device: int = 0, table_capacity: int = 0,
                 frontier_capacity: int = 0, send_capacity: int = 1 << 20, keep_trace: bool = True,
                 check_deadlock: bool = False, collect_levels: bool = False):
        self.mc, self.rank, self.world = mc, rank, world
        self.lib = mc._lib
        self.dev = torch.device("cuda", device)
        o = mc.run_opts(deadlock=check_deadlock, device=device, table_capacity=table_capacity,
                        frontier_capacity=frontier_capacity, keep_trace=keep_trace, collect_levels=collect_levels)
        self._e = C.c_void_p()
        err = C.create_string_buffer(512)
        rc = self.lib.vsr_engine_create(mc._h, C.byref(o), rank, world, C.byref(self._e), err, len(err))
        if rc:
            raise ck.VsrError(rc, err.value.decode())
        self.record_bytes = int(self.lib.vsr_engine_record_bytes(self._e))
        self.send_capacity = send_capacity if world > 1 else 1
        self.send = torch.empty((world, self.send_capacity, self.record_bytes), dtype=torch.uint8, device=self.dev)
        self.send_count = torch.zeros(world, dtype=torch.int32, device=self.dev)
        self.lib.vsr_engine_set_send_buffers(self._e, self.send.data_ptr(), self.send_capacity, self.send_count.data_ptr())

    def _ck(self, rc):
        if rc:
            raise ck.VsrError(rc, self.lib.vsr_engine_last_error(self._e).decode())

    def reset(self):
        self._ck(self.lib.vsr_engine_reset(self._e))

    def seed(self):
        self._ck(self.lib.vsr_engine_seed_init(self._e))

    def expand(self):
        self._ck(self.lib.vsr_engine_expand(self._e))

    def expand_part(self, first: int, count: int):
        self._ck(self.lib.vsr_engine_expand_part(self._e, first, count))

    def send_counts(self) -> torch.Tensor:
        torch.cuda.current_stream(self.dev).synchronize()
        return self.send_count.to(torch.int64)

    def send_slice(self, dest: int, n: int) -> torch.Tensor:
        return self.send[dest, :n].reshape(-1)

    def new_recv(self, n: int) -> torch.Tensor:
        return torch.empty((max(n, 1), self.record_bytes), dtype=torch.uint8, device=self.dev)

    def insert(self, recs: torch.Tensor, n: int):
        if n:
            torch.cuda.current_stream(self.dev).synchronize()
            self._ck(self.lib.vsr_engine_insert_records(self._e, recs.data_ptr(), n))

    def finish(self) -> "ck.VsrLevelInfo":
        li = ck.VsrLevelInfo()
        self._ck(self.lib.vsr_engine_finish_level(self._e, C.byref(li)))
        return li

    def frontier_size(self) -> int:
        return int(self.lib.vsr_engine_frontier_size(self._e))

    def stats(self) -> "ck.VsrStats":
        st = ck.VsrStats()
        self.lib.vsr_engine_stats(self._e, C.byref(st))
        return st

    def trace_record(self, local_id: int) -> Tuple[int, int]:
        parent, cand = C.c_uint64(), C.c_uint32()
        self._ck(self.lib.vsr_engine_trace_record(self._e, local_id, C.byref(parent), C.byref(cand)))
        return int(parent.value), int(cand.value)

    def lookup(self, state: bytes) -> Tuple[int, int]:
        """(depth at which this canonical packed state was first seen on THIS rank's shard or 0, owner rank)"""
        lvl, owner = C.c_int(), C.c_int()
        buf = (C.c_uint8 * self.mc.state_bytes).from_buffer_copy(state)
        self._ck(self.lib.vsr_engine_lookup(self._e, buf, C.byref(lvl), C.byref(owner)))
        return int(lvl.value), int(owner.value)

    def collected(self, level: int) -> bytes:
        n = int(self.lib.vsr_engine_collected(self._e, level, None, 0))
        buf = (C.c_uint8 * max(n * self.mc.state_bytes, 1))()
        if n:
            self.lib.vsr_engine_collected(self._e, level, buf, n)
        return bytes(buf)[: n * self.mc.state_bytes]

    def close(self):
        if self._e:
            self.lib.vsr_engine_destroy(self._e)
            self._e = None

    def sync(self):
        torch.cuda.synchronize(self.dev)


@dataclass
class ShardedResult:
    rc: int = 0
    generated: int = 0
    distinct: int = 0
    queue: int = 0
    depth: int = 0
    complete: bool = False
    level_sizes: List[int] = field(default_factory=list)
    level_generated: List[int] = field(default_factory=list)
    h2_ties: int = 0
    fp_collisions: int = 0
    violation_level: int = 0
    violation_gid: int = -1
    seconds: float = 0.0
    kernel_ms_max: float = 0.0       # sum over levels of the slowest rank's kernel time
    insert_ms_max: float = 0.0       # of which insert_kernel (records from peers), slowest rank per level
    exchanged_records: int = 0       # records this rank sent
    # this rank's host wall clock by phase (every engine call returns after its kernel has finished): expand, exchange
    # (counts + records, until the last record has arrived), insert, finish (tie resolution, counters, the level's all-reduces)
    phase_seconds: dict = field(default_factory=lambda: {"expand": 0.0, "exchange": 0.0, "insert": 0.0, "finish": 0.0})
    launches: int = 0
    trace_cands: List[int] = field(default_factory=list)
    trace: List[Tuple[str, bytes]] = field(default_factory=list)


class ShardedBfs:
    """Level-synchronous BFS over `world` engines; every rank runs this same loop."""

    ROOT_PARENT = (1 << 52) - 1

    def __init__(self, engine, rank: int, world: int, group=None, part_states: int = 0):
        self.e, self.rank, self.world, self.group = engine, rank, world, group
        self.part_states = part_states  # frontier states per sub-wavefront and rank (0 = whole level at once)
        # NCCL moves device tensors; gloo (CPU tests, and ranks that share one GPU in a test) gets host tensors
        self._nccl = world > 1 and dist.get_backend(group) == "nccl"
        self._cdev = getattr(engine, "dev", torch.device("cpu")) if self._nccl else torch.device("cpu")
        self._phase = {"expand": 0.0, "exchange": 0.0, "insert": 0.0, "finish": 0.0}

    # -- collectives (no-ops when world == 1) ---------------------------------------------------
    def _allreduce(self, vals: List[int], op) -> List[int]:
        if self.world == 1:
            return list(vals)
        t = torch.tensor(vals, dtype=torch.int64, device=self._cdev)
        dist.all_reduce(t, op=op, group=self.group)
        return [int(x) for x in t.cpu().tolist()]

    def _reduce_level(self, sums: List[int], mins: List[int], maxs: List[int]):
        """the level's sums, minima and maxima over ranks in ONE collective (an all-gather of a short vector reduced on the
        host) instead of three all-reduces: each collective is a host-synchronous round trip, 47 levels deep at cfg2"""
        if self.world == 1:
            return list(sums), list(mins), list(maxs)
        v = torch.tensor(list(sums) + list(mins) + list(maxs), dtype=torch.int64, device=self._cdev)
        parts = [torch.empty_like(v) for _ in range(self.world)]
        dist.all_gather(parts, v, group=self.group)
        h = torch.stack(parts).cpu()
        ns, nm = len(sums), len(mins)
        return (h[:, :ns].sum(0).tolist(), h[:, ns:ns + nm].min(0).values.tolist(), h[:, ns + nm:].max(0).values.tolist())

    def _exchange(self) -> int:
        """counts all-to-all, then the records; returns the number of records this rank sent"""
        if self.world == 1:
            return 0
        tx = time.time()
        counts = self.e.send_counts().to(self._cdev)  # int64[world]
        if int(counts.max()) > self.e.send_capacity:
            raise ck.VsrError(152, f"send buffer overflow: {int(counts.max())} records for one destination, capacity "
                                   f"{self.e.send_capacity}")
        recv_counts = torch.empty_like(counts)
        dist.all_to_all_single(recv_counts, counts, group=self.group)
        sc = [int(x) for x in counts.cpu().tolist()]
        rcnt = [int(x) for x in recv_counts.cpu().tolist()]
        total = sum(rcnt)
        recv = self.e.new_recv(total)
        rb = self.e.record_bytes
        parts = [self.e.send_slice(p, sc[p]) for p in range(self.world)]
        out = recv.reshape(-1)[: total * rb]
        if self._nccl:
            # grouped ncclSend/ncclRecv straight out of the per-destination send buffers over NVLink: no staging copy;
            # pairs with nothing to move are skipped on both sides (both know the counts)
            ops, off = [], 0
            for p in range(self.world):
                if rcnt[p]:
                    ops.append(dist.P2POp(dist.irecv, out[off * rb:(off + rcnt[p]) * rb], p, self.group))
                off += rcnt[p]
                if sc[p]:
                    ops.append(dist.P2POp(dist.isend, parts[p], p, self.group))
            if ops:
                for w in dist.batch_isend_irecv(ops):
                    w.wait()
        else:
            # gloo (CPU tests) has no list all-to-all: one variable-size all_to_all_single over a concatenation
            inp = (torch.cat(parts) if sum(sc) else parts[0][:0]).to(self._cdev)
            hout = out if out.device == self._cdev else torch.empty(total * rb, dtype=torch.uint8)
            dist.all_to_all_single(hout, inp, output_split_sizes=[c * rb for c in rcnt], input_split_sizes=[c * rb for c in sc],
                                   group=self.group)
            if hout is not out:
                out.copy_(hout)
        if self._nccl:
            torch.cuda.current_stream(self._cdev).synchronize()  # insert() waits for the records anyway: wait here, so the clock splits
        ti = time.time()
        self.e.insert(recv, total)
        self._phase["exchange"] += ti - tx
        self._phase["insert"] += time.time() - ti
        return sum(sc)

    # -- the loop -----------------------------------------------------------------------------------
    def run(self, max_depth: int = 0, max_seconds: float = 0.0, max_states: int = 0, stop_on_violation: bool = True,
            want_trace: bool = True) -> ShardedResult:
        r = ShardedResult()
        SUM, MAX = dist.ReduceOp.SUM, dist.ReduceOp.MAX
        t0 = time.time()
        self.e.reset()
        self.e.seed()
        self._phase = r.phase_seconds
        level = 0
        bad_gid, result = -1, 0
        while True:
            tf = time.time()
            li = self.e.finish()
            level += 1
            (new, gen, ties, coll, viol, dead, err, ovf, fin), (vmin, dmin), (kms, ims) = self._reduce_level(
                [int(li.new_states), int(li.generated), int(li.ties), int(li.collisions), int(li.violation), int(li.deadlock),
                 1 if li.error_code else 0, 1 if li.overflow else 0, int(self.e.frontier_size())],
                [(self.rank << GID_SHIFT) | int(li.violation_id) if li.violation else I64_MAX,
                 (self.rank << GID_SHIFT) | int(li.deadlock_id) if li.deadlock else I64_MAX],
                [int(li.ms * 1e6), int(getattr(li, "ms_insert", 0.0) * 1e6)])
            r.kernel_ms_max += kms / 1e6
            r.insert_ms_max += ims / 1e6
            self._phase["finish"] += time.time() - tf
            r.generated += gen
            r.distinct += new
            r.h2_ties += ties
            r.fp_collisions += coll
            if level >= 2:
                r.level_generated.append(gen)
            if new:
                r.level_sizes.append(new)
            if err:
                result = 255
                break
            if ovf:
                result = 152
                break
            if viol and not r.violation_level:
                r.violation_level, r.violation_gid = level, vmin
                result, bad_gid = 12, vmin
                if stop_on_violation:
                    break
            if dead:
                result, bad_gid = 11, dmin
                break
            if fin == 0:
                r.complete = True
                break
            if max_depth and level >= max_depth:
                break
            if level >= 254:  # the seen-set tags entries with an 8-bit depth
                result = 152
                break
            if max_states and r.distinct >= max_states:
                break
            if max_seconds:
                (late,) = self._allreduce([1 if time.time() - t0 >= max_seconds else 0], MAX)
                if late:
                    break
            if self.world == 1 or not self.part_states:
                te = time.time()
                self.e.expand()
                self._phase["expand"] += time.time() - te
                r.exchanged_records += self._exchange()
            else:
                # wide level: pump it in sub-wavefronts so the exchange buffers stay bounded
                (nparts,) = self._allreduce([(self.e.frontier_size() + self.part_states - 1) // self.part_states], MAX)
                for k in range(max(nparts, 1)):
                    te = time.time()
                    self.e.expand_part(k * self.part_states, self.part_states)
                    self._phase["expand"] += time.time() - te
                    r.exchanged_records += self._exchange()
        r.rc = result
        r.depth = len(r.level_sizes)
        (r.queue,) = self._allreduce([0 if r.complete else self.e.frontier_size()], SUM)
        if bad_gid >= 0 and want_trace:
            r.trace_cands = self._walk_trace(bad_gid)
        r.seconds = time.time() - t0
        return r

    def _walk_trace(self, gid: int) -> List[int]:
        """follow (parent, candidate) records across ranks from a state back to Init"""
        cands: List[int] = []
        for _ in range(4096):
            owner = gid >> GID_SHIFT
            buf = torch.zeros(2, dtype=torch.int64, device=self._cdev)
            if owner == self.rank:
                parent, cand = self.e.trace_record(gid & ((1 << GID_SHIFT) - 1))
                buf[0], buf[1] = parent, cand
            if self.world > 1:
                dist.broadcast(buf, src=owner, group=self.group)
            parent, cand = int(buf[0]), int(buf[1])
            if parent == self.ROOT_PARENT:
                break
            cands.append(cand)
            gid = parent
        return cands[::-1]


def replay_trace(mc: "ck.ModelChecker", cands: List[int]) -> List[Tuple[str, bytes]]:
    """Literal behaviour (fixed value names) from the candidate chain of a counterexample."""
    n = len(cands)
    arr = (C.c_uint32 * max(n, 1))(*cands)
    cap = n + 1
    out = (C.c_uint8 * (cap * mc.state_bytes))()
    acts = (C.c_uint8 * cap)()
    m = mc._lib.vsr_replay_candidates(mc._h, arr, n, out, acts, cap)
    if m < 0:
        raise ck.VsrError(255, "trace replay failed")
    raw = bytes(out)
    sb = mc.state_bytes
    return [(ck.ACTION_NAMES[acts[i]], raw[i * sb:(i + 1) * sb]) for i in range(m)]
