"""Test double for dist.GpuEngine: the same stepwise interface, driven on the CPU through the product's
host-side single-state C ABI (vsr_successors / vsr_fingerprint / vsr_aux_key / vsr_invariant) so that the
N>1 control flow of dist.ShardedBfs (ownership by high fingerprint bits, counts + records all-to-all,
termination, violation min-reduce, cross-rank trace walk) can be exercised over gloo without a GPU.
Lives in tests/ on purpose: it is not a product path (the product BFS has no CPU fallback)."""
import struct
from types import SimpleNamespace

import torch

HDR = struct.Struct("<QQ")  # fp, parent gid << 12 | cand | mult << 56  (vsr_gpu.cuh RecHdr)
GID_SHIFT = 40
ROOT_PARENT = (1 << 44) - 1


class HostEngine:
    def __init__(self, mc, rank, world, send_capacity=1 << 16, check_deadlock=False):
        self.mc, self.rank, self.world = mc, rank, world
        self.dev = torch.device("cpu")
        self.sb = mc.state_bytes
        self.record_bytes = self.sb + HDR.size
        self.inbox_records = send_capacity
        self.check_deadlock = check_deadlock
        lg = world.bit_length() - 1
        self.shift = 64 - lg if world > 1 else 64
        self.reset()

    def owner(self, fp):
        # the product's rule (vsr_owner_rank = csrc/vsr_layout.h owner_of): high bits of the fingerprint times an odd constant
        return int(self.mc._lib.vsr_owner_rank(fp, self.world)) if self.world > 1 else self.rank

    def reset(self):
        self.seen = {}          # fp -> (level, auxkey)
        self.states = []        # local id -> packed bytes
        self.trace = []         # local id -> (parent gid, cand)
        self.frontier = []      # local ids of the current level
        self.next = []
        self.level = 0
        self.out = [[] for _ in range(self.world)]
        self.inbox = [[b"" for _ in range(self.world)] for _ in range(2)]
        self._li = self._blank()
        self.levels = []

    def _blank(self):
        return SimpleNamespace(new_states=0, generated=0, frontier_in=0, ties=0, collisions=0, violation=0, deadlock=0,
                               error_code=0, overflow=0, violation_id=0, deadlock_id=0, ms=0.0)

    def _insert(self, state, fp, aux, parent, cand, mult):
        li = self._li
        li.generated += mult
        hit = self.seen.get(fp)
        if hit is not None:
            if hit[0] == self.level + 1 and hit[1] != aux:
                li.ties += 1
            return
        self.seen[fp] = (self.level + 1, aux)
        lid = len(self.states)
        self.states.append(state)
        self.trace.append((parent, cand))
        self.next.append(lid)
        li.new_states += 1
        if self.mc.invariant(state) != 0 and not li.violation:
            li.violation, li.violation_id = 1, lid

    def seed(self):
        s = self.mc.init_state()
        fp = self.mc.fingerprint(s) or 1
        if self.owner(fp) == self.rank:
            self._insert(s, fp, self.mc.aux_key(s), ROOT_PARENT, 0, 1)

    def step(self, first, count, parity, drain_counts):
        """the engine's step: expand frontier[first : first + count] (own successors inserted, the others queued per
        destination), then insert what the peers sent in the previous step (the other half of the inbox)"""
        full = self.frontier
        self.frontier = full[first:first + count]
        self.out = [[] for _ in range(self.world)]
        try:
            self.expand()
        finally:
            self.frontier = full
        if drain_counts is not None:
            rb, sb = self.record_bytes, self.sb
            for src in range(self.world):
                raw = self.inbox[(parity ^ 1) & 1][src]
                for i in range(drain_counts[src]):
                    r = raw[i * rb:(i + 1) * rb]
                    fp, tm = HDR.unpack(r[sb:])
                    assert self.owner(fp) == self.rank
                    self._insert(r[:sb], fp, self.mc.aux_key(r[:sb]), (tm >> 12) & ROOT_PARENT, tm & 0xFFF, (tm >> 56) & 0xF)
        return [len(x) for x in self.out]

    def expand(self):
        import ctypes as C
        lib, h, sb = self.mc._lib, self.mc._h, self.sb
        ops_cap = 1024
        out = (C.c_uint8 * (sb * ops_cap))()
        acts = (C.c_uint8 * ops_cap)()
        mult = (C.c_uint32 * ops_cap)()
        self._li.frontier_in = len(self.frontier)
        for lid in self.frontier:
            st = self.states[lid]
            src = (C.c_uint8 * sb).from_buffer_copy(st)
            n = lib.vsr_successors(h, src, out, ops_cap, acts, mult)
            assert n >= 0
            cands = self._enabled_candidates(st)
            assert len(cands) == n
            if n == 0 and self.check_deadlock and not self._li.deadlock:
                self._li.deadlock, self._li.deadlock_id = 1, lid
            raw = bytes(out)
            for i in range(n):
                t = raw[i * sb:(i + 1) * sb]
                fp = self.mc.fingerprint(t) or 1
                gid = (self.rank << GID_SHIFT) | lid
                o = self.owner(fp)
                if o == self.rank:
                    self._insert(t, fp, self.mc.aux_key(t), gid, cands[i], int(mult[i]))
                else:
                    self.out[o].append(t + HDR.pack(fp, (gid << 12) | cands[i] | (int(mult[i]) << 56)))

    def _enabled_candidates(self, st):
        """true candidate indices (what trace records store) of the enabled bindings, in vsr_successors order"""
        import ctypes as C
        lib = self.mc._lib
        lib.vsr_enabled_candidates.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_uint32), C.c_size_t]
        buf = (C.c_uint32 * 1024)()
        n = lib.vsr_enabled_candidates(self.mc._h, (C.c_uint8 * self.sb).from_buffer_copy(st), buf, 1024)
        return [int(buf[i]) for i in range(n)]

    def outgoing(self, dest, n):
        data = b"".join(self.out[dest][:n])
        return torch.frombuffer(bytearray(data), dtype=torch.uint8) if data else torch.empty(0, dtype=torch.uint8)

    def put_incoming(self, parity, src, data, n):
        self.inbox[parity & 1][src] = data.reshape(-1).numpy().tobytes()[: n * self.record_bytes]

    def finish(self):
        li = self._li
        self.frontier, self.next = self.next, []
        self.levels.append([self.states[i] for i in self.frontier])
        self.level += 1
        self.out = [[] for _ in range(self.world)]
        self._li = self._blank()
        return li

    def frontier_size(self):
        return len(self.frontier)

    def trace_record(self, local_id):
        return self.trace[local_id]

    def sync(self):
        pass

    def close(self):
        pass
