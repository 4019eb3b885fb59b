"""ctypes access to the CPU oracle (oracle/_build/liboracle.so).  Test infrastructure only."""
import ctypes as C
import os
import struct
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_SO = os.path.join(ROOT, "oracle", "_build", "liboracle.so")

_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(ORACLE_SO):
            import subprocess
            subprocess.check_call(["make", "-j8"], cwd=os.path.join(ROOT, "oracle"))
        L = C.CDLL(ORACLE_SO)
        vp = C.c_void_p
        L.orc_init_flat.argtypes = [vp, vp]
        L.orc_successors_flat.argtypes = [vp, vp, vp, vp, C.c_int]
        L.orc_digest_flat.argtypes = [vp, vp, C.c_int, vp, vp]
        L.orc_digest_full_flat.argtypes = [vp, vp, C.c_int, vp]
        L.orc_invariant_flat.argtypes = [vp, vp]
        L.orc_check_assumptions_flat.argtypes = [vp, vp]
        L.orc_check_assumptions_flat.restype = C.c_uint64
        L.orc_print_flat.argtypes = [vp, vp, C.c_int, C.c_char_p, C.c_int]
        L.orc_parse_trace.argtypes = [C.c_char_p, vp, vp, vp, C.c_int]
        L.orc_reprint_trace.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_int]
        L.orc_bfs.argtypes = [vp, C.c_int, C.c_int, C.c_uint64, C.c_double, C.c_int, C.c_int, C.c_int, C.c_char_p, vp, vp, vp,
                              C.c_int, vp, vp, C.c_int]
        _lib = L
    return _lib


def params(R, V, L, symmetry=True, view=True, invariant=1, C_=1, restart=0):
    """q = {R, C, V, L, restart_limit, symmetry, view, invariant(1..4)}"""
    return (C.c_int * 8)(R, C_, V, L, restart, int(symmetry), int(view), invariant)


class OracleBfs:
    def __init__(self, scal, levels, levgen, level_digests, trace):
        self.generated, self.distinct, self.queue, self.depth = int(scal[0]), int(scal[1]), int(scal[2]), int(scal[3])
        self.rc, self.complete, self.h2_ties = int(scal[4]), bool(scal[5]), int(scal[6])
        self.assumptions = [int(scal[i]) for i in range(7, 16)]
        self.trace_len = int(scal[16])
        self.seconds = int(scal[17]) / 1e6
        self.level_sizes = levels
        self.level_generated = levgen
        self.level_digests = level_digests  # list of sorted lists of 16-byte digests
        self.trace = trace


def bfs(q, workers=4, max_depth=0, max_states=0, max_seconds=0.0, check_deadlock=False, keep_trace=True, check_assumptions=True,
        digests=False, flat_cls=None):
    L = lib()
    scal = (C.c_uint64 * 32)()
    lv = (C.c_uint64 * 512)()
    lg = (C.c_uint64 * 512)()
    path = None
    if digests:
        fd, path = tempfile.mkstemp(suffix=".dig")
        os.close(fd)
    tcap = 256 if (flat_cls and keep_trace) else 0
    tr = (flat_cls * tcap)() if tcap else None
    ta = (C.c_int * max(tcap, 1))()
    L.orc_bfs(q, workers, max_depth, max_states, max_seconds, int(check_deadlock), int(keep_trace), int(check_assumptions),
              path.encode() if path else None, scal, lv, lg, 512, tr, ta, tcap)
    n = int(scal[18])
    level_digests = []
    if path:
        with open(path, "rb") as f:
            data = f.read()
        os.unlink(path)
        off = 0
        while off < len(data):
            (cnt,) = struct.unpack_from("<Q", data, off)
            off += 8
            level_digests.append([data[off + 16 * i: off + 16 * i + 16] for i in range(cnt)])
            off += 16 * cnt
    trace = []
    if tr is not None:
        trace = [(int(ta[i]), tr[i]) for i in range(min(int(scal[16]), tcap))]
    return OracleBfs(scal, [int(lv[i]) for i in range(n)], [int(lg[i]) for i in range(n)], level_digests, trace)


def digests_of(q, flats, n=None):
    """canonical VIEW digests (16 bytes each) + aux keys of (the first n of) an array of VsrFlatState"""
    n = len(flats) if n is None else n
    out = (C.c_uint64 * (2 * n))()
    aux = (C.c_uint32 * n)()
    lib().orc_digest_flat(q, flats, n, out, aux)
    raw = bytes(out)
    return [raw[16 * i:16 * i + 16] for i in range(n)], [int(aux[i]) for i in range(n)]


def digests_full_of(q, flats):
    n = len(flats)
    out = (C.c_uint64 * (2 * n))()
    lib().orc_digest_full_flat(q, flats, n, out)
    raw = bytes(out)
    return [raw[16 * i:16 * i + 16] for i in range(n)]


def print_flat(q, flat, with_rec_vars=True):
    buf = C.create_string_buffer(1 << 18)
    n = lib().orc_print_flat(q, C.byref(flat), int(with_rec_vars), buf, len(buf))
    assert n >= 0
    return buf.value.decode()
