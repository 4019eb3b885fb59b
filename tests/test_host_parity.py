"""CPU parity (no GPU): the product's packed Next / canonical labels / fingerprint vs the oracle, and the oracle
against hand-derivable facts.  Sized to run in about a minute."""
import ctypes as C
import itertools
import json
import os
import subprocess

import pytest

import orc
from conftest import ROOT

DIFF = os.path.join(ROOT, "build", "diff_host")


def run_diff(*args):
    if not os.path.exists(DIFF):
        import __graft_entry__
        __graft_entry__.build()
    r = subprocess.run([DIFF] + [str(a) for a in args], capture_output=True, text=True, timeout=900)
    out = json.loads(r.stdout.strip().splitlines()[-1])
    assert r.returncode == 0, (out, r.stderr[-2000:])
    return out


@pytest.mark.parametrize("R,V,L,sym,n", [(2, 1, 1, 0, 10**6), (2, 2, 2, 1, 10**6), (2, 2, 2, 0, 10**6), (3, 1, 1, 0, 10**6)])
def test_packed_next_equals_oracle_on_complete_spaces(R, V, L, sym, n):
    out = run_diff(R, V, L, sym, n)
    assert out["complete"] == 1 and out["mismatches"] == 0 and out["assumption_violations"] == 0


@pytest.mark.parametrize("R,V,L,sym", [(3, 2, 2, 1), (3, 3, 3, 1), (5, 2, 2, 1), (3, 3, 3, 0), (4, 2, 2, 1), (3, 3, 2, 1), (5, 3, 2, 1),
                                        (4, 3, 2, 1), (3, 2, 3, 1), (2, 3, 2, 0), (5, 1, 1, 0)])
def test_packed_next_equals_oracle_bfs_prefix(R, V, L, sym):
    out = run_diff(R, V, L, sym, 8000)
    assert out["checked"] == 8000 and out["mismatches"] == 0


@pytest.mark.parametrize("R,V,L", [(2, 2, 1), (3, 1, 2), (3, 2, 1), (3, 3, 1), (4, 1, 1), (4, 2, 1), (5, 2, 1)])
def test_every_other_builtin_layout_against_the_oracle(R, V, L):
    """the built-in layouts (VSR_FOR_EACH_CONFIG) not named in the tests above: BFS prefix + random walks each"""
    out = run_diff(R, V, L, 1 if V > 1 else 0, 4000, 1, 50000, 11)
    assert out["checked"] == 4000 and out["mismatches"] == 0 and out["assumption_violations"] == 0


@pytest.mark.parametrize("R,V,L,sym,seed", [(3, 2, 2, 1, 1), (3, 3, 3, 1, 2), (5, 2, 2, 1, 3), (3, 3, 3, 0, 4), (3, 2, 2, 0, 5), (5, 3, 2, 1, 6),
                                             (4, 3, 2, 1, 7)])
def test_packed_next_equals_oracle_on_random_walks(R, V, L, sym, seed):
    """simulation-style deep coverage: state transfer, view-change completion and log truncation only happen 15+ steps in"""
    out = run_diff(R, V, L, sym, 8000, 1, 100000, seed)
    assert out["mismatches"] == 0 and out["max_walk_depth"] >= 30


def test_oracle_small_configs_ground_truth():
    """BASELINE configs[0] and friends, full BFS under the oracle: the numbers every other test leans on"""
    o = orc.bfs(orc.params(2, 1, 1, symmetry=False), workers=2)
    assert (o.generated, o.distinct, o.queue, o.depth, o.complete, o.rc) == (100, 76, 0, 14, True, 0)
    assert o.level_sizes == [1, 2, 3, 5, 8, 9, 9, 9, 9, 6, 5, 5, 4, 1]
    o = orc.bfs(orc.params(2, 2, 2), workers=4)
    assert (o.generated, o.distinct, o.depth, o.complete) == (2812, 2073, 27, True)
    o = orc.bfs(orc.params(3, 1, 1, symmetry=False), workers=8)
    assert (o.generated, o.distinct, o.depth, o.complete) == (118746, 43941, 24, True)
    assert o.h2_ties == 0 and sum(o.assumptions) == 0
    # level-1/2 sizes are hand-derivable: (R-1)+1 distinct successors of Init under symmetry
    for (R, V, L) in [(2, 1, 1), (3, 2, 2), (3, 3, 3), (5, 2, 2)]:
        o = orc.bfs(orc.params(R, V, L), workers=2, max_depth=2)
        assert o.level_sizes == [1, (R - 1) + 1] and o.level_generated[0] == (R - 1) + V


def test_deadlock_exists_and_is_reported_by_the_oracle():
    """VSR.tla has reachable terminal states, so TLC's default deadlock check would stop the run"""
    o = orc.bfs(orc.params(2, 1, 1, symmetry=False), workers=1, check_deadlock=True)
    assert o.rc == 11


def test_canonical_labelling_is_a_canonical_form(pkg):
    """for states reached by random exploration and EVERY permutation pi of Values:
    canon(pack(pi(s))) == canon(pack(s)) — the fast key-sorted labelling picks one representative per orbit"""
    mc = pkg.ModelChecker.from_constants(3, 3, 3, symmetry=True)
    raw = pkg.ModelChecker.from_constants(3, 3, 3, symmetry=False)
    import random
    rnd = random.Random(7)
    checked = 0
    for walk in range(60):
        s = raw.init_state()
        for step in range(45):
            succ = raw.successors(s)
            if not succ:
                break
            s = rnd.choice(succ)[0]
            f = raw.unpack(s)
            base = mc.pack(f)
            for perm in itertools.permutations([1, 2, 3]):
                g = permute_flat(pkg, f, perm)
                assert mc.pack(g) == base
                checked += 1
    assert checked > 5000


def permute_flat(pkg, f, perm):
    """apply a permutation of value ids to a VsrFlatState (test-side, independent of product and oracle code)"""
    Flat = pkg.checker.VsrFlatState
    g = Flat.from_buffer_copy(bytes(f))
    m = lambda x: perm[x - 1] if 1 <= x <= len(perm) else x

    def fix_msg(k):
        if k.has_entry:
            k.entry.operation = m(k.entry.operation)
        for i in range(k.log_n):
            k.log[i].operation = m(k.log[i].operation)

    for r in range(g.R):
        rep = g.rep[r]
        for i in range(rep.log_n):
            rep.log[i].operation = m(rep.log[i].operation)
        for i in range(rep.n_svc):
            fix_msg(rep.svc_recv[i])
        for i in range(rep.n_dvc):
            fix_msg(rep.dvc_recv[i])
    for i in range(g.n_msgs):
        fix_msg(g.msgs[i])
    acked = [0] * len(perm)
    for v in range(len(perm)):
        acked[perm[v] - 1] = f.acked[v]
    for v in range(len(perm)):
        g.acked[v] = acked[v]
    return g


def gf2_mulmod(a, b, poly_full, deg=64):
    r = 0
    while b:
        if b & 1:
            r ^= a
        b >>= 1
        a <<= 1
        if a >> deg:
            a ^= poly_full
    return r


def test_fp64_polynomial_is_irreducible_and_table_is_rabin(pkg):
    """FP64_POLY is TLC's Polys[0] recalled from memory (no TLC source here): check it IS an irreducible degree-64
    polynomial over GF(2) (Rabin's test) and that the byte table implements polynomial reduction by it."""
    poly = 0x911498AE0E66BAD6  # bit 63 = x^0 ... bit 0 = x^63, x^64 implicit
    full = 1 << 64
    for i in range(64):
        if (poly >> (63 - i)) & 1:
            full |= 1 << i
    # x^(2^64) == x (mod p)  and  gcd(x^(2^32) - x, p) == 1
    x = 2
    t = x
    powers = {}
    for k in range(1, 65):
        t = gf2_mulmod(t, t, full)
        powers[k] = t
    assert powers[64] == x

    def gf2_mod(a, b):
        db = b.bit_length()
        while a.bit_length() >= db:
            a ^= b << (a.bit_length() - db)
        return a

    def gf2_gcd(a, b):
        while b:
            a, b = b, gf2_mod(a, b)
        return a
    assert gf2_gcd(full, powers[32] ^ x) == 1
    # fingerprint linearity over GF(2) (a Rabin fingerprint is affine: fp(a)^fp(b)^fp(c) == fp(a^b^c))
    mc = pkg.ModelChecker.from_cfg_text(pkg.cfg_text(3, ["v1", "v2"], 2, view=False, symmetry=False))
    s0 = mc.init_state()
    succ = [t for t, _, _ in mc.successors(s0)]
    a, b, c = s0, succ[0], succ[1]
    x3 = bytes(p ^ q ^ r for p, q, r in zip(a, b, c))
    assert mc.fingerprint(a) ^ mc.fingerprint(b) ^ mc.fingerprint(c) == mc.fingerprint(x3)


def test_view_masks_aux_variables_out_of_the_fingerprint(pkg):
    """VIEW view (VSR.tla:149-150) drops aux_svc / aux_client_acked: states differing only there share a fingerprint"""
    mc = pkg.ModelChecker.from_constants(3, 2, 2, symmetry=True, view=True)
    nv = pkg.ModelChecker.from_constants(3, 2, 2, symmetry=True, view=False)
    s = mc.init_state()
    f = mc.unpack(s)
    f.aux_svc = 1
    t = mc.pack(f)
    assert t != s
    assert mc.fingerprint(t) == mc.fingerprint(s)
    assert nv.fingerprint(t) != nv.fingerprint(s)
    assert mc.aux_key(t) != mc.aux_key(s)


def test_sliced_fingerprint_equals_bytewise_definition(pkg):
    """vsr_fingerprint (slicing-by-8, what the GPU computes) == the byte-at-a-time FP64 definition, on real states,
    with and without VIEW, for layouts with an even and an odd number of VIEW words"""
    import random
    for (R, V, L, view) in [(3, 2, 2, True), (3, 2, 2, False), (3, 3, 3, True), (5, 2, 2, True), (2, 1, 1, True), (2, 2, 2, True)]:
        mc = pkg.ModelChecker.from_constants(R, V, L, view=view)
        rnd = random.Random(R * 100 + V * 10 + L)
        s = mc.init_state()
        for _ in range(300):
            buf = (C.c_uint8 * mc.state_bytes).from_buffer_copy(s)
            assert mc._lib.vsr_fingerprint(mc._h, buf) == mc._lib.vsr_fingerprint_bytewise(mc._h, buf)
            succ = mc.successors(s)
            if not succ:
                s = mc.init_state()
                continue
            s = rnd.choice(succ)[0]


@pytest.mark.parametrize("sym", [1, 0])
def test_every_action_is_compared_with_the_oracle(pkg, tmp_path, sym):
    """Uniform random walks from Init and shallow BFS prefixes never enable the state-transfer actions (SendGetState needs a
    Prepare from a higher view with a gap: 16+ steps in), and the golden trace never takes ReceiveHigherDVC, ReceiveGetState or
    ReceiveNewState.  Walks started from the 24 golden-trace states reach all of them: every one of the 15 actions
    that can fire with RestartEmptyLimit = 0 must have successors compared with the oracle's, with no mismatch."""
    import base64, zlib
    fx = json.load(open(os.path.join(ROOT, "tests", "golden", "state_transfer_trace.json")))
    Flat = pkg.checker.VsrFlatState
    mc = pkg.ModelChecker.from_constants(3, 3, 3, symmetry=bool(sym))
    seeds = tmp_path / "seeds.hex"
    seeds.write_text("".join(mc.pack(Flat.from_buffer_copy(zlib.decompress(base64.b64decode(s["flat_zlib_b64"])))).hex() + "\n"
                             for s in fx["states"]))
    out = run_diff(3, 3, 3, sym, 40000, 1, 1000000, 5, str(seeds))
    assert out["mismatches"] == 0 and out["checked"] == 40000
    cover = dict(zip(pkg.ACTION_NAMES[1:16], out["action_coverage"]))
    assert all(n > 0 for n in cover.values()), cover
    assert cover["SendGetState"] >= 50 and cover["ReceiveGetState"] >= 100 and cover["ReceiveNewState"] >= 10, cover


def test_owner_ranks_spread_every_senders_successors_evenly(pkg):
    """Sharding: FP64 is linear over GF(2), so with the fingerprint's own high bits as the owner a rank's successors
    would go to `owner(parent) ^ c` for a handful of constants c — on the shipped VSR.cfg with 8 ranks some (sender, owner) pairs
    carried 10x the records of others and overflowed their inbox segment (the first 8-GPU run).  vsr_owner_rank must give every
    pair about the same share, and a uniform split of the states themselves."""
    mc = pkg.ModelChecker.from_constants(3, 2, 2)
    lib = mc._lib
    seen, frontier = {mc.init_state()}, [mc.init_state()]
    for _ in range(10):  # depth 11: 44,840 states
        nxt = []
        for s in frontier:
            for t, _, _ in mc.successors(s):
                if t not in seen:
                    seen.add(t)
                    nxt.append(t)
        frontier = nxt
    for world in (2, 4, 8):
        mat = [[0] * world for _ in range(world)]
        plain = [[0] * world for _ in range(world)]
        shift = 64 - (world.bit_length() - 1)
        for s in frontier:
            f = mc.fingerprint(s)
            a = lib.vsr_owner_rank(f, world)
            for t, _, _ in mc.successors(s):
                g = mc.fingerprint(t)
                mat[a][lib.vsr_owner_rank(g, world)] += 1
                plain[f >> shift][g >> shift] += 1
        cells = [c for row in mat for c in row]
        assert max(cells) < 1.25 * min(cells), mat
        owners = [sum(row) for row in mat]
        assert max(owners) < 1.1 * min(owners)
        if world == 8:  # what the rule replaces, for the record: the plain high bits are far from even
            pc = [c for row in plain for c in row]
            assert max(pc) > 4 * min(pc)
    assert lib.vsr_owner_rank(12345, 1) == 0 and lib.vsr_owner_rank(12345, 3) == -1
