/*
 * vsr_gpu.cuh — device side of the BFS wavefront (sm_100a).
 *
 * One launch of expand_kernel<L> = TLC's worker loop (SURVEY §3.1 / §8a stages E1-E9) over one BFS
 * level of packed VSR.tla states:
 *   E1  successor enumeration   Ops<L>::step over the candidate (action, binding) index space
 *   E2  SYMMETRY                kept incrementally by step (canonical value labels)
 *   E3  VIEW                    mask of the aux bits inside fp64_view
 *   E4  fingerprint             FP64 (Rabin, slicing-by-8 tables in shared memory) of the packed VIEW bytes + a 32-bit check hash
 *   E5  seen-set                open-addressed HBM table of 16-byte {fp, meta} entries, one 128-bit load per
 *                               probe (issued as soon as the fingerprint is known), insertion by one 128-bit CAS
 *                               (ATOMG.E.CAS.128); probing is bounded, a full table is an error, never a spin
 *   E6  queue                   next frontier staged per warp in shared memory, flushed 32 states at a time
 *                               by a TMA bulk store (cp.async.bulk.global.shared::cta, UBLKCP)
 *   E7  invariant               evaluated inline on every newly inserted state
 *   E8  trace                   (parent id, candidate) per new state
 *   E9  deadlock                states with no enabled candidate
 * Work shape: a block takes 32*WARPS frontier states, one per thread.  SCAN: every thread copies its state
 * into registers and evaluates all guards of Next on it with compile-time candidate indices (Ops::enabled_group): one bit
 * per (action, binding); the enabled (state, candidate) pairs are written to a block pool grouped by action (packed warp
 * prefix sums, one shared atomic per warp and action group, one barrier).  APPLY: warps take batches of 32 pairs of ONE
 * action and apply them one per lane — the successor is built in a rotated shared-memory row, read once into registers,
 * fingerprinted, probed, inserted — so the expensive part runs with full lanes and without divergence between actions.
 * See profiles/round1_expand_kernel.md for the measurements that led here.
 */
#ifndef VSR_GPU_CUH
#define VSR_GPU_CUH

#include <cuda_runtime.h>
#include <stdint.h>

#include "vsr_actions.h"

namespace vsr {

struct DevCounters {
    unsigned long long out_count;   /* states appended to the next frontier this level */
    unsigned long long generated;   /* successors generated (TLC's count: one per binding) */
    unsigned long long ties;        /* same-level VIEW ties with a different aux key */
    unsigned long long collisions;  /* fp equal, check hash different */
    unsigned long long probes;      /* table entries inspected */
    unsigned long long viol_id;     /* smallest global id of a violating new state (~0 = none) */
    unsigned long long dead_id;     /* smallest global id of an expanded state without successors */
    unsigned long long tie_count;   /* entries in the tie list */
    int error;                      /* first E_* raised */
    int overflow;                   /* next frontier / send buffer / tie list full */
    int viol_which;                 /* mask bit of the violated invariant */
    int _pad;
    /* per LAUNCH (one memset clears them): */
    unsigned long long work_next;   /* next round of frontier states to hand out */
    unsigned long long drain_next;  /* next chunk of inbox records to hand out */
    unsigned int send_count[8];     /* records pushed to each rank by this launch (MAX_WORLD) */
};

/* a same-level VIEW tie: header, then the candidate's L::NW packed words */
struct TieRec {
    uint64_t fp;
    uint64_t parent;
    uint32_t auxkey, cand, check, _pad;
};

/* record shipped to the owner rank of a successor: the state's words, then this 16-byte header.  The owner recomputes the
   check hash and the aux key from the words (a few dozen instructions); the 64-bit fingerprint travels because computing
   it is the expensive part and the sender needs it anyway to find the owner. */
struct RecHdr {
    uint64_t fp;
    uint64_t tm; /* trace record (parent global id << 12 | candidate, 56 bits) | mult << 56 */
};
constexpr int MAX_WORLD = 8;


struct ExpandParams {
    const uint32_t* in;          /* current frontier, n_in states of L::NW words */
    unsigned long long n_in;
    unsigned long long in_base;  /* local id of in[0] */
    /* frontier spill (BASELINE configs[3]): a frontier buffer may continue in pinned host memory once its part in HBM is full.
       States [0, in_split) of this launch's view are at `in`, the rest at in_hi (NULL / ~0: no spill); same for out. */
    const uint32_t* in_hi;
    unsigned long long in_split;
    uint32_t* out;               /* next frontier */
    uint32_t* out_hi;
    unsigned long long out_split;
    unsigned long long out_cap;  /* states the next frontier holds in all (HBM part + host part) */
    unsigned long long out_base; /* local id of out[0] */
    uint64_t* table;             /* capacity entries of {fp, meta} */
    unsigned long long table_cap; /* entries: any multiple of VSR_BUCKET (not only powers of two: memory-bound configs size the seen-set to what is left) */
    uint64_t* trace;             /* per local id: make_trec(parent global id, candidate); may be null */
    unsigned long long trace_cap;
    DevCounters* ctr;
    uint8_t* ties;               /* tie_cap entries of sizeof(TieRec) + L::BYTES */
    unsigned long long tie_cap;
    const uint64_t* fp_tab;      /* 8 x 256 FP64 slicing tables */
    RunCfg run;
    int level;                   /* depth of the states being GENERATED (Init = 1) */
    int check_deadlock;
    int rank, world, owner_shift;/* owner(fp) = owner_of(fp, owner_shift) (world a power of two; shift 64 when world = 1) */
    /* world > 1.  push[d] = where THIS rank's records for rank d go: its segment of rank d's inbox, in rank d's memory,
       mapped here over NVLink (CUDA IPC / peer access) — the expand kernel stores them there itself — or a local staging
       buffer when the host moves them with a collective.  Slots are taken from the local counters ctr->send_count[d]. */
    uint8_t* push[MAX_WORLD];
    unsigned long long push_cap; /* records per segment */
    int push_direct;             /* 1: every lane stores its own record with 16-byte stores (VSR_B200_PUSH=direct); 0: staged + TMA bulk store */
    /* records received from rank s in the previous step (the other half of the double-buffered inbox): inserted by this
       launch after its share of the frontier */
    const uint8_t* drain[MAX_WORLD];
    unsigned int drain_n[MAX_WORLD];
    unsigned long long drain_total;
};

struct InsertParams {
    const uint8_t* recs;
    unsigned long long n;
    ExpandParams e;              /* table / out / trace / counters as above */
};

/* ------------------------------------------------------------------ primitives */

__device__ __forceinline__ void cas128(uint64_t* p, uint64_t s0, uint64_t s1, uint64_t& o0, uint64_t& o1) {
    /* compare with {0,0} (empty slot), swap in {s0,s1}; returns the previous contents */
    asm volatile(
        "{\n\t.reg .b128 cmp, swp, old;\n\tmov.b128 cmp, {%2, %3};\n\tmov.b128 swp, {%4, %5};\n\t"
        "atom.global.relaxed.gpu.cas.b128 old, [%6], cmp, swp;\n\tmov.b128 {%0, %1}, old;\n\t}"
        : "=l"(o0), "=l"(o1)
        : "l"(0ull), "l"(0ull), "l"(s0), "l"(s1), "l"(p)
        : "memory");
}
__device__ __forceinline__ void ld128_cg(const uint64_t* p, uint64_t& a, uint64_t& b) {
    asm volatile("ld.global.cg.v2.u64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory");
}
__device__ __forceinline__ uint64_t mix64(uint64_t x) { /* splitmix64 finaliser */
    x ^= x >> 30; x *= 0xbf58476d1ce4e5b9ULL;
    x ^= x >> 27; x *= 0x94d049bb133111ebULL;
    x ^= x >> 31;
    return x;
}
template <class L, bool USE_VIEW, class W> VSR_HD uint32_t check_hash_t(const W& w) {
    /* second, independent 32-bit hash of the VIEW words (3 instructions per word + finaliser): lets the seen-set tell
       fp64 collisions apart instead of silently merging two states as a bare fingerprint set would */
    constexpr int full = L::VIEW_BITS >> 5, rem = L::VIEW_BITS & 31;
    constexpr int nw = USE_VIEW ? (full + (rem ? 1 : 0)) : L::NW;
    uint32_t h = 0x9747b28cu;
VSR_UNROLL
    for (int i = 0; i < nw; i++) {
        uint32_t k = rdw(w, i);
        if (USE_VIEW && i == full) k &= (1u << rem) - 1u;
        h ^= k;
        h = ((h << 13) | (h >> 19)) * 5u + 0xe6546b64u;
    }
    h ^= h >> 16; h *= 0x85ebca6bu; h ^= h >> 13; h *= 0xc2b2ae35u; h ^= h >> 16;
    return h;
}
template <class L, class W> VSR_HD uint32_t check_hash(const W& w, bool use_view) {
    return use_view ? check_hash_t<L, true>(w) : check_hash_t<L, false>(w);
}

enum { INS_NEW = 0, INS_DUP = 1, INS_TIE = 2, INS_FULL = 3 };

/* global state id = rank << 40 | local id (44 bits; all ones = "no parent": Init); trace record = global id of the parent
   << 12 | candidate index (56 bits; bit 63 is a transient "violates the invariant" mark inside the staging area, bits
   56..59 carry mult in records that travel between ranks) */
constexpr unsigned long long GID_MASK = (1ull << 44) - 1ull, ROOT_GID = GID_MASK;
__host__ __device__ __forceinline__ uint64_t make_gid(int rank, unsigned long long local_id) { return ((uint64_t)rank << 40) | local_id; }
__host__ __device__ __forceinline__ uint64_t make_trec(uint64_t parent_gid, uint32_t cand) { return (parent_gid << 12) | (cand & 0xFFFu); }

__device__ __forceinline__ uint64_t make_meta(int level, uint32_t auxkey, uint32_t check) {
    return ((uint64_t)(uint32_t)level << 56) | ((uint64_t)(auxkey & 0xFFFFFFu) << 32) | check;
}

/* lock-free insert-if-absent over 16-byte entries {fp, meta}.  A probe reads one BUCKET of VSR_BUCKET consecutive entries
   (1: one 128-bit load; 2: one 256-bit load = a whole 32-byte sector; 4: two 256-bit loads issued together) and walks to
   the next bucket only when every slot of this one holds another state: the kernel is bound by the LATENCY of dependent
   probes (profiles/round2_expand_kernel.md), so a wider first probe is paid for in bandwidth the kernel does not use.
   Slots of a bucket fill in order (no deletions), so a lookup may stop at the first empty slot.  The bucket's entries
   are loaded by the caller as early as the fingerprint is known so that the HBM round trip overlaps the rest of the
   successor's work. */
#ifndef VSR_BUCKET
#define VSR_BUCKET 2
#endif
struct Probe { uint64_t e[2 * VSR_BUCKET]; };
__device__ __forceinline__ void ld256_cg(const uint64_t* p, uint64_t& a, uint64_t& b, uint64_t& c, uint64_t& d) {
    asm volatile("ld.global.cg.v4.u64 {%0, %1, %2, %3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p) : "memory");
}
__device__ __forceinline__ unsigned long long table_home(unsigned long long cap, uint64_t fp) {
    /* bucket = floor(hash * nbuckets / 2^64): any capacity, no division.  The hash is the fingerprint times an odd constant
       (the owner rank is the fingerprint's HIGH bits, so they must not select the bucket on their own) */
    return __umul64hi(fp * 0x9E3779B97F4A7C15ULL, cap / VSR_BUCKET) * VSR_BUCKET;
}
__device__ __forceinline__ void probe_load(const uint64_t* table, unsigned long long h, Probe& p) {
#if VSR_BUCKET == 1
    ld128_cg(table + 2 * h, p.e[0], p.e[1]);
#elif VSR_BUCKET == 2
    ld256_cg(table + 2 * h, p.e[0], p.e[1], p.e[2], p.e[3]);
#elif VSR_BUCKET == 4
    ld256_cg(table + 2 * h, p.e[0], p.e[1], p.e[2], p.e[3]);
    ld256_cg(table + 2 * h + 4, p.e[4], p.e[5], p.e[6], p.e[7]);
#else
#error "VSR_BUCKET must be 1, 2 or 4"
#endif
}
__device__ __forceinline__ int table_insert_from(uint64_t* table, unsigned long long cap, unsigned long long h, Probe p, uint64_t fp, uint64_t meta,
                                                 unsigned& probes, unsigned& collisions) {
    for (unsigned tries = 0;; tries++) {
        if (tries > (1u << 16)) return INS_FULL; /* the table is (nearly) full: never spin forever, the host aborts with 152 */
        probes++;
VSR_UNROLL
        for (int j = 0; j < VSR_BUCKET; j++) {
            uint64_t e0 = p.e[2 * j], e1 = p.e[2 * j + 1];
            if (e0 == 0) {
                cas128(table + 2 * (h + j), fp, meta, e0, e1);
                if (e0 == 0 && e1 == 0) return INS_NEW;
            }
            for (int again = 0; e1 == 0 && again < 64; again++) ld128_cg(table + 2 * (h + j), e0, e1); /* half-visible entry: look again */
            if (e0 == fp) {
                if ((uint32_t)e1 == (uint32_t)meta) {
                    const bool same_level = (e1 >> 56) == (meta >> 56);
                    const bool same_aux = ((e1 >> 32) & 0xFFFFFF) == ((meta >> 32) & 0xFFFFFF);
                    return (same_level && !same_aux) ? INS_TIE : INS_DUP;
                }
                collisions++;
            }
        }
        h += VSR_BUCKET;
        if (h >= cap) h = 0;
        probe_load(table, h, p);
    }
}
#if defined(VSR_EXP_CASFIRST) && VSR_BUCKET == 2
/* experiment (tools/variants.sh casfirst): no probe load — the first access of the home bucket IS the compare-and-swap of its
   first slot.  A new state whose home slot is free is in after ONE round trip instead of two (load, then CAS), a duplicate
   sitting in the home slot is recognised from the CAS's return value; only a home slot held by another state costs the second
   access (the bucket's other slot).  Every access becomes an atomic. */
__device__ __forceinline__ int table_insert_casfirst(uint64_t* table, unsigned long long cap, unsigned long long h, uint64_t fp, uint64_t meta,
                                                     unsigned& probes, unsigned& collisions) {
    Probe p;
    cas128(table + 2 * h, fp, meta, p.e[0], p.e[1]);
    if (p.e[0] == 0 && p.e[1] == 0) { probes++; return INS_NEW; }
    if (p.e[0] == fp && (uint32_t)p.e[1] == (uint32_t)meta) {
        probes++;
        const bool same_level = (p.e[1] >> 56) == (meta >> 56);
        const bool same_aux = ((p.e[1] >> 32) & 0xFFFFFF) == ((meta >> 32) & 0xFFFFFF);
        return (same_level && !same_aux) ? INS_TIE : INS_DUP;
    }
    ld128_cg(table + 2 * (h + 1), p.e[2], p.e[3]);
    return table_insert_from(table, cap, h, p, fp, meta, probes, collisions);
}
#endif
__device__ __forceinline__ int table_insert(uint64_t* table, unsigned long long cap, uint64_t fp, uint64_t meta,
                                            unsigned& probes, unsigned& collisions) {
    const unsigned long long h = table_home(cap, fp);
    Probe p;
    probe_load(table, h, p);
    return table_insert_from(table, cap, h, p, fp, meta, probes, collisions);
}

/* TMA bulk store shared -> global of `bytes` (multiple of 16), issued by one lane; waits until the
   shared source may be overwritten */
__device__ __forceinline__ void bulk_store(void* gdst, const void* ssrc, uint32_t bytes) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    const uint32_t s = (uint32_t)__cvta_generic_to_shared(ssrc);
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(s), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}

template <int NW> __device__ __forceinline__ uint32_t* out_state(const ExpandParams& P, unsigned long long i) {
    return i < P.out_split ? P.out + i * NW : P.out_hi + (i - P.out_split) * NW;
}

/* ------------------------------------------------------------------ expand kernel */

constexpr int SCAP = 64;      /* staged new states per warp */
#ifdef VSR_QPS
constexpr int QPS = VSR_QPS;  /* tools/variants.sh "qps1": a pool so small that the overflow path (leftovers) runs all the time */
#else
constexpr int QPS = 10;       /* pool entries per parent state (pool = QPS * states per block round) */
#endif

template <class L> struct WarpStage {
    alignas(16) uint32_t stage[SCAP * L::NW + 32]; /* new states, packed back to back for the bulk store (+ room for the skew of the scratch rows, Expander::scratch) */
    unsigned long long tstage[SCAP];             /* their trace records */
    /* per-warp running state.  It lives here, not in the Expander object: the big per-action routines are real calls
       (one copy of each in the instruction cache), and an object whose address is passed to them would be kept in
       local memory — 1024 threads x a few hundred bytes does not fit the L1 left beside 220 KB of shared memory. */
    int sn;                                      /* states currently staged */
};
template <class L, int WARPS> struct BlockSmemT {
    static constexpr int NS = WARPS * 32;        /* parent states per block round: one per thread */
    static constexpr int NG = 13;                /* action groups (Ops<L>::NGRP) */
    uint64_t fp_tab[8 * 256];                    /* FP64 slicing-by-8 tables */
    uint32_t par[NS * (L::NW + 1)];              /* parents, row stride NW+1 (odd: bank-conflict-free column reads) */
    static constexpr int QCAP = QPS * NS;
    uint16_t pool[QCAP];                         /* enabled (state, candidate) pairs, grouped by action */
    int qcount[NG];                              /* pairs found per group (may exceed what the pool holds) */
    int take;
    unsigned long long round_first;
    WarpStage<L> w[WARPS];
};
/* Block shape.  ONE block of up to 32 warps per SM when its shared memory fits (227 KB), else two blocks of 16 / 12 / 8 warps
   (2 x <= 113 KB).  Measured on the shipped VSR.cfg (profiles/round2_expand_kernel.md section 5): one block of 32 warps — rounds of
   1024 parents — takes 13 % less kernel time than two blocks of 16 with the same 32 resident warps: the end-of-round tail (the
   barrier stall of section 4) halves, and all warps of the SM run the same phase, so they share the instruction cache lines of the
   scan.  The pool item keeps 16 bits: thread (10) | candidate offset in its group (6), which bounds a group at 63 candidates. */
template <class L> struct ExpandCfg {
    static constexpr size_t SMEM_ONE = 227 * 1024 - 512, SMEM_TWO = 113 * 1024;
    static constexpr int max_grp() { int m = 0; for (int g = 0; g < Ops<L>::NGRP; g++) m = Ops<L>::grp_size(g) > m ? Ops<L>::grp_size(g) : m; return m; }
    template <int W> static constexpr int pick_one() { /* most warps (even) of ONE block per SM; 0: not even 18 fit */
        if constexpr (W < 18) return 0;
        else if constexpr (sizeof(BlockSmemT<L, W>) <= SMEM_ONE) return W;
        else return pick_one<W - 2>();
    }
    static constexpr int W2 = sizeof(BlockSmemT<L, 16>) <= SMEM_TWO ? 16 : (sizeof(BlockSmemT<L, 12>) <= SMEM_TWO ? 12 : 8);
    static constexpr int W1 = max_grp() < 64 ? pick_one<32>() : 0;
#ifdef VSR_FORCE_WARPS
    static constexpr int WARPS = VSR_FORCE_WARPS; /* tuning experiments only */
#else
    static constexpr int WARPS = W1 >= 2 * W2 - 4 ? W1 : W2; /* one block unless it would cost more than 4 resident warps */
#endif
    static constexpr int BLOCKS = WARPS > 16 ? 1 : 2;
    typedef BlockSmemT<L, WARPS> Smem;
};

/*
 * Block-synchronous scan, action-pure apply.  A block takes NS = 32*WARPS frontier states (one per thread):
 *   scan   every thread copies its own state into registers and evaluates all of Next's guards on it with compile-time
 *          candidate indices: one bit per (action, binding).  The enabled (state, candidate) pairs are then laid out in
 *          a block pool grouped by action (packed warp prefix sums, one shared atomic per warp and group, one barrier).
 *   apply  after the second barrier, warps take batches of 32 pairs of ONE group and apply them one per lane: no
 *          divergence between actions inside a warp, full lanes except one partial batch per group.
 * History, all measured on the shipped VSR.cfg (profiles/round1_expand_kernel.md): (1) every warp walking guards and
 * effects on its own: 88 kB of SASS against the instruction cache, 55 % `no_instruction` stall samples; (2) guards in a
 * run-time loop over candidates with one ballot + barrier per group: 75 warp instructions per (32 states, candidate)
 * for slot decoding and shared-memory field reads; (3) this form: 16 per candidate, 46 % fewer instructions overall.
 */
/* MULTI: the instantiation for several GPUs (push_records in emit, drain after the rounds).  The one-GPU instantiation has none
   of that code: it costs the hot path registers (380 vs 140 bytes of spills in emit at the 64-register budget). */
template <class L, bool MULTI> struct Expander {
    typedef Ops<L> O_;
    typedef typename ExpandCfg<L>::Smem Smem;
    static constexpr int WARPS = ExpandCfg<L>::WARPS, NS = Smem::NS;
    const ExpandParams& P;
    Smem& B;
    WarpStage<L>& S;
    const int lane, warp, tid;
    const uint32_t* mine = nullptr;
    bool have = false;

    __device__ Expander(const ExpandParams& p, Smem& b) : P(p), B(b), S(b.w[threadIdx.x >> 5]), lane(threadIdx.x & 31), warp(threadIdx.x >> 5), tid(threadIdx.x) {}

    /* flush the first n staged states (n <= 32) to the next frontier: one atomicAdd for the block of
       ids, one TMA bulk store for the states, then move the remainder (< 32 states) down */
    static __device__ __noinline__ void flush(const ExpandParams& P, WarpStage<L>& S, int lane, int n) {
        const int sn = S.sn;
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd(&P.ctr->out_count, (unsigned long long)n);
        base = __shfl_sync(0xffffffffu, base, 0);
        if (lane < n) {
            const unsigned long long t = S.tstage[lane];
            if (t >> 63) { /* marked by the inline invariant check; now the id is known */
                atomicMin(&P.ctr->viol_id, P.out_base + base + lane);
                S.tstage[lane] = t & ~(1ull << 63);
            }
        }
        __syncwarp();
        if (base + n > P.out_cap) {
            if (lane == 0) atomicExch(&P.ctr->overflow, 1);
        } else {
            if (lane == 0) {
                if (base + n <= P.out_split) bulk_store(P.out + base * L::NW, S.stage, (uint32_t)(n * L::BYTES));
                else if (base >= P.out_split) bulk_store(P.out_hi + (base - P.out_split) * L::NW, S.stage, (uint32_t)(n * L::BYTES));
                else { /* the block of ids straddles the end of the HBM part */
                    const int n1 = (int)(P.out_split - base);
                    bulk_store(P.out + base * L::NW, S.stage, (uint32_t)(n1 * L::BYTES));
                    bulk_store(P.out_hi, S.stage + n1 * L::NW, (uint32_t)((n - n1) * L::BYTES));
                }
            }
            if (P.trace && lane < n && P.out_base + base + lane < P.trace_cap) P.trace[P.out_base + base + lane] = S.tstage[lane];
        }
        __syncwarp();
        const int rest = sn - n;
        for (int i0 = 0; i0 < rest * L::NW; i0 += 32) {
            const int i = i0 + lane;
            uint32_t v = 0;
            if (i < rest * L::NW) v = S.stage[n * L::NW + i];
            __syncwarp();
            if (i < rest * L::NW) S.stage[i] = v;
        }
        unsigned long long t = 0;
        if (lane < rest) t = S.tstage[n + lane];
        __syncwarp();
        if (lane < rest) S.tstage[lane] = t;
        if (lane == 0) S.sn = rest;
        __syncwarp();
    }

    /* fingerprint, route, insert, stage: the part of apply that does not depend on the action */
    /* this lane's scratch row for the successor it builds: staging rows 32..63 are free whenever a batch starts (fewer than
       32 states are staged then).  Plain rows, bank conflicts avoided by skewing the row STARTS (measured against rotating
       every access: 7.5 % less kernel time on the shipped VSR.cfg, profiles/round2_expand_kernel.md).  Rows of
       NW words collide every p = 32 / gcd(NW, 32) lanes; shifting lane l's row by l / p words puts the 32 lanes' word i in
       32 different banks, and an access is base + i: no per-access arithmetic. */
    typedef uint32_t* Row;
    static constexpr int gcd32(int a) { int g = 32; while (a % g) g >>= 1; return g; }
    static constexpr int SKEW_P = 32 / gcd32(L::NW);
    static __device__ __forceinline__ Row scratch(WarpStage<L>& S, int lane) { return &S.stage[(32 + lane) * L::NW + lane / SKEW_P]; }

    /* Records for peer ranks (world > 1), pushed by the kernel itself: the lanes of the batch whose successor belongs to
       another rank lay their records out in destination order in the free half of the warp's staging area (every lane
       has read its scratch row into registers by now), each destination's run takes its slots in that rank's inbox with
       ONE atomicAdd on a local counter, and the run leaves as ONE TMA bulk store (cp.async.bulk.global.shared::cta) to
       the peer's memory — over NVLink when push[] is a peer mapping.  Fire and forget: nothing waits for the remote
       write, the owner inserts the records in its next launch (drain).  The scratch half holds CAPREC records; a batch
       with more senders goes in two passes. */
    static __device__ __forceinline__ void push_records(const ExpandParams& P, WarpStage<L>& S, int lane, const RegRow<L::NW>& v, int send_to, uint64_t fp,
                                                        uint64_t tm) {
        const unsigned senders = __ballot_sync(0xffffffffu, send_to >= 0);
        if (!senders) return;
        constexpr int RB = L::BYTES + (int)sizeof(RecHdr), RW = RB / 4;
        constexpr int CAPREC = (32 * L::BYTES) / RB;
        static_assert(CAPREC >= 16, "two passes must cover a batch");
        int off = 0, cnt = 0, rnk = 0; /* start of my destination's run in destination order, its length, my place in it */
        unsigned mymask = 0;
        for (int d = 0; d < P.world; d++) {
            const unsigned m = __ballot_sync(0xffffffffu, send_to == d);
            if (send_to > d) off += __popc(m);
            if (send_to == d) { mymask = m; cnt = __popc(m); rnk = __popc(m & ((1u << lane) - 1u)); }
        }
        unsigned base = 0;
        if (send_to >= 0 && rnk == 0) base = atomicAdd(&P.ctr->send_count[send_to], (unsigned)cnt);
        base = __shfl_sync(0xffffffffu, base, mymask ? __ffs(mymask) - 1 : 0);
        const bool fits = send_to >= 0 && (unsigned long long)base + (unsigned)cnt <= P.push_cap;
        if (send_to >= 0 && !fits && rnk == 0) atomicExch(&P.ctr->overflow, 3);
        if (P.push_direct) { /* fallback / A-B: no staging, four half-sector stores per lane */
            if (fits) {
                uint4* d = reinterpret_cast<uint4*>(P.push[send_to] + (size_t)(base + (unsigned)rnk) * RB);
                VSR_UNROLL
                for (int q = 0; q < L::NW / 4; q++) d[q] = make_uint4(v.w[4 * q], v.w[4 * q + 1], v.w[4 * q + 2], v.w[4 * q + 3]);
                d[L::NW / 4] = make_uint4((uint32_t)fp, (uint32_t)(fp >> 32), (uint32_t)tm, (uint32_t)(tm >> 32));
            }
            return;
        }
        const int pos = off + rnk, total = __popc(senders);
        uint32_t* sbuf = &S.stage[32 * L::NW];
        for (int lo = 0; lo < total; lo += CAPREC) {
            const int hi = lo + CAPREC < total ? lo + CAPREC : total;
            if (send_to >= 0 && pos >= lo && pos < hi) {
                uint32_t* r = sbuf + (pos - lo) * RW;
                VSR_UNROLL
                for (int j = 0; j < L::NW; j++) r[j] = v.w[j];
                r[L::NW] = (uint32_t)fp; r[L::NW + 1] = (uint32_t)(fp >> 32);
                r[L::NW + 2] = (uint32_t)tm; r[L::NW + 3] = (uint32_t)(tm >> 32);
            }
            __syncwarp();
            if (fits) {
                const int st = off > lo ? off : lo, en = off + cnt < hi ? off + cnt : hi;
                if (pos == st && st < en)
                    bulk_store(P.push[send_to] + (size_t)(base + (unsigned)(st - off)) * RB, sbuf + (st - lo) * RW, (uint32_t)((en - st) * RB));
            }
            __syncwarp();
        }
    }

    /* the seen-set insert of one state per lane and everything after it — tie list, inline invariant, compaction of the
       survivors into the warp's staging area, flush — shared by the expansion (emit) and by the records received from
       peers (drain).  Returns this lane's counts for the run's statistics: successors generated (low half) | seen-set
       probes (high half); the caller keeps the running sums in registers (a warp reduction per batch cost 25 shuffles) */
    static __device__ __forceinline__ unsigned long long commit(const ExpandParams& P, WarpStage<L>& S, int lane, const RegRow<L::NW>& v, bool live, uint64_t fp,
                                                                uint32_t chk, uint32_t auxkey, unsigned long long home, const Probe& first, unsigned long long trec,
                                                                unsigned mult, bool check_inv = true) {
        unsigned gen = 0, probes = 0, coll = 0;
        int sn = S.sn;
        bool isnew = false;
        int bad = 0;
        if (live) {
            const uint64_t meta = make_meta(P.level, auxkey, chk);
            gen = mult;
#if defined(VSR_EXP_CASFIRST) && VSR_BUCKET == 2
            (void)first;
            const int r = table_insert_casfirst(P.table, P.table_cap, home, fp, meta, probes, coll);
#else
            const int r = table_insert_from(P.table, P.table_cap, home, first, fp, meta, probes, coll);
#endif
            isnew = r == INS_NEW;
            if (r == INS_FULL) atomicExch(&P.ctr->overflow, 4);
            if (isnew && check_inv) bad = O_::invariant(P.run, v);
            if (coll) atomicAdd(&P.ctr->collisions, (unsigned long long)coll); /* never seen so far */
            if (r == INS_TIE) {
                atomicAdd(&P.ctr->ties, 1ull);
                const unsigned long long t = atomicAdd(&P.ctr->tie_count, 1ull);
                if (t < P.tie_cap) {
                    TieRec rec;
                    rec.fp = fp; rec.parent = trec >> 12; rec.auxkey = auxkey; rec.cand = (uint32_t)(trec & 0xFFFu); rec.check = chk; rec._pad = 0;
                    uint8_t* dst = P.ties + t * (sizeof(TieRec) + L::BYTES);
                    *(TieRec*)dst = rec;
                    VSR_UNROLL
                    for (int j = 0; j < L::NW; j++) ((uint32_t*)(dst + sizeof(TieRec)))[j] = v.w[j];
                } else atomicExch(&P.ctr->overflow, 2);
            }
        }
        /* compaction of the survivors into the warp's staging area (one ballot, all lanes).  The survivors' final rows
           may overlap other lanes' scratch rows: every lane has read its row before anybody writes */
        const unsigned newmask = __ballot_sync(0xffffffffu, isnew);
        __syncwarp();
        if (isnew) {
            const int slot = sn + __popc(newmask & ((1u << lane) - 1u));
            VSR_UNROLL
            for (int j = 0; j < L::NW; j++) S.stage[slot * L::NW + j] = v.w[j];
            if (bad) {
                trec |= 1ull << 63;
                atomicOr(&P.ctr->viol_which, bad);
            }
            S.tstage[slot] = trec;
        }
        sn += __popc(newmask);
        __syncwarp();
        if (lane == 0) S.sn = sn;
        __syncwarp();
        while (sn >= 32) {
            flush(P, S, lane, 32);
            sn -= 32;
        }
        return (unsigned long long)gen | ((unsigned long long)probes << 32);
    }

    /* fingerprint and route one successor per lane: the part of apply that does not depend on the action */
    static __device__ __noinline__ unsigned long long emit(const ExpandParams& P, Smem& B, WarpStage<L>& S, int lane, const Row n, int mult,
                                                           int cand, int si, bool act, bool check_inv) {
        int send_to = -1;
        uint64_t fp = 0;
        uint32_t chk = 0, auxkey = 0;
        unsigned long long home = 0, trec = 0;
        Probe first = {};
        bool live = false;
        /* the successor's words, read once from the lane's scratch row: fingerprint, check hash, aux key, invariant and
           the copies to the staging area / a peer's inbox all work on registers (constant word indices) */
        RegRow<L::NW> v;
        if (act && mult > 0) {
            VSR_UNROLL
            for (int j = 0; j < L::NW; j++) v.w[j] = rdw(n, j);
        }
        if (act) {
            if (mult < 0) {
                atomicCAS(&P.ctr->error, 0, mult);
            } else if (mult > 0) {
                fp = fp64_view8<L>(B.fp_tab, v, P.run.use_view != 0);
                if (fp == 0) fp = 1;
                const int owner = MULTI ? owner_of(fp, P.owner_shift) : P.rank;
                trec = make_trec(make_gid(P.rank, P.in_base + B.round_first + si), (uint32_t)cand);
                if (owner == P.rank) {
                    /* start the seen-set probe now; the check hash, aux key and tags are computed under its latency */
                    live = true;
                    home = table_home(P.table_cap, fp);
#if !(defined(VSR_EXP_CASFIRST) && VSR_BUCKET == 2)
                    probe_load(P.table, home, first);
#endif
                    chk = check_hash<L>(v, P.run.use_view != 0);
                    auxkey = O_::aux_key(v);
                } else {
                    send_to = owner; /* counted ("states generated") where it is inserted */
                }
            }
        }
        if (MULTI) {
            __syncwarp(); /* every lane has read its scratch row: that half of the staging area may now carry outgoing records */
            push_records(P, S, lane, v, send_to, fp, trec | ((uint64_t)(unsigned)mult << 56));
        }
        return commit(P, S, lane, v, live, fp, chk, auxkey, home, first, trec, (unsigned)mult, check_inv);
    }

    /* ---- drain: one record received from a peer per lane (world > 1), after this block's share of the frontier.  The
       sender computed the fingerprint; check hash and aux key are recomputed from the words; then the same seen-set insert
       / invariant / staging as a local successor.  drain_begin issues the header and bucket loads, drain_end consumes them.
       Measured (tools/drain_bench.py, profiles/round2_multi_gpu.md): push + drain cost 53 ps per record, i.e. the drain
       runs close to the seen-set's random-access ceiling; what did NOT help: 2 or 4 records in flight per lane (+15 % /
       +60 %: register spills), and pipelining a chunk under every batch of the expansion (+19 %, and +5 % on ONE GPU,
       again through spills in the hot loop). */
    struct DrainPre {
        uint64_t fp, tm;
        unsigned long long home;
        Probe first;
        const uint4* rec;
        bool have;
    };
    unsigned long long dchunk = ~0ull; /* inbox chunk this warp has claimed (>= nchunks: none left) */
    __device__ __forceinline__ unsigned long long drain_chunks() const { return (P.drain_total + 31) / 32; }
    __device__ __forceinline__ unsigned long long drain_claim() {
        unsigned long long c = 0;
        if (lane == 0) c = atomicAdd(&P.ctr->drain_next, 1ull);
        return c; /* lane 0's value; broadcast by the caller when it is needed (the atomic's latency hides under other work) */
    }
    __device__ __forceinline__ DrainPre drain_begin(unsigned long long chunk) {
        constexpr int RB = L::BYTES + (int)sizeof(RecHdr);
        DrainPre d;
        unsigned long long i = chunk * 32 + lane;
        d.have = chunk < drain_chunks() && i < P.drain_total;
        d.fp = d.tm = 0;
        d.home = 0;
        d.first = Probe{};
        d.rec = nullptr;
        if (d.have) {
            int s = 0;
            while (s < P.world - 1 && i >= P.drain_n[s]) { i -= P.drain_n[s]; s++; }
            d.rec = reinterpret_cast<const uint4*>(P.drain[s] + i * RB);
            const uint4 h = __ldcs(d.rec + L::NW / 4);
            d.fp = ((uint64_t)h.y << 32) | h.x;
            d.tm = ((uint64_t)h.w << 32) | h.z;
            d.home = table_home(P.table_cap, d.fp);
#if !(defined(VSR_EXP_CASFIRST) && VSR_BUCKET == 2)
            probe_load(P.table, d.home, d.first);
#endif
        }
        return d;
    }
    __device__ __forceinline__ void drain_end(const DrainPre& d) {
        RegRow<L::NW> v;
        uint32_t chk = 0, auxkey = 0;
        if (d.have) {
            VSR_UNROLL
            for (int q = 0; q < L::NW / 4; q++) {
                const uint4 x = __ldcs(d.rec + q);
                v.w[4 * q] = x.x; v.w[4 * q + 1] = x.y; v.w[4 * q + 2] = x.z; v.w[4 * q + 3] = x.w;
            }
            chk = check_hash<L>(v, P.run.use_view != 0);
            auxkey = O_::aux_key(v);
        }
        tally(commit(P, S, lane, v, d.have, d.fp, chk, auxkey, d.home, d.first, d.tm & ((1ull << 56) - 1ull), (unsigned)((d.tm >> 56) & 0xFu)));
    }
    __device__ void drain() {
        const unsigned long long nchunks = drain_chunks();
        if (dchunk == ~0ull) dchunk = __shfl_sync(0xffffffffu, drain_claim(), 0);
        while (dchunk < nchunks) {
            const unsigned long long nx = drain_claim(); /* the next claim's latency hides under this chunk */
            const DrainPre d = drain_begin(dchunk);
            drain_end(d);
            dchunk = __shfl_sync(0xffffffffu, nx, 0);
        }
    }

    /* ---- scan: guards only, from registers.  Each thread copies its own state into registers and evaluates every guard
       of Next on it with compile-time candidate indices (Ops::enabled_group): a guard is a few bit tests on registers,
       not a decode of a run-time slot index plus shared-memory reads.  The result is one bit per candidate.  Then the
       (state, candidate) pairs are laid out in the block pool grouped by action: per-lane counts -> one packed warp
       prefix sum -> one shared atomic per (warp, group) -> barrier -> each lane writes its own pairs.  Two barriers per
       round instead of one per group. */
    static constexpr int NG = O_::NGRP;
    static __host__ __device__ constexpr int moff(int g) { int o = 0; for (int h = 0; h < g; h++) o += O_::grp_words(h); return o; }
    static constexpr int MW = moff(NG);       /* mask words per state */
    static constexpr int PW = (NG + 1) / 2;   /* packed 16-bit counters, two groups per word */
    static __host__ __device__ constexpr int max_grp() { int m = 0; for (int g = 0; g < NG; g++) m = O_::grp_size(g) > m ? O_::grp_size(g) : m; return m; }
    static constexpr int TBITS = NS <= 512 ? 9 : 10; /* pool item = thread (TBITS bits) | candidate offset in its group (the other 16 - TBITS) */
    static_assert(NS <= (1 << TBITS) && max_grp() < (1 << (16 - TBITS)), "pool item does not fit 16 bits");
    static_assert(max_grp() * NS < 65536, "16-bit packed counters");

    template <int G> __device__ __forceinline__ void guards(const RegRow<L::NW>& st, uint32_t* m, uint32_t* pc) {
        O_::template enabled_group<G>(P.run, st, m + moff(G));
        uint32_t c = 0;
        VSR_UNROLL
        for (int k = 0; k < O_::grp_words(G); k++) c += (uint32_t)__popc(m[moff(G) + k]);
        pc[G >> 1] += c << (16 * (G & 1));
        if constexpr (G + 1 < NG) guards<G + 1>(st, m, pc);
    }
    /* write this lane's pairs of group G (and the following groups) to the pool; pairs that do not fit stay in m */
    /* FAST: the whole round's pairs fit the pool (the caller has checked): no bound test per pair, nothing left over */
    template <int G, bool FAST> __device__ __forceinline__ void push(uint32_t* m, const uint32_t* ex, const uint32_t* wb, int st) {
        int pos = st + (int)((wb[G >> 1] >> (16 * (G & 1))) & 0xFFFFu) + (int)((ex[G >> 1] >> (16 * (G & 1))) & 0xFFFFu);
        VSR_UNROLL
        for (int k = 0; k < O_::grp_words(G); k++) {
            uint32_t mm = m[moff(G) + k], left = 0;
            while (mm) {
                const int bit = __ffs(mm) - 1;
                mm &= mm - 1;
                if (FAST || pos < Smem::QCAP) B.pool[pos] = (uint16_t)(tid | ((k * 32 + bit) << TBITS));
                else left |= 1u << bit;
                pos++;
            }
            if (!FAST) m[moff(G) + k] = left;
        }
        if constexpr (G + 1 < NG) {
            const int nst = FAST ? st + B.qcount[G] : (st + B.qcount[G] < Smem::QCAP ? st + B.qcount[G] : Smem::QCAP);
            push<G + 1, FAST>(m, ex, wb, nst);
        }
    }
    /* pool full (rare): the pairs left in m are applied right here by their own lanes, one group at a time */
    template <int G> __device__ __forceinline__ void leftovers(uint32_t* m) {
        VSR_UNROLL
        for (int k = 0; k < O_::grp_words(G); k++) {
            while (__any_sync(0xffffffffu, m[moff(G) + k] != 0)) {
                const bool inl = m[moff(G) + k] != 0;
                int cand = 0;
                if (inl) {
                    const int bit = __ffs(m[moff(G) + k]) - 1;
                    m[moff(G) + k] &= m[moff(G) + k] - 1;
                    cand = O_::grp_begin(G) + k * 32 + bit;
                }
                tally(apply<G>(P, B, S, lane, mine, cand, tid, inl));
            }
        }
        if constexpr (G + 1 < NG) leftovers<G + 1>(m);
    }
    __device__ __forceinline__ void scan_all() {
        uint32_t m[MW], pc[PW];
        VSR_UNROLL
        for (int i = 0; i < MW; i++) m[i] = 0;
        VSR_UNROLL
        for (int i = 0; i < PW; i++) pc[i] = 0;
        if (have) {
            RegRow<L::NW> st;
            VSR_UNROLL
            for (int i = 0; i < L::NW; i++) st.w[i] = mine[i];
            guards<0>(st, m, pc);
        }
        /* warp prefix sums of the per-lane counts, two groups per word */
        uint32_t ex[PW], wb[PW];
        VSR_UNROLL
        for (int i = 0; i < PW; i++) ex[i] = pc[i];
        VSR_UNROLL
        for (int o = 1; o < 32; o <<= 1) {
            VSR_UNROLL
            for (int i = 0; i < PW; i++) {
                const uint32_t t = __shfl_up_sync(0xffffffffu, ex[i], o);
                if (lane >= o) ex[i] += t;
            }
        }
        /* lane g reserves the warp's share of group g's pool segment */
        uint32_t mytot = 0, anyc = 0;
        VSR_UNROLL
        for (int g = 0; g < NG; g++) {
            const uint32_t t = (__shfl_sync(0xffffffffu, ex[g >> 1], 31) >> (16 * (g & 1))) & 0xFFFFu;
            if (lane == g) mytot = t;
        }
        VSR_UNROLL
        for (int i = 0; i < PW; i++) { anyc |= pc[i]; ex[i] -= pc[i]; } /* inclusive -> exclusive */
        uint32_t mybase = 0;
        if (lane < NG && mytot) mybase = (uint32_t)atomicAdd(&B.qcount[lane], (int)mytot);
        VSR_UNROLL
        for (int i = 0; i < PW; i++) wb[i] = 0;
        VSR_UNROLL
        for (int g = 0; g < NG; g++) wb[g >> 1] |= (__shfl_sync(0xffffffffu, mybase, g) & 0xFFFFu) << (16 * (g & 1));
        if (P.check_deadlock && have && !anyc) atomicMin(&P.ctr->dead_id, P.in_base + B.round_first + tid);
        __syncthreads(); /* qcount[] final: group g's segment starts at min(sum of the groups before it, QCAP) */
#ifndef VSR_EXP_NO_PUSHFAST
        int all = 0;
        VSR_UNROLL
        for (int g = 0; g < NG; g++) all += B.qcount[g];
        if (all <= Smem::QCAP) { /* block-uniform: the usual case */
            push<0, true>(m, ex, wb, 0);
            return;
        }
#endif
        push<0, false>(m, ex, wb, 0);
        uint32_t rest = 0;
        VSR_UNROLL
        for (int i = 0; i < MW; i++) rest |= m[i];
        if (__any_sync(0xffffffffu, rest != 0)) leftovers<0>(m);
    }
    /* apply one (parent, candidate) pair of group G per lane; the only copy of that action's effect in the kernel */
    template <int G> static __device__ __noinline__ unsigned long long apply(const ExpandParams& P, Smem& B, WarpStage<L>& S, int lane,
                                                                             const uint32_t* parent, int cand, int si, bool act) {
        const Row n = scratch(S, lane);
        int mult = 0;
        if (act) mult = O_::template step_grp<true, G>(P.run, parent, cand, n);
#ifdef VSR_EXP_INVSKIP
        /* the invariants read the replicas' logs and the acknowledgements only (VSR.tla:933-950): a successor of a state that
           satisfies them can violate them only through an action that rewrites a log or acknowledges a value (Ops::may_falsify) */
        return emit(P, B, S, lane, n, mult, cand, si, act, O_::may_falsify(G));
#else
        return emit(P, B, S, lane, n, mult, cand, si, act, true);
#endif
    }
    /* one batch of <= 32 queued pairs of group G, pool[b .. b + k) */
    template <int G> static __device__ __forceinline__ unsigned long long batch(const ExpandParams& P, Smem& B, WarpStage<L>& S, int lane, int b,
                                                                                int k) {
        const bool act = lane < k;
        int cand = 0, si = 0;
        if (act) {
            const unsigned item = B.pool[b + lane];
            si = item & ((1 << TBITS) - 1);
            cand = O_::grp_begin(G) + (int)(item >> TBITS);
        }
        return apply<G>(P, B, S, lane, &B.par[si * (L::NW + 1)], cand, si, act);
    }
    unsigned long long acc_gen = 0, acc_probes = 0; /* this lane's share of the statistics, summed over the launch */
    __device__ __forceinline__ void tally(unsigned long long r) {
        acc_gen += (unsigned)r;
        acc_probes += r >> 32;
    }

    __device__ void run_round(unsigned long long first, int count) {
        /* coalesced load of `count` parent states into padded rows */
        const uint32_t* src = P.in + first * L::NW;
        if (first + count <= P.in_split) {
            for (int i = tid; i < count * L::NW; i += NS) B.par[(i / L::NW) * (L::NW + 1) + (i % L::NW)] = __ldg(src + i);
        } else { /* (part of) this round's parents are in the host part of the frontier */
            for (int i = tid; i < count * L::NW; i += NS) {
                const unsigned long long st = first + i / L::NW;
                const uint32_t* row = st < P.in_split ? P.in + st * L::NW : P.in_hi + (st - P.in_split) * L::NW;
                B.par[(i / L::NW) * (L::NW + 1) + (i % L::NW)] = __ldg(row + i % L::NW);
            }
        }
        if (tid < Smem::NG) B.qcount[tid] = 0;
        if (tid == 0) { B.round_first = first; B.take = 0; }
        __syncthreads();
        have = tid < count;
        mine = &B.par[(have ? tid : 0) * (L::NW + 1)];
        scan_all();
        __syncthreads();
        /* batches: group g has ceil(|group g's pool segment| / 32) of them.  Lane g keeps group g's segment [st, en) and
           the index of its first batch, so mapping a batch number to (group, offset) is one ballot and three shuffles. */
        const int q = lane < Smem::NG ? B.qcount[lane] : 0;
        int inc = q;
        VSR_UNROLL
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        const int seg_st = inc - q < Smem::QCAP ? inc - q : Smem::QCAP, seg_en = inc < Smem::QCAP ? inc : Smem::QCAP;
        const int nb = (seg_en - seg_st + 31) >> 5;
        int binc = nb;
        VSR_UNROLL
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, binc, o);
            if (lane >= o) binc += t;
        }
        const int total = __shfl_sync(0xffffffffu, binc, 31);
        for (;;) {
            int t = 0;
            if (lane == 0) t = atomicAdd(&B.take, 1);
            t = __shfl_sync(0xffffffffu, t, 0);
            if (t >= total) break;
            const int g = __popc(__ballot_sync(0xffffffffu, lane < Smem::NG && binc <= t)); /* groups that end before batch t */
            const int st = __shfl_sync(0xffffffffu, seg_st, g), en = __shfl_sync(0xffffffffu, seg_en, g);
            const int b = st + (t - __shfl_sync(0xffffffffu, binc - nb, g)) * 32;
            const int k = en - b < 32 ? en - b : 32;
            unsigned long long r;
            switch (g) {
            case 0: r = batch<0>(P, B, S, lane, b, k); break;   case 1: r = batch<1>(P, B, S, lane, b, k); break;
            case 2: r = batch<2>(P, B, S, lane, b, k); break;   case 3: r = batch<3>(P, B, S, lane, b, k); break;
            case 4: r = batch<4>(P, B, S, lane, b, k); break;   case 5: r = batch<5>(P, B, S, lane, b, k); break;
            case 6: r = batch<6>(P, B, S, lane, b, k); break;   case 7: r = batch<7>(P, B, S, lane, b, k); break;
            case 8: r = batch<8>(P, B, S, lane, b, k); break;   case 9: r = batch<9>(P, B, S, lane, b, k); break;
            case 10: r = batch<10>(P, B, S, lane, b, k); break; case 11: r = batch<11>(P, B, S, lane, b, k); break;
            default: r = batch<12>(P, B, S, lane, b, k); break;
            }
            tally(r);
        }
    }

    __device__ void finish() {
        while (S.sn > 0) flush(P, S, lane, S.sn < 32 ? S.sn : 32);
        for (int o = 16; o; o >>= 1) {
            acc_gen += __shfl_xor_sync(0xffffffffu, acc_gen, o);
            acc_probes += __shfl_xor_sync(0xffffffffu, acc_probes, o);
        }
        if (lane == 0) {
            atomicAdd(&P.ctr->generated, acc_gen);
            atomicAdd(&P.ctr->probes, acc_probes);
        }
    }
};

template <class L, bool MULTI> __global__ void __launch_bounds__(ExpandCfg<L>::WARPS * 32, ExpandCfg<L>::BLOCKS) expand_kernel(const __grid_constant__ ExpandParams P) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    typedef typename ExpandCfg<L>::Smem Smem;
    Smem& B = *reinterpret_cast<Smem*>(smem_raw);
    for (int i = threadIdx.x; i < 8 * 256; i += blockDim.x) B.fp_tab[i] = P.fp_tab[i];
    __shared__ unsigned long long next_round;
    Expander<L, MULTI> X(P, B);
    if ((threadIdx.x & 31) == 0) {
        WarpStage<L>& S = B.w[threadIdx.x >> 5];
        S.sn = 0;
    }
    const unsigned long long nrounds = (P.n_in + Smem::NS - 1) / Smem::NS;
    if (threadIdx.x == 0) next_round = atomicAdd(&P.ctr->work_next, 1ull);
    for (;;) {
        __syncthreads();
        const unsigned long long c = next_round;
        __syncthreads();
        if (c >= nrounds) break;
        /* claim the round after this one now: the global atomic's latency hides under this round's work */
        if (threadIdx.x == 0) {
            next_round = atomicAdd(&P.ctr->work_next, 1ull);
            /* pull the next round's parents into L2 while this round runs */
            const unsigned long long nx = next_round;
            if (nx < nrounds && (nx + 1) * Smem::NS <= P.in_split) {
                const unsigned long long nfirst = nx * Smem::NS;
                const unsigned long long ncount = (P.n_in - nfirst) < (unsigned long long)Smem::NS ? (P.n_in - nfirst) : Smem::NS;
                asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(P.in + nfirst * L::NW), "r"((uint32_t)(ncount * L::BYTES)) : "memory");
            }
        }
        const unsigned long long first = c * Smem::NS;
        const int count = (int)((P.n_in - first) < (unsigned long long)Smem::NS ? (P.n_in - first) : Smem::NS);
        X.run_round(first, count);
    }
    if (MULTI && P.drain_total) X.drain();
    X.finish();
}

/* ------------------------------------------------------------------ insert kernel (records from peers, and Init) */

template <class L> __global__ void __launch_bounds__(256) insert_kernel(const InsertParams Q) {
    const ExpandParams& P = Q.e;
    const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const bool have = i < Q.n;
    bool isnew = false;
    const uint8_t* rec = Q.recs + (have ? i : 0) * (L::BYTES + sizeof(RecHdr));
    const uint32_t* n = (const uint32_t*)rec;
    const RecHdr* h = (const RecHdr*)(rec + L::BYTES);
    unsigned probes = 0, coll = 0, nties = 0;
    unsigned long long gen = 0;
    uint64_t parent = 0;
    uint32_t cand = 0;
    if (have) {
        /* check hash and aux key are computed here from the words; the header carries the fingerprint, the trace record and mult */
        const uint64_t meta = make_meta(P.level, Ops<L>::aux_key(n), check_hash<L>(n, P.run.use_view != 0));
        const int r = table_insert(P.table, P.table_cap, h->fp, meta, probes, coll);
        isnew = r == INS_NEW;
        if (r == INS_FULL) atomicExch(&P.ctr->overflow, 4);
        gen = (h->tm >> 56) & 0xFu;
        parent = (h->tm >> 12) & ((1ull << 44) - 1ull);
        cand = (uint32_t)(h->tm & 0xFFFu);
        if (r == INS_TIE) {
            nties = 1;
            const unsigned long long t = atomicAdd(&P.ctr->tie_count, 1ull);
            if (t < P.tie_cap) {
                TieRec tr;
                tr.fp = h->fp; tr.parent = parent; tr.auxkey = (uint32_t)((meta >> 32) & 0xFFFFFF); tr.cand = cand;
                tr.check = (uint32_t)meta; tr._pad = 0;
                uint8_t* dst = P.ties + t * (sizeof(TieRec) + L::BYTES);
                *(TieRec*)dst = tr;
                for (int j = 0; j < L::NW; j++) ((uint32_t*)(dst + sizeof(TieRec)))[j] = n[j];
            } else atomicExch(&P.ctr->overflow, 2);
        }
    }
    /* per-warp totals: one atomic per counter per warp, not per record */
    for (int o = 16; o; o >>= 1) {
        gen += __shfl_xor_sync(0xffffffffu, gen, o);
        probes += __shfl_xor_sync(0xffffffffu, probes, o);
        coll += __shfl_xor_sync(0xffffffffu, coll, o);
        nties += __shfl_xor_sync(0xffffffffu, nties, o);
    }
    if (lane == 0) {
        if (gen) atomicAdd(&P.ctr->generated, gen);
        if (probes) atomicAdd(&P.ctr->probes, (unsigned long long)probes);
        if (coll) atomicAdd(&P.ctr->collisions, (unsigned long long)coll);
        if (nties) atomicAdd(&P.ctr->ties, (unsigned long long)nties);
    }
    const unsigned newmask = __ballot_sync(0xffffffffu, isnew);
    if (newmask) {
        unsigned long long base = 0;
        const int leader = __ffs(newmask) - 1;
        if (lane == leader) base = atomicAdd(&P.ctr->out_count, (unsigned long long)__popc(newmask));
        base = __shfl_sync(0xffffffffu, base, leader);
        if (isnew) {
            const unsigned long long pos = base + __popc(newmask & ((1u << lane) - 1u));
            if (pos < P.out_cap) {
                uint32_t* dst = out_state<L::NW>(P, pos);
                for (int j = 0; j < L::NW; j++) dst[j] = n[j];
                if (P.trace && P.out_base + pos < P.trace_cap) P.trace[P.out_base + pos] = make_trec(parent, cand);
                const int bad = Ops<L>::invariant(P.run, n);
                if (bad) {
                    atomicMin(&P.ctr->viol_id, P.out_base + pos);
                    atomicOr(&P.ctr->viol_which, bad);
                }
            } else atomicExch(&P.ctr->overflow, 1);
        }
    }
}

/* VIEW-tie patch pass (only launched for a level that reported ties).  `ties` holds, sorted by fp, ONE
   record per tied fingerprint: the smallest (aux_key, parent, candidate) among the late arrivals.  Every state of the
   new level looks itself up; if a tie record beats the first arrival's aux_key, the state and its trace record are
   replaced, so the survivor is "smallest aux_key wins" whatever the arrival order — the rule the oracle applies.
   The invariant is re-evaluated on every state of the level (the verdict may change with the aux variables). */
template <class L> __global__ void patch_ties_kernel(const ExpandParams P, const uint8_t* ties, unsigned long long ntie, unsigned long long n_out) {
    const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_out) return;
    uint32_t w[L::NW];
    uint32_t* st = out_state<L::NW>(P, i);
    for (int j = 0; j < L::NW; j++) w[j] = st[j];
    uint64_t fp = fp64_view8<L>(P.fp_tab, w, P.run.use_view != 0);
    if (fp == 0) fp = 1;
    const uint32_t chk = check_hash<L>(w, P.run.use_view != 0);
    const size_t stride = sizeof(TieRec) + L::BYTES;
    unsigned long long lo = 0, hi = ntie;
    while (lo < hi) {
        const unsigned long long mid = (lo + hi) / 2;
        if (((const TieRec*)(ties + mid * stride))->fp < fp) lo = mid + 1; else hi = mid;
    }
    for (; lo < ntie; lo++) {
        const TieRec* t = (const TieRec*)(ties + lo * stride);
        if (t->fp != fp) break;
        if (t->check != chk) continue;
        if (t->auxkey < Ops<L>::aux_key(w)) {
            const uint32_t* tw = (const uint32_t*)((const uint8_t*)t + sizeof(TieRec));
            for (int j = 0; j < L::NW; j++) { w[j] = tw[j]; st[j] = tw[j]; }
            if (P.trace && P.out_base + i < P.trace_cap) P.trace[P.out_base + i] = make_trec(t->parent, t->cand);
        }
    }
    const int bad = Ops<L>::invariant(P.run, w);
    if (bad) {
        atomicMin(&P.ctr->viol_id, P.out_base + i);
        atomicOr(&P.ctr->viol_which, bad);
    }
}

/* membership query (tests / golden-trace cross-check): meta of the entry holding (fp, check), 0 if absent */
static __global__ void lookup_kernel(const uint64_t* table, unsigned long long cap, uint64_t fp, uint32_t check, unsigned long long* meta_out) {
    unsigned long long h = table_home(cap, fp);
    for (unsigned long long i = 0; i < cap; i++) {
        const uint64_t e0 = table[2 * h], e1 = table[2 * h + 1];
        if (e0 == 0) { *meta_out = 0; return; }
        if (e0 == fp && (uint32_t)e1 == check) { *meta_out = e1; return; }
        if (++h >= cap) h = 0;
    }
    *meta_out = 0;
}

/* ------------------------------------------------------------------ simulation mode (TLC `-simulate`)
   One thread per random walk from Init, `depth` states long at most; the invariant is checked on every state reached.
   The first violating (walk, depth) is kept (smallest walk index wins) and re-walked on the host for the trace. */
struct SimParams {
    unsigned long long num_walks, seed;
    int depth;
    RunCfg run;
    unsigned long long* first_bad; /* walk << 16 | depth of the state that violates (~0 = none) */
    unsigned long long* steps;     /* transitions taken */
    unsigned long long* dead_ends; /* walks that stopped in a state without successors */
    unsigned long long* probe_out; /* optional: for walks 0 .. probe_walks-1, fingerprint of the last state and transitions taken */
    unsigned long long probe_walks;
    const uint64_t* fp_tab;
};
template <class L> __global__ void simulate_kernel(const SimParams Q) {
    unsigned long long steps = 0, dead = 0;
    for (unsigned long long wk = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; wk < Q.num_walks; wk += (unsigned long long)gridDim.x * blockDim.x) {
        uint64_t rng = Q.seed ^ (wk * 0xD1B54A32D192ED03ULL);
        uint32_t a[L::NW], b[L::NW];
        Ops<L>::init((uint32_t*)a);
        unsigned long long mysteps = 0;
        for (int d = 2; d <= Q.depth; d++) {
            const int cand = Ops<L>::random_enabled(Q.run, (const uint32_t*)a, rng);
            if (cand < 0) { dead++; break; }
            if (Ops<L>::template step<true>(Q.run, (const uint32_t*)a, cand, (uint32_t*)b) <= 0) break;
            for (int j = 0; j < L::NW; j++) a[j] = b[j];
            steps++;
            mysteps++;
            if (Ops<L>::invariant(Q.run, (const uint32_t*)a)) {
                atomicMin(Q.first_bad, (wk << 16) | (unsigned long long)d);
                break;
            }
        }
        if (wk < Q.probe_walks) {
            Q.probe_out[2 * wk] = fp64_view<L>(Q.fp_tab, (const uint32_t*)a, false);
            Q.probe_out[2 * wk + 1] = mysteps;
        }
    }
    for (int o = 16; o; o >>= 1) {
        steps += __shfl_xor_sync(0xffffffffu, steps, o);
        dead += __shfl_xor_sync(0xffffffffu, dead, o);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(Q.steps, steps);
        atomicAdd(Q.dead_ends, dead);
    }
}

/* seen-set micro-benchmark: n splitmix64 keys, a fraction of them duplicates, inserted with the same
   table_insert the BFS uses; nothing else in the loop, so its rate is the random-probe ceiling of this table design */
static __global__ void probe_bench_kernel(uint64_t* table, unsigned long long cap, unsigned long long n, unsigned long long distinct,
                                   unsigned long long seed, unsigned long long* new_count, unsigned long long* probe_count) {
    unsigned long long mine_new = 0;
    unsigned probes = 0, coll = 0;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (unsigned long long)gridDim.x * blockDim.x) {
        unsigned long long z = seed + (i % distinct) * 0x9E3779B97F4A7C15ULL; /* splitmix64 of the key index */
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
        z ^= z >> 31;
        if (z == 0) z = 1;
        mine_new += table_insert(table, cap, z, make_meta(1, 0, (uint32_t)(z >> 32) | 1u), probes, coll) == INS_NEW;
    }
    for (int o = 16; o; o >>= 1) {
        mine_new += __shfl_xor_sync(0xffffffffu, mine_new, o);
        probes += __shfl_xor_sync(0xffffffffu, probes, o);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(new_count, mine_new);
        atomicAdd(probe_count, (unsigned long long)probes);
    }
}

} // namespace vsr
#endif
