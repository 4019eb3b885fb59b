/*
 * bfs.cpp — level-synchronous BFS of the oracle (test infrastructure / reported CPU baseline only).
 *
 * Engine stages TLC wraps around Next, restated without packing:
 *   successors -> canonical() under SYMMETRY -> VIEW projection -> exact 128-bit digest of the
 *   serialised view -> seen-set -> invariant on new states -> next level.
 * Same-level states with equal VIEW but different aux variables are resolved by the
 * label-independent rule "smallest aux_key wins" and counted in h2_ties.
 */
#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <functional>
#include <mutex>
#include <thread>
#include <unordered_map>
#ifdef __GLIBC__
#include <malloc.h>
#endif

#include "vsr_oracle.h"

namespace orc {

namespace {
struct Dig {
    uint64_t a, b;
    bool operator==(const Dig& o) const { return a == o.a && b == o.b; }
    bool operator<(const Dig& o) const { return a != o.a ? a < o.a : b < o.b; }
};
struct DigHash {
    size_t operator()(const Dig& d) const { return (size_t)d.a; }
};
struct Cand {
    Dig d;
    uint32_t auxkey;
    uint32_t action;
    uint64_t parent;
    std::string ser;
};
struct Slot {
    uint32_t level;
    uint32_t pending; /* index into the shard's pending list while level == current */
};
constexpr int NSHARD = 256; /* >= the thread count of a big host: the insert phase hands out whole shards */
struct Shard {
    std::unordered_map<Dig, Slot, DigHash> map;
    std::vector<Cand> pending;
};
/* W workers that live as long as the BFS (worker 0 is the caller).  Persistent on purpose: a thread keeps its malloc arena,
   so what worker w allocated in one phase it can free in a later one without taking another thread's arena lock, and a
   many-core host does not start 2 x W threads per batch. */
class Pool {
public:
    explicit Pool(int w) : W(w) {
        for (int i = 1; i < W; i++) th.emplace_back([this, i] { loop(i); });
    }
    ~Pool() {
        {
            std::lock_guard<std::mutex> g(m);
            stop = true;
            gen++;
        }
        cv.notify_all();
        for (auto& t : th) t.join();
    }
    /* runs f(w) on every worker w in 0..W-1 and returns when all have finished */
    void run(const std::function<void(int)>& f) {
        {
            std::lock_guard<std::mutex> g(m);
            fn = &f;
            pending = W - 1;
            gen++;
        }
        cv.notify_all();
        f(0);
        std::unique_lock<std::mutex> g(m);
        done.wait(g, [this] { return pending == 0; });
        fn = nullptr;
    }
private:
    void loop(int w) {
        uint64_t seen = 0;
        for (;;) {
            const std::function<void(int)>* f;
            {
                std::unique_lock<std::mutex> g(m);
                cv.wait(g, [&] { return gen != seen; });
                seen = gen;
                if (stop) return;
                f = fn;
            }
            (*f)(w);
            {
                std::lock_guard<std::mutex> g(m);
                if (--pending == 0) done.notify_one();
            }
        }
    }
    const int W;
    std::vector<std::thread> th;
    std::mutex m;
    std::condition_variable cv, done;
    const std::function<void(int)>* fn = nullptr;
    uint64_t gen = 0;
    int pending = 0;
    bool stop = false;
};

double now_s() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
} // namespace

BfsResult bfs(const Params& p, const BfsOptions& o) {
    BfsResult res;
    double t0 = now_s();
    std::vector<Shard> shards(NSHARD);
    int W = std::max(1, o.workers);
#ifdef __GLIBC__
    if (const char* mo = getenv("ORC_BFS_MALLOPT")) {
        /* allocator tuning for the timed CPU baseline only (bench.py tries both settings and keeps the faster): every successor
           is a fresh graph of small heap objects; "1" keeps the arenas from handing pages back and asking for them again (each
           shrink/grow is an mprotect/madvise under the process-wide mmap lock, which a many-thread run queues on), "0" restores
           glibc's defaults.  Process-wide and sticky, hence opt-in. */
        const bool on = mo[0] == '1';
        mallopt(M_TRIM_THRESHOLD, on ? 1 << 30 : 128 << 10);
        mallopt(M_MMAP_THRESHOLD, on ? 64 << 20 : 128 << 10);
    }
#endif
    Pool pool(W);

    /* global per-state trace records */
    std::vector<uint64_t> parent_of;
    std::vector<uint8_t> action_of;
    std::vector<std::string> all_states; /* only when keep_trace */

    FILE* dig_file = nullptr;
    if (!o.level_digest_path.empty()) dig_file = fopen(o.level_digest_path.c_str(), "wb");

    auto make_cand = [&](const State& raw, uint64_t parent, int action, Cand& c) {
        State cs = canonical(p, raw);
        std::string key;
        serialize(cs, !p.use_view, key);
        uint64_t d[2];
        digest128(key, d);
        c.d = Dig{d[0], d[1]};
        c.auxkey = aux_key(p, cs);
        c.action = (uint32_t)action;
        c.parent = parent;
        serialize(cs, true, c.ser);
    };

    /* level 1: Init */
    std::vector<std::string> frontier;
    uint64_t frontier_base = 0;
    {
        State s0 = init_state(p);
        Cand c;
        make_cand(s0, (uint64_t)-1, VSR_ACT_INIT, c);
        shards[c.d.a % NSHARD].map[c.d] = Slot{1, 0};
        frontier.push_back(c.ser);
        parent_of.push_back((uint64_t)-1);
        action_of.push_back(VSR_ACT_INIT);
        if (o.keep_trace) all_states.push_back(c.ser);
        res.generated = 1;
        res.distinct = 1;
        res.level_sizes.push_back(1);
        if (dig_file) {
            uint64_t n = 1;
            fwrite(&n, 8, 1, dig_file);
            fwrite(&c.d, 16, 1, dig_file);
        }
        if (o.check_assumptions) check_assumptions(p, s0, res.as);
        if (!invariant_holds(p, s0)) {
            res.rc = 12;
            res.trace.push_back({VSR_ACT_INIT, s0});
        }
    }
    int level = 1;
    double t_expand = 0, t_insert = 0, t_collect = 0; /* ORC_BFS_PROFILE=1: where the wall time of a run went */
    uint64_t violating_id = (uint64_t)-1, deadlock_id = (uint64_t)-1;
    bool stopped_early = false;

    while (!frontier.empty() && res.rc == 0) {
        if (o.max_depth && level >= o.max_depth) { stopped_early = true; break; }
        if (o.max_states && res.distinct >= o.max_states) { stopped_early = true; break; }
        if (o.max_seconds > 0 && now_s() - t0 >= o.max_seconds) { stopped_early = true; break; }
        const int cur = level + 1; /* depth of the states generated now */
        uint64_t gen_this = 0;
        /* states per parallel region: enough per thread that starting and joining the threads does not dominate on a
           many-core host (bench.py's CPU baseline runs this with every core of the box) */
        const size_t BATCH = std::max<size_t>(1 << 15, (size_t)W << 12);
        std::atomic<uint64_t> h2{0};
        std::atomic<uint64_t> dead{(uint64_t)-1};
        std::vector<Assumptions> was(W);
        size_t expanded = 0;
        for (size_t b0 = 0; b0 < frontier.size() && !stopped_early; b0 += BATCH) {
            size_t b1 = std::min(frontier.size(), b0 + BATCH);
            std::vector<std::vector<std::vector<Cand>>> buckets(W, std::vector<std::vector<Cand>>(NSHARD));
            std::atomic<size_t> next{b0};
            std::atomic<uint64_t> gen{0};
            auto expand = [&](int w) {
                std::vector<Succ> succ;
                for (;;) {
                    size_t i0 = next.fetch_add(64);
                    if (i0 >= b1) break;
                    size_t i1 = std::min(b1, i0 + 64);
                    for (size_t i = i0; i < i1; i++) {
                        State s = deserialize(p, frontier[i]);
                        succ.clear();
                        successors(p, s, succ, o.check_assumptions ? &was[w] : nullptr);
                        gen.fetch_add(succ.size(), std::memory_order_relaxed);
                        if (succ.empty() && o.check_deadlock) {
                            uint64_t id = frontier_base + i, exp = dead.load();
                            while (id < exp && !dead.compare_exchange_weak(exp, id)) {}
                        }
                        for (Succ& sc : succ) {
                            Cand c;
                            make_cand(sc.s, frontier_base + i, sc.action, c);
                            buckets[w][c.d.a % NSHARD].push_back(std::move(c));
                        }
                    }
                }
            };
            double tp0 = now_s();
            pool.run(expand);
            double tp1 = now_s();
            t_expand += tp1 - tp0;
            gen_this += gen.load();
            expanded = b1;
            /* insert phase: one shard at a time per thread */
            std::atomic<int> next_shard{0};
            auto insert = [&](int) {
                for (;;) {
                    int sh = next_shard.fetch_add(1);
                    if (sh >= NSHARD) break;
                    Shard& S = shards[sh];
                    for (int w = 0; w < W; w++)
                        for (Cand& c : buckets[w][sh]) {
                            auto it = S.map.find(c.d);
                            if (it == S.map.end()) {
                                S.map.emplace(c.d, Slot{(uint32_t)cur, (uint32_t)S.pending.size()});
                                S.pending.push_back(std::move(c));
                            } else if (it->second.level == (uint32_t)cur) {
                                Cand& inc = S.pending[it->second.pending];
                                if (c.auxkey != inc.auxkey) h2.fetch_add(1, std::memory_order_relaxed);
                                /* deterministic winner: smallest (aux_key, parent, action) */
                                if (c.auxkey < inc.auxkey ||
                                    (c.auxkey == inc.auxkey && (c.parent < inc.parent || (c.parent == inc.parent && c.action < inc.action))))
                                    inc = std::move(c);
                            }
                        }
                }
            };
            pool.run(insert);
            /* the losers (most candidates are duplicates) go back to the arena of the worker that made them */
            pool.run([&](int w) { std::vector<std::vector<Cand>>().swap(buckets[w]); });
            t_insert += now_s() - tp1;
            if (o.max_seconds > 0 && now_s() - t0 >= o.max_seconds && b1 < frontier.size()) stopped_early = true;
        }
        for (const Assumptions& a : was) {
            res.as.bag_count_gt1 += a.bag_count_gt1; res.as.op_ne_loglen += a.op_ne_loglen;
            res.as.choose_tie_diff_logs += a.choose_tie_diff_logs;
        }
        res.h2_ties += h2.load();
        res.generated += gen_this;
        res.level_generated.push_back(gen_this);

        /* collect the new level in digest order */
        double tc0 = now_s();
        std::vector<Cand*> fresh;
        for (Shard& S : shards)
            for (Cand& c : S.pending) fresh.push_back(&c);
        std::sort(fresh.begin(), fresh.end(), [](const Cand* x, const Cand* y) { return x->d < y->d; });
        uint64_t new_base = frontier_base + frontier.size();
        std::vector<std::string> next_frontier;
        next_frontier.reserve(fresh.size());
        if (dig_file && !stopped_early && !fresh.empty()) {
            uint64_t n = fresh.size();
            fwrite(&n, 8, 1, dig_file);
            for (Cand* c : fresh) fwrite(&c->d, 16, 1, dig_file);
        }
        for (size_t k = 0; k < fresh.size(); k++) {
            Cand* c = fresh[k];
            if (o.keep_trace) {
                parent_of.push_back(c->parent);
                action_of.push_back((uint8_t)c->action);
                all_states.push_back(c->ser);
            }
            if (violating_id == (uint64_t)-1 && (p.invariant != 0 || o.check_assumptions)) {
                State s = deserialize(p, c->ser);
                if (o.check_assumptions) check_assumptions(p, s, res.as);
                if (!invariant_holds(p, s) && o.stop_on_violation) violating_id = new_base + k;
                else if (!invariant_holds(p, s)) res.rc = 12;
            }
            next_frontier.push_back(std::move(c->ser));
        }
        for (Shard& S : shards) { S.pending.clear(); S.pending.shrink_to_fit(); }
        res.distinct += next_frontier.size();
        t_collect += now_s() - tc0;
        if (stopped_early) {
            /* partial level: report what was found, queue = unexpanded part + new states */
            res.queue = (frontier.size() - expanded) + next_frontier.size();
            res.level_sizes.push_back(next_frontier.size());
            level = cur;
            break;
        }
        if (!next_frontier.empty()) {
            res.level_sizes.push_back(next_frontier.size());
            level = cur;
        }
        if (dead.load() != (uint64_t)-1) { deadlock_id = dead.load(); res.rc = 11; }
        if (violating_id != (uint64_t)-1) res.rc = 12;
        frontier_base = new_base;
        frontier.swap(next_frontier);
        if (res.rc != 0) break;
    }
    if (dig_file) fclose(dig_file);
    if (getenv("ORC_BFS_PROFILE"))
        fprintf(stderr, "orc bfs: %d threads, expand %.2f s, insert %.2f s, collect (serial) %.2f s\n", W, t_expand, t_insert, t_collect);
    res.depth = level;
    if (res.rc == 0 && !stopped_early) res.complete = true;
    if (res.rc != 0 || (stopped_early && res.queue == 0)) res.queue = frontier.size();
    if (res.complete) res.queue = 0;

    uint64_t bad = violating_id != (uint64_t)-1 ? violating_id : deadlock_id;
    if (bad != (uint64_t)-1 && o.keep_trace) {
        std::vector<uint64_t> chain;
        for (uint64_t id = bad; id != (uint64_t)-1; id = parent_of[id]) chain.push_back(id);
        std::reverse(chain.begin(), chain.end());
        /* the stored states are canonical representatives; re-execute from Init so that every
           consecutive pair of the reported trace is a literal step of Next (TLC does the same) */
        State cur = init_state(p);
        res.trace.push_back({VSR_ACT_INIT, cur});
        for (size_t k = 1; k < chain.size(); k++) {
            std::vector<Succ> succ;
            successors(p, cur, succ, nullptr);
            bool found = false;
            for (Succ& sc : succ) {
                std::string ser;
                serialize(canonical(p, sc.s), true, ser);
                if (ser == all_states[chain[k]] && sc.action == (int)action_of[chain[k]]) {
                    cur = sc.s;
                    res.trace.push_back({sc.action, cur});
                    found = true;
                    break;
                }
            }
            if (!found) { res.trace.clear(); break; }
        }
    }
    res.seconds = now_s() - t0;
    return res;
}

} // namespace orc
