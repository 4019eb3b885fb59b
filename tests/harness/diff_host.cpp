/*
 * diff_host.cpp — differential test: the product's packed Next (C ABI, host side) against the CPU
 * oracle, on EVERY state of a bounded exploration.
 *
 * Walks the full (no VIEW merging) state graph from Init using the product's own successors, and
 * for each state s checks, through the oracle's general value model:
 *   - multiset { (action, canonical digest of t) : t in oracle.Next(unpack(s)) }
 *       ==  multiset of the product's successors (each counted `mult` times),
 *   - pack(unpack(s)) == s, invariant verdicts equal, aux tie-break keys equal,
 *   - the oracle's audit of the slot-encoding assumptions stays at zero.
 * usage: diff_host R V L sym(0/1) max_states [inv_mask [walks seed [start_states.hex]]]   (walks > 0: random walks instead of
 *        BFS, optionally started from given packed states — e.g. the golden trace's, to reach the state-transfer actions)
 * prints one JSON line; exit 0 iff no mismatch.
 */
#include <algorithm>
#include <cctype>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <unordered_set>
#include <vector>

#include "../../include/vsr_b200.h"
#include "../../oracle/vsr_oracle.h"

int main(int argc, char** argv) {
    if (argc < 6) { fprintf(stderr, "usage: diff_host R V L sym max_states [inv_mask]\n"); return 2; }
    const int R = atoi(argv[1]), V = atoi(argv[2]), Lm = atoi(argv[3]), sym = atoi(argv[4]);
    const size_t max_states = strtoull(argv[5], 0, 10);
    const int inv_mask = argc > 6 ? atoi(argv[6]) : 1;
    const int walks = argc > 7 ? atoi(argv[7]) : 0;
    const uint64_t seed = argc > 8 ? strtoull(argv[8], 0, 10) : 1;
    const char* seeds_path = argc > 9 ? argv[9] : nullptr; /* optional: hex-encoded packed states to start the walks from */
    char err[512];
    VsrModel* m = nullptr;
    int rc = vsr_model_create(R, 1, V, Lm, 0, sym, 1, inv_mask, &m, err, sizeof err);
    if (rc) { fprintf(stderr, "model: %s\n", err); return 2; }
    VsrModelInfo info;
    vsr_model_info(m, &info);
    const int SB = info.state_bytes;
    orc::Params p;
    p.R = R; p.C = 1; p.V = V; p.L = Lm; p.symmetry = info.symmetry != 0; p.use_view = true;
    p.invariant = (inv_mask & 1) ? 1 : ((inv_mask & 2) ? 2 : 4);

    std::vector<std::string> frontier, next;
    std::unordered_set<std::string> seen;
    std::string s0(SB, '\0');
    vsr_init(m, &s0[0]);
    frontier.push_back(s0);
    seen.insert(s0);
    size_t checked = 0, mism = 0, succ_total = 0, depth = 1, pack_bad = 0, inv_bad = 0, aux_bad = 0, canon_bad = 0;
    uint64_t assump = 0;
    uint64_t action_cover[VSR_NUM_ACTIONS] = {0}; /* successors compared, per action of Next */
    std::vector<char> succbuf((size_t)SB * 1024);
    std::vector<uint8_t> acts(1024);
    std::vector<uint32_t> mult(1024);
    VsrFlatState* f = new VsrFlatState;
    VsrFlatState* g = new VsrFlatState;
    std::string first_bad;
    bool stop = false;
    auto check = [&](const std::string& s, std::vector<std::string>& succs) {
        succs.clear();
        checked++;
        if (vsr_unpack(m, s.data(), f) != 0) { mism++; if (first_bad.empty()) first_bad = "unpack failed"; return; }
        /* round trip */
        std::string back(SB, '\0');
        int prc = vsr_pack(m, f, &back[0]);
        if (prc != 0 || back != s) { pack_bad++; if (first_bad.empty()) first_bad = "pack(unpack(s)) != s rc=" + std::to_string(prc); }
        orc::State os = orc::from_flat(f);
        {
            orc::Assumptions as;
            orc::check_assumptions(p, os, as);
            assump += as.bag_count_gt1 + as.op_ne_loglen + as.recv_view_mismatch + as.dup_value_in_log + as.entry_not_unique +
                      as.prepare_key_clash + as.slot_clash + as.view_gt_max;
        }
        if ((vsr_invariant(m, s.data()) == 0) != orc::invariant_holds(p, os)) { inv_bad++; if (first_bad.empty()) first_bad = "invariant verdict differs"; }
        if (vsr_aux_key(m, s.data()) != orc::aux_key(p, os)) { aux_bad++; if (first_bad.empty()) first_bad = "aux_key differs"; }
        if (sym) { /* canonical form is idempotent on states the engine produces */
            std::string c = s;
            vsr_canon(m, &c[0]);
            if (c != s) { canon_bad++; if (first_bad.empty()) first_bad = "state produced by step() is not canonical"; }
        }
        std::vector<orc::Succ> osucc;
        orc::successors(p, os, osucc, nullptr);
        std::map<std::pair<int, std::pair<uint64_t, uint64_t>>, int> want, got;
        for (orc::Succ& sc : osucc) {
            std::string key;
            orc::serialize(orc::canonical(p, sc.s), true, key);
            uint64_t d[2];
            orc::digest128(key, d);
            want[{sc.action, {d[0], d[1]}}]++;
        }
        int n = vsr_successors(m, s.data(), succbuf.data(), 1024, acts.data(), mult.data());
        if (n < 0) { mism++; if (first_bad.empty()) first_bad = "vsr_successors error " + std::to_string(n); return; }
        {   /* the register-mask form of the guards (the GPU scan's) against the one-candidate form: checked inside the call */
            uint32_t en[1024];
            const int ne = vsr_enabled_candidates(m, s.data(), en, 1024);
            if (ne != n) { mism++; if (first_bad.empty()) first_bad = "vsr_enabled_candidates " + std::to_string(ne) + " vs successors " + std::to_string(n); return; }
        }
        for (int i = 0; i < n; i++) {
            const char* t = succbuf.data() + (size_t)i * SB;
            if (vsr_unpack(m, t, g) != 0) { mism++; continue; }
            std::string key;
            orc::serialize(orc::canonical(p, orc::from_flat(g)), true, key);
            uint64_t d[2];
            orc::digest128(key, d);
            got[{(int)acts[i], {d[0], d[1]}}] += (int)mult[i];
            action_cover[acts[i] < VSR_NUM_ACTIONS ? acts[i] : 0] += mult[i];
            succ_total += mult[i];
            succs.emplace_back(t, SB);
        }
        if (want != got) {
            mism++;
            if (first_bad.empty()) {
                char b[256];
                snprintf(b, sizeof b, "successor multiset differs at depth %zu (oracle %zu, product %d)", depth, osucc.size(), n);
                first_bad = b;
                char* txt = new char[1 << 16];
                vsr_state_to_tla(m, s.data(), txt, 1 << 16);
                fprintf(stderr, "first mismatching state:\n%s\n", txt);
                for (auto& kv : want) if (!got.count(kv.first) || got[kv.first] != kv.second) fprintf(stderr, "  oracle-only/mismatch: action %s x%d\n", orc::action_name(kv.first.first), kv.second);
                for (auto& kv : got) if (!want.count(kv.first) || want[kv.first] != kv.second) fprintf(stderr, "  product-only/mismatch: action %s x%d\n", orc::action_name(kv.first.first), kv.second);
                delete[] txt;
            }
        }
    };
    std::vector<std::string> succs;
    size_t max_walk_depth = 0, violations_seen = 0;
    if (walks > 0) {
        /* random walks (simulation): reaches the deep states a bounded BFS cannot */
        uint64_t rng = seed * 0x9E3779B97F4A7C15ULL + 1;
        auto rnd = [&]() { rng ^= rng << 13; rng ^= rng >> 7; rng ^= rng << 17; return rng; };
        std::vector<std::string> starts;
        if (seeds_path) {
            FILE* sf = fopen(seeds_path, "r");
            char line[2048];
            while (sf && fgets(line, sizeof line, sf)) {
                std::string st;
                for (size_t i = 0; i + 1 < strlen(line) && isxdigit((unsigned char)line[i]); i += 2) {
                    unsigned v;
                    sscanf(line + i, "%2x", &v);
                    st.push_back((char)v);
                }
                if ((int)st.size() == SB) starts.push_back(st);
            }
            if (sf) fclose(sf);
        }
        for (int wk = 0; wk < walks && !stop; wk++) {
            std::string cur = starts.empty() ? s0 : starts[rnd() % starts.size()];
            for (size_t d = 1;; d++) {
                if (checked >= max_states) { stop = true; break; }
                depth = d;
                check(cur, succs);
                if (vsr_invariant(m, cur.data()) != 0) violations_seen++;
                max_walk_depth = std::max(max_walk_depth, d);
                if (succs.empty()) break;
                cur = succs[rnd() % succs.size()];
                seen.insert(cur);
            }
        }
        stop = true;
    }
    while (!frontier.empty() && !stop) {
        next.clear();
        for (const std::string& s : frontier) {
            if (checked >= max_states) { stop = true; break; }
            check(s, succs);
            for (const std::string& ts : succs)
                if (seen.insert(ts).second) next.push_back(ts);
        }
        if (!stop) { frontier.swap(next); depth++; }
    }
    std::string cover;
    for (int a = 1; a <= 15; a++) cover += (a > 1 ? ", " : "") + std::to_string(action_cover[a]);
    printf("{\"R\": %d, \"V\": %d, \"L\": %d, \"sym\": %d, \"state_bytes\": %d, \"checked\": %zu, \"distinct_full\": %zu, \"successors\": %zu, "
           "\"depth\": %zu, \"complete\": %d, \"mismatches\": %zu, \"pack_roundtrip_bad\": %zu, \"invariant_bad\": %zu, \"aux_key_bad\": %zu, "
           "\"canon_bad\": %zu, \"assumption_violations\": %llu, \"walks\": %d, \"max_walk_depth\": %zu, \"violating_states_seen\": %zu, \"first_bad\": \"%s\", \"action_coverage\": [%s]}\n",
           R, V, Lm, sym, SB, checked, seen.size(), succ_total, depth, stop ? 0 : 1, mism, pack_bad, inv_bad, aux_bad, canon_bad,
           (unsigned long long)assump, walks, max_walk_depth, violations_seen, first_bad.c_str(), cover.c_str());
    return (mism || pack_bad || inv_bad || aux_bad || canon_bad || assump) ? 1 : 0;
}
