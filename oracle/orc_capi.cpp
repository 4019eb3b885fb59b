/*
 * orc_capi.cpp — extern "C" surface of the oracle for tests/, smoke() and bench.py's CPU-baseline
 * legs (ctypes).  TEST INFRASTRUCTURE ONLY: the product never links or calls this.
 */
#include <cstring>
#include <fstream>
#include <sstream>

#include "vsr_oracle.h"

using namespace orc;

static Params mk(const int* q) {
    Params p;
    p.R = q[0]; p.C = q[1]; p.V = q[2]; p.L = q[3]; p.restart_limit = q[4];
    p.symmetry = q[5] != 0; p.use_view = q[6] != 0; p.invariant = q[7];
    return p;
}

extern "C" {

/* q = {R, C, V, L, restart_limit, symmetry, view, invariant} everywhere */

void orc_init_flat(const int* q, VsrFlatState* out) {
    Params p = mk(q);
    to_flat(p, init_state(p), out);
}

int orc_successors_flat(const int* q, const VsrFlatState* in, VsrFlatState* out, int* actions, int cap) {
    Params p = mk(q);
    State s = from_flat(in);
    std::vector<Succ> succ;
    successors(p, s, succ, nullptr);
    int n = 0;
    for (Succ& sc : succ) {
        if (n < cap) {
            to_flat(p, sc.s, &out[n]);
            actions[n] = sc.action;
        }
        n++;
    }
    return n;
}

/* 128-bit digest of the canonical (SYMMETRY) VIEW projection, plus the aux tie-break key */
void orc_digest_flat(const int* q, const VsrFlatState* in, int n, uint64_t* dig_out, uint32_t* auxkey_out) {
    Params p = mk(q);
    for (int i = 0; i < n; i++) {
        State cs = canonical(p, from_flat(&in[i]));
        std::string key;
        serialize(cs, !p.use_view, key);
        digest128(key, dig_out + 2 * i);
        if (auxkey_out) auxkey_out[i] = aux_key(p, cs);
    }
}

/* digest of the canonical FULL state (aux included) */
void orc_digest_full_flat(const int* q, const VsrFlatState* in, int n, uint64_t* dig_out) {
    Params p = mk(q);
    for (int i = 0; i < n; i++) {
        State cs = canonical(p, from_flat(&in[i]));
        std::string key;
        serialize(cs, true, key);
        digest128(key, dig_out + 2 * i);
    }
}

int orc_invariant_flat(const int* q, const VsrFlatState* in) {
    Params p = mk(q);
    return invariant_holds(p, from_flat(in)) ? 1 : 0;
}

/* returns total assumption violations for this state */
uint64_t orc_check_assumptions_flat(const int* q, const VsrFlatState* in) {
    Params p = mk(q);
    Assumptions a;
    check_assumptions(p, from_flat(in), a);
    return a.bag_count_gt1 + a.op_ne_loglen + a.recv_view_mismatch + a.dup_value_in_log + a.entry_not_unique +
           a.prepare_key_clash + a.slot_clash + a.view_gt_max;
}

int orc_print_flat(const int* q, const VsrFlatState* in, int with_rec_vars, char* buf, int cap) {
    Params p = mk(q);
    std::string t = print_state(p, from_flat(in), with_rec_vars != 0);
    if ((int)t.size() + 1 > cap) return -(int)t.size() - 1;
    memcpy(buf, t.c_str(), t.size() + 1);
    return (int)t.size();
}

/* parse a TLC dumpTrace file; q_out receives {R,C,V,L,restart} inferred from it.  Returns number of
   states (<0 on error).  actions[i] = VSR_ACT_* of the step that produced state i. */
int orc_parse_trace(const char* text, int* q_out, VsrFlatState* out, int* actions, int cap) {
    Params p;
    std::vector<TraceState> ts;
    std::string err = parse_trace_text(text, p, ts);
    if (!err.empty()) return -1;
    q_out[0] = p.R; q_out[1] = p.C; q_out[2] = p.V; q_out[3] = p.L; q_out[4] = p.restart_limit;
    int n = 0;
    for (TraceState& t : ts) {
        if (n < cap) {
            to_flat(p, t.s, &out[n]);
            int a = -1;
            for (int k = 0; k < VSR_NUM_ACTIONS; k++)
                if (t.action_name == action_name(k)) a = k;
            actions[n] = a;
        }
        n++;
    }
    return n;
}

/* Re-print a parsed trace in the file's own format (17-variable form when with_rec_vars = 0) with
   the original location strings, for byte comparison against the golden file. */
int orc_reprint_trace(const char* text, int with_rec_vars, char* buf, int cap) {
    Params p;
    std::vector<TraceState> ts;
    std::string err = parse_trace_text(text, p, ts);
    if (!err.empty()) return -1;
    std::string o = "<<\n";
    for (size_t i = 0; i < ts.size(); i++) {
        o += print_trace_entry(p, ts[i].s, ts[i].position, ts[i].action_name.c_str(), ts[i].location.c_str(), with_rec_vars != 0);
        o += (i + 1 < ts.size()) ? ",\n" : "\n";
    }
    o += ">>";
    if ((int)o.size() + 1 > cap) return -(int)o.size() - 1;
    memcpy(buf, o.c_str(), o.size() + 1);
    return (int)o.size();
}

/* scalars_out: [0] generated [1] distinct [2] queue [3] depth [4] rc [5] complete [6] h2_ties
   [7..15] assumption counters [16] trace length [17] seconds*1e6 [18] nlevels */
int orc_bfs(const int* q, int workers, int max_depth, uint64_t max_states, double max_seconds, int check_deadlock,
            int keep_trace, int check_assump, const char* digest_path, uint64_t* scalars_out, uint64_t* level_sizes,
            uint64_t* level_generated, int level_cap, VsrFlatState* trace_out, int* trace_actions, int trace_cap) {
    Params p = mk(q);
    BfsOptions o;
    o.workers = workers;
    o.max_depth = max_depth;
    o.max_states = max_states;
    o.max_seconds = max_seconds;
    o.check_deadlock = check_deadlock != 0;
    o.keep_trace = keep_trace != 0;
    o.check_assumptions = check_assump != 0;
    if (digest_path) o.level_digest_path = digest_path;
    BfsResult r = bfs(p, o);
    scalars_out[0] = r.generated; scalars_out[1] = r.distinct; scalars_out[2] = r.queue; scalars_out[3] = (uint64_t)r.depth;
    scalars_out[4] = (uint64_t)r.rc; scalars_out[5] = r.complete ? 1 : 0; scalars_out[6] = r.h2_ties;
    scalars_out[7] = r.as.bag_count_gt1; scalars_out[8] = r.as.op_ne_loglen; scalars_out[9] = r.as.recv_view_mismatch;
    scalars_out[10] = r.as.dup_value_in_log; scalars_out[11] = r.as.entry_not_unique; scalars_out[12] = r.as.choose_tie_diff_logs;
    scalars_out[13] = r.as.prepare_key_clash; scalars_out[14] = r.as.slot_clash; scalars_out[15] = r.as.view_gt_max;
    scalars_out[16] = r.trace.size();
    scalars_out[17] = (uint64_t)(r.seconds * 1e6);
    scalars_out[18] = r.level_sizes.size();
    for (size_t i = 0; i < r.level_sizes.size() && (int)i < level_cap; i++) level_sizes[i] = r.level_sizes[i];
    if (level_generated)
        for (size_t i = 0; i < r.level_generated.size() && (int)i < level_cap; i++) level_generated[i] = r.level_generated[i];
    for (size_t i = 0; i < r.trace.size() && (int)i < trace_cap; i++) {
        to_flat(p, r.trace[i].second, &trace_out[i]);
        trace_actions[i] = r.trace[i].first;
    }
    return r.rc;
}

} /* extern "C" */
