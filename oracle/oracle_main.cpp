/*
 * oracle_main.cpp — command line for the CPU oracle (test infrastructure / reported CPU baseline).
 *   vsr_oracle bfs R V L [--workers N] [--depth D] [--seconds S] [--states N] [--nosym] [--noview]
 *                        [--inv K] [--deadlock-check] [--restart N] [--digests FILE] [--notrace] [--noassume]
 *   vsr_oracle replay TRACE.txt      replay a TLC dumpTrace file through Next
 */
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>

#include "vsr_oracle.h"

using namespace orc;

static int cmd_replay(const char* path) {
    std::ifstream f(path);
    if (!f) { fprintf(stderr, "cannot open %s\n", path); return 2; }
    std::stringstream ss;
    ss << f.rdbuf();
    Params p;
    std::vector<TraceState> ts;
    std::string err = parse_trace_text(ss.str(), p, ts);
    if (!err.empty()) { fprintf(stderr, "parse error: %s\n", err.c_str()); return 2; }
    printf("parsed %zu states: R=%d C=%d V=%d L=%d\n", ts.size(), p.R, p.C, p.V, p.L);
    int bad = 0;
    for (size_t i = 0; i + 1 < ts.size(); i++) {
        std::vector<Succ> succ;
        successors(p, ts[i].s, succ, nullptr);
        bool ok = false;
        for (Succ& sc : succ)
            if (cmp_state(sc.s, ts[i + 1].s, true) == 0 && ts[i + 1].action_name == action_name(sc.action)) ok = true;
        printf("  %2zu -> %2zu %-24s %s (%zu successors)\n", i + 1, i + 2, ts[i + 1].action_name.c_str(), ok ? "ok" : "NOT A STEP", succ.size());
        if (!ok) bad++;
    }
    p.invariant = 1;
    bool last_ok = invariant_holds(p, ts.back().s);
    printf("AcknowledgedWriteNotLost on last state: %s\n", last_ok ? "holds (unexpected)" : "violated");
    return (bad == 0 && !last_ok) ? 0 : 1;
}

int main(int argc, char** argv) {
    if (argc >= 3 && !strcmp(argv[1], "replay")) return cmd_replay(argv[2]);
    if (argc >= 5 && !strcmp(argv[1], "bfs")) {
        Params p;
        p.R = atoi(argv[2]); p.V = atoi(argv[3]); p.L = atoi(argv[4]);
        BfsOptions o;
        for (int i = 5; i < argc; i++) {
            if (!strcmp(argv[i], "--workers")) o.workers = atoi(argv[++i]);
            else if (!strcmp(argv[i], "--depth")) o.max_depth = atoi(argv[++i]);
            else if (!strcmp(argv[i], "--seconds")) o.max_seconds = atof(argv[++i]);
            else if (!strcmp(argv[i], "--states")) o.max_states = strtoull(argv[++i], 0, 10);
            else if (!strcmp(argv[i], "--nosym")) p.symmetry = false;
            else if (!strcmp(argv[i], "--noview")) p.use_view = false;
            else if (!strcmp(argv[i], "--inv")) p.invariant = atoi(argv[++i]);
            else if (!strcmp(argv[i], "--restart")) p.restart_limit = atoi(argv[++i]);
            else if (!strcmp(argv[i], "--deadlock-check")) o.check_deadlock = true;
            else if (!strcmp(argv[i], "--digests")) o.level_digest_path = argv[++i];
            else if (!strcmp(argv[i], "--notrace")) o.keep_trace = false;
            else if (!strcmp(argv[i], "--noassume")) o.check_assumptions = false;
        }
        BfsResult r = bfs(p, o);
        printf("R=%d V=%d L=%d sym=%d view=%d inv=%d workers=%d\n", p.R, p.V, p.L, p.symmetry, p.use_view, p.invariant, o.workers);
        for (size_t i = 0; i < r.level_sizes.size(); i++)
            printf("  depth %3zu: %12llu new, %12llu generated from it\n", i + 1, (unsigned long long)r.level_sizes[i],
                   i < r.level_generated.size() ? (unsigned long long)r.level_generated[i] : 0ULL);
        printf("%llu states generated, %llu distinct states found, %llu states left on queue.\n", (unsigned long long)r.generated,
               (unsigned long long)r.distinct, (unsigned long long)r.queue);
        printf("depth %d  rc %d  complete %d  %.2f s  %.0f distinct/s\n", r.depth, r.rc, r.complete, r.seconds, r.distinct / std::max(r.seconds, 1e-9));
        printf("h2_ties %llu  choose_tie_diff_logs %llu\n", (unsigned long long)r.h2_ties, (unsigned long long)r.as.choose_tie_diff_logs);
        printf("assumptions: count>1 %llu, op!=len %llu, recv_view %llu, dup_value %llu, entry_unique %llu, prepkey %llu, slot %llu, view>K %llu\n",
               (unsigned long long)r.as.bag_count_gt1, (unsigned long long)r.as.op_ne_loglen, (unsigned long long)r.as.recv_view_mismatch,
               (unsigned long long)r.as.dup_value_in_log, (unsigned long long)r.as.entry_not_unique, (unsigned long long)r.as.prepare_key_clash,
               (unsigned long long)r.as.slot_clash, (unsigned long long)r.as.view_gt_max);
        if (!r.trace.empty()) {
            printf("trace (%zu states):\n", r.trace.size());
            for (size_t i = 0; i < r.trace.size(); i++) printf("  %2zu %s\n", i + 1, action_name(r.trace[i].first));
        }
        return r.rc;
    }
    fprintf(stderr, "usage: vsr_oracle bfs R V L [opts] | replay TRACE\n");
    return 2;
}
