/*
 * vsr_flat_conv.h — packed state <-> VsrFlatState (include/vsr_flat.h).  Host only.
 * unpack() is the definition of what a packed word vector MEANS as a VSR.tla state
 * (vsr-revisited/paper/VSR.tla:119-138); pack() is its inverse and rejects (E_UNSUPPORTED /
 * E_SLOT_OCCUPIED) any state the slot encoding of vsr_layout.h cannot hold.
 */
#ifndef VSR_FLAT_CONV_H
#define VSR_FLAT_CONV_H

#include <string.h>

#include "../../include/vsr_flat.h"
#include "vsr_actions.h"

namespace vsr {

template <class L> struct Conv {
    typedef Ops<L> O_;
    static constexpr int R = L::R, V = L::V, K = L::K, O = L::O;

    static VsrEntry entry_of(const uint32_t* w, int x /*1-based value id*/) {
        VsrEntry e;
        e.view = (uint8_t)VGET(L, PR_VIEW, w, x - 1);
        e.operation = (uint8_t)x;
        e.client = 1;
        e.req = (uint8_t)VGET(L, PR_REQ, w, x - 1);
        return e;
    }
    static void blank(VsrMsg* m, int type) {
        memset(m, 0, sizeof(*m));
        m->type = (uint8_t)type;
        m->view = m->src = m->dest = m->op = m->commit = m->lnv = m->first_op = m->x = VSR_ABSENT;
    }
    template <int B> static int put_log(const uint32_t* w, int row, VsrMsg* m, int lo0 /*0-based first pos*/) {
        int n = 0;
        for (int i = lo0; i < V; i++) {
            const int x = (int)fget<B, L::OB>(w, row * V + i);
            if (!x) break;
            m->log[n++] = entry_of(w, x);
        }
        m->has_log = 1;
        m->log_lo = (uint8_t)(lo0 + 1);
        m->log_n = (uint8_t)n;
        return n;
    }

    static int unpack(const uint32_t* w, VsrFlatState* f) {
        memset(f, 0, sizeof(*f));
        f->R = R; f->C = 1; f->V = V;
        f->aux_svc = (uint8_t)VGET(L, AUX_SVC, w, 0);
        f->aux_restart = 0;
        for (int x = 0; x < V; x++) f->acked[x] = (uint8_t)VGET(L, ACKED, w, x);
        for (int r = 0; r < R; r++) {
            VsrReplica& q = f->rep[r];
            q.status = (uint8_t)VGET(L, STATUS, w, r);
            q.view = (uint8_t)VGET(L, VIEWN, w, r);
            q.commit = (uint8_t)VGET(L, COMMIT, w, r);
            q.lnv = (uint8_t)VGET(L, LNV, w, r);
            q.sent_dvc = (uint8_t)VGET(L, SENT_DVC, w, r);
            q.sent_sv = (uint8_t)VGET(L, SENT_SV, w, r);
            q.rec_number = 0;
            int n = 0;
            for (int i = 0; i < V; i++) {
                const int x = (int)VGET(L, LOG, w, r * V + i);
                if (!x) break;
                q.log[n++] = entry_of(w, x);
            }
            q.log_n = (uint8_t)n;
            q.op = (uint8_t)n; /* rep_op_number = Len(rep_log) */
            for (int p = 0; p < R; p++) q.peer_op[p] = (uint8_t)VGET(L, PEER, w, r * R + p);
            q.client_table[0].req = (uint8_t)VGET(L, CT_REQ, w, r);
            q.client_table[0].op = (uint8_t)VGET(L, CT_OP, w, r);
            q.client_table[0].executed = (uint8_t)VGET(L, CT_EXEC, w, r);
            for (int s = 0; s < R; s++) {
                if (VGET(L, SVC_MASK, w, r * R + s)) {
                    VsrMsg* m = &q.svc_recv[q.n_svc++];
                    blank(m, VSR_MT_SVC);
                    m->view = q.view; m->dest = (uint8_t)(r + 1); m->src = (uint8_t)(s + 1);
                }
                if (VGET(L, DVC_MASK, w, r * R + s)) {
                    VsrMsg* m = &q.dvc_recv[q.n_dvc++];
                    blank(m, VSR_MT_DVC);
                    m->view = q.view; m->dest = (uint8_t)(r + 1); m->src = (uint8_t)(s + 1);
                    if (s == r) {
                        m->op = (uint8_t)put_log<L::SELF_LOG_B>(w, r, m, 0);
                        m->lnv = (uint8_t)VGET(L, SELF_LNV, w, r);
                        m->commit = (uint8_t)VGET(L, SELF_COMMIT, w, r);
                    } else {
                        if (q.view < 2 || O_::primary(q.view) != r) return E_MISSING_PAYLOAD;
                        const int di = (q.view - 2) * O + O_::oidx(r, s);
                        if (VGET(L, DVC_ST, w, di) == ST_ABSENT) return E_MISSING_PAYLOAD;
                        m->op = (uint8_t)put_log<L::DVC_LOG_B>(w, di, m, 0);
                        m->lnv = (uint8_t)VGET(L, DVC_LNV, w, di);
                        m->commit = (uint8_t)VGET(L, DVC_COMMIT, w, di);
                    }
                }
            }
        }
        /* the bag */
        auto add = [&](int st) -> VsrMsg* {
            if (f->n_msgs >= VSR_MAX_MSGS) return nullptr;
            VsrMsg* m = &f->msgs[f->n_msgs++];
            m->count = st == ST_PENDING ? 1 : 0;
            return m;
        };
        for (int idx = 0; idx < L::NSVC; idx++) {
            const int st = (int)VGET(L, SVC_ST, w, idx);
            if (!st) continue;
            const int dp = idx % O, src = (idx / O) % R, v = idx / (O * R) + 2;
            VsrMsg* m = add(st); if (!m) return E_OVERFLOW;
            const uint8_t c = m->count; blank(m, VSR_MT_SVC); m->count = c;
            m->view = (uint8_t)v; m->src = (uint8_t)(src + 1); m->dest = (uint8_t)(O_::oinv(src, dp) + 1);
        }
        for (int idx = 0; idx < L::NDVC; idx++) {
            const int st = (int)VGET(L, DVC_ST, w, idx);
            if (!st) continue;
            const int v = idx / O + 2, p = O_::primary(v), src = O_::oinv(p, idx % O);
            VsrMsg* m = add(st); if (!m) return E_OVERFLOW;
            const uint8_t c = m->count; blank(m, VSR_MT_DVC); m->count = c;
            m->view = (uint8_t)v; m->src = (uint8_t)(src + 1); m->dest = (uint8_t)(p + 1);
            m->op = (uint8_t)put_log<L::DVC_LOG_B>(w, idx, m, 0);
            m->lnv = (uint8_t)VGET(L, DVC_LNV, w, idx);
            m->commit = (uint8_t)VGET(L, DVC_COMMIT, w, idx);
        }
        for (int idx = 0; idx < L::NSV; idx++) {
            const int st = (int)VGET(L, SV_ST, w, idx);
            if (!st) continue;
            const int v = idx / O + 2, p = O_::primary(v), d = O_::oinv(p, idx % O);
            VsrMsg* m = add(st); if (!m) return E_OVERFLOW;
            const uint8_t c = m->count; blank(m, VSR_MT_SV); m->count = c;
            m->view = (uint8_t)v; m->src = (uint8_t)(p + 1); m->dest = (uint8_t)(d + 1);
            m->op = (uint8_t)put_log<L::SV_LOG_B>(w, v - 2, m, 0);
            m->commit = (uint8_t)VGET(L, SV_COMMIT, w, v - 2);
        }
        for (int x = 0; x < V; x++) {
            const int pv = (int)VGET(L, PR_VIEW, w, x);
            if (!pv) continue;
            const int p = O_::primary(pv);
            for (int dp = 0; dp < O; dp++) {
                VsrMsg* m = add(VGET(L, PR_CONS, w, x * O + dp) ? ST_CONSUMED : ST_PENDING); if (!m) return E_OVERFLOW;
                const uint8_t c = m->count; blank(m, VSR_MT_PREPARE); m->count = c;
                m->view = (uint8_t)pv; m->src = (uint8_t)(p + 1); m->dest = (uint8_t)(O_::oinv(p, dp) + 1);
                m->has_entry = 1; m->entry = entry_of(w, x + 1);
                m->op = (uint8_t)VGET(L, PR_OP, w, x);
                m->commit = (uint8_t)VGET(L, PR_COMMIT, w, x);
            }
        }
        for (int idx = 0; idx < L::NPOK; idx++) {
            const int st = (int)VGET(L, POK_ST, w, idx);
            if (!st) continue;
            const int sp = idx % O, nn = (idx / O) % V + 1, v = idx / (O * V) + 1, p = O_::primary(v);
            VsrMsg* m = add(st); if (!m) return E_OVERFLOW;
            const uint8_t c = m->count; blank(m, VSR_MT_PREPAREOK); m->count = c;
            m->view = (uint8_t)v; m->op = (uint8_t)nn; m->dest = (uint8_t)(p + 1); m->src = (uint8_t)(O_::oinv(p, sp) + 1);
        }
        for (int gi = 0; gi < L::NGS; gi++) {
            const int v = gi / O + 2, p = O_::primary(v), src = O_::oinv(p, gi % O);
            const int st = (int)VGET(L, GS_ST, w, gi);
            if (st) {
                VsrMsg* m = add(st); if (!m) return E_OVERFLOW;
                const uint8_t c = m->count; blank(m, VSR_MT_GETSTATE); m->count = c;
                m->view = (uint8_t)v; m->op = (uint8_t)VGET(L, GS_T, w, gi);
                m->dest = (uint8_t)(VGET(L, GS_DEST, w, gi) + 1); m->src = (uint8_t)(src + 1);
            }
            const int ns = (int)VGET(L, NS_ST, w, gi);
            if (ns) {
                VsrMsg* m = add(ns); if (!m) return E_OVERFLOW;
                const uint8_t c = m->count; blank(m, VSR_MT_NEWSTATE); m->count = c;
                const int t = (int)VGET(L, GS_T, w, gi);
                m->view = (uint8_t)v;
                const int n = put_log<L::NS_LOG_B>(w, gi, m, t);
                m->first_op = (uint8_t)(t + 1);
                m->op = (uint8_t)(t + n);
                m->commit = (uint8_t)VGET(L, NS_COMMIT, w, gi);
                m->dest = (uint8_t)(src + 1);                          /* back to the GetState's source */
                m->src = (uint8_t)(VGET(L, GS_DEST, w, gi) + 1);
            }
        }
        return 0;
    }

    /* ------------------------------------------------------------------ pack */
    static bool log_to_ids(const VsrEntry* lg, int n, int lo0, uint32_t ids[/*V*/]) {
        for (int i = 0; i < V; i++) ids[i] = 0;
        if (lo0 + n > V) return false;
        for (int i = 0; i < n; i++) {
            if (lg[i].operation < 1 || lg[i].operation > V || lg[i].client != 1) return false;
            ids[lo0 + i] = lg[i].operation;
        }
        return true;
    }
    template <int B> static bool same_log(const uint32_t* w, int row, const uint32_t ids[]) {
        for (int i = 0; i < V; i++)
            if (fget<B, L::OB>(w, row * V + i) != ids[i]) return false;
        return true;
    }
    template <int B> static void set_log(uint32_t* w, int row, const uint32_t ids[]) {
        for (int i = 0; i < V; i++) fset<B, L::OB>(w, row * V + i, ids[i]);
    }

    static int pack(const VsrFlatState* f, uint32_t* w, bool symmetry) {
        if (f->R != R || f->V != V || f->C != 1 || f->aux_restart != 0) return E_UNSUPPORTED;
        for (int i = 0; i < L::NW; i++) w[i] = 0;
        uint32_t ids[V];
        /* the bag first: it holds the Prepare slots that define log entries */
        for (int i = 0; i < f->n_msgs; i++) {
            const VsrMsg& m = f->msgs[i];
            if (m.count > 1) return E_UNSUPPORTED;
            const uint32_t st = m.count ? ST_PENDING : ST_CONSUMED;
            if (m.view == VSR_ABSENT || m.view < 1 || m.view > K) return E_UNSUPPORTED;
            const int v = m.view, p = O_::primary(v), src = m.src - 1, dest = m.dest - 1;
            if (src < 0 || src >= R || dest < 0 || dest >= R || src == dest) return E_UNSUPPORTED;
            switch (m.type) {
            case VSR_MT_SVC: {
                if (v < 2) return E_UNSUPPORTED;
                const int idx = O_::svc_slot(v, src, dest);
                if (VGET(L, SVC_ST, w, idx)) return E_SLOT_OCCUPIED;
                VSET(L, SVC_ST, w, idx, st);
                break;
            }
            case VSR_MT_DVC: {
                if (v < 2 || dest != p || m.has_log != 1 || m.op != m.log_n || m.lnv > K || m.commit > V) return E_UNSUPPORTED;
                const int idx = (v - 2) * O + O_::oidx(p, src);
                if (VGET(L, DVC_ST, w, idx)) return E_SLOT_OCCUPIED;
                if (!log_to_ids(m.log, m.log_n, 0, ids)) return E_UNSUPPORTED;
                VSET(L, DVC_ST, w, idx, st);
                set_log<L::DVC_LOG_B>(w, idx, ids);
                VSET(L, DVC_LNV, w, idx, m.lnv);
                VSET(L, DVC_COMMIT, w, idx, m.commit);
                break;
            }
            case VSR_MT_SV: {
                if (v < 2 || src != p || m.has_log != 1 || m.op != m.log_n || m.commit > V) return E_UNSUPPORTED;
                const int idx = (v - 2) * O + O_::oidx(p, dest);
                if (VGET(L, SV_ST, w, idx)) return E_SLOT_OCCUPIED;
                if (!log_to_ids(m.log, m.log_n, 0, ids)) return E_UNSUPPORTED;
                bool first = true;
                for (int dp = 0; dp < O; dp++) first = first && VGET(L, SV_ST, w, (v - 2) * O + dp) == 0;
                if (!first && (!same_log<L::SV_LOG_B>(w, v - 2, ids) || VGET(L, SV_COMMIT, w, v - 2) != m.commit)) return E_SLOT_OCCUPIED;
                VSET(L, SV_ST, w, idx, st);
                set_log<L::SV_LOG_B>(w, v - 2, ids);
                VSET(L, SV_COMMIT, w, v - 2, m.commit);
                break;
            }
            case VSR_MT_PREPARE: {
                const int x = m.entry.operation;
                if (!m.has_entry || x < 1 || x > V || src != p || m.entry.view != v || m.entry.client != 1 || m.op < 1 || m.op > V ||
                    m.commit > V || m.entry.req < 1 || m.entry.req > V)
                    return E_UNSUPPORTED;
                if (VGET(L, PR_VIEW, w, x - 1)) {
                    if ((int)VGET(L, PR_VIEW, w, x - 1) != v || VGET(L, PR_REQ, w, x - 1) != m.entry.req || VGET(L, PR_OP, w, x - 1) != m.op ||
                        VGET(L, PR_COMMIT, w, x - 1) != m.commit)
                        return E_SLOT_OCCUPIED;
                } else {
                    VSET(L, PR_VIEW, w, x - 1, v);
                    VSET(L, PR_REQ, w, x - 1, m.entry.req);
                    VSET(L, PR_OP, w, x - 1, m.op);
                    VSET(L, PR_COMMIT, w, x - 1, m.commit);
                    for (int dp = 0; dp < O; dp++) VSET(L, PR_CONS, w, (x - 1) * O + dp, 1); /* until a copy is seen */
                    VSET(L, ACKED, w, x - 1, 3); /* marker: need every dest */
                }
                VSET(L, PR_CONS, w, (x - 1) * O + O_::oidx(p, dest), m.count ? 0 : 1);
                break;
            }
            case VSR_MT_PREPAREOK: {
                if (dest != p || m.op < 1 || m.op > V) return E_UNSUPPORTED;
                const int idx = ((v - 1) * V + (m.op - 1)) * O + O_::oidx(p, src);
                if (VGET(L, POK_ST, w, idx)) return E_SLOT_OCCUPIED;
                VSET(L, POK_ST, w, idx, st);
                break;
            }
            case VSR_MT_GETSTATE: {
                if (v < 2 || src == p || m.op > V) return E_UNSUPPORTED;
                const int gi = (v - 2) * O + O_::oidx(p, src);
                if (VGET(L, GS_ST, w, gi)) return E_SLOT_OCCUPIED;
                VSET(L, GS_ST, w, gi, st);
                VSET(L, GS_T, w, gi, m.op);
                VSET(L, GS_DEST, w, gi, dest);
                break;
            }
            case VSR_MT_NEWSTATE: break; /* second pass: needs its GetState */
            default: return E_UNSUPPORTED;
            }
        }
        for (int i = 0; i < f->n_msgs; i++) {
            const VsrMsg& m = f->msgs[i];
            if (m.type != VSR_MT_NEWSTATE) continue;
            const int v = m.view, p = O_::primary(v), src = m.src - 1, dest = m.dest - 1;
            if (v < 2 || dest == p || m.has_log != 1 || m.first_op != m.log_lo || m.op != m.log_lo + m.log_n - 1 || m.commit > V) return E_UNSUPPORTED;
            const int gi = (v - 2) * O + O_::oidx(p, dest);
            if (!VGET(L, GS_ST, w, gi) || (int)VGET(L, GS_T, w, gi) != m.first_op - 1 || (int)VGET(L, GS_DEST, w, gi) != src) return E_UNSUPPORTED;
            if (VGET(L, NS_ST, w, gi)) return E_SLOT_OCCUPIED;
            if (!log_to_ids(m.log, m.log_n, m.log_lo - 1, ids)) return E_UNSUPPORTED;
            VSET(L, NS_ST, w, gi, m.count ? ST_PENDING : ST_CONSUMED);
            set_log<L::NS_LOG_B>(w, gi, ids);
            VSET(L, NS_COMMIT, w, gi, m.commit);
        }
        /* Prepare: every destination must have been seen (BroadcastFunc sends to all at once) */
        for (int x = 0; x < V; x++) {
            if (!VGET(L, PR_VIEW, w, x)) continue;
            int seen = 0;
            const int p = O_::primary((int)VGET(L, PR_VIEW, w, x));
            for (int i = 0; i < f->n_msgs; i++)
                if (f->msgs[i].type == VSR_MT_PREPARE && f->msgs[i].entry.operation == x + 1 && f->msgs[i].dest - 1 != p) seen++;
            if (seen != O) return E_UNSUPPORTED;
            VSET(L, ACKED, w, x, 0);
        }
        /* aux */
        if (f->aux_svc > L::L) return E_UNSUPPORTED;
        VSET(L, AUX_SVC, w, 0, f->aux_svc);
        for (int x = 0; x < V; x++) {
            if ((f->acked[x] != 0) != (VGET(L, PR_VIEW, w, x) != 0)) return E_UNSUPPORTED; /* DOMAIN acked = requested values */
            VSET(L, ACKED, w, x, f->acked[x]);
        }
        /* replicas */
        auto entries_ok = [&](const VsrEntry* lg, int n) {
            for (int i = 0; i < n; i++) {
                const int x = lg[i].operation;
                if (x < 1 || x > V || !VGET(L, PR_VIEW, w, x - 1)) return false;
                if (VGET(L, PR_VIEW, w, x - 1) != lg[i].view || VGET(L, PR_REQ, w, x - 1) != lg[i].req) return false;
            }
            return true;
        };
        for (int i = 0; i < f->n_msgs; i++)
            if (f->msgs[i].has_log == 1 && !entries_ok(f->msgs[i].log, f->msgs[i].log_n)) return E_UNSUPPORTED;
        for (int r = 0; r < R; r++) {
            const VsrReplica& q = f->rep[r];
            if (q.status > 1 || q.view < 1 || q.view > K || q.commit > V || q.lnv > K || q.rec_number || q.n_rec) return E_UNSUPPORTED;
            if (q.op != q.log_n || !log_to_ids(q.log, q.log_n, 0, ids) || !entries_ok(q.log, q.log_n)) return E_UNSUPPORTED;
            VSET(L, STATUS, w, r, q.status);
            VSET(L, VIEWN, w, r, q.view);
            VSET(L, COMMIT, w, r, q.commit);
            VSET(L, LNV, w, r, q.lnv);
            VSET(L, SENT_DVC, w, r, q.sent_dvc ? 1 : 0);
            VSET(L, SENT_SV, w, r, q.sent_sv ? 1 : 0);
            set_log<L::LOG_B>(w, r, ids);
            for (int p = 0; p < R; p++) {
                if (q.peer_op[p] > V) return E_UNSUPPORTED;
                VSET(L, PEER, w, r * R + p, q.peer_op[p]);
            }
            if (q.client_table[0].req > V || q.client_table[0].op > V) return E_UNSUPPORTED;
            VSET(L, CT_REQ, w, r, q.client_table[0].req);
            VSET(L, CT_OP, w, r, q.client_table[0].op);
            VSET(L, CT_EXEC, w, r, q.client_table[0].executed ? 1 : 0);
            for (int i = 0; i < q.n_svc; i++) {
                const VsrMsg& m = q.svc_recv[i];
                if (m.type != VSR_MT_SVC || m.view != q.view || m.dest != r + 1 || m.src < 1 || m.src > R || m.src == r + 1) return E_UNSUPPORTED;
                VSET(L, SVC_MASK, w, r * R + (m.src - 1), 1);
            }
            for (int i = 0; i < q.n_dvc; i++) {
                const VsrMsg& m = q.dvc_recv[i];
                if (m.type != VSR_MT_DVC || m.view != q.view || m.dest != r + 1 || m.src < 1 || m.src > R || m.has_log != 1 || m.op != m.log_n) return E_UNSUPPORTED;
                if (!log_to_ids(m.log, m.log_n, 0, ids) || !entries_ok(m.log, m.log_n)) return E_UNSUPPORTED;
                if (VGET(L, DVC_MASK, w, r * R + (m.src - 1))) return E_SLOT_OCCUPIED;
                VSET(L, DVC_MASK, w, r * R + (m.src - 1), 1);
                if (m.src - 1 == r) {
                    set_log<L::SELF_LOG_B>(w, r, ids);
                    VSET(L, SELF_LNV, w, r, m.lnv);
                    VSET(L, SELF_COMMIT, w, r, m.commit);
                } else {
                    if (q.view < 2 || O_::primary(q.view) != r) return E_UNSUPPORTED;
                    const int di = (q.view - 2) * O + O_::oidx(r, m.src - 1);
                    if (!VGET(L, DVC_ST, w, di) || !same_log<L::DVC_LOG_B>(w, di, ids) || VGET(L, DVC_LNV, w, di) != m.lnv ||
                        VGET(L, DVC_COMMIT, w, di) != m.commit)
                        return E_UNSUPPORTED;
                }
            }
        }
        if (symmetry) return O_::canonicalize(w);
        return 0;
    }
};

} // namespace vsr
#endif
