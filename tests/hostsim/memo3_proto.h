// memo3_proto.h -- TEST / ANALYSIS INFRASTRUCTURE ONLY (host build).  Semantic model of the "merged planes" schedule:
// re-spawned rows live in the SAME per-direction front planes as the original rows (a re-spawned backward row has the
// virtual row of the front that spawned it), and the reference's processing order (original rows in row order, then
// re-spawned rows in creation order; simulator.py:149-365) is kept as ONE order bit per adjacent pair of virtual rows:
// a front only ever affects the front on the neighbouring sample (it writes that front's source sample), so the order
// between non-adjacent fronts is immaterial.  A sub-pass runs as Jacobi rounds over the fronts whose predecessor (the
// neighbour that is earlier in the reference's order) is resolved.  Must equal the oracle bit for bit.
#pragma once
#include <algorithm>
#include <vector>

namespace sto {

struct Memo3Stats {
    long long rounds[2] = {0, 0};        // Jacobi rounds with >= 1 evaluation, per direction (unbounded lanes)
    long long rounds_g8[2] = {0, 0};     // sum over rounds of ceil(evaluations / 8)
    long long evals[2] = {0, 0};
    long long dup_created = 0;           // new forward front landing on a live forward front of the same virtual row
    long long dup_diverged = 0;          // (not modelled) kept for the caller
    long long reborn = 0;
};
static Memo3Stats g_m3;

inline void qss_memo3_proto(const QssArgs& A, const sto_vehicle_f64& V, int b) {
    const int N = A.N, ld = A.ld;
    const double lat0 = max_lat_acc(V, 0.0);
    double* rec = A.rec + (size_t)b * N * 4;
    for (int i = 0; i < N; ++i) {
        const double Ri = A.R[at(i, ld, b)];
        rec[4 * (size_t)i + 0] = init_speed(lat0, Ri, gsb_at(A, i), V.max_speed);
        rec[4 * (size_t)i + 1] = 0.0;
        rec[4 * (size_t)i + 2] = A.dd[at(i, ld, b)];
        rec[4 * (size_t)i + 3] = Ri;
    }
    std::vector<char> live[2], cont[2], stop[2], ord[2];
    std::vector<int> mult(N, 1);
    for (int d = 0; d < 2; ++d) { live[d].assign(N, 1); cont[d].assign(N, 0); stop[d].assign(N, 0); }
    // ord[0][j]: pair (j-1, j) of backward fronts, 1 = the UPPER front (j) comes first in the reference's order (so j-1,
    //            whose source sample j writes, waits for it); originals: row j-1 before row j, except the seam (N-1, 0)
    // ord[1][i]: pair (i, i+1) of forward fronts, 1 = the LOWER front (i) comes first (i+1 waits); originals: all but the seam
    ord[0].assign(N, 0); ord[0][0] = 1;
    ord[1].assign(N, 1); ord[1][N - 1] = 0;
    int status = 0, s = 0, iters = 0;
    int64_t steps = 0;
    std::vector<int> reborn;
    std::vector<char> unres(N), startlive(N), isnew(N), isrb(N);
    struct Ev { int j, p, q; EvalRes r; };
    std::vector<Ev> evs;
    for (;;) {
        int nl = 0;
        for (int d = 0; d < 2; ++d) for (int i = 0; i < N; ++i) nl += live[d][i];
        if (nl == 0 || status != 0) break;
        reborn.clear();
        for (int d = 0; d < 2; ++d) {
            const bool fwd = d == 1;
            int nun = 0;
            for (int i = 0; i < N; ++i) {
                unres[i] = startlive[i] = live[d][i];
                if (live[d][i]) { steps += fwd ? mult[i] : 1; ++nun; }
            }
            while (nun > 0 && status == 0) {
                evs.clear();
                std::vector<int> ready;
                for (int j = 0; j < N; ++j) {
                    if (!unres[j]) continue;
                    // the neighbour that can affect front j: backward j+1 (it writes sample p_j), forward j-1
                    const int nb = fwd ? ((j == 0) ? N - 1 : j - 1) : ((j + 1 == N) ? 0 : j + 1);
                    const bool nb_first = fwd ? ord[1][nb] : ord[0][nb];   // pair (nb, j) forward / (j, nb) backward
                    if (unres[nb] && nb_first) continue;
                    ready.push_back(j);
                }
                if (ready.empty()) { status |= STO_CAND_NO_CONVERGENCE; break; }   // (a cycle of waits: must not happen)
                for (int j : ready) {
                    int p = fwd ? j + s : j - s;
                    if (p >= N) p -= N;
                    if (p < 0) p += N;
                    const int q = fwd ? ((p + 1 == N) ? 0 : p + 1) : ((p == 0) ? N - 1 : p - 1);
                    if (cont[d][p]) continue;
                    if (stop[d][p]) { live[d][j] = 0; continue; }
                    const double* rp = rec + 4 * (size_t)p;
                    const double* rq = rec + 4 * (size_t)q;
                    const double dd = fwd ? rp[2] : rq[2];
                    Ev e{j, p, q, eval_core(V, fwd, rp[0], rp[1], rq[0], rq[1], dd, rq[3], gsb_at(A, q), lat0)};
                    evs.push_back(e);
                }
                for (int j : ready) { unres[j] = 0; --nun; }
                if (!evs.empty()) {
                    ++g_m3.rounds[d];
                    g_m3.rounds_g8[d] += ((long long)evs.size() + 7) / 8;
                    g_m3.evals[d] += (long long)evs.size();
                }
                // two-phase commit (sto_qss_memo2.cuh): state + own-edge memo, then the invalidation minus the own edge
                for (const Ev& e : evs) {
                    const int k = e.r.kind;
                    if (k == EV_WRITE || k == EV_SPAWN) { rec[4 * (size_t)e.q] = e.r.v_new; rec[4 * (size_t)e.q + 1] = e.r.a_new; }
                    if (k == EV_WRITE || k == EV_KEEP) { cont[d][e.p] = 1; stop[d][e.p] = 0; }
                    else if (k == EV_STOP) { stop[d][e.p] = 1; cont[d][e.p] = 0; }
                    else if (k == EV_SPAWN) { cont[d][e.p] = 0; stop[d][e.p] = 0; }
                    else if (k == EV_ZERO) status |= STO_CAND_ZERO_SPEED;
                    if (k == EV_STOP || k == EV_SPAWN || k == EV_RESPAWN || k == EV_ZERO) live[d][e.j] = 0;
                    if (k == EV_SPAWN || k == EV_RESPAWN) reborn.push_back(e.j);
                }
                for (const Ev& e : evs) {
                    const int k = e.r.kind;
                    if (k != EV_WRITE && k != EV_SPAWN) continue;
                    const int q = e.q, qn = (q + 1 == N) ? 0 : q + 1, qp = (q == 0) ? N - 1 : q - 1;
                    if (fwd) {   // own edge p -> q is bit p = qp of the forward planes
                        cont[1][q] = 0; stop[1][q] = 0;
                        cont[0][q] = 0; stop[0][q] = 0; cont[0][qn] = 0; stop[0][qn] = 0;
                    } else {     // own edge p -> q is bit p = qn of the backward planes
                        cont[0][q] = 0; stop[0][q] = 0;
                        cont[1][q] = 0; stop[1][q] = 0; cont[1][qp] = 0; stop[1][qp] = 0;
                    }
                }
            }
        }
        if (status != 0) break;
        // ---- fold: the rows re-spawned in this iteration's backward sub-pass first act in the next iteration
        g_m3.reborn += (long long)reborn.size();
        std::fill(isrb.begin(), isrb.end(), 0);
        std::fill(isnew.begin(), isnew.end(), 0);
        for (int j : reborn) isrb[j] = 1;
        const std::vector<char> ord0_old = ord[0];
        for (int j : reborn) {
            const int jm = (j == 0) ? N - 1 : j - 1, jp = (j + 1 == N) ? 0 : j + 1;
            live[0][j] = 1;
            if (!isrb[jm]) ord[0][j] = 0;    // pair (j-1, j): j is the newest -> j-1 first
            if (!isrb[jp]) ord[0][jp] = 1;   // pair (j, j+1): j+1 is older -> upper first
        }
        std::vector<int> newf;
        for (int j : reborn) {
            int ivf = j - 2 * (s + 1);
            while (ivf < 0) ivf += N;
            if (live[1][ivf]) {
                ++g_m3.dup_created; ++mult[ivf];
                const int im = (ivf == 0) ? N - 1 : ivf - 1;
                if (live[1][im] && ord[1][im] == 0) ++g_m3.dup_diverged;   // a front on the sample below sits between the two in order
                continue;
            }
            live[1][ivf] = 1;
            mult[ivf] = 1;
            isnew[ivf] = 1;
            newf.push_back(ivf);
        }
        for (int j : reborn) {
            int ivf = j - 2 * (s + 1);
            while (ivf < 0) ivf += N;
            if (!isnew[ivf]) continue;
            const int im = (ivf == 0) ? N - 1 : ivf - 1, ip = (ivf + 1 == N) ? 0 : ivf + 1;
            const int jm = (j == 0) ? N - 1 : j - 1;
            // pair (ivf-1, ivf): the lower one is older unless it is new too (then: the order their backward fronts had)
            if (!isnew[im]) ord[1][im] = 1;
            else ord[1][im] = isrb[jm] ? !ord0_old[j] : 1;
            if (!isnew[ip]) ord[1][ivf] = 0;   // pair (ivf, ivf+1): the upper one is older
        }
        s = (s + 1 == N) ? 0 : s + 1;
        ++iters;
        if (iters > 64 * N + 1024) status |= STO_CAND_NO_CONVERGENCE;
    }
    qss_finish(A, StateRec{rec}, true, b, status, steps, iters);
}

// ---------------------------------------------------------------------------------------------------------------------
// Prototype 2 (analysis / test only): the re-spawned LISTS walked out of order.  A sub-pass over a list is cut into
// chunks of <= QCAP "open" entries (Q): entries that need an evaluation now, or whose memo an earlier open entry may
// still change (it sits on the sample that entry writes).  Every other entry is settled while scanning (kept / dropped),
// which is order-independent.  Inside a chunk an entry is processed (memo test, evaluation, commit) as soon as no EARLIER
// open entry within one sample of it is unfinished; entries processed in the same round are therefore >= 2 samples
// apart and commute, entries closer than that keep the list order: the result equals the sequential walk bit for bit.
// Lane k mod G owns open entry k and evaluates at most one per round.  Rows spawned by a chunk are appended in list order.
struct MemoQStats { long long eval_rounds[2] = {0, 0}, idle_rounds[2] = {0, 0}, evals[2] = {0, 0}, chunks[2] = {0, 0},
                    qsum[2] = {0, 0}, visits[2] = {0, 0}, walks[2] = {0, 0}; };
static MemoQStats g_mq;
static int g_mq_cap = 32, g_mq_lanes = 8, g_mq_war = 1, g_mq_fpar = 0, g_mq_wr = 0;

template <bool FWD>
inline int memo_spawned_rows_q(const QssArgs& A, const MemoWork& W, const MemoCtx& C, const sto_vehicle_f64& V, int b,
                               bool skip, int s, double lat0, int nlist, int nB, int& nnew, int64_t& steps, int& status,
                               std::vector<char>& PB, int& ndead) {
    const int N = A.N, ld = A.ld, d = FWD ? 1 : 0, QCAP = g_mq_cap, G = g_mq_lanes;
    const Ring cont = C.cont(d), stop = C.stop(d);
    int32_t* list = FWD ? W.spF : W.spB;
    if (skip) nlist = 0;
    if (nlist) ++g_mq.walks[d];
    int r = 0, w = 0;
    // g_mq_wr: open entries = entries on WORK-RUN rows (rows of live entries from the first row of a run of occupied adjacent
    // rows whose edge memo is not CONT, in the direction of travel); everything else is kept without a memo test
    std::vector<char> WR;
    if (g_mq_wr && nlist) {
        std::vector<char> occ(N, 0);
        WR.assign(N, 0);
        for (int k = 0; k < nlist; ++k) { const int iv = list[at(k, ld, b)]; if (iv >= 0) occ[iv] = 1; }
        for (int iv0 = 0; iv0 < N; ++iv0) {
            if (!occ[iv0] || WR[iv0]) continue;
            int p0 = FWD ? iv0 + s : iv0 - s;
            if (p0 >= N) p0 -= N;
            if (p0 < 0) p0 += N;
            if (cont.test(p0)) continue;
            int guard = 0;
            for (int iv = iv0; occ[iv] && !WR[iv] && guard < N; iv = FWD ? ((iv + 1 == N) ? 0 : iv + 1) : ((iv == 0) ? N - 1 : iv - 1), ++guard) WR[iv] = 1;
        }
    }
    int qp[64], qslot[64], qiv[64];
    unsigned long long dep[64], war[64];
    int dead = 0;
    while (r < nlist && status == 0) {
        int nq = 0;
        // the device scans G entries per step and closes the chunk when a whole step may no longer fit; an open entry
        // depends on the LATEST earlier open entry on its own sample / on the sample whose entry writes its source (dep),
        // and must not commit before the latest earlier open entry whose source it writes (war)
        while (r < nlist && nq <= QCAP - G) {
            for (int t = 0; t < G && r < nlist; ++t) {
                const int iv = list[at(r, ld, b)];
                ++r;
                if (iv < 0) continue;   // tombstone of an earlier walk
                ++steps;
                ++g_mq.visits[d];
                int p = FWD ? iv + s : iv - s;
                if (p >= N) p -= N;
                if (p < 0) p += N;
                const bool c0 = cont.test(p), s0 = stop.test(p);
                if (g_mq_wr) {
                    if (!WR[iv]) { list[at(w, ld, b)] = iv; ++w; continue; }
                } else
                if ((c0 || s0) && !PB[p]) { if (c0) { list[at(w, ld, b)] = iv; ++w; } continue; }
                list[at(w, ld, b)] = iv;
                qslot[nq] = w; ++w;
                qp[nq] = p; qiv[nq] = iv;
                unsigned long long m = 0, mw = 0;
                const int praw = FWD ? ((p == 0) ? N - 1 : p - 1) : ((p + 1 == N) ? 0 : p + 1);   // an earlier entry there writes my source
                const int pwar = FWD ? ((p + 1 == N) ? 0 : p + 1) : ((p == 0) ? N - 1 : p - 1);   // I write the source of an earlier entry there
                int lr = -1, ls = -1, lw = -1;
                for (int i = 0; i < nq; ++i) {
                    if (qp[i] == p) ls = i;
                    else if (qp[i] == praw) lr = i;
                    else if (qp[i] == pwar) lw = i;
                }
                if (lr >= 0) m |= 1ull << lr;
                if (ls >= 0) m |= 1ull << ls;
                if (lw >= 0) { if (g_mq_war) mw |= 1ull << lw; else m |= 1ull << lw; }
                dep[nq] = m;
                war[nq] = mw;
                const int aff = FWD ? ((p + 1 == N) ? 0 : p + 1) : ((p == 0) ? N - 1 : p - 1);   // the sample this entry may write
                PB[aff] = 1;
                ++nq;
            }
        }
        if (!nq) break;
        ++g_mq.chunks[d];
        g_mq.qsum[d] += nq;
        unsigned long long unfinished = (nq >= 64) ? ~0ull : ((1ull << nq) - 1ull), spawn_mask = 0;
        while (unfinished && status == 0) {
            const unsigned long long snap = unfinished;
            int members[64], nm = 0;
            unsigned long long processed = 0;
            for (int g = 0; g < G; ++g)
                for (int k = g; k < nq; k += G) {
                    if (!((unfinished >> k) & 1ull) || (dep[k] & snap)) continue;
                    const int p = qp[k];
                    if (cont.test(p)) { unfinished &= ~(1ull << k); processed |= 1ull << k; continue; }
                    if (stop.test(p)) { list[at(qslot[k], ld, b)] = -1; ++dead; unfinished &= ~(1ull << k); processed |= 1ull << k; continue; }
                    members[nm++] = k;
                    processed |= 1ull << k;
                    break;
                }
            for (bool again = true; again;) {   // a member that would overwrite the source of an earlier entry not processed in this round waits
                again = false;
                for (int m = 0; m < nm; ++m) {
                    const int k = members[m];
                    if (war[k] & snap & ~processed) {
                        processed &= ~(1ull << k);
                        members[m] = members[--nm];
                        again = true;
                        break;
                    }
                }
            }
            std::sort(members, members + nm);   // commits in list order (the device commits in two phases instead)
            EvalRes res[64];
            for (int m = 0; m < nm; ++m) {
                const int p = qp[members[m]];
                const int q = FWD ? ((p + 1 == N) ? 0 : p + 1) : ((p == 0) ? N - 1 : p - 1);
                res[m] = eval_pure(A, V, b, FWD, p, q, lat0);
            }
            for (int m = 0; m < nm; ++m) {
                const int k = members[m], p = qp[k];
                const int q = FWD ? ((p + 1 == N) ? 0 : p + 1) : ((p == 0) ? N - 1 : p - 1);
                bool spawn, changed;
                const bool stopped = apply_res(A, C, b, FWD, p, q, res[m], status, spawn, changed);
                if (spawn) spawn_mask |= 1ull << k;
                if (stopped) { list[at(qslot[k], ld, b)] = -1; ++dead; }
                unfinished &= ~(1ull << k);
            }
            if (nm) { ++g_mq.eval_rounds[d]; g_mq.evals[d] += nm; } else ++g_mq.idle_rounds[d];
        }
        for (int k = 0; k < nq; ++k) {
            const int p = qp[k];
            const int aff = FWD ? ((p + 1 == N) ? 0 : p + 1) : ((p == 0) ? N - 1 : p - 1);
            PB[aff] = 0;
            if ((spawn_mask >> k) & 1ull) { memo_spawn(A, W, b, aff, s, nB, nnew, status); ++nnew; }
        }
    }
    ndead = dead;
    return w;
}


// ---------------------------------------------------------------------------------------------------------------------
// Prototype 3 (analysis / test only): the forward sub-pass over the ORIGINAL rows with independent live runs in parallel.
// The sub-pass is a sequential sweep in row order, but a row only affects the NEXT row, and only if that row is live: a
// dirty front is READY when no dirty front sits below it in its own run of consecutive live rows.  Per round the G lowest
// words with dirty fronts offer their lowest dirty front if it is ready; all are evaluated on the state as it stands and
// committed together; a changing step makes the next live row dirty.
struct MemoFStats { long long rounds = 0, evals = 0, subpasses = 0; };
static MemoFStats g_mf;
static int g_mf_allruns = 0;

inline void memo_forward_rows_par(const QssArgs& A, const MemoWork& W, const MemoCtx& C, const sto_vehicle_f64& V, int b,
                                  bool skip, int s, double lat0, int& nlive, u64& words, int64_t& steps, int& status) {
    const int N = A.N, NW = W.W, G = g_mq_lanes;
    if (skip) return;
    const Ring live = C.live(1), cont = C.cont(1), stop = C.stop(1);
    std::vector<u64> att(NW, 0), Lw(NW, 0);
    u64 todo = 0;
    for (int w = 0; w < NW; ++w) {
        if (!((words >> w) & 1ull)) continue;
        const u64 L = live.word(w);
        Lw[w] = L;
        if (!L) { words &= ~(1ull << w); continue; }
        steps += popc64(L);
        int start = 64 * w + s;
        if (start >= N) start -= N;
        att[w] = L & ~cont.window(start);
        if (att[w]) todo |= 1ull << w;
    }
    ++g_mf.subpasses;
    while (todo && status == 0) {
        struct Cd { int w, t, p; };
        Cd cd[64];
        int nc = 0;
        u64 tw = todo;
        if (g_mf_allruns) {
            // every run of consecutive live rows offers its lowest dirty front (several per word), lowest rows first
            bool below_dirty = false;   // a dirty front sits below in the current run
            for (int w = 0; w < NW && nc < G; ++w) {
                const int nbits = (N - 64 * w >= 64) ? 64 : N - 64 * w;
                if (!att[w] && Lw[w] == ((nbits == 64) ? ~0ull : ((1ull << nbits) - 1ull))) continue;   // fully live, clean: run goes on
                for (int t = 0; t < nbits && nc < G; ++t) {
                    const u64 bit = 1ull << t;
                    if (!(Lw[w] & bit)) { below_dirty = false; continue; }
                    if (!(att[w] & bit)) continue;
                    if (below_dirty) continue;
                    int p = 64 * w + t + s;
                    if (p >= N) p -= N;
                    if (stop.test(p)) { Lw[w] &= ~bit; live.set_word(w, Lw[w]); --nlive; att[w] &= ~bit; below_dirty = false; continue; }
                    cd[nc++] = Cd{w, t, p};
                    below_dirty = true;
                }
                if (!att[w]) todo &= ~(1ull << w);
            }
        } else
        for (int g = 0; g < G && tw; ++g) {
            const int w = ctz64(tw);
            tw &= tw - 1ull;
            for (;;) {
                if (!att[w]) { todo &= ~(1ull << w); break; }
                const int t = ctz64(att[w]);
                const u64 bit = 1ull << t, low = bit - 1ull;
                bool blocked = false;
                if ((Lw[w] & low) == low) {           // every row below in this word is live: the run goes on below
                    for (int v = w - 1; v >= 0; --v) {
                        const int nbits = (N - 64 * v >= 64) ? 64 : N - 64 * v;
                        const u64 fullm = (nbits == 64) ? ~0ull : ((1ull << nbits) - 1ull);
                        const u64 dead = ~Lw[v] & fullm;
                        if (!dead) { if (att[v]) { blocked = true; break; } continue; }
                        const int z = 63 - __builtin_clzll(dead);            // highest dead row of the word
                        if (z < 63 && (att[v] >> (z + 1))) blocked = true;
                        break;
                    }
                }
                if (blocked) break;
                int p = 64 * w + t + s;
                if (p >= N) p -= N;
                if (stop.test(p)) { Lw[w] &= ~bit; live.set_word(w, Lw[w]); --nlive; att[w] &= ~bit; continue; }
                cd[nc++] = Cd{w, t, p};
                break;
            }
        }
        if (!nc) continue;
        ++g_mf.rounds;
        g_mf.evals += nc;
        EvalRes res[64];
        for (int m = 0; m < nc; ++m) {
            const int p = cd[m].p, q = (p + 1 == N) ? 0 : p + 1;
            res[m] = eval_pure(A, V, b, true, p, q, lat0);
        }
        for (int m = 0; m < nc; ++m) {
            const int p = cd[m].p, q = (p + 1 == N) ? 0 : p + 1, w = cd[m].w, t = cd[m].t;
            const u64 bit = 1ull << t;
            bool spawn, changed;
            const bool stopped = apply_res(A, C, b, true, p, q, res[m], status, spawn, changed);
            att[w] &= ~bit;
            if (stopped) { Lw[w] &= ~bit; live.set_word(w, Lw[w]); --nlive; }
            if (changed) {
                if (t < 63) att[w] |= Lw[w] & (bit << 1);
                else if (w + 1 < NW && (Lw[w + 1] & 1ull)) { att[w + 1] |= 1ull; todo |= 1ull << (w + 1); }
            }
            if (!att[w]) todo &= ~(1ull << w); else todo |= 1ull << w;
        }
    }
    for (int w = 0; w < NW; ++w) if (!Lw[w]) words &= ~(1ull << w);
}


// ---------------------------------------------------------------------------------------------------------------------
// Prototype 4 (analysis / test only): a re-spawned list processed by RUNS OF ROWS.  The live entries of a list occupy
// virtual rows; an entry only affects entries on the next row (forward) / the previous row (backward).  Rows whose edge
// memo is not CONT are open; a maximal run of occupied adjacent rows is cut at its first open row (in the direction of
// travel) and everything from there on is a WORK RUN.  Work runs are independent of each other; the entries of one work
// run are processed one at a time in list order (which IS the reference order restricted to entries that can interact).
struct MemoRStats { long long walks = 0, runs = 0, evals = 0, rounds_est = 0, longest = 0, entries_in_runs = 0, visits = 0; };
static MemoRStats g_mr[2];
static int g_mq_rows = 0;

template <bool FWD>
inline int memo_spawned_rows_runs(const QssArgs& A, const MemoWork& W, const MemoCtx& C, const sto_vehicle_f64& V, int b,
                                  bool skip, int s, double lat0, int nlist, int nB, int& nnew, int64_t& steps, int& status,
                                  int& ndead) {
    const int N = A.N, ld = A.ld, d = FWD ? 1 : 0, G = g_mq_lanes;
    const Ring cont = C.cont(d), stop = C.stop(d);
    int32_t* list = FWD ? W.spF : W.spB;
    if (skip) nlist = 0;
    if (!nlist) return 0;
    ++g_mr[d].walks;
    // live entries by virtual row
    std::vector<std::vector<int>> at_row(N);     // list indices, ascending
    std::vector<char> occ(N, 0), openr(N, 0), inrun(N, 0);
    for (int r = 0; r < nlist; ++r) {
        const int iv = list[at(r, ld, b)];
        if (iv < 0) continue;
        ++steps;
        ++g_mr[d].visits;
        at_row[iv].push_back(r);
        occ[iv] = 1;
    }
    for (int iv = 0; iv < N; ++iv) {
        if (!occ[iv]) continue;
        int p = FWD ? iv + s : iv - s;
        if (p >= N) p -= N;
        if (p < 0) p += N;
        openr[iv] = !cont.test(p);
    }
    // work runs: start at an open row whose predecessor row (against the direction of travel) is not in a run
    auto nextrow = [&](int iv) { return FWD ? ((iv + 1 == N) ? 0 : iv + 1) : ((iv == 0) ? N - 1 : iv - 1); };
    auto prevrow = [&](int iv) { return FWD ? ((iv == 0) ? N - 1 : iv - 1) : ((iv + 1 == N) ? 0 : iv + 1); };
    std::vector<std::vector<int>> runs;
    for (int iv0 = 0; iv0 < N; ++iv0) {
        if (!openr[iv0] || inrun[iv0]) continue;
        // walk back: is there an open occupied row connected behind me?  then I belong to its run (found from there)
        int st = iv0, guard = 0;
        for (int pv = prevrow(st); occ[pv] && guard < N; pv = prevrow(pv), ++guard) if (openr[pv]) st = pv;
        if (inrun[st]) continue;
        std::vector<int> rows;
        guard = 0;
        for (int iv = st; occ[iv] && !inrun[iv] && guard < N; iv = nextrow(iv), ++guard) { rows.push_back(iv); inrun[iv] = 1; }
        runs.push_back(rows);
    }
    std::vector<std::pair<int, int>> spawns;   // (list index, q)
    long long tot = 0, longest = 0;
    for (const auto& rows : runs) {
        std::vector<int> ents;
        for (int iv : rows) for (int r : at_row[iv]) ents.push_back(r);
        std::sort(ents.begin(), ents.end());
        g_mr[d].entries_in_runs += (long long)ents.size();
        long long ne = 0;
        for (int r : ents) {
            if (status != 0) break;
            const int iv = list[at(r, ld, b)];
            int p = FWD ? iv + s : iv - s;
            if (p >= N) p -= N;
            if (p < 0) p += N;
            if (cont.test(p)) continue;
            if (stop.test(p)) { list[at(r, ld, b)] = -1; ++ndead; continue; }
            const int q = FWD ? ((p + 1 == N) ? 0 : p + 1) : ((p == 0) ? N - 1 : p - 1);
            bool spawn, changed;
            const bool stopped = memo_step(A, C, V, b, FWD, p, q, lat0, status, spawn, changed);
            ++ne;
            if (spawn) spawns.push_back(std::make_pair(r, q));
            if (stopped) { list[at(r, ld, b)] = -1; ++ndead; }
        }
        tot += ne;
        if (ne > longest) longest = ne;
    }
    g_mr[d].runs += (long long)runs.size();
    g_mr[d].evals += tot;
    g_mr[d].rounds_est += std::max(longest, (tot + G - 1) / G);
    g_mr[d].longest = std::max(g_mr[d].longest, longest);
    std::sort(spawns.begin(), spawns.end());     // new rows in list order of their spawners
    for (const auto& sp : spawns) { memo_spawn(A, W, b, sp.second, s, nB, nnew, status); ++nnew; }
    return nlist;   // (no compaction: tombstones stay; the model's caller compacts)
}

inline void qss_memo_q_proto(const QssArgs& A, const MemoWork& W, const MemoCtx& C, const sto_vehicle_f64& V, int b) {
    const int N = A.N, ld = A.ld;
    const double lat0 = max_lat_acc(V, 0.0);
    int status = 0;
    double* rec = A.rec + (size_t)b * N * 4;
    for (int i = 0; i < N; ++i) {
        const double Ri = A.R[at(i, ld, b)];
        rec[4 * (size_t)i + 0] = init_speed(lat0, Ri, gsb_at(A, i), V.max_speed);
        rec[4 * (size_t)i + 1] = 0.0;
        rec[4 * (size_t)i + 2] = A.dd[at(i, ld, b)];
        rec[4 * (size_t)i + 3] = Ri;
    }
    for (int w = 0; w < W.W; ++w) {
        const int nb = N - 64 * w;
        const u64 ones = (nb >= 64) ? ~0ull : ((1ull << nb) - 1ull);
        for (int d = 0; d < 2; ++d) { C.live(d).set_word(w, ones); C.cont(d).set_word(w, 0); C.stop(d).set_word(w, 0); }
    }
    std::vector<char> PB(N, 0);
    int nliveB = N, nliveF = N;
    u64 wordsB = (W.W >= 64) ? ~0ull : ((1ull << W.W) - 1ull), wordsF = wordsB;
    int nB = 0, nF = 0, s = 0, iters = 0, liveLB = 0, liveLF = 0;   // nB, nF: list lengths (tombstones included)
    int64_t steps = 0;
    for (;;) {
        const bool done = (nliveB == 0 && nliveF == 0 && liveLB == 0 && liveLF == 0) || status != 0;
        if (done) break;
        int nnew = 0, none = 0;
        memo_original_rows<false>(A, W, C, V, b, nliveB == 0, s, lat0, nB, nnew, nliveB, wordsB, steps, status);
        int wB = 0, wF = 0, deadB = 0, deadF = 0;
        auto compact = [&](int32_t* list, int n) { int w2 = 0; for (int r2 = 0; r2 < n; ++r2) { const int iv = list[at(r2, ld, b)]; if (iv >= 0) list[at(w2++, ld, b)] = iv; } return w2; };
        if (nB > 0 && g_mq_rows) {
            // (the spawned rows sit behind the list: keep them there while compacting)
            std::vector<int32_t> tail;
            memo_spawned_rows_runs<false>(A, W, C, V, b, false, s, lat0, nB, nB, nnew, steps, status, deadB);
            for (int j = 0; j < nnew; ++j) tail.push_back(W.spB[at(nB + j, ld, b)]);
            wB = compact(W.spB, nB); deadB = 0;
            for (int j = 0; j < nnew; ++j) W.spB[at(nB + j, ld, b)] = tail[j];
        } else
        if (nB > 0) wB = memo_spawned_rows_q<false>(A, W, C, V, b, false, s, lat0, nB, nB, nnew, steps, status, PB, deadB);
        if (iters == 0 && nliveF == N) memo_forward_sweep0(A, W, C, V, b, lat0, nliveF, steps, status);
        if (g_mq_fpar) memo_forward_rows_par(A, W, C, V, b, nliveF == 0, s, lat0, nliveF, wordsF, steps, status);
        else memo_original_rows<true>(A, W, C, V, b, nliveF == 0, s, lat0, nB, none, nliveF, wordsF, steps, status);
        if (nF > 0 && g_mq_rows) {
            memo_spawned_rows_runs<true>(A, W, C, V, b, false, s, lat0, nF, nB, none, steps, status, deadF);
            wF = compact(W.spF, nF); deadF = 0;
        } else
        if (nF > 0) wF = memo_spawned_rows_q<true>(A, W, C, V, b, false, s, lat0, nF, nB, none, steps, status, PB, deadF);
        if (wF + nnew > A.cap) { status |= STO_CAND_ROW_OVERFLOW; nnew = 0; }
        for (int j = 0; j < nnew; ++j) {
            const int ivb = W.spB[at(nB + j, ld, b)];
            int ivf = ivb - 2 * (s + 1);
            if (ivf < 0) ivf += N;
            if (ivf < 0) ivf += N;
            W.spB[at(wB + j, ld, b)] = ivb;
            W.spF[at(wF + j, ld, b)] = ivf;
        }
        nB = wB + nnew;
        nF = wF + nnew;
        liveLB = nB - deadB;
        liveLF = nF - deadF;
        s = (s + 1 == N) ? 0 : s + 1;
        ++iters;
        if (iters > 64 * N + 1024) status |= STO_CAND_NO_CONVERGENCE;
    }
    qss_finish(A, StateRec{rec}, true, b, status, steps, iters);
}

}  // namespace sto
