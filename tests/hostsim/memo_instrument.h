// memo_instrument.h -- TEST INFRASTRUCTURE ONLY (tests/hostsim/): probes of the memoised QSS schedule for host-side analysis.
// Part 1 (before csrc/sto_qss_memo.cuh is included): counters, the event log and the probe macros the product header
// leaves empty.  Part 2 (STO_INSTRUMENT_PART2, after it): the forward-list batching PROTOTYPE, which needs the header's types.
#if !defined(STO_INSTRUMENT_PART2)
#include <stdint.h>
#include <vector>
#include "../../include/sto_b200.h"
namespace sto {
struct QssArgs;
struct MemoWork;
struct MemoCtx;
template <int G>
inline int memo_spawned_fwd_batch(const QssArgs& A, const MemoWork& W, const MemoCtx& C, const sto_vehicle_f64& V, int b,
                                  bool skip, int s, double lat0, int nlist, int64_t& steps, int& status);
static long long g_memo_evals[2] = {0, 0};
static long long g_memo_words[2] = {0, 0};
static long long g_att_hist[2][16] = {};
static long long g_sp_visits[2] = {0, 0}, g_sp_distinct[2] = {0, 0}, g_sp_on_live_orig[2] = {0, 0};
static long long g_sp_evals[2] = {0, 0}, g_sp_changed[2] = {0, 0}, g_sp_maxlist[2] = {0, 0};
static int g_fwd_batch = 0;                    // prototype switch: forward list by memo_spawned_fwd_batch
static long long g_fb[4] = {0, 0, 0, 0};       // its batches, committed members, cuts, list entries resolved
static int g_log_phase = 0, g_log_iter = 0, g_log_on = 0;   // event log: {iteration, sub-pass, sample p, outcome}
static std::vector<int> g_log;
}  // namespace sto
#define STO_PROBE_EVAL(fwd) ++g_memo_evals[(fwd) ? 1 : 0];
#define STO_PROBE_WORD(d, att) { ++g_memo_words[d]; if (att) { int c_ = popc64(att); ++g_att_hist[d][c_ > 15 ? 15 : c_]; } }
#define STO_PROBE_EVENT(x, kind) if (g_log_on) { g_log.push_back(g_log_iter); g_log.push_back(g_log_phase); g_log.push_back(x); g_log.push_back(kind); }
#define STO_PROBE_LIST_EVAL(d, n, batches) { g_sp_evals[d] += (n); g_sp_changed[d] += (batches); }
#define STO_PROBE_PHASE(k, iters) { g_log_phase = (k); g_log_iter = (iters); }
#define STO_PROBE_FWD_LIST(G, A, W, C, V, b, done, s, lat0, nF, steps, status, wF) \
    (g_fwd_batch && (G) > 1 && ((wF) = memo_spawned_fwd_batch<G>(A, W, C, V, b, done, s, lat0, nF, steps, status), true))
#else
namespace sto {
// PROTOTYPE (host analysis only, not compiled into the library): the forward re-spawned list with up to G evaluations at a
// time and NO static conflict rule.  A segment of the list is scanned until G entries need an
// evaluation (duplicates of a member - same sample - are not members); the members are evaluated against the state as
// it stands; the segment is then resolved in list order from the memo bits read at the START of the segment plus the
// members' outcomes, without looking at the planes again - which is what a device version can do 8 entries per round:
//   * an earlier committed member on the same sample p decides the entry: CONT memo -> kept, STOP memo -> dropped;
//   * an earlier committed member that WROTE the sample p (its q == p) cleared the entry's memo and changed its source:
//     the entry needs an evaluation on the new state -> the segment is CUT here (the next segment starts at this entry);
//   * otherwise the entry is what the start-of-segment memo bits say; a member then commits its speculative result (its
//     inputs are untouched: the only writer of p or q would have triggered one of the two rules above).
// tests/test_hostsim.py runs this against the one-at-a-time walk (bit-identical state, list, step count) and reports the
// evaluations per batch it achieves.
template <int G>
inline int memo_spawned_fwd_batch(const QssArgs& A, const MemoWork& W, const MemoCtx& C, const sto_vehicle_f64& V, int b,
                                  bool skip, int s, double lat0, int nlist, int64_t& steps, int& status) {
    const int N = A.N, ld = A.ld;
    const Ring cont = C.cont(1), stop = C.stop(1);
    int32_t* list = W.spF;
    if (skip) nlist = 0;
    int r = 0, w = 0;
    std::vector<int> cls;   // start-of-segment class per entry: 0 CONT, 1 STOP, 2 needs evaluation
    while (r < nlist) {
        int mi[G], mp[G], mq[G], cnt = 0;
        EvalRes mres[G];
        cls.clear();
        int e = r;
        while (e < nlist && cnt < G) {   // pass A: snapshot classification, member collection
            int p = list[at(e, ld, b)] + s;
            if (p >= N) p -= N;
            const int c = cont.test(p) ? 0 : (stop.test(p) ? 1 : 2);
            cls.push_back(c);
            bool dup = false;
            for (int k = 0; k < cnt; ++k) dup = dup || mp[k] == p;
            if (c == 2 && !dup) { mi[cnt] = e; mp[cnt] = p; mq[cnt] = (p + 1 == N) ? 0 : p + 1; ++cnt; }
            ++e;
        }
        const int seg_end = e;
        for (int k = 0; k < cnt; ++k) mres[k] = eval_pure(A, V, b, true, mp[k], mq[k], lat0);   // pass B: all at once
        ++g_fb[0];
        int ncommit = 0;                 // members committed so far in this segment (they are the first ncommit ones)
        int i = r;
        for (; i < seg_end; ++i) {       // pass C: resolution in list order, planes not consulted
            const int iv = list[at(i, ld, b)];
            int p = iv + s;
            if (p >= N) p -= N;
            int c = cls[i - r];
            bool cut = false;
            for (int k = 0; k < ncommit; ++k) {          // in commit order: later members override
                if (mres[k].kind == EV_WRITE && mq[k] == p) { c = 2; cut = true; }
                if (mp[k] == p) { c = (mres[k].kind == EV_STOP || mres[k].kind == EV_ZERO) ? 1 : 0; cut = false; }
            }
            if (cut) break;                              // source rewritten: evaluate in the next segment
            ++g_fb[3];
            ++steps;
            if (c == 0) { if (w != i) list[at(w, ld, b)] = iv; ++w; continue; }
            if (c == 1) continue;
            // needs an evaluation: it must be the next uncommitted member
            if (!(ncommit < cnt && mi[ncommit] == i)) { --steps; --g_fb[3]; break; }   // (a duplicate whose member stopped short)
            bool spawn = false, changed = false;
            const bool stopped = apply_res(A, C, b, true, mp[ncommit], mq[ncommit], mres[ncommit], status, spawn, changed);
            if (mres[ncommit].kind == EV_ZERO) { mres[ncommit].kind = EV_ZERO; }
            if (!stopped) { if (w != i) list[at(w, ld, b)] = iv; ++w; }
            ++ncommit;
            ++g_fb[1];
        }
        if (i < seg_end) ++g_fb[2];
        r = i;
    }
    return w;
}
}  // namespace sto
#endif
