// hostsim.cpp -- TEST INFRASTRUCTURE ONLY.  Compiles the per-candidate device functions of
// spline_trajectory_optimization_b200/csrc/*.cuh for the host (g++, -ffp-contract=off) so the schedule logic
// (row tables, ring windows, memoisation) can be unit-tested in the GPU-less authoring container.  Each
// candidate runs as a "warp of one".  The product never builds, loads or falls back to this file.
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "memo_instrument.h"   // probes + counters: part 1, before the product headers
#include "../../spline_trajectory_optimization_b200/csrc/sto_common.cuh"
#include "../../spline_trajectory_optimization_b200/csrc/sto_eval.cuh"
#include "../../spline_trajectory_optimization_b200/csrc/sto_fit.cuh"
#include "../../spline_trajectory_optimization_b200/csrc/sto_fast.cuh"
#include "../../spline_trajectory_optimization_b200/csrc/sto_fit_lsq.cuh"
#include "../../spline_trajectory_optimization_b200/csrc/sto_qss.cuh"
#include "../../spline_trajectory_optimization_b200/csrc/sto_qss_memo.cuh"
#define STO_INSTRUMENT_PART2 1
#include "memo_instrument.h"   // part 2: the forward-list batching prototype (needs the product header's types)
#include "memo3_proto.h"       // semantic model of the merged-planes schedule (analysis / test only)

extern "C" {

int hostsim_fit(const double* cenx, const double* ceny, const double* nrmx, const double* nrmy, const double* off,
                const double* px, const double* py, int M, int B, int ld, double* u, double* cx, double* cy,
                int32_t* status, int split) {
    std::vector<double> w((size_t)5 * M * ld);
    sto::FitArgs A{};
    A.cenx = cenx; A.ceny = ceny; A.nrmx = nrmx; A.nrmy = nrmy; A.off = off; A.px = px; A.py = py;
    A.M = M; A.B = B; A.ld = ld; A.u = u; A.cx = cx; A.cy = cy; A.status = status;
    A.cp = w.data(); A.zx = A.cp + (size_t)M * ld; A.zy = A.zx + (size_t)M * ld; A.zz = A.zy + (size_t)M * ld;
    A.ze = A.zz + (size_t)M * ld;
    // split > 0: Thomas; -32 <= split < 0: partitioned solve over |split| lanes; split <= -100: FITPACK's Givens sweep, |split| - 100 lanes
    for (int b = 0; b < B; ++b) {
        if (split <= -100) sto::fit_candidate(A, b, -split - 100 > 0 ? -split - 100 : 1, false, true);
        else sto::fit_candidate(A, b, split < 0 ? -split : split, split < 0);
    }
    return 0;
}

// least-squares periodic fit on fixed knots (sto_fit_lsq.cuh); px, py [M][ld]; outputs u [M+1][ld], cx, cy [nt-k-1][ld]
int hostsim_fit_lsq(const double* px, const double* py, int M, int B, int ld, const double* t, int nt, int k, double* u,
                    double* cx, double* cy, int32_t* status) {
    if (!sto::lsq_sizes_ok(M, nt, k)) return -1;
    std::vector<double> w(sto::lsq_work_doubles(nt, k) * (size_t)ld);
    const size_t g = (size_t)(nt - 2 * k - 1);
    sto::LsqArgs A{};
    A.F.px = px; A.F.py = py; A.F.M = M; A.F.B = B; A.F.ld = ld; A.F.u = u; A.F.status = status;
    A.t = t; A.nt = nt; A.k = k; A.cx = cx; A.cy = cy;
    A.band = w.data();
    A.rx = A.band + g * (size_t)(k + 1) * ld;
    A.ry = A.rx + g * ld;
    A.Y = A.ry + g * ld;
    for (int b = 0; b < B; ++b) sto::lsq_candidate_k(A, b);
    return 0;
}

int hostsim_eval(const double* u, const double* cx, const double* cy, int M, const double* ts, int N, int B, int ld,
                 double* x, double* y, double* yaw, double* radius, double* chord_qss, double* chord_norm) {
    sto::EvalArgs A{u, cx, cy, ts, M, N, B, ld, x, y, yaw, radius, chord_qss, chord_norm};
    for (int b = 0; b < B; ++b) sto::eval_candidate(A, b);
    return 0;
}

// the same with every candidate's sample range split into `split` slices (what the GPU's lane split does)
int hostsim_eval_split(const double* u, const double* cx, const double* cy, int M, const double* ts, int N, int B, int ld,
                       int split, double* x, double* y, double* yaw, double* radius, double* chord_qss,
                       double* chord_norm) {
    sto::EvalArgs A{u, cx, cy, ts, M, N, B, ld, x, y, yaw, radius, chord_qss, chord_norm};
    const int per = (N + split - 1) / split;
    for (int b = 0; b < B; ++b)
        for (int g = split - 1; g >= 0; --g) {   // any order: slices are independent
            const int j0 = g * per, j1 = (j0 + per < N) ? j0 + per : N;
            sto::eval_range(A, b, j0, j1);
        }
    return 0;
}

int hostsim_arc_sections(const double* t, int nt, const double* cx, const double* cy, int k, const double* ts, int N,
                         double* sec) {
    sto::ArcArgs A{t, cx, cy, ts, nt, k, N, sec};
    for (int i = 0; i < N; ++i) sto::arc_section(A, i);
    return 0;
}

int hostsim_eval_spline_batch(const double* t, int nt, int k, const double* cx, const double* cy, const double* ts, int N,
                              int B, int ld, double* x, double* y, double* yaw, double* radius, double* chord_qss,
                              double* chord_norm) {
    sto::SplineBatchArgs A{t, nt, k, cx, cy, ts, N, B, ld, x, y, yaw, radius, chord_qss, chord_norm};
    for (int j = 0; j < N; ++j)
        for (int b = 0; b < B; ++b) sto::eval_spline_batch_sample(A, j, b);
    return 0;
}

int hostsim_eval_spline(const double* t, int nt, const double* cx, const double* cy, int k, const double* ts, int N,
                        double* x, double* y, double* yaw, double* radius) {
    sto::SplineEvalArgs A{t, cx, cy, ts, nt, k, N, x, y, yaw, radius};
    for (int j = 0; j < N; ++j) sto::eval_spline_sample(A, j);
    return 0;
}

// x, y, radius: [N][ld]; outputs v, a, lat, tseg [N][ld], owner [N][ld] (plain only, may be NULL), lap[B],
// summary[8][ld], status[B]
int hostsim_qss(int impl, const double* x, const double* y, const double* radius, const double* sinb, int N, int B,
                int ld, const sto_vehicle_f64* V, double* v, double* a, double* lat, double* tseg, int32_t* owner,
                double* lap, double* summary, int32_t* status) {
    const int cap = 2 * N + 64;
    std::vector<double> dd((size_t)N * ld), df((size_t)N * ld);
    for (int i = 0; i < N; ++i)
        for (int b = 0; b < B; ++b) {
            int n = (i + 1 == N) ? 0 : i + 1;
            double x0 = x[sto::at(i, ld, b)], y0 = y[sto::at(i, ld, b)], x1 = x[sto::at(n, ld, b)],
                   y1 = y[sto::at(n, ld, b)];
            dd[sto::at(i, ld, b)] = sto::chord_qss(x0, y0, x1, y1);
            df[sto::at(i, ld, b)] = sto::chord_norm(x0, y0, x1, y1);
        }
    std::vector<uint8_t> rowflag((size_t)N * ld), spf((size_t)cap * ld);
    std::vector<int32_t> e1((size_t)cap * ld), e2((size_t)cap * ld), e3((size_t)cap * ld);
    sto::QssArgs A{};
    A.dd = dd.data(); A.df = df.data(); A.R = radius; A.sinb = sinb; A.N = N; A.B = B; A.ld = ld; A.cap = cap;
    A.v = v; A.a = a; A.rowflag = rowflag.data(); A.sp_ent = e1.data(); A.sp_ext = e2.data(); A.sp_turn = e3.data();
    A.sp_flag = spf.data(); A.owner = owner; A.lat = lat; A.tseg = tseg; A.lap = lap; A.summary = summary;
    A.status = status;
    for (int b = 0; b < B; ++b) status[b] = 0;
    if (impl == STO_QSS_PLAIN || N < 128) {
        for (int b = 0; b < B; ++b) {
            if (owner) sto::qss_plain_candidate<true>(A, *V, b, true);
            else sto::qss_plain_candidate<false>(A, *V, b, true);
        }
    } else {
        std::vector<std::vector<char>> keep;
        sto::MemoWork W = sto::carve_memo([&](size_t n) { keep.emplace_back(n); return (void*)keep.back().data(); },
                                          N, (size_t)ld, cap);
        std::vector<double> rec((size_t)N * ld * 4);
        A.rec = rec.data();
        std::vector<unsigned long long> planes((size_t)6 * W.W + STO_LIST_RING);   // a "warp of one": stride 1, lane 0
        std::vector<int32_t> ring(STO_LIST_RING);
        const sto::MemoCtx C = sto::memo_bind(planes.data(), 1, 0, ring.data(), 1, 0, N, W.W);
        // impl 1: one lane per candidate; impl 100 + G: emulated groups of G lanes (backward rows G at a time)
        for (int b = 0; b < B; ++b) {
            switch (impl) {
                case 300: sto::qss_memo3_proto(A, *V, b); break;
                case 301: sto::qss_memo_q_proto(A, W, C, *V, b); break;
                case 102: sto::qss_memo_candidate<2>(A, W, C, *V, b, true, 0, 0); break;
                case 104: sto::qss_memo_candidate<4>(A, W, C, *V, b, true, 0, 0); break;
                case 108: sto::qss_memo_candidate<8>(A, W, C, *V, b, true, 0, 0); break;
                default: sto::qss_memo_candidate<1>(A, W, C, *V, b, true, 0, 0); break;
            }
        }
    }
    return 0;
}

// fast mode (sto_fast.cuh): offsets [M][ld] -> lap[B]; optional v, a, tseg [N][ld]
int hostsim_fast(const double* cenx, const double* ceny, const double* nrmx, const double* nrmy, const double* sinb,
                 const double* ts, const double* off, int M, int N, int B, int ld, const sto_vehicle_f64* V, int rounds,
                 double* v, double* a, double* tseg, double* lap, int32_t* status) {
    std::vector<double> w(((size_t)(M + 1) + 2 * (size_t)(M + 3) + 3 * (size_t)N + 4 * (size_t)M) * ld);
    double* p = w.data();
    sto::FastArgs A{};
    A.rounds = rounds;
    A.F.cenx = cenx; A.F.ceny = ceny; A.F.nrmx = nrmx; A.F.nrmy = nrmy; A.F.off = off;
    A.F.M = M; A.F.B = B; A.F.ld = ld; A.F.status = status;
    A.F.u = p; p += (size_t)(M + 1) * ld;
    A.F.cx = p; p += (size_t)(M + 3) * ld;
    A.F.cy = p; p += (size_t)(M + 3) * ld;
    double* R = p; p += (size_t)N * ld;
    double* dd = p; p += (size_t)N * ld;
    double* df = p; p += (size_t)N * ld;
    A.F.cp = p; p += (size_t)M * ld;
    A.F.zx = p; p += (size_t)M * ld;
    A.F.zy = p; p += (size_t)M * ld;
    A.F.zz = p;
    A.E = sto::EvalArgs{A.F.u, A.F.cx, A.F.cy, ts, M, N, B, ld, nullptr, nullptr, nullptr, R, dd, df};
    A.Q.dd = dd; A.Q.df = df; A.Q.R = R; A.Q.sinb = sinb; A.Q.N = N; A.Q.B = B; A.Q.ld = ld;
    A.Q.v = v; A.Q.a = a; A.Q.tseg = tseg; A.Q.lap = lap; A.Q.status = status;
    for (int b = 0; b < B; ++b) sto::fast_candidate(A, *V, b, true, 0);
    return 0;
}

void hostsim_sp_counters(long long* out12, int reset) {
    long long* src[6] = {sto::g_sp_visits, sto::g_sp_distinct, sto::g_sp_on_live_orig, sto::g_sp_evals, sto::g_sp_changed,
                         sto::g_sp_maxlist};
    for (int k = 0; k < 6; ++k) { out12[2 * k] = src[k][0]; out12[2 * k + 1] = src[k][1]; if (reset) src[k][0] = src[k][1] = 0; }
}

// event log of the memoised schedule (analysis tooling): 4 ints per applied outcome
void hostsim_log_enable(int on) { sto::g_log_on = on; sto::g_log.clear(); }
long long hostsim_log_size(void) { return (long long)sto::g_log.size(); }
void hostsim_log_copy(int* out) { memcpy(out, sto::g_log.data(), sto::g_log.size() * sizeof(int)); }

void hostsim_att_hist(long long* out32) {
    for (int d = 0; d < 2; ++d) for (int k = 0; k < 16; ++k) { out32[d * 16 + k] = sto::g_att_hist[d][k]; sto::g_att_hist[d][k] = 0; }
}

// prototype switch + counters of memo_spawned_fwd_batch (analysis tooling)
void hostsim_fwd_batch(int on) { sto::g_fwd_batch = on; }
void hostsim_fwd_batch_counters(long long* out4, int reset) {
    for (int k = 0; k < 4; ++k) { out4[k] = sto::g_fb[k]; if (reset) sto::g_fb[k] = 0; }
}

void hostsim_m3_stats(long long* out9, int reset) {
    out9[0] = sto::g_m3.rounds[0]; out9[1] = sto::g_m3.rounds[1]; out9[2] = sto::g_m3.rounds_g8[0];
    out9[3] = sto::g_m3.rounds_g8[1]; out9[4] = sto::g_m3.evals[0]; out9[5] = sto::g_m3.evals[1];
    out9[6] = sto::g_m3.dup_created; out9[7] = sto::g_m3.dup_diverged; out9[8] = sto::g_m3.reborn;
    if (reset) sto::g_m3 = sto::Memo3Stats();
}

void hostsim_mq_config(int cap, int lanes, int war) { sto::g_mq_cap = cap; sto::g_mq_lanes = lanes; sto::g_mq_war = war; }
void hostsim_mf_config(int on) { sto::g_mq_fpar = on ? 1 : 0; sto::g_mf_allruns = on == 2; }
void hostsim_mf_stats(long long* out3, int reset) { out3[0] = sto::g_mf.rounds; out3[1] = sto::g_mf.evals; out3[2] = sto::g_mf.subpasses; if (reset) sto::g_mf = sto::MemoFStats(); }
void hostsim_mr_config(int on) { sto::g_mq_rows = on == 1; sto::g_mq_wr = on == 2; }
void hostsim_mr_stats(long long* out14, int reset) { for (int d = 0; d < 2; ++d) { const sto::MemoRStats& m = sto::g_mr[d]; long long v[7] = {m.walks, m.runs, m.evals, m.rounds_est, m.longest, m.entries_in_runs, m.visits}; for (int k = 0; k < 7; ++k) out14[7 * d + k] = v[k]; if (reset) sto::g_mr[d] = sto::MemoRStats(); } }
void hostsim_mq_stats(long long* out14, int reset) {
    for (int d = 0; d < 2; ++d) {
        out14[d] = sto::g_mq.eval_rounds[d]; out14[2 + d] = sto::g_mq.idle_rounds[d]; out14[4 + d] = sto::g_mq.evals[d];
        out14[6 + d] = sto::g_mq.chunks[d]; out14[8 + d] = sto::g_mq.qsum[d]; out14[10 + d] = sto::g_mq.visits[d];
        out14[12 + d] = sto::g_mq.walks[d];
    }
    if (reset) sto::g_mq = sto::MemoQStats();
}

void hostsim_counters(long long* out4, int reset) {
    out4[0] = sto::g_memo_evals[0]; out4[1] = sto::g_memo_evals[1];
    out4[2] = sto::g_memo_words[0]; out4[3] = sto::g_memo_words[1];
    if (reset) { sto::g_memo_evals[0] = sto::g_memo_evals[1] = sto::g_memo_words[0] = sto::g_memo_words[1] = 0; }
}

}  // extern "C"
