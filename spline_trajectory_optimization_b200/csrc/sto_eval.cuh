// sto_eval.cuh -- arc-parameter resampling of a fitted closed cubic B-spline with heading and turn radius.
//
// Replaces BSplineTrajectory.sample_along(ts=...) and its helpers __get_yaw / __get_turn_radius
// (spline_traj_optm/models/trajectory.py:250-260,268-281).  The six evaluations (x, y and their first two
// derivatives) follow SciPy's evaluate_spline -> _deBoor_D operation order exactly, so that on identical
// coefficients the samples are bit-identical to scipy.interpolate.splev (a textbook "differentiate the
// control polygon" evaluator differs by ~4e-6 absolute in the second derivative).
//
// One candidate per thread; every candidate has its own knot vector (chord-length parameter of its own
// points), so the interval search is a per-thread monotone walk that stays within +-1 knot of its warp
// neighbours (all candidates have u_i ~ i/M): knot / coefficient loads stay coalesced.
#pragma once
#include "sto_common.cuh"

namespace sto {

struct EvalArgs {
    const double *u, *cx, *cy;  // [M+1][ld], [M+3][ld], [M+3][ld]
    const double* ts;           // [N] shared parameters
    int M, N, B, ld;
    double *x, *y, *yaw, *radius, *chord_qss, *chord_norm;  // [N][ld], each optional
};

// Knot U(i), i in [-3, M+3], from the normalised chord parameter (periodic extension, period 1).
STO_HD double eval_knot(const EvalArgs& A, int i, int b) {
    if (i < 0) return A.u[at(A.M + i, A.ld, b)] - 1.0;
    if (i > A.M) return A.u[at(i - A.M, A.ld, b)] + 1.0;
    return A.u[at(i, A.ld, b)];
}

// Cubic basis values and first/second derivative bases at x for the knot window tk[0..5] = t[ell-2..ell+3],
// transcribing _deBoor_D (k = 3) for m = 0, 1, 2.  The leading "standard" iterations are common to the three
// derivative orders and are computed once; each is the identical operation sequence SciPy runs.
STO_HD void cubic_basis_d012(const double* tk, double x, double* h0, double* h1, double* h2) {
    // t[ell + n] = tk[2 + n],  t[ell + n - j] = tk[2 + n - j]
    // j = 1 (standard): from h = {1}
    double a0, a1;
    {
        double xb = tk[3], xa = tk[2];
        double w = 1.0 / (xb - xa);
        a0 = 0.0 + w * (xb - x);
        a1 = w * (x - xa);
    }
    // m = 2: derivative iterations j = 2, 3 applied to {a0, a1}
    {
        double c0, c1, c2;
        {   // j = 2
            double w1 = 2 * a0 / (tk[3] - tk[1]);
            c0 = 0.0 - w1;
            c1 = w1;
            double w2 = 2 * a1 / (tk[4] - tk[2]);
            c1 = c1 - w2;
            c2 = w2;
        }
        {   // j = 3
            double w1 = 3 * c0 / (tk[3] - tk[0]);
            double r0 = 0.0 - w1, r1 = w1;
            double w2 = 3 * c1 / (tk[4] - tk[1]);
            r1 = r1 - w2;
            double r2 = w2;
            double w3 = 3 * c2 / (tk[5] - tk[2]);
            r2 = r2 - w3;
            h2[0] = r0; h2[1] = r1; h2[2] = r2; h2[3] = w3;
        }
    }
    // j = 2 (standard): from {a0, a1}
    double b0, b1, b2;
    {
        double w1 = a0 / (tk[3] - tk[1]);
        b0 = 0.0 + w1 * (tk[3] - x);
        b1 = w1 * (x - tk[1]);
        double w2 = a1 / (tk[4] - tk[2]);
        b1 = b1 + w2 * (tk[4] - x);
        b2 = w2 * (x - tk[2]);
    }
    // m = 1: derivative iteration j = 3 applied to {b0, b1, b2}
    {
        double w1 = 3 * b0 / (tk[3] - tk[0]);
        double r0 = 0.0 - w1, r1 = w1;
        double w2 = 3 * b1 / (tk[4] - tk[1]);
        r1 = r1 - w2;
        double r2 = w2;
        double w3 = 3 * b2 / (tk[5] - tk[2]);
        r2 = r2 - w3;
        h1[0] = r0; h1[1] = r1; h1[2] = r2; h1[3] = w3;
    }
    // m = 0: j = 3 (standard)
    {
        double w1 = b0 / (tk[3] - tk[0]);
        double r0 = 0.0 + w1 * (tk[3] - x);
        double r1 = w1 * (x - tk[0]);
        double w2 = b1 / (tk[4] - tk[1]);
        r1 = r1 + w2 * (tk[4] - x);
        double r2 = w2 * (x - tk[1]);
        double w3 = b2 / (tk[5] - tk[2]);
        r2 = r2 + w3 * (tk[5] - x);
        h0[0] = r0; h0[1] = r1; h0[2] = r2; h0[3] = w3 * (x - tk[2]);
    }
}

STO_HD double dot4(const double* c, const double* h) {
    double acc = 0.0;
    acc = acc + c[0] * h[0];
    acc = acc + c[1] * h[1];
    acc = acc + c[2] * h[2];
    acc = acc + c[3] * h[3];
    return acc;
}

// trajectory.py:253-260: turn radius = 1 / | |x'y'' - y'x''| / sqrt((x'^2 + y'^2)^3) |
STO_HD double turn_radius(double dx, double dy, double ddx, double ddy) {
    double num = fabs(dx * ddy - dy * ddx);
    double s2 = dx * dx + dy * dy;
    double s6 = (s2 * s2) * s2;
    double curv = num / sqrt(s6);
    return 1.0 / fabs(curv);
}

// Register window of one lane over its candidate's spline: knots tk[0..5] = U(i-2..i+3), coefficients c[i..i+3].
struct EvalCursor {
    int i, loaded;
    double tk[6], ccx[4], ccy[4];
};

// Position and first two derivatives at parameter x (scipy _find_interval walk from the previous interval).
STO_HD void eval_at(const EvalArgs& A, int b, EvalCursor& c, double x, double* pos, double* d1, double* d2) {
    const int M = A.M, ld = A.ld;
    int i = c.i;
    while (i > 0 && x < ((c.loaded == i) ? c.tk[2] : eval_knot(A, i, b))) { --i; }
    while (i < M - 1 && x >= ((c.loaded == i) ? c.tk[3] : eval_knot(A, i + 1, b))) {
        ++i;
        if (c.loaded == i - 1) {  // slide the register window by one knot
            c.tk[0] = c.tk[1]; c.tk[1] = c.tk[2]; c.tk[2] = c.tk[3]; c.tk[3] = c.tk[4]; c.tk[4] = c.tk[5];
            c.tk[5] = eval_knot(A, i + 3, b);
            c.ccx[0] = c.ccx[1]; c.ccx[1] = c.ccx[2]; c.ccx[2] = c.ccx[3]; c.ccx[3] = A.cx[at(i + 3, ld, b)];
            c.ccy[0] = c.ccy[1]; c.ccy[1] = c.ccy[2]; c.ccy[2] = c.ccy[3]; c.ccy[3] = A.cy[at(i + 3, ld, b)];
            c.loaded = i;
        }
    }
    if (c.loaded != i) {
        for (int n = 0; n < 6; ++n) c.tk[n] = eval_knot(A, i - 2 + n, b);
        for (int n = 0; n < 4; ++n) { c.ccx[n] = A.cx[at(i + n, ld, b)]; c.ccy[n] = A.cy[at(i + n, ld, b)]; }
        c.loaded = i;
    }
    c.i = i;
    double h0[4], h1[4], h2[4];
    cubic_basis_d012(c.tk, x, h0, h1, h2);
    pos[0] = dot4(c.ccx, h0); pos[1] = dot4(c.ccy, h0);
    d1[0] = dot4(c.ccx, h1);  d1[1] = dot4(c.ccy, h1);
    d2[0] = dot4(c.ccx, h2);  d2[1] = dot4(c.ccy, h2);
}

// Samples j0 <= j < j1 of candidate b.  Samples are independent, so a candidate's range may be split over the
// lanes of a group (each lane re-evaluates the one neighbour position its first / last chord needs).
STO_HD void eval_range(const EvalArgs& A, int b, int j0, int j1) {
    const int M = A.M, N = A.N, ld = A.ld;
    if (j0 >= j1) return;
    const bool want_chord = A.chord_qss || A.chord_norm;
    EvalCursor c;
    c.loaded = -1000;
    double pos[2], d1[2], d2[2];
    double x_prev = 0.0, y_prev = 0.0;
    {   // start the interval walk near the right knot: candidates have u_i ~ i / M
        double guess = A.ts[j0 > 0 ? j0 - 1 : 0] * (double)M;
        int i = (guess > 0.0) ? (int)guess : 0;
        c.i = (i > M - 1) ? M - 1 : i;
    }
    if (want_chord && j0 > 0) {
        eval_at(A, b, c, A.ts[j0 - 1], pos, d1, d2);
        x_prev = pos[0]; y_prev = pos[1];
    }
    for (int j = j0; j < j1; ++j) {
        eval_at(A, b, c, A.ts[j], pos, d1, d2);
        if (A.x) A.x[at(j, ld, b)] = pos[0];
        if (A.y) A.y[at(j, ld, b)] = pos[1];
        if (A.yaw) A.yaw[at(j, ld, b)] = atan2(d1[1], d1[0]);
        if (A.radius) A.radius[at(j, ld, b)] = turn_radius(d1[0], d1[1], d2[0], d2[1]);
        if (want_chord) {
            if (j > 0) {
                if (A.chord_qss) A.chord_qss[at(j - 1, ld, b)] = chord_qss(x_prev, y_prev, pos[0], pos[1]);
                if (A.chord_norm) A.chord_norm[at(j - 1, ld, b)] = chord_norm(x_prev, y_prev, pos[0], pos[1]);
            }
            x_prev = pos[0]; y_prev = pos[1];
        }
    }
    if (want_chord && j1 == N) {  // closing chord N-1 -> 0
        EvalCursor c0;
        c0.loaded = -1000;
        c0.i = 0;
        eval_at(A, b, c0, A.ts[0], pos, d1, d2);
        if (A.chord_qss) A.chord_qss[at(N - 1, ld, b)] = chord_qss(x_prev, y_prev, pos[0], pos[1]);
        if (A.chord_norm) A.chord_norm[at(N - 1, ld, b)] = chord_norm(x_prev, y_prev, pos[0], pos[1]);
    }
}

STO_HD void eval_candidate(const EvalArgs& A, int b) { eval_range(A, b, 0, A.N); }

// ---- generic degree (1..5), shared knots: one thread per SAMPLE ------------------------------------------
// scipy _deBoor_D verbatim semantics for any k; used for the track's own splines (k = 3 or 5, smoothing fits
// produced on the host by FITPACK) so their sampling also runs on the device.
STO_HD void deboor_d(const double* t, double x, int k, int ell, int m, double* h /*k+1*/) {
    double hh[6];
    h[0] = 1.0;
    for (int j = 1; j <= k - m; ++j) {
        for (int n = 0; n < j; ++n) hh[n] = h[n];
        h[0] = 0.0;
        for (int n = 1; n <= j; ++n) {
            int ind = ell + n;
            double xb = t[ind], xa = t[ind - j];
            if (xb == xa) { h[n] = 0.0; continue; }
            double w = hh[n - 1] / (xb - xa);
            h[n - 1] += w * (xb - x);
            h[n] = w * (x - xa);
        }
    }
    for (int j = k - m + 1; j <= k; ++j) {
        for (int n = 0; n < j; ++n) hh[n] = h[n];
        h[0] = 0.0;
        for (int n = 1; n <= j; ++n) {
            int ind = ell + n;
            double xb = t[ind], xa = t[ind - j];
            if (xb == xa) { h[m] = 0.0; continue; }
            double w = j * hh[n - 1] / (xb - xa);
            h[n - 1] -= w;
            h[n] = w;
        }
    }
}

struct SplineEvalArgs {
    const double *t, *cx, *cy, *ts;
    int nt, k, N;
    double *x, *y, *yaw, *radius;
};

STO_HD void eval_spline_sample(const SplineEvalArgs& A, int j) {
    const int k = A.k, n = A.nt - k - 1;
    const double x = A.ts[j];
    // interval by bisection: t[ell] <= x < t[ell+1], clamped to [k, n-1] (extrapolate=True)
    int ell;
    if (!(x == x)) {
        const double qnan = nan("");
        if (A.x) A.x[j] = qnan;
        if (A.y) A.y[j] = qnan;
        if (A.yaw) A.yaw[j] = qnan;
        if (A.radius) A.radius[j] = qnan;
        return;
    }
    if (x < A.t[k + 1] || n - 1 == k) ell = k;
    else if (x >= A.t[n - 1]) ell = n - 1;
    else {
        int lo = k, hi = n - 1;  // t[lo] <= x < t[hi]
        while (hi - lo > 1) {
            int mid = (lo + hi) >> 1;
            if (x >= A.t[mid]) lo = mid; else hi = mid;
        }
        ell = lo;
    }
    double h[6], v[6];
    for (int m = 0; m < 3; ++m) {
        deboor_d(A.t, x, k, ell, m, h);
        double ax = 0.0, ay = 0.0;
        for (int a = 0; a <= k; ++a) {
            ax = ax + A.cx[ell + a - k] * h[a];
            ay = ay + A.cy[ell + a - k] * h[a];
        }
        v[2 * m] = ax; v[2 * m + 1] = ay;
    }
    if (A.x) A.x[j] = v[0];
    if (A.y) A.y[j] = v[1];
    if (A.yaw) A.yaw[j] = atan2(v[3], v[2]);
    if (A.radius) A.radius[j] = turn_radius(v[2], v[3], v[4], v[5]);
}

// ---- batch of coefficient sets on SHARED knots (any degree <= 5): one thread per (sample, candidate) -----------------
// The optimiser-loop shape of the path (reference optimization/optimizer.py:196-211,276-289): many variants of one
// spline that differ only in control points are resampled at the same parameters.  Knot interval and basis values
// depend on the sample only; the coefficient loads are coalesced across candidates.
struct SplineBatchArgs {
    const double* t;            // [nt] shared knots
    int nt, k;
    const double *cx, *cy;      // [nt-k-1][ld] coefficient sets, sample-major
    const double* ts;           // [N]
    int N, B, ld;
    double *x, *y, *yaw, *radius, *chord_qss, *chord_norm;  // [N][ld], each optional
};

STO_HD int spline_interval(const double* t, int nt, int k, double x) {
    const int n = nt - k - 1;
    if (x < t[k + 1] || n - 1 == k) return k;
    if (x >= t[n - 1]) return n - 1;
    int lo = k, hi = n - 1;  // t[lo] <= x < t[hi]
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (x >= t[mid]) lo = mid; else hi = mid;
    }
    return lo;
}

STO_HD void eval_spline_batch_sample(const SplineBatchArgs& A, int j, int b) {
    const int k = A.k, ld = A.ld;
    const double x = A.ts[j];
    const int ell = spline_interval(A.t, A.nt, k, x);
    double h[6], v[6];
    for (int m = 0; m < 3; ++m) {
        deboor_d(A.t, x, k, ell, m, h);
        double ax = 0.0, ay = 0.0;
        for (int a = 0; a <= k; ++a) {
            ax = ax + A.cx[at(ell + a - k, ld, b)] * h[a];
            ay = ay + A.cy[at(ell + a - k, ld, b)] * h[a];
        }
        v[2 * m] = ax; v[2 * m + 1] = ay;
    }
    if (A.x) A.x[at(j, ld, b)] = v[0];
    if (A.y) A.y[at(j, ld, b)] = v[1];
    if (A.yaw) A.yaw[at(j, ld, b)] = atan2(v[3], v[2]);
    if (A.radius) A.radius[at(j, ld, b)] = turn_radius(v[2], v[3], v[4], v[5]);
    if (A.chord_qss || A.chord_norm) {  // chord j -> j+1 (cyclic): the neighbour's position is re-evaluated here
        const int jn = (j + 1 == A.N) ? 0 : j + 1;
        const double xn = A.ts[jn];
        const int elln = spline_interval(A.t, A.nt, k, xn);
        deboor_d(A.t, xn, k, elln, 0, h);
        double bx = 0.0, by = 0.0;
        for (int a = 0; a <= k; ++a) {
            bx = bx + A.cx[at(elln + a - k, ld, b)] * h[a];
            by = by + A.cy[at(elln + a - k, ld, b)] * h[a];
        }
        if (A.chord_qss) A.chord_qss[at(j, ld, b)] = chord_qss(v[0], v[1], bx, by);
        if (A.chord_norm) A.chord_norm[at(j, ld, b)] = chord_norm(v[0], v[1], bx, by);
    }
}

// ---- arc length of the sample intervals (DIST_TO_SF_BWD / FWD columns), one spline, one thread per interval ----------
// The reference integrates |r'(t)| over [ts[i-1], ts[i]] with adaptive QUADPACK, one Python call per sample
// (models/trajectory.py:228-230,283-289; >99 % of sample_along's time, and the QSS never reads the result).  Here each
// interval is split at the knots it contains and every piece (where x', y' are polynomials) gets a fixed 8-point
// Gauss-Legendre rule: ~1e-13 relative quadrature error, i.e. agreement with quad at its own tolerance (1.5e-8).
struct ArcArgs {
    const double *t, *cx, *cy, *ts;
    int nt, k, N;
    double* sec;  // [N]: sec[0] = 0, sec[i] = length of the curve between ts[i-1] and ts[i]
};

STO_HD double spline_speed(const ArcArgs& A, double x) {
    const int k = A.k;
    const int ell = spline_interval(A.t, A.nt, k, x);
    double h[6];
    deboor_d(A.t, x, k, ell, 1, h);
    double dx = 0.0, dy = 0.0;
    for (int a = 0; a <= k; ++a) {
        dx = dx + A.cx[ell + a - k] * h[a];
        dy = dy + A.cy[ell + a - k] * h[a];
    }
    return sqrt(dx * dx + dy * dy);
}

STO_HD void arc_section(const ArcArgs& A, int i) {
    if (i == 0) { A.sec[0] = 0.0; return; }
    // 8-point Gauss-Legendre nodes / weights on [-1, 1]
    const double gx[4] = {0.1834346424956498, 0.5255324099163290, 0.7966664774136267, 0.9602898564975363};
    const double gw[4] = {0.3626837833783620, 0.3137066458778873, 0.2223810344533745, 0.1012285362903763};
    const double a = A.ts[i - 1], b = A.ts[i];
    const int n = A.nt - A.k - 1;
    double total = 0.0, lo = a;
    int ell = spline_interval(A.t, A.nt, A.k, a);
    while (lo < b) {
        double hi = b;
        if (ell + 1 <= n - 1 && A.t[ell + 1] < b) hi = A.t[ell + 1];   // stop at the next knot inside the interval
        if (hi <= lo) { ++ell; if (ell > n - 1) break; continue; }
        const double c = 0.5 * (lo + hi), hw = 0.5 * (hi - lo);
        double acc = 0.0;
        for (int q = 0; q < 4; ++q)
            acc += gw[q] * (spline_speed(A, c - hw * gx[q]) + spline_speed(A, c + hw * gx[q]));
        total += hw * acc;
        lo = hi;
        ++ell;
    }
    A.sec[i] = total;
}

}  // namespace sto
