// sto_fit_lsq.cuh -- periodic least-squares B-spline fit of a closed line on FIXED knots, one candidate per thread.
//
// non-interpolating candidates.  The reference builds smoothed lines with
// BSplineTrajectory(coords, s > 0, k) (spline_traj_optm/models/trajectory.py:213-223 -> scipy splprep(per=True) ->
// FITPACK clocur), whose knot selection is an adaptive, data-dependent iteration.  With the knot vector held fixed -
// the knots of the track's own smoothing fit (models/race_track.py:23-29), shared by every candidate - the fit is
// FITPACK's task = -1: the weighted least-squares closed spline on the given knots, chord-length parameter u as in the
// interpolating fit.  That is what this file computes, for any degree 1 <= k <= 5:
//
//     minimise  sum_i | p_i - sum_j c_j B_j(u_i) |^2   subject to  c_{g+j} = c_j (j < k),   g = nt - 2k - 1
//
// FITPACK (fpclos) triangularises the observation matrix with Givens rotations; here the (k+1)-banded cyclic normal
// equations are accumulated (points of one knot interval share a (k+1) x (k+1) block that lives in registers) and
// solved by a bordered banded Cholesky factorisation: the last k unknowns form the border, the rest is a plain band.
// The least-squares solution is unique, so the two agree up to rounding (normal equations: ~cond(A)^2 eps; 1e-14 on
// the Monza lines, cond ~ 2e2; BASELINE tolerance 1e-9).  A knot interval without data makes a pivot vanish: the
// candidate is flagged STO_CAND_DEGENERATE_FIT (FITPACK returns ier = 10 there) and gets NaN coefficients.
//
// Memory: everything per candidate is sample-major ([row][ld]); every access below is coalesced across the 32
// candidates of a warp.  Scratch per candidate: g (k+1) + 2 g + (g - k) k doubles (7.6 KB at g = 112, k = 3).
#pragma once
#include "sto_common.cuh"
#include "sto_fit.cuh"

namespace sto {

struct LsqArgs {
    FitArgs F;          // points (centre + offset * normal, or px/py), M, B, ld, u[M+1][ld] (output), status
    const double* t;    // [nt] shared knot vector (t[k] = 0, t[nt-k-1] = 1, periodic extension on both sides)
    int nt, k;
    double *cx, *cy;    // outputs [nt-k-1][ld]
    double* band;       // work [g (k+1)][ld]: lower cyclic band a[i][d] = A[i][(i - d) mod g]; the Cholesky factor later
    double *rx, *ry;    // work [g][ld]: right-hand sides -> solution
    double* Y;          // work [(g - k) k][ld]: border columns U, then L^-1 U
};

STO_HD int lsq_free_coefficients(int nt, int k) { return nt - 2 * k - 1; }
// the bordered solver needs the two corner blocks of the cyclic band to stay apart
STO_HD bool lsq_sizes_ok(int M, int nt, int k) {
    const int g = nt - 2 * k - 1;
    return k >= 1 && k <= 5 && g >= 3 * k + 1 && M >= g;
}
STO_HD size_t lsq_work_doubles(int nt, int k) {
    const size_t g = (size_t)(nt - 2 * k - 1);
    return g * (size_t)(k + 1) + 2 * g + (g - (size_t)k) * (size_t)k;
}

template <int K>
STO_HD void lsq_fail(const LsqArgs& A, int b) {
    if (A.F.status) A.F.status[b] |= STO_CAND_DEGENERATE_FIT;
    const double qnan = nan("");
    for (int j = 0; j < A.nt - K - 1; ++j) { A.cx[at(j, A.F.ld, b)] = qnan; A.cy[at(j, A.F.ld, b)] = qnan; }
}

template <int K>
STO_HD void lsq_candidate(const LsqArgs& A, int b) {
    const FitArgs& F = A.F;
    const int M = F.M, ld = F.ld, nt = A.nt, g = nt - 2 * K - 1, n1 = g - K;
    const double* t = A.t;
    // chord-length parameter, exactly as the interpolating fit (FITPACK parcur / clocur, ipar = 0)
    fit_phase_segments(F, b, 0, 1);
    const double total = fit_phase_cumsum(F, b);
    if (!(total > 0.0)) {
        const double qnan = nan("");
        for (int i = 0; i <= M; ++i) F.u[at(i, ld, b)] = qnan;
        lsq_fail<K>(A, b);
        return;
    }
    fit_phase_normalise(F, b, 0, 1, total);

    for (int i = 0; i < g * (K + 1); ++i) A.band[at(i, ld, b)] = 0.0;
    for (int i = 0; i < g; ++i) { A.rx[at(i, ld, b)] = 0.0; A.ry[at(i, ld, b)] = 0.0; }

    // ---- normal equations: points of one knot interval share the block of coefficients l-K .. l ----------------
    double blk[K + 1][K + 1], bx[K + 1], by[K + 1];
#pragma unroll
    for (int a = 0; a <= K; ++a) {
        bx[a] = by[a] = 0.0;
#pragma unroll
        for (int c = 0; c <= K; ++c) blk[a][c] = 0.0;
    }
    int l = K;
    for (int i = 0; i <= M; ++i) {   // i == M only flushes the last block
        int li = l;
        double x = 0.0, px = 0.0, py = 0.0;
        if (i < M) {
            x = F.u[at(i, ld, b)];
            fit_point(F, i, b, px, py);
            while (li < nt - K - 2 && x >= t[li + 1]) ++li;   // t[li] <= x < t[li+1] (u is non-decreasing)
        }
        if (i == M || li != l) {
#pragma unroll
            for (int a = 0; a <= K; ++a) {
                int row = l - K + a;
                if (row >= g) row -= g;
#pragma unroll
                for (int c = 0; c <= K; ++c)
                    if (c <= a) {
                        A.band[at(row * (K + 1) + (a - c), ld, b)] += blk[a][c];
                        blk[a][c] = 0.0;
                    }
                A.rx[at(row, ld, b)] += bx[a];
                A.ry[at(row, ld, b)] += by[a];
                bx[a] = by[a] = 0.0;
            }
            l = li;
        }
        if (i == M) break;
        // the K+1 non-zero B-splines at x (FITPACK fpbspl / de Boor-Cox recurrence)
        double h[K + 1], hh[K + 1];
        h[0] = 1.0;
#pragma unroll
        for (int j = 1; j <= K; ++j) {
#pragma unroll
            for (int m = 0; m < j; ++m) hh[m] = h[m];
            h[0] = 0.0;
#pragma unroll
            for (int m = 1; m <= j; ++m) {
                const double tl = t[l + m], tj = t[l + m - j];
                const double f = hh[m - 1] / (tl - tj);
                h[m - 1] = h[m - 1] + f * (tl - x);
                h[m] = f * (x - tj);
            }
        }
#pragma unroll
        for (int a = 0; a <= K; ++a) {
#pragma unroll
            for (int c = 0; c <= K; ++c)
                if (c <= a) blk[a][c] = blk[a][c] + h[a] * h[c];
            bx[a] = bx[a] + h[a] * px;
            by[a] = by[a] + h[a] * py;
        }
    }

    // ---- bordered banded Cholesky: A = [[T, U], [U^T, D]], T = rows/cols < n1 (plain band), border = last K ---------
#define STO_BAND(i, d) A.band[at((i) * (K + 1) + (d), ld, b)]
#define STO_Y(i, c) A.Y[at((i) * K + (c), ld, b)]
    for (int i = 0; i < n1; ++i) {
#pragma unroll
        for (int c = 0; c < K; ++c) {
            double v = 0.0;
            if (i < K && c >= i) v = STO_BAND(i, K + i - c);                        // wrapped entries of the first rows
            if (i >= n1 - K && c <= K + i - n1) v = STO_BAND(n1 + c, n1 + c - i);   // band of the border rows
            STO_Y(i, c) = v;
        }
    }
    bool ok = true;
    for (int i = 0; i < n1; ++i) {   // T = L L^T in place (lower band), then L z = [r, U] row by row
        const int j0 = (i - K > 0) ? i - K : 0;
        for (int j = j0; j <= i; ++j) {
            double s = STO_BAND(i, i - j);
            for (int m = j0; m < j; ++m) s = s - STO_BAND(i, i - m) * STO_BAND(j, j - m);
            if (j == i) {
                if (!(s > 0.0)) { ok = false; s = 1.0; }
                STO_BAND(i, 0) = sqrt(s);
            } else {
                STO_BAND(i, i - j) = s / STO_BAND(j, 0);
            }
        }
        const double dii = STO_BAND(i, 0);
        double vx = A.rx[at(i, ld, b)], vy = A.ry[at(i, ld, b)];
        double vu[K];
#pragma unroll
        for (int c = 0; c < K; ++c) vu[c] = STO_Y(i, c);
        for (int m = j0; m < i; ++m) {
            const double lim = STO_BAND(i, i - m);
            vx = vx - lim * A.rx[at(m, ld, b)];
            vy = vy - lim * A.ry[at(m, ld, b)];
#pragma unroll
            for (int c = 0; c < K; ++c) vu[c] = vu[c] - lim * STO_Y(m, c);
        }
        A.rx[at(i, ld, b)] = vx / dii;
        A.ry[at(i, ld, b)] = vy / dii;
#pragma unroll
        for (int c = 0; c < K; ++c) STO_Y(i, c) = vu[c] / dii;
    }
    // Schur complement of the border: S = D - Y^T Y, right-hand side r_D - Y^T z
    double S[K][K], wx[K], wy[K];
#pragma unroll
    for (int c1 = 0; c1 < K; ++c1) {
        wx[c1] = A.rx[at(n1 + c1, ld, b)];
        wy[c1] = A.ry[at(n1 + c1, ld, b)];
#pragma unroll
        for (int c2 = 0; c2 < K; ++c2) S[c1][c2] = (c2 <= c1) ? STO_BAND(n1 + c1, c1 - c2) : 0.0;
    }
    for (int i = 0; i < n1; ++i) {
        double yi[K];
#pragma unroll
        for (int c = 0; c < K; ++c) yi[c] = STO_Y(i, c);
        const double zx = A.rx[at(i, ld, b)], zy = A.ry[at(i, ld, b)];
#pragma unroll
        for (int c1 = 0; c1 < K; ++c1) {
            wx[c1] = wx[c1] - yi[c1] * zx;
            wy[c1] = wy[c1] - yi[c1] * zy;
#pragma unroll
            for (int c2 = 0; c2 < K; ++c2)
                if (c2 <= c1) S[c1][c2] = S[c1][c2] - yi[c1] * yi[c2];
        }
    }
#pragma unroll
    for (int c1 = 0; c1 < K; ++c1) {   // dense Cholesky of S, forward substitution
#pragma unroll
        for (int c2 = 0; c2 < K; ++c2) {
            if (c2 > c1) continue;
            double s = S[c1][c2];
#pragma unroll
            for (int m = 0; m < K; ++m)
                if (m < c2) s = s - S[c1][m] * S[c2][m];
            if (c2 == c1) {
                if (!(s > 0.0)) { ok = false; s = 1.0; }
                S[c1][c1] = sqrt(s);
            } else {
                S[c1][c2] = s / S[c2][c2];
            }
        }
#pragma unroll
        for (int m = 0; m < K; ++m)
            if (m < c1) { wx[c1] = wx[c1] - S[c1][m] * wx[m]; wy[c1] = wy[c1] - S[c1][m] * wy[m]; }
        wx[c1] = wx[c1] / S[c1][c1];
        wy[c1] = wy[c1] / S[c1][c1];
    }
#pragma unroll
    for (int c1 = K - 1; c1 >= 0; --c1) {   // back substitution: border unknowns
#pragma unroll
        for (int m = K - 1; m >= 0; --m)
            if (m > c1) { wx[c1] = wx[c1] - S[m][c1] * wx[m]; wy[c1] = wy[c1] - S[m][c1] * wy[m]; }
        wx[c1] = wx[c1] / S[c1][c1];
        wy[c1] = wy[c1] / S[c1][c1];
    }
    if (!ok) { lsq_fail<K>(A, b); return; }
    // L^T x = z - Y w, rows n1-1 .. 0
    for (int i = n1 - 1; i >= 0; --i) {
        double vx = A.rx[at(i, ld, b)], vy = A.ry[at(i, ld, b)];
#pragma unroll
        for (int c = 0; c < K; ++c) {
            const double y = STO_Y(i, c);
            vx = vx - y * wx[c];
            vy = vy - y * wy[c];
        }
        for (int d = 1; d <= K && i + d < n1; ++d) {
            const double lji = STO_BAND(i + d, d);
            vx = vx - lji * A.rx[at(i + d, ld, b)];
            vy = vy - lji * A.ry[at(i + d, ld, b)];
        }
        const double dii = STO_BAND(i, 0);
        A.rx[at(i, ld, b)] = vx / dii;
        A.ry[at(i, ld, b)] = vy / dii;
    }
#undef STO_BAND
#undef STO_Y
    // coefficients in SciPy's order: c[0 .. g-1], then the periodic wrap c[g + j] = c[j], j < K
    for (int j = 0; j < n1; ++j) {
        A.cx[at(j, ld, b)] = A.rx[at(j, ld, b)];
        A.cy[at(j, ld, b)] = A.ry[at(j, ld, b)];
    }
#pragma unroll
    for (int c = 0; c < K; ++c) {
        A.cx[at(n1 + c, ld, b)] = wx[c];
        A.cy[at(n1 + c, ld, b)] = wy[c];
    }
    for (int j = 0; j < K; ++j) {
        A.cx[at(g + j, ld, b)] = A.cx[at(j, ld, b)];
        A.cy[at(g + j, ld, b)] = A.cy[at(j, ld, b)];
    }
}

STO_HD void lsq_candidate_k(const LsqArgs& A, int b) {
    switch (A.k) {
        case 1: lsq_candidate<1>(A, b); break;
        case 2: lsq_candidate<2>(A, b); break;
        case 3: lsq_candidate<3>(A, b); break;
        case 4: lsq_candidate<4>(A, b); break;
        default: lsq_candidate<5>(A, b); break;
    }
}

}  // namespace sto
