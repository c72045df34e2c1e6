// sto_fit.cuh -- periodic cubic interpolating B-spline fit of a closed line, one candidate per thread.
//
// Replaces BSplineTrajectory.__init__ (spline_traj_optm/models/trajectory.py:213-223) for s=0, k=3, i.e.
// scipy.interpolate.splprep(per=True) = FITPACK clocur: chord-length parameter u (normalised, u[M] = 1),
// knots at every data parameter with period-1 extension, interpolation conditions -> cyclic tridiagonal
// system (identical matrix for x and y), solved here by Thomas elimination + Sherman-Morrison.
//
// Memory: everything is sample-major [i][ld]; the five passes below stream those arrays once each, fully
// coalesced across the 32 candidates of a warp.  Algorithmic HBM bytes per candidate: read 8*M (offsets),
// write 8*(M+1) + 16*(M+3) (u, cx, cy); scratch traffic (4 work arrays) is ~5x that and stays in L2 for
// small batches.
#pragma once
#include "sto_common.cuh"

#ifndef STO_FIT_CHUNK
#define STO_FIT_CHUNK 8
#endif

namespace sto {

struct FitArgs {
    const double *cenx, *ceny, *nrmx, *nrmy;  // [M] track centre line and left normal (NULL: use px/py)
    const double* off;                          // [M][ld] lateral offsets
    const double *px, *py;                      // [M][ld] explicit points (when cenx == NULL)
    int M, B, ld;
    double *u, *cx, *cy;                        // outputs: [M+1][ld], [M+3][ld], [M+3][ld]
    int32_t* status;                            // [B] (may be NULL)
    double *cp, *zx, *zy, *zz;                  // work: [M][ld] each
};

STO_HD void fit_point(const FitArgs& A, int i, int b, double& x, double& y) {
    if (A.cenx) {
        double o = A.off[at(i, A.ld, b)];
        x = A.cenx[i] + o * A.nrmx[i];
        y = A.ceny[i] + o * A.nrmy[i];
    } else {
        x = A.px[at(i, A.ld, b)];
        y = A.py[at(i, A.ld, b)];
    }
}

// Periodic knot extension: U(i) for i in [-3, M+3] given the normalised u[0..M].
STO_HD double fit_knot(const FitArgs& A, int i, int b) {
    const int M = A.M;
    if (i < 0) return A.u[at(M + i, A.ld, b)] - 1.0;   // t[3-i] = t[M+3-i] - per
    if (i > M) return A.u[at(i - M, A.ld, b)] + 1.0;   // t[M+3+i] = t[3+i] + per
    return A.u[at(i, A.ld, b)];
}

// The fit as phases that a group of G lanes runs together (G = 1: one lane does everything; the arithmetic and its
// order are identical for every G, so the coefficients are bit-identical).  Per-row work that does not depend on the
// recurrences - segment lengths, normalisation, the collocation coefficients b0/b2, the final e -> c map - is split
// over the lanes; the three Thomas recurrences (x, y and the Sherman-Morrison vector, which share the pivots) run on
// three lanes side by side.  `sync` is __syncwarp on the device and nothing on the host.
#if defined(__CUDA_ARCH__)
#define STO_FIT_SYNC() __syncwarp()
#else
#define STO_FIT_SYNC()
#endif

// phase A: segment lengths (FITPACK clocur: dist = sum_dim (x_i - x_{i-1})**2), rows split over lanes
STO_HD void fit_phase_segments(const FitArgs& A, int b, int g, int G) {
    const int M = A.M, ld = A.ld;
    const int per = (M + G - 1) / G;
    const int i0 = 1 + g * per, i1 = (i0 + per <= M + 1) ? i0 + per : M + 1;   // rows i0 <= i < i1 of 1..M
    if (i0 >= i1) return;
    double xp, yp;
    fit_point(A, i0 - 1, b, xp, yp);
    for (int c0 = i0; c0 < i1; c0 += STO_FIT_CHUNK) {
        double xs[STO_FIT_CHUNK], ys[STO_FIT_CHUNK];
#pragma unroll
        for (int k = 0; k < STO_FIT_CHUNK; ++k) {
            const int i = c0 + k;
            if (i < i1) fit_point(A, (i < M) ? i : 0, b, xs[k], ys[k]);   // closed loop: point M == point 0
            else { xs[k] = ys[k] = 0.0; }
        }
#pragma unroll
        for (int k = 0; k < STO_FIT_CHUNK; ++k) {
            const int i = c0 + k;
            if (i < i1) {
                double dx = xs[k] - xp, dy = ys[k] - yp;
                double dist = 0.0;
                dist = dist + dx * dx;
                dist = dist + dy * dy;
                A.u[at(i, ld, b)] = sqrt(dist);
                xp = xs[k]; yp = ys[k];
            }
        }
    }
}

// phase B (one lane): running sum u(i) = u(i-1) + seg(i) in row order; returns the total
STO_HD double fit_phase_cumsum(const FitArgs& A, int b) {
    const int M = A.M, ld = A.ld;
    double acc = 0.0;
    A.u[at(0, ld, b)] = 0.0;
    for (int c0 = 1; c0 <= M; c0 += STO_FIT_CHUNK) {
        double seg[STO_FIT_CHUNK];
#pragma unroll
        for (int k = 0; k < STO_FIT_CHUNK; ++k) seg[k] = (c0 + k <= M) ? A.u[at(c0 + k, ld, b)] : 0.0;
#pragma unroll
        for (int k = 0; k < STO_FIT_CHUNK; ++k)
            if (c0 + k <= M) { acc = acc + seg[k]; A.u[at(c0 + k, ld, b)] = acc; }
    }
    return acc;
}

// phase C: normalise (rows split over lanes); phase D: collocation coefficients b0 -> cx[j], b2 -> cy[j] (scratch use)
STO_HD void fit_phase_normalise(const FitArgs& A, int b, int g, int G, double total) {
    const int M = A.M, ld = A.ld;
    for (int i = 1 + g; i < M; i += G) A.u[at(i, ld, b)] = A.u[at(i, ld, b)] / total;
    if (g == 0) A.u[at(M, ld, b)] = 1.0;
}
STO_HD void fit_phase_rows(const FitArgs& A, int b, int g, int G) {
    const int M = A.M, ld = A.ld;
    for (int j = g; j < M; j += G) {
        const double t0 = fit_knot(A, j - 2, b), t1 = fit_knot(A, j - 1, b), t2 = fit_knot(A, j, b),
                     t3 = fit_knot(A, j + 1, b), t4 = fit_knot(A, j + 2, b);
        A.cx[at(j, ld, b)] = ((t3 - t2) * (t3 - t2)) / ((t3 - t0) * (t3 - t1));
        A.cy[at(j, ld, b)] = ((t2 - t1) * (t2 - t1)) / ((t4 - t1) * (t3 - t1));
    }
}

// phases E + F: Thomas forward elimination and back substitution of T = A - w v^T (Sherman-Morrison split of the two
// corner entries) for right-hand side `which` (0: x, 1: y, 2: the vector w); every caller recomputes the shared pivots.
STO_HD void fit_phase_thomas(const FitArgs& A, int b, int which) {
    const int M = A.M, ld = A.ld;
    double* z = (which == 0) ? A.zx : (which == 1) ? A.zy : A.zz;
    double beta = 0.0, gamma = 0.0, alpha = 0.0, cpp = 0.0, zp = 0.0;
    for (int j0 = 0; j0 < M; j0 += STO_FIT_CHUNK) {
        double b0s[STO_FIT_CHUNK], b2s[STO_FIT_CHUNK], rs[STO_FIT_CHUNK];
#pragma unroll
        for (int k = 0; k < STO_FIT_CHUNK; ++k) {
            const int j = j0 + k;
            b0s[k] = b2s[k] = rs[k] = 0.0;
            if (j < M) {
                b0s[k] = A.cx[at(j, ld, b)];
                b2s[k] = A.cy[at(j, ld, b)];
                if (which < 2) { double px, py; fit_point(A, j, b, px, py); rs[k] = which ? py : px; }
            }
        }
#pragma unroll
        for (int k = 0; k < STO_FIT_CHUNK; ++k) {
            const int j = j0 + k;
            if (j >= M) break;
            const double b0 = b0s[k], b2 = b2s[k];
            const double b1 = (1.0 - b0) - b2;
            double den, ncp, nz;
            if (j == 0) {
                beta = b0;       // corner: row 0, column M-1
                gamma = -b1;
                den = b1 - gamma;
                ncp = b2 / den;
                nz = ((which == 2) ? gamma : rs[k]) / den;
            } else {
                double dj = b1, wj = 0.0;
                if (j == M - 1) {
                    alpha = b2;  // corner: row M-1, column 0
                    dj = b1 - alpha * beta / gamma;
                    wj = alpha;
                }
                den = dj - b0 * cpp;
                ncp = b2 / den;
                nz = (((which == 2) ? wj : rs[k]) - b0 * zp) / den;
            }
            if (which == 0) A.cp[at(j, ld, b)] = ncp;
            z[at(j, ld, b)] = nz;
            cpp = ncp; zp = nz;
        }
    }
    // back substitution (every lane reads lane 0's cp: written above by the lane with which == 0, made visible by the
    // caller's sync between the two halves when lanes differ)
}
STO_HD void fit_phase_backsub(const FitArgs& A, int b, int which) {
    const int M = A.M, ld = A.ld;
    double* z = (which == 0) ? A.zx : (which == 1) ? A.zy : A.zz;
    double zn = z[at(M - 1, ld, b)];
    for (int j0 = M - 2; j0 >= 0; j0 -= STO_FIT_CHUNK) {
        double cs[STO_FIT_CHUNK], zs[STO_FIT_CHUNK];
#pragma unroll
        for (int k = 0; k < STO_FIT_CHUNK; ++k) {
            const int j = j0 - k;
            cs[k] = (j >= 0) ? A.cp[at(j, ld, b)] : 0.0;
            zs[k] = (j >= 0) ? z[at(j, ld, b)] : 0.0;
        }
#pragma unroll
        for (int k = 0; k < STO_FIT_CHUNK; ++k) {
            const int j = j0 - k;
            if (j < 0) break;
            zn = zs[k] - cs[k] * zn;
            z[at(j, ld, b)] = zn;
        }
    }
}

// One-lane variants of phases E + F: the three right-hand sides share every pivot, so a single lane runs them in one
// loop (4 divisions per row instead of 3 x 2).
STO_HD void fit_phase_thomas_all(const FitArgs& A, int b) {
    const int M = A.M, ld = A.ld;
    double beta = 0.0, gamma = 0.0, alpha = 0.0, cpp = 0.0, zxp = 0.0, zyp = 0.0, zzp = 0.0;
    for (int j0 = 0; j0 < M; j0 += STO_FIT_CHUNK) {
        double b0s[STO_FIT_CHUNK], b2s[STO_FIT_CHUNK], pxs[STO_FIT_CHUNK], pys[STO_FIT_CHUNK];
#pragma unroll
        for (int k = 0; k < STO_FIT_CHUNK; ++k) {
            const int j = j0 + k;
            b0s[k] = b2s[k] = pxs[k] = pys[k] = 0.0;
            if (j < M) {
                b0s[k] = A.cx[at(j, ld, b)];
                b2s[k] = A.cy[at(j, ld, b)];
                fit_point(A, j, b, pxs[k], pys[k]);
            }
        }
#pragma unroll
        for (int k = 0; k < STO_FIT_CHUNK; ++k) {
            const int j = j0 + k;
            if (j >= M) break;
            const double b0 = b0s[k], b2 = b2s[k];
            const double b1 = (1.0 - b0) - b2;
            double den, ncp, nzx, nzy, nzz;
            if (j == 0) {
                beta = b0;
                gamma = -b1;
                den = b1 - gamma;
                ncp = b2 / den;
                nzx = pxs[k] / den;
                nzy = pys[k] / den;
                nzz = gamma / den;
            } else {
                double dj = b1, wj = 0.0;
                if (j == M - 1) {
                    alpha = b2;
                    dj = b1 - alpha * beta / gamma;
                    wj = alpha;
                }
                den = dj - b0 * cpp;
                ncp = b2 / den;
                nzx = (pxs[k] - b0 * zxp) / den;
                nzy = (pys[k] - b0 * zyp) / den;
                nzz = (wj - b0 * zzp) / den;
            }
            A.cp[at(j, ld, b)] = ncp;
            A.zx[at(j, ld, b)] = nzx;
            A.zy[at(j, ld, b)] = nzy;
            A.zz[at(j, ld, b)] = nzz;
            cpp = ncp; zxp = nzx; zyp = nzy; zzp = nzz;
        }
    }
}
STO_HD void fit_phase_backsub_all(const FitArgs& A, int b) {
    const int M = A.M, ld = A.ld;
    double zxn = A.zx[at(M - 1, ld, b)], zyn = A.zy[at(M - 1, ld, b)], zzn = A.zz[at(M - 1, ld, b)];
    for (int j0 = M - 2; j0 >= 0; j0 -= STO_FIT_CHUNK) {
        double cs[STO_FIT_CHUNK], xs[STO_FIT_CHUNK], ys[STO_FIT_CHUNK], zs[STO_FIT_CHUNK];
#pragma unroll
        for (int k = 0; k < STO_FIT_CHUNK; ++k) {
            const int j = j0 - k;
            cs[k] = xs[k] = ys[k] = zs[k] = 0.0;
            if (j >= 0) {
                cs[k] = A.cp[at(j, ld, b)]; xs[k] = A.zx[at(j, ld, b)];
                ys[k] = A.zy[at(j, ld, b)]; zs[k] = A.zz[at(j, ld, b)];
            }
        }
#pragma unroll
        for (int k = 0; k < STO_FIT_CHUNK; ++k) {
            const int j = j0 - k;
            if (j < 0) break;
            zxn = xs[k] - cs[k] * zxn;
            zyn = ys[k] - cs[k] * zyn;
            zzn = zs[k] - cs[k] * zzn;
            A.zx[at(j, ld, b)] = zxn;
            A.zy[at(j, ld, b)] = zyn;
            A.zz[at(j, ld, b)] = zzn;
        }
    }
}

// phase G + H: Sherman-Morrison correction, e -> c with the one-slot rotation and the periodic wrap (rows split)
STO_HD void fit_corner(const FitArgs& A, int b, double& beta, double& gamma) {
    const double b0_first = A.cx[at(0, A.ld, b)], b2_first = A.cy[at(0, A.ld, b)];  // still the b0/b2 scratch
    beta = b0_first;
    gamma = -((1.0 - b0_first) - b2_first);
}
STO_HD void fit_phase_coefficients(const FitArgs& A, int b, int g, int G, double beta, double gamma) {
    const int M = A.M, ld = A.ld;
    const double zx0 = A.zx[at(0, ld, b)], zy0 = A.zy[at(0, ld, b)], zz0 = A.zz[at(0, ld, b)];
    const double zxl = A.zx[at(M - 1, ld, b)], zyl = A.zy[at(M - 1, ld, b)], zzl = A.zz[at(M - 1, ld, b)];
    const double denom = (1.0 + zz0) + beta * zzl / gamma;
    const double fx = (zx0 + beta * zxl / gamma) / denom;
    const double fy = (zy0 + beta * zyl / gamma) / denom;
    // cx/cy stop being the b0/b2 scratch here: the Thomas phases are done and the corner was read by the caller
    for (int i = g; i < M; i += G) {
        const double zzi = A.zz[at(i, ld, b)];
        const double ex = A.zx[at(i, ld, b)] - fx * zzi;
        const double ey = A.zy[at(i, ld, b)] - fy * zzi;
        const int ci = (i + 1 == M) ? 0 : i + 1;
        A.cx[at(ci, ld, b)] = ex;
        A.cy[at(ci, ld, b)] = ey;
        if (ci < 3) {
            A.cx[at(M + ci, ld, b)] = ex;
            A.cy[at(M + ci, ld, b)] = ey;
        }
    }
}

STO_HD void fit_degenerate(const FitArgs& A, int b) {  // FITPACK returns ier=10; scipy raises
    const int M = A.M, ld = A.ld;
    if (A.status) A.status[b] |= STO_CAND_DEGENERATE_FIT;
    const double qnan = nan("");
    for (int i = 0; i <= M; ++i) A.u[at(i, ld, b)] = qnan;
    for (int i = 0; i < M + 3; ++i) { A.cx[at(i, ld, b)] = qnan; A.cy[at(i, ld, b)] = qnan; }
}

// One lane of a group of G (device).  Lanes 0..2 run the three recurrences when G >= 3.
STO_HD void fit_candidate_lane(const FitArgs& A, int b, bool active, int g, int G, int lane0) {
    (void)lane0;
    if (active) fit_phase_segments(A, b, g, G);
    STO_FIT_SYNC();
    double total = 0.0;
    if (active && g == 0) total = fit_phase_cumsum(A, b);
#if defined(__CUDA_ARCH__)
    total = __shfl_sync(0xffffffffu, total, lane0);
#endif
    const bool ok = total > 0.0;
    STO_FIT_SYNC();
    if (active && !ok) { if (g == 0) fit_degenerate(A, b); }
    if (active && ok) fit_phase_normalise(A, b, g, G, total);
    STO_FIT_SYNC();
    if (active && ok) fit_phase_rows(A, b, g, G);
    STO_FIT_SYNC();
    if (active && ok) {
        if (G >= 3) { if (g < 3) fit_phase_thomas(A, b, g); }
        else if (g == 0) fit_phase_thomas_all(A, b);
    }
    STO_FIT_SYNC();
    if (active && ok) {
        if (G >= 3) { if (g < 3) fit_phase_backsub(A, b, g); }
        else if (g == 0) fit_phase_backsub_all(A, b);
    }
    STO_FIT_SYNC();
    double beta = 0.0, gamma = 1.0;
    if (active && ok) fit_corner(A, b, beta, gamma);
    STO_FIT_SYNC();  // every lane has read row 0's b0/b2 before anybody overwrites cx/cy
    if (active && ok) fit_phase_coefficients(A, b, g, G, beta, gamma);
}

// Whole fit by one caller: G = 1 is the plain one-lane fit; G > 1 plays the lanes of a group one after another (host
// emulation of the device schedule, used by the tests).
STO_HD void fit_candidate(const FitArgs& A, int b, int G = 1) {
    for (int g = 0; g < G; ++g) fit_phase_segments(A, b, g, G);
    const double total = fit_phase_cumsum(A, b);
    if (!(total > 0.0)) { fit_degenerate(A, b); return; }
    for (int g = 0; g < G; ++g) fit_phase_normalise(A, b, g, G, total);
    for (int g = 0; g < G; ++g) fit_phase_rows(A, b, g, G);
    if (G >= 3) {
        for (int w = 0; w < 3; ++w) fit_phase_thomas(A, b, w);
        for (int w = 0; w < 3; ++w) fit_phase_backsub(A, b, w);
    } else {
        fit_phase_thomas_all(A, b);
        fit_phase_backsub_all(A, b);
    }
    double beta, gamma;
    fit_corner(A, b, beta, gamma);
    for (int g = 0; g < G; ++g) fit_phase_coefficients(A, b, g, G, beta, gamma);
}

}  // namespace sto
