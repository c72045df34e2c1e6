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
    double *cp, *zx, *zy, *zz, *ze;             // work: [M][ld] each (ze: the FITPACK solver only)
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

#if defined(__CUDA_ARCH__)
// phase B for a group of G lanes: the additions stay one chain in row order (u is bit-identical to the one-lane sum and
// to FITPACK's), but the loads and stores are spread over the lanes: lane g fetches row c0 + g of every chunk of G
// rows, the values travel by __shfl_sync, every lane carries the same running sum and keeps the prefix of its own row.
// The chain is then bound by the DADD latency instead of a global round trip per STO_FIT_CHUNK rows.
STO_D double fit_phase_cumsum_group(const FitArgs& A, int b, bool active, int g, int G, int lane0) {
    const int M = A.M, ld = A.ld;
    double acc = 0.0;
    if (active && g == 0) A.u[at(0, ld, b)] = 0.0;
    double vn = (active && 1 + g <= M) ? A.u[at(1 + g, ld, b)] : 0.0;
    for (int c0 = 1; c0 <= M; c0 += G) {
        const double v = vn;
        const int i2 = c0 + G + g;
        vn = (active && i2 <= M) ? A.u[at(i2, ld, b)] : 0.0;
        double mine = 0.0;
        for (int k = 0; k < G; ++k) {
            acc = acc + __shfl_sync(0xffffffffu, v, lane0 + k);   // rows past M contribute + 0.0 (exact)
            if (k == g) mine = acc;
        }
        if (active && c0 + g <= M) A.u[at(c0 + g, ld, b)] = mine;
    }
    return acc;
}
#endif

// phase C: normalise (rows split over lanes); phase D: collocation coefficients b0 -> cx[j], b2 -> cy[j] (scratch use)
STO_HD void fit_phase_normalise(const FitArgs& A, int b, int g, int G, double total) {
    const int M = A.M, ld = A.ld;
    const int per = (M - 1 + G - 1) / G;                       // rows 1 .. M-1 in contiguous blocks, chunked loads
    const int i0 = 1 + g * per, i1 = (i0 + per < M) ? i0 + per : M;
    for (int c0 = i0; c0 < i1; c0 += STO_FIT_CHUNK) {
        double us[STO_FIT_CHUNK];
#pragma unroll
        for (int k = 0; k < STO_FIT_CHUNK; ++k) us[k] = (c0 + k < i1) ? A.u[at(c0 + k, ld, b)] : 1.0;
#pragma unroll
        for (int k = 0; k < STO_FIT_CHUNK; ++k)
            if (c0 + k < i1) A.u[at(c0 + k, ld, b)] = us[k] / total;
    }
    if (g == 0) A.u[at(M, ld, b)] = 1.0;
}
STO_HD void fit_phase_rows(const FitArgs& A, int b, int g, int G) {
    // every lane takes a contiguous block of rows and slides a five-knot window along it: one new knot per row,
    // fetched STO_FIT_CHUNK rows ahead (the value of each row is a pure function of its knots: any split is bit-identical)
    const int M = A.M, ld = A.ld;
    const int per = (M + G - 1) / G;
    const int j0 = g * per, j1 = (j0 + per < M) ? j0 + per : M;
    if (j0 >= j1) return;
    double t0 = fit_knot(A, j0 - 2, b), t1 = fit_knot(A, j0 - 1, b), t2 = fit_knot(A, j0, b), t3 = fit_knot(A, j0 + 1, b);
    for (int c0 = j0; c0 < j1; c0 += STO_FIT_CHUNK) {
        double nk[STO_FIT_CHUNK];
#pragma unroll
        for (int k = 0; k < STO_FIT_CHUNK; ++k) nk[k] = (c0 + k < j1) ? fit_knot(A, c0 + k + 2, b) : 0.0;
#pragma unroll
        for (int k = 0; k < STO_FIT_CHUNK; ++k) {
            const int j = c0 + k;
            if (j < j1) {
                const double t4 = nk[k];
                A.cx[at(j, ld, b)] = ((t3 - t2) * (t3 - t2)) / ((t3 - t0) * (t3 - t1));
                A.cy[at(j, ld, b)] = ((t2 - t1) * (t2 - t1)) / ((t4 - t1) * (t3 - t1));
                t0 = t1; t1 = t2; t2 = t3; t3 = t4;
            }
        }
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

// ---- partitioned cyclic solve: a GROUP of P lanes per candidate (long, finely sampled tracks) ------------------------
// The Thomas recurrences above are one dependent chain of M divisions per candidate; at M ~ 23 k (0.25 m sampling)
// that chain, not the batch, bounds the fit.  Here the M rows are cut into P contiguous blocks, one per lane:
//   1. every lane eliminates the interior of its block (rows s .. e-2) with the two neighbouring interface unknowns
//      X_{k-1} = e_{s-1} and X_k = e_{e-1} kept symbolic:  e_j + cp_j e_{j+1} = z_j - zl_j X_{k-1};
//   2. a backward recurrence gives the block's first unknown as e_s = y - wl X_{k-1} - wr X_k;
//   3. row e-1 of every block, with e_{e-2} from (1) and e_e from the NEXT lane's (2) (one __shfl_sync), is a cyclic
//      tridiagonal equation in (X_{k-1}, X_k, X_{k+1}): P equations, solved across the lanes by parallel cyclic
//      reduction on warp shuffles (log2 P - 1 steps and a final 2 x 2 solve; the period is handled by the ring of
//      lanes itself, no Sherman-Morrison correction);
//   4. every lane back-substitutes its block and writes the coefficients (e -> c rotation and wrap as above).
// Critical path: 3 sweeps over M / P rows + log2 P shuffle rounds instead of 2 sweeps over M rows.  The arithmetic
// differs from the one-lane Thomas solve in rounding only (coefficients agree to ~1e-15 relative; tests/).
struct PartFwd { double cp, zx, zy, zl; };     // forward-eliminated row e-2 of a block
struct PartFirst { double yx, yy, wl, wr; };   // first unknown of a block: e_s = y - wl X_{k-1} - wr X_k
struct PartEq { double a, d, c, rx, ry; };     // a X_{k-h} + d X_k + c X_{k+h} = r

STO_HD int fit_part_begin(int M, int k, int P) { return (int)(((long long)k * M) / P); }   // balanced blocks, >= M/P rows

STO_HD PartFwd fit_part_forward(const FitArgs& A, int b, int s, int e) {
    const int ld = A.ld;
    double cpp = 0.0, zxp = 0.0, zyp = 0.0, zlp = 0.0;
    for (int j0 = s; j0 < e - 1; j0 += STO_FIT_CHUNK) {
        double b0s[STO_FIT_CHUNK], b2s[STO_FIT_CHUNK], pxs[STO_FIT_CHUNK], pys[STO_FIT_CHUNK];
#pragma unroll
        for (int k = 0; k < STO_FIT_CHUNK; ++k) {
            const int j = j0 + k;
            b0s[k] = b2s[k] = pxs[k] = pys[k] = 0.0;
            if (j < e - 1) {
                b0s[k] = A.cx[at(j, ld, b)];
                b2s[k] = A.cy[at(j, ld, b)];
                fit_point(A, j, b, pxs[k], pys[k]);
            }
        }
#pragma unroll
        for (int k = 0; k < STO_FIT_CHUNK; ++k) {
            const int j = j0 + k;
            if (j >= e - 1) break;
            const double b0 = b0s[k], b2 = b2s[k];
            const double b1 = (1.0 - b0) - b2;
            const double inv = 1.0 / (b1 - b0 * cpp);          // one division per row; cpp = 0 on the block's first row
            const double ncp = b2 * inv;
            const double nzx = (pxs[k] - b0 * zxp) * inv;
            const double nzy = (pys[k] - b0 * zyp) * inv;
            const double nzl = ((j == s) ? b0 : 0.0 - b0 * zlp) * inv;   // coupling to X_{k-1} enters on the first row only
            A.cp[at(j, ld, b)] = ncp;
            A.zx[at(j, ld, b)] = nzx;
            A.zy[at(j, ld, b)] = nzy;
            A.zz[at(j, ld, b)] = nzl;
            cpp = ncp; zxp = nzx; zyp = nzy; zlp = nzl;
        }
    }
    return PartFwd{cpp, zxp, zyp, zlp};
}

STO_HD PartFirst fit_part_first(const FitArgs& A, int b, int s, int e, const PartFwd& last) {
    const int ld = A.ld;
    double yx = last.zx, yy = last.zy, wl = last.zl, wr = last.cp;
    for (int j0 = e - 3; j0 >= s; j0 -= STO_FIT_CHUNK) {
        double cs[STO_FIT_CHUNK], xs[STO_FIT_CHUNK], ys[STO_FIT_CHUNK], ls[STO_FIT_CHUNK];
#pragma unroll
        for (int k = 0; k < STO_FIT_CHUNK; ++k) {
            const int j = j0 - k;
            cs[k] = xs[k] = ys[k] = ls[k] = 0.0;
            if (j >= s) {
                cs[k] = A.cp[at(j, ld, b)]; xs[k] = A.zx[at(j, ld, b)];
                ys[k] = A.zy[at(j, ld, b)]; ls[k] = A.zz[at(j, ld, b)];
            }
        }
#pragma unroll
        for (int k = 0; k < STO_FIT_CHUNK; ++k) {
            if (j0 - k < s) break;
            yx = xs[k] - cs[k] * yx;
            yy = ys[k] - cs[k] * yy;
            wl = ls[k] - cs[k] * wl;
            wr = 0.0 - cs[k] * wr;
        }
    }
    return PartFirst{yx, yy, wl, wr};
}

// Row e-1 of a block: b0 e_{e-2} + b1 X_k + b2 e_e = p, with e_{e-2} from this block and e_e from the next one.
STO_HD PartEq fit_part_equation(const FitArgs& A, int b, int e, const PartFwd& own, const PartFirst& next) {
    const int i = e - 1;
    const double b0 = A.cx[at(i, A.ld, b)], b2 = A.cy[at(i, A.ld, b)];
    const double b1 = (1.0 - b0) - b2;
    double px, py;
    fit_point(A, i, b, px, py);
    PartEq q;
    q.a = 0.0 - b0 * own.zl;
    q.d = (b1 - b0 * own.cp) - b2 * next.wl;
    q.c = 0.0 - b2 * next.wr;
    q.rx = (px - b0 * own.zx) - b2 * next.yx;
    q.ry = (py - b0 * own.zy) - b2 * next.yy;
    return q;
}

// One step of parallel cyclic reduction: eliminate X_{k-h} and X_{k+h} with the equations of lanes k-h and k+h.
STO_HD PartEq fit_pcr_combine(const PartEq& me, const PartEq& lo, const PartEq& hi) {
    const double al = 0.0 - me.a / lo.d, ga = 0.0 - me.c / hi.d;
    PartEq r;
    r.a = al * lo.a;
    r.c = ga * hi.c;
    r.d = (me.d + al * lo.c) + ga * hi.a;
    r.rx = (me.rx + al * lo.rx) + ga * hi.rx;
    r.ry = (me.ry + al * lo.ry) + ga * hi.ry;
    return r;
}
// Last step (h = P/2): lanes k-h and k+h are the same lane `o`; a 2 x 2 system between k and o.
STO_HD void fit_pcr_final(const PartEq& me, const PartEq& o, double& x, double& y) {
    const double S = me.a + me.c, So = o.a + o.c;
    const double det = me.d * o.d - S * So;
    x = (me.rx * o.d - S * o.rx) / det;
    y = (me.ry * o.d - S * o.ry) / det;
}
STO_HD void fit_pcr_single(const PartEq& me, double& x, double& y) {   // P = 1: X couples to itself on both sides
    const double d = (me.a + me.d) + me.c;
    x = me.rx / d;
    y = me.ry / d;
}

STO_HD void fit_store_coef(const FitArgs& A, int b, int i, double ex, double ey) {
    const int M = A.M, ld = A.ld;
    const int ci = (i + 1 == M) ? 0 : i + 1;
    A.cx[at(ci, ld, b)] = ex;
    A.cy[at(ci, ld, b)] = ey;
    if (ci < 3) {
        A.cx[at(M + ci, ld, b)] = ex;
        A.cy[at(M + ci, ld, b)] = ey;
    }
}

STO_HD void fit_part_backsub(const FitArgs& A, int b, int s, int e, double xpx, double xpy, double xkx, double xky) {
    const int ld = A.ld;
    fit_store_coef(A, b, e - 1, xkx, xky);
    double xn = xkx, yn = xky;
    for (int j0 = e - 2; j0 >= s; j0 -= STO_FIT_CHUNK) {
        double cs[STO_FIT_CHUNK], xs[STO_FIT_CHUNK], ys[STO_FIT_CHUNK], ls[STO_FIT_CHUNK];
#pragma unroll
        for (int k = 0; k < STO_FIT_CHUNK; ++k) {
            const int j = j0 - k;
            cs[k] = xs[k] = ys[k] = ls[k] = 0.0;
            if (j >= s) {
                cs[k] = A.cp[at(j, ld, b)]; xs[k] = A.zx[at(j, ld, b)];
                ys[k] = A.zy[at(j, ld, b)]; ls[k] = A.zz[at(j, ld, b)];
            }
        }
#pragma unroll
        for (int k = 0; k < STO_FIT_CHUNK; ++k) {
            const int j = j0 - k;
            if (j < s) break;
            xn = (xs[k] - ls[k] * xpx) - cs[k] * xn;
            yn = (ys[k] - ls[k] * xpy) - cs[k] * yn;
            fit_store_coef(A, b, j, xn, yn);
        }
    }
}

// Number of blocks: a function of M ONLY (never of the batch size or the lane count), so a line's coefficients do not
// depend on how many lanes happened to share it: 32 blocks from M = 256 on, one block below.
STO_HD int fit_part_blocks(int M) { return (M >= 256) ? 32 : 1; }

#if defined(__CUDA_ARCH__)
STO_D PartEq fit_shfl_eq(const PartEq& q, int src) {
    PartEq r;
    r.a = __shfl_sync(0xffffffffu, q.a, src);
    r.d = __shfl_sync(0xffffffffu, q.d, src);
    r.c = __shfl_sync(0xffffffffu, q.c, src);
    r.rx = __shfl_sync(0xffffffffu, q.rx, src);
    r.ry = __shfl_sync(0xffffffffu, q.ry, src);
    return r;
}
STO_D PartFirst fit_shfl_first(const PartFirst& h, int src) {
    PartFirst r;
    r.yx = __shfl_sync(0xffffffffu, h.yx, src);
    r.yy = __shfl_sync(0xffffffffu, h.yy, src);
    r.wl = __shfl_sync(0xffffffffu, h.wl, src);
    r.wr = __shfl_sync(0xffffffffu, h.wr, src);
    return r;
}
// Device: lane l of a group of L = P / E lanes that shares candidate b; the lane owns blocks l, l + L, l + 2 L, ...
// (slot j = block l + j L) and their interface equations.  A PCR step of stride h pairs equation k with k -/+ h: for
// h < L those sit in lanes l -/+ h (same slot, or the neighbouring slot when the lane index wraps - the SOURCE lane
// knows whether its one reader wraps and offers the right slot, so each partner costs one set of shuffles); for
// h >= L they are other slots of the same lane.  Every equation goes through exactly the combines of the emulation
// below whatever E is, so the result is bit-identical for every lane count.  Whole warps call this.
template <int E, int P>
STO_D void fit_solve_partitioned_lane(const FitArgs& A, int b, bool active, int l, int lane0) {
    constexpr int L = P / E;   // everything below indexes the slot arrays with compile-time values: they stay in registers
    const int M = A.M;
    // few slots: every loop over them is unrolled and the slot arrays stay in registers; many slots (big batches, one or
    // two lanes per line): plain loops over local-memory arrays - the sweeps dominate there and stay compact
    constexpr int kSweepUnroll = (E <= 4) ? E : 1;
    constexpr int kStepUnroll = (E <= 4) ? 5 : 1;
    PartFwd f[E];
    PartFirst h[E];
#pragma unroll kSweepUnroll
    for (int j = 0; j < E; ++j) {
        f[j] = PartFwd{0.0, 0.0, 0.0, 0.0};
        h[j] = PartFirst{0.0, 0.0, 0.0, 0.0};
        if (active) {
            const int k = l + j * L;
            const int s = fit_part_begin(M, k, P), e = fit_part_begin(M, k + 1, P);
            f[j] = fit_part_forward(A, b, s, e);
            h[j] = fit_part_first(A, b, s, e, f[j]);
        }
    }
    PartEq q[E];
    {   // block k + 1 is slot j of lane l + 1, or slot j + 1 of lane 0 when l is the last lane
        const int nl = lane0 + ((l + 1) & (L - 1));
#pragma unroll kSweepUnroll
        for (int j = 0; j < E; ++j) {
            const PartFirst offer = (l == 0) ? h[(j + 1) % E] : h[j];   // lane 0's reader is the last lane: it wraps
            const PartFirst hn = fit_shfl_first(offer, nl);
            q[j] = PartEq{0.0, 1.0, 0.0, 0.0, 0.0};
            if (active) q[j] = fit_part_equation(A, b, fit_part_begin(M, l + j * L + 1, P), f[j], hn);
        }
    }
#pragma unroll kStepUnroll
    for (int hh = 1; 2 * hh < P; hh <<= 1) {
        PartEq n[E];
        if (hh < L) {
            const int lo_l = lane0 + ((l + L - hh) & (L - 1)), hi_l = lane0 + ((l + hh) & (L - 1));
            const bool lo_reader_wraps = l >= L - hh;   // my reader l + hh - L < hh reads me as its k - h partner
            const bool hi_reader_wraps = l < hh;        // my reader l - hh + L reads me as its k + h partner
#pragma unroll kSweepUnroll
            for (int j = 0; j < E; ++j) {
                const PartEq lo = fit_shfl_eq(lo_reader_wraps ? q[(j + E - 1) % E] : q[j], lo_l);
                const PartEq hi = fit_shfl_eq(hi_reader_wraps ? q[(j + 1) % E] : q[j], hi_l);
                n[j] = fit_pcr_combine(q[j], lo, hi);
            }
        } else {
            const int m = hh / L;   // partners are m slots away in this lane (compile-time after unrolling)
#pragma unroll kSweepUnroll
            for (int j = 0; j < E; ++j) n[j] = fit_pcr_combine(q[j], q[(j + E - m) % E], q[(j + m) % E]);
        }
#pragma unroll kSweepUnroll
        for (int j = 0; j < E; ++j) q[j] = n[j];
    }
    double x[E], y[E];
    if (P >= 2) {
        constexpr int hh = P / 2;
        if (hh < L) {   // E == 1: the partner is lane l + L/2, no wrap distinction (both directions reach the same lane)
#pragma unroll kSweepUnroll
            for (int j = 0; j < E; ++j) {
                const PartEq o = fit_shfl_eq(q[j], lane0 + ((l + hh) & (L - 1)));
                fit_pcr_final(q[j], o, x[j], y[j]);
            }
        } else {
            const int m = hh / L;
#pragma unroll kSweepUnroll
            for (int j = 0; j < E; ++j) fit_pcr_final(q[j], q[(j + m) % E], x[j], y[j]);
        }
    } else {
        fit_pcr_single(q[0], x[0], y[0]);
    }
    {   // X_{k-1}: slot j of lane l - 1, or slot j - 1 of the last lane when l == 0
        const int pl = lane0 + ((l + L - 1) & (L - 1));
#pragma unroll kSweepUnroll
        for (int j = 0; j < E; ++j) {
            const bool reader_wraps = (l == L - 1);     // my reader is lane 0
            const double ox = reader_wraps ? x[(j + E - 1) % E] : x[j], oy = reader_wraps ? y[(j + E - 1) % E] : y[j];
            const double xpx = __shfl_sync(0xffffffffu, ox, pl), xpy = __shfl_sync(0xffffffffu, oy, pl);
            if (active) {
                const int k = l + j * L;
                fit_part_backsub(A, b, fit_part_begin(M, k, P), fit_part_begin(M, k + 1, P), xpx, xpy, x[j], y[j]);
            }
        }
    }
}
// E = blocks per lane (the caller launches G = 32 / E lanes per candidate when M >= 256, one lane below).
template <int E>
STO_D void fit_solve_partitioned_group(const FitArgs& A, int b, bool active, int g, int lane0) {
    if (fit_part_blocks(A.M) == 1) {   // tiny tracks: one block, one lane (any other lanes of the group idle)
        fit_solve_partitioned_lane<1, 1>(A, b, active && g == 0, 0, lane0 + g);
        return;
    }
    fit_solve_partitioned_lane<E, 32>(A, b, active, g, lane0);
}
#endif

// The same schedule played by one caller (host emulation for the tests; also the reference for what the lanes do).
STO_HD void fit_solve_partitioned_emulated(const FitArgs& A, int b, int P) {
    const int M = A.M;
    PartFwd f[32];
    PartFirst h[32];
    PartEq q[32], qn[32];
    double x[32], y[32];
    for (int g = 0; g < P; ++g) {
        const int s = fit_part_begin(M, g, P), e = fit_part_begin(M, g + 1, P);
        f[g] = fit_part_forward(A, b, s, e);
        h[g] = fit_part_first(A, b, s, e, f[g]);
    }
    for (int g = 0; g < P; ++g) q[g] = fit_part_equation(A, b, fit_part_begin(M, g + 1, P), f[g], h[(g + 1) & (P - 1)]);
    for (int hh = 1; 2 * hh < P; hh <<= 1) {
        for (int g = 0; g < P; ++g) qn[g] = fit_pcr_combine(q[g], q[(g + P - hh) & (P - 1)], q[(g + hh) & (P - 1)]);
        for (int g = 0; g < P; ++g) q[g] = qn[g];
    }
    for (int g = 0; g < P; ++g) {
        if (P >= 2) fit_pcr_final(q[g], q[(g + P / 2) & (P - 1)], x[g], y[g]);
        else fit_pcr_single(q[g], x[g], y[g]);
    }
    for (int g = 0; g < P; ++g) {
        const int s = fit_part_begin(M, g, P), e = fit_part_begin(M, g + 1, P);
        const int pv = (g + P - 1) & (P - 1);
        fit_part_backsub(A, b, s, e, x[pv], y[pv], x[g], y[g]);
    }
}

// ---- FITPACK's own arithmetic: fpclos for s = 0, k = 3 (the reference's fit, bit for bit) ---------------------------
// BSplineTrajectory.__init__ (models/trajectory.py:219-220) is scipy.interpolate.splprep(per=True) = FITPACK clocur ->
// fpclos (Dierckx; SciPy 1.18.1, un-vendored).  For an interpolating closed cubic, fpclos puts a knot at every data
// parameter, evaluates the three non-zero B-splines of every point with fpbspl (de Boor-Cox recurrence), rotates the
// M observation rows one by one into an upper triangular band matrix a1 (bandwidth 3) with Givens rotations WITHOUT
// square-root-free tricks (fpgivs / fprota), carries the last two columns - which the periodicity condition couples to
// the first rows - in a dense M x 2 block a2, rotates the last two rows through ALL of a1 (the periodic wrap), and
// back-substitutes (fpbacp).  This is a restatement of that published algorithm, operation by operation, from the
// FITPACK routines' documented structure; scipy's coefficients are reproduced bit for bit on every golden line and on
// 4,096 bench lines (tests/), so laps downstream equal the reference's to the last bit as well.
// The sweep is one dependent chain of ~5 M rotations per line (one lane per line); only the B-spline values are
// computed ahead by all lanes of the group.
struct FpRot { double c, s; };
STO_HD FpRot fp_givs(double piv, double& ww) {           // fpgivs: dd = sqrt(piv^2 + ww^2) without overflow
    // if (store >= ww) dd = store * sqrt(1 + (ww / piv)^2) else dd = ww * sqrt(1 + (piv / ww)^2): the same operations
    // on selected operands, so the lanes of a warp (different lines) do not diverge
    const double store = fabs(piv);
    const bool big = store >= ww;
    const double num = big ? ww : piv, den = big ? piv : ww, base = big ? store : ww;
    FpRot g;
#if defined(__CUDA_ARCH__) && !defined(STO_NO_FAST_FP64)
    {   // straight-line copy of the operators' fast paths (sto_common.cuh): cos and sin overlap; redone below if flagged
        bool slow = false;
        const double r = div_fast(num, den, slow);
        const double dd = base * sqrt_fast(1.0 + r * r, slow);
        g.c = div_fast(ww, dd, slow);
        g.s = div_fast(piv, dd, slow);
        if (!slow) { ww = dd; return g; }
    }
#endif
    const double r = num / den;
    const double dd = base * sqrt(1.0 + r * r);
    g.c = ww / dd;
    g.s = piv / dd;
    ww = dd;
    return g;
}
// fpgivs for a periodic row on its way through the band.  The row's fill-in shrinks by ~0.27 per band row (the decay
// of the cyclic coupling), so from band row ~16 on |piv| < 2^-30 ww, and from row ~510 on it sits in the denormals
// for the rest of the sweep (it never reaches 0: cos ~ 0.97 times one denormal unit rounds back to one unit).  There
// fpgivs takes its second form and collapses, operation by operation: r = piv / ww; r * r <= 2^-60, so 1 + r * r
// rounds to 1; sqrt(1) = 1; dd = ww * 1 = ww; cos = ww / dd = ww / ww = 1; sin = piv / dd = piv / ww = r.  One division
// instead of a division, a square root and two divisions - and the denormal rows no longer fall out of every fast
// path (they were 82 % of the sweep: ncu showed both slow-path calls taken on 604 k of 741 k visits).
STO_HD FpRot fp_givs_decayed(double piv, double& ww) {
    const double store = fabs(piv);
    if (store < 0x1p-30 * ww && ww >= 0x1p-500 && ww < INFINITY) {
        FpRot g;
        g.c = 1.0;
#if defined(__CUDA_ARCH__) && !defined(STO_NO_FAST_FP64)
        if (store >= 0x1p-960) {
            bool slow = false;
            g.s = div_fast(piv, ww, slow);
            if (slow) g.s = piv / ww;
        } else {
            g.s = piv / ww;                      // denormal numerator or quotient: the operator's own slow path
        }
#else
        g.s = piv / ww;
#endif
        return g;                                // ww stays: dd == ww
    }
    return fp_givs(piv, ww);
}
STO_HD void fp_rota(const FpRot& g, double& a, double& b) {   // fprota
    const double s1 = a, s2 = b;
    b = g.c * s2 + g.s * s1;
    a = g.c * s1 - g.s * s2;
}

// fpbspl at a knot: the (k + 1 = 4) B-spline values at x = t(l) for row i (0-based), l = i + 4; the fourth is 0.
// Results go to the slots where row i of a1 will be stored once the row has been rotated in.
STO_HD void fit_phase_bspl_rows(const FitArgs& A, int b, int g, int G) {
    const int M = A.M, ld = A.ld;
    const int per = (M + G - 1) / G;
    const int i0 = g * per, i1 = (i0 + per < M) ? i0 + per : M;
    for (int i = i0; i < i1; ++i) {
        double t[7];   // t[m] = t(l - 3 + m), m = 0..6 (1-based FITPACK knot index l = i + 4)
#pragma unroll
        for (int m = 0; m < 7; ++m) t[m] = fit_knot(A, i - 3 + m, b);
        const double x = t[3];
        double h[5], hh[4];
        h[0] = 1.0;
#pragma unroll
        for (int j = 1; j <= 3; ++j) {
#pragma unroll
            for (int m = 0; m < j; ++m) hh[m] = h[m];
            h[0] = 0.0;
#pragma unroll
            for (int m = 1; m <= j; ++m) {
                const double tli = t[3 + m], tlj = t[3 + m - j];
                if (tli == tlj) { h[m] = 0.0; continue; }
                const double f = hh[m - 1] / (tli - tlj);
                h[m - 1] = h[m - 1] + f * (tli - x);
                h[m] = f * (x - tlj);
            }
        }
        A.cp[at(i, ld, b)] = h[0];
        A.zx[at(i, ld, b)] = h[1];
        A.zy[at(i, ld, b)] = h[2];
        // ... and the point itself goes where row i of the right-hand side z will be stored (one strided load per
        // row in the rotation chain instead of an offset, four track values and the arithmetic)
        double px, py;
        fit_point(A, i, b, px, py);
        A.cx[at(i, ld, b)] = px;
        A.cy[at(i, ld, b)] = py;
    }
}

#ifndef STO_FP_CHUNK
#define STO_FP_CHUNK 4   // rows fetched ahead of the rotation chain (global round trips off the critical path)
#endif

// One periodic row on its way through the band (fpclos labels 160-240): h1(1..3) = (ha, hb, hc), h2(1..2) = (g1, g2).
struct FpWrapRow { double ha, hb, hc, g1, g2, x, y; };

// fpclos "rotation with the rows 1,2,...n10 of matrix a", one row jj (1-based) held in registers by the caller.
STO_HD void fp_wrap_step(FpWrapRow& r, int jj, int n10, double& a1, double& a2, double& a3, double& b1, double& b2,
                         double& zx, double& zy) {
    const int kk = 2;
    const double piv = r.ha;
    if (piv == 0.0) { r.ha = r.hb; r.hb = r.hc; r.hc = 0.0; return; }
    const FpRot g = fp_givs_decayed(piv, a1);
    fp_rota(g, r.x, zx); fp_rota(g, r.y, zy);
    fp_rota(g, r.g1, b1); fp_rota(g, r.g2, b2);
    if (jj == n10) return;
    const int i2 = (n10 - jj < kk) ? n10 - jj : kk;
    if (i2 >= 1) { fp_rota(g, r.hb, a2); r.ha = r.hb; }
    if (i2 >= 2) { fp_rota(g, r.hc, a3); r.hb = r.hc; }
    if (i2 >= 2) r.hc = 0.0; else r.hb = 0.0;   // h1(i2 + 1) = 0
}

// Part 1 (one lane): rows 1 .. M-2 into the band, the a2 block, the two periodic rows ready to travel.
STO_HD void fitpack_regular(const FitArgs& A, int b, FpWrapRow* wr, double* init21, double* init22) {
    const int M = A.M, ld = A.ld;
    const int n7 = M, kk = 2, n10 = n7 - kk;   // kk1 = 3
    double* const a11 = A.cp; double* const a12 = A.zx; double* const a13 = A.zy;   // a1(j, 1..3), rows 0-based
    double* const a21 = A.zz; double* const a22 = A.ze;                            // a2(j, 1..2)
    double* const z1 = A.cx; double* const z2 = A.cy;                              // right-hand sides, then c in place
    // B-spline values of the two periodic rows (their slots are overwritten at the end of the regular sweep)
    double hp[2][3], pp[2][2];
    for (int r = 0; r < 2; ++r) {
        const int i = M - 2 + r;
        hp[r][0] = a11[at(i, ld, b)]; hp[r][1] = a12[at(i, ld, b)]; hp[r][2] = a13[at(i, ld, b)];
        pp[r][0] = z1[at(i, ld, b)]; pp[r][1] = z2[at(i, ld, b)];
    }
    // ---- rows 1 .. M-2: rotate into a1 (three rows of a1 and z live in registers; row `it` is final afterwards) ------
    double w1[3] = {0.0, 0.0, 0.0}, w2[3] = {0.0, 0.0, 0.0}, w3[3] = {0.0, 0.0, 0.0};
    double q1[2] = {0.0, 0.0}, q2[2] = {0.0, 0.0}, q3[2] = {0.0, 0.0};
    // the rows of the NEXT chunk are requested before the current chunk's rotation chain starts, so a global round trip
    // overlaps ~STO_FP_CHUNK rows of arithmetic instead of preceding them (30 % of the stall samples before)
    double nhs[STO_FP_CHUNK][3], nps[STO_FP_CHUNK][2];
#pragma unroll
    for (int c = 0; c < STO_FP_CHUNK; ++c) {
        nhs[c][0] = nhs[c][1] = nhs[c][2] = nps[c][0] = nps[c][1] = 0.0;
        if (c < M - 2) {
            nhs[c][0] = a11[at(c, ld, b)]; nhs[c][1] = a12[at(c, ld, b)]; nhs[c][2] = a13[at(c, ld, b)];
            nps[c][0] = z1[at(c, ld, b)]; nps[c][1] = z2[at(c, ld, b)];
        }
    }
    for (int i0 = 0; i0 < M - 2; i0 += STO_FP_CHUNK) {
        double hs[STO_FP_CHUNK][3], ps[STO_FP_CHUNK][2];
#pragma unroll
        for (int c = 0; c < STO_FP_CHUNK; ++c) {
            hs[c][0] = nhs[c][0]; hs[c][1] = nhs[c][1]; hs[c][2] = nhs[c][2]; ps[c][0] = nps[c][0]; ps[c][1] = nps[c][1];
        }
#pragma unroll
        for (int c = 0; c < STO_FP_CHUNK; ++c) {
            const int i = i0 + STO_FP_CHUNK + c;
            if (i < M - 2) {
                nhs[c][0] = a11[at(i, ld, b)]; nhs[c][1] = a12[at(i, ld, b)]; nhs[c][2] = a13[at(i, ld, b)];
                nps[c][0] = z1[at(i, ld, b)]; nps[c][1] = z2[at(i, ld, b)];
            }
        }
#pragma unroll
        for (int c = 0; c < STO_FP_CHUNK; ++c) {
            const int i = i0 + c;
            if (i >= M - 2) break;
            double h1 = hs[c][0], h2 = hs[c][1], h3 = hs[c][2];
            double xi = ps[c][0], yi = ps[c][1];
            if (h1 != 0.0) {
                const FpRot g = fp_givs(h1, w1[0]);
                fp_rota(g, xi, q1[0]); fp_rota(g, yi, q1[1]);
                fp_rota(g, h2, w1[1]); fp_rota(g, h3, w1[2]);
            }
            if (h2 != 0.0) {
                const FpRot g = fp_givs(h2, w2[0]);
                fp_rota(g, xi, q2[0]); fp_rota(g, yi, q2[1]);
                fp_rota(g, h3, w2[1]);
            }
            if (h3 != 0.0) {
                // band row i + 2 is untouched so far (all zero): fpgivs gives dd = |h3| * sqrt(1 + (0 / h3)^2) = |h3|,
                // cos = 0 / dd = 0, sin = h3 / dd = 1 (B-spline values are >= 0), and fprota leaves
                // z = 0 * 0 + 1 * x = 0 + x - the rotation costs nothing, bit for bit
                w3[0] = fabs(h3);
                q3[0] = 0.0 + xi; q3[1] = 0.0 + yi;
            }
            a11[at(i, ld, b)] = w1[0]; a12[at(i, ld, b)] = w1[1]; a13[at(i, ld, b)] = w1[2];
            z1[at(i, ld, b)] = q1[0]; z2[at(i, ld, b)] = q1[1];
#pragma unroll
            for (int k = 0; k < 3; ++k) { w1[k] = w2[k]; w2[k] = w3[k]; w3[k] = 0.0; }
#pragma unroll
            for (int k = 0; k < 2; ++k) { q1[k] = q2[k]; q2[k] = q3[k]; q3[k] = 0.0; }
        }
    }
    // the window now holds rows M-2 and M-1 (0-based) of a1 / z
    a11[at(M - 2, ld, b)] = w1[0]; a12[at(M - 2, ld, b)] = w1[1]; a13[at(M - 2, ld, b)] = w1[2];
    z1[at(M - 2, ld, b)] = q1[0]; z2[at(M - 2, ld, b)] = q1[1];
    a11[at(M - 1, ld, b)] = w2[0]; a12[at(M - 1, ld, b)] = w2[1]; a13[at(M - 1, ld, b)] = w2[2];
    z1[at(M - 1, ld, b)] = q2[0]; z2[at(M - 1, ld, b)] = q2[1];
    // ---- a2: the last two columns of the band, as a dense M x 2 block (zero except six entries; the zeros of rows
    //      1 .. n10 are supplied on the fly below, where those rows are first touched) ---------------------------------
    for (int k = 0; k < 3; ++k) { init21[k] = 0.0; init22[k] = 0.0; }   // a2(n10 - 1 + k, 1..2), k = 0..2 (rows M-3 .. M-1, 0-based)
    {
        int jk = n10 + 1;                      // 1-based as in fpclos
        for (int i = 1; i <= kk; ++i) {
            int ik = jk;
            for (int j = 1; j <= 3; ++j) {
                if (ik <= 0) break;
                const double v = (j == 1 ? a11 : j == 2 ? a12 : a13)[at(ik - 1, ld, b)];
                const int slot = ik - (n10 - 1);      // rows n10-1 .. n10+2 (1-based) -> 0..3
                if (slot >= 0 && slot < 3) { if (i == 1) init21[slot] = v; else init22[slot] = v; }
                else if (slot == 3) { if (i == 1) a21[at(ik - 1, ld, b)] = v; else a22[at(ik - 1, ld, b)] = v; }
                --ik;
            }
            ++jk;
        }
        // rows n10+1, n10+2 (1-based) hold the border block: written here; row n10+2 = M only gets a2(M, 2)
        a21[at(M - 1, ld, b)] = 0.0;
        a21[at(M - 2, ld, b)] = init21[2]; a22[at(M - 2, ld, b)] = init22[2];
        if (kk >= 2) { /* a22(M) was stored through slot == 3 above */ }
    }
    // ---- rows M-1 and M: the periodicity condition couples them to the first columns.  Both rows travel through the
    //      band together: row M-1 rotates against band row jj, then row M against the updated band row (exactly the
    //      order of two successive sweeps, since neither revisits a band row) ----------------------------------------
    for (int r = 0; r < 2; ++r) {
        const int it = M - 1 + r;              // 1-based row number
        const int l5 = it - 1;
        const double h[4] = {0.0, hp[r][0], hp[r][1], hp[r][2]};
        double h1[5] = {0.0, 0.0, 0.0, 0.0, 0.0}, h2[4] = {0.0, 0.0, 0.0, 0.0};
        wr[r].x = pp[r][0]; wr[r].y = pp[r][1];
        int j = l5 - n10;
        for (int i = 1; i <= 3; ++i) {
            ++j;
            int l0 = j;
            for (;;) {
                const int l1 = l0 - kk;
                if (l1 <= 0) { h2[l0] = h2[l0] + h[i]; break; }
                if (l1 <= n10) { h1[l1] = h[i]; break; }
                l0 = l1 - n10;
            }
        }
        wr[r].ha = h1[1]; wr[r].hb = h1[2]; wr[r].hc = h1[3]; wr[r].g1 = h2[1]; wr[r].g2 = h2[2];
    }
}

// Part 2, one-lane form: both periodic rows through band rows 1 .. n10 (host build; device when a line has one lane).
STO_HD void fitpack_sweep(const FitArgs& A, int b, FpWrapRow* wr, const double* init21, const double* init22) {
    const int M = A.M, ld = A.ld;
    const int n7 = M, kk = 2, n10 = n7 - kk;   // kk1 = 3
    double* const a11 = A.cp; double* const a12 = A.zx; double* const a13 = A.zy;   // a1(j, 1..3), rows 0-based
    double* const a21 = A.zz; double* const a22 = A.ze;                            // a2(j, 1..2)
    double* const z1 = A.cx; double* const z2 = A.cy;                              // right-hand sides, then c in place
    (void)n7; (void)a11; (void)a12; (void)a13; (void)a21; (void)a22; (void)z1; (void)z2; (void)n10;
    double na1[STO_FP_CHUNK], na2[STO_FP_CHUNK], na3[STO_FP_CHUNK], nzx[STO_FP_CHUNK], nzy[STO_FP_CHUNK];
#pragma unroll
    for (int c = 0; c < STO_FP_CHUNK; ++c) {
        const int jj = 1 + c;
        na1[c] = na2[c] = na3[c] = nzx[c] = nzy[c] = 0.0;
        if (jj <= n10) {
            na1[c] = a11[at(jj - 1, ld, b)]; na2[c] = a12[at(jj - 1, ld, b)]; na3[c] = a13[at(jj - 1, ld, b)];
            nzx[c] = z1[at(jj - 1, ld, b)]; nzy[c] = z2[at(jj - 1, ld, b)];
        }
    }
    for (int j0 = 1; j0 <= n10; j0 += STO_FP_CHUNK) {
        double ra1[STO_FP_CHUNK], ra2[STO_FP_CHUNK], ra3[STO_FP_CHUNK], rzx[STO_FP_CHUNK], rzy[STO_FP_CHUNK];
#pragma unroll
        for (int c = 0; c < STO_FP_CHUNK; ++c) { ra1[c] = na1[c]; ra2[c] = na2[c]; ra3[c] = na3[c]; rzx[c] = nzx[c]; rzy[c] = nzy[c]; }
#pragma unroll
        for (int c = 0; c < STO_FP_CHUNK; ++c) {
            const int jj = j0 + STO_FP_CHUNK + c;
            if (jj <= n10) {
                na1[c] = a11[at(jj - 1, ld, b)]; na2[c] = a12[at(jj - 1, ld, b)]; na3[c] = a13[at(jj - 1, ld, b)];
                nzx[c] = z1[at(jj - 1, ld, b)]; nzy[c] = z2[at(jj - 1, ld, b)];
            }
        }
#pragma unroll
        for (int c = 0; c < STO_FP_CHUNK; ++c) {
            const int jj = j0 + c;
            if (jj > n10) break;
            const int slot = jj - (n10 - 1);   // the two band rows that already carry a2 entries
            double b1 = (slot >= 0) ? init21[slot] : 0.0, b2 = (slot >= 0) ? init22[slot] : 0.0;
            fp_wrap_step(wr[0], jj, n10, ra1[c], ra2[c], ra3[c], b1, b2, rzx[c], rzy[c]);
            fp_wrap_step(wr[1], jj, n10, ra1[c], ra2[c], ra3[c], b1, b2, rzx[c], rzy[c]);
            a11[at(jj - 1, ld, b)] = ra1[c]; a12[at(jj - 1, ld, b)] = ra2[c]; a13[at(jj - 1, ld, b)] = ra3[c];
            z1[at(jj - 1, ld, b)] = rzx[c]; z2[at(jj - 1, ld, b)] = rzy[c];
            a21[at(jj - 1, ld, b)] = b1; a22[at(jj - 1, ld, b)] = b2;
        }
    }
}

// Part 3 (one lane): the border rows of a2, then fpbacp.
STO_HD void fitpack_finish(const FitArgs& A, int b, FpWrapRow* wr) {
    const int M = A.M, ld = A.ld;
    const int n7 = M, kk = 2, n10 = n7 - kk;   // kk1 = 3
    double* const a11 = A.cp; double* const a12 = A.zx; double* const a13 = A.zy;   // a1(j, 1..3), rows 0-based
    double* const a21 = A.zz; double* const a22 = A.ze;                            // a2(j, 1..2)
    double* const z1 = A.cx; double* const z2 = A.cy;                              // right-hand sides, then c in place
    (void)n7; (void)a11; (void)a12; (void)a13; (void)a21; (void)a22; (void)z1; (void)z2; (void)n10;
    for (int r = 0; r < 2; ++r) {              // rotation with rows n10+1 .. n7 of a2 (row M-1 completely, then row M)
        double gg[3] = {0.0, wr[r].g1, wr[r].g2};
        for (int jj = 1; jj <= kk; ++jj) {
            const int ij = n10 + jj;
            if (ij <= 0) continue;
            const double piv = gg[jj];
            if (piv == 0.0) continue;
            double* const col = (jj == 1) ? a21 : a22;
            double ww = col[at(ij - 1, ld, b)];
            const FpRot g = fp_givs(piv, ww);
            col[at(ij - 1, ld, b)] = ww;
            double zz1 = z1[at(ij - 1, ld, b)], zz2 = z2[at(ij - 1, ld, b)];
            fp_rota(g, wr[r].x, zz1); fp_rota(g, wr[r].y, zz2);
            z1[at(ij - 1, ld, b)] = zz1; z2[at(ij - 1, ld, b)] = zz2;
            if (jj == kk) break;
            double v = a22[at(ij - 1, ld, b)];
            fp_rota(g, gg[2], v);
            a22[at(ij - 1, ld, b)] = v;
        }
    }
    // ---- fpbacp: back substitution, c in place of z -----------------------------------------------------------------
    const int n = n7, n2 = n - kk;
    double cn1x, cn1y, cn2x, cn2y;            // c(n2+1), c(n2+2) = the two border unknowns
    {
        cn2x = z1[at(n - 1, ld, b)] / a22[at(n - 1, ld, b)];
        cn2y = z2[at(n - 1, ld, b)] / a22[at(n - 1, ld, b)];
        double sx = z1[at(n - 2, ld, b)], sy = z2[at(n - 2, ld, b)];
        sx = sx - cn2x * a22[at(n - 2, ld, b)];
        sy = sy - cn2y * a22[at(n - 2, ld, b)];
        cn1x = sx / a21[at(n - 2, ld, b)];
        cn1y = sy / a21[at(n - 2, ld, b)];
        z1[at(n - 1, ld, b)] = cn2x; z2[at(n - 1, ld, b)] = cn2y;
        z1[at(n - 2, ld, b)] = cn1x; z2[at(n - 2, ld, b)] = cn1y;
    }
    double px1 = 0.0, py1 = 0.0, px2 = 0.0, py2 = 0.0;   // c(i+1), c(i+2) of the band part
    double ma1[STO_FP_CHUNK], ma2[STO_FP_CHUNK], ma3[STO_FP_CHUNK], mb1[STO_FP_CHUNK], mb2[STO_FP_CHUNK],
           mzx[STO_FP_CHUNK], mzy[STO_FP_CHUNK];
#pragma unroll
    for (int c = 0; c < STO_FP_CHUNK; ++c) {
        const int i = n2 - c;
        ma1[c] = 1.0; ma2[c] = ma3[c] = mb1[c] = mb2[c] = mzx[c] = mzy[c] = 0.0;
        if (i >= 1) {
            ma1[c] = a11[at(i - 1, ld, b)]; ma2[c] = a12[at(i - 1, ld, b)]; ma3[c] = a13[at(i - 1, ld, b)];
            mb1[c] = a21[at(i - 1, ld, b)]; mb2[c] = a22[at(i - 1, ld, b)];
            mzx[c] = z1[at(i - 1, ld, b)]; mzy[c] = z2[at(i - 1, ld, b)];
        }
    }
    for (int i0 = n2; i0 >= 1; i0 -= STO_FP_CHUNK) {
        double ra1[STO_FP_CHUNK], ra2[STO_FP_CHUNK], ra3[STO_FP_CHUNK], rb1[STO_FP_CHUNK], rb2[STO_FP_CHUNK],
               rzx[STO_FP_CHUNK], rzy[STO_FP_CHUNK];
#pragma unroll
        for (int c = 0; c < STO_FP_CHUNK; ++c) {
            ra1[c] = ma1[c]; ra2[c] = ma2[c]; ra3[c] = ma3[c]; rb1[c] = mb1[c]; rb2[c] = mb2[c]; rzx[c] = mzx[c]; rzy[c] = mzy[c];
        }
#pragma unroll
        for (int c = 0; c < STO_FP_CHUNK; ++c) {
            const int i = i0 - STO_FP_CHUNK - c;
            if (i >= 1) {
                ma1[c] = a11[at(i - 1, ld, b)]; ma2[c] = a12[at(i - 1, ld, b)]; ma3[c] = a13[at(i - 1, ld, b)];
                mb1[c] = a21[at(i - 1, ld, b)]; mb2[c] = a22[at(i - 1, ld, b)];
                mzx[c] = z1[at(i - 1, ld, b)]; mzy[c] = z2[at(i - 1, ld, b)];
            }
        }
#pragma unroll
        for (int c = 0; c < STO_FP_CHUNK; ++c) {
            const int i = i0 - c;
            if (i < 1) break;
            double sx = rzx[c], sy = rzy[c];
            sx = sx - cn1x * rb1[c]; sx = sx - cn2x * rb2[c];
            sy = sy - cn1y * rb1[c]; sy = sy - cn2y * rb2[c];
            const int jb = n2 - i + 1;         // fpbacp's j: 1 for the last band row
            const int i1 = (jb <= kk) ? jb - 1 : kk;
            if (i1 >= 1) { sx = sx - px1 * ra2[c]; sy = sy - py1 * ra2[c]; }
            if (i1 >= 2) { sx = sx - px2 * ra3[c]; sy = sy - py2 * ra3[c]; }
#if defined(__CUDA_ARCH__) && !defined(STO_NO_FAST_FP64)
            {
                bool slow = false;
                const double qx = div_fast(sx, ra1[c], slow), qy = div_fast(sy, ra1[c], slow);
                if (slow) { sx = sx / ra1[c]; sy = sy / ra1[c]; } else { sx = qx; sy = qy; }
            }
#else
            sx = sx / ra1[c]; sy = sy / ra1[c];
#endif
            z1[at(i - 1, ld, b)] = sx; z2[at(i - 1, ld, b)] = sy;
            px2 = px1; py2 = py1; px1 = sx; py1 = sy;
        }
    }
    for (int i = 0; i < 3; ++i) {              // c(n7 + i) = c(i)
        z1[at(M + i, ld, b)] = z1[at(i, ld, b)];
        z2[at(M + i, ld, b)] = z2[at(i, ld, b)];
    }
}

#if defined(__CUDA_ARCH__)
// Part 2, two-lane form: lane g = 0 carries row M-1, lane g = 1 row M one band row behind; the band row travels from
// lane 0 to lane 1 by __shfl_sync after lane 0's rotation and is stored by lane 1.  Each lane performs exactly the
// rotations fitpack_sweep performs for its row, on the same operands, so the result is bit-identical; the chain is
// n10 + 1 rotations long instead of 2 n10.  Whole warps call this (lanes g >= 2 only take part in the shuffles).
STO_D void fitpack_sweep_pair(const FitArgs& A, int b, bool active, int g, int lane0, FpWrapRow& mine,
                              const double* init21, const double* init22) {
    const int M = A.M, ld = A.ld, n10 = M - 2;
    double* const a11 = A.cp; double* const a12 = A.zx; double* const a13 = A.zy;
    double* const a21 = A.zz; double* const a22 = A.ze;
    double* const z1 = A.cx; double* const z2 = A.cy;
    const bool lead = active && g == 0, follow = active && g == 1;
    double na1[STO_FP_CHUNK], na2[STO_FP_CHUNK], na3[STO_FP_CHUNK], nzx[STO_FP_CHUNK], nzy[STO_FP_CHUNK];
#pragma unroll
    for (int c = 0; c < STO_FP_CHUNK; ++c) {
        const int jj = 1 + c;
        na1[c] = na2[c] = na3[c] = nzx[c] = nzy[c] = 0.0;
        if (lead && jj <= n10) {
            na1[c] = a11[at(jj - 1, ld, b)]; na2[c] = a12[at(jj - 1, ld, b)]; na3[c] = a13[at(jj - 1, ld, b)];
            nzx[c] = z1[at(jj - 1, ld, b)]; nzy[c] = z2[at(jj - 1, ld, b)];
        }
    }
    double p1 = 0.0, p2 = 0.0, p3 = 0.0, pb1 = 0.0, pb2 = 0.0, pzx = 0.0, pzy = 0.0;   // band row handed to lane 1
    for (int j0 = 1; j0 <= n10 + 1; j0 += STO_FP_CHUNK) {
        double ra1[STO_FP_CHUNK], ra2[STO_FP_CHUNK], ra3[STO_FP_CHUNK], rzx[STO_FP_CHUNK], rzy[STO_FP_CHUNK];
#pragma unroll
        for (int c = 0; c < STO_FP_CHUNK; ++c) { ra1[c] = na1[c]; ra2[c] = na2[c]; ra3[c] = na3[c]; rzx[c] = nzx[c]; rzy[c] = nzy[c]; }
#pragma unroll
        for (int c = 0; c < STO_FP_CHUNK; ++c) {
            const int jj = j0 + STO_FP_CHUNK + c;
            if (lead && jj <= n10) {
                na1[c] = a11[at(jj - 1, ld, b)]; na2[c] = a12[at(jj - 1, ld, b)]; na3[c] = a13[at(jj - 1, ld, b)];
                nzx[c] = z1[at(jj - 1, ld, b)]; nzy[c] = z2[at(jj - 1, ld, b)];
            }
        }
#pragma unroll
        for (int c = 0; c < STO_FP_CHUNK; ++c) {
            const int step = j0 + c;
            if (step > n10 + 1) break;               // warp-uniform
            const int jj = step - g;                 // lane 0: band row `step`; lane 1: the row lane 0 did one step ago
            double r1, r2, r3, b1, b2, zx, zy;
            if (g == 0) {
                const int slot = jj - (n10 - 1);     // the two band rows that already carry a2 entries
                r1 = ra1[c]; r2 = ra2[c]; r3 = ra3[c]; zx = rzx[c]; zy = rzy[c];
                b1 = (slot >= 0 && slot < 3) ? init21[slot] : 0.0;
                b2 = (slot >= 0 && slot < 3) ? init22[slot] : 0.0;
            } else {
                r1 = p1; r2 = p2; r3 = p3; b1 = pb1; b2 = pb2; zx = pzx; zy = pzy;
            }
            const bool on = (lead || follow) && jj >= 1 && jj <= n10;
            if (on) fp_wrap_step(mine, jj, n10, r1, r2, r3, b1, b2, zx, zy);
            if (on && g == 1) {
                a11[at(jj - 1, ld, b)] = r1; a12[at(jj - 1, ld, b)] = r2; a13[at(jj - 1, ld, b)] = r3;
                z1[at(jj - 1, ld, b)] = zx; z2[at(jj - 1, ld, b)] = zy;
                a21[at(jj - 1, ld, b)] = b1; a22[at(jj - 1, ld, b)] = b2;
            }
            p1 = __shfl_sync(0xffffffffu, r1, lane0); p2 = __shfl_sync(0xffffffffu, r2, lane0);
            p3 = __shfl_sync(0xffffffffu, r3, lane0); pb1 = __shfl_sync(0xffffffffu, b1, lane0);
            pb2 = __shfl_sync(0xffffffffu, b2, lane0); pzx = __shfl_sync(0xffffffffu, zx, lane0);
            pzy = __shfl_sync(0xffffffffu, zy, lane0);
        }
    }
}

STO_D FpWrapRow fp_shfl_row(const FpWrapRow& r, int src) {
    FpWrapRow o;
    o.ha = __shfl_sync(0xffffffffu, r.ha, src); o.hb = __shfl_sync(0xffffffffu, r.hb, src);
    o.hc = __shfl_sync(0xffffffffu, r.hc, src); o.g1 = __shfl_sync(0xffffffffu, r.g1, src);
    o.g2 = __shfl_sync(0xffffffffu, r.g2, src); o.x = __shfl_sync(0xffffffffu, r.x, src);
    o.y = __shfl_sync(0xffffffffu, r.y, src);
    return o;
}

// The FITPACK solver for lane g of a group of G lanes (whole warps call this): one lane per line, except that the two
// periodic rows travel through the band on two lanes when the group has them.
STO_D void fit_solve_fitpack_lane(const FitArgs& A, int b, bool active, int g, int G, int lane0) {
    FpWrapRow wr[2];
    wr[0] = FpWrapRow{0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    wr[1] = wr[0];
    double init21[3] = {0.0, 0.0, 0.0}, init22[3] = {0.0, 0.0, 0.0};
    if (active && g == 0) fitpack_regular(A, b, wr, init21, init22);
    if (G >= 2) {
        const FpWrapRow second = fp_shfl_row(wr[1], lane0);          // row M goes to lane 1
        FpWrapRow mine = (g == 0) ? wr[0] : second;
        fitpack_sweep_pair(A, b, active, g, lane0, mine, init21, init22);
        const FpWrapRow back = fp_shfl_row(mine, lane0 + 1);         // ... and comes back for the border rows
        if (g == 0) { wr[0] = mine; wr[1] = back; }
        __syncwarp();                                                // lane 1's band rows are visible to lane 0
    } else if (active) {
        fitpack_sweep(A, b, wr, init21, init22);
    }
    if (active && g == 0) fitpack_finish(A, b, wr);
}
#endif

// The FITPACK solver played by one caller (host build: the reference for what the lanes do).
STO_HD void fit_solve_fitpack(const FitArgs& A, int b) {
    FpWrapRow wr[2];
    double init21[3], init22[3];
    fitpack_regular(A, b, wr, init21, init22);
    fitpack_sweep(A, b, wr, init21, init22);
    fitpack_finish(A, b, wr);
}

STO_HD void fit_degenerate(const FitArgs& A, int b) {  // FITPACK returns ier=10; scipy raises
    const int M = A.M, ld = A.ld;
    if (A.status) A.status[b] |= STO_CAND_DEGENERATE_FIT;
    const double qnan = nan("");
    for (int i = 0; i <= M; ++i) A.u[at(i, ld, b)] = qnan;
    for (int i = 0; i < M + 3; ++i) { A.cx[at(i, ld, b)] = qnan; A.cy[at(i, ld, b)] = qnan; }
}

// One lane of a group of G (device).  Lanes 0..2 run the three recurrences when G >= 3.
// PART_E = 0: Thomas recurrences; PART_E > 0: partitioned solve with PART_E blocks per lane (device only);
// PART_E < 0: FITPACK's Givens sweep (one lane per line; the other lanes of the group help with the row phases).
template <int PART_E = 0>
STO_HD void fit_candidate_lane(const FitArgs& A, int b, bool active, int g, int G, int lane0) {
    (void)lane0;
    if (active) fit_phase_segments(A, b, g, G);
    STO_FIT_SYNC();
    double total = 0.0;
#if defined(__CUDA_ARCH__)
    if (G > 1) total = fit_phase_cumsum_group(A, b, active, g, G, lane0);
    else if (active) total = fit_phase_cumsum(A, b);
#else
    if (active && g == 0) total = fit_phase_cumsum(A, b);
#endif
    const bool ok = total > 0.0;
    STO_FIT_SYNC();
    if (active && !ok) { if (g == 0) fit_degenerate(A, b); }
    if (active && ok) fit_phase_normalise(A, b, g, G, total);
    STO_FIT_SYNC();
    if (PART_E < 0) {
        if (active && ok) fit_phase_bspl_rows(A, b, g, G);
        STO_FIT_SYNC();
#if defined(__CUDA_ARCH__)
        fit_solve_fitpack_lane(A, b, active && ok, g, G, lane0);
#else
        if (active && ok && g == 0) fit_solve_fitpack(A, b);
#endif
        return;
    }
    if (active && ok) fit_phase_rows(A, b, g, G);
    STO_FIT_SYNC();
#if defined(__CUDA_ARCH__)
    if (PART_E > 0) {   // the G lanes solve the cyclic system together
        fit_solve_partitioned_group<(PART_E > 0) ? PART_E : 1>(A, b, active && ok, g, lane0);
        return;
    }
#endif
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
STO_HD void fit_candidate(const FitArgs& A, int b, int G = 1, bool part = false, bool fitpack = false) {
    for (int g = 0; g < G; ++g) fit_phase_segments(A, b, g, G);
    const double total = fit_phase_cumsum(A, b);
    if (!(total > 0.0)) { fit_degenerate(A, b); return; }
    for (int g = 0; g < G; ++g) fit_phase_normalise(A, b, g, G, total);
    if (fitpack) {
        for (int g = 0; g < G; ++g) fit_phase_bspl_rows(A, b, g, G);
        fit_solve_fitpack(A, b);
        return;
    }
    for (int g = 0; g < G; ++g) fit_phase_rows(A, b, g, G);
    if (part) { fit_solve_partitioned_emulated(A, b, fit_part_blocks(A.M)); return; }
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
