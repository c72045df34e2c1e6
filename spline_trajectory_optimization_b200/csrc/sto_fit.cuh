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

STO_HD void fit_candidate(const FitArgs& A, int b) {
    const int M = A.M, ld = A.ld;
    // pass 1: cumulative chord length (FITPACK clocur: dist = sum_dim (x_i - x_{i-1})**2)
    double x0, y0;
    fit_point(A, 0, b, x0, y0);
    double xp = x0, yp = y0, acc = 0.0;
    A.u[at(0, ld, b)] = 0.0;
    // Rows are processed in chunks of STO_FIT_CHUNK with all loads of a chunk issued before its (serially
    // dependent) arithmetic: one exposed memory round trip per chunk instead of one per row.
    for (int i0 = 1; i0 <= M; i0 += STO_FIT_CHUNK) {
        double xs[STO_FIT_CHUNK], ys[STO_FIT_CHUNK];
#pragma unroll
        for (int k = 0; k < STO_FIT_CHUNK; ++k) {
            const int i = i0 + k;
            if (i < M) fit_point(A, i, b, xs[k], ys[k]); else { xs[k] = x0; ys[k] = y0; }  // closed loop (trajectory.py:217-218)
        }
#pragma unroll
        for (int k = 0; k < STO_FIT_CHUNK; ++k) {
            const int i = i0 + k;
            if (i <= M) {
                double dx = xs[k] - xp, dy = ys[k] - yp;
                double dist = 0.0;
                dist = dist + dx * dx;
                dist = dist + dy * dy;
                acc = acc + sqrt(dist);
                A.u[at(i, ld, b)] = acc;
                xp = xs[k]; yp = ys[k];
            }
        }
    }
    const double total = acc;
    if (!(total > 0.0)) {  // FITPACK returns ier=10; scipy raises
        if (A.status) A.status[b] |= STO_CAND_DEGENERATE_FIT;
        const double qnan = nan("");
        for (int i = 0; i <= M; ++i) A.u[at(i, ld, b)] = qnan;
        for (int i = 0; i < M + 3; ++i) { A.cx[at(i, ld, b)] = qnan; A.cy[at(i, ld, b)] = qnan; }
        return;
    }
    // pass 2: normalise
    for (int i = 1; i <= M; ++i) A.u[at(i, ld, b)] = A.u[at(i, ld, b)] / total;
    A.u[at(M, ld, b)] = 1.0;

    // pass 3: collocation rows b0*e[j-1] + b1*e[j] + b2*e[j+1] = p_j (e_i = c[(i+1) mod M]) and Thomas
    // forward elimination of T = A - w v^T (Sherman-Morrison split of the two corner entries).
    double t0 = fit_knot(A, -2, b), t1 = fit_knot(A, -1, b), t2 = fit_knot(A, 0, b), t3 = fit_knot(A, 1, b),
           t4 = fit_knot(A, 2, b);
    double beta = 0.0, gamma = 0.0, alpha = 0.0;
    double cpp = 0.0, zxp = 0.0, zyp = 0.0, zzp = 0.0;
    double zx_last = 0.0, zy_last = 0.0, zz_last = 0.0;
    for (int j0 = 0; j0 < M; j0 += STO_FIT_CHUNK) {
        double pxs[STO_FIT_CHUNK], pys[STO_FIT_CHUNK], tn[STO_FIT_CHUNK];
#pragma unroll
        for (int k = 0; k < STO_FIT_CHUNK; ++k) {
            const int j = j0 + k;
            if (j < M) { fit_point(A, j, b, pxs[k], pys[k]); tn[k] = fit_knot(A, j + 3, b); }
            else { pxs[k] = pys[k] = tn[k] = 0.0; }
        }
#pragma unroll
        for (int k = 0; k < STO_FIT_CHUNK; ++k) {
            const int j = j0 + k;
            if (j >= M) break;
            double b0 = ((t3 - t2) * (t3 - t2)) / ((t3 - t0) * (t3 - t1));
            double b2 = ((t2 - t1) * (t2 - t1)) / ((t4 - t1) * (t3 - t1));
            double b1 = (1.0 - b0) - b2;
            const double pxj = pxs[k], pyj = pys[k];
            double den, ncp, nzx, nzy, nzz;
            if (j == 0) {
                beta = b0;       // corner: row 0, column M-1
                gamma = -b1;
                den = b1 - gamma;
                ncp = b2 / den;
                nzx = pxj / den;
                nzy = pyj / den;
                nzz = gamma / den;
            } else {
                double dj = b1, wj = 0.0;
                if (j == M - 1) {
                    alpha = b2;  // corner: row M-1, column 0
                    dj = b1 - alpha * beta / gamma;
                    wj = alpha;
                }
                den = dj - b0 * cpp;
                ncp = b2 / den;
                nzx = (pxj - b0 * zxp) / den;
                nzy = (pyj - b0 * zyp) / den;
                nzz = (wj - b0 * zzp) / den;
            }
            A.cp[at(j, ld, b)] = ncp;
            A.zx[at(j, ld, b)] = nzx;
            A.zy[at(j, ld, b)] = nzy;
            A.zz[at(j, ld, b)] = nzz;
            cpp = ncp; zxp = nzx; zyp = nzy; zzp = nzz;
            t0 = t1; t1 = t2; t2 = t3; t3 = t4;
            t4 = tn[k];
        }
    }
    zx_last = zxp; zy_last = zyp; zz_last = zzp;
    // pass 4: back substitution
    double zxn = zx_last, zyn = zy_last, zzn = zz_last;
    for (int j0 = M - 2; j0 >= 0; j0 -= STO_FIT_CHUNK) {
        double cs[STO_FIT_CHUNK], xs[STO_FIT_CHUNK], ys[STO_FIT_CHUNK], zs[STO_FIT_CHUNK];
#pragma unroll
        for (int k = 0; k < STO_FIT_CHUNK; ++k) {
            const int j = j0 - k;
            if (j >= 0) {
                cs[k] = A.cp[at(j, ld, b)]; xs[k] = A.zx[at(j, ld, b)];
                ys[k] = A.zy[at(j, ld, b)]; zs[k] = A.zz[at(j, ld, b)];
            } else { cs[k] = xs[k] = ys[k] = zs[k] = 0.0; }
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
    // pass 5: Sherman-Morrison correction, e -> c with the one-slot rotation and the periodic wrap
    double denom = (1.0 + zzn) + beta * zz_last / gamma;
    double fx = (zxn + beta * zx_last / gamma) / denom;
    double fy = (zyn + beta * zy_last / gamma) / denom;
    for (int i = 0; i < M; ++i) {
        double zzi = A.zz[at(i, ld, b)];
        double ex = A.zx[at(i, ld, b)] - fx * zzi;
        double ey = A.zy[at(i, ld, b)] - fy * zzi;
        int ci = (i + 1 == M) ? 0 : i + 1;
        A.cx[at(ci, ld, b)] = ex;
        A.cy[at(ci, ld, b)] = ey;
        if (ci < 3) {
            A.cx[at(M + ci, ld, b)] = ex;
            A.cy[at(M + ci, ld, b)] = ey;
        }
    }
}

}  // namespace sto
