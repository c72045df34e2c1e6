/* sto_oracle.h -- declarations for the CPU restatement (TEST INFRASTRUCTURE ONLY; see sto_oracle.c). */
#ifndef STO_ORACLE_H
#define STO_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define STO_ORACLE_MAX_BREAKS 32
#define STO_ORACLE_ZERO_SPEED 1 /* reference raises FloatingPointError (simulator.py:164-165) */
#define STO_ORACLE_NAN 2
#define STO_ORACLE_NO_CONVERGENCE 3 /* safety cap of 64*N outer iterations, never hit on real data */

/* models/vehicle.py:6-24: the six scalars + the two PPoly tables SciPy's CubicSpline produced
 * (x = breakpoints, c = PPoly.c with shape [4][n-1]).  Same field order as sto_vehicle_f64. */
typedef struct sto_oracle_vehicle {
    double max_lon_acc, max_lon_dcc, max_left_acc, max_right_acc, max_speed, max_jerk;
    int32_t n_acc, n_dcc;
    double acc_x[STO_ORACLE_MAX_BREAKS];
    double acc_c[4][STO_ORACLE_MAX_BREAKS - 1];
    double dcc_x[STO_ORACLE_MAX_BREAKS];
    double dcc_c[4][STO_ORACLE_MAX_BREAKS - 1];
} sto_oracle_vehicle;

double sto_oracle_ppoly(const double* x, const double* c, int n_break, int order, double v);
double sto_oracle_maxlat(const sto_oracle_vehicle* V, double lon, int ref_pow);
/* solver of the interpolation system: 2 (default) FITPACK's fpclos Givens sweep (the reference's arithmetic, bit for bit),
 * 1 block elimination + cyclic reduction, 0 Thomas + Sherman-Morrison */
void sto_oracle_set_fit_solver(int solver);
int sto_oracle_fit_periodic_cubic(const double* px, const double* py, int M, double* t, double* cx,
                                  double* cy);
void sto_oracle_bspline_eval(const double* t, int nt, const double* c, int k, const double* xs, int nx,
                             int der, double* out);
int sto_oracle_sample(const double* t, int nt, const double* cx, const double* cy, int k, const double* ts,
                      int N, double* X, double* Y, double* YAW, double* R, int ref_pow);
int sto_oracle_qss(const double* X, const double* Y, const double* R, const double* sinb, int N,
                   const sto_oracle_vehicle* V, double* v, double* a, double* lat, double* flag,
                   double* tseg, double* lap, int64_t* stats, int ref_pow);
/* fast-mode checker (not the reference's algorithm; see sto_oracle.c) */
int sto_oracle_qss_sweep(const double* X, const double* Y, const double* R, const double* sinb, int N,
                         const sto_oracle_vehicle* V, int rounds, double* v, double* a, double* lat, double* tseg,
                         double* lap, int ref_pow);
int sto_oracle_lap_from_offsets(const double* centre_x, const double* centre_y, const double* normal_x,
                                const double* normal_y, const double* offsets, int M, const double* ts,
                                const double* sinb, int N, const sto_oracle_vehicle* V, double* lap,
                                int ref_pow);
int sto_oracle_lap_batch(const double* centre_x, const double* centre_y, const double* normal_x,
                         const double* normal_y, const double* offsets, int M, int B, const double* ts,
                         const double* sinb, int N, const sto_oracle_vehicle* V, double* lap,
                         int32_t* status, int n_threads, int ref_pow);
#ifdef __cplusplus
}
#endif
#endif
