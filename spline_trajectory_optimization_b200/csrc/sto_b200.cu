// sto_b200.cu -- __global__ wrappers and the extern "C" boundary of libsto_b200.so (see include/sto_b200.h).
//
// Build (sm_100a only, no FMA contraction so every operation rounds like the reference's NumPy/Python):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false -shared -Xcompiler -fPIC ...
#include <cuda_runtime.h>

#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>

#include "sto_common.cuh"
#include "sto_eval.cuh"
#include "sto_fast.cuh"
#include "sto_fit.cuh"
#include "sto_fit_lsq.cuh"
#include "sto_qss.cuh"
#include "sto_qss_memo.cuh"
#include "sto_qss_memo2.cuh"
#include "sto_qss_memo3.cuh"

namespace {

thread_local std::string g_err;
thread_local const char* g_last_qss_kernel = "";   // which QSS kernel the last launch of this thread used (sto_last_qss_kernel)

int fail(int code, const char* what) {
    g_err = what;
    return code;
}
int fail_cuda(cudaError_t e, const char* where) {
    g_err = std::string(where) + ": " + cudaGetErrorString(e);
    return STO_ERR_CUDA;
}
#define STO_CUDA(call)                                        \
    do {                                                      \
        cudaError_t e_ = (call);                              \
        if (e_ != cudaSuccess) return fail_cuda(e_, #call);   \
    } while (0)

// Developer tuning overrides (0 = automatic).  Read from the STO_* environment variables ONCE, when the library is
// loaded, and settable afterwards through sto_set_tuning(); the launch paths only read these atomics.
struct Tuning {
    std::atomic<int> fit_split{0};    // lanes per candidate of the fit kernels (power of two <= 32)
    std::atomic<int> qss_lanes{0};    // candidates per warp of the QSS kernels (power of two <= 32)
    std::atomic<int> qss_group{0};    // lanes per candidate of the memoised QSS kernel (1, 2, 4, 8, 16, 32)
    std::atomic<int> qss_planes{0};   // 1 = bit planes in shared memory, 2 = global (mixed decided by residency),
                                      // 3 = all global, 4 = CONT planes shared + live / STOP planes global
    std::atomic<int> qss_kernel{0};   // small batches: 1 = the four-walker kernel (sto_qss_memo.cuh), 2 = the one-loop kernel
                                      // (sto_qss_memo2.cuh), 3 = the one-loop kernel with the re-spawned lists walked out of
                                      // order (sto_qss_memo3.cuh), 4 = the one-loop kernel with the forward original-row
                                      // sub-pass run-parallel; 0 = automatic (N <= 4096: see launch_qss)
    std::atomic<int> fit_solver{STO_FIT_FITPACK};   // sto_set_fit_solver
};
Tuning g_tune;
inline bool pow2_le32(int v) { return v >= 1 && v <= 32 && (v & (v - 1)) == 0; }
struct TuningFromEnv {
    TuningFromEnv() {
        if (const char* e = getenv("STO_FIT_SPLIT")) { const int v = atoi(e); if (pow2_le32(v)) g_tune.fit_split = v; }
        if (const char* e = getenv("STO_FIT_SOLVER")) { const int v = atoi(e); if (v >= 0 && v <= 2) g_tune.fit_solver = v; }
        if (const char* e = getenv("STO_QSS_LANES")) { const int v = atoi(e); if (v >= 1 && v <= 32) g_tune.qss_lanes = v; }
        if (const char* e = getenv("STO_QSS_GROUP")) { const int v = atoi(e); if (pow2_le32(v)) g_tune.qss_group = v; }
        if (const char* e = getenv("STO_QSS_KERNEL")) { const int v = atoi(e); if (v >= 0 && v <= 4) g_tune.qss_kernel = v; }
        if (const char* e = getenv("STO_QSS_PLANES"))
            g_tune.qss_planes = (e[0] == 's') ? 1 : (e[0] != 'g') ? 0 : (e[1] == '\0') ? 2 : (e[1] == '0') ? 3 : 4;
    }
} g_tuning_from_env;

// 32-thread CTAs until every SM has a few warps: the QSS is latency-bound, so spreading warps over all 148
// SMs matters more than CTA size; larger batches use fuller CTAs to cut launch/scheduling overhead.
int pick_block(int B) {
    if (B <= 148 * 32 * 8) return 32;
    if (B <= 148 * 64 * 16) return 64;
    return 128;
}
inline int grid_for(int B, int block) { return (B + block - 1) / block; }
// Lanes per candidate for the fit: 4 (three recurrences + split row work) while the batch leaves SMs idle.
int pick_fit_split(int B) {
    if (const int v = g_tune.fit_split.load()) return v;
    // measured on B200 (Monza, M = 2895): 4,096 candidates 10.2 / 7.6 / 5.5 / 4.5 / 4.4 ms at 1 / 2 / 4 / 8 / 16 lanes;
    // 32,768 candidates 12.1 / 9.5 / 12.8 / 20.0 ms at 1 / 2 / 4 / 8 lanes
    if ((long long)B * 8 <= 148LL * 8 * 32) return 8;
    if (B <= 16384) return 4;
    if (B <= 65536) return 2;
    return 1;
}
// Solver of the fit's interpolation system (include/sto_b200.h, sto_set_fit_solver):
//   STO_FIT_FITPACK (default)  FITPACK's own Givens sweep, restated operation by operation (sto_fit.cuh): coefficients
//                              bit-identical to the reference's scipy splprep, one dependent chain per line;
//   STO_FIT_BLOCKS             a group of lanes shares each line: 32-block elimination + PCR over warp shuffles, one
//                              division per row; coefficients equal up to rounding (1e-15);
//   STO_FIT_THOMAS             one-lane Thomas + Sherman-Morrison recurrences (same accuracy class as BLOCKS).
// Lanes per candidate of the BLOCKS solver measured on B200 (tools/fit_part_bench.py, ms at M = 2895 | 23,160):
//   B =   256: Thomas x8 2.7 | 30.7   blocks x32 0.27 |  2.7
//   B =  1024: Thomas x8 3.6 | 32.8   blocks x16 0.64 |  7.7   (x32 0.58 | 11.3)
//   B =  4096: Thomas x8 4.2 | 34.3   blocks x8  1.81 | 16.1
//   B = 16384: Thomas x1 8.8 | 71.2   blocks x4  3.60 | 29.8
//   B = 32768: Thomas x2 8.7          blocks x2  5.9          B = 65536: Thomas x1 11.4, blocks x1 9.2
struct FitPlan { int split; int solver; };
FitPlan pick_fit_plan(int M, int B) {
    const int solver = g_tune.fit_solver.load();
    if (solver == STO_FIT_THOMAS) return FitPlan{pick_fit_split(B), solver};
    if (solver == STO_FIT_FITPACK) {
        // the rotation chain runs on one lane (two for the periodic rows); the other lanes of a group only speed up the
        // row phases (segments, running sum, fpbspl values).  Measured (M = 2895): B = 64: 5.1 / 5.4 / 5.7 ms at
        // 32 / 16 / 8 lanes; B = 4,096: 7.1 ms at 8 lanes, 12.7 at 16
        int lanes = (B <= 256) ? 32 : (B <= 1024) ? 16 : pick_fit_split(B);
        if (const int v = g_tune.fit_split.load()) lanes = v;
        return FitPlan{lanes, solver};
    }
    int lanes = (B <= 512) ? 32 : (B <= 2048) ? 16 : (B <= 8192) ? 8 : (B <= 24576) ? 4 : (B <= 49152) ? 2 : 1;
    if (const int v = g_tune.fit_split.load()) lanes = v;
    if (M < 256) lanes = 1;   // one block (sto::fit_part_blocks): nothing to share
    return FitPlan{lanes, STO_FIT_BLOCKS};
}

// Lanes per candidate for kernels whose samples are independent: fill ~8 warps per SM before going one-per-candidate.
int pick_split(int B, int N) {
    int split = 1;
    while (split < 32 && (long long)B * split < 148LL * 8 * 32 && N / (split * 2) >= 64) split <<= 1;
    return split;
}

// ---- workspace carving ---------------------------------------------------------------------------------
struct Carver {
    char* base;
    size_t off = 0;
    explicit Carver(void* p) : base(static_cast<char*>(p)) {}
    template <class T>
    T* take(size_t n) {
        off = (off + 255) & ~size_t(255);
        T* p = base ? reinterpret_cast<T*>(base + off) : nullptr;
        off += n * sizeof(T);
        return p;
    }
    size_t bytes() const { return (off + 255) & ~size_t(255); }
};

inline size_t ldof(int B) { return (size_t)((B + 31) & ~31); }  // internal leading dimension

struct FitWork { double *cp, *zx, *zy, *zz, *ze; };
FitWork carve_fit(Carver& c, int M, size_t ld) {
    FitWork w;
    w.cp = c.take<double>((size_t)M * ld);
    w.zx = c.take<double>((size_t)M * ld);
    w.zy = c.take<double>((size_t)M * ld);
    w.zz = c.take<double>((size_t)M * ld);
    w.ze = c.take<double>((size_t)M * ld);
    return w;
}

struct QssWork {
    double *dd, *df, *v, *a, *rec;
    uint8_t *rowflag, *sp_flag;
    int32_t *sp_ent, *sp_ext, *sp_turn;
    sto::MemoWork memo;
    int cap;
};
constexpr size_t kMemoSmemBudget = 200 * 1024;  // of the 227 KB a CTA may use on sm_100a
inline int effective_impl(int N, int impl) {
    // the memoised kernel keeps six bit planes per candidate in shared memory: 48 * ceil(N/64) bytes
    return (impl == STO_QSS_MEMO && N >= STO_QSS_MEMO_MIN_N && sto::memo_smem_bytes(N, 1) <= kMemoSmemBudget)
               ? STO_QSS_MEMO : STO_QSS_PLAIN;
}
QssWork carve_qss(Carver& c, int N, size_t ld, int impl, bool need_chords, bool need_state) {
    QssWork w{};
    impl = effective_impl(N, impl);
    w.cap = 2 * N + 64;
    if (need_chords) {
        w.dd = c.take<double>((size_t)N * ld);
        w.df = c.take<double>((size_t)N * ld);
    }
    if (need_state && impl == STO_QSS_PLAIN) {
        w.v = c.take<double>((size_t)N * ld);
        w.a = c.take<double>((size_t)N * ld);
    }
    if (impl == STO_QSS_MEMO) w.rec = c.take<double>((size_t)N * ld * 4);
    if (impl == STO_QSS_PLAIN) {
        w.rowflag = c.take<uint8_t>((size_t)N * ld);
        w.sp_flag = c.take<uint8_t>((size_t)w.cap * ld);
        w.sp_ent = c.take<int32_t>((size_t)w.cap * ld);
        w.sp_ext = c.take<int32_t>((size_t)w.cap * ld);
        w.sp_turn = c.take<int32_t>((size_t)w.cap * ld);
    } else {
        w.memo = sto::carve_memo([&](size_t nbytes) { return (void*)c.take<char>(nbytes); }, N, ld, w.cap);
    }
    return w;
}

// ---- kernels ---------------------------------------------------------------------------------------------
// `split` (a power of two <= 32) lanes share a candidate: per-row work is split over them and the three Thomas
// recurrences run side by side (sto_fit.cuh).  Whole warps stay alive so the group syncs are warp-uniform.
// PART_E > 0: the `split` lanes solve the cyclic system together (partitioned elimination + PCR over warp shuffles),
// PART_E blocks per lane; PART_E = 0: Thomas recurrences.
template <int PART_E>
__global__ void fit_kernel(sto::FitArgs A, int split) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int b = t / split, g = t % split;
    const int lane0 = (threadIdx.x & 31) & ~(split - 1);
    sto::fit_candidate_lane<PART_E>(A, b < A.B ? b : A.B - 1, b < A.B, g, split, lane0);
}

void launch_fit(const sto::FitArgs& A, const FitPlan& plan, cudaStream_t st) {
    const int split = plan.split;
    const int block = pick_block(A.B * split);
    const int grid = grid_for(A.B * split, block);
    if (plan.solver == STO_FIT_THOMAS) { fit_kernel<0><<<grid, block, 0, st>>>(A, split); return; }
    if (plan.solver == STO_FIT_FITPACK) { fit_kernel<-1><<<grid, block, 0, st>>>(A, split); return; }
    switch (32 / split) {   // blocks per lane (one block in all when M < 256: plan.split == 1, any instantiation)
        case 1: fit_kernel<1><<<grid, block, 0, st>>>(A, split); break;
        case 2: fit_kernel<2><<<grid, block, 0, st>>>(A, split); break;
        case 4: fit_kernel<4><<<grid, block, 0, st>>>(A, split); break;
        case 8: fit_kernel<8><<<grid, block, 0, st>>>(A, split); break;
        case 16: fit_kernel<16><<<grid, block, 0, st>>>(A, split); break;
        default: fit_kernel<32><<<grid, block, 0, st>>>(A, split); break;
    }
}

// Least-squares closed spline on fixed knots (sto_fit_lsq.cuh), one candidate per thread, coalesced sample-major.
__global__ void fit_lsq_kernel(sto::LsqArgs A) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < A.F.B) sto::lsq_candidate_k(A, b);
}

// `split` lanes share a candidate, each taking a contiguous slice of its samples (samples are independent): small
// batches get split x more threads, which is what a latency-bound kernel needs.
__global__ void eval_kernel(sto::EvalArgs A, int split) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int b = t / split, g = t % split;
    if (b >= A.B) return;
    const int per = (A.N + split - 1) / split;
    const int j0 = g * per, j1 = (j0 + per < A.N) ? j0 + per : A.N;
    sto::eval_range(A, b, j0, j1);
}

__global__ void eval_spline_kernel(sto::SplineEvalArgs A) {
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < A.N) sto::eval_spline_sample(A, j);
}

__global__ void eval_spline_batch_kernel(sto::SplineBatchArgs A) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= A.B) return;
    for (int j = blockIdx.y; j < A.N; j += gridDim.y) sto::eval_spline_batch_sample(A, j, b);
}

// Trajectory.fill_bounds (reference models/trajectory.py:83-141): nearest intersection of the two-sided segment
// P -/+ max_dist * n with a closed polyline; no hit -> the point itself.  One thread per point, edges streamed from L1/L2.
__global__ void fill_bounds_kernel(const double* __restrict__ px, const double* __restrict__ py,
                                   const double* __restrict__ nx, const double* __restrict__ ny, int n,
                                   const double* __restrict__ ring, int m, double max_dist, double* bx, double* by,
                                   int32_t* found) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double Px = px[i], Py = py[i], dx = nx[i], dy = ny[i];
    double best = INFINITY, tbest = 0.0;
    for (int j = 0; j < m; ++j) {
        const int jn = (j + 1 == m) ? 0 : j + 1;
        const double ax = ring[2 * j], ay = ring[2 * j + 1];
        const double ex = ring[2 * jn] - ax, ey = ring[2 * jn + 1] - ay;
        const double wx = ax - Px, wy = ay - Py;
        const double den = dx * ey - dy * ex;
        if (den == 0.0) continue;
        const double t = (wx * ey - wy * ex) / den;   // along the normal
        const double s = (wx * dy - wy * dx) / den;   // along the edge
        if (s >= 0.0 && s <= 1.0 && fabs(t) <= max_dist && fabs(t) < best) { best = fabs(t); tbest = t; }
    }
    const bool hit = best < INFINITY;
    bx[i] = hit ? Px + tbest * dx : Px;
    by[i] = hit ? Py + tbest * dy : Py;
    if (found) found[i] = hit ? 1 : 0;
}

__global__ void arc_sections_kernel(sto::ArcArgs A) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < A.N) sto::arc_section(A, i);
}

// dd / df from explicit x, y columns (the sto_qss_f64 entry); one thread per (sample, candidate)
__global__ void chord_kernel(const double* x, const double* y, int N, int B, int ld_in, double* dd, double* df,
                             int ld) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    for (int i = blockIdx.y; i < N; i += gridDim.y) {   // grid.y is capped at 65,535: longer tracks stride
        int n = (i + 1 == N) ? 0 : i + 1;
        double x0 = x[sto::at(i, ld_in, b)], y0 = y[sto::at(i, ld_in, b)];
        double x1 = x[sto::at(n, ld_in, b)], y1 = y[sto::at(n, ld_in, b)];
        dd[sto::at(i, ld, b)] = sto::chord_qss(x0, y0, x1, y1);
        df[sto::at(i, ld, b)] = sto::chord_norm(x0, y0, x1, y1);
    }
}

// The QSS kernels are latency bound (one dependent chain of FP64 divisions / square roots and scattered loads per
// candidate), so what matters is how many warps each SM can interleave.  `lanes` candidates share a warp
// (lanes = 32 for big batches); small batches are spread over more warps by leaving lanes idle.
__device__ __forceinline__ int candidate_of_thread(int lanes, int B, bool& active) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int warp = t >> 5, lane = t & 31;
    const int b = warp * lanes + lane;
    active = lane < lanes && b < B;
    return active ? b : B - 1;
}

template <bool OWNER>
__global__ void qss_plain_kernel(sto::QssArgs A, int lanes, const __grid_constant__ sto_vehicle_f64 V) {
    bool active;
    const int b = candidate_of_thread(lanes, A.B, active);
    sto::qss_plain_candidate<OWNER>(A, V, b, active);
}

extern __shared__ unsigned long long sto_planes[];  // [6 planes][W words][cpw candidates], then the per-lane rings

// One warp per CTA hosting `cpw` candidates, each run by a group of G lanes (see sto_qss_memo.cuh, "lane groups").
template <int G>
__global__ void qss_memo_kernel(sto::QssArgs A, sto::MemoWork W, int cpw, const __grid_constant__ sto_vehicle_f64 V) {
    const int lane = threadIdx.x & 31, warp = blockIdx.x;
    const int grp = lane / G, g = lane % G;
    const int b = warp * cpw + grp;
    const bool active = grp < cpw && b < A.B;
    int32_t* ring = reinterpret_cast<int32_t*>(sto_planes + (size_t)6 * W.W * cpw);
    const sto::MemoCtx C = sto::memo_bind(sto_planes, cpw, grp < cpw ? grp : 0, ring, 32, lane, A.N, W.W);
    sto::qss_memo_candidate<G>(A, W, C, V, active ? b : A.B - 1, active, g, grp * G);
}

// ---- fast mode (sto_fast.cuh): one kernel, one candidate per thread ------------------------------------------------
// STAGE: the six track tables are copied once per CTA into shared memory by the bulk-copy engine (cp.async.bulk with an
// mbarrier transaction count - the 1-D TMA path; SASS UBLKCP) and every lane reads them from there as broadcasts.
struct FastTables { int n_cen, n_ts; bool have_sinb; };
extern __shared__ __align__(16) double sto_fast_smem[];
template <bool STAGE>
__global__ void fast_kernel(sto::FastArgs Ain, FastTables T, const __grid_constant__ sto_vehicle_f64 V) {
    sto::FastArgs A = Ain;
    if (STAGE) {
        __shared__ __align__(8) unsigned long long bar;
        const int pc = (T.n_cen + 1) & ~1, pt = (T.n_ts + 1) & ~1;        // table strides: multiples of 16 bytes
        double* s_cenx = sto_fast_smem;
        double* s_ceny = s_cenx + pc;
        double* s_nrmx = s_ceny + pc;
        double* s_nrmy = s_nrmx + pc;
        double* s_ts = s_nrmy + pc;
        double* s_sb = s_ts + pt;
        const unsigned bar_s = (unsigned)__cvta_generic_to_shared(&bar);
        if (threadIdx.x == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_s) : "memory");
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            const unsigned bc = (unsigned)(T.n_cen / 2) * 16u, bt = (unsigned)(T.n_ts / 2) * 16u;   // whole 16-byte units
            const unsigned total = 4u * bc + bt + (T.have_sinb ? bt : 0u);
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_s), "r"(total) : "memory");
            const double* src[6] = {Ain.F.cenx, Ain.F.ceny, Ain.F.nrmx, Ain.F.nrmy, Ain.E.ts, Ain.Q.sinb};
            double* dst[6] = {s_cenx, s_ceny, s_nrmx, s_nrmy, s_ts, s_sb};
#pragma unroll
            for (int k = 0; k < 6; ++k) {
                const unsigned nb = (k < 4) ? bc : bt;
                if (nb == 0u || (k == 5 && !T.have_sinb)) continue;
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"((unsigned)__cvta_generic_to_shared(dst[k])), "l"(src[k]), "r"(nb), "r"(bar_s) : "memory");
                const int n = (k < 4) ? T.n_cen : T.n_ts;
                if (n & 1) dst[k][n - 1] = src[k][n - 1];                   // odd tail element: a plain copy
            }
        }
        {   // every thread waits for the transaction count (phase 0)
            unsigned done = 0;
            while (!done)
                asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\nselp.u32 %0, 1, 0, p;\n}"
                             : "=r"(done) : "r"(bar_s) : "memory");
        }
        __syncthreads();   // the tail elements stored by thread 0
        A.F.cenx = s_cenx; A.F.ceny = s_ceny; A.F.nrmx = s_nrmx; A.F.nrmy = s_nrmy;
        A.E.ts = s_ts;
        if (T.have_sinb) A.Q.sinb = s_sb;
    }
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const bool active = t < A.Q.B;
    sto::fast_candidate(A, V, active ? t : A.Q.B - 1, active, threadIdx.x & ~31);
}

// The same launch shape with the one-loop kernel (sto_qss_memo2.cuh): shared evaluate-and-commit, four small search stages.
// MAXR = register budget per thread: 168 leaves the allocation free (~160 registers, at most 12 one-warp CTAs per SM),
// 144 / 128 cap it for batches that put 13 / more warps on an SM.
template <int G, int MAXR, bool FP>
__global__ void __maxnreg__(MAXR) qss_memo2_kernel(sto::QssArgs A, sto::MemoWork W, int cpw, const __grid_constant__ sto_vehicle_f64 V) {
    const int lane = threadIdx.x & 31, warp = blockIdx.x;
    const int grp = lane / G, g = lane % G;
    const int b = warp * cpw + grp;
    const bool active = grp < cpw && b < A.B;
    const sto::MemoCtx C = sto::memo_bind(sto_planes, cpw, grp < cpw ? grp : 0, nullptr, 0, lane, A.N, W.W);
    sto::qss_memo2_candidate<G, FP>(A, W, C, V, active ? b : A.B - 1, active, g, grp * G, grp < cpw ? grp : 0, cpw);
}

// The one-loop kernel with the re-spawned lists walked out of order (sto_qss_memo3.cuh).
template <int G, int MAXR>
__global__ void __maxnreg__(MAXR) qss_memo3_kernel(sto::QssArgs A, sto::MemoWork W, int cpw, const __grid_constant__ sto_vehicle_f64 V) {
    const int lane = threadIdx.x & 31, warp = blockIdx.x;
    const int grp = lane / G, g = lane % G;
    const int b = warp * cpw + grp;
    const bool active = grp < cpw && b < A.B;
    const sto::MemoCtx C = sto::memo_bind(sto_planes, cpw, grp < cpw ? grp : 0, nullptr, 0, lane, A.N, W.W);
    sto::qss_memo3_candidate<G>(A, W, C, V, active ? b : A.B - 1, active, g, grp * G, grp < cpw ? grp : 0, cpw);
}

// FP64 pipe peak for the roofline report: 8 independent DFMA chains per thread, every SM full.
__global__ void fp64_peak_kernel(double* out, int iters, double a, double b) {
    double x0 = a + threadIdx.x * 1e-9, x1 = x0 + 1e-3, x2 = x0 + 2e-3, x3 = x0 + 3e-3, x4 = x0 + 4e-3, x5 = x0 + 5e-3,
           x6 = x0 + 6e-3, x7 = x0 + 7e-3;
    for (int i = 0; i < iters; ++i) {
        x0 = __fma_rn(x0, b, a); x1 = __fma_rn(x1, b, a); x2 = __fma_rn(x2, b, a); x3 = __fma_rn(x3, b, a);
        x4 = __fma_rn(x4, b, a); x5 = __fma_rn(x5, b, a); x6 = __fma_rn(x6, b, a); x7 = __fma_rn(x7, b, a);
    }
    const double r = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
    if (r == 123.456) out[0] = r;   // never true: keeps the chains alive
}

// Self-test of the flagged fast-path division / square root (sto_common.cuh): every thread draws operand pairs from a
// counter-based generator - all bit patterns, mid-range magnitudes, near-equal pairs, exact zeros - and counts results
// that differ from the plain operators while the flag is down, and how often the flag is up.
__device__ __forceinline__ unsigned long long splitmix64(unsigned long long x) {
    x += 0x9e3779b97f4a7c15ull;
    x = (x ^ (x >> 30)) * 0xbf58476d1ce4e5b9ull;
    x = (x ^ (x >> 27)) * 0x94d049bb133111ebull;
    return x ^ (x >> 31);
}
__global__ void fp64_selftest_kernel(unsigned long long seed, int per_thread, unsigned long long* counts) {
#if defined(__CUDA_ARCH__)
    const unsigned long long tid = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long bad = 0, flagged = 0;
    for (int i = 0; i < per_thread; ++i) {
        const unsigned long long r0 = splitmix64(seed + (tid * (unsigned long long)per_thread + i) * 3ull);
        const unsigned long long r1 = splitmix64(r0), r2 = splitmix64(r1);
        double a, b;
        const int mode = (int)(r2 & 7);
        if (mode < 3) {            // any bit pattern: all exponents, signs, NaN / inf / denormals
            a = __longlong_as_double((long long)r0);
            b = __longlong_as_double((long long)r1);
        } else if (mode < 6) {     // mid-range magnitudes, both signs
            a = (double)(r0 >> 11) * (1.0 / 9007199254740992.0) * 2e6 - 1e6;
            b = exp2((double)((int)(r1 >> 58) - 32)) * (1.0 + (double)(r1 & 0xfffffffffffffull) * (1.0 / 4503599627370496.0));
            if (r2 & 8) b = -b;
        } else {                   // near-equal operands, exact zeros
            a = (double)(r0 >> 11) * (1.0 / 9007199254740992.0) * 100.0;
            b = ((r1 & 3) == 0) ? a : a + (double)((long long)(r1 >> 60) - 8) * 1e-12;
            if ((r1 & 12) == 0) a = (r1 & 16) ? 0.0 : -0.0;
            if (r2 & 8) b = -b;
        }
        bool s1 = false, s2 = false;
        const double q = sto::div_fast(a, b, s1);
        const double qr = a / b;
        const double sa = fabs(a);
        const double w = sto::sqrt_fast(sa, s2);
        const double wr = sqrt(sa);
        if (s1) ++flagged; else if (__double_as_longlong(q) != __double_as_longlong(qr) && !(q != q && qr != qr)) ++bad;
        if (s2) ++flagged; else if (__double_as_longlong(w) != __double_as_longlong(wr) && !(w != w && wr != wr)) ++bad;
    }
    atomicAdd(&counts[0], bad);
    atomicAdd(&counts[1], flagged);
#endif
}

__global__ void zero_status_kernel(int32_t* s, int B) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < B) s[b] = 0;
}

// 32x32 tile transpose through shared memory (+1 padding: conflict-free): dst[c][r] = src[r][c]
__global__ void transpose_kernel(const double* __restrict__ src, int rows, int cols, int ld_src,
                                 double* __restrict__ dst, int ld_dst) {
    __shared__ double tile[32][33];
    int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    for (int k = threadIdx.y; k < 32; k += blockDim.y) {
        int r = r0 + k, c = c0 + threadIdx.x;
        if (r < rows && c < cols) tile[k][threadIdx.x] = src[(size_t)r * ld_src + c];
    }
    __syncthreads();
    for (int k = threadIdx.y; k < 32; k += blockDim.y) {
        int c = c0 + k, r = r0 + threadIdx.x;
        if (r < rows && c < cols) dst[(size_t)c * ld_dst + r] = tile[threadIdx.x][k];
    }
}

__global__ void argmin_kernel(const double* lap, const int32_t* status, int B, double* best_lap,
                              long long* best_idx) {
    __shared__ double s_v[32];
    __shared__ long long s_i[32];
    double bv = INFINITY;
    long long bi = -1;
    for (int b = threadIdx.x; b < B; b += blockDim.x) {
        double v = lap[b];
        if (!(v == v) || (status && status[b] != 0)) continue;
        if (v < bv || bi < 0) { bv = v; bi = b; }
    }
    for (int o = 16; o > 0; o >>= 1) {
        double ov = __shfl_down_sync(0xffffffffu, bv, o);
        long long oi = __shfl_down_sync(0xffffffffu, bi, o);
        if (oi >= 0 && (bi < 0 || ov < bv || (ov == bv && oi < bi))) { bv = ov; bi = oi; }
    }
    if ((threadIdx.x & 31) == 0) { s_v[threadIdx.x >> 5] = bv; s_i[threadIdx.x >> 5] = bi; }
    __syncthreads();
    if (threadIdx.x < 32) {
        int nw = (blockDim.x + 31) >> 5;
        bv = threadIdx.x < nw ? s_v[threadIdx.x] : INFINITY;
        bi = threadIdx.x < nw ? s_i[threadIdx.x] : -1;
        for (int o = 16; o > 0; o >>= 1) {
            double ov = __shfl_down_sync(0xffffffffu, bv, o);
            long long oi = __shfl_down_sync(0xffffffffu, bi, o);
            if (oi >= 0 && (bi < 0 || ov < bv || (ov == bv && oi < bi))) { bv = ov; bi = oi; }
        }
        if (threadIdx.x == 0) { *best_lap = (bi >= 0) ? bv : nan(""); *best_idx = bi; }
    }
}

// Multi-GPU reduction helpers (sharding.py): a rank's (best lap, GLOBAL candidate index) as one 16-byte pair, ready for a
// single all-gather, and the argmin over the gathered pairs (lowest global index wins ties, NaN / -1 pairs ignored).
__global__ void argmin_pair_kernel(const double* best_lap, const long long* best_idx, long long index_base, double* pair) {
    const long long i = best_idx[0];
    pair[0] = (i >= 0) ? best_lap[0] : nan("");
    pair[1] = (i >= 0) ? (double)(i + index_base) : -1.0;   // indices < 2^53 are exact in a double
}
__global__ void argmin_pairs_kernel(const double* pairs, int n, double* best_lap, long long* best_idx) {
    double bv = INFINITY;
    long long bi = -1;
    for (int r = threadIdx.x; r < n; r += 32) {
        const double v = pairs[2 * r];
        const long long i = (long long)pairs[2 * r + 1];
        if (!(v == v) || i < 0) continue;
        if (bi < 0 || v < bv || (v == bv && i < bi)) { bv = v; bi = i; }
    }
    for (int o = 16; o > 0; o >>= 1) {
        const double ov = __shfl_down_sync(0xffffffffu, bv, o);
        const long long oi = __shfl_down_sync(0xffffffffu, bi, o);
        if (oi >= 0 && (bi < 0 || ov < bv || (ov == bv && oi < bi))) { bv = ov; bi = oi; }
    }
    if (threadIdx.x == 0) { *best_lap = (bi >= 0) ? bv : nan(""); *best_idx = bi; }
}

// Optional stage timing of the fused path (bench.py's roofline leg): CUDA events recorded on the launching
// stream between the stages; nothing synchronises until sto_last_stage_ms() is called.
struct StageTimer {
    bool on = false;
    cudaEvent_t ev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    bool recorded = false;
};
thread_local StageTimer g_timer;
inline void stage_mark(int i, cudaStream_t st) {
    if (!g_timer.on) return;
    if (!g_timer.ev[i]) cudaEventCreate(&g_timer.ev[i]);
    cudaEventRecord(g_timer.ev[i], st);
    if (i == 4) g_timer.recorded = true;
}

int check_vehicle(const sto_vehicle_f64* v) {
    if (!v) return fail(STO_ERR_INVALID, "vehicle is NULL");
    if (v->n_acc < 2 || v->n_acc > STO_MAX_BREAKS || v->n_dcc < 2 || v->n_dcc > STO_MAX_BREAKS)
        return fail(STO_ERR_INVALID, "vehicle table sizes must be in [2, STO_MAX_BREAKS]");
    return STO_OK;
}

// Large-batch variant: bit planes in global memory (candidate-major, L1/L2 cached), only the 2 KB prefetch ring in
// shared memory, so residency is bounded by registers (~19 warps/SM) instead of 2.2 KB of planes per candidate.
__global__ void qss_memo_gplanes_kernel(sto::QssArgs A, sto::MemoWork W, int lanes,
                                        const __grid_constant__ sto_vehicle_f64 V) {
    __shared__ int32_t ring[STO_LIST_RING * 32];
    bool active;
    const int b = candidate_of_thread(lanes, A.B, active);
    const int lane = threadIdx.x & 31;
    const sto::MemoCtx C = sto::memo_bind_global(W.gplanes + (size_t)b * 6 * W.W, ring, 32, lane, A.N, W.W);
    sto::qss_memo_candidate<1>(A, W, C, V, b, active, 0, lane);
}

// Large-batch variant 2: the two CONT planes in shared memory (736 B per Monza candidate -> every warp of a 32 k batch
// stays resident), live / STOP planes in global memory.
__global__ void qss_memo_mixed_kernel(sto::QssArgs A, sto::MemoWork W, int lanes,
                                      const __grid_constant__ sto_vehicle_f64 V) {
    bool active;
    const int b = candidate_of_thread(lanes, A.B, active);
    const int lane = threadIdx.x & 31;
    int32_t* ring = reinterpret_cast<int32_t*>(sto_planes + (size_t)2 * W.W * 32);
    const sto::MemoCtx C = sto::memo_bind_mixed(W.gplanes + (size_t)b * 6 * W.W, sto_planes, 32, lane, ring, 32, lane,
                                                A.N, W.W);
    sto::qss_memo_candidate<1>(A, W, C, V, b, active, 0, lane);
}

// Candidates per warp for the QSS kernels: aim for a few warps on each of the 148 SMs before filling warps.
int pick_lanes(int B, size_t smem_per_candidate) {
    int lanes = 32;
    if (const int v = g_tune.qss_lanes.load()) lanes = v;
    else while (lanes > 4 && (B + lanes - 1) / lanes < 148 * 2) lanes >>= 1;
    while (lanes > 1 && smem_per_candidate * lanes > kMemoSmemBudget) lanes >>= 1;
    return lanes;
}

int launch_qss(const sto::QssArgs& A, const QssWork& w, const sto_vehicle_f64* vehicle, int impl, bool owner,
               cudaStream_t st) {
    impl = effective_impl(A.N, impl);
    if (impl == STO_QSS_PLAIN) {
        const int lanes = pick_lanes(A.B, 0);
        const int warps = (A.B + lanes - 1) / lanes;
        const int block = 64;
        const int grid = (warps * 32 + block - 1) / block;
        g_last_qss_kernel = "qss_plain_kernel";
        if (owner) qss_plain_kernel<true><<<grid, block, 0, st>>>(A, lanes, *vehicle);
        else qss_plain_kernel<false><<<grid, block, 0, st>>>(A, lanes, *vehicle);
    } else {
        // Where the bit planes live: shared memory while the whole batch is resident in one wave (small batches are
        // bound by the per-candidate dependent chain, and LDS beats an L1 hit), global memory (L1/L2 cached) beyond
        // that, where throughput is bound by how many candidates an SM can keep in flight (measured on B200, Monza:
        // 4,096 candidates 86 ms smem vs 92 ms global; 32,768 candidates 299 ms smem vs 186 ms global).
        const int pl = g_tune.qss_planes.load();   // tuning override (see Tuning)
        const bool global_planes = pl ? (pl >= 2)
                                      : ((size_t)A.B * sto::memo_plane_bytes(A.N) > (size_t)148 * kMemoSmemBudget);
        if (global_planes) {
            const int lanes = pick_lanes(A.B, 0);
            const int warps = (A.B + lanes - 1) / lanes;
            const size_t mixed_smem = (size_t)2 * w.memo.W * 32 * sizeof(unsigned long long) + STO_LIST_RING * 32 * sizeof(int32_t);
            // mixed placement while the warps of TWO such launches stay resident with their CONT planes in shared memory.
            // Measured on B200, Monza: one launch of 32,768 candidates 154 ms mixed vs 164 ms all-global (131,072: 662 vs
            // 559 ms) - but 32,768 candidates are only 7 warps per SM, a caller that keeps two launches in flight (steps
            // pipelined over two streams, two host threads on the *_host entry) fills the other half of the SM only if the
            // 25.6 KB of shared memory per warp do not stand in the way: 124.6 ms per step all-global against 156.2 ms
            // mixed (263 k vs 210 k candidates/s).  So the mixed kernel is left to launches of at most half the SMs' capacity.
            const int resident = 148 * (int)((size_t)227 * 1024 / (mixed_smem + 1024));
            const bool mixed = (pl >= 3) ? (pl == 4) : (2 * warps <= resident);
            if (mixed && mixed_smem <= kMemoSmemBudget / 4) {
                STO_CUDA(cudaFuncSetAttribute(qss_memo_mixed_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                              (int)mixed_smem));
                g_last_qss_kernel = "qss_memo_mixed_kernel";
                qss_memo_mixed_kernel<<<warps, 32, mixed_smem, st>>>(A, w.memo, lanes, *vehicle);
            } else {
                g_last_qss_kernel = "qss_memo_gplanes_kernel";
                qss_memo_gplanes_kernel<<<warps, 32, 0, st>>>(A, w.memo, lanes, *vehicle);
            }
            STO_CUDA(cudaGetLastError());
            return STO_OK;
        }
        // Small batch: bound by the per-candidate dependent chain -> spend idle lanes on it.  A group of G lanes per
        // candidate evaluates the backward sub-pass G fronts at a time; cpw <= 32 / G candidates share a warp.
        // measured on B200 (Monza, 4,096 candidates): 4 candidates x 8 lanes per warp 63.3 ms, 8 x 4 lanes 65.1 ms,
        // 8 x 2 lanes 72.9 ms; larger batches fill the SMs with 8 candidates per warp
        int G = ((A.B + 3) / 4 <= 148 * 8) ? 8 : 4;
        // the one-loop kernel (N <= 4096), measured on B200, Monza: 1,024 candidates 27.1 ms at 1 x 32 lanes per warp (32.8 at
        // 2 x 16); 2,048: 33.1 ms at 2 x 16 (43.8 at 4 x 8); 4,096: 41.3 ms at 4 x 8, 40.9 at 2 x 16 - but 2 x 16 is 2,048
        // warps = 14 per SM in the 128-register build and leaves no room for a second launch, while 4 x 8 (7 warps per SM,
        // 168 registers) lets the next step's kernels start under this one's tail: 35.8 vs 43.8 ms per step when steps are
        // pipelined over two streams.  So: as many lanes per candidate as keep the launch at <= 7 warps per SM.
        if (g_tune.qss_kernel.load() != 1 && w.memo.W <= 64) {
            if (A.B <= 148 * 7) G = 32;
            else if ((A.B + 1) / 2 <= 148 * 7) G = 16;
            else G = 8;
        }
        if (const int v = g_tune.qss_group.load()) G = v;
        const size_t per_cand = sto::memo_plane_bytes(A.N);
        int cpw = pick_lanes(A.B, per_cand);
        if (cpw > 32 / G) cpw = 32 / G;
        while (cpw > 1 && sto::memo_smem_bytes(A.N, cpw) > kMemoSmemBudget) cpw >>= 1;
        const int warps = (A.B + cpw - 1) / cpw;
        const size_t smem = sto::memo_smem_bytes(A.N, cpw);
        const int which = g_tune.qss_kernel.load();
        if (which != 1 && w.memo.W <= 64 && G >= 8) {   // the one-loop kernels: lane groups of 8 / 16 / 32, N <= 4096
            // default (0, 4): sto_qss_memo2.cuh with the forward original-row sub-pass run-parallel (measured on B200, Monza:
            // 4,096 lines 30.9 ms against 33.2 ms with the sequential sweep, 8,192 lines 41.1 / 43.2, 2,048 lines 23.9 / 25.7,
            // 1,024 lines 19.6 / 20.7); 2 = the sequential sweep; 3 = sto_qss_memo3.cuh, the re-spawned lists walked out of
            // order (fewer rounds, 1,882 instead of 5,244 on the forward list, but the list scan that finds them costs more
            // than they save: 40.2 ms) - kept selectable for A/B runs, with the 168-register build only
            const int per_sm = (warps + 147) / 148;
            const int mb = (per_sm <= 12) ? 168 : (per_sm <= 13) ? 144 : 128;
            if (which == 3) {
                const size_t smem3 = sto::memo3_smem_bytes(w.memo.W, cpw);
#define STO_LAUNCH_MEMO3(GG)                                                                                               \
    do {                                                                                                                   \
        STO_CUDA(cudaFuncSetAttribute(qss_memo3_kernel<GG, 168>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem3)); \
        qss_memo3_kernel<GG, 168><<<warps, 32, smem3, st>>>(A, w.memo, cpw, *vehicle);                                     \
    } while (0)
                g_last_qss_kernel = "qss_memo3_kernel";
                if (G == 32) STO_LAUNCH_MEMO3(32);
                else if (G == 16) STO_LAUNCH_MEMO3(16);
                else STO_LAUNCH_MEMO3(8);
#undef STO_LAUNCH_MEMO3
                STO_CUDA(cudaGetLastError());
                return STO_OK;
            }
            // six planes + the BLOCKED plane + the TODO word of the run-parallel forward phase
            const size_t smem2 = ((size_t)7 * w.memo.W + 1) * sizeof(unsigned long long) * cpw;
#define STO_LAUNCH_MEMO2(GG, MB, FP)                                                                                            \
    do {                                                                                                                        \
        STO_CUDA(cudaFuncSetAttribute(qss_memo2_kernel<GG, MB, FP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));  \
        qss_memo2_kernel<GG, MB, FP><<<warps, 32, smem2, st>>>(A, w.memo, cpw, *vehicle);                                       \
    } while (0)
#define STO_LAUNCH_MEMO2_G(GG)                                      \
    do {                                                            \
        if (which == 2) STO_LAUNCH_MEMO2(GG, 168, false);           \
        else if (mb == 168) STO_LAUNCH_MEMO2(GG, 168, true);        \
        else if (mb == 144) STO_LAUNCH_MEMO2(GG, 144, true);        \
        else STO_LAUNCH_MEMO2(GG, 128, true);                       \
    } while (0)
            g_last_qss_kernel = "qss_memo2_kernel";
            if (G == 32) STO_LAUNCH_MEMO2_G(32);
            else if (G == 16) STO_LAUNCH_MEMO2_G(16);
            else STO_LAUNCH_MEMO2_G(8);
#undef STO_LAUNCH_MEMO2_G
#undef STO_LAUNCH_MEMO2
            STO_CUDA(cudaGetLastError());
            return STO_OK;
        }
#define STO_LAUNCH_MEMO(GG)                                                                                          \
    do {                                                                                                             \
        STO_CUDA(cudaFuncSetAttribute(qss_memo_kernel<GG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        qss_memo_kernel<GG><<<warps, 32, smem, st>>>(A, w.memo, cpw, *vehicle);                                      \
    } while (0)
        g_last_qss_kernel = "qss_memo_kernel";
        if (G == 32) STO_LAUNCH_MEMO(32);
        else if (G == 16) STO_LAUNCH_MEMO(16);
        else if (G == 8) STO_LAUNCH_MEMO(8);
        else if (G == 4) STO_LAUNCH_MEMO(4);
        else if (G == 2) STO_LAUNCH_MEMO(2);
        else STO_LAUNCH_MEMO(1);
#undef STO_LAUNCH_MEMO
    }
    STO_CUDA(cudaGetLastError());
    return STO_OK;
}

}  // namespace

// =============================================================================================================
extern "C" {

int sto_abi_version(void) { return STO_B200_ABI_VERSION; }
const char* sto_last_qss_kernel(void) { return g_last_qss_kernel; }
const char* sto_last_error(void) { return g_err.c_str(); }

int sto_set_fit_solver(int solver) {
    if (solver != STO_FIT_THOMAS && solver != STO_FIT_BLOCKS && solver != STO_FIT_FITPACK)
        return fail(STO_ERR_INVALID, "unknown fit solver");
    g_tune.fit_solver.store(solver);
    return STO_OK;
}
int sto_get_fit_solver(void) { return g_tune.fit_solver.load(); }

int sto_set_tuning(const char* key, int value) {
    if (!key) return fail(STO_ERR_INVALID, "key is NULL");
    const std::string k(key);
    if (k == "fit_split" && (value == 0 || pow2_le32(value))) { g_tune.fit_split = value; return STO_OK; }
    if (k == "qss_lanes" && value >= 0 && value <= 32) { g_tune.qss_lanes = value; return STO_OK; }
    if (k == "qss_group" && (value == 0 || pow2_le32(value))) { g_tune.qss_group = value; return STO_OK; }
    if (k == "qss_planes" && value >= 0 && value <= 4) { g_tune.qss_planes = value; return STO_OK; }
    if (k == "qss_kernel" && value >= 0 && value <= 4) { g_tune.qss_kernel = value; return STO_OK; }
    return fail(STO_ERR_INVALID, "unknown tuning key or value");
}
int sto_fit_solver_lanes(int M, int B) {
    if (M < 3 || B < 1) return 0;
    return pick_fit_plan(M, B).split;
}

int sto_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    int ok = 0;
    for (int d = 0; d < n; ++d) {
        cudaDeviceProp p;
        if (cudaGetDeviceProperties(&p, d) == cudaSuccess && p.major == 10) ++ok;
    }
    return ok;
}

size_t sto_fit_workspace_bytes(int M, int B) {
    if (M < 3 || B < 1) return 0;
    Carver c(nullptr);
    carve_fit(c, M, ldof(B));
    return c.bytes();
}

int sto_fit_periodic_cubic_f64(const double* centre_x, const double* centre_y, const double* normal_x,
                               const double* normal_y, const double* offsets, const double* px,
                               const double* py, int M, int B, int ld, double* u, double* cx, double* cy,
                               int32_t* status, void* work, size_t work_bytes, void* stream) {
    if (M < 3) return fail(STO_ERR_INVALID, "fit needs M >= 3 points (trajectory.py:214)");
    if (B < 1 || ld < B) return fail(STO_ERR_INVALID, "need B >= 1 and ld >= B");
    if (!u || !cx || !cy || !work) return fail(STO_ERR_INVALID, "u, cx, cy, work must be non-NULL");
    const bool frenet = centre_x != nullptr;
    if (frenet && (!centre_y || !normal_x || !normal_y || !offsets))
        return fail(STO_ERR_INVALID, "centre/normal/offsets must all be given");
    if (!frenet && (!px || !py)) return fail(STO_ERR_INVALID, "either centre+normal+offsets or px+py");
    // the work arrays use the caller's leading dimension
    Carver c(work);
    sto::FitArgs A{};
    A.cp = c.take<double>((size_t)M * ld);
    A.zx = c.take<double>((size_t)M * ld);
    A.zy = c.take<double>((size_t)M * ld);
    A.zz = c.take<double>((size_t)M * ld);
    A.ze = c.take<double>((size_t)M * ld);
    if (c.bytes() > work_bytes && (size_t)ld > ldof(B)) return fail(STO_ERR_WORKSPACE, "fit workspace: ld too large");
    if (c.bytes() > work_bytes) return fail(STO_ERR_WORKSPACE, "fit workspace too small");
    A.cenx = centre_x; A.ceny = centre_y; A.nrmx = normal_x; A.nrmy = normal_y; A.off = offsets;
    A.px = px; A.py = py; A.M = M; A.B = B; A.ld = ld; A.u = u; A.cx = cx; A.cy = cy; A.status = status;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    launch_fit(A, pick_fit_plan(M, B), st);
    STO_CUDA(cudaGetLastError());
    return STO_OK;
}

size_t sto_fit_lsq_workspace_bytes(int M, int nt, int k, int B) {
    if (B < 1 || !sto::lsq_sizes_ok(M, nt, k)) return 0;
    return sto::lsq_work_doubles(nt, k) * ldof(B) * sizeof(double) + 256;
}

int sto_fit_periodic_lsq_f64(const double* centre_x, const double* centre_y, const double* normal_x,
                             const double* normal_y, const double* offsets, const double* px, const double* py,
                             int M, int B, int ld, const double* t, int nt, int k, double* u, double* cx, double* cy,
                             int32_t* status, void* work, size_t work_bytes, void* stream) {
    if (!sto::lsq_sizes_ok(M, nt, k))
        return fail(STO_ERR_INVALID, "least-squares fit needs 1 <= k <= 5, nt - 2k - 1 >= 3k + 1 free coefficients, M >= that");
    if (B < 1 || ld < B) return fail(STO_ERR_INVALID, "need B >= 1 and ld >= B");
    if (!t || !u || !cx || !cy || !work) return fail(STO_ERR_INVALID, "t, u, cx, cy, work must be non-NULL");
    const bool frenet = centre_x != nullptr;
    if (frenet && (!centre_y || !normal_x || !normal_y || !offsets))
        return fail(STO_ERR_INVALID, "centre/normal/offsets must all be given");
    if (!frenet && (!px || !py)) return fail(STO_ERR_INVALID, "either centre+normal+offsets or px+py");
    const size_t g = (size_t)(nt - 2 * k - 1);
    if (sto::lsq_work_doubles(nt, k) * (size_t)ld * sizeof(double) + 256 > work_bytes)
        return fail(STO_ERR_WORKSPACE, "least-squares fit workspace too small (sto_fit_lsq_workspace_bytes; ld = round_up(B, 32))");
    Carver c(work);
    sto::LsqArgs A{};
    A.F.cenx = centre_x; A.F.ceny = centre_y; A.F.nrmx = normal_x; A.F.nrmy = normal_y; A.F.off = offsets;
    A.F.px = px; A.F.py = py; A.F.M = M; A.F.B = B; A.F.ld = ld; A.F.u = u; A.F.status = status;
    A.t = t; A.nt = nt; A.k = k; A.cx = cx; A.cy = cy;
    A.band = c.take<double>(sto::lsq_work_doubles(nt, k) * (size_t)ld);
    A.rx = A.band + g * (size_t)(k + 1) * ld;
    A.ry = A.rx + g * ld;
    A.Y = A.ry + g * ld;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int block = (B >= 148 * 128) ? 128 : 32;
    fit_lsq_kernel<<<grid_for(B, block), block, 0, st>>>(A);
    STO_CUDA(cudaGetLastError());
    return STO_OK;
}

int sto_sample_f64(const double* u, const double* cx, const double* cy, int M, const double* ts, int N, int B,
                   int ld, double* x, double* y, double* yaw, double* radius, double* chord_qss,
                   double* chord_norm, void* stream) {
    if (M < 3 || N < 1 || B < 1 || ld < B) return fail(STO_ERR_INVALID, "bad sizes");
    if (!u || !cx || !cy || !ts) return fail(STO_ERR_INVALID, "u, cx, cy, ts must be non-NULL");
    sto::EvalArgs A{u, cx, cy, ts, M, N, B, ld, x, y, yaw, radius, chord_qss, chord_norm};
    const int split = pick_split(B, N);
    const int block = pick_block(B * split);
    eval_kernel<<<grid_for(B * split, block), block, 0, static_cast<cudaStream_t>(stream)>>>(A, split);
    STO_CUDA(cudaGetLastError());
    return STO_OK;
}

int sto_sample_spline_f64(const double* t, int nt, const double* cx, const double* cy, int k, const double* ts,
                          int N, double* x, double* y, double* yaw, double* radius, void* stream) {
    if (k < 1 || k > 5 || nt < 2 * (k + 1) || N < 1) return fail(STO_ERR_INVALID, "bad spline sizes");
    if (!t || !cx || !cy || !ts) return fail(STO_ERR_INVALID, "t, cx, cy, ts must be non-NULL");
    sto::SplineEvalArgs A{t, cx, cy, ts, nt, k, N, x, y, yaw, radius};
    eval_spline_kernel<<<(N + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(A);
    STO_CUDA(cudaGetLastError());
    return STO_OK;
}

static int launch_spline_batch(const sto::SplineBatchArgs& A, cudaStream_t st) {
    const int block = (A.B >= 128) ? 128 : 32;
    dim3 grid((A.B + block - 1) / block, A.N < 65535 ? A.N : 65535);
    eval_spline_batch_kernel<<<grid, block, 0, st>>>(A);
    STO_CUDA(cudaGetLastError());
    return STO_OK;
}

int sto_sample_splines_f64(const double* t, int nt, int k, const double* cx, const double* cy, const double* ts, int N,
                           int B, int ld, double* x, double* y, double* yaw, double* radius, double* chord_qss,
                           double* chord_norm, void* stream) {
    if (k < 1 || k > 5 || nt < 2 * (k + 1) || N < 1 || B < 1 || ld < B)
        return fail(STO_ERR_INVALID, "bad spline batch sizes");
    if (!t || !cx || !cy || !ts) return fail(STO_ERR_INVALID, "t, cx, cy, ts must be non-NULL");
    sto::SplineBatchArgs A{t, nt, k, cx, cy, ts, N, B, ld, x, y, yaw, radius, chord_qss, chord_norm};
    return launch_spline_batch(A, static_cast<cudaStream_t>(stream));
}

int sto_fill_bounds_f64(const double* px, const double* py, const double* nx, const double* ny, int n,
                        const double* ring_xy, int m, double max_dist, double* bx, double* by, int32_t* found,
                        void* stream) {
    if (n < 1 || m < 2 || !px || !py || !nx || !ny || !ring_xy || !bx || !by)
        return fail(STO_ERR_INVALID, "bad fill_bounds arguments");
    fill_bounds_kernel<<<(n + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(px, py, nx, ny, n, ring_xy, m,
                                                                                     max_dist, bx, by, found);
    STO_CUDA(cudaGetLastError());
    return STO_OK;
}

int sto_arc_sections_f64(const double* t, int nt, const double* cx, const double* cy, int k, const double* ts, int N,
                         double* sec, void* stream) {
    if (k < 1 || k > 5 || nt < 2 * (k + 1) || N < 1) return fail(STO_ERR_INVALID, "bad spline sizes");
    if (!t || !cx || !cy || !ts || !sec) return fail(STO_ERR_INVALID, "NULL argument");
    sto::ArcArgs A{t, cx, cy, ts, nt, k, N, sec};
    arc_sections_kernel<<<(N + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(A);
    STO_CUDA(cudaGetLastError());
    return STO_OK;
}

size_t sto_qss_workspace_bytes(int N, int B, int impl) {
    if (N < 2 || B < 1) return 0;
    Carver c(nullptr);
    carve_qss(c, N, ldof(B), impl, true, true);
    c.take<int32_t>((size_t)N * ldof(B));  // owner, if requested through a narrower ld
    return c.bytes();
}

int sto_qss_f64(const double* x, const double* y, const double* radius, const double* sin_bank, int N, int B,
                int ld, const sto_vehicle_f64* vehicle, int impl, double* lap, double* summary,
                const sto_profile_out_f64* out, int32_t* status, void* work, size_t work_bytes, void* stream) {
    if (N < 2 || B < 1 || ld < B) return fail(STO_ERR_INVALID, "bad sizes");
    if (!x || !y || !radius || !lap || !work) return fail(STO_ERR_INVALID, "x, y, radius, lap, work must be non-NULL");
    if (impl != STO_QSS_PLAIN && impl != STO_QSS_MEMO) return fail(STO_ERR_INVALID, "unknown impl");
    if (int rc = check_vehicle(vehicle)) return rc;
    if (impl == STO_QSS_MEMO && out && out->owner)
        return fail(STO_ERR_INVALID, "ITERATION_FLAG (owner) is only tracked by STO_QSS_PLAIN");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const size_t wld = ldof(B);
    // the kernels work on one leading dimension (the workspace's); checked before anything is enqueued
    if ((size_t)ld != wld) return fail(STO_ERR_INVALID, "sto_qss_f64 needs ld == round_up(B, 32)");
    Carver c(work);
    QssWork w = carve_qss(c, N, wld, impl, true, true);
    if (c.bytes() > work_bytes) return fail(STO_ERR_WORKSPACE, "qss workspace too small");
    {
        dim3 g((B + 127) / 128, N < 65535 ? N : 65535);
        chord_kernel<<<g, 128, 0, st>>>(x, y, N, B, ld, w.dd, w.df, (int)wld);
        STO_CUDA(cudaGetLastError());
    }
    const double* Rw = radius;
    sto::QssArgs A{};
    A.dd = w.dd; A.df = w.df; A.R = Rw; A.sinb = sin_bank; A.N = N; A.B = B; A.ld = (int)wld; A.cap = w.cap;
    A.v = (out && out->speed) ? out->speed : w.v;      // memo kernel: outputs only (NULL = not materialised)
    A.a = (out && out->lon_acc) ? out->lon_acc : w.a;
    A.rec = w.rec;
    A.rowflag = w.rowflag; A.sp_ent = w.sp_ent; A.sp_ext = w.sp_ext; A.sp_turn = w.sp_turn; A.sp_flag = w.sp_flag;
    A.owner = out ? out->owner : nullptr;
    A.lat = out ? out->lat_acc : nullptr;
    A.tseg = out ? out->time : nullptr;
    A.lap = lap; A.summary = summary; A.status = status;
    if (status) {
        zero_status_kernel<<<(B + 255) / 256, 256, 0, st>>>(status, B);
        STO_CUDA(cudaGetLastError());
    }
    return launch_qss(A, w, vehicle, impl, A.owner != nullptr, st);
}

struct LapWork {
    double *u, *cx, *cy, *R;
    FitWork fit;
    QssWork qss;
};
static LapWork carve_lap(Carver& c, int M, int N, size_t ld, int impl) {
    LapWork w;
    w.u = c.take<double>((size_t)(M + 1) * ld);
    w.cx = c.take<double>((size_t)(M + 3) * ld);
    w.cy = c.take<double>((size_t)(M + 3) * ld);
    w.R = c.take<double>((size_t)N * ld);
    w.fit = carve_fit(c, M, ld);
    w.qss = carve_qss(c, N, ld, impl, true, true);
    return w;
}

size_t sto_lap_workspace_bytes(int M, int N, int B, int impl) {
    if (M < 3 || N < 2 || B < 1) return 0;
    Carver c(nullptr);
    carve_lap(c, M, N, ldof(B), impl);
    return c.bytes();
}

int sto_lap_time_f64(const double* centre_x, const double* centre_y, const double* normal_x,
                     const double* normal_y, const double* sin_bank, const double* ts, const double* offsets,
                     int M, int N, int B, int ld, const sto_vehicle_f64* vehicle, int impl, double* lap,
                     int32_t* status, void* work, size_t work_bytes, void* stream) {
    if (M < 3 || N < 2 || B < 1 || ld < B) return fail(STO_ERR_INVALID, "bad sizes");
    if (!centre_x || !centre_y || !normal_x || !normal_y || !ts || !offsets || !lap || !status || !work)
        return fail(STO_ERR_INVALID, "NULL argument");
    if (impl != STO_QSS_PLAIN && impl != STO_QSS_MEMO) return fail(STO_ERR_INVALID, "unknown impl");
    if ((size_t)ld != ldof(B)) return fail(STO_ERR_INVALID, "sto_lap_time_f64 needs ld == round_up(B, 32)");
    if (int rc = check_vehicle(vehicle)) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    Carver c(work);
    LapWork w = carve_lap(c, M, N, (size_t)ld, impl);
    if (c.bytes() > work_bytes) return fail(STO_ERR_WORKSPACE, "lap workspace too small");
    const int block = pick_block(B), grid = grid_for(B, block);
    stage_mark(0, st);
    zero_status_kernel<<<(B + 255) / 256, 256, 0, st>>>(status, B);
    stage_mark(1, st);
    sto::FitArgs F{};
    F.cenx = centre_x; F.ceny = centre_y; F.nrmx = normal_x; F.nrmy = normal_y; F.off = offsets;
    F.M = M; F.B = B; F.ld = ld; F.u = w.u; F.cx = w.cx; F.cy = w.cy; F.status = status;
    F.cp = w.fit.cp; F.zx = w.fit.zx; F.zy = w.fit.zy; F.zz = w.fit.zz; F.ze = w.fit.ze;
    {
        launch_fit(F, pick_fit_plan(M, B), st);
    }
    STO_CUDA(cudaGetLastError());
    stage_mark(2, st);
    // lap-only: x, y, yaw are never materialised; the chords and the radius are all the QSS reads
    sto::EvalArgs E{w.u, w.cx, w.cy, ts, M, N, B, ld, nullptr, nullptr, nullptr, w.R, w.qss.dd, w.qss.df};
    {
        const int split = pick_split(B, N);
        const int eblock = pick_block(B * split);
        eval_kernel<<<grid_for(B * split, eblock), eblock, 0, st>>>(E, split);
    }
    STO_CUDA(cudaGetLastError());
    stage_mark(3, st);
    sto::QssArgs A{};
    A.dd = w.qss.dd; A.df = w.qss.df; A.R = w.R; A.sinb = sin_bank; A.N = N; A.B = B; A.ld = ld; A.cap = w.qss.cap;
    A.v = w.qss.v; A.a = w.qss.a; A.rec = w.qss.rec; A.rowflag = w.qss.rowflag;
    A.sp_ent = w.qss.sp_ent; A.sp_ext = w.qss.sp_ext; A.sp_turn = w.qss.sp_turn; A.sp_flag = w.qss.sp_flag;
    A.lap = lap; A.status = status;
    int rc = launch_qss(A, w.qss, vehicle, impl, false, st);
    stage_mark(4, st);
    return rc;
}

size_t sto_lap_splines_workspace_bytes(int N, int B, int impl) {
    if (N < 2 || B < 1) return 0;
    Carver c(nullptr);
    c.take<double>((size_t)N * ldof(B));
    carve_qss(c, N, ldof(B), impl, true, true);
    return c.bytes();
}

int sto_lap_time_splines_f64(const double* t, int nt, int k, const double* cx, const double* cy, const double* ts,
                             const double* sin_bank, int N, int B, int ld, const sto_vehicle_f64* vehicle, int impl,
                             double* lap, int32_t* status, void* work, size_t work_bytes, void* stream) {
    if (k < 1 || k > 5 || nt < 2 * (k + 1) || N < 2 || B < 1)
        return fail(STO_ERR_INVALID, "bad spline batch sizes");
    if (!t || !cx || !cy || !ts || !lap || !status || !work) return fail(STO_ERR_INVALID, "NULL argument");
    if (impl != STO_QSS_PLAIN && impl != STO_QSS_MEMO) return fail(STO_ERR_INVALID, "unknown impl");
    if ((size_t)ld != ldof(B)) return fail(STO_ERR_INVALID, "sto_lap_time_splines_f64 needs ld == round_up(B, 32)");
    if (int rc = check_vehicle(vehicle)) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    Carver c(work);
    double* R = c.take<double>((size_t)N * ld);
    QssWork w = carve_qss(c, N, (size_t)ld, impl, true, true);
    if (c.bytes() > work_bytes) return fail(STO_ERR_WORKSPACE, "spline lap workspace too small");
    zero_status_kernel<<<(B + 255) / 256, 256, 0, st>>>(status, B);
    sto::SplineBatchArgs E{t, nt, k, cx, cy, ts, N, B, ld, nullptr, nullptr, nullptr, R, w.dd, w.df};
    if (int rc = launch_spline_batch(E, st)) return rc;
    sto::QssArgs A{};
    A.dd = w.dd; A.df = w.df; A.R = R; A.sinb = sin_bank; A.N = N; A.B = B; A.ld = ld; A.cap = w.cap;
    A.v = w.v; A.a = w.a; A.rec = w.rec; A.rowflag = w.rowflag;
    A.sp_ent = w.sp_ent; A.sp_ext = w.sp_ext; A.sp_turn = w.sp_turn; A.sp_flag = w.sp_flag;
    A.lap = lap; A.status = status;
    return launch_qss(A, w, vehicle, impl, false, st);
}

// ---- fast mode ------------------------------------------------------------------------------------------------------
struct FastWork { double *u, *cx, *cy, *R, *dd, *df, *v, *a; FitWork fit; };
static FastWork carve_fast(Carver& c, int M, int N, size_t ld) {
    FastWork w;
    w.u = c.take<double>((size_t)(M + 1) * ld);
    w.cx = c.take<double>((size_t)(M + 3) * ld);
    w.cy = c.take<double>((size_t)(M + 3) * ld);
    w.R = c.take<double>((size_t)N * ld);
    w.dd = c.take<double>((size_t)N * ld);
    w.df = c.take<double>((size_t)N * ld);
    w.v = c.take<double>((size_t)N * ld);
    w.a = c.take<double>((size_t)N * ld);
    w.fit.cp = c.take<double>((size_t)M * ld);
    w.fit.zx = c.take<double>((size_t)M * ld);
    w.fit.zy = c.take<double>((size_t)M * ld);
    w.fit.zz = c.take<double>((size_t)M * ld);
    w.fit.ze = nullptr;   // FITPACK solver only
    return w;
}

size_t sto_fast_workspace_bytes(int M, int N, int B) {
    if (M < 3 || N < 2 || B < 1) return 0;
    Carver c(nullptr);
    carve_fast(c, M, N, ldof(B));
    return c.bytes();
}

int sto_lap_time_fast_f64(const double* centre_x, const double* centre_y, const double* normal_x,
                          const double* normal_y, const double* sin_bank, const double* ts, const double* offsets,
                          int M, int N, int B, int ld, const sto_vehicle_f64* vehicle, int rounds,
                          const sto_fast_out_f64* out, double* lap, int32_t* status, void* work, size_t work_bytes,
                          int stage_tables, void* stream) {
    if (M < 3 || N < 2 || B < 1 || ld < B || rounds < 1) return fail(STO_ERR_INVALID, "bad sizes");
    if (!centre_x || !centre_y || !normal_x || !normal_y || !ts || !offsets || !lap || !status || !work)
        return fail(STO_ERR_INVALID, "NULL argument");
    if ((size_t)ld != ldof(B)) return fail(STO_ERR_INVALID, "sto_lap_time_fast_f64 needs ld == round_up(B, 32)");
    if (int rc = check_vehicle(vehicle)) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    Carver c(work);
    FastWork w = carve_fast(c, M, N, (size_t)ld);
    if (c.bytes() > work_bytes) return fail(STO_ERR_WORKSPACE, "fast-mode workspace too small");
    sto::FastArgs A{};
    A.rounds = rounds;
    sto::FitArgs& F = A.F;
    F.cenx = centre_x; F.ceny = centre_y; F.nrmx = normal_x; F.nrmy = normal_y; F.off = offsets;
    F.M = M; F.B = B; F.ld = ld; F.status = status;
    F.u = w.u; F.cx = (out && out->cx) ? out->cx : w.cx; F.cy = (out && out->cy) ? out->cy : w.cy;
    F.cp = w.fit.cp; F.zx = w.fit.zx; F.zy = w.fit.zy; F.zz = w.fit.zz; F.ze = nullptr;
    A.E = sto::EvalArgs{F.u, F.cx, F.cy, ts, M, N, B, ld, out ? out->x : nullptr, out ? out->y : nullptr,
                        out ? out->yaw : nullptr, (out && out->radius) ? out->radius : w.R, w.dd, w.df};
    sto::QssArgs& Q = A.Q;
    Q.dd = w.dd; Q.df = w.df; Q.R = A.E.radius; Q.sinb = sin_bank; Q.N = N; Q.B = B; Q.ld = ld;
    Q.v = (out && out->speed) ? out->speed : w.v;
    Q.a = (out && out->lon_acc) ? out->lon_acc : w.a;
    Q.lat = out ? out->lat_acc : nullptr;
    Q.tseg = out ? out->time : nullptr;
    Q.lap = lap; Q.status = status;
    // Track tables in shared memory (one bulk copy per table and CTA) when they fit and are 16-byte aligned.
    const size_t pc = (size_t)((M + 1) & ~1), pt = (size_t)((N + 1) & ~1);
    const size_t smem = (4 * pc + 2 * pt) * sizeof(double);
    auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; };
    const bool can_stage = smem <= kMemoSmemBudget && al16(centre_x) && al16(centre_y) && al16(normal_x) &&
                           al16(normal_y) && al16(ts) && (!sin_bank || al16(sin_bank));
    if (stage_tables > 0 && !can_stage)
        return fail(STO_ERR_INVALID, "track tables cannot be staged (need 16-byte aligned pointers and <= 200 KB in all)");
    // automatic: stage while the launch is a single wave of one-CTA-per-SM blocks (measured on B200, Monza: 1.02x at 4,096
    // candidates; at 131,072 the 512-thread blocks the tables force run 0.91-1.0x the speed of 128-thread blocks)
    const bool stage = can_stage && (stage_tables > 0 || (stage_tables < 0 && B <= 148 * 512));
    FastTables T{M, N, sin_bank != nullptr};
    if (stage) {
        // one CTA per SM (the tables take most of its shared memory): as many threads as the batch gives each SM
        int block = ((B + 147) / 148 + 31) & ~31;
        block = block < 32 ? 32 : (block > 512 ? 512 : block);
        STO_CUDA(cudaFuncSetAttribute(fast_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        fast_kernel<true><<<grid_for(B, block), block, smem, st>>>(A, T, *vehicle);
    } else {
        const int block = pick_block(B);
        fast_kernel<false><<<grid_for(B, block), block, 0, st>>>(A, T, *vehicle);
    }
    STO_CUDA(cudaGetLastError());
    return STO_OK;
}

int sto_set_stage_timing(int on) {
    g_timer.on = on != 0;
    g_timer.recorded = false;
    return STO_OK;
}

int sto_last_stage_ms(float* ms4) {
    if (!ms4) return fail(STO_ERR_INVALID, "ms4 is NULL");
    if (!g_timer.on || !g_timer.recorded) return fail(STO_ERR_INVALID, "no timed sto_lap_time_f64 call on this thread");
    STO_CUDA(cudaEventSynchronize(g_timer.ev[4]));
    for (int i = 0; i < 4; ++i) STO_CUDA(cudaEventElapsedTime(&ms4[i], g_timer.ev[i], g_timer.ev[i + 1]));
    return STO_OK;
}

int sto_transpose_f64(const double* src, int rows, int cols, int ld_src, double* dst, int ld_dst, void* stream) {
    if (!src || !dst || rows < 1 || cols < 1 || ld_src < cols || ld_dst < rows)
        return fail(STO_ERR_INVALID, "bad transpose arguments");
    dim3 g((cols + 31) / 32, (rows + 31) / 32), blk(32, 8);
    transpose_kernel<<<g, blk, 0, static_cast<cudaStream_t>(stream)>>>(src, rows, cols, ld_src, dst, ld_dst);
    STO_CUDA(cudaGetLastError());
    return STO_OK;
}

int sto_selftest_fp64(unsigned long long seed, long long* tested, long long* mismatches, long long* flagged) {
    if (!tested || !mismatches || !flagged) return fail(STO_ERR_INVALID, "NULL argument");
    unsigned long long* counts = nullptr;
    STO_CUDA(cudaMalloc(&counts, 2 * sizeof(unsigned long long)));
    STO_CUDA(cudaMemset(counts, 0, 2 * sizeof(unsigned long long)));
    const int block = 256, grid = 1024, per_thread = 256;   // 2^26 operand pairs, two operations each
    fp64_selftest_kernel<<<grid, block>>>(seed, per_thread, counts);
    unsigned long long h[2] = {0, 0};
    STO_CUDA(cudaMemcpy(h, counts, sizeof(h), cudaMemcpyDeviceToHost));
    cudaFree(counts);
    *tested = 2LL * block * grid * per_thread;
    *mismatches = (long long)h[0];
    *flagged = (long long)h[1];
    return STO_OK;
}

int sto_measure_fp64_peak(double* tflops) {
    if (!tflops) return fail(STO_ERR_INVALID, "tflops is NULL");
    int dev = 0, sms = 0;
    STO_CUDA(cudaGetDevice(&dev));
    STO_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    double* out = nullptr;
    STO_CUDA(cudaMalloc(&out, sizeof(double)));
    cudaEvent_t e0, e1;
    STO_CUDA(cudaEventCreate(&e0));
    STO_CUDA(cudaEventCreate(&e1));
    const int iters = 1 << 15, block = 256, grid = sms * 8;
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {   // first pass warms up
        cudaEventRecord(e0);
        fp64_peak_kernel<<<grid, block>>>(out, iters, 1e-9, 0.999999);
        cudaEventRecord(e1);
        STO_CUDA(cudaEventSynchronize(e1));
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0 && ms < best) best = ms;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(out);
    *tflops = 2.0 * 8.0 * (double)iters * (double)block * (double)grid / ((double)best * 1e-3) / 1e12;
    return STO_OK;
}

int sto_argmin_f64(const double* lap, const int32_t* status, int B, double* best_lap, int64_t* best_idx,
                   void* stream) {
    if (!lap || !best_lap || !best_idx || B < 1) return fail(STO_ERR_INVALID, "bad argmin arguments");
    argmin_kernel<<<1, 1024, 0, static_cast<cudaStream_t>(stream)>>>(lap, status, B, best_lap,
                                                                     reinterpret_cast<long long*>(best_idx));
    STO_CUDA(cudaGetLastError());
    return STO_OK;
}

int sto_argmin_pair_f64(const double* lap, const int32_t* status, int B, int64_t index_base, double* pair,
                        void* scratch, void* stream) {
    if (!lap || !pair || !scratch || B < 1) return fail(STO_ERR_INVALID, "bad argmin_pair arguments");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    double* bl = static_cast<double*>(scratch);
    long long* bi = reinterpret_cast<long long*>(bl + 1);
    argmin_kernel<<<1, 1024, 0, st>>>(lap, status, B, bl, bi);
    argmin_pair_kernel<<<1, 1, 0, st>>>(bl, bi, (long long)index_base, pair);
    STO_CUDA(cudaGetLastError());
    return STO_OK;
}

int sto_argmin_pairs_f64(const double* pairs, int n, double* best_lap, int64_t* best_idx, void* stream) {
    if (!pairs || !best_lap || !best_idx || n < 1) return fail(STO_ERR_INVALID, "bad argmin_pairs arguments");
    argmin_pairs_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(pairs, n, best_lap,
                                                                        reinterpret_cast<long long*>(best_idx));
    STO_CUDA(cudaGetLastError());
    return STO_OK;
}

// ---- host-buffer entry: the call a ctypes binding makes ---------------------------------------------------
namespace {
// kHostSlots contexts per device, each a grow-only arena (cudaMalloc is not free), a stream and a mutex.  Threads driving
// different GPUs never meet; up to kHostSlots threads driving the SAME GPU run concurrently on their own streams (a QSS
// launch ends with its slowest candidates: a second call's kernels fill the SMs the others left), further callers queue.
struct HostCtx {
    std::mutex mu;
    void* p = nullptr;
    size_t n = 0;
    cudaStream_t stream = nullptr;
};
constexpr int kMaxDevices = 64;
constexpr int kHostSlots = 2;
HostCtx g_host[kMaxDevices][kHostSlots];
std::atomic<unsigned> g_host_ticket[kMaxDevices];
int arena_get(HostCtx& ctx, size_t bytes, void** out) {   // caller holds ctx.mu and has made the device current
    if (ctx.n < bytes) {
        if (ctx.p) { cudaFree(ctx.p); ctx.p = nullptr; ctx.n = 0; }
        STO_CUDA(cudaMalloc(&ctx.p, bytes));
        ctx.n = bytes;
    }
    *out = ctx.p;
    return STO_OK;
}
}  // namespace

int sto_release(void) {
    int cur = 0;
    const bool have_cur = cudaGetDevice(&cur) == cudaSuccess;
    for (int d = 0; d < kMaxDevices; ++d)
        for (int k = 0; k < kHostSlots; ++k) {
            HostCtx& ctx = g_host[d][k];
            std::lock_guard<std::mutex> lk(ctx.mu);
            if (!ctx.p && !ctx.stream) continue;
            cudaSetDevice(d);
            if (ctx.p) { cudaFree(ctx.p); ctx.p = nullptr; ctx.n = 0; }
            if (ctx.stream) { cudaStreamDestroy(ctx.stream); ctx.stream = nullptr; }
        }
    if (have_cur) cudaSetDevice(cur);
    return STO_OK;
}

int sto_lap_time_host_f64(const double* centre_x, const double* centre_y, const double* normal_x,
                          const double* normal_y, const double* sin_bank, const double* ts,
                          const double* offsets_host, int M, int N, int B, const sto_vehicle_f64* vehicle,
                          int impl, double* lap_host, int32_t* status_host, int device,
                          size_t max_work_bytes) {
    if (M < 3 || N < 2 || B < 1) return fail(STO_ERR_INVALID, "bad sizes");
    if (!centre_x || !centre_y || !normal_x || !normal_y || !ts || !offsets_host || !lap_host || !status_host)
        return fail(STO_ERR_INVALID, "NULL argument");
    if (int rc = check_vehicle(vehicle)) return rc;
    if (device < 0 || device >= kMaxDevices) return fail(STO_ERR_INVALID, "device index out of range");
    // a free context of this device, else wait for one (round robin)
    std::unique_lock<std::mutex> lk;
    int slot = -1;
    for (int k = 0; k < kHostSlots && slot < 0; ++k) {
        std::unique_lock<std::mutex> t(g_host[device][k].mu, std::try_to_lock);
        if (t.owns_lock()) { lk = std::move(t); slot = k; }
    }
    if (slot < 0) {
        slot = (int)(g_host_ticket[device].fetch_add(1u) % kHostSlots);
        lk = std::unique_lock<std::mutex>(g_host[device][slot].mu);
    }
    HostCtx& ctx = g_host[device][slot];
    STO_CUDA(cudaSetDevice(device));
    // chunk size: the largest multiple of 32 candidates whose buffers fit the budget
    auto need = [&](int bc) {
        size_t ld = ldof(bc);
        return sto_lap_workspace_bytes(M, N, bc, impl) + 2 * (size_t)M * ld * sizeof(double)  // staged + transposed
               + (size_t)(5 * M + 2 * N) * sizeof(double) + ld * (sizeof(double) + sizeof(int32_t)) + 8192;
    };
    size_t budget = max_work_bytes ? max_work_bytes : (size_t)32 << 30;
    // Steady state (same shape as the previous call: the cached arena already holds the whole batch) touches no
    // driver-wide state: no cudaMemGetInfo, no allocation, no stream creation.
    const bool arena_fits = ctx.n >= need(B) && need(B) <= budget;
    if (!arena_fits) {
        size_t free_b = 0, total_b = 0;
        STO_CUDA(cudaMemGetInfo(&free_b, &total_b));
        size_t avail = free_b + ctx.n;
        if (budget > avail / 10 * 8) budget = avail / 10 * 8;
    }
    int bc = B;
    while (bc > 32 && need(bc) > budget) bc = ((bc / 2) + 31) & ~31;
    if (need(bc) > budget) return fail(STO_ERR_WORKSPACE, "device memory budget too small for 32 candidates");
    void* base = nullptr;
    if (int rc = arena_get(ctx, need(bc), &base)) return rc;
    const size_t ld = ldof(bc);
    Carver c(base);
    double* d_cx = c.take<double>(M); double* d_cy = c.take<double>(M);
    double* d_nx = c.take<double>(M); double* d_ny = c.take<double>(M);
    double* d_sb = c.take<double>(N); double* d_ts = c.take<double>(N);
    double* d_stage = c.take<double>((size_t)M * ld);  // [bc][M] as the user holds it
    double* d_off = c.take<double>((size_t)M * ld);    // [M][ld]
    double* d_lap = c.take<double>(ld);
    int32_t* d_status = c.take<int32_t>(ld);
    void* d_work = c.take<char>(sto_lap_workspace_bytes(M, N, bc, impl));
    const size_t work_bytes = sto_lap_workspace_bytes(M, N, bc, impl);
    if (!ctx.stream) STO_CUDA(cudaStreamCreateWithFlags(&ctx.stream, cudaStreamNonBlocking));
    cudaStream_t st = ctx.stream;
    int rc = STO_OK;
    auto H2D = [&](void* d, const void* h, size_t n) { return cudaMemcpyAsync(d, h, n, cudaMemcpyHostToDevice, st); };
    cudaError_t e = H2D(d_cx, centre_x, sizeof(double) * M);
    if (e == cudaSuccess) e = H2D(d_cy, centre_y, sizeof(double) * M);
    if (e == cudaSuccess) e = H2D(d_nx, normal_x, sizeof(double) * M);
    if (e == cudaSuccess) e = H2D(d_ny, normal_y, sizeof(double) * M);
    if (e == cudaSuccess) e = H2D(d_ts, ts, sizeof(double) * N);
    if (e == cudaSuccess && sin_bank) e = H2D(d_sb, sin_bank, sizeof(double) * N);
    if (e != cudaSuccess) rc = fail_cuda(e, "H2D of track tables");
    for (int b0 = 0; rc == STO_OK && b0 < B; b0 += bc) {
        const int nb = (B - b0 < bc) ? B - b0 : bc;
        const int nld = (int)ldof(nb);
        e = H2D(d_stage, offsets_host + (size_t)b0 * M, sizeof(double) * (size_t)nb * M);
        if (e != cudaSuccess) { rc = fail_cuda(e, "H2D of offsets"); break; }
        if (nld != nb) {  // padding lanes of the last warp read defined data
            e = cudaMemsetAsync(d_off, 0, sizeof(double) * (size_t)M * nld, st);
            if (e != cudaSuccess) { rc = fail_cuda(e, "memset"); break; }
        }
        rc = sto_transpose_f64(d_stage, nb, M, M, d_off, nld, st);
        if (rc) break;
        rc = sto_lap_time_f64(d_cx, d_cy, d_nx, d_ny, sin_bank ? d_sb : nullptr, d_ts, d_off, M, N, nb, nld, vehicle,
                              impl, d_lap, d_status, d_work, work_bytes, st);
        if (rc) break;
        e = cudaMemcpyAsync(lap_host + b0, d_lap, sizeof(double) * nb, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess)
            e = cudaMemcpyAsync(status_host + b0, d_status, sizeof(int32_t) * nb, cudaMemcpyDeviceToHost, st);
        if (e != cudaSuccess) { rc = fail_cuda(e, "D2H of lap times"); break; }
    }
    e = cudaStreamSynchronize(st);
    if (rc == STO_OK && e != cudaSuccess) rc = fail_cuda(e, "stream synchronize");
    return rc;
}

}  // extern "C"
