// csrc/mm.cu -- Machado-Mata quantile decomposition on the device.
//
// Reference: QuantileDecompositionBuilder::run (quantile_decomposition.rs:281-421): the point pass and every bootstrap
// pass fit `simulations` quantile regressions per group at random quantiles (run_single_pass, :173-279; each one an LP
// handed to the clarabel interior-point solver, math/quantile_regression.rs:22-135), simulate three outcome vectors from
// the fitted coefficients and difference their empirical quantiles.
//
// Here a pass is a column of the multiplicity matrix the OLS bootstrap already builds (resample.cu): a regression on a
// resampled frame is the weighted regression  min sum_i c_i rho_tau(y_i - x_i'beta)  on the original rows.  Every
// (pass, simulation, group) is one independent problem, solved by ONE thread block from start to finish:
//
//   * Frisch-Newton primal-dual interior point (Portnoy & Koenker 1997) on the bounded dual
//       max y'a  s.t.  X'a = (1 - tau) X'c,  0 <= a <= c,
//     Mehrotra predictor-corrector with one step length for both iterates.  Per iteration three sweeps over the group's
//     rows (apply the step + Newton matrix; affine step + the two vectors the corrector's right-hand side is linear in;
//     corrected step); the weighted Gram X'QX and the right-hand sides are FP64 DMMA contractions
//     (mma.sync.m8n8k4.f64) with the rows as the contraction dimension, each warp owning every eighth 32-row block; the
//     row-wise vector algebra runs lane-per-row on the same 32 rows, fed by the same fragment loads (a 4-row x 8-column
//     fragment is four 64-byte segments of the row-major design).  Primal/dual iterates live in a per-block slab of HBM
//     (6 doubles per row, loaded one block ahead of their use; L2-resident for small groups).
//   * polish to the LP's vertex: the rows with a numerically zero residual are compacted in row order, beta is refined on
//     them (normal equations, double-double residuals), and the result is verified (zero residuals there, unchanged
//     residual signs elsewhere).  The answer is then independent of the interior-point path -- which is what makes a
//     1e-10 comparison with the oracle (and with an LP simplex solver) meaningful.
//
// Blocks take problems from an atomic counter; a problem's arithmetic is self-contained and in fixed order, so results do
// not depend on which block solves what.  No CPU fallback.
#include "common.cuh"
#include "internal.h"

#include <cassert>
#include <cstdlib>

// compute-sanitizer is closed on this pool; `make EXTRA=-DOB_MM_BOUNDS_CHECK` compiles device-side bounds checks into every
// global access of the Machado-Mata kernels whose index is computed (iterate slabs, design rows, multiplicity columns,
// candidate lists, outputs); a violation traps with file:line.  tools/mm_bounds_check.sh runs the GPU tests on that build.
#ifdef OB_MM_BOUNDS_CHECK
#define MM_CHECK(cond) assert(cond)
#else
#define MM_CHECK(cond) ((void)0)
#endif

namespace ob {
namespace {

constexpr int MM_THREADS = 256;
constexpr int MM_WARPS = MM_THREADS / 32;
constexpr double MM_RANK_TOL = 1e-14; // a Cholesky pivot below this fraction of its diagonal: rank-deficient up to rounding
constexpr int MM_MAXC = 512;          // zero-residual rows the polish works on (the first ones in row order)
constexpr int MM_MAXK8 = 6;           // design columns + 1 <= 48
constexpr int MM_MAXKP = 8 * MM_MAXK8;

struct MmShared {
    double G[MM_MAXKP * MM_MAXKP];    // Gram of the current Newton system, then its Cholesky factor (lower triangle)
    double vec[4][MM_MAXKP];          // [0] yd (dual of the equality constraints; beta = -yd)  [1] step / rhs  [2] beta  [3] polish iterate
    double ts[MM_THREADS], ys[MM_THREADS], qs[MM_THREADS], vs[MM_THREADS], vs2[MM_THREADS];
    double diag[MM_MAXKP], rhs[MM_MAXKP];   // diagonal of the matrix being factored (rank test), column K of the Gram
    double red[8][MM_WARPS];
    double ah[MM_MAXC];
    int cand[MM_MAXC];
    unsigned wmask[MM_WARPS];
    int ired[16][MM_WARPS];
    int prob, flag;
};

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_min(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_down_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_down_sync(0xffffffffu, v, o));
    return v;
}

// Block reductions in a fixed order (thread partials in row order -> shuffle tree -> warps 0..7): the same bits on every
// run.  op: 0 sum, 1 min, 2 max.  All threads return the result.
template <int N>
__device__ __forceinline__ void block_reduce(MmShared& sh, double (&v)[N], const int (&op)[N]) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < N; ++k) {
        const double r = op[k] == 0 ? warp_sum(v[k]) : (op[k] == 1 ? warp_min(v[k]) : warp_max(v[k]));
        if (lane == 0) sh.red[k][w] = r;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < N; ++k) {
        double r = sh.red[k][0];
#pragma unroll
        for (int i = 1; i < MM_WARPS; ++i) r = op[k] == 0 ? r + sh.red[k][i] : (op[k] == 1 ? fmin(r, sh.red[k][i]) : fmax(r, sh.red[k][i]));
        v[k] = r;
    }
    __syncthreads();
}

// in-place Cholesky of the K x K lower triangle of sh.G (row stride MM_MAXKP) by warp 0; rel_pivot: a pivot below
// rel_pivot * its diagonal counts as rank deficiency.  Returns 0 when fine (same value in every thread of the block).
__device__ int block_chol(MmShared& sh, int K, double rel_pivot) {
    if (threadIdx.x == 0) sh.flag = 0;
    __syncthreads();
    if (threadIdx.x < 32) {
        const int lane = threadIdx.x;
        bool bad = false;
        for (int j = 0; j < K && !bad; ++j) {
            const double d0 = sh.G[j * MM_MAXKP + j];
            // the diagonal has been updated by the trailing updates of the previous columns; rank test against the original
            const double dorig = sh.diag[j];
            if (!(d0 > rel_pivot * dorig) || !(d0 > 0.0) || !isfinite(d0)) { bad = true; break; }
            const double d = sqrt(d0);
            __syncwarp();
            if (lane == 0) sh.G[j * MM_MAXKP + j] = d;
            for (int i = j + 1 + lane; i < K; i += 32) sh.G[i * MM_MAXKP + j] /= d;
            __syncwarp();
            for (int i = j + 1; i < K; ++i) {
                const double lij = sh.G[i * MM_MAXKP + j];
                for (int k = j + 1 + lane; k <= i; k += 32) sh.G[i * MM_MAXKP + k] -= lij * sh.G[k * MM_MAXKP + j];
            }
            __syncwarp();
        }
        if (bad && lane == 0) sh.flag = 1;
    }
    __syncthreads();
    return sh.flag;
}

// b <- (L L')^{-1} b on warp 0, column-oriented substitutions; b in shared memory
__device__ void block_chol_solve(MmShared& sh, int K, double* b) {
    if (threadIdx.x < 32) {
        const int lane = threadIdx.x;
        for (int k = 0; k < K; ++k) {
            __syncwarp();
            const double bk = b[k] / sh.G[k * MM_MAXKP + k];
            __syncwarp();
            if (lane == 0) b[k] = bk;
            for (int i = k + 1 + lane; i < K; i += 32) b[i] -= sh.G[i * MM_MAXKP + k] * bk;
        }
        for (int k = K - 1; k >= 0; --k) {
            __syncwarp();
            const double bk = b[k] / sh.G[k * MM_MAXKP + k];
            __syncwarp();
            if (lane == 0) b[k] = bk;
            for (int i = lane; i < k; i += 32) b[i] -= sh.G[k * MM_MAXKP + i] * bk;
        }
    }
    __syncthreads();
}

struct MmKernelArgs {
    const double* X[2]; const void* C[2]; long long n[2], n_pad[2];
    int ldx, K, count_bytes;
    int sims; long long slots;                   // passes of this batch (slot = column of the multiplicity matrix)
    const double* taus;                          // [slots][sims]
    double* state; long long state_stride;       // per block: 6 vectors of state_stride doubles
    double* betas;                               // [slots][sims][2][K]: problem-major (problem = (slot sims + sim) 2 + group)
    int* info;                                   // [slots][sims][2]: status | iterations << 8 | min(candidates, 65535) << 16
    int* counter;                                // work queue
    long long p_begin, p_end;                    // this launch solves problems [p_begin, p_end) (mode R: the rank's share)
};

constexpr int QR_VERTEX = 0, QR_APPROX = 1, QR_FAILED = 2;

// 1/x, 1/s and the scaling q = 1 / (z/x + w/s) = x s / (z s + w x) from ONE division (an FP64 division is a dozen
// instructions with a slow-path branch; the row algebra of a sweep used three to five of them)
__device__ __forceinline__ void recips(double x, double s, double z, double w, double& rx, double& rs, double& q) {
    const double xs = x * s, d = z * s + w * x;
    const double t = 1.0 / (xs * d);
    rx = s * d * t; rs = x * d * t; q = xs * xs * t;
}

// The per-row iterate of a problem: six vectors in the block's slab (x, s = u - x, z, w, the affine dx, the corrected dx).
struct St { double x, s, z, w, a, c; };
enum : int { LX = 1, LS = 2, LZ = 4, LW = 8, LA = 16, LC = 32 };
struct StPtr { double *x, *s, *z, *w, *a, *c; long long len, n_pad; };      // len: doubles per vector of the slab; n_pad: rows of the design incl. zero pad rows
template <int MASK>
__device__ __forceinline__ void st_load(St& v, const StPtr& p, long long i, bool in) {
    MM_CHECK(!in || (i >= 0 && i < p.len));
    v.x = (MASK & LX) && in ? p.x[i] : 0.0;
    v.s = (MASK & LS) && in ? p.s[i] : 0.0;
    v.z = (MASK & LZ) && in ? p.z[i] : 0.0;
    v.w = (MASK & LW) && in ? p.w[i] : 0.0;
    v.a = (MASK & LA) && in ? p.a[i] : 0.0;
    v.c = (MASK & LC) && in ? p.c[i] : 0.0;
}

// One sweep over the rows of the problem's group.  Per 32-row block (a warp owns every eighth one):
//   DOT  : t_i = x_i'vec (vec: shared, zero beyond column K-1) and, with NEEDY, the outcome y_i, from the fragment loads
//   rowf : lane-per-row work on the row's iterate (the vectors named by MASK, loaded one block ahead so that their
//          latency hides under the previous block's work); returns the weight q_i and the extra column v_i
//   GRAM : 1 = acc += sum_i q_i [x_i | v_i][x_i | v_i]' (upper-triangle tiles),
//          3 = two right-hand sides at once: acc[jt] = (X'Q v, X'Q v2) for the design columns of tile jt
template <int K8, bool DOT, int GRAM, int MASK, bool NEEDY, typename RowF>
__device__ __forceinline__ void sweep(MmShared& sh, const double* __restrict__ X, int ldx, long long n, int K, const double* vec,
                                      const StPtr& sp, double (&acc)[K8 * (K8 + 1) / 2][2], RowF&& rowf) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int cg = lane >> 2, r4 = lane & 3;
    const int wb = w * 32;
    constexpr int UNR = K8 <= 2 ? 8 : (K8 <= 4 ? 4 : 2);
    if (GRAM) {
#pragma unroll
        for (int t = 0; t < K8 * (K8 + 1) / 2; ++t) { acc[t][0] = 0.0; acc[t][1] = 0.0; }
    }
    double vv[K8];
    if (DOT && vec) {
#pragma unroll
        for (int t = 0; t < K8; ++t) vv[t] = vec[8 * t + cg];
    }
    // The outcome column K sits in the last fragment tile (K8 = K / 8 + 1), at lane group cy; the tiles before it are
    // full, only the last one is masked (columns beyond K, possibly beyond the row).  Rows need no mask: a 32-row block
    // that starts below n ends below n_pad (a multiple of 32), and the design's pad rows are zero.
    const int cy = K & 7;
    const bool clast = 8 * (K8 - 1) + cg <= K;
    St nxt;
    st_load<MASK>(nxt, sp, (long long)wb + lane, (long long)wb + lane < n);
    for (long long base = (long long)wb; base < n; base += MM_THREADS) {
        const St cur = nxt;
        st_load<MASK>(nxt, sp, base + MM_THREADS + lane, base + MM_THREADS + lane < n);
        const double* xp = X + (base + r4) * ldx + cg;
        MM_CHECK(base + 31 < sp.n_pad);          // the block's 32 design rows exist (pad rows are zero)
        double t_own = 0.0;                       // x_i'vec of this lane's own row (butterfly path)
        if (DOT && K8 <= 2) {
            // All eight 4-row steps first (16 fragment loads in flight), then ONE butterfly over the eight lane groups that
            // hold the pieces of a row's dot product: each exchange halves the number of sums a lane carries (4 + 2 + 1
            // shuffles instead of 3 per step), and the step a lane ends up with is its own lane group -- lane l holds the
            // product of row l, no trip through shared memory.
            double part[8];
#pragma unroll
            for (int ks = 0; ks < 8; ++ks) {
                double xv[K8];
                const double* xr = xp + (size_t)(4 * ks) * ldx;
#pragma unroll
                for (int t = 0; t < K8 - 1; ++t) xv[t] = __ldg(xr + 8 * t);
                xv[K8 - 1] = clast ? __ldg(xr + 8 * (K8 - 1)) : 0.0;
                double p = 0.0;
                if (vec) {
#pragma unroll
                    for (int t = 0; t < K8; ++t) p = fma(xv[t], vv[t], p);
                }
                part[ks] = p;
                if (NEEDY && cg == cy) sh.ys[wb + 4 * ks + r4] = xv[K8 - 1];
            }
            if (vec) {
                const bool hi = cg & 4, mid = cg & 2, lo = cg & 1;
                double p4[4], p2[2];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const double recv = __shfl_xor_sync(0xffffffffu, hi ? part[j] : part[j + 4], 16);
                    p4[j] = (hi ? part[j + 4] : part[j]) + recv;
                }
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const double recv = __shfl_xor_sync(0xffffffffu, mid ? p4[j] : p4[j + 2], 8);
                    p2[j] = (mid ? p4[j + 2] : p4[j]) + recv;
                }
                const double recv = __shfl_xor_sync(0xffffffffu, lo ? p2[0] : p2[1], 4);
                t_own = (lo ? p2[1] : p2[0]) + recv;
            }
            if (NEEDY) __syncwarp();
        } else if (DOT) {
#pragma unroll UNR
            for (int ks = 0; ks < 8; ++ks) {
                double xv[K8];
                const double* xr = xp + (size_t)(4 * ks) * ldx;
#pragma unroll
                for (int t = 0; t < K8 - 1; ++t) xv[t] = __ldg(xr + 8 * t);
                xv[K8 - 1] = clast ? __ldg(xr + 8 * (K8 - 1)) : 0.0;
                double part = 0.0;
                const double yv = xv[K8 - 1];
                if (vec) {
#pragma unroll
                    for (int t = 0; t < K8; ++t) part = fma(xv[t], vv[t], part);
                    part += __shfl_xor_sync(0xffffffffu, part, 4);
                    part += __shfl_xor_sync(0xffffffffu, part, 8);
                    part += __shfl_xor_sync(0xffffffffu, part, 16);
                }
                if (cg == 0) sh.ts[wb + 4 * ks + r4] = part;
                if (NEEDY && cg == cy) sh.ys[wb + 4 * ks + r4] = yv;
            }
            __syncwarp();
            t_own = sh.ts[wb + lane];
        }
        {
            const long long i = base + lane;
            double q = 0.0, v = 0.0, v2 = 0.0;
            rowf(i, i < n, t_own, (DOT && NEEDY) ? sh.ys[wb + lane] : 0.0, cur, q, v, v2);
            if (GRAM) { sh.qs[wb + lane] = q; sh.vs[wb + lane] = v; }
            if (GRAM == 3) sh.vs2[wb + lane] = v2;
        }
        if (GRAM) {
            __syncwarp();
#pragma unroll UNR
            for (int ks = 0; ks < 8; ++ks) {
                double xv[K8];
                const double* xr = xp + (size_t)(4 * ks) * ldx;
#pragma unroll
                for (int t = 0; t < K8 - 1; ++t) xv[t] = __ldg(xr + 8 * t);
                xv[K8 - 1] = clast ? __ldg(xr + 8 * (K8 - 1)) : 0.0;
                const double q = sh.qs[wb + 4 * ks + r4];
                const double ve = sh.vs[wb + 4 * ks + r4];
                if (cg == cy) xv[K8 - 1] = ve;       // the extra column replaces the outcome column
                if (GRAM == 1) {
                    int tt = 0;
#pragma unroll
                    for (int jt = 0; jt < K8; ++jt) {
                        const double a = q * xv[jt];
#pragma unroll
                        for (int lt = jt; lt < K8; ++lt, ++tt) dmma884(acc[tt][0], acc[tt][1], a, xv[lt]);
                    }
                } else {
                    // B tile: column 0 = v, column 1 = v2 of the row (lane holds B[row r4][column cg])
                    const double b = cg == 0 ? ve : (cg == 1 ? sh.vs2[wb + 4 * ks + r4] : 0.0);
#pragma unroll
                    for (int jt = 0; jt < K8; ++jt) dmma884(acc[jt][0], acc[jt][1], q * xv[jt], b);
                }
            }
        }
        __syncwarp();
    }
}

// warps add their accumulator tiles into shared memory one after the other (fixed order)
template <int K8>
__device__ void gram_to_shared(MmShared& sh, const double (&acc)[K8 * (K8 + 1) / 2][2]) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < MM_MAXKP * MM_MAXKP; i += MM_THREADS) sh.G[i] = 0.0;
    __syncthreads();
    for (int turn = 0; turn < MM_WARPS; ++turn) {
        if (w == turn) {
            int tt = 0;
#pragma unroll
            for (int jt = 0; jt < K8; ++jt)
#pragma unroll
                for (int lt = jt; lt < K8; ++lt, ++tt) {
                    const int row = 8 * jt + (lane >> 2), col = 8 * lt + 2 * (lane & 3);
                    sh.G[row * MM_MAXKP + col] += acc[tt][0];
                    sh.G[row * MM_MAXKP + col + 1] += acc[tt][1];
                }
        }
        __syncthreads();
    }
}
// warps add their partial right-hand sides one after the other (fixed order).  GRAM == 3: lanes with (lane & 3) == 0 hold columns 0 and 1 of D rows 8 jt + lane / 4 -> out1, out2 [0..K)
template <int K8>
__device__ void rhs2_to_shared(MmShared& sh, const double (&acc)[K8 * (K8 + 1) / 2][2], int K, double* out1, double* out2) {
    (void)sh;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (threadIdx.x < MM_MAXKP) { out1[threadIdx.x] = 0.0; out2[threadIdx.x] = 0.0; }
    __syncthreads();
    for (int turn = 0; turn < MM_WARPS; ++turn) {
        if (w == turn && (lane & 3) == 0) {
#pragma unroll
            for (int jt = 0; jt < K8; ++jt) {
                const int row = 8 * jt + (lane >> 2);
                if (row < K) { out1[row] += acc[jt][0]; out2[row] += acc[jt][1]; }
            }
        }
        __syncthreads();
    }
}

// after gram_to_shared: mirror the upper triangle into the lower one, keep the diagonal (rank test) and column K (rhs)
__device__ void gram_finish(MmShared& sh, int K) {
    for (int idx = threadIdx.x; idx < K * K; idx += MM_THREADS) {
        const int i = idx / K, k = idx % K;
        if (k < i) sh.G[i * MM_MAXKP + k] = sh.G[k * MM_MAXKP + i];
    }
    if (threadIdx.x < MM_MAXKP) {
        sh.rhs[threadIdx.x] = threadIdx.x < K ? sh.G[threadIdx.x * MM_MAXKP + K] : 0.0;
        sh.diag[threadIdx.x] = threadIdx.x < K ? sh.G[threadIdx.x * MM_MAXKP + threadIdx.x] : 0.0;
    }
    __syncthreads();
}

__device__ __forceinline__ double dd_residual(const double* __restrict__ xrow, double y, const double* b, int K) {
    double hi = y, lo = 0.0;
    for (int j = 0; j < K; ++j) {
        const double x = __ldg(xrow + j), p = x * b[j], e = fma(x, b[j], -p);
        const double t = hi - p, bb = t - hi;
        const double err = (hi - (t - bb)) + (-p - bb);
        hi = t; lo += err - e;
    }
    return hi + lo;
}

template <int K8, int MINB>
__global__ void __launch_bounds__(MM_THREADS, MINB) mm_qr_kernel(const MmKernelArgs a) {
    __shared__ MmShared sh;
    constexpr int NT = K8 * (K8 + 1) / 2;
    const int tid = threadIdx.x;
    const int K = a.K, ldx = a.ldx;
    double* const st = a.state + (size_t)blockIdx.x * 6 * a.state_stride;
    double *st_x = st, *st_s = st + a.state_stride, *st_z = st + 2 * a.state_stride, *st_w = st + 3 * a.state_stride,
           *st_dxa = st + 4 * a.state_stride, *st_dxc = st + 5 * a.state_stride;
    StPtr sp{st_x, st_s, st_z, st_w, st_dxa, st_dxc, a.state_stride, 0};
    MM_CHECK(a.n[0] <= a.state_stride && a.n[1] <= a.state_stride && a.n[0] <= a.n_pad[0] && a.n[1] <= a.n_pad[1]);
    double acc[NT][2];

    for (;;) {
        __syncthreads();
        if (tid == 0) sh.prob = atomicAdd(a.counter, 1);
        __syncthreads();
        const long long p = a.p_begin + sh.prob;
        if (p >= a.p_end) break;
        const int g = (int)(p & 1);
        const long long rest = p >> 1;
        const int sim = (int)(rest % a.sims);
        const long long slot = rest / a.sims;
        const long long n = a.n[g];
        sp.n_pad = a.n_pad[g];
        const double* __restrict__ X = a.X[g];
        const long long out_idx = p;
        double tau = a.taus[slot * a.sims + sim];
        tau = fmin(fmax(tau, 1e-6), 1.0 - 1e-6);
        const unsigned char* C8 = (const unsigned char*)a.C[g] + ((size_t)(slot / BM) * a.n_pad[g]) * BM * a.count_bytes;
        const int ccol = (int)(slot % BM);
        for (int j = tid; j < 4 * MM_MAXKP; j += MM_THREADS) (&sh.vec[0][0])[j] = 0.0;
        __syncthreads();

        int status = QR_FAILED, iters = 0, ncand = 0;
        // ---- starting point: u = multiplicities, yd = weighted least-squares fit of the cost -y ----
        double yscale, scale, usum, nact;
        {
            double r4[4] = {0.0, 0.0, 0.0, 0.0};       // max |y|, sum u |y|, sum u, active rows
            sweep<K8, true, 1, 0, true>(sh, X, ldx, n, K, nullptr, sp, acc, [&](long long i, bool in, double, double y, const St&, double& q, double& v, double&) {
                double u = 0.0;
                MM_CHECK(!in || i < a.n_pad[g]);
                if (in) u = a.count_bytes == 1 ? (double)C8[(size_t)i * BM + ccol] : (double)((const unsigned short*)C8)[(size_t)i * BM + ccol];
                if (in) st_s[i] = u;
                if (u > 0.0) { r4[0] = fmax(r4[0], fabs(y)); r4[1] += u * fabs(y); r4[2] += u; r4[3] += 1.0; }
                q = u; v = -y;
            });
            const int ops[4] = {2, 0, 0, 0};
            block_reduce<4>(sh, r4, ops);
            yscale = r4[0]; scale = r4[1]; usum = r4[2]; nact = r4[3];
        }
        bool dead = nact < 1.0;
        if (yscale == 0.0) yscale = 1.0;
        if (scale == 0.0) scale = yscale;
        gram_to_shared<K8>(sh, acc);
        gram_finish(sh, K);
        if (!dead && block_chol(sh, K, MM_RANK_TOL)) dead = true;
        double gap = 0.0;
        if (!dead) {
            if (tid < MM_MAXKP) sh.vec[0][tid] = sh.rhs[tid];
            __syncthreads();
            block_chol_solve(sh, K, sh.vec[0]);
            double r1[1] = {0.0};
            sweep<K8, true, 0, LS, true>(sh, X, ldx, n, K, sh.vec[0], sp, acc, [&](long long i, bool in, double t, double y, const St& c, double&, double&, double&) {
                if (!in) return;
                const double u = c.s;
                const double r = -y - t;
                st_z[i] = r;
                if (u > 0.0) r1[0] += u * fabs(r);
            });
            const int ops1[1] = {0};
            block_reduce<1>(sh, r1, ops1);
            const double delta = 0.01 * (r1[0] / usum) + 1e-10 * yscale;
            const double tol = 1e-12 * scale;
            // ---- Mehrotra predictor-corrector iterations ----
            bool first = true, corr = false, last = false;
            double ap = 0.0, ad = 0.0, mu = 0.0;
            for (;;) {
                // P1: apply the previous step (or build the starting point), new q and r, gap, Newton matrix
                double rg[1] = {0.0};
                sweep<K8, false, 1, LX | LS | LZ | LW | LA | LC, false>(sh, X, ldx, n, K, nullptr, sp, acc,
                                                                 [&](long long i, bool in, double, double, const St& c, double& q, double& v, double&) {
                    if (!in) return;
                    double x, s, z, w;
                    if (first) {
                        const double u = c.s;
                        if (!(u > 0.0)) { st_x[i] = 0.0; st_s[i] = 0.0; return; }
                        const double r0 = c.z;
                        x = (1.0 - tau) * u; s = u - x;
                        z = fmax(r0, 0.0) + delta; w = fmax(-r0, 0.0) + delta;
                    } else {
                        x = c.x; s = c.s;
                        if (!(x + s > 0.0)) return;
                        z = c.z; w = c.w;
                        double rx, rs, qold;
                        recips(x, s, z, w, rx, rs, qold);
                        const double dxa = c.a;
                        const double dza = -z * (1.0 + dxa * rx), dwa = -w * (1.0 - dxa * rs);
                        double dx = dxa, dz = dza, dw = dwa;
                        if (corr) {
                            dx = c.c;
                            dz = (mu - dxa * dza) * rx - z - z * rx * dx;
                            dw = (mu + dxa * dwa) * rs - w + w * rs * dx;
                        }
                        x += ap * dx; s -= ap * dx; z += ad * dz; w += ad * dw;
                    }
                    MM_CHECK(i < sp.len);
                    st_x[i] = x; st_s[i] = s; st_z[i] = z; st_w[i] = w;
                    rg[0] += z * x + w * s;
                    q = (x * s) / (z * s + w * x);
                    v = z - w;
                });
                block_reduce<1>(sh, rg, ops1);
                gap = rg[0];
                first = false;
                if (!(gap > tol) || !isfinite(gap) || iters >= 100 || last) break;
                ++iters;
                gram_to_shared<K8>(sh, acc);
                gram_finish(sh, K);
                if (block_chol(sh, K, MM_RANK_TOL)) break;
                if (tid < MM_MAXKP) sh.vec[1][tid] = sh.rhs[tid];
                __syncthreads();
                block_chol_solve(sh, K, sh.vec[1]);
                // P2: affine step, ratio test, the sums that give the gap after the step -- and the two vectors the
                // corrector's right-hand side is linear in: X'Q xi = X'Q r + mu X'Q (1/s - 1/x) + X'Q (dx dz / x + dx dw / s),
                // so the corrector needs no sweep of its own for its right-hand side
                double r2[4] = {0.0, 0.0, 0.0, 0.0};        // 1 / primal step limit, 1 / dual step limit, S1, S3
                sweep<K8, true, 3, LX | LS | LZ | LW, false>(sh, X, ldx, n, K, sh.vec[1], sp, acc,
                                                      [&](long long i, bool in, double t, double, const St& c, double& q, double& v, double& v2) {
                    if (!in) return;
                    const double x = c.x, s = c.s;
                    if (!(x + s > 0.0)) return;
                    const double z = c.z, w = c.w;
                    double rx, rs;
                    recips(x, s, z, w, rx, rs, q);
                    const double r = z - w;
                    const double dx = q * (t - r);
                    const double dz = -z * (1.0 + dx * rx), dw = -w * (1.0 - dx * rs);
                    st_dxa[i] = dx;
                    v = rs - rx;
                    v2 = dx * dz * rx + dx * dw * rs;
                    // ratio tests: the largest step keeping x, s, z, w positive (as reciprocals: no division per row)
                    r2[0] = fmax(r2[0], fmax(-dx * rx, dx * rs));
                    r2[1] = fmax(r2[1], fmax(1.0 + dx * rx, 1.0 - dx * rs));
                    r2[2] += dx * r; r2[3] += dx * (dz - dw);
                });
                const int ops2[4] = {2, 2, 0, 0};
                block_reduce<4>(sh, r2, ops2);
                ap = fmin(0.99995 / r2[0], 1.0); ad = fmin(0.99995 / r2[1], 1.0);
                corr = fmin(ap, ad) < 1.0;
                if (corr) {
                    const double gaff = gap + ap * r2[2] + ad * (-gap - r2[2]) + ap * ad * r2[3];
                    const double ratio = gaff / gap;
                    mu = gap * ratio * ratio * ratio / (2.0 * nact);
                    if (!(mu >= 0.0)) mu = 0.0;
                    rhs2_to_shared<K8>(sh, acc, K, sh.vec[1], sh.ah);
                    if (tid < K) sh.vec[1][tid] = sh.rhs[tid] + mu * sh.vec[1][tid] + sh.ah[tid];
                    __syncthreads();
                    block_chol_solve(sh, K, sh.vec[1]);
                    // P4: corrected step and its ratio test
                    double r3[2] = {0.0, 0.0};
                    sweep<K8, true, 0, LX | LS | LZ | LW | LA, false>(sh, X, ldx, n, K, sh.vec[1], sp, acc,
                                                               [&](long long i, bool in, double t, double, const St& c, double&, double&, double&) {
                        if (!in) return;
                        const double x = c.x, s = c.s;
                        if (!(x + s > 0.0)) return;
                        const double z = c.z, w = c.w, dxa = c.a;
                        double rx, rs, q;
                        recips(x, s, z, w, rx, rs, q);
                        const double dza = -z * (1.0 + dxa * rx), dwa = -w * (1.0 - dxa * rs);
                        const double xi = (z - w) + mu * (rs - rx) + dxa * dza * rx + dxa * dwa * rs;
                        const double dx = q * (t - xi);
                        const double dz = (mu - dxa * dza) * rx - z - z * rx * dx;
                        const double dw = (mu + dxa * dwa) * rs - w + w * rs * dx;
                        st_dxc[i] = dx;
                        r3[0] = fmax(r3[0], fmax(-dx * rx, dx * rs));
                        const double tzw = 1.0 / (z * w);
                        r3[1] = fmax(r3[1], fmax(-dz * w * tzw, -dw * z * tzw));
                    });
                    const int ops3[2] = {2, 2};
                    block_reduce<2>(sh, r3, ops3);
                    ap = fmin(0.99995 / r3[0], 1.0); ad = fmin(0.99995 / r3[1], 1.0);
                }
                // one step length for both iterates: separate ones lose centrality at the extreme quantiles (tau near 0.01 /
                // 0.99: iteration counts of 100 and more instead of ~20)
                ap = ad = fmin(ap, ad);
                if (tid < K) sh.vec[0][tid] += ad * sh.vec[1][tid];
                __syncthreads();
                if (ap < 1e-12) last = true;
            }
            if (isfinite(gap) && gap <= 1e-7 * scale) status = QR_APPROX;
        }

        if (status == QR_APPROX) {
            // ---- polish to the LP's vertex ----
            if (tid < MM_MAXKP) sh.vec[2][tid] = tid < K ? -sh.vec[0][tid] : 0.0;
            __syncthreads();
            int cnt[12];
#pragma unroll
            for (int j = 0; j < 12; ++j) cnt[j] = 0;
            sweep<K8, true, 0, LX | LS, true>(sh, X, ldx, n, K, sh.vec[2], sp, acc, [&](long long i, bool in, double t, double y, const St& c, double&, double&, double&) {
                if (!in) return;
                if (!(c.x + c.s > 0.0)) return;
                const double res = y - t;
                st_dxa[i] = res;
                double th = yscale * 1e-3;
#pragma unroll
                for (int j = 0; j < 12; ++j, th *= 0.1) cnt[j] += fabs(res) < th;
            });
            {
                const int lane = tid & 31, w = tid >> 5;
#pragma unroll
                for (int j = 0; j < 12; ++j) {
                    int c = cnt[j];
                    for (int o = 16; o > 0; o >>= 1) c += __shfl_down_sync(0xffffffffu, c, o);
                    if (lane == 0) sh.ired[j][w] = c;
                }
                __syncthreads();
#pragma unroll
                for (int j = 0; j < 12; ++j) { int c = 0; for (int i = 0; i < MM_WARPS; ++i) c += sh.ired[j][i]; cnt[j] = c; }
                __syncthreads();
            }
            int jstar = -1;                              // thresholds yscale 10^-(3 + j), j = 0..11
            for (int j = 11; j >= 0; --j) if (cnt[j] >= K) { jstar = j; break; }
            long long m_prev = -1;
            for (int jt = jstar; jt >= 1 && status != QR_VERTEX; --jt) {
                const double thr = yscale * pow(10.0, -(double)(jt + 2));
                // candidates in row order (the first MM_MAXC of them)
                long long m_total = 0;
                for (long long base = 0; base < n; base += MM_THREADS) {
                    const long long i = base + tid;
                    bool c = false;
                    if (i < n) { const double x = st_x[i], s = st_s[i]; c = (x + s > 0.0) && fabs(st_dxa[i]) < thr; }
                    const unsigned m = __ballot_sync(0xffffffffu, c);
                    if ((tid & 31) == 0) sh.wmask[tid >> 5] = m;
                    __syncthreads();
                    long long off = m_total;
                    int tot = 0;
                    for (int w = 0; w < MM_WARPS; ++w) { const int pc = __popc(sh.wmask[w]); if (w < (tid >> 5)) off += pc; tot += pc; }
                    if (c) { const long long pos = off + __popc(m & ((1u << (tid & 31)) - 1u)); if (pos < MM_MAXC) { MM_CHECK(pos >= 0 && i < n); sh.cand[pos] = (int)i; } }
                    m_total += tot;
                    __syncthreads();
                }
                if (m_total == m_prev) continue;
                m_prev = m_total;
                const int m = (int)(m_total < MM_MAXC ? m_total : MM_MAXC);
                // normal equations over the candidates
                for (int idx = tid; idx < K * K; idx += MM_THREADS) {
                    const int i = idx / K, k = idx % K;
                    if (k > i) continue;
                    double sum = 0.0;
                    for (int h = 0; h < m; ++h) { const double* xr = X + (size_t)sh.cand[h] * ldx; sum = fma(__ldg(xr + i), __ldg(xr + k), sum); }
                    sh.G[i * MM_MAXKP + k] = sum;
                }
                __syncthreads();
                if (tid < K) sh.diag[tid] = sh.G[tid * MM_MAXKP + tid];
                if (tid < MM_MAXKP) sh.vec[3][tid] = sh.vec[2][tid];
                __syncthreads();
                if (block_chol(sh, K, 1e-12)) continue;
                for (int round = 0; round < 3; ++round) {
                    for (int h = tid; h < m; h += MM_THREADS) {
                        const double* xr = X + (size_t)sh.cand[h] * ldx;
                        sh.ah[h] = dd_residual(xr, __ldg(xr + K), sh.vec[3], K);
                    }
                    __syncthreads();
                    if (tid < K) {
                        double sum = 0.0;
                        for (int h = 0; h < m; ++h) sum = fma(__ldg(X + (size_t)sh.cand[h] * ldx + tid), sh.ah[h], sum);
                        sh.vec[1][tid] = sum;
                    }
                    __syncthreads();
                    block_chol_solve(sh, K, sh.vec[1]);
                    if (tid < K) sh.vec[3][tid] += sh.vec[1][tid];
                    __syncthreads();
                }
                double bad[1] = {0.0};
                sweep<K8, true, 0, LX | LS | LA, true>(sh, X, ldx, n, K, sh.vec[3], sp, acc, [&](long long, bool in, double t, double y, const St& c, double&, double&, double&) {
                    if (!in) return;
                    if (!(c.x + c.s > 0.0)) return;
                    const double res0 = c.a, res = y - t;
                    if (fabs(res0) < thr) { if (fabs(res) > 1e-11 * yscale) bad[0] += 1.0; }
                    else if ((res > 0.0) != (res0 > 0.0)) bad[0] += 1.0;
                });
                const int ops1[1] = {0};
                block_reduce<1>(sh, bad, ops1);
                if (bad[0] == 0.0) {
                    if (tid < K) sh.vec[2][tid] = sh.vec[3][tid];
                    __syncthreads();
                    status = QR_VERTEX;
                    ncand = (int)(m_total < 65535 ? m_total : 65535);
                }
            }
        }
        MM_CHECK(out_idx >= 0 && out_idx < 2 * a.slots * a.sims && slot < a.slots);
        if (tid < K) a.betas[(size_t)out_idx * K + tid] = status == QR_FAILED ? __longlong_as_double(0x7ff8000000000000LL) : sh.vec[2][tid];
        if (tid == 0) a.info[out_idx] = status | (iters << 8) | (ncand << 16);
    }
}


// ---- the random quantiles and simulated rows of every pass (native stream) ----
// quantile_decomposition.rs:221-225: tau ~ Uniform(0.01, 0.99), `simulations` per pass, shared by both groups;
// :251-255: one uniformly drawn row of each (resampled) group per simulation -- a row of the resampled frame is original
// row i with probability c_i / n, drawn here by rejection against the pass's multiplicity column.
constexpr uint32_t MM_STREAM_TAU = 0x4d4d5441u, MM_STREAM_ROW = 0x4d4d524fu;
__global__ void mm_streams_kernel(long long slots, int sims, long long pass0, int first_slot, uint64_t seed, const void* Ca, const void* Cb,
                                  long long na, long long nb, long long npa, long long npb, int count_bytes,
                                  double* taus, uint32_t* rows_a, uint32_t* rows_b) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= slots * sims) return;
    const long long slot = idx / sims; const int s = (int)(idx % sims);
    // global pass id: 0 = point estimates (slot 0 of the first batch), r + 1 = replicate r
    const unsigned long long pass = (first_slot && slot == 0) ? 0ull : (unsigned long long)(pass0 + slot);
    const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    {
        const Philox4 r = philox4x32_10((uint32_t)pass, (uint32_t)(pass >> 32), (uint32_t)s, MM_STREAM_TAU, k0, k1);
        const double u = (double)(((unsigned long long)r.x << 21) ^ (unsigned long long)(r.y >> 11)) * (1.0 / 9007199254740992.0);
        taus[idx] = 0.01 + (0.99 - 0.01) * u;
    }
    for (int g = 0; g < 2; ++g) {
        const long long n = g ? nb : na, np = g ? npb : npa;
        const unsigned char* C = (const unsigned char*)(g ? Cb : Ca) + ((size_t)(slot / BM) * np) * BM * count_bytes;
        const int col = (int)(slot % BM);
        uint32_t row = 0;
        for (uint32_t t = 0;; ++t) {
            const Philox4 r = philox4x32_10((uint32_t)pass, (uint32_t)(pass >> 32) ^ (t << 8), (uint32_t)s | ((uint32_t)g << 31), MM_STREAM_ROW, k0, k1);
            const unsigned long long u64 = ((unsigned long long)r.x << 32) | r.y;
            row = (uint32_t)__umul64hi(u64, (unsigned long long)n);
            const unsigned c = count_bytes == 1 ? C[(size_t)row * BM + col] : ((const unsigned short*)C)[(size_t)row * BM + col];
            if ((r.z & 63u) < (c < 64u ? c : 64u) || t > 100000u) break;        // accept with probability min(c, 64) / 64
        }
        (g ? rows_b : rows_a)[idx] = row;
    }
}

// ---- simulation and effects of a pass: quantile_decomposition.rs:244-277 ----
// One block per pass.  betas [slots][sims][2][K], info [slots][sims][2]; rows_* [slots][sims] = the simulated ORIGINAL
// row of each group for the i-th pairing; stats [slots][3 nq]; status [slots].
__global__ void __launch_bounds__(256) mm_effects_kernel(const double* Xa, const double* Xb, int ldx, int K, int sims, long long slots,
                                                         const double* betas, const int* info, const uint32_t* rows_a, const uint32_t* rows_b,
                                                         int nq, const double* quantiles, double* stats, int* status, int* nsucc) {
    extern __shared__ double dyn[];
    int pow2 = 1;
    while (pow2 < sims) pow2 <<= 1;
    double* yv = dyn;                                   // [3][pow2]
    int* la = reinterpret_cast<int*>(dyn + 3 * (size_t)pow2);   // [sims] successful fits of A, in order
    int* lb = la + sims;
    __shared__ int s_na, s_nb;
    const long long slot = blockIdx.x;
    const int tid = threadIdx.x;
    const int* ii = info + (size_t)slot * sims * 2;          // [sims][2]
    if (tid == 0) {          // filter_map(.ok()): order-preserving compaction (sims is a few hundred)
        int ca = 0, cb = 0;
        for (int s = 0; s < sims; ++s) {
            if ((ii[2 * s] & 0xff) != QR_FAILED) la[ca++] = s;
            if ((ii[2 * s + 1] & 0xff) != QR_FAILED) lb[cb++] = s;
        }
        s_na = ca; s_nb = cb;
    }
    __syncthreads();
    const int ns = s_na < s_nb ? s_na : s_nb;
    const int S = 3 * nq;
    if (tid == 0 && nsucc) nsucc[slot] = ns;
    if (s_na < sims / 2 || s_nb < sims / 2) {           // :238-242
        for (int j = tid; j < S; j += blockDim.x) stats[(size_t)slot * S + j] = __longlong_as_double(0x7ff8000000000000LL);
        if (tid == 0) status[slot] = OB_ERR_NALGEBRA;
        return;
    }
    const double inf = __longlong_as_double(0x7ff0000000000000LL);
    for (int i = tid; i < pow2; i += blockDim.x) {
        double yaa = inf, ybb = inf, yab = inf;
        if (i < ns) {
            MM_CHECK(la[i] < sims && lb[i] < sims);
            const double* xa = Xa + (size_t)rows_a[(size_t)slot * sims + i] * ldx;
            const double* xb = Xb + (size_t)rows_b[(size_t)slot * sims + i] * ldx;
            const double* ba = betas + (((size_t)slot * sims + la[i]) * 2 + 0) * K;
            const double* bb = betas + (((size_t)slot * sims + lb[i]) * 2 + 1) * K;
            yaa = ybb = yab = 0.0;
            for (int j = 0; j < K; ++j) { yaa += xa[j] * ba[j]; ybb += xb[j] * bb[j]; yab += xa[j] * bb[j]; }
        }
        yv[i] = yaa; yv[pow2 + i] = ybb; yv[2 * pow2 + i] = yab;
    }
    __syncthreads();
    for (int k = 2; k <= pow2; k <<= 1)                 // bitonic sort of the three vectors
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = tid; t < 3 * pow2; t += blockDim.x) {
                const int v = t / pow2, i = t % pow2, l = i ^ j;
                if (l > i) {
                    double* y = yv + (size_t)v * pow2;
                    const bool up = (i & k) == 0;
                    const double a = y[i], b = y[l];
                    if ((a > b) == up) { y[i] = b; y[l] = a; }
                }
            }
            __syncthreads();
        }
    for (int q = tid; q < nq; q += blockDim.x) {
        double val[3];
        for (int v = 0; v < 3; ++v) {
            if (ns == 0) { val[v] = 0.0; continue; }    // empirical_quantile of an empty vector (:165-167)
            int index = (int)((double)ns * quantiles[q]);
            if (index > ns - 1) index = ns - 1;
            if (index < 0) index = 0;
            val[v] = yv[(size_t)v * pow2 + index];
        }
        stats[(size_t)slot * S + 3 * q + 0] = val[0] - val[1];      // gap             = q_aa - q_bb
        stats[(size_t)slot * S + 3 * q + 1] = val[2] - val[1];      // characteristics = q_ab - q_bb
        stats[(size_t)slot * S + 3 * q + 2] = val[0] - val[2];      // coefficients    = q_aa - q_ab
    }
    if (tid == 0) status[slot] = OB_OK;
}

}  // namespace

int mm_state_vectors() { return 6; }

// resident blocks per SM the kernel is compiled for (register budget 65536 / (256 MINB)).  Measured on the B200 with the
// final kernel (8400 regressions, K = 10 unless noted): 1e5 rows per group 1235 / 1331 / 1364 ms at 2 / 3 / 4 blocks
// (no spills at 128 registers; 144 and 380 bytes of spills at 80 and 64), 3e4 rows 340 / 358 / 358 ms, 1e4 rows
// 111 / - / 111 ms, 1e4 rows with K = 6 87 / 83 / 81 ms: 2 blocks from ~1.5e4 rows on, 4 below.
static int mm_minb(int K, int64_t rows) {
    const int K8 = (K + 1 + 7) / 8;
    int minb = K8 <= 3 ? (rows <= 15000 ? 4 : 2) : 2;
    if (const char* e = getenv("OBBOOT_MM_MINB")) { const int v = atoi(e); if (v >= 2 && v <= 4 && K8 <= 3) minb = v; }
    return minb;
}
int mm_blocks_per_sm(int K, int64_t rows) { return mm_minb(K, rows); }

void mm_qr_launch(const MmArgs& m, int grid, cudaStream_t st) {
    MmKernelArgs a;
    for (int g = 0; g < 2; ++g) { a.X[g] = m.X[g]; a.C[g] = m.C[g]; a.n[g] = m.n[g]; a.n_pad[g] = m.n_pad[g]; }
    a.ldx = m.ldx; a.K = m.K; a.count_bytes = m.count_bytes; a.sims = m.sims; a.slots = m.slots; a.taus = m.taus;
    a.state = m.state; a.state_stride = m.state_stride; a.betas = m.betas; a.info = m.info; a.counter = m.counter;
    a.p_begin = m.p_begin; a.p_end = m.p_end;
    const int K8 = (m.K + 1 + 7) / 8, minb = mm_minb(m.K, m.n[0] > m.n[1] ? m.n[0] : m.n[1]);
#define OB_MM(K8_, MB_) mm_qr_kernel<K8_, MB_><<<grid, MM_THREADS, 0, st>>>(a)
#define OB_MM3(K8_) do { if (minb == 4) OB_MM(K8_, 4); else if (minb == 3) OB_MM(K8_, 3); else OB_MM(K8_, 2); } while (0)
    switch (K8) {
    case 1: OB_MM3(1); break;
    case 2: OB_MM3(2); break;
    case 3: OB_MM3(3); break;
    case 4: OB_MM(4, 2); break;
    case 5: OB_MM(5, 2); break;
    case 6: OB_MM(6, 2); break;
    default: throw StatusError{OB_ERR_UNSUPPORTED, "Machado-Mata: more than 47 design columns"};
    }
#undef OB_MM3
#undef OB_MM
    OB_CUDA(cudaGetLastError());
}

}  // namespace ob

namespace ob {

void mm_streams_launch(const MmArgs& m, long long pass0, int first_slot, uint64_t seed, double* taus, uint32_t* rows_a, uint32_t* rows_b, cudaStream_t st) {
    const long long total = m.slots * m.sims;
    if (total <= 0) return;
    mm_streams_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(m.slots, m.sims, pass0, first_slot, seed, m.C[0], m.C[1], m.n[0], m.n[1],
                                                                     m.n_pad[0], m.n_pad[1], m.count_bytes, taus, rows_a, rows_b);
    OB_CUDA(cudaGetLastError());
}

size_t mm_effects_smem(int sims) {
    int pow2 = 1;
    while (pow2 < sims) pow2 <<= 1;
    return sizeof(double) * 3 * (size_t)pow2 + sizeof(int) * 2 * (size_t)sims;
}

void mm_effects_launch(const MmArgs& m, const uint32_t* rows_a, const uint32_t* rows_b, int nq, const double* d_quantiles,
                       double* stats, int* status, int* nsucc, cudaStream_t st) {
    const size_t smem = mm_effects_smem(m.sims);
    // (per device, so not cached in a process-wide flag: several contexts of one process may sit on different GPUs)
    if (smem > 48 * 1024) OB_CUDA(cudaFuncSetAttribute(mm_effects_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    mm_effects_kernel<<<(unsigned)m.slots, 256, smem, st>>>(m.X[0], m.X[1], m.ldx, m.K, m.sims, m.slots, m.betas, m.info, rows_a, rows_b, nq,
                                                           d_quantiles, stats, status, nsucc);
    OB_CUDA(cudaGetLastError());
}

}  // namespace ob
