// csrc/heckman.cu -- the Heckman two-step replicate on the multiplicity matrix.
//
// Reference, per replicate and group (estimation.rs:114-269 -> heckman.rs:38-108 -> math/probit.rs:25-175):
//   gather the resampled rows; probit of the selection outcome on [1 | selection predictors] over ALL rows of the
//   group by Fisher scoring from 0 (<= 100 steps, ||step|| < 1e-6, -H + 1e-9 I factored by Cholesky, LU fallback);
//   IMR = phi/Phi on the selected rows; OLS of y on [X | IMR] over the selected rows; delta = mean(-IMR (IMR + z'gamma)).
//
// Here nothing is gathered: row i enters replicate b with weight c[i,b] (the multiplicity), so every sum over the
// resampled rows is a sum over the group's rows weighted by the count tile.  Unlike the OLS path the row weights of
// the probit depend on the replicate's own coefficients -- there is no shared Z to contract against, the work is
// exp/erfc-bound CUDA-core work, "one thread per replicate slot, rows streamed":
//   hk_probit_accum    per (row chunk, panel): gradient and expected Hessian of the 128 slots of the panel over the
//                      chunk's rows (counts read coalesced, the selection row broadcast from shared memory)
//   hk_probit_update   per slot: chunk partials summed in chunk order, (-H) solved, gamma += step, convergence
//   hk_terms           after convergence, per (chunk, panel): column sums of Z (all rows); on the selected rows the IMR,
//                      its sums (IMR, IMR^2, IMR y, -IMR (IMR + z'gamma)) and L[i,b] = c IMR for the cross term
//   hk_xterm           X' (c IMR): 16 design columns per pass over L
//   (gram.cu)          X'CX, X'Cy, column sums over the selected rows: the ordinary DMMA contraction on a copy of the
//                      design whose unselected rows are zeroed -- intercept included, so G[0][0] is the number of
//                      selected rows drawn
//   hk_solve           per slot: [X | IMR]'[X | IMR] assembled, Cholesky, coefficients, means, beta*, two/three-fold,
//                      detailed rows over K+1 columns, detailed_selection (builder.rs:464-534)
// Chunk partials are summed in chunk order by one thread: bit-identical run to run and across panel batching.
#include "common.cuh"
#include "internal.h"

#include <algorithm>
#include <cmath>

namespace ob {

constexpr int HK_CHUNK = 2048;     // rows per block
constexpr int HK_STAGE = 256;      // rows staged in shared memory at a time
constexpr int HK_NH = HK_MAX_SEL * (HK_MAX_SEL + 1) / 2;

__device__ __forceinline__ double hk_pdf(double x) { return exp(-0.5 * x * x) / 2.5066282746310002; }   // statrs Normal::pdf
__device__ __forceinline__ double hk_cdf(double x) { return 0.5 * erfc(-x / 1.4142135623730951); }        // statrs Normal::cdf

// ------------------------------------------------------------------------------------------------ probit: accumulate
// partial layout: [chunk][panel][acc][BM] with acc = K1 gradient entries then K1 (K1+1)/2 Hessian entries (lower, row-major)
template <typename CountT>
__global__ void __launch_bounds__(BM) hk_probit_accum_kernel(const double* __restrict__ Z, const uint8_t* __restrict__ sel, long long n,
                                                             long long n_pad, int K1, const CountT* __restrict__ C,
                                                             const double* __restrict__ gamma, const int* __restrict__ active,
                                                             long long slots_pad, double* __restrict__ partial, int nacc) {
    __shared__ double zs[HK_STAGE * HK_MAX_SEL];
    __shared__ uint8_t ss[HK_STAGE];
    const int chunk = blockIdx.x, panel = blockIdx.y, slot = threadIdx.x;
    const long long gs = (long long)panel * BM + slot;
    const bool on = active[gs] != 0;
    double g[HK_MAX_SEL], H[HK_NH], b[HK_MAX_SEL];
#pragma unroll
    for (int j = 0; j < HK_MAX_SEL; ++j) { g[j] = 0.0; b[j] = j < K1 ? gamma[gs * HK_MAX_SEL + j] : 0.0; }
#pragma unroll
    for (int j = 0; j < HK_NH; ++j) H[j] = 0.0;
    const long long r0 = (long long)chunk * HK_CHUNK, r1 = min(r0 + HK_CHUNK, n);
    const CountT* Cp = C + (size_t)panel * n_pad * BM;
    for (long long s0 = r0; s0 < r1; s0 += HK_STAGE) {
        const int rows = (int)min((long long)HK_STAGE, r1 - s0);
        __syncthreads();
        for (int e = threadIdx.x; e < rows * K1; e += BM) zs[(e / K1) * HK_MAX_SEL + e % K1] = Z[s0 * K1 + e];
        for (int e = threadIdx.x; e < rows; e += BM) ss[e] = sel[s0 + e];
        __syncthreads();
        if (!on) continue;
        for (int r = 0; r < rows; ++r) {
            const unsigned c = Cp[(s0 + r) * BM + slot];
            if (c == 0) continue;
            const double* z = zs + r * HK_MAX_SEL;
            double eta = 0.0;
#pragma unroll
            for (int j = 0; j < HK_MAX_SEL; ++j) if (j < K1) eta += z[j] * b[j];        // probit.rs:54
            const double phi = hk_pdf(eta);
            double Phi = hk_cdf(eta);
            Phi = fmin(fmax(Phi, 1e-10), 1.0 - 1e-10);                                  // :71
            const double lam = ss[r] ? phi / Phi : -phi / (1.0 - Phi);                  // :73-77 (y > 0.5)
            const double sw = sqrt((phi * phi) / (Phi * (1.0 - Phi)));                  // :82-83
            const double cw = (double)c * (sw * sw), cl = (double)c * lam;              // :95-98; the row is drawn c times
            int q = 0;
#pragma unroll
            for (int j = 0; j < HK_MAX_SEL; ++j) {
                if (j < K1) {
                    g[j] += cl * z[j];                                                  // :87
#pragma unroll
                    for (int l = 0; l <= j; ++l) H[q + l] += cw * z[j] * z[l];          // :100-116 (sign applied in the update)
                }
                q += j + 1;
            }
        }
    }
    double* out = partial + (((size_t)chunk * gridDim.y + panel) * nacc) * BM + slot;
    int a = 0;
#pragma unroll
    for (int j = 0; j < HK_MAX_SEL; ++j) if (j < K1) out[(size_t)(a++) * BM] = g[j];
    int q = 0;
#pragma unroll
    for (int j = 0; j < HK_MAX_SEL; ++j) {
        if (j < K1) {
#pragma unroll
            for (int l = 0; l <= j; ++l) out[(size_t)(a++) * BM] = H[q + l];
        }
        q += j + 1;
    }
    (void)slots_pad;
}

// small dense solves in registers / local memory (K1 <= HK_MAX_SEL)
__device__ bool hk_chol_solve(double* A, double* x, int n) {      // A row-major n x n SPD (lower used), x = rhs -> solution
    for (int j = 0; j < n; ++j) {
        for (int k = 0; k < j; ++k) {
            const double f = A[j * HK_MAX_SEL + k];
            for (int i = j; i < n; ++i) A[i * HK_MAX_SEL + j] -= A[i * HK_MAX_SEL + k] * f;
        }
        const double d = A[j * HK_MAX_SEL + j];
        if (!(d > 0.0)) return false;
        const double r = sqrt(d);
        A[j * HK_MAX_SEL + j] = r;
        for (int i = j + 1; i < n; ++i) A[i * HK_MAX_SEL + j] /= r;
    }
    for (int i = 0; i < n; ++i) { double s = x[i]; for (int k = 0; k < i; ++k) s -= A[i * HK_MAX_SEL + k] * x[k]; x[i] = s / A[i * HK_MAX_SEL + i]; }
    for (int i = n - 1; i >= 0; --i) { double s = x[i]; for (int k = i + 1; k < n; ++k) s -= A[k * HK_MAX_SEL + i] * x[k]; x[i] = s / A[i * HK_MAX_SEL + i]; }
    return true;
}
__device__ bool hk_lu_solve(double* A, double* x, int n) {        // partial pivoting; false on an exactly zero pivot
    for (int c = 0; c < n; ++c) {
        int piv = c; double best = fabs(A[c * HK_MAX_SEL + c]);
        for (int r = c + 1; r < n; ++r) if (fabs(A[r * HK_MAX_SEL + c]) > best) { best = fabs(A[r * HK_MAX_SEL + c]); piv = r; }
        if (A[piv * HK_MAX_SEL + c] == 0.0) return false;
        if (piv != c) {
            for (int k = 0; k < n; ++k) { const double t = A[c * HK_MAX_SEL + k]; A[c * HK_MAX_SEL + k] = A[piv * HK_MAX_SEL + k]; A[piv * HK_MAX_SEL + k] = t; }
            const double t = x[c]; x[c] = x[piv]; x[piv] = t;
        }
        for (int r = c + 1; r < n; ++r) {
            const double f = A[r * HK_MAX_SEL + c] / A[c * HK_MAX_SEL + c];
            for (int k = c; k < n; ++k) A[r * HK_MAX_SEL + k] -= f * A[c * HK_MAX_SEL + k];
            x[r] -= f * x[c];
        }
    }
    for (int i = n - 1; i >= 0; --i) { double s = x[i]; for (int k = i + 1; k < n; ++k) s -= A[i * HK_MAX_SEL + k] * x[k]; x[i] = s / A[i * HK_MAX_SEL + i]; }
    return true;
}

// ------------------------------------------------------------------------------------------------ probit: update
__global__ void __launch_bounds__(BM) hk_probit_update_kernel(const double* __restrict__ partial, int nchunks, int panels, int nacc, int K1,
                                                              long long slots, double tol, int last_iter, double* __restrict__ gamma,
                                                              int* __restrict__ active, int* __restrict__ pstatus, int* __restrict__ n_active) {
    const int panel = blockIdx.x, slot = threadIdx.x;
    const long long gs = (long long)panel * BM + slot;
    if (gs >= slots || !active[gs]) return;
    double acc[HK_MAX_SEL + HK_NH];
    for (int a = 0; a < nacc; ++a) acc[a] = 0.0;
    for (int ch = 0; ch < nchunks; ++ch) {                   // chunk order: the summation order of every slot is fixed
        const double* p = partial + (((size_t)ch * panels + panel) * nacc) * BM + slot;
        for (int a = 0; a < nacc; ++a) acc[a] += p[(size_t)a * BM];
    }
    double M[HK_MAX_SEL * HK_MAX_SEL], step[HK_MAX_SEL];
    int q = K1;
    for (int j = 0; j < K1; ++j)
        for (int l = 0; l <= j; ++l) { const double v = acc[q++]; M[j * HK_MAX_SEL + l] = v; M[l * HK_MAX_SEL + j] = v; }   // = -H without the ridge
    for (int j = 0; j < K1; ++j) { M[j * HK_MAX_SEL + j] += 1e-9; step[j] = acc[j]; }                                          // probit.rs:124-126, :132
    double Mc[HK_MAX_SEL * HK_MAX_SEL];
    for (int e = 0; e < HK_MAX_SEL * HK_MAX_SEL; ++e) Mc[e] = M[e];
    bool ok = hk_chol_solve(Mc, step, K1);                                                                                      // :133-134
    if (!ok) {                                                // :135-147: LU of H, step = -(H^-1 g) = (-H)^-1 g
        for (int e = 0; e < HK_MAX_SEL * HK_MAX_SEL; ++e) Mc[e] = M[e];
        for (int j = 0; j < K1; ++j) step[j] = acc[j];
        ok = hk_lu_solve(Mc, step, K1);
    }
    if (!ok) { pstatus[gs] = OB_ERR_NALGEBRA; active[gs] = 0; return; }
    double nrm = 0.0;
    for (int j = 0; j < K1; ++j) { gamma[gs * HK_MAX_SEL + j] += step[j]; nrm += step[j] * step[j]; }                           // :150
    if (sqrt(nrm) < tol || last_iter) {                                                                                          // :152-155 / max_iter reached
        // probit.rs:166-173: vcov = -(H^-1) of the LAST Hessian must exist (fails only on an exactly singular H)
        for (int e = 0; e < HK_MAX_SEL * HK_MAX_SEL; ++e) Mc[e] = M[e];
        for (int j = 0; j < K1; ++j) step[j] = 0.0;
        if (!hk_lu_solve(Mc, step, K1)) pstatus[gs] = OB_ERR_NALGEBRA;
        active[gs] = 0;
    } else {
        atomicAdd(n_active, 1);
    }
}

// ------------------------------------------------------------------------------------------------ IMR terms
// acc: [0,K1) sum c z_j (all rows); K1+0 sum c IMR; +1 sum c IMR^2; +2 sum c IMR y; +3 sum c (-IMR (IMR + z'gamma)); +4 sum c y (all rows)
template <typename CountT>
__global__ void __launch_bounds__(BM) hk_terms_kernel(const double* __restrict__ Z, const uint8_t* __restrict__ sel, const double* __restrict__ X,
                                                      int ldx, int ycol, long long n, long long n_pad, int K1, const CountT* __restrict__ C,
                                                      const double* __restrict__ gamma, double* __restrict__ L, double* __restrict__ partial, int nacc) {
    __shared__ double zs[HK_STAGE * HK_MAX_SEL];
    __shared__ double ys[HK_STAGE];
    __shared__ uint8_t ss[HK_STAGE];
    const int chunk = blockIdx.x, panel = blockIdx.y, slot = threadIdx.x;
    const long long gs = (long long)panel * BM + slot;
    double zsum[HK_MAX_SEL], b[HK_MAX_SEL], t0 = 0.0, t1 = 0.0, t2 = 0.0, t3 = 0.0, t4 = 0.0;
#pragma unroll
    for (int j = 0; j < HK_MAX_SEL; ++j) { zsum[j] = 0.0; b[j] = j < K1 ? gamma[gs * HK_MAX_SEL + j] : 0.0; }
    const long long r0 = (long long)chunk * HK_CHUNK, r1 = min(r0 + HK_CHUNK, n);
    const CountT* Cp = C + (size_t)panel * n_pad * BM;
    double* Lp = L + (size_t)panel * n_pad * BM;
    for (long long s0 = r0; s0 < r1; s0 += HK_STAGE) {
        const int rows = (int)min((long long)HK_STAGE, r1 - s0);
        __syncthreads();
        for (int e = threadIdx.x; e < rows * K1; e += BM) zs[(e / K1) * HK_MAX_SEL + e % K1] = Z[s0 * K1 + e];
        for (int e = threadIdx.x; e < rows; e += BM) { ss[e] = sel[s0 + e]; ys[e] = X[(s0 + e) * ldx + ycol]; }
        __syncthreads();
        for (int r = 0; r < rows; ++r) {
            const double c = (double)Cp[(s0 + r) * BM + slot];
            double l = 0.0;
            if (c != 0.0) {
                const double* z = zs + r * HK_MAX_SEL;
                t4 += c * ys[r];
#pragma unroll
                for (int j = 0; j < HK_MAX_SEL; ++j) if (j < K1) zsum[j] += c * z[j];
                if (ss[r]) {
                    double eta = 0.0;
#pragma unroll
                    for (int j = 0; j < HK_MAX_SEL; ++j) if (j < K1) eta += z[j] * b[j];   // heckman.rs:54
                    const double phi = hk_pdf(eta), Phi = hk_cdf(eta);
                    const double imr = Phi < 1e-10 ? 0.0 : phi / Phi;                     // :58-66
                    l = c * imr;
                    t0 += l; t1 += l * imr; t2 += l * ys[r]; t3 += c * (-imr * (imr + eta));   // :92-97
                }
            }
            Lp[(s0 + r) * BM + slot] = l;
        }
    }
    double* out = partial + (((size_t)chunk * gridDim.y + panel) * nacc) * BM + slot;
    int a = 0;
#pragma unroll
    for (int j = 0; j < HK_MAX_SEL; ++j) if (j < K1) out[(size_t)(a++) * BM] = zsum[j];
    out[(size_t)(a++) * BM] = t0; out[(size_t)(a++) * BM] = t1; out[(size_t)(a++) * BM] = t2; out[(size_t)(a++) * BM] = t3;
    out[(size_t)(a++) * BM] = t4;
}

// X' (c IMR) for design columns [j0, j0 + 16): partial [chunk][panel][K][BM]
__global__ void __launch_bounds__(BM) hk_xterm_kernel(const double* __restrict__ X, int ldx, int K, long long n, long long n_pad,
                                                      const double* __restrict__ L, double* __restrict__ partial) {
    __shared__ double xs[HK_STAGE * 16];
    const int chunk = blockIdx.x, panel = blockIdx.y, j0 = blockIdx.z * 16, slot = threadIdx.x;
    const int nj = min(16, K - j0);
    double acc[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) acc[j] = 0.0;
    const long long r0 = (long long)chunk * HK_CHUNK, r1 = min(r0 + HK_CHUNK, n);
    const double* Lp = L + (size_t)panel * n_pad * BM;
    for (long long s0 = r0; s0 < r1; s0 += HK_STAGE) {
        const int rows = (int)min((long long)HK_STAGE, r1 - s0);
        __syncthreads();
        for (int e = threadIdx.x; e < rows * 16; e += BM) { const int r = e / 16, j = e % 16; xs[e] = j < nj ? X[(s0 + r) * ldx + j0 + j] : 0.0; }
        __syncthreads();
        for (int r = 0; r < rows; ++r) {
            const double l = Lp[(s0 + r) * BM + slot];
            if (l == 0.0) continue;
#pragma unroll
            for (int j = 0; j < 16; ++j) acc[j] += l * xs[r * 16 + j];
        }
    }
    double* out = partial + (((size_t)chunk * gridDim.y + panel) * K + j0) * BM + slot;
#pragma unroll
    for (int j = 0; j < 16; ++j) if (j < nj) out[(size_t)j * BM] = acc[j];
}

// out [acc][slots_pad] = sum over chunks, in chunk order, of partial [chunk][panel][acc][BM]
__global__ void __launch_bounds__(BM) hk_reduce_kernel(const double* __restrict__ partial, int nchunks, int panels, int nacc,
                                                       double* __restrict__ out, long long slots_pad) {
    const int panel = blockIdx.x, a = blockIdx.y, slot = threadIdx.x;
    double s = 0.0;
    for (int ch = 0; ch < nchunks; ++ch) s += partial[(((size_t)ch * panels + panel) * nacc + a) * BM + slot];
    out[(size_t)a * slots_pad + (size_t)panel * BM + slot] = s;
}

// zero the rows of the outcome design that are not selected (intercept included): the Gram contraction over all rows
// then equals the one over the selected rows, and its intercept entries count the selected rows drawn
__global__ void __launch_bounds__(256) hk_mask_kernel(const double* __restrict__ X, const uint8_t* __restrict__ sel, double* __restrict__ Xm,
                                                      long long rows, int ldx) {
    const long long total = rows * ldx;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x)
        Xm[e] = sel[e / ldx] ? X[e] : 0.0;
}

// frame-order selection columns -> packed group order: Z[r][0] = 1, Z[r][1+j] = pred_j[src[r]]; sel[r] = (outcome[src[r]] == 1)
__global__ void __launch_bounds__(256) hk_gather_kernel(const uint32_t* __restrict__ src, long long n, int K1, const double* const* __restrict__ pred,
                                                        const double* __restrict__ outcome, double* __restrict__ Z, uint8_t* __restrict__ sel,
                                                        int* __restrict__ flags) {
    for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += (long long)gridDim.x * blockDim.x) {
        const uint32_t f = src[r];
        Z[r * K1] = 1.0;                                              // estimation.rs:190-195: intercept first
        for (int j = 1; j < K1; ++j) Z[r * K1 + j] = pred[j - 1][f];
        const double s = outcome[f];
        sel[r] = s == 1.0 ? 1 : 0;                                    // estimation.rs:213 equal(1)
        if (s != s) atomicOr(&flags[0], 1);                           // "Selection outcome contains nulls" (estimation.rs:179-185)
    }
}

// ------------------------------------------------------------------------------------------------ solve + epilogue
struct HkSolveParams {
    const double* gram; long long slots_pad; int Pld; long long slots;
    int K, K1, ref_kind;
    const double* terms[2];     // [nacc_t][slots_pad] per group: zsum[K1], imr, imr^2, imr y, delta sum, sum c y
    const double* xterm[2];     // [K][slots_pad] per group
    const double* gamma[2];     // [slots_pad][HK_MAX_SEL]
    const int* pstatus[2];      // probit status per slot
    double na, nb;              // rows of each group (all rows: every replicate draws exactly that many)
    int S;
    double* stats; int* status; double* beta_a; double* beta_b; double* point_extra;
};

__device__ bool hk_chol_factor_sm(double* G, int N, int ld, int* fail) {
    for (int j = 0; j < N; ++j) {
        for (int i = j + threadIdx.x; i < N; i += blockDim.x) {
            double s = G[i * ld + j];
            for (int k = 0; k < j; ++k) s -= G[i * ld + k] * G[j * ld + k];
            G[i * ld + j] = s;
        }
        __syncthreads();
        const double d = G[j * ld + j];
        if (!(d > 0.0)) { if (threadIdx.x == 0) *fail = 1; __syncthreads(); return false; }
        const double r = sqrt(d);
        __syncthreads();
        for (int i = j + threadIdx.x; i < N; i += blockDim.x) G[i * ld + j] = (i == j) ? r : G[i * ld + j] / r;
        __syncthreads();
    }
    return true;
}

__global__ void __launch_bounds__(128) hk_solve_kernel(const HkSolveParams p) {
    extern __shared__ __align__(16) double sm[];
    const int K = p.K, Ka = K + 1, K1 = p.K1, ld = Ka | 1;
    const long long slot = blockIdx.x;
    const int tid = threadIdx.x;
    double* M = sm;                       // [Ka][ld]
    double* BA = M + Ka * ld;             // [Ka] coefficients of group A (IMR last)
    double* BB = BA + Ka;
    double* xa = BB + Ka; double* xb = xa + Ka; double* bs = xb + Ka;
    __shared__ int fail;
    __shared__ double dl[2];              // delta per group
    if (tid == 0) fail = 0;
    int status = OB_OK;
    for (int g = 0; g < 2; ++g) {
        const double* G = p.gram + ((size_t)g * p.slots_pad + slot) * p.Pld;
        const double* T = p.terms[g];
        double* B = g ? BB : BA; double* xm = g ? xb : xa;
        __syncthreads();
        if (status != OB_OK) break;
        if (p.pstatus[g][slot] != OB_OK) { status = p.pstatus[g][slot]; break; }
        const double m = G[0];                                    // selected rows drawn (intercept x intercept over the masked design)
        if (!(m > 0.0)) { status = OB_ERR_INVALID_GROUP; break; } // estimation.rs:236-240 "No observed outcomes in group"
        if (m <= (double)Ka) { status = OB_ERR_INSUFFICIENT_DATA; break; }   // ols.rs:98-105 on the augmented design
        for (int i = 0; i < K; ++i) {
            const long long base_idx = (long long)i * (K + 1) - (long long)i * (i - 1) / 2 - i;
            for (int l = i + tid; l < K; l += blockDim.x) { const double v = G[base_idx + l]; M[i * ld + l] = v; M[l * ld + i] = v; }
        }
        for (int j = tid; j < K; j += blockDim.x) {
            const double v = p.xterm[g][(size_t)j * p.slots_pad + slot];          // sum c IMR x_j
            M[K * ld + j] = v; M[j * ld + K] = v;
            B[j] = G[(long long)j * (K + 1) - (long long)j * (j - 1) / 2 + (K - j)];   // X'Cy over the selected rows
            xm[j] = G[j] / m;                                                     // estimation.rs:143-144
        }
        if (tid == 0) {
            M[K * ld + K] = T[(size_t)(K1 + 1) * p.slots_pad + slot];              // sum c IMR^2
            B[K] = T[(size_t)(K1 + 2) * p.slots_pad + slot];                       // sum c IMR y
            xm[K] = T[(size_t)(K1 + 0) * p.slots_pad + slot] / m;                  // estimation.rs:146-151
            dl[g] = T[(size_t)(K1 + 3) * p.slots_pad + slot] / m;                  // heckman.rs:92-97
        }
        __syncthreads();
        if (!hk_chol_factor_sm(M, Ka, ld, &fail)) { status = OB_ERR_NALGEBRA; break; }
        if (tid < 32) {     // L L' x = b by warp 0
            const int lane = tid;
            for (int i = 0; i < Ka; ++i) {
                double s = 0.0;
                for (int k = lane; k < i; k += 32) s += M[i * ld + k] * B[k];
                for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
                if (lane == 0) B[i] = (B[i] - s) / M[i * ld + i];
                __syncwarp();
            }
            for (int i = Ka - 1; i >= 0; --i) {
                double s = 0.0;
                for (int k = i + 1 + lane; k < Ka; k += 32) s += M[k * ld + i] * B[k];
                for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
                if (lane == 0) B[i] = (B[i] - s) / M[i * ld + i];
                __syncwarp();
            }
        }
        __syncthreads();
    }
    __syncthreads();
    double* out = p.stats + (size_t)slot * p.S;
    if (tid == 0) {
        const int D = Ka;
        if (status == OB_OK) {
            const bool ref_a = p.ref_kind == OB_REF_GROUP_A;
            if (ref_a) for (int j = 0; j < Ka; ++j) bs[j] = BA[j];
            else if (p.ref_kind == OB_REF_GROUP_B) for (int j = 0; j < Ka; ++j) bs[j] = BB[j];
            else {   // Weighted | Cotton: builder.rs:591-620 with df_a.height(), df_b.height()
                const double wA = p.na / (p.na + p.nb), wB = 1.0 - wA;
                for (int j = 0; j < Ka; ++j) bs[j] = BA[j] * wA + BB[j] * wB;
            }
            double en = 0.0, co = 0.0, in = 0.0, ex = 0.0, fa = 0.0, fb = 0.0;
            for (int j = 0; j < Ka; ++j) {
                const double dx = xa[j] - xb[j], db = BA[j] - BB[j];
                en += dx * BB[j]; co += xb[j] * db; in += dx * db;
                ex += dx * bs[j]; fa += xa[j] * BA[j]; fb += xb[j] * BB[j];
                out[5 + j] = dx * bs[j];
                out[5 + D + j] = xa[j] * (BA[j] - bs[j]) + xb[j] * (bs[j] - BB[j]);
            }
            out[0] = ex; out[1] = (fa - fb) - ex; out[2] = en; out[3] = co; out[4] = in;
            // detailed_selection (builder.rs:477-534): theta_ref delta_ref gamma_ref[i] (zbar_a[i] - zbar_b[i])
            const int gr = ref_a ? 0 : 1;
            const double theta = ref_a ? BA[K] : BB[K];
            for (int j = 0; j < K1; ++j) {
                const double za = p.terms[0][(size_t)j * p.slots_pad + slot] / p.na, zb = p.terms[1][(size_t)j * p.slots_pad + slot] / p.nb;
                out[5 + 2 * D + j] = theta * dl[gr] * p.gamma[gr][slot * HK_MAX_SEL + j] * (za - zb);
            }
            if (p.beta_a) for (int j = 0; j < Ka; ++j) p.beta_a[(size_t)slot * Ka + j] = BA[j];
            if (p.beta_b) for (int j = 0; j < Ka; ++j) p.beta_b[(size_t)slot * Ka + j] = BB[j];
            if (p.point_extra && slot == 0) {
                double* pe = p.point_extra;   // [xa Ka | xb Ka | beta* Ka | gamma_a K1 | gamma_b K1 | total_gap]
                for (int j = 0; j < Ka; ++j) { pe[j] = xa[j]; pe[Ka + j] = xb[j]; pe[2 * Ka + j] = bs[j]; }
                for (int j = 0; j < K1; ++j) { pe[3 * Ka + j] = p.gamma[0][slot * HK_MAX_SEL + j]; pe[3 * Ka + K1 + j] = p.gamma[1][slot * HK_MAX_SEL + j]; }
                // builder.rs:676-684: mean outcome over ALL rows of each group
                pe[3 * Ka + 2 * K1] = p.terms[0][(size_t)(K1 + 4) * p.slots_pad + slot] / p.na - p.terms[1][(size_t)(K1 + 4) * p.slots_pad + slot] / p.nb;
            }
        } else {
            const double nan = __longlong_as_double(0x7ff8000000000000LL);
            for (int j = 0; j < p.S; ++j) out[j] = nan;
            if (p.beta_a) for (int j = 0; j < Ka; ++j) p.beta_a[(size_t)slot * Ka + j] = nan;
            if (p.beta_b) for (int j = 0; j < Ka; ++j) p.beta_b[(size_t)slot * Ka + j] = nan;
        }
        p.status[slot] = status;
    }
}

// ------------------------------------------------------------------------------------------------ host launchers
int hk_num_chunks(int64_t n) { return (int)std::max<int64_t>(1, (n + HK_CHUNK - 1) / HK_CHUNK); }

void hk_gather_launch(const uint32_t* src, int64_t n, int K1, const double* const* d_pred, const double* d_outcome, double* Z, uint8_t* sel,
                      int* d_flags, cudaStream_t st) {
    if (n == 0) return;
    hk_gather_kernel<<<(unsigned)std::min<int64_t>((n + 255) / 256, 148 * 16), 256, 0, st>>>(src, n, K1, d_pred, d_outcome, Z, sel, d_flags);
    OB_CUDA(cudaGetLastError());
}

void hk_mask_launch(const double* X, const uint8_t* sel, double* Xm, int64_t rows, int ldx, cudaStream_t st) {
    if (rows == 0) return;
    const long long total = rows * (long long)ldx;
    hk_mask_kernel<<<(unsigned)std::min<long long>((total + 255) / 256, 148 * 32), 256, 0, st>>>(X, sel, Xm, rows, ldx);
    OB_CUDA(cudaGetLastError());
}

void hk_probit_accum_launch(const HkGroup& g, int K1, const void* C, int count_bytes, int panels, const double* gamma, const int* active,
                            int64_t slots_pad, double* partial, cudaStream_t st) {
    const int nacc = K1 + K1 * (K1 + 1) / 2;
    dim3 grid(hk_num_chunks(g.n), panels);
    if (count_bytes == 1) hk_probit_accum_kernel<uint8_t><<<grid, BM, 0, st>>>(g.Z, g.sel, g.n, g.n_pad, K1, (const uint8_t*)C, gamma, active, slots_pad, partial, nacc);
    else hk_probit_accum_kernel<uint16_t><<<grid, BM, 0, st>>>(g.Z, g.sel, g.n, g.n_pad, K1, (const uint16_t*)C, gamma, active, slots_pad, partial, nacc);
    OB_CUDA(cudaGetLastError());
}

void hk_probit_update_launch(const double* partial, int nchunks, int panels, int K1, int64_t slots, double tol, int last_iter, double* gamma,
                             int* active, int* pstatus, int* n_active, cudaStream_t st) {
    const int nacc = K1 + K1 * (K1 + 1) / 2;
    hk_probit_update_kernel<<<panels, BM, 0, st>>>(partial, nchunks, panels, nacc, K1, slots, tol, last_iter, gamma, active, pstatus, n_active);
    OB_CUDA(cudaGetLastError());
}

void hk_terms_launch(const HkGroup& g, const double* X, int ldx, int ycol, int K1, const void* C, int count_bytes, int panels, const double* gamma,
                     double* L, double* partial, cudaStream_t st) {
    const int nacc = K1 + 5;
    dim3 grid(hk_num_chunks(g.n), panels);
    if (count_bytes == 1) hk_terms_kernel<uint8_t><<<grid, BM, 0, st>>>(g.Z, g.sel, X, ldx, ycol, g.n, g.n_pad, K1, (const uint8_t*)C, gamma, L, partial, nacc);
    else hk_terms_kernel<uint16_t><<<grid, BM, 0, st>>>(g.Z, g.sel, X, ldx, ycol, g.n, g.n_pad, K1, (const uint16_t*)C, gamma, L, partial, nacc);
    OB_CUDA(cudaGetLastError());
}

void hk_xterm_launch(const HkGroup& g, const double* X, int ldx, int K, int panels, const double* L, double* partial, cudaStream_t st) {
    dim3 grid(hk_num_chunks(g.n), panels, (K + 15) / 16);
    hk_xterm_kernel<<<grid, BM, 0, st>>>(X, ldx, K, g.n, g.n_pad, L, partial);
    OB_CUDA(cudaGetLastError());
}

void hk_reduce_launch(const double* partial, int nchunks, int panels, int nacc, double* out, int64_t slots_pad, cudaStream_t st) {
    dim3 grid(panels, nacc);
    hk_reduce_kernel<<<grid, BM, 0, st>>>(partial, nchunks, panels, nacc, out, slots_pad);
    OB_CUDA(cudaGetLastError());
}

void hk_solve_launch(const HkSolveArgs& a, cudaStream_t st) {
    HkSolveParams p;
    p.gram = a.gram; p.slots_pad = a.slots_pad; p.Pld = a.Pld; p.slots = a.slots; p.K = a.K; p.K1 = a.K1; p.ref_kind = a.ref_kind;
    for (int g = 0; g < 2; ++g) { p.terms[g] = a.terms[g]; p.xterm[g] = a.xterm[g]; p.gamma[g] = a.gamma[g]; p.pstatus[g] = a.pstatus[g]; }
    p.na = a.na; p.nb = a.nb; p.S = a.S; p.stats = a.stats; p.status = a.status; p.beta_a = a.beta_a; p.beta_b = a.beta_b; p.point_extra = a.point_extra;
    const int Ka = a.K + 1;
    const size_t smem = sizeof(double) * ((size_t)Ka * (Ka | 1) + 5 * (size_t)Ka);
    if (smem > 227 * 1024) throw StatusError{OB_ERR_UNSUPPORTED, "design too wide for the Heckman solve kernel (K + 1 <= 160)"};
    OB_CUDA(cudaFuncSetAttribute(hk_solve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    hk_solve_kernel<<<(unsigned)a.slots, 128, smem, st>>>(p);
    OB_CUDA(cudaGetLastError());
}

}  // namespace ob
