// csrc/rif.cu -- RIF outcome pre-step of decompose_quantile (builder.rs:721-737 -> math/rif.rs:14-88),
// on the packed design's outcome column, once per group before the bootstrap (the RIF vector is
// fixed across replicates; only rows are resampled).
//
//   q_tau   type-7 sample quantile (rif.rs:23-35)           -> 64-bit MSB radix SELECT of the four order
//   IQR     sorted[ceil(.75n)-1] - sorted[ceil(.25n)-1]        statistics needed (no full sort), 8 passes of
//                                                              8 bits over an orderable-key copy of y
//   sd      two-pass sample SD (rif.rs:39-41)                -> fixed-order partials, two per leaf segment of the group's
//                                                              GLOBAL row segmentation (internal.h): the sums are the same
//                                                              whether one GPU holds the group or its rows are sharded
//   h       0.9 min(sd, IQR/1.34) n^-0.2 with fallbacks (rif.rs:51-59)
//   f(q)    Gaussian KDE at one point (rif.rs:65-72), floor 1e-8 (:75)
//   RIF_i   q + (tau - 1[y_i <= q]) / f (rif.rs:79-85), written back in place
// HBM-bound: ~12 passes over n doubles.  No host synchronisation: scalars stay in a device state block.
//
// Row-sharded designs (mode N): every rank runs the same kernels over its rows; the leaf partials (disjoint support
// across ranks, so the all-reduce adds zeros: exact) and the radix histograms (integers) are all-reduced over the
// communicator, 11 small collectives per transform.  Quantile, bandwidth, density and hence every RIF value are
// bit-identical to the unsharded transform.
#include "common.cuh"
#include "internal.h"

namespace ob {

constexpr int RIF_BLOCKS = 592;   // 4 x 148 SMs (histogram / apply passes)
constexpr int RIF_THREADS = 256;
constexpr int RIF_SUB = 2;                        // reduction blocks per leaf
constexpr int RIF_PARTS = MAX_SEGS * RIF_SUB;     // fixed-order partial sums of a group

struct RifState {
    unsigned long long prefix[4];
    long long rank[4];
    unsigned int hist[4][256];
    double partial[RIF_PARTS];
    double mean, sd, q, dens, bw;
    int active;   // 0 when n < 2: series returned unchanged (rif.rs:18-20)
};

__device__ __forceinline__ unsigned long long to_key(double v) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(v);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double from_key(unsigned long long k) {
    const unsigned long long b = (k >> 63) ? (k & 0x7FFFFFFFFFFFFFFFull) : ~k;
    return __longlong_as_double((long long)b);
}

__device__ __forceinline__ double block_reduce_sum(double v, double* red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0.0;
    if (threadIdx.x == 0) for (int i = 0; i < RIF_THREADS / 32; ++i) t += red[i];
    return t;  // valid on thread 0
}

// contiguous row range of each block (histogram pass: integer counts, any partition will do)
__device__ __forceinline__ void block_range(long long n, long long& lo, long long& hi) {
    const long long per = (n + gridDim.x - 1) / gridDim.x;
    lo = (long long)blockIdx.x * per; hi = min(lo + per, n);
}

// where a group's rows sit: the global segmentation and the part held here
struct RifRows { long long n_global, row_begin, n_local; int seg_rows; };

// Reduction blocks: block b sums half (b % RIF_SUB) of leaf b / RIF_SUB of the group's GLOBAL segmentation -- the local
// rows [lo, hi) of it, empty when the leaf lives on another rank.  The partition depends on the group size only.
__device__ __forceinline__ void leaf_range(const RifRows& r, long long& lo, long long& hi) {
    const long long leaf = blockIdx.x / RIF_SUB, sub = blockIdx.x % RIF_SUB;
    const long long g0 = leaf * r.seg_rows, g1 = min(g0 + r.seg_rows, r.n_global);
    if (g0 >= g1) { lo = hi = 0; return; }
    const long long half = ((g1 - g0) / RIF_SUB + 31) / 32 * 32;
    const long long a = sub == 0 ? g0 : min(g0 + half, g1), b = sub == 0 ? min(g0 + half, g1) : g1;
    lo = min(max(a - r.row_begin, 0ll), r.n_local); hi = min(max(b - r.row_begin, 0ll), r.n_local);
}

// y_src / ystride: where the RAW outcome lives (the outcome column of X on the first transform, the saved copy after)
__global__ void __launch_bounds__(RIF_THREADS) rif_extract_kernel(const double* __restrict__ y_src, long long ystride, RifRows rows,
                                                                  unsigned long long* __restrict__ keys, RifState* s) {
    __shared__ double red[RIF_THREADS / 32];
    long long lo, hi; leaf_range(rows, lo, hi);
    double sum = 0.0;
    for (long long i = lo + threadIdx.x; i < hi; i += RIF_THREADS) {
        const double y = y_src[i * ystride];
        keys[i] = to_key(y);
        sum += y;
    }
    const double t = block_reduce_sum(sum, red);
    if (threadIdx.x == 0) s->partial[blockIdx.x] = t;
}

__global__ void rif_init_kernel(RifState* s, long long n, double tau) {
    // runs after extract: mean from the ordered partials; select targets (rif.rs:25-28, :43-47)
    double sum = 0.0;
    for (int i = 0; i < RIF_PARTS; ++i) sum += s->partial[i];
    const double nf = (double)n;
    s->mean = sum / nf;
    const double h = (nf - 1.0) * tau;
    long long i75 = (long long)ceil(0.75 * nf); i75 = i75 == 0 ? 0 : i75 - 1;
    long long i25 = (long long)ceil(0.25 * nf); i25 = i25 == 0 ? 0 : i25 - 1;
    s->rank[0] = (long long)floor(h); s->rank[1] = (long long)ceil(h);
    s->rank[2] = min(i25, n - 1); s->rank[3] = min(i75, n - 1);
    for (int t = 0; t < 4; ++t) { s->prefix[t] = 0; s->rank[t] = max(0ll, min(s->rank[t], n - 1)); }
    for (int t = 0; t < 4; ++t) for (int b = 0; b < 256; ++b) s->hist[t][b] = 0;
}

__global__ void __launch_bounds__(RIF_THREADS) rif_ss_kernel(const unsigned long long* __restrict__ keys, RifRows rows, RifState* s) {
    __shared__ double red[RIF_THREADS / 32];
    long long lo, hi; leaf_range(rows, lo, hi);
    const double mean = s->mean;
    double ss = 0.0;
    for (long long i = lo + threadIdx.x; i < hi; i += RIF_THREADS) { const double d = from_key(keys[i]) - mean; ss += d * d; }
    const double t = block_reduce_sum(ss, red);
    if (threadIdx.x == 0) s->partial[blockIdx.x] = t;
}

__global__ void rif_sd_kernel(RifState* s, long long n) {
    double ss = 0.0;
    for (int i = 0; i < RIF_PARTS; ++i) ss += s->partial[i];
    s->sd = sqrt(ss / ((double)n - 1.0));
}

// pass = 0 (most significant byte) .. 7
__global__ void __launch_bounds__(RIF_THREADS) rif_hist_kernel(const unsigned long long* __restrict__ keys, long long n,
                                                               RifState* s, int pass) {
    __shared__ unsigned int h[4][256];
    for (int i = threadIdx.x; i < 1024; i += RIF_THREADS) (&h[0][0])[i] = 0;
    __syncthreads();
    const int shift = 56 - 8 * pass;
    unsigned long long pf[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) pf[t] = s->prefix[t];
    long long lo, hi; block_range(n, lo, hi);
    for (long long i = lo + threadIdx.x; i < hi; i += RIF_THREADS) {
        const unsigned long long k = keys[i];
        const unsigned d = (unsigned)(k >> shift) & 0xFFu;
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const bool match = pass == 0 ? true : ((k ^ pf[t]) >> (shift + 8)) == 0;
            if (match) atomicAdd(&h[t][d], 1u);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 1024; i += RIF_THREADS) {
        const unsigned v = (&h[0][0])[i];
        if (v) atomicAdd(&(&s->hist[0][0])[i], v);
    }
}

__global__ void rif_pick_kernel(RifState* s, int pass) {
    const int t = threadIdx.x;
    if (t < 4) {
        const int shift = 56 - 8 * pass;
        long long r = s->rank[t];
        int b = 0;
        for (; b < 255; ++b) {
            const long long c = s->hist[t][b];
            if (r < c) break;
            r -= c;
        }
        s->rank[t] = r;
        s->prefix[t] |= (unsigned long long)b << shift;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) (&s->hist[0][0])[i] = 0;
}

__global__ void rif_params_kernel(RifState* s, long long n, double tau) {
    const double nf = (double)n;
    const double h = (nf - 1.0) * tau;
    const double hf = floor(h), hc = ceil(h), frac = h - hf;
    const double y0 = from_key(s->prefix[0]), y1 = from_key(s->prefix[1]);
    s->q = (hf == hc) ? y0 : y0 + frac * (y1 - y0);                      // rif.rs:29-35
    const double iqr = from_key(s->prefix[3]) - from_key(s->prefix[2]);   // rif.rs:49
    double spread = (iqr > 1e-8) ? fmin(s->sd, iqr / 1.34) : s->sd;       // rif.rs:51-55
    if (spread < 1e-8) spread = 1.0;                                      // rif.rs:57
    s->bw = 0.9 * spread * pow(nf, -0.2);                                 // rif.rs:59
}

__global__ void __launch_bounds__(RIF_THREADS) rif_density_kernel(const unsigned long long* __restrict__ keys, RifRows rows, RifState* s) {
    __shared__ double red[RIF_THREADS / 32];
    long long lo, hi; leaf_range(rows, lo, hi);
    const double q = s->q, bw = s->bw;
    const double c = 1.0 / sqrt(2.0 * 3.14159265358979323846);
    double acc = 0.0;
    for (long long i = lo + threadIdx.x; i < hi; i += RIF_THREADS) {
        const double u = (q - from_key(keys[i])) / bw;
        acc += c * exp(-0.5 * (u * u));                                   // rif.rs:65-71
    }
    const double t = block_reduce_sum(acc, red);
    if (threadIdx.x == 0) s->partial[blockIdx.x] = t;
}

__global__ void rif_dens_final_kernel(RifState* s, long long n) {
    double d = 0.0;
    for (int i = 0; i < RIF_PARTS; ++i) d += s->partial[i];
    d /= ((double)n * s->bw);
    s->dens = d < 1e-8 ? 1e-8 : d;                                        // rif.rs:75
}

__global__ void __launch_bounds__(RIF_THREADS) rif_apply_kernel(double* __restrict__ X, long long n, int K, int ldx,
                                                                const double* __restrict__ y_src, long long ystride,
                                                                const RifState* s, double tau) {
    const double q = s->q, dens = s->dens;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const double y = y_src[i * ystride];
        X[i * ldx + K] = q + (tau - (y <= q ? 1.0 : 0.0)) / dens;         // rif.rs:79-85
    }
}

size_t rif_scratch_bytes(int64_t n) { return sizeof(unsigned long long) * (size_t)std::max<int64_t>(n, 1) + sizeof(RifState) + 256; }

void rif_transform(const GroupData& g, int ycol, int ldx, double tau, void* d_scratch, size_t scratch_bytes, cudaStream_t st, Comm* comm) {
    const long long n = g.n, n_glob = g.shard.n_global;
    if ((double)n_glob < 2.0) return;  // rif.rs:18-20
    if (scratch_bytes < rif_scratch_bytes(n)) throw StatusError{OB_ERR_INVALID_ARG, "rif scratch too small"};
    unsigned long long* keys = static_cast<unsigned long long*>(d_scratch);
    RifState* s = reinterpret_cast<RifState*>(reinterpret_cast<char*>(d_scratch) + ((sizeof(unsigned long long) * (size_t)std::max<long long>(n, 1) + 255) / 256) * 256);
    const double* y_src = g.y_raw ? g.y_raw : g.X + ycol;
    const long long ystride = g.y_raw ? 1 : ldx;
    const RifRows rows{n_glob, g.shard.row_begin, n, g.shard.seg_rows};
    // leaf partials: every rank writes all RIF_PARTS entries (zeros for leaves it does not hold), so the all-reduce is exact
    auto share_partials = [&] { if (comm) comm->allreduce(s->partial, RIF_PARTS, CommDType::F64, CommOp::SUM, st); };
    rif_extract_kernel<<<RIF_PARTS, RIF_THREADS, 0, st>>>(y_src, ystride, rows, keys, s);
    share_partials();
    rif_init_kernel<<<1, 1, 0, st>>>(s, n_glob, tau);
    rif_ss_kernel<<<RIF_PARTS, RIF_THREADS, 0, st>>>(keys, rows, s);
    share_partials();
    rif_sd_kernel<<<1, 1, 0, st>>>(s, n_glob);
    for (int pass = 0; pass < 8; ++pass) {
        rif_hist_kernel<<<RIF_BLOCKS, RIF_THREADS, 0, st>>>(keys, n, s, pass);
        if (comm) comm->allreduce(&s->hist[0][0], 4 * 256, CommDType::I32, CommOp::SUM, st);
        rif_pick_kernel<<<1, 256, 0, st>>>(s, pass);
    }
    rif_params_kernel<<<1, 1, 0, st>>>(s, n_glob, tau);
    rif_density_kernel<<<RIF_PARTS, RIF_THREADS, 0, st>>>(keys, rows, s);
    share_partials();
    rif_dens_final_kernel<<<1, 1, 0, st>>>(s, n_glob);
    if (n > 0) rif_apply_kernel<<<RIF_BLOCKS, RIF_THREADS, 0, st>>>(g.X, n, ycol, ldx, y_src, ystride, s, tau);
    OB_CUDA(cudaGetLastError());
}

}  // namespace ob
