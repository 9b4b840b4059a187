// csrc/reduce_stats.cu -- replicate statistics -> standard errors, p-values, percentile CIs.
//
// Restates bootstrap_stats (inference.rs:4-34) and process_component's t statistic
// (builder.rs:849-865) for all S statistics at once: one CTA per statistic; failed replicates are
// masked out (filter_map, builder.rs:816-839), B' = number of successes.
//   mean, sample SD (B'-1)      two-pass, fixed-order block reduction (deterministic)
//   p = min(1, 2 min(#>=0, #<=0)/B')
//   CI = sorted[floor(.025 B')], sorted[min(floor(.975 B'), B'-1)]   (bitonic sort: in shared memory up to 16384
//        replicates, in a global scratch row per statistic beyond that)
//   t = point/se if |se| > 1e-9 else 0;   B' = 0 -> all NaN, t = 0
#include "common.cuh"
#include "internal.h"

namespace ob {

constexpr int RS_THREADS = 1024;

__device__ __forceinline__ double block_sum(double v, double* red) {
    // fixed-order tree: warp shuffle then 32 partials summed by warp 0
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = (threadIdx.x < 32) ? red[threadIdx.x] : 0.0;
    if (threadIdx.x < 32) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        if (threadIdx.x == 0) red[32] = t;
    }
    __syncthreads();
    return red[32];
}

__global__ void __launch_bounds__(RS_THREADS) reduce_stats_kernel(const double* __restrict__ stats,
                                                                  const int* __restrict__ status, long long reps,
                                                                  int S, const double* __restrict__ point,
                                                                  double* __restrict__ out, long long* __restrict__ n_ok_out,
                                                                  int npow2, double* __restrict__ gscratch) {
    extern __shared__ __align__(16) double vs[];  // [npow2] when it fits
    double* v = gscratch ? gscratch + (size_t)blockIdx.x * npow2 : vs;
    __shared__ double red[33];
    const int j = blockIdx.x, tid = threadIdx.x;
    const double inf = __longlong_as_double(0x7ff0000000000000LL);
    const double nan = __longlong_as_double(0x7ff8000000000000LL);

    double cnt = 0.0, sum = 0.0, pos = 0.0, neg = 0.0;
    for (int b = tid; b < npow2; b += RS_THREADS) {
        double x = inf;
        if (b < reps && status[b] == OB_OK) {
            x = stats[(size_t)b * S + j];
            cnt += 1.0; sum += x; pos += (x >= 0.0) ? 1.0 : 0.0; neg += (x <= 0.0) ? 1.0 : 0.0;
        }
        v[b] = x;
    }
    const double n = block_sum(cnt, red);
    const double total = block_sum(sum, red);
    const double npos = block_sum(pos, red);
    const double nneg = block_sum(neg, red);
    const double mean = total / n;
    double ss = 0.0;
    for (int b = tid; b < npow2; b += RS_THREADS) {
        const double x = v[b];
        if (b < reps && status[b] == OB_OK) { const double d = x - mean; ss += d * d; }
    }
    const double sst = block_sum(ss, red);

    // bitonic sort ascending
    for (int k = 2; k <= npow2; k <<= 1) {
        for (int s = k >> 1; s > 0; s >>= 1) {
            __syncthreads();
            for (int i = tid; i < npow2; i += RS_THREADS) {
                const int l = i ^ s;
                if (l > i) {
                    const double a = v[i], b = v[l];
                    const bool up = (i & k) == 0;
                    if ((a > b) == up) { v[i] = b; v[l] = a; }
                }
            }
        }
    }
    __syncthreads();
    if (tid == 0) {
        double se = nan, pv = nan, lo = nan, hi = nan;
        if (n > 0.0) {
            se = sqrt(sst / (n - 1.0));
            pv = fmin(2.0 * fmin(npos / n, nneg / n), 1.0);
            const long long nn = (long long)n;
            long long li = (long long)floor(0.025 * n), hi_i = (long long)floor(0.975 * n);
            if (hi_i > nn - 1) hi_i = nn - 1;
            lo = (li < nn) ? v[li] : nan;
            hi = v[hi_i];
        }
        out[0 * S + j] = se; out[1 * S + j] = pv; out[2 * S + j] = lo; out[3 * S + j] = hi;
        out[4 * S + j] = (fabs(se) > 1e-9) ? point[j] / se : 0.0;   // builder.rs:851-855 (NaN compares false)
        if (j == 0) *n_ok_out = (long long)n;
    }
}

void reduce_stats_launch(const double* stats, const int* status, int64_t reps, int S, const double* point_stats,
                         double* out5S, long long* n_ok, cudaStream_t st, double* d_scratch) {
    if (reps > (1ll << 30)) throw StatusError{OB_ERR_UNSUPPORTED, "more than 2^30 replicates in one reduction"};
    int npow2 = 2;
    while (npow2 < reps) npow2 <<= 1;
    if (reps > REDUCE_MAX_REPS && !d_scratch) throw StatusError{OB_ERR_INVALID_ARG, "reduction scratch missing"};
    const bool in_smem = reps <= REDUCE_MAX_REPS;
    const size_t smem = in_smem ? sizeof(double) * (size_t)npow2 : 0;
    OB_CUDA(cudaFuncSetAttribute(reduce_stats_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(double) * REDUCE_MAX_REPS)));
    reduce_stats_kernel<<<S, RS_THREADS, smem, st>>>(stats, status, (long long)reps, S, point_stats, out5S, n_ok, npow2,
                                                     in_smem ? nullptr : d_scratch);
    OB_CUDA(cudaGetLastError());
}

size_t reduce_stats_scratch_bytes(int64_t reps, int S) {
    if (reps <= REDUCE_MAX_REPS) return 0;
    size_t npow2 = 2;
    while ((int64_t)npow2 < reps) npow2 <<= 1;
    return sizeof(double) * npow2 * (size_t)S;
}

}  // namespace ob
