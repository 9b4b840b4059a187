// tools/fp64_peaks.cu -- measures the FP64 rooflines the Gram contraction is judged against.
//
//   (1) register-only DMMA.8x8x4 issue rate (mma.sync.m8n8k4.f64), swept over warps/SM and
//       independent accumulator chains per warp  -> the FP64 tensor peak of this B200
//   (2) register-only DFMA rate                  -> the FP64 vector peak
//   (3) a DMMA loop with one shared-memory fragment load per operand -> the smem-fed ceiling
// cuBLAS DGEMM is timed from Python (tools/fp64_peaks.py) in the same gpurun call.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/fp64_peaks tools/fp64_peaks.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
    fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int CHAINS>
__global__ void __launch_bounds__(512) dmma_regs(double* out, int iters, double seed) {
    double c0[CHAINS], c1[CHAINS];
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) { c0[i] = 0.0; c1[i] = 0.0; }
    double a = seed + threadIdx.x * 1e-9, b = seed - threadIdx.x * 1e-9;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < CHAINS; ++i) dmma884(c0[i], c1[i], a, b);
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) s += c0[i] + c1[i];
    if (s == 123.456) out[0] = s;
}

template <int CHAINS>
__global__ void __launch_bounds__(512) dfma_regs(double* out, int iters, double seed) {
    double c[CHAINS];
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) c[i] = i;
    double a = seed + threadIdx.x * 1e-9, b = seed * 1e-3;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < CHAINS; ++i) c[i] = fma(c[i], a, b);
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) s += c[i];
    if (s == 123.456) out[0] = s;
}

// warp tile MT x NT sub-tiles of 8x8; per k-step: MT + NT smem loads, MT*NT DMMA
template <int MT, int NT>
__global__ void __launch_bounds__(MT * NT >= 32 ? 256 : 512) dmma_smem(double* out, int iters) {
    extern __shared__ double sm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 8192; i += blockDim.x) sm[i] = 1e-3 * i;
    __syncthreads();
    double c0[MT][NT], c1[MT][NT];
#pragma unroll
    for (int i = 0; i < MT; ++i)
#pragma unroll
        for (int j = 0; j < NT; ++j) { c0[i][j] = 0; c1[i][j] = 0; }
    const double* pa = sm + (lane & 3) * 132 + (lane >> 2) + warp * 8;
    const double* pb = sm + 3968 + (lane & 3) * 132 + (lane >> 2);
    for (int it = 0; it < iters; ++it) {
        double a[MT], b[NT];
        const int off = (it & 7) * 4 * 132;
#pragma unroll
        for (int i = 0; i < MT; ++i) a[i] = pa[off + i * 8];
#pragma unroll
        for (int j = 0; j < NT; ++j) b[j] = pb[off + j * 8];
#pragma unroll
        for (int i = 0; i < MT; ++i)
#pragma unroll
            for (int j = 0; j < NT; ++j) dmma884(c0[i][j], c1[i][j], a[i], b[j]);
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < MT; ++i)
#pragma unroll
        for (int j = 0; j < NT; ++j) s += c0[i][j] + c1[i][j];
    if (s == 123.456) out[0] = s;
}

template <typename F>
static double time_ms(F launch, int reps = 5) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    launch(); launch();
    CK(cudaDeviceSynchronize());
    double best = 1e30;
    for (int r = 0; r < reps; ++r) {
        CK(cudaEventRecord(e0));
        launch();
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    CK(cudaGetLastError());
    return best;
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    const int sms = p.multiProcessorCount;
    double* out; CK(cudaMalloc(&out, 64));
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"rows\": [\n", p.name, sms);
    const int iters = 20000;
    bool first = true;
    auto row = [&](const char* kind, int warps, int chains, double ms, double flop) {
        printf("%s  {\"kind\": \"%s\", \"warps_per_sm\": %d, \"chains\": %d, \"ms\": %.4f, \"tflops\": %.3f}",
               first ? "" : ",\n", kind, warps, chains, ms, flop / ms * 1e-9);
        first = false;
    };
    for (int warps : {4, 8, 16}) {
#define RUN_DMMA(CH) { double ms = time_ms([&] { dmma_regs<CH><<<sms, warps * 32>>>(out, iters, 1.0); }); \
        row("dmma_regs", warps, CH, ms, 512.0 * CH * iters * warps * sms); }
        RUN_DMMA(1) RUN_DMMA(2) RUN_DMMA(4) RUN_DMMA(8) RUN_DMMA(16)
#define RUN_DFMA(CH) { double ms = time_ms([&] { dfma_regs<CH><<<sms, warps * 32>>>(out, iters, 1.0); }); \
        row("dfma_regs", warps, CH, ms, 64.0 * CH * iters * warps * sms); }
        RUN_DFMA(4) RUN_DFMA(8) RUN_DFMA(16)
    }
    CK(cudaFuncSetAttribute(dmma_smem<8, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
    CK(cudaFuncSetAttribute(dmma_smem<4, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
    CK(cudaFuncSetAttribute(dmma_smem<4, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
    for (int warps : {4, 8, 16}) {
        if (warps <= 8) { double ms = time_ms([&] { dmma_smem<8, 4><<<sms, warps * 32, 65536>>>(out, iters); });
          row("dmma_smem_8x4", warps, 32, ms, 512.0 * 32 * iters * warps * sms); }
        { double ms = time_ms([&] { dmma_smem<4, 4><<<sms, warps * 32, 65536>>>(out, iters); });
          row("dmma_smem_4x4", warps, 16, ms, 512.0 * 16 * iters * warps * sms); }
        { double ms = time_ms([&] { dmma_smem<4, 2><<<sms, warps * 32, 65536>>>(out, iters); });
          row("dmma_smem_4x2", warps, 8, ms, 512.0 * 8 * iters * warps * sms); }
    }
    printf("\n]}\n");
    return 0;
}
