// tools/gram_probe.cu -- isolates the cost of each ingredient of gram.cu's inner loop on static
// shared-memory data (no TMA): MODE bit 0 = B fragment as product of two loads (+DMUL) instead of one load,
// bit 1 = interleaved count-widening step, bit 2 = __syncthreads per 32-row stage, bit 3 = pair offsets via
// runtime registers (else compile-time consecutive columns).
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "CUDA %s %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)
constexpr int LDA = 132, KT = 32, LDX = 52;
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
template <int MODE>
__global__ void __launch_bounds__(256, 1) probe(double* out, const uint16_t* pairs, int stages) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* As = reinterpret_cast<double*>(smem_raw);
    double* Xs = As + 2 * KT * LDA;
    uint8_t* Cr = reinterpret_cast<uint8_t*>(Xs + 4 * KT * LDX);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, wm = warp >> 2, wn = warp & 3, lk = lane & 3, lg = lane >> 2;
    for (int i = tid; i < 2 * KT * LDA; i += 256) As[i] = (i % 7) * 0.25;
    for (int i = tid; i < 4 * KT * LDX; i += 256) Xs[i] = 1.0 + (i % 5) * 0.125;
    for (int i = tid; i < 4 * KT * 128; i += 256) Cr[i] = i % 3;
    __syncthreads();
    int oj[4], ol[4];
    for (int s = 0; s < 4; ++s) {
        const int col = wn * 32 + s * 8 + lg;
        if (MODE & 8) { oj[s] = pairs[2 * col]; ol[s] = pairs[2 * col + 1]; } else { oj[s] = 0; ol[s] = col % LDX; }
    }
    double acc[8][4][2];
    for (int i = 0; i < 8; ++i) for (int s = 0; s < 4; ++s) { acc[i][s][0] = 0; acc[i][s][1] = 0; }
    const int cr = tid >> 3, cq = tid & 7;
    for (int s = 0; s < stages; ++s) {
        const int slot = s & 3;
        const double* abase = As + (s & 1) * KT * LDA + lk * LDA + wm * 64 + lg;
        const double* xbase = Xs + slot * KT * LDX + lk * LDX;
        const uint8_t* nsrc = Cr + ((s + 1) & 3) * KT * 128 + cr * 128 + cq * 2;
        double* ndst = As + ((s + 1) & 1) * KT * LDA + cr * LDA + cq * 2;
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
            double a[8], b[4];
            const double* arow = abase + kk * 4 * LDA;
            const double* xrow = xbase + kk * 4 * LDX;
#pragma unroll
            for (int i = 0; i < 8; ++i) a[i] = arow[i * 8];
#pragma unroll
            for (int t = 0; t < 4; ++t) b[t] = (MODE & 1) ? xrow[oj[t]] * xrow[ol[t]] : xrow[ol[t]];
            if (MODE & 2) {
                const unsigned v = *reinterpret_cast<const uint16_t*>(nsrc + kk * 16);
                double2 o;
                o.x = fma(__hiloint2double(0x43300000, (int)(v & 0xFF)), 1.5, -4503599627370496.0 * 1.5);
                o.y = fma(__hiloint2double(0x43300000, (int)(v >> 8)), 1.5, -4503599627370496.0 * 1.5);
                *reinterpret_cast<double2*>(ndst + kk * 16) = o;
            }
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int t = 0; t < 4; ++t) dmma884(acc[i][t][0], acc[i][t][1], a[i], b[t]);
        }
        if (MODE & 4) __syncthreads();
    }
    double sum = 0;
    for (int i = 0; i < 8; ++i) for (int s = 0; s < 4; ++s) sum += acc[i][s][0] + acc[i][s][1];
    if (sum == 123.456) out[0] = sum;
}
template <int MODE> void run(double* out, const uint16_t* pairs, int sms) {
    const int stages = 4000; const size_t smem = 8 * (2 * KT * LDA + 4 * KT * LDX) + 4 * KT * 128;
    CK(cudaFuncSetAttribute(probe<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    probe<MODE><<<sms, 256, smem>>>(out, pairs, stages); CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < 3; ++r) { cudaEventRecord(e0); probe<MODE><<<sms, 256, smem>>>(out, pairs, stages); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); best = ms < best ? ms : best; }
    CK(cudaGetLastError());
    const double flop = 2.0 * 128 * 128 * 32 * stages * sms;
    printf("{\"mode\": %d, \"b_product\": %d, \"widen\": %d, \"barrier\": %d, \"runtime_pairs\": %d, \"ms\": %.3f, \"tflops\": %.2f}\n", MODE, MODE & 1, (MODE >> 1) & 1, (MODE >> 2) & 1, (MODE >> 3) & 1, best, flop / best * 1e-9);
}
int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    double* out; CK(cudaMalloc(&out, 64));
    uint16_t h[256]; int c = 0;
    for (int j = 0; j < 52 && c < 128; ++j) for (int l = j; l < 52 && c < 128; ++l) { h[2 * c] = j; h[2 * c + 1] = l; ++c; }
    uint16_t* d; CK(cudaMalloc(&d, sizeof h)); CK(cudaMemcpy(d, h, sizeof h, cudaMemcpyHostToDevice));
    run<0>(out, d, p.multiProcessorCount); run<1>(out, d, p.multiProcessorCount); run<2>(out, d, p.multiProcessorCount);
    run<4>(out, d, p.multiProcessorCount); run<5>(out, d, p.multiProcessorCount); run<7>(out, d, p.multiProcessorCount);
    run<8>(out, d, p.multiProcessorCount); run<9>(out, d, p.multiProcessorCount); run<15>(out, d, p.multiProcessorCount);
    return 0;
}
