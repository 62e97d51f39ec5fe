// Microbenchmark: FP32 FMA issue rate with scalar FFMA vs packed fma.rn.f32x2 (FFMA2), and warp
// shuffle rate, per SM on B200.  Not part of the product.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

__global__ void k_ffma(float* out, float a, float b, int iters) {
    float acc[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = threadIdx.x + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) acc[i] = fmaf(acc[i], a, b);
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_ffma2(float* out, float a, float b, int iters) {
    unsigned long long acc[8];
    unsigned long long av, bv;
    asm("mov.b64 %0, {%1, %1};" : "=l"(av) : "f"(a));
    asm("mov.b64 %0, {%1, %1};" : "=l"(bv) : "f"(b));
#pragma unroll
    for (int i = 0; i < 8; ++i) { float lo = threadIdx.x + i, hi = threadIdx.x - i; asm("mov.b64 %0, {%1, %2};" : "=l"(acc[i]) : "f"(lo), "f"(hi)); }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(acc[i]) : "l"(av), "l"(bv));
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) { float lo, hi; asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(acc[i])); s += lo + hi; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_shfl(float* out, int iters) {
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = threadIdx.x * 1.5f + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = __shfl_xor_sync(0xffffffffu, v[i], 1 + (i & 15));
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
    float* out; CK(cudaMalloc(&out, 148 * 1024 * 8 * 4));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    int clk_khz = 0; CK(cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0));
    printf("clock rate attr %d kHz\n", clk_khz);
    const int iters = 20000;
    for (int threads : {128, 256, 512, 1024}) {
        float ms;
        for (int rep = 0; rep < 2; ++rep) { CK(cudaEventRecord(e0)); k_ffma<<<148, threads>>>(out, 1.0001f, 0.5f, iters); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1)); }
        double fma = 148.0 * threads * 16.0 * iters;
        printf("FFMA  threads/SM=%4d: %.3f ms  %.1f TFMA/s  (%.1f FMA/clk/SM @1.965GHz)\n", threads, ms, fma / ms / 1e9, fma / (ms * 1e-3) / 148 / 1.965e9);
        for (int rep = 0; rep < 2; ++rep) { CK(cudaEventRecord(e0)); k_ffma2<<<148, threads>>>(out, 1.0001f, 0.5f, iters); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1)); }
        printf("FFMA2 threads/SM=%4d: %.3f ms  %.1f TFMA/s  (%.1f FMA/clk/SM @1.965GHz)\n", threads, ms, fma / ms / 1e9, fma / (ms * 1e-3) / 148 / 1.965e9);
        for (int rep = 0; rep < 2; ++rep) { CK(cudaEventRecord(e0)); k_shfl<<<148, threads>>>(out, iters); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1)); }
        double sh = 148.0 * threads * 8.0 * iters;
        printf("SHFL  threads/SM=%4d: %.3f ms  (%.1f lanes/clk/SM @1.965GHz)\n", threads, ms, sh / (ms * 1e-3) / 148 / 1.965e9);
    }
    return 0;
}
