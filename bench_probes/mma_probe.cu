// Microbenchmark: issue rate of the legacy tensor-core path (mma.sync.m16n8k8 TF32, m16n8k16 BF16) per SM on B200,
// to decide whether 3xTF32 mma.sync beats the FP32 pipe for the 32x32 / 32x64 projections of the block glue.
// Not part of the product.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

__global__ void k_tf32(float* out, int iters) {
    float c[8][4];
    unsigned a[4], b[2];
    for (int i = 0; i < 4; ++i) a[i] = __float_as_uint(1.0f + threadIdx.x * 1e-3f + i);
    for (int i = 0; i < 2; ++i) b[i] = __float_as_uint(0.5f + threadIdx.x * 1e-3f + i);
    for (int j = 0; j < 8; ++j) for (int i = 0; i < 4; ++i) c[j][i] = 0.f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < 8; ++j)
            asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+f"(c[j][0]), "+f"(c[j][1]), "+f"(c[j][2]), "+f"(c[j][3])
                         : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
    }
    float s = 0;
    for (int j = 0; j < 8; ++j) for (int i = 0; i < 4; ++i) s += c[j][i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_bf16(float* out, int iters) {
    float c[8][4];
    unsigned a[4], b[2];
    for (int i = 0; i < 4; ++i) a[i] = 0x3f803f80u + threadIdx.x + i;
    for (int i = 0; i < 2; ++i) b[i] = 0x3f003f00u + threadIdx.x + i;
    for (int j = 0; j < 8; ++j) for (int i = 0; i < 4; ++i) c[j][i] = 0.f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < 8; ++j)
            asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+f"(c[j][0]), "+f"(c[j][1]), "+f"(c[j][2]), "+f"(c[j][3])
                         : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
    }
    float s = 0;
    for (int j = 0; j < 8; ++j) for (int i = 0; i < 4; ++i) s += c[j][i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
    float* out; CK(cudaMalloc(&out, 148 * 1024 * 4));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    const int iters = 20000;
    for (int threads : {128, 256, 512, 1024}) {
        float ms;
        k_tf32<<<148, threads>>>(out, 100); CK(cudaDeviceSynchronize());
        CK(cudaEventRecord(e0)); k_tf32<<<148, threads>>>(out, iters); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        CK(cudaEventElapsedTime(&ms, e0, e1));
        double fma = 148.0 * (threads / 32) * (double)iters * 8 * 1024;   // m16n8k8 = 1024 FMA
        printf("tf32 m16n8k8  threads/SM=%4d: %.3f ms  %.1f TFMA/s  %.0f FMA/clk/SM @1.965GHz\n", threads, ms, fma / ms / 1e9, fma / (ms * 1e-3) / 148 / 1.965e9);
        k_bf16<<<148, threads>>>(out, 100); CK(cudaDeviceSynchronize());
        CK(cudaEventRecord(e0)); k_bf16<<<148, threads>>>(out, iters); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        CK(cudaEventElapsedTime(&ms, e0, e1));
        fma = 148.0 * (threads / 32) * (double)iters * 8 * 2048;          // m16n8k16 = 2048 FMA
        printf("bf16 m16n8k16 threads/SM=%4d: %.3f ms  %.1f TFMA/s  %.0f FMA/clk/SM @1.965GHz\n", threads, ms, fma / ms / 1e9, fma / (ms * 1e-3) / 148 / 1.965e9);
    }
    return 0;
}
