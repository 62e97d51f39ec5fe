// Microbenchmark: the inner loop of the block glue's matrix-vector products (per input channel: 8 broadcast LDS.128 of
// weights + 32 FFMA2 on 32 output-pair accumulators, scalar-broadcast input) as a function of warps per SM and of the
// LDS : FFMA2 ratio.  Question: what FP32-pipe utilisation can this instruction mix reach with 1, 2, 3, 4 warps per
// scheduler?  Not part of the product.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)
typedef float2 f2;

template <int NOUT, bool LDS>
__global__ void k(float* out, const float* w, int iters) {
    __shared__ __align__(16) float W[32 * 64];
    for (int i = threadIdx.x; i < 32 * 64; i += blockDim.x) W[i] = w[i];
    __syncthreads();
    f2 a0[NOUT / 2], a1[NOUT / 2], in[32];
#pragma unroll
    for (int i = 0; i < NOUT / 2; ++i) a0[i] = a1[i] = make_float2(threadIdx.x, i);
#pragma unroll
    for (int c = 0; c < 32; ++c) in[c] = make_float2(threadIdx.x * 1e-3f + c, c * 1e-3f);
    float4 wr[NOUT / 4];
#pragma unroll
    for (int q = 0; q < NOUT / 4; ++q) wr[q] = *reinterpret_cast<const float4*>(W + 4 * q);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int c = 0; c < 32; ++c) {
            const f2 x0 = make_float2(in[c].x, in[c].x), x1 = make_float2(in[c].y, in[c].y);
#pragma unroll
            for (int q = 0; q < NOUT / 4; ++q) {
                float4 ww = wr[q];
                if (LDS) ww = *reinterpret_cast<const float4*>(W + c * 64 + 4 * q + (it & 1) * 32);
                const f2 wa = make_float2(ww.x, ww.y), wb = make_float2(ww.z, ww.w);
                a0[2 * q] = __ffma2_rn(wa, x0, a0[2 * q]);
                a0[2 * q + 1] = __ffma2_rn(wb, x0, a0[2 * q + 1]);
                a1[2 * q] = __ffma2_rn(wa, x1, a1[2 * q]);
                a1[2 * q + 1] = __ffma2_rn(wb, x1, a1[2 * q + 1]);
            }
        }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < NOUT / 2; ++i) s += a0[i].x + a0[i].y + a1[i].x + a1[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NOUT, bool LDS>
void run(const char* name, float* out, const float* w) {
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    const int iters = 2000;
    for (int threads : {128, 256, 384, 512}) {
        float ms;
        k<NOUT, LDS><<<148, threads>>>(out, w, 10); CK(cudaDeviceSynchronize());
        CK(cudaEventRecord(e0)); k<NOUT, LDS><<<148, threads>>>(out, w, iters); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        CK(cudaEventElapsedTime(&ms, e0, e1));
        const double fma = 148.0 * threads * (double)iters * 32 * NOUT * 2;
        printf("%s warps/SM=%2d: %.3f ms  %.1f FMA/clk/SM @1.965GHz\n", name, threads / 32, ms, fma / (ms * 1e-3) / 148 / 1.965e9);
    }
}

int main() {
    float *out, *w; CK(cudaMalloc(&out, 148 * 1024 * 4)); CK(cudaMalloc(&w, 32 * 64 * 4)); CK(cudaMemset(w, 0, 32 * 64 * 4));
    run<32, false>("32 outputs, weights in registers  ", out, w);
    run<32, true>("32 outputs, 8 LDS.128 per channel ", out, w);
    run<8, true>(" 8 outputs, 2 LDS.128 per channel ", out, w);
    return 0;
}
