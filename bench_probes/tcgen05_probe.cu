// Probe: one tcgen05.mma kind::tf32 tile, D[128 voxels x 32 outputs] = A[128 x 32] * W[32 x 32]^T, with
//   A  = activations as they lie in an NCDHW tensor: [channel k][voxel m], voxel-contiguous = "MN-major", staged in
//        shared memory as SWIZZLE_128B_BASE32B atoms of 4 channels x 32 voxels (what a voxel-owning warp writes
//        conflict-free, one 128-byte row per channel);
//   W  = [output n][channel k] row-major = "K-major", no swizzle, 8 x 16-byte core matrices;
//   D  = TMEM, lane = voxel, column = output: tcgen05.ld 32x32b hands every thread ITS voxel's 32 outputs.
// Checks (1) the plain TF32 product against a TF32-truncated host product, (2) the 3xTF32 split
// (a_hi b_hi + a_lo b_hi + a_hi b_lo, lo = x - trunc_tf32(x) computed elementwise on the staged buffer) against fp64.
// Purpose: pin down the descriptor encodings for a tensor-core version of the block glue.  Not part of the product.
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

constexpr int M = 128, N = 32, K = 32;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout_type) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3fff);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;
    d |= (uint64_t)1 << 46;                       // descriptor version (Blackwell)
    d |= (uint64_t)(layout_type & 7) << 61;
    return d;
}

__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
        :: "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}

// byte offset of A element (voxel m, channel k) in the swizzled staging buffer.  MN-major TF32 operands have exactly one
// legal shared-memory layout, SWIZZLE_128B_BASE32B (cutlass sm100_common.inl:92): atoms of 4 channels x 32 voxels (512 B),
// a channel row = 128 contiguous bytes, its four 32-byte chunks XOR-ed with (channel % 4) (Swizzle<2,5,2> on byte addresses).
__device__ __forceinline__ uint32_t a_offset(int m, int k) {
    const int r = k & 3, j = m & 31;
    return (uint32_t)((k >> 2) * 2048 + (m >> 5) * 512 + r * 128 + (((j >> 3) ^ r) << 5) + (j & 7) * 4);
}
// byte offset of W element (output n, channel k): core matrices of 8 outputs x 4 channels (128 B), K-adjacent
__device__ __forceinline__ uint32_t b_offset(int n, int k) {
    return (uint32_t)((n >> 3) * 1024 + (k >> 2) * 128 + (n & 7) * 16 + (k & 3) * 4);
}

// K-major, no swizzle A: core matrices of 8 voxels x 4 channels (128 B); K-adjacent cores 128 B apart, M groups 1 KiB apart
__device__ __forceinline__ uint32_t a_offset_k(int m, int k) {
    return (uint32_t)((m >> 3) * 1024 + (k >> 2) * 128 + (m & 7) * 16 + (k & 3) * 4);
}

__global__ void __launch_bounds__(128) probe(const float* __restrict__ A /*[K][M]*/, const float* __restrict__ W /*[N][K]*/,
                                              float* __restrict__ D1 /*[M][N] plain tf32*/, float* __restrict__ D3 /*3xTF32*/,
                                              int variant, uint32_t* __restrict__ info) {
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* a_hi = smem;                 // 16 KiB
    unsigned char* a_lo = smem + 16384;         // 16 KiB
    unsigned char* b_hi = smem + 32768;         // 4 KiB
    unsigned char* b_lo = smem + 36864;         // 4 KiB
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5;

    // stage A: thread = voxel, loops channels (conflict-free: a warp writes one 128-byte swizzled row per channel)
    for (int k = 0; k < K; ++k) {
        const float v = A[k * M + tid];
        const float hi = __uint_as_float(__float_as_uint(v) & 0xffffe000u);     // what the tensor core keeps of a TF32 input
        const uint32_t off = variant == 0 ? a_offset(tid, k) : a_offset_k(tid, k);
        *reinterpret_cast<float*>(a_hi + off) = v;
        *reinterpret_cast<float*>(a_lo + off) = v - hi;
    }
    for (int e = tid; e < N * K; e += 128) {
        const int n = e / K, k = e % K;
        const float v = W[e];
        const float hi = __uint_as_float(__float_as_uint(v) & 0xffffe000u);
        *reinterpret_cast<float*>(b_hi + b_offset(n, k)) = v;
        *reinterpret_cast<float*>(b_lo + b_offset(n, k)) = v - hi;
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" :: "r"(smem_u32(&tmem_base)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // generic-proxy writes to shared memory must be visible to the tensor core (async proxy)
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base;

    // instruction descriptor: D = F32, A = B = TF32, A MN-major, B K-major, N = 32, M = 128
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((variant == 0 ? 1u : 0u) << 15) | (0u << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
    if (tid == 0) { info[0] = tmem; info[1] = idesc; info[2] = smem_u32(a_hi); info[3] = smem_u32(b_hi); }
    if (tid == 0) {
        // columns [0, 32): plain product; columns [32, 64): 3xTF32
        for (int s = 0; s < K / 8; ++s) {
            const uint64_t ah = variant == 0 ? make_desc(smem_u32(a_hi) + s * 4096, 512, 2048, 1)
                                             : make_desc(smem_u32(a_hi) + s * 256, 128, 1024, 0);
            const uint64_t bh = make_desc(smem_u32(b_hi) + s * 256, 128, 1024, 0);
            mma_tf32(tmem, ah, bh, idesc, s > 0);
        }
        for (int s = 0; s < K / 8; ++s) {
            const uint64_t ah = variant == 0 ? make_desc(smem_u32(a_hi) + s * 4096, 512, 2048, 1)
                                             : make_desc(smem_u32(a_hi) + s * 256, 128, 1024, 0);
            const uint64_t al = variant == 0 ? make_desc(smem_u32(a_lo) + s * 4096, 512, 2048, 1)
                                             : make_desc(smem_u32(a_lo) + s * 256, 128, 1024, 0);
            const uint64_t bh = make_desc(smem_u32(b_hi) + s * 256, 128, 1024, 0);
            const uint64_t bl = make_desc(smem_u32(b_lo) + s * 256, 128, 1024, 0);
            mma_tf32(tmem + 32, al, bh, idesc, s > 0);
            mma_tf32(tmem + 32, ah, bl, idesc, 1);
            mma_tf32(tmem + 32, ah, bh, idesc, 1);
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(&bar)) : "memory");
    }
    // wait for the MMAs
    {
        uint32_t ok = 0;
        while (!ok) {
            asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                         : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
        }
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // each warp reads its 32 lanes (= voxels): 32 columns per load
    for (int half = 0; half < 2; ++half) {
        uint32_t r[32];
        const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + half * 32;
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
            "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
            "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
              "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
              "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
              "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
            : "r"(taddr) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        float* out = half ? D3 : D1;
        for (int n = 0; n < 32; ++n) out[tid * N + n] = __uint_as_float(r[n]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" :: "r"(tmem) : "memory");
}

static float trunc_tf32(float v) { uint32_t u; memcpy(&u, &v, 4); u &= 0xffffe000u; memcpy(&v, &u, 4); return v; }

int main() {
    float hA[K * M], hW[N * K];
    srand(1);
    for (int i = 0; i < K * M; ++i) hA[i] = (float)rand() / RAND_MAX * 2.f - 1.f;
    for (int i = 0; i < N * K; ++i) hW[i] = (float)rand() / RAND_MAX * 2.f - 1.f;
    float *dA, *dW, *d1, *d3;
    CK(cudaMalloc(&dA, sizeof(hA))); CK(cudaMalloc(&dW, sizeof(hW))); CK(cudaMalloc(&d1, M * N * 4)); CK(cudaMalloc(&d3, M * N * 4));
    CK(cudaMemcpy(dA, hA, sizeof(hA), cudaMemcpyHostToDevice)); CK(cudaMemcpy(dW, hW, sizeof(hW), cudaMemcpyHostToDevice));
    uint32_t* dinfo; CK(cudaMalloc(&dinfo, 16));
    CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 40960 + 1024));
    for (int variant = 0; variant < 2; ++variant) {
        CK(cudaMemset(d1, 0xff, M * N * 4)); CK(cudaMemset(d3, 0xff, M * N * 4));
        probe<<<1, 128, 40960 + 1024>>>(dA, dW, d1, d3, variant, dinfo);
        CK(cudaDeviceSynchronize());
        static float h1[M * N], h3[M * N];
        uint32_t hinfo[4];
        CK(cudaMemcpy(h1, d1, sizeof(h1), cudaMemcpyDeviceToHost)); CK(cudaMemcpy(h3, d3, sizeof(h3), cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(hinfo, dinfo, 16, cudaMemcpyDeviceToHost));
        double e1 = 0, e3 = 0, e1f = 0;
        for (int m = 0; m < M; ++m)
            for (int n = 0; n < N; ++n) {
                double rt = 0, rf = 0;
                for (int k = 0; k < K; ++k) {
                    rt += (double)trunc_tf32(hA[k * M + m]) * (double)trunc_tf32(hW[n * K + k]);
                    rf += (double)hA[k * M + m] * (double)hW[n * K + k];
                }
                e1 = fmax(e1, fabs(h1[m * N + n] - rt));
                e1f = fmax(e1f, fabs(h1[m * N + n] - rf));
                e3 = fmax(e3, fabs(h3[m * N + n] - rf));
            }
        printf("variant %d (%s): tmem base 0x%08x idesc 0x%08x smem A 0x%x B 0x%x\n", variant,
               variant == 0 ? "A MN-major SWIZZLE_128B_BASE32B" : "A K-major no swizzle", hinfo[0], hinfo[1], hinfo[2], hinfo[3]);
        printf("  plain TF32: max |D - trunc-TF32 reference| = %.3g (vs exact product %.3g)\n", e1, e1f);
        printf("  3xTF32    : max |D - exact product|        = %.3g\n", e3);
        int nz1 = 0, nz3 = 0;
        for (int i = 0; i < M * N; ++i) { nz1 += h1[i] != 0.f; nz3 += h3[i] != 0.f; }
        printf("  nonzeros D1 %d D3 %d of %d\n", nz1, nz3, M * N);
        for (int m : {0, 1, 33, 127}) {
            double rt = 0; for (int k = 0; k < K; ++k) rt += (double)trunc_tf32(hA[k * M + m]) * (double)trunc_tf32(hW[0 * K + k]);
            printf("  row %3d: D1 = %9.5f %9.5f %9.5f %9.5f   ref[0] = %9.5f\n", m, h1[m * N], h1[m * N + 1], h1[m * N + 2], h1[m * N + 3], rt);
        }
        printf(e1 < 1e-4 && e3 < 2e-5 ? "  PROBE OK\n" : "  PROBE MISMATCH\n");
    }
    return 0;
}
