// Probe for the two operand forms of the tensor-core backward glue kernel (kind::tf32).  Not part of the product.
//  Part A: per-voxel GEMM with the A operand in TENSOR MEMORY: D[128 x 64] = A[128 x 32] W[64 x 32]^T, A written by
//          tcgen05.st (thread = lane = row, column = k), W in shared memory K-major without swizzle; 3xTF32 with the
//          lo parts in further TMEM columns.  Reports the error against the truncated-TF32 and the exact product.
//  Part B: contraction over voxels with MN-major (SWIZZLE_128B_BASE32B) operands whose ROWS are voxels:
//          D[128 x 64] = sum over 96 voxels of A[m][v] B[n][v]; A = 4 atoms of 32 rows m, B = 2 atoms of 32 columns n,
//          a voxel's 32 values = one 128-byte row, 32-byte chunks XOR-ed with (voxel % 4).
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t lt) {
    return (uint64_t)((saddr >> 4) & 0x3fff) | (uint64_t)((lbo >> 4) & 0x3fff) << 16 | (uint64_t)((sbo >> 4) & 0x3fff) << 32 | (uint64_t)1 << 46 | (uint64_t)(lt & 7) << 61;
}
__device__ __forceinline__ uint32_t idesc(int m, int n, int a_mn, int b_mn) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ float hi(float v) { return __uint_as_float(__float_as_uint(v) & 0xffffe000u); }

#define TMEM_LD32(taddr, r)                                                                                              \
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "                                                               \
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "                               \
                 "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"               \
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),       \
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), \
                   "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), \
                   "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]) \
                 : "r"(taddr) : "memory");                                                                               \
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory")
#define TMEM_ST32(taddr, r)                                                                                              \
    asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "                                                         \
                 "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "                              \
                 "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"                      \
                 :: "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),  \
                    "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]),        \
                    "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]),      \
                    "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])       \
                 : "memory")

// mode 0: part A single pass (hi only); mode 1: part A 3xTF32; mode 2: part B
__global__ void __launch_bounds__(128) probe(int mode, const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ D) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (tid == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&bar))); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" :: "r"(smem_u32(&tmem_base)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base;
    const uint32_t lane_addr = tmem + ((uint32_t)(warp * 32) << 16);
    if (mode < 2) {
        // A (128 x 32) row-major in global; W = B (64 x 32) row-major
        unsigned char* w_hi = smem;            // K-major no swizzle: (n >> 3) * 1024 + (k >> 2) * 128 + (n & 7) * 16 + (k & 3) * 4
        unsigned char* w_lo = smem + 8192;
        for (int e = tid; e < 64 * 32; e += 128) {
            const int n = e >> 5, k = e & 31;
            const uint32_t o = (n >> 3) * 1024 + (k >> 2) * 128 + (n & 7) * 16 + (k & 3) * 4;
            const float w = B[e];
            *reinterpret_cast<float*>(w_hi + o) = w;
            *reinterpret_cast<float*>(w_lo + o) = w - hi(w);
        }
        uint32_t rh[32], rl[32];
        for (int k = 0; k < 32; ++k) {
            const float a = A[tid * 32 + k];
            rh[k] = __float_as_uint(a);
            rl[k] = __float_as_uint(a - hi(a));
        }
        TMEM_ST32(lane_addr + 64, rh);         // A hi: columns 64..95, A lo: 96..127, D: 0..63
        TMEM_ST32(lane_addr + 96, rl);
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (tid == 0) {
            const uint32_t id = idesc(128, 64, 0, 0);
            for (int s = 0; s < 4; ++s) {
                const uint64_t bh = make_desc(smem_u32(w_hi) + s * 256, 128, 1024, 0);
                const uint64_t bl = make_desc(smem_u32(w_lo) + s * 256, 128, 1024, 0);
                const uint32_t a_hi = tmem + 64 + s * 8, a_lo = tmem + 96 + s * 8;
                if (mode == 1) {
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n"
                                 :: "r"(tmem), "r"(a_lo), "l"(bh), "r"(id), "r"((uint32_t)(s > 0)) : "memory");
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n"
                                 :: "r"(tmem), "r"(a_hi), "l"(bl), "r"(id), "r"(1u) : "memory");
                }
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n"
                             :: "r"(tmem), "r"(a_hi), "l"(bh), "r"(id), "r"((uint32_t)(mode == 1 || s > 0)) : "memory");
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(&bar)) : "memory");
        }
    } else {
        // A: (128 rows m) x (96 voxels), B: (64 rows n) x (96 voxels), both row-major [row][96] in global.
        // staged as voxel rows: atom a (32 values) of voxel v at a * ATOM + v * 128, chunk (i >> 3) ^ (v & 3), word i & 7
        constexpr int ATOM = 96 * 128;        // 12 KiB
        unsigned char* sa = smem;              // 4 atoms
        unsigned char* sb = smem + 4 * ATOM;   // 2 atoms
        for (int e = tid; e < 128 * 96; e += 128) {
            const int m = e / 96, v = e % 96;
            *reinterpret_cast<float*>(sa + (m >> 5) * ATOM + v * 128 + ((((m & 31) >> 3) ^ (v & 3)) << 5) + (m & 7) * 4) = A[e];
        }
        for (int e = tid; e < 64 * 96; e += 128) {
            const int n = e / 96, v = e % 96;
            *reinterpret_cast<float*>(sb + (n >> 5) * ATOM + v * 128 + ((((n & 31) >> 3) ^ (v & 3)) << 5) + (n & 7) * 4) = B[e];
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (tid == 0) {
            const uint32_t id = idesc(128, 64, 1, 1);
            for (int s = 0; s < 12; ++s) {
                const uint64_t ad = make_desc(smem_u32(sa) + s * 1024, ATOM, 512, 1);
                const uint64_t bd = make_desc(smem_u32(sb) + s * 1024, ATOM, 512, 1);
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
                             :: "r"(tmem), "l"(ad), "l"(bd), "r"(id), "r"((uint32_t)(s > 0)) : "memory");
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(&bar)) : "memory");
        }
    }
    uint32_t ok = 0;
    while (!ok) asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int half = 0; half < 2; ++half) {
        uint32_t r[32];
        TMEM_LD32(lane_addr + half * 32, r);
        for (int n = 0; n < 32; ++n) D[tid * 64 + half * 32 + n] = __uint_as_float(r[n]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" :: "r"(tmem) : "memory");
}

static float trunc_tf32(float v) { uint32_t u; memcpy(&u, &v, 4); u &= 0xffffe000u; memcpy(&v, &u, 4); return v; }

int main(int argc, char** argv) {
    const int mode = argc > 1 ? atoi(argv[1]) : 0;
    static float hA[128 * 96], hB[64 * 96], hD[128 * 64];
    srand(5);
    for (auto& v : hA) v = (float)rand() / RAND_MAX * 2.f - 1.f;
    for (auto& v : hB) v = (float)rand() / RAND_MAX * 2.f - 1.f;
    float *dA, *dB, *dD;
    CK(cudaMalloc(&dA, sizeof(hA))); CK(cudaMalloc(&dB, sizeof(hB))); CK(cudaMalloc(&dD, sizeof(hD)));
    CK(cudaMemcpy(dA, hA, sizeof(hA), cudaMemcpyHostToDevice)); CK(cudaMemcpy(dB, hB, sizeof(hB), cudaMemcpyHostToDevice));
    CK(cudaMemset(dD, 0, sizeof(hD)));
    const int smem = 6 * 96 * 128 + 1024;
    CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    probe<<<1, 128, smem>>>(mode, dA, dB, dD);
    cudaError_t e = cudaDeviceSynchronize();
    printf("mode %d: %s\n", mode, cudaGetErrorString(e));
    if (e != cudaSuccess) return 1;
    CK(cudaMemcpy(hD, dD, sizeof(hD), cudaMemcpyDeviceToHost));
    double err_exact = 0, err_trunc = 0;
    const int K = mode < 2 ? 32 : 96;
    for (int m = 0; m < 128; ++m)
        for (int n = 0; n < 64; ++n) {
            double ex = 0, tr = 0;
            for (int k = 0; k < K; ++k) {
                const float a = hA[m * K + k], b = hB[n * K + k];
                ex += (double)a * b;
                tr += (double)trunc_tf32(a) * trunc_tf32(b);
            }
            err_exact = fmax(err_exact, fabs(hD[m * 64 + n] - ex));
            err_trunc = fmax(err_trunc, fabs(hD[m * 64 + n] - tr));
        }
    printf("  max |D - exact| = %.3g   max |D - truncated-TF32 product| = %.3g   D[0][0..3] = %g %g %g %g\n", err_exact, err_trunc,
           hD[0], hD[1], hD[2], hD[3]);
    return 0;
}
