// Address discovery for tcgen05 shared-memory descriptors (kind::tf32): which 32-bit word of the B buffer does the tensor
// core read as element (n, k)?  A = identity on rows 0..7 (K-major, no swizzle: a layout already pinned), the B buffer
// holds its own word index in every word, so D[k][n] = index of the word read as B(n, k).  One MMA (M = 128, K = 8) per
// configuration.  Not part of the product.
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t lt) {
    return (uint64_t)((saddr >> 4) & 0x3fff) | (uint64_t)((lbo >> 4) & 0x3fff) << 16 | (uint64_t)((sbo >> 4) & 0x3fff) << 32 | (uint64_t)1 << 46 | (uint64_t)(lt & 7) << 61;
}

struct Cfg { int b_mn_major, type, lbo, sbo, start, N; };

__global__ void __launch_bounds__(128) probe(Cfg c, float* __restrict__ D /*[128][64]*/) {
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* a = smem;                        // 128 rows x 8 k, K-major no swizzle: core matrices (8 rows x 16 B)
    unsigned char* b = smem + 8192;                 // 8 KiB: word i holds (float)i
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5;
    // A(m, k) at (m >> 3) * 256 + (k >> 2) * 128 + (m & 7) * 16 + (k & 3) * 4  (LBO = 128 along K, SBO = 256 along M)
    for (int k = 0; k < 8; ++k)
        *reinterpret_cast<float*>(a + (tid >> 3) * 256 + (k >> 2) * 128 + (tid & 7) * 16 + (k & 3) * 4) = (tid == k) ? 1.f : 0.f;
    for (int i = tid; i < 2048; i += 128) reinterpret_cast<float*>(b)[i] = (float)i;
    if (tid == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&bar))); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" :: "r"(smem_u32(&tmem_base)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base;
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(c.b_mn_major ? 1 : 0) << 16) | ((uint32_t)(c.N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    if (tid == 0) {
        const uint64_t ad = make_desc(smem_u32(a), 128, 256, 0);
        const uint64_t bd = make_desc(smem_u32(b) + c.start, c.lbo, c.sbo, c.type);
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
                     :: "r"(tmem), "l"(ad), "l"(bd), "r"(idesc), "r"(0u) : "memory");
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(&bar)) : "memory");
    }
    uint32_t ok = 0;
    while (!ok) asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
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
        if (half * 32 < c.N) for (int n = 0; n < 32; ++n) D[tid * 64 + half * 32 + n] = __uint_as_float(r[n]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" :: "r"(tmem) : "memory");
}

int main(int argc, char** argv) {
    const int only = argc > 1 ? atoi(argv[1]) : -1;
    static float hD[128 * 64];
    float* dD;
    CK(cudaMalloc(&dD, sizeof(hD)));
    const int smem = 8192 + 8192 + 1024;
    CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    const Cfg cfgs[] = {
        // b_mn_major, type, lbo, sbo, start, N
        {0, 0, 128, 256, 0, 32},      // control: K-major, no swizzle
        {0, 2, 16, 1024, 0, 32},      // control: K-major SWIZZLE_128B
        {0, 2, 16, 1024, 32, 32},     //          ... second K step
        {0, 1, 16, 1024, 0, 32},      // H1: K-major SWIZZLE_128B_BASE32B
        {0, 1, 16, 1024, 32, 32},     //          ... second K step
        {0, 1, 16, 1024, 96, 32},     //          ... fourth K step
        {0, 1, 16, 2048, 0, 64},      //          ... N = 64, row groups 2 KiB apart
        {1, 1, 4096, 512, 0, 32},     // MN-major B, BASE32B: lbo = N-atom stride?, sbo = K-atom stride?
        {1, 1, 512, 4096, 0, 32},     //          ... swapped
        {1, 1, 2048, 512, 0, 64},     //          ... N = 64: two N atoms
        {1, 2, 4096, 1024, 0, 32},    // MN-major B with the plain 128B swizzle
        {0, 6, 16, 256, 0, 32},       // K-major SWIZZLE_32B
        {1, 0, 128, 1024, 0, 32},     // MN-major B without swizzle: lbo between groups of 4 n?, sbo between groups of 8 k?
        {1, 0, 1024, 128, 0, 32},     //          ... swapped
        {1, 0, 128, 1024, 0, 64},     //          ... N = 64
    };
    int idx = -1;
    for (const Cfg& c : cfgs) {
        if (++idx != only && only >= 0) continue;
        CK(cudaMemset(dD, 0, sizeof(hD)));
        probe<<<1, 128, smem>>>(c, dD);
        cudaError_t e = cudaDeviceSynchronize();
        printf("cfg mn=%d type=%d lbo=%d sbo=%d start=%d N=%d : %s\n", c.b_mn_major, c.type, c.lbo, c.sbo, c.start, c.N, cudaGetErrorString(e));
        if (e != cudaSuccess) return 1;
        CK(cudaMemcpy(hD, dD, sizeof(hD), cudaMemcpyDeviceToHost));
        for (int k = 0; k < 8; ++k) {
            printf("  k=%d:", k);
            for (int n = 0; n < c.N; ++n) printf(" %d", (int)hD[k * 64 + n]);
            printf("\n");
        }
    }
    return 0;
}
