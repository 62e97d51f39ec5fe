// How long does a chain of small kind::tf32 MMAs take?  One CTA, one issuing thread; 96 MMAs per case, clock64 from the
// first issue to the completion of the commit.  Operand contents are irrelevant (zeros).  Not part of the product.
//   case: A source (tmem / smem K-major no swizzle / smem MN-major voxel rows), N, number of accumulators rotated through
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t lt) {
    return (uint64_t)((saddr >> 4) & 0x3fff) | (uint64_t)((lbo >> 4) & 0x3fff) << 16 | (uint64_t)((sbo >> 4) & 0x3fff) << 32 | (uint64_t)1 << 46 | (uint64_t)(lt & 7) << 61;
}
__device__ __forceinline__ uint32_t idesc(int m, int n, int a_mn, int b_mn) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

// ASRC: 0 tmem, 1 smem K-major no swizzle, 2 smem MN-major voxel rows (B MN-major too).  BSW: B K-major with SWIZZLE_128B.
template <int ASRC, int N, int NACC, int BSW>
__global__ void __launch_bounds__(128) probe(long long* out) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    for (int i = tid; i < 160 * 1024 / 4; i += 128) reinterpret_cast<float*>(smem)[i] = 0.f;
    if (tid == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&bar))); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" :: "r"(smem_u32(&tmem_base)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = __shfl_sync(0xffffffffu, tmem_base, 0);
    if (warp == 0) {
        const uint32_t id = idesc(128, N, ASRC == 2, ASRC == 2);
        const uint32_t sb = smem_u32(smem);
        const uint64_t bd = ASRC == 2 ? make_desc(sb + 65536, 12288, 512, 1) : BSW ? make_desc(sb + 65536, 16, 1024, 2) : make_desc(sb + 65536, 128, 1024, 0);
        const uint64_t ad = ASRC == 2 ? make_desc(sb, 12288, 512, 1) : make_desc(sb, 128, 256, 0);
        uint32_t el;
        asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(el));
        const long long t0 = clock64();
        if (el) {
#pragma unroll
            for (int i = 0; i < 48; ++i) {
                const uint32_t d = tmem + 128 + (uint32_t)(((i % NACC) * N) % 384);
                const uint64_t a2 = ad + (uint64_t)((i % 4) * (ASRC == 2 ? 64 : 16));
                const uint64_t b2 = bd + (uint64_t)((i % 4) * (ASRC == 2 ? 64 : BSW ? 2 : 16));
                if (ASRC == 0)
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n"
                                 :: "r"(d), "r"(tmem + (uint32_t)((i % 16) * 8)), "l"(b2), "r"(id), "r"((uint32_t)(i >= NACC)) : "memory");
                else
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
                                 :: "r"(d), "l"(a2), "l"(b2), "r"(id), "r"((uint32_t)(i >= NACC)) : "memory");
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(&bar)) : "memory");
        }
        __syncwarp();
        const long long t1 = clock64();
        uint32_t ok = 0;
        while (!ok) asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
        const long long t2 = clock64();
        if (tid == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" :: "r"(tmem) : "memory");
}

template <int ASRC, int N, int NACC, int BSW>
void run(long long* d) {
    const int smem = 160 * 1024 + 1024;
    CK(cudaFuncSetAttribute(probe<ASRC, N, NACC, BSW>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    long long h[2];
    for (int rep = 0; rep < 2; ++rep) {
        probe<ASRC, N, NACC, BSW><<<1, 128, smem>>>(d);
        CK(cudaDeviceSynchronize());
    }
    CK(cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost));
    printf("A=%s N=%3d accumulators=%d Bswz=%d : issue %5lld cycles, done %5lld cycles  (%.1f / MMA; floor %d)\n",
           ASRC == 0 ? "tmem" : ASRC == 1 ? "smemK" : "smemMN", N, NACC, BSW, h[0], h[1], h[1] / 48.0, N / 2);
}

int main() {
    long long* d;
    CK(cudaMalloc(&d, 16));
    run<0, 64, 1, 0>(d); run<0, 64, 2, 0>(d); run<0, 64, 4, 0>(d); run<0, 32, 1, 0>(d); run<0, 32, 4, 0>(d); run<0, 128, 1, 0>(d); run<0, 256, 1, 0>(d);
    run<0, 64, 1, 1>(d); run<0, 32, 1, 1>(d); run<0, 128, 1, 1>(d);
    run<1, 64, 1, 0>(d); run<1, 64, 4, 0>(d); run<1, 32, 1, 0>(d); run<1, 128, 1, 0>(d); run<1, 256, 1, 0>(d); run<1, 64, 1, 1>(d);
    run<2, 64, 1, 0>(d); run<2, 64, 2, 0>(d); run<2, 32, 1, 0>(d); run<2, 128, 1, 0>(d);
    return 0;
}
