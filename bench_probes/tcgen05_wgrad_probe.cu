// Probe for a tensor-core weight gradient: D[64 x 32] = sum over K = 128 voxels of A[r][v] B[c][v], A rows 32..63 zero
// (a 32 x 32 gradient padded to the minimum M = 64), both operands K-major = voxel-contiguous rows, no swizzle, core
// matrices 144 B apart along K (padding that makes a voxel-owning warp's stores conflict-free).  Dumps where the 64 rows
// land in TMEM.  Not part of the product.
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)
constexpr int R = 32, NC = 32, KV = 128, LBO = 144, SBO = (KV / 4) * LBO;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t lt) {
    return (uint64_t)((saddr >> 4) & 0x3fff) | (uint64_t)((lbo >> 4) & 0x3fff) << 16 | (uint64_t)((sbo >> 4) & 0x3fff) << 32 | (uint64_t)1 << 46 | (uint64_t)(lt & 7) << 61;
}
__device__ __forceinline__ uint32_t k_off(int row, int v) { return (uint32_t)((row >> 3) * SBO + (v >> 2) * LBO + (row & 7) * 16 + (v & 3) * 4); }

__global__ void __launch_bounds__(128) probe(const float* __restrict__ A /*[R][KV]*/, const float* __restrict__ B /*[NC][KV]*/, float* __restrict__ D /*[128][32]*/) {
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* a = smem;                       // 64 rows
    unsigned char* b = smem + 8 * SBO;             // 32 rows
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int r = 0; r < 64; ++r) *reinterpret_cast<float*>(a + k_off(r, tid)) = r < R ? A[r * KV + tid] : 0.f;
    for (int c = 0; c < NC; ++c) *reinterpret_cast<float*>(b + k_off(c, tid)) = B[c * KV + tid];
    if (tid == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&bar))); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 32;" :: "r"(smem_u32(&tmem_base)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base;
    // D = F32, A = B = TF32, both K-major, N = 32, M = 64
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(NC >> 3) << 17) | ((uint32_t)(64 >> 4) << 24);
    if (tid == 0) {
        for (int s = 0; s < KV / 8; ++s) {
            const uint64_t ad = make_desc(smem_u32(a) + s * 2 * LBO, LBO, SBO, 0);
            const uint64_t bd = make_desc(smem_u32(b) + s * 2 * LBO, LBO, SBO, 0);
            const uint32_t acc = s > 0;
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
                         :: "r"(tmem), "l"(ad), "l"(bd), "r"(idesc), "r"(acc) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(&bar)) : "memory");
    }
    uint32_t ok = 0;
    while (!ok) asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t r[32];
    const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
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
    for (int n = 0; n < 32; ++n) D[tid * 32 + n] = __uint_as_float(r[n]);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 32;" :: "r"(tmem) : "memory");
}

int main() {
    static float hA[R * KV], hB[NC * KV], hD[128 * 32];
    srand(2);
    for (auto& v : hA) v = (float)rand() / RAND_MAX * 2.f - 1.f;
    for (auto& v : hB) v = (float)rand() / RAND_MAX * 2.f - 1.f;
    float *dA, *dB, *dD;
    CK(cudaMalloc(&dA, sizeof(hA))); CK(cudaMalloc(&dB, sizeof(hB))); CK(cudaMalloc(&dD, sizeof(hD)));
    CK(cudaMemcpy(dA, hA, sizeof(hA), cudaMemcpyHostToDevice)); CK(cudaMemcpy(dB, hB, sizeof(hB), cudaMemcpyHostToDevice));
    CK(cudaMemset(dD, 0, sizeof(hD)));
    const int smem = 12 * SBO + 1024;
    CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    probe<<<1, 128, smem>>>(dA, dB, dD);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(hD, dD, sizeof(hD), cudaMemcpyDeviceToHost));
    // which TMEM lane holds which row?
    int found = 0;
    for (int r = 0; r < R; ++r) {
        double ref[32];
        for (int c = 0; c < NC; ++c) { ref[c] = 0; for (int v = 0; v < KV; ++v) ref[c] += (double)hA[r * KV + v] * hB[c * KV + v]; }
        int best = -1; double beste = 1e9;
        for (int lane = 0; lane < 128; ++lane) {
            double e = 0; for (int c = 0; c < NC; ++c) e = fmax(e, fabs(hD[lane * 32 + c] - ref[c]));
            if (e < beste) { beste = e; best = lane; }
        }
        if (r < 4 || r == 15 || r == 16 || r == 31) printf("row %2d -> lane %3d (max err %.3g)\n", r, best, beste);
        found += beste < 0.05;
    }
    int nzl = 0; for (int lane = 0; lane < 128; ++lane) { double s = 0; for (int c = 0; c < 32; ++c) s += fabs(hD[lane * 32 + c]); if (s > 1e-6) { if (nzl < 40) printf("%d ", lane); ++nzl; } }
    printf("\nnonzero lanes %d; rows matched (TF32 tolerance) %d of %d\n", nzl, found, R);
    return 0;
}
