// Microbenchmark: per-SM rate of TMA tile LOADS and STORES of one 8-channel 8x8x(8*wpt) window box
// (32 / 64 / 128-byte rows) when the data streams from HBM vs. when it is L2-resident, as a function of
// the number of tiles in flight per SM.  Question answered: is the fused kernels' tile traffic
// (~15 B/clk/SM, mostly L2 hits) limited by the TMA / L2 path for 32-byte rows?  Not part of the product.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)
typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeFn get_encode() {
    void* fn = nullptr; cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
    return (EncodeFn)fn;
}
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, int n) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(n)); }
__device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ bool mbar_try(uint64_t* b, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(smem_u32(b)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) { while (!mbar_try(b, parity)) {} }
__device__ __forceinline__ void tma_load5(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2, int c3, int c4) {
    asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(m), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
}
__device__ __forceinline__ void tma_store5(const CUtensorMap* m, const void* src, int c0, int c1, int c2, int c3, int c4) {
    asm volatile("cp.async.bulk.tensor.5d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4, %5, %6}], [%1];"
                 ::"l"(m), "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// One thread per "stream": each issues loads (or stores) of its tiles back to back with `depth` = 2 in
// flight.  tiles are numbered over a region of d_tiles x 16 x (16/wpt) x chan_groups boxes, visited
// `reps` times (reps > 1 with a small region = L2-resident).
__global__ void probe(const __grid_constant__ CUtensorMap map, int tiles_w, int tiles_h, int tiles_d, int chan_groups,
                      int tile_bytes, int wpt, int store, int reps, int streams_total) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint64_t bars[64];
    const int s = threadIdx.x;            // stream within the CTA (blockDim.x streams, one thread each)
    unsigned char* buf = smem + (size_t)s * 2 * tile_bytes;
    uint64_t* bar = bars + 2 * s;
    mbar_init(&bar[0], 1); mbar_init(&bar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
    const int total = tiles_w * tiles_h * tiles_d * chan_groups;
    const int gs = blockIdx.x * blockDim.x + s;
    int phase[2] = {0, 0};
    for (int r = 0; r < reps; ++r) {
        int stage = 0, issued = 0;
        for (int t = gs; t < total; t += streams_total) {
            int q = t;
            const int c0 = (q % tiles_w) * 8 * wpt; q /= tiles_w;
            const int c1 = (q % tiles_h) * 8; q /= tiles_h;
            const int c2 = (q % tiles_d) * 8; q /= tiles_d;
            const int c3 = q * 8;
            if (store) {
                if (issued >= 2) bulk_wait_read1();
                tma_store5(&map, buf + stage * tile_bytes, c0, c1, c2, c3, 0);
                bulk_commit();
            } else {
                if (issued >= 2) { mbar_wait(&bar[stage], phase[stage]); phase[stage] ^= 1; }
                mbar_expect(&bar[stage], tile_bytes);
                tma_load5(buf + stage * tile_bytes, &map, &bar[stage], c0, c1, c2, c3, 0);
            }
            ++issued; stage ^= 1;
        }
        if (store) bulk_wait0();
        else {
            // drain: wait for the (up to two) loads still in flight
            const int pending = issued < 2 ? issued : 2;
            for (int k = 0; k < pending; ++k) {
                const int st = (stage + (2 - pending) + k) & 1;
                mbar_wait(&bar[st], phase[st]); phase[st] ^= 1;
            }
        }
    }
}

int main() {
    const int C = 32, D = 128, H = 128, W = 128;
    const size_t n = (size_t)C * D * H * W;
    float* x;
    CK(cudaMalloc(&x, n * 4));
    CK(cudaMemset(x, 0, n * 4));
    EncodeFn enc = get_encode();
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    int dev = 0, khz = 0; CK(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev));
    printf("clock %d MHz (attr)\n", khz / 1000);
    for (int wpt : {1, 2, 4}) {
        CUtensorMap map;
        cuuint64_t dims[5] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)C, 1};
        cuuint64_t strides[4] = {(cuuint64_t)W * 4, (cuuint64_t)W * H * 4, (cuuint64_t)W * H * D * 4, (cuuint64_t)W * H * D * C * 4};
        cuuint32_t box[5] = {(cuuint32_t)(8 * wpt), 8, 8, 8, 1};
        cuuint32_t es[5] = {1, 1, 1, 1, 1};
        if (enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, x, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE)) { printf("encode failed\n"); return 1; }
        const int tile_bytes = 16384 * wpt;
        for (int store : {0, 1}) {
            for (int resident : {0, 1}) {
                for (int streams : {2, 4, 6}) {
                    size_t smem = (size_t)streams * 2 * tile_bytes;
                    if (smem > 200 * 1024) continue;
                    CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                    // region: full volume once (HBM) or 1 channel group x 32 planes = 16.8 MB, 16 times (L2)
                    const int td = resident ? 4 : D / 8, cg = resident ? 1 : C / 8, reps = resident ? 16 : 1;
                    float best = 1e9, ms;
                    for (int it = 0; it < 4; ++it) {
                        CK(cudaEventRecord(e0));
                        probe<<<148, streams, smem>>>(map, W / (8 * wpt), H / 8, td, cg, tile_bytes, wpt, store, reps, 148 * streams);
                        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
                        CK(cudaGetLastError());
                        CK(cudaEventElapsedTime(&ms, e0, e1));
                        if (ms < best) best = ms;
                    }
                    const double bytes = (double)(W / (8 * wpt)) * (H / 8) * td * cg * tile_bytes * reps;
                    printf("rows=%3dB %s %s tiles-in-flight/SM=%2d: %8.1f us  %6.0f GB/s  %5.1f B/clk/SM @1.9GHz\n", 32 * wpt, store ? "store" : "load ",
                           resident ? "L2 " : "HBM", streams * 2, best * 1e3, bytes / (best * 1e-3) / 1e9, bytes / (best * 1e-3) / 1.9e9 / 148);
                }
            }
        }
    }
    return 0;
}
