// Microbenchmark: what does a TMA-tiled window copy sustain on B200 as a function of the box width
// (windows per tile), warps per SM, and store flavour (plain / reduce-add)?  Also checks that
// out-of-bounds (negative / past-the-end) box coordinates clip on store.  Not part of the product.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeFn get_encode() {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
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
__device__ __forceinline__ void tma_red5(const CUtensorMap* m, const void* src, int c0, int c1, int c2, int c3, int c4) {
    asm volatile("cp.reduce.async.bulk.tensor.5d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3, %4, %5, %6}], [%1];"
                 ::"l"(m), "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
template <int N> __device__ __forceinline__ void bulk_wait() { asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory"); }

// Each warp: private 2-stage ring of tiles; copy tiles in -> out through smem with TMA only.
// mode 0: store, 1: reduce-add.  shift: box start offset (tests unaligned element coords + clipping).
__global__ void probe(const __grid_constant__ CUtensorMap min, const __grid_constant__ CUtensorMap mout,
                      int tiles_w, int tiles_h, int tiles_d, int chan_groups, int tile_bytes, int wpt, int mode, int shift, int shift_out,
                      int* counter) {
    extern __shared__ __align__(1024) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned char* buf = smem + (size_t)warp * 2 * tile_bytes;
    __shared__ uint64_t bars[32 * 2];
    uint64_t* bar = bars + warp * 2;
    if (lane == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    __syncwarp();
    const int total = tiles_w * tiles_h * tiles_d * chan_groups;
    int phase[2] = {0, 0};
    int cur = -1, nxt = -1, stage = 0;
    auto coords_s = [&](int t, int sh, int& c0, int& c1, int& c2, int& c3) {
        c0 = (t % tiles_w) * 8 * wpt - sh; t /= tiles_w;
        c1 = (t % tiles_h) * 8 - sh; t /= tiles_h;
        c2 = (t % tiles_d) * 8 - sh; t /= tiles_d;
        c3 = t * 8;
    };
    auto coords = [&](int t, int& c0, int& c1, int& c2, int& c3) { coords_s(t, shift, c0, c1, c2, c3); };
    if (lane == 0) {
        cur = atomicAdd(counter, 1);
        if (cur < total) { int c0, c1, c2, c3; coords(cur, c0, c1, c2, c3); mbar_expect(&bar[0], tile_bytes); tma_load5(buf, &min, &bar[0], c0, c1, c2, c3, 0); }
    }
    cur = __shfl_sync(0xffffffffu, cur, 0);
    while (cur < total) {
        if (lane == 0) {
            nxt = atomicAdd(counter, 1);
            bulk_wait_read<0>();  // previous store from the other stage has finished reading smem
            if (nxt < total) { int c0, c1, c2, c3; coords(nxt, c0, c1, c2, c3); mbar_expect(&bar[stage ^ 1], tile_bytes); tma_load5(buf + (stage ^ 1) * tile_bytes, &min, &bar[stage ^ 1], c0, c1, c2, c3, 0); }
        }
        nxt = __shfl_sync(0xffffffffu, nxt, 0);
        mbar_wait(&bar[stage], phase[stage]);
        phase[stage] ^= 1;
        if (lane == 0) {
            int c0, c1, c2, c3; coords_s(cur, shift_out, c0, c1, c2, c3);
            if (mode == 0) tma_store5(&mout, buf + stage * tile_bytes, c0, c1, c2, c3, 0);
            else tma_red5(&mout, buf + stage * tile_bytes, c0, c1, c2, c3, 0);
            bulk_commit();
        }
        __syncwarp();
        cur = nxt; stage ^= 1;
    }
    if (lane == 0) bulk_wait<0>();
}

int main(int argc, char** argv) {
    const int C = 32, D = 128, H = 128, W = 128;
    const size_t n = (size_t)C * D * H * W;
    float *x, *y; int* counter;
    CK(cudaMalloc(&x, n * 4)); CK(cudaMalloc(&y, n * 4)); CK(cudaMalloc(&counter, 4));
    std::vector<float> hx(n);
    for (size_t i = 0; i < n; ++i) hx[i] = (float)(i % 1000003) * 1e-3f;
    CK(cudaMemcpy(x, hx.data(), n * 4, cudaMemcpyHostToDevice));
    EncodeFn enc = get_encode();
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    // reference: plain D2D memcpy
    for (int i = 0; i < 3; ++i) CK(cudaMemcpyAsync(y, x, n * 4, cudaMemcpyDeviceToDevice));
    CK(cudaEventRecord(e0));
    for (int i = 0; i < 10; ++i) CK(cudaMemcpyAsync(y, x, n * 4, cudaMemcpyDeviceToDevice));
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    printf("memcpy D2D: %.1f us  %.0f GB/s (r+w)\n", ms * 100, 2.0 * n * 4 / (ms / 10 * 1e-3) / 1e9);

    if (argc > 1) {
        // clipping tests, one per process: argv[1] = 0 load-OOB, 1 store-OOB, 2 reduce-OOB; argv[2] = wpt
        int test = atoi(argv[1]); int wpt = argc > 2 ? atoi(argv[2]) : 1;
        CUtensorMap min, mout;
        cuuint64_t dims[5] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)C, 1};
        cuuint64_t strides[4] = {(cuuint64_t)W * 4, (cuuint64_t)W * H * 4, (cuuint64_t)W * H * D * 4, (cuuint64_t)W * H * D * C * 4};
        cuuint32_t box[5] = {(cuuint32_t)(8 * wpt), 8, 8, 8, 1};
        cuuint32_t es[5] = {1, 1, 1, 1, 1};
        CUresult r1 = enc(&min, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, x, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        CUresult r2 = enc(&mout, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, y, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        printf("encode %d %d\n", r1, r2);
        const int tile_bytes = 16384 * wpt;
        CK(cudaMemset(y, 0, n * 4)); CK(cudaMemset(counter, 0, 4));
        size_t smem = (size_t)2 * 2 * tile_bytes;
        CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int sin = test == 0 ? 4 : (test == 3 ? 4 : 0), sout = test == 0 ? 0 : 4, mode = test >= 2 ? 1 : 0;
        int extra = test == 3 ? 1 : 0;
        probe<<<148, 64, smem>>>(min, mout, W / (8 * wpt) + extra, H / 8 + extra, D / 8 + extra, C / 8, tile_bytes, wpt, mode, sin, sout, counter);
        cudaError_t e = cudaDeviceSynchronize();
        printf("test %d wpt %d: %s\n", test, wpt, cudaGetErrorString(e));
        if (e == cudaSuccess) {
            std::vector<float> hy(n);
            CK(cudaMemcpy(hy.data(), y, n * 4, cudaMemcpyDeviceToHost));
            // expected: test 0: y[i] = x[i-4 per axis] (zero where OOB) ; test 1/2: y[i-4] = x[i]; test 3: y = x
            size_t bad = 0;
            for (int c = 0; c < C; ++c) for (int d = 0; d < D; ++d) for (int h = 0; h < H; ++h) for (int w = 0; w < W; ++w) {
                size_t i = (((size_t)c * D + d) * H + h) * W + w;
                float want;
                if (test == 3) want = hx[i];
                else if (test == 0) { int dd = d - 4, hh = h - 4, ww = w - 4; want = (dd < 0 || hh < 0 || ww < 0) ? 0.f : hx[(((size_t)c * D + dd) * H + hh) * W + ww]; }
                else { int dd = d + 4, hh = h + 4, ww = w + 4; want = (dd >= D || hh >= H || ww >= W) ? 0.f : hx[(((size_t)c * D + dd) * H + hh) * W + ww]; }
                if (hy[i] != want) ++bad;
            }
            printf("test %d wpt %d: %zu mismatches of %zu\n", test, wpt, bad, n);
        }
        return 0;
    }
    for (int wpt : {1, 2, 4}) {
        for (int promo : {0, 2}) {  // none / 128B
            CUtensorMap min, mout;
            cuuint64_t dims[5] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)C, 1};
            cuuint64_t strides[4] = {(cuuint64_t)W * 4, (cuuint64_t)W * H * 4, (cuuint64_t)W * H * D * 4, (cuuint64_t)W * H * D * C * 4};
            cuuint32_t box[5] = {(cuuint32_t)(8 * wpt), 8, 8, 8, 1};
            cuuint32_t es[5] = {1, 1, 1, 1, 1};
            CUtensorMapL2promotion pr = promo == 0 ? CU_TENSOR_MAP_L2_PROMOTION_NONE : CU_TENSOR_MAP_L2_PROMOTION_L2_128B;
            CUresult r1 = enc(&min, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, x, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, pr, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            CUresult r2 = enc(&mout, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, y, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, pr, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r1 || r2) { printf("encode failed %d %d\n", r1, r2); return 1; }
            const int tile_bytes = 16384 * wpt;
            for (int mode : {0, 1}) {
                for (int warps_per_sm : {2, 4, 6}) {
                    int nw = warps_per_sm;  // one CTA per SM with nw warps
                    size_t smem = (size_t)nw * 2 * tile_bytes;
                    if (smem > 227 * 1024) continue;
                    CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                    float best = 1e9;
                    for (int it = 0; it < 5; ++it) {
                        CK(cudaMemsetAsync(counter, 0, 4));
                        if (mode == 1) CK(cudaMemsetAsync(y, 0, n * 4));
                        CK(cudaEventRecord(e0));
                        probe<<<148, nw * 32, smem>>>(min, mout, W / (8 * wpt), H / 8, D / 8, C / 8, tile_bytes, wpt, mode, 0, 0, counter);
                        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
                        CK(cudaGetLastError());
                        CK(cudaEventElapsedTime(&ms, e0, e1));
                        if (ms < best) best = ms;
                    }
                    printf("wpt=%d promo=%d mode=%s warps/SM=%d: %.1f us  %.0f GB/s (r+w)\n", wpt, promo, mode ? "red.add" : "store", nw, best * 1e3, 2.0 * n * 4 / (best * 1e-3) / 1e9);
                }
            }
        }
    }
    return 0;
}
