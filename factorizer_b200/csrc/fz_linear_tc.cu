// Weight gradient of a pointwise channel map on the tensor cores (tcgen05 / TMEM, 3xTF32):
//     dW[o][c] = sum over (batch, voxel) of dy[o][v] x[c][v],   db[o] = sum of dy[o][v]
// (reference factorizer/layers/linear.py:53-58: a k = 1 Conv1d; the wide blocks' linears, the patch convolutions on their
// space-to-depth view, the stem on its unfolded input).
//
// A contraction over voxels of two operands that lie in memory exactly as the tensor core wants a K-major operand:
// rows = channels, K = voxels, contiguous.  TMA brings a (64 channels x 32 voxels) box of dy and of x per stage straight into
// the SWIZZLE_128B layout; that fp32 tile is the "hi" operand as it stands (the tensor core reads the top 19 bits of a word).
// Four warps compute the remainder tiles lo = v - hi(v) element by element (same shared-memory offsets, so the swizzle is
// never decoded), and ONE MMA per 8 voxels with A = [dy ; dy_lo] (M = 128) and B = [x ; x_lo ; ones] (N = 144) produces all
// four hi / lo blocks of a 64 x 64 block of dW plus the bias column; they accumulate in TMEM over the CTA's share of the
// voxels and are added up, then added to the global gradient with one atomic per element and CTA.
// Warp roles: warp 0 = TMA producer, warp 1 = MMA issue (warp-uniform code under elect.sync), warps 2-5 = remainder tiles and
// epilogue.  Three stages of 32 KB, two CTAs per SM.
#include <cuda.h>

#include "fz_tc.cuh"

namespace fz {
namespace {

using namespace tc;

constexpr int kBlk = 64;                       // channels per block of dW, both ways
constexpr int kVT = 32;                        // voxels per stage = one 128-byte swizzle row
constexpr int kStages = 3;
constexpr int kThreads = 192;
constexpr uint32_t kTile = kBlk * 128;         // one (64 x 32) fp32 tile: 8 KiB
// stage: dy hi | dy lo | x hi | x lo | ones block (16 rows: the first all ones)
constexpr uint32_t oA = 0, oB = 2 * kTile, oOnes = 4 * kTile, kStage = 4 * kTile + 16 * 128;
constexpr uint32_t oBarFull = kStages * kStage, oBarLo = oBarFull + 8 * kStages, oBarEmpty = oBarLo + 8 * kStages,
                   oBarDone = oBarEmpty + 8 * kStages, oTmemSlot = oBarDone + 8, kSmemW = oTmemSlot + 8;
constexpr int kN = 2 * kBlk + 16;              // 144 accumulator columns

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}
// (batch, channels, voxels) fp32, box (1, 64, 32): rows beyond `channels` read as zeros, so do voxels beyond the end
int make_map(CUtensorMap* m, const float* ptr, long long batch, int channels, long long voxels) {
    EncodeTiledFn enc = encode_fn();
    if (!enc) return fail(FZ_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
    cuuint64_t dims[3] = {(cuuint64_t)voxels, (cuuint64_t)channels, (cuuint64_t)batch};
    cuuint64_t strides[2] = {(cuuint64_t)voxels * 4, (cuuint64_t)voxels * channels * 4};
    cuuint32_t box[3] = {kVT, kBlk, 1};
    cuuint32_t es[3] = {1, 1, 1};
    const CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(ptr), dims, strides, box, es,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(FZ_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    return FZ_OK;
}

__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 :: "r"(dst), "l"(m), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void bar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("{\n.reg .b64 st;\nmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n}" :: "r"(bar), "r"(bytes) : "memory");
}

__global__ void __launch_bounds__(kThreads, 2)
linear_wgrad_tc(const __grid_constant__ CUtensorMap map_dy, const __grid_constant__ CUtensorMap map_x, float* __restrict__ dW,
                float* __restrict__ db, int cout, int cin, int tiles_per_sample, long long total_tiles) {
    extern __shared__ __align__(1024) unsigned char smem[];
    const uint32_t sbase = smem_u32(smem);
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = (int)uniform_u32((uint32_t)(tid >> 5));
    const int ob = blockIdx.y, ib = blockIdx.z;
    const long long my_tiles = blockIdx.x < total_tiles ? (total_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

    if (tid == 0) {
        for (int s = 0; s < kStages; ++s) {
            bar_init(sbase + oBarFull + 8 * s, 1);
            bar_init(sbase + oBarLo + 8 * s, 4);
            bar_init(sbase + oBarEmpty + 8 * s, 1);
        }
        bar_init(sbase + oBarDone, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // the ones block of every stage (constant): row 0 = ones, rows 1 .. 15 = zeros (a 128-byte row of equal words: no swizzle to mind)
    for (int i = tid; i < kStages * 16 * 32; i += kThreads) {
        const int s = i / (16 * 32), r = (i / 32) % 16;
        reinterpret_cast<float*>(smem + s * kStage + oOnes)[i % (16 * 32)] = r == 0 ? 1.f : 0.f;
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" :: "r"(sbase + oTmemSlot) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = uniform_u32(*reinterpret_cast<const uint32_t*>(smem + oTmemSlot));

    if (warp == 0) {
        // ---- TMA producer ----
        if (elect_one()) {
            long long nb = (long long)blockIdx.x / tiles_per_sample;
            int nt = (int)((long long)blockIdx.x - nb * tiles_per_sample);
            for (long long it = 0; it < my_tiles; ++it) {
                const int s = (int)(it % kStages);
                const uint32_t round = (uint32_t)(it / kStages);
                if (round > 0) bar_wait(sbase + oBarEmpty + 8 * s, (round - 1) & 1);
                const uint32_t full = sbase + oBarFull + 8 * s;
                bar_expect_tx(full, 2 * kTile);
                tma_load_3d(sbase + s * kStage + oA, &map_dy, full, nt * kVT, ob * kBlk, (int)nb);
                tma_load_3d(sbase + s * kStage + oB, &map_x, full, nt * kVT, ib * kBlk, (int)nb);
                nt += (int)gridDim.x;
                while (nt >= tiles_per_sample) { nt -= tiles_per_sample; ++nb; }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ---- MMA issue: D[128 x 144] += [dy ; dy_lo] (rows) x [x ; x_lo ; ones] (columns) over the stage's 4 groups of 8 voxels ----
        const uint32_t idesc = make_idesc(128, kN, false, false);
        for (long long it = 0; it < my_tiles; ++it) {
            const int s = (int)(it % kStages);
            const uint32_t round = (uint32_t)(it / kStages);
            bar_wait(sbase + oBarLo + 8 * s, round & 1);
            tc_fence_after();
            if (elect_one()) {
                const uint64_t ad = make_desc(sbase + s * kStage + oA, 16, 1024, 2);
                const uint64_t bd = make_desc(sbase + s * kStage + oB, 16, 1024, 2);
#pragma unroll
                for (int k = 0; k < kVT / 8; ++k) mma_tf32(tmem, desc_at(ad, k * 32), desc_at(bd, k * 32), idesc, it > 0 || k > 0);
                commit(sbase + oBarEmpty + 8 * s);
                if (it + 1 == my_tiles) commit(sbase + oBarDone);
            }
            __syncwarp();
        }
    } else {
        // ---- remainder tiles: lo = v - hi(v), word by word at the same shared-memory offsets ----
        const int t = tid - 64;                                // 0 .. 127
        for (long long it = 0; it < my_tiles; ++it) {
            const int s = (int)(it % kStages);
            const uint32_t round = (uint32_t)(it / kStages);
            bar_wait(sbase + oBarFull + 8 * s, round & 1);
            unsigned char* st = smem + s * kStage;
#pragma unroll
            for (int half = 0; half < 2; ++half) {             // dy tile, then x tile
                const float4* hi = reinterpret_cast<const float4*>(st + (half ? oB : oA));
                float4* lo = reinterpret_cast<float4*>(st + (half ? oB : oA) + kTile);
#pragma unroll
                for (int i = 0; i < (int)(kTile / 16) / 128; ++i) {
                    const float4 v = hi[i * 128 + t];
                    lo[i * 128 + t] = make_float4(tf32_lo(v.x), tf32_lo(v.y), tf32_lo(v.z), tf32_lo(v.w));
                }
            }
            fence_async_smem();
            __syncwarp();
            if (lane == 0) bar_arrive(sbase + oBarLo + 8 * s);
        }
    }
    // ---- epilogue: accumulator row r (TMEM lane r) = output channel r % 64 (hi block for r < 64, lo block above); column n =
    //      input channel n % 64 (hi for n < 64, lo for 64 <= n < 128), column 128 = the bias sums ----
    if (my_tiles > 0) {
        bar_wait(sbase + oBarDone, 0);
        tc_fence_after();
    }
    __syncthreads();                                           // the stages are free: they take the scratch copy of the accumulator
    float* S = reinterpret_cast<float*>(smem);                 // [128][145]
    if (my_tiles > 0 && warp >= 2) {
        const int q = warp & 3;                                // the TMEM lane quarter this warp may read
        const int r = q * 32 + lane;
        const uint32_t row_addr = tmem + ((uint32_t)(q * 32) << 16);
#pragma unroll
        for (int c0 = 0; c0 < 128; c0 += 32) {
            float d[32];
            tmem_ld32(row_addr + c0, d);
#pragma unroll
            for (int c = 0; c < 32; ++c) S[r * 145 + c0 + c] = d[c];
        }
        uint32_t e[16];
        tmem_ld16_nowait(row_addr + 128, e);
        tmem_ld_wait();
        S[r * 145 + 128] = __uint_as_float(e[0]);
    }
    tc_fence_before();
    __syncthreads();
    if (my_tiles > 0) {
        for (int i = tid; i < kBlk * kBlk; i += kThreads) {
            const int o = i >> 6, c = i & 63;
            const int go = ob * kBlk + o, gc = ib * kBlk + c;
            if (go < cout && gc < cin) {
                const float v = (S[o * 145 + c] + S[o * 145 + 64 + c]) + (S[(64 + o) * 145 + c] + S[(64 + o) * 145 + 64 + c]);
                atomicAdd(dW + (size_t)go * cin + gc, v);
            }
        }
        if (db && ib == 0 && tid < kBlk && ob * kBlk + tid < cout) atomicAdd(db + ob * kBlk + tid, S[tid * 145 + 128] + S[(64 + tid) * 145 + 128]);
    }
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" :: "r"(tmem) : "memory");
}

}  // namespace

bool linear_wgrad_tc_supported(const float* dy, const float* x, long long batch, int cout, int cin, long long voxels) {
    return voxels % 4 == 0 && voxels < (1LL << 31) && batch < (1LL << 31) && ((reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(x)) & 15) == 0 &&
           cout > 0 && cin > 0;
}

// dW / db must be zeroed by the caller (the kernel adds its CTA totals with atomics)
int linear_wgrad_tc_launch(const float* dy, const float* x, float* dW, float* db, long long batch, int cout, int cin, long long voxels,
                           cudaStream_t st) {
    CUtensorMap map_dy, map_x;
    if (int e = make_map(&map_dy, dy, batch, cout, voxels)) return e;
    if (int e = make_map(&map_x, x, batch, cin, voxels)) return e;
    static SmemConfig cfg;
    FZ_CUDA_CHECK(cfg.ensure(linear_wgrad_tc, kSmemW));
    const int tps = (int)((voxels + kVT - 1) / kVT);
    const long long tiles = batch * tps;
    const int bo = (cout + kBlk - 1) / kBlk, bi = (cin + kBlk - 1) / kBlk;
    long long gx = (2LL * num_sms() + bo * bi - 1) / (bo * bi);
    // The tensor core adds into its fp32 accumulator with truncation, so the error of a TMEM accumulation chain grows linearly
    // with its length (measured 2.6e-5 of max |dW| after 1 770 MMAs, 6e-7 for the FP32-pipe kernel): at most kMaxChain stages
    // (4 MMAs each) per CTA, more CTAs along the voxel axis instead
    constexpr long long kMaxChain = 64;
    if (gx * kMaxChain < tiles) gx = (tiles + kMaxChain - 1) / kMaxChain;
    if (gx > tiles) gx = tiles;
    if (gx < 1) gx = 1;
    linear_wgrad_tc<<<dim3((unsigned)gx, bo, bi), kThreads, kSmemW, st>>>(map_dy, map_x, dW, db, cout, cin, tps, tiles);
    FZ_LAUNCH_CHECK();
    return FZ_OK;
}

}  // namespace fz
