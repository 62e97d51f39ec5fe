// Weight gradient of the pointwise (k = 1) channel map y = W x + b on channels-first activations
// (factorizer/layers/linear.py:53-58 runs it as a Conv1d; its weight gradient is the contraction over voxels):
//     dW[o][i] = sum_{b,v} dy[b][o][v] x[b][i][v],     db[o] = sum_{b,v} dy[b][o][v]
// The Swin Factorizer's wider stages give this a 64..1024 x 64..512 result over 512..262144 voxels, a shape library
// SGEMMs handle badly (sgemm_largek: 340 us for 64 x 64 x 262144, 10 % of the FP32 pipe).  Here a CTA owns one
// 32 x 32 block of dW and a share of the voxel tiles, stages 256 voxels of its 32 + 32 rows in shared memory, keeps
// the 32 x 32 partial sums in registers (a warp takes every 4th float4 column, a lane an 8 x 4 sub-block read with
// broadcast LDS.128) and adds them to dW once at the end.  FP32 pipe, exact fp32 products.
#include "fz_common.cuh"
#include "fz_internal.cuh"

namespace fz {
namespace {

constexpr int kWT = 128;            // threads
constexpr int kWV = 256;            // voxels per tile
constexpr int kWRS = kWV + 4;       // staged row stride (floats), = 4 mod 32: conflict-free row-broadcast reads
constexpr int kWWarps = kWT / 32;

__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
    return __ffma2_rn(a, b, c);
}

// rows [r0, r0+32) x voxels [v0, v0+256) of a (rows_total, vox) matrix -> S[32][kWRS]; past-the-end voxels and rows
// (rows_left = rows_total - r0 may be below 32) read as 0
__device__ __forceinline__ void stage_rows(const float* __restrict__ g, long long vox, long long v0, int rows_left, float* __restrict__ S,
                                           int tid, float (&rowsum)[16], bool want_sums) {
    const int col4 = tid & 63, rbase = tid >> 6;
    const long long v = v0 + 4 * col4;
    const bool in_v = v < vox;                    // vox % 4 == 0: a float4 is wholly inside or outside
#pragma unroll
    for (int h = 0; h < 2; ++h) {                 // 8 rows in flight at a time (register budget of 3 CTAs per SM)
        float4 val[8];
#pragma unroll
        for (int i = 0; i < 8; ++i)
            val[i] = (in_v && rbase + 2 * (8 * h + i) < rows_left)
                         ? __ldg(reinterpret_cast<const float4*>(g + (long long)(rbase + 2 * (8 * h + i)) * vox + v))
                         : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            *reinterpret_cast<float4*>(S + (rbase + 2 * (8 * h + i)) * kWRS + 4 * col4) = val[i];
            if (want_sums) rowsum[8 * h + i] += (val[i].x + val[i].y) + (val[i].z + val[i].w);
        }
    }
}

__global__ void __launch_bounds__(kWT, 3) linear_wgrad(const float* __restrict__ dy, const float* __restrict__ x, float* __restrict__ dW,
                                                       float* __restrict__ db, int cout, int cin, long long vox, int tiles_per_sample,
                                                       long long total_tiles) {
    extern __shared__ __align__(16) float sm[];
    float* SA = sm;                     // dy rows [32][kWRS]
    float* SB = sm + 32 * kWRS;         // x rows  [32][kWRS]
    const int tid = threadIdx.x, lane = tid & 31, grp = tid >> 5, ro = lane & 3, co = lane >> 2;
    const int o0 = blockIdx.y * 32, i0 = blockIdx.z * 32;
    const bool sums = db != nullptr && blockIdx.z == 0;
    float2 acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int k = 0; k < 4; ++k) acc[i][k] = make_float2(0.f, 0.f);
    float rowsum[16], unused[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) rowsum[i] = 0.f;
    for (long long tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const long long b = tile / tiles_per_sample;
        const long long v0 = (tile - b * tiles_per_sample) * kWV;
        __syncthreads();                // the previous tile has been consumed
        stage_rows(dy + (b * cout + o0) * vox, vox, v0, cout - o0, SA, tid, rowsum, sums);
        stage_rows(x + (b * cin + i0) * vox, vox, v0, cin - i0, SB, tid, unused, false);
        __syncthreads();
#pragma unroll 2
        for (int it = 0; it < kWV / 4 / kWWarps; ++it) {
            const int v = (it * kWWarps + grp) * 4;
            float4 B[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) B[k] = *reinterpret_cast<const float4*>(SB + (co + 8 * k) * kWRS + v);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float4 A = *reinterpret_cast<const float4*>(SA + (ro + 4 * i) * kWRS + v);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    acc[i][k] = ffma2(make_float2(A.x, A.y), make_float2(B[k].x, B[k].y), acc[i][k]);
                    acc[i][k] = ffma2(make_float2(A.z, A.w), make_float2(B[k].z, B[k].w), acc[i][k]);
                }
            }
        }
    }
    __syncthreads();
    // per-warp partial blocks -> shared memory (the staging area is free now) -> one atomicAdd per element and CTA
    float* scr = sm;                    // [warp][32*32]
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int k = 0; k < 4; ++k) scr[grp * 1024 + (ro + 4 * i) * 32 + co + 8 * k] = acc[i][k].x + acc[i][k].y;
    float* rs = sm + kWWarps * 1024;    // [32] row sums of dy
    if (tid < 32) rs[tid] = 0.f;
    __syncthreads();
    for (int e = tid; e < 1024; e += kWT) {
        float t = 0.f;
#pragma unroll
        for (int w = 0; w < kWWarps; ++w) t += scr[w * 1024 + e];
        if (o0 + (e >> 5) < cout && i0 + (e & 31) < cin) atomicAdd(dW + (long long)(o0 + (e >> 5)) * cin + i0 + (e & 31), t);
    }
    if (sums) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const float t = warp_sum(rowsum[i]);
            if (lane == 0) atomicAdd(rs + (tid >> 6) + 2 * i, t);      // once per kernel: the CAS loop does not matter here
        }
        __syncthreads();
        if (tid < 32 && o0 + tid < cout) atomicAdd(db + o0 + tid, rs[tid]);
    }
}

}  // namespace
}  // namespace fz

using namespace fz;

extern "C" {

int fz_linear_wgrad_supported(int32_t cout, int32_t cin, int64_t voxels) {
    return cout > 0 && cin > 0 && cout <= 32 * 65535 && cin <= 32 * 65535 && voxels > 0 && voxels % 4 == 0;
}

int fz_linear_wgrad(const float* dy, const float* x, float* dW, float* db, int64_t batch, int32_t cout, int32_t cin, int64_t voxels,
                    void* stream) {
    if (batch < 0 || cout <= 0 || cin <= 0 || voxels <= 0) return fail(FZ_ERR_INVALID, "linear wgrad: bad sizes");
    if (!fz_linear_wgrad_supported(cout, cin, voxels))
        return fail(FZ_ERR_UNSUPPORTED, "linear wgrad kernel needs a voxel count divisible by 4 (got %d x %d x %lld)", cout, cin,
                    (long long)voxels);
    if (!dW) return fail(FZ_ERR_INVALID, "linear wgrad: null dW");
    cudaStream_t st = (cudaStream_t)stream;
    FZ_CUDA_CHECK(cudaMemsetAsync(dW, 0, sizeof(float) * (size_t)cout * cin, st));
    if (db) FZ_CUDA_CHECK(cudaMemsetAsync(db, 0, sizeof(float) * (size_t)cout, st));
    if (batch == 0) return FZ_OK;
    if (!dy || !x) return fail(FZ_ERR_INVALID, "linear wgrad: null buffer");
    const int tps = (int)((voxels + kWV - 1) / kWV);
    const long long tiles = batch * tps;
    const int bo = (cout + 31) / 32, bi = (cin + 31) / 32;       // channel counts are padded with zero rows inside the kernel
    const int blocks = bo * bi;
    int dev = 0, sms = 148;
    FZ_CUDA_CHECK(cudaGetDevice(&dev));
    FZ_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    long long gx = (3LL * sms + blocks - 1) / blocks;
    if (gx > tiles) gx = tiles;
    if (gx < 1) gx = 1;
    const size_t smem = sizeof(float) * 2 * 32 * kWRS;          // 66.5 KB >= the 16.1 KB epilogue scratch
    static SmemConfig cfg;
    FZ_CUDA_CHECK(cfg.ensure(linear_wgrad, smem));
    linear_wgrad<<<dim3((unsigned)gx, bo, bi), kWT, smem, st>>>(dy, x, dW, db, cout, cin, voxels, tps, tiles);
    FZ_LAUNCH_CHECK();
    return FZ_OK;
}

}  // extern "C"
