// Specialised fused kernels for the production geometry of every Swin-Factorizer config:
// head_dim 8, patch 8x8x8 (an 8 x 512 matrix per window), rank 1, HALS, any number of window sets
// with any integer shifts.  One kernel per direction does
//     gather (roll + window partition + ReLU)  ->  T unrolled HALS sweeps  ->  U V^T  ->  scatter-mean
// (reference factorizer/factorizer.py:41-50, operations.py:266-280/417-434,
//  matrix_factorization.py:210-229/514-533), reading X once and writing Y once.
//
// Design (B200, sm_100a):
//  * work unit = one window = one 16 KiB tile; a PAIR of warps (64 lanes) owns it: lane p holds
//    columns 4p..4p+3 and 256+4p..+3 of all 8 rows in registers, so X^T u, the V update and the rank-1
//    reconstruction are lane-local and only the 8 row sums (X v) need a 64-lane reduction
//    (recursive-halving shuffles + one named barrier);
//  * X (and dY) tiles arrive by TMA: a 5-D tensor map over (W,H,D,C,B) with box (8,8,8,8,1) lands a
//    window in shared memory in exactly the matrix's column order, prefetched one window ahead.
//    Windows that wrap around the volume (the roll) or start at a W offset that is not 16-byte
//    aligned cannot be one TMA box; they use per-lane global accesses with modular addressing;
//  * the mean over window sets never round-trips a full tile through memory in the forward: the
//    early sets write only their rank-1 factors (u: 8, v: 512 floats = 13 % of a tile) and the last
//    set -- chosen to be the unshifted one, so all of its windows are clean boxes -- adds the factor
//    products of the (up to 8 per set) windows it overlaps and writes Y exactly once.  TMA
//    reduce-add was measured at ~1.6 TB/s on B200 and is not used;
//  * the backward's dX is a dense sum, so set j reads the partial sum left by set j-1 (L2-resident),
//    adds its own contribution and stores it back (plain loads / stores, no atomics);
//  * persistent CTAs pull windows from an atomic counter (claimed one window ahead, so its latency
//    is hidden) in a host-built order that walks the volume row of windows by row of windows, every
//    dependent row a few rows behind the rows it needs: second touches of X / dY / partial dX hit L2
//    and HBM sees each tensor once.  Per-(set,row) completion counters guard the dependencies; they
//    are deadlock-free because a window only ever waits for windows claimed earlier, and the head
//    of every dependency chain never waits;
//  * forward saves u_t (8 floats) and b_t per sweep and window (1.2 % extra traffic); backward
//    recomputes every v_t from them with one lane-local GEMV instead of re-running the solver, keeps
//    X and the dX accumulator in registers, and streams dY through shared memory once.
#include <cuda.h>
#include <stdlib.h>
#include <string.h>

#include "fz_internal.cuh"

namespace fz {

constexpr int kMaxOrder = 2048;       // rows of windows (all sets) the ordering table can hold
constexpr int kTileBytes = 16384;     // 8 x 512 fp32
constexpr int kMaxT = 8;
constexpr int kFwdPairs = 8;          // 16 warps / CTA, 1 CTA / SM
constexpr int kBwdPairs = 4;          // 8 warps / CTA, 1 CTA / SM
constexpr int kFacFloats = 520;       // u (8) + v (512) per early-set window

struct alignas(64) FastParams {
    CUtensorMap tm_x;     // X volume
    CUtensorMap tm_g;     // dY volume (backward)
    CUtensorMap tm_out;   // dX volume (backward: partial-sum load and store)
    const float* x;
    const float* gy;
    float* out;
    const float* v0;
    float* saved;
    float* fac;           // forward: factors of the early sets
    int* ctr;             // [0..1] 64-bit work counter, [2 + (set*B + b)*NR + row] finished windows
    int n0, n1, n2, G0, G1, G2, heads, B, S, C;
    int sh[FZ_MAX_SHIFTS][3];   // shifts normalised into [0, n)
    int dep_of[FZ_MAX_SHIFTS];  // backward: set whose partial sum this set continues (-1: chain head)
    int fac_idx[FZ_MAX_SHIFTS]; // forward: slot of an early set in `fac` (-1: the final set)
    int signals[FZ_MAX_SHIFTS]; // does a finished window of this set bump its row counter?
    int final_set;
    long long vox;
    int NR, TPR, entries;
    long long total_items;
    int T, K, relu, rec_floats;
    float eps, inv_S;
    int debug;            // FZ_DEBUG_FLAGS: 1 = skip dependency waits, 2 = skip output (timing experiments only)
    unsigned order[kMaxOrder];  // (set << 24) | row, in processing order
};

// ---- PTX helpers ------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, int n) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(n));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* b) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* b, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                 : "=r"(ok) : "r"(smem_u32(b)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) { while (!mbar_try_wait(b, parity)) {} }
__device__ __forceinline__ void tma_load_tile(void* dst, const CUtensorMap* m, uint64_t* bar, int cw, int ch, int cd, int cc, int cb) {
    asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(m), "r"(smem_u32(bar)), "r"(cw), "r"(ch), "r"(cd), "r"(cc), "r"(cb) : "memory");
}
__device__ __forceinline__ void tma_store_tile(const CUtensorMap* m, const void* src, int cw, int ch, int cd, int cc, int cb) {
    asm volatile("cp.async.bulk.tensor.5d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4, %5, %6}], [%1];"
                 ::"l"(m), "r"(smem_u32(src)), "r"(cw), "r"(ch), "r"(cd), "r"(cc), "r"(cb) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read_all() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ void pair_bar(int id) { asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory"); }
__device__ __forceinline__ int ld_acquire(const int* p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// ---- work decoding ----------------------------------------------------------------------------------
struct Info {             // one claimed window, decoded by the pair leader and shared through smem
    long long item;       // >= total_items: no more work
    long long win_id;     // canonical index ((set*B + b)*heads + head)*G + window
    long long chan_base;  // element offset of channel head*8 of sample b
    int b, set, row, head;
    int c0, c1, c2;       // first voxel of the window in the un-rolled volume
    int flags;            // 1: one contiguous, 16-byte aligned box (TMA);  2: c2 % 4 == 0 (float4 direct access)
};
constexpr int kInterior = 1, kAligned4 = 2;

__device__ __forceinline__ void decode(const FastParams& P, long long item, Info* o) {
    o->item = item;
    if (item >= P.total_items) return;
    const long long per_sample = (long long)P.entries * P.TPR;
    const int b = (int)(item / per_sample);
    const int rem = (int)(item - (long long)b * per_sample);
    const int e = rem / P.TPR, wi = rem - e * P.TPR;
    const unsigned entry = P.order[e];
    const int set = (int)(entry >> 24), row = (int)(entry & 0xffffffu);
    const int k = row / P.G1, g1 = row - k * P.G1;
    const int head = wi / P.G2, g2 = wi - head * P.G2;
    int c0 = k * 8 - P.sh[set][0]; if (c0 < 0) c0 += P.n0;
    int c1 = g1 * 8 - P.sh[set][1]; if (c1 < 0) c1 += P.n1;
    int c2 = g2 * 8 - P.sh[set][2]; if (c2 < 0) c2 += P.n2;
    const bool al = (c2 & 3) == 0;
    const bool inside = (c0 + 8 <= P.n0) && (c1 + 8 <= P.n1) && (c2 + 8 <= P.n2);
    o->b = b; o->set = set; o->row = row; o->head = head;
    o->c0 = c0; o->c1 = c1; o->c2 = c2;
    o->flags = ((inside && al) ? kInterior : 0) | (al ? kAligned4 : 0);
    const long long G = (long long)P.NR * P.G2;
    o->win_id = (((long long)set * P.B + b) * P.heads + head) * G + (long long)row * P.G2 + g2;
    o->chan_base = ((long long)b * P.C + (long long)head * 8) * P.vox;
}

// Lane p owns chunks r = p and p + 64 of every row; chunk r <-> (q0 = r >> 4, q1 = (r >> 1) & 7,
// q2 = 4 * (r & 1) + e), i.e. matrix columns 4r .. 4r+3 (reference column order, operations.py:321-325).
struct LaneAddr {
    int rowoff[2];   // (i0 * n1 + i1) * n2 for the two chunks
    int i2[2];       // W coordinate of the chunk's first column (the following ones may wrap)
};
__device__ __forceinline__ LaneAddr lane_addr(const FastParams& P, const Info& it, int p) {
    LaneAddr a;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int r = p + 64 * h;
        int i0 = it.c0 + (r >> 4); if (i0 >= P.n0) i0 -= P.n0;
        int i1 = it.c1 + ((r >> 1) & 7); if (i1 >= P.n1) i1 -= P.n1;
        int i2 = it.c2 + 4 * (r & 1); if (i2 >= P.n2) i2 -= P.n2;
        a.rowoff[h] = (i0 * P.n1 + i1) * P.n2;
        a.i2[h] = i2;
    }
    return a;
}
__device__ __forceinline__ int wrap2(const FastParams& P, int i2) { return i2 >= P.n2 ? i2 - P.n2 : i2; }

__device__ __forceinline__ void load_rows_direct(const FastParams& P, const float* base, const Info& it, int p, float (&x)[8][8]) {
    const LaneAddr a = lane_addr(P, it, p);
    if (it.flags & kAligned4) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const float4 v = __ldcg(reinterpret_cast<const float4*>(base + (long long)i * P.vox + a.rowoff[h] + a.i2[h]));
                x[i][4 * h] = v.x; x[i][4 * h + 1] = v.y; x[i][4 * h + 2] = v.z; x[i][4 * h + 3] = v.w;
            }
    } else {
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int h = 0; h < 2; ++h)
#pragma unroll
                for (int e = 0; e < 4; ++e)
                    x[i][4 * h + e] = __ldcg(base + (long long)i * P.vox + a.rowoff[h] + wrap2(P, a.i2[h] + e));
    }
}

// out = val (+ previous content when `accumulate`); plain stores, the caller owns these voxels.
__device__ __forceinline__ void store_rows_direct(const FastParams& P, float* base, const Info& it, int p,
                                                  const float (&x)[8][8], bool accumulate) {
    const LaneAddr a = lane_addr(P, it, p);
    if (it.flags & kAligned4) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                float4* dst = reinterpret_cast<float4*>(base + (long long)i * P.vox + a.rowoff[h] + a.i2[h]);
                float4 v = make_float4(x[i][4 * h], x[i][4 * h + 1], x[i][4 * h + 2], x[i][4 * h + 3]);
                if (accumulate) { const float4 o = __ldcg(dst); v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w; }
                *dst = v;
            }
    } else {
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int h = 0; h < 2; ++h)
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    float* dst = base + (long long)i * P.vox + a.rowoff[h] + wrap2(P, a.i2[h] + e);
                    float v = x[i][4 * h + e];
                    if (accumulate) v += __ldcg(dst);
                    *dst = v;
                }
    }
}

// Wait until every row of window set `dep` that this window overlaps has been finished.
__device__ __forceinline__ void wait_rows(const FastParams& P, const Info& it, int dep) {
    int r0 = it.c0 + P.sh[dep][0]; if (r0 >= P.n0) r0 -= P.n0;
    int r1 = it.c1 + P.sh[dep][1]; if (r1 >= P.n1) r1 -= P.n1;
    int r0b = r0 + 7; if (r0b >= P.n0) r0b -= P.n0;
    int r1b = r1 + 7; if (r1b >= P.n1) r1b -= P.n1;
    const int ka = r0 >> 3, kb = r0b >> 3, ga = r1 >> 3, gb = r1b >> 3;
    const int* done = P.ctr + 2 + ((long long)dep * P.B + it.b) * P.NR;
    const int rows[4] = {ka * P.G1 + ga, ka * P.G1 + gb, kb * P.G1 + ga, kb * P.G1 + gb};
#pragma unroll
    for (int q = 0; q < 4; ++q)
        while (ld_acquire(done + rows[q]) < P.TPR) __nanosleep(40);
}
__device__ __forceinline__ void signal_row(const FastParams& P, int set, int b, int row) {
    __threadfence();
    atomicAdd(P.ctr + 2 + ((long long)set * P.B + b) * P.NR + row, 1);
}

// 64-lane all-reduce of 8 row sums and one scalar.  red: 2 slots x 2 warps x 12 floats.
__device__ __forceinline__ void pair_reduce9(float (&v)[8], float& e, float* red, int& slot, int wip, int lane, int barid) {
    const bool h16 = lane & 16, h8 = lane & 8, h4 = lane & 4;
    float k4[4], k2[2], k1;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float keep = h16 ? v[j + 4] : v[j], send = h16 ? v[j] : v[j + 4];
        k4[j] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        const float keep = h8 ? k4[j + 2] : k4[j], send = h8 ? k4[j] : k4[j + 2];
        k2[j] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    }
    {
        const float keep = h4 ? k2[1] : k2[0], send = h4 ? k2[0] : k2[1];
        k1 = keep + __shfl_xor_sync(0xffffffffu, send, 4);
    }
    k1 += __shfl_xor_sync(0xffffffffu, k1, 2);
    k1 += __shfl_xor_sync(0xffffffffu, k1, 1);
    e = warp_sum(e);
    float* mine = red + slot * 24 + wip * 12;
    if ((lane & 3) == 0) mine[lane >> 2] = k1;   // row (lane >> 2) & 7
    if (lane == 0) mine[8] = e;
    pair_bar(barid);
    const float4* r4 = reinterpret_cast<const float4*>(red + slot * 24);
    const float4 a0 = r4[0], a1 = r4[1], b0 = r4[3], b1 = r4[4];
    v[0] = a0.x + b0.x; v[1] = a0.y + b0.y; v[2] = a0.z + b0.z; v[3] = a0.w + b0.w;
    v[4] = a1.x + b1.x; v[5] = a1.y + b1.y; v[6] = a1.z + b1.z; v[7] = a1.w + b1.w;
    e = red[slot * 24 + 8] + red[slot * 24 + 20];
    slot ^= 1;
}

// X tile -> registers.  x[i][0..3] = chunk p, x[i][4..7] = chunk p+64.
__device__ __forceinline__ void load_rows_smem(const float* tile, int p, float (&x)[8][8]) {
    const float4* t4 = reinterpret_cast<const float4*>(tile);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const float4 a = t4[i * 128 + p], b = t4[i * 128 + 64 + p];
        x[i][0] = a.x; x[i][1] = a.y; x[i][2] = a.z; x[i][3] = a.w;
        x[i][4] = b.x; x[i][5] = b.y; x[i][6] = b.z; x[i][7] = b.w;
    }
}
__device__ __forceinline__ void get8(const float4* src, int idx_a, int idx_b, float (&v)[8]) {
    const float4 a = src[idx_a], b = src[idx_b];
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}

struct PairCtx {
    int pair, wip, lane, p, barid;
    bool leader;
};
__device__ __forceinline__ PairCtx pair_ctx() {
    PairCtx c;
    const int warp = threadIdx.x >> 5;
    c.pair = warp >> 1; c.wip = warp & 1; c.lane = threadIdx.x & 31; c.p = c.wip * 32 + c.lane;
    c.barid = 1 + c.pair; c.leader = (c.p == 0);
    return c;
}

struct alignas(16) PairShared {       // static shared memory, one per pair
    float red[48];
    Info info[2];
    uint64_t xfull, xfree, gfull;
};

__device__ __forceinline__ long long claim(const FastParams& P) {
    return (long long)atomicAdd(reinterpret_cast<unsigned long long*>(P.ctr), 1ULL);
}
__device__ __forceinline__ void prefetch_x(const FastParams& P, const Info& nt, float* xin, uint64_t* xfull) {
    if (nt.item < P.total_items && (nt.flags & kInterior)) {
        mbar_expect_tx(xfull, kTileBytes);
        tma_load_tile(xin, &P.tm_x, xfull, nt.c2, nt.c1, nt.c0, nt.head * 8, nt.b);
    }
}

// =====================================================================================================
// forward
// =====================================================================================================
__global__ void __launch_bounds__(kFwdPairs * 64, 1) swnmf_fwd_fast(const __grid_constant__ FastParams P) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    __shared__ __align__(16) float v0s[512];
    __shared__ __align__(16) PairShared ps_all[kFwdPairs];

    const PairCtx c = pair_ctx();
    float* xin = reinterpret_cast<float*>(smem_raw) + (size_t)c.pair * 4096;
    PairShared& ps = ps_all[c.pair];
    const long long G = (long long)P.NR * P.G2;

    for (int j = threadIdx.x; j < 512; j += blockDim.x) v0s[j] = P.v0[j];
    long long claimed = 0;     // leader: item claimed one window ahead, not yet decoded
    if (c.leader) {
        mbar_init(&ps.xfull, 1);
        mbar_init(&ps.xfree, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        decode(P, claim(P), &ps.info[0]);
        decode(P, claim(P), &ps.info[1]);
        prefetch_x(P, ps.info[0], xin, &ps.xfull);
        claimed = claim(P);
    }
    __syncthreads();

    uint32_t xparity = 0, fparity = 0;
    int slot = 0;
    for (int n = 0;; ++n) {
        const Info it = ps.info[n & 1];
        if (it.item >= P.total_items) break;
        float x[8][8];
        if (it.flags & kInterior) {
            mbar_wait(&ps.xfull, xparity);
            xparity ^= 1;
            load_rows_smem(xin, c.p, x);
        } else {
            load_rows_direct(P, P.x + it.chan_base, it, c.p, x);
        }
        // xin is free once both warps have pulled their registers: warp 1 tells the leader
        __syncwarp();
        if (c.wip == 1 && c.lane == 0) mbar_arrive(&ps.xfree);
        if (c.leader) {
            mbar_wait(&ps.xfree, fparity);
            prefetch_x(P, ps.info[(n + 1) & 1], xin, &ps.xfull);
        }
        fparity ^= 1;
        if (P.relu) {
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int k = 0; k < 8; ++k) x[i][k] = fmaxf(x[i][k], 0.f);
        }

        // ---- T HALS sweeps, rank 1 (matrix_factorization.py:224-227 twice per sweep, :122-136) ----
        float v[8], u[8];
        get8(reinterpret_cast<const float4*>(v0s), c.p, 64 + c.p, v);
        float* rec = P.saved ? P.saved + it.win_id * P.rec_floats : nullptr;
        for (int t = 0; t < P.T; ++t) {
            float a[8], bb = 0.f;
#pragma unroll
            for (int i = 0; i < 8; ++i) a[i] = 0.f;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
#pragma unroll
                for (int i = 0; i < 8; ++i) a[i] = fmaf(x[i][k], v[k], a[i]);
                bb = fmaf(v[k], v[k], bb);
            }
            pair_reduce9(a, bb, ps.red, slot, c.wip, c.lane, c.barid);
            if (t == 0 && c.leader) {
                // every lane has copied info[n & 1] (barrier above): decode the item claimed a window
                // ago into that slot and claim the next one (its latency hides behind this window)
                decode(P, claimed, &ps.info[n & 1]);
                claimed = claim(P);
            }
            const float rb = __fdiv_rn(1.f, bb + P.eps);
            float d = 0.f;
#pragma unroll
            for (int i = 0; i < 8; ++i) { u[i] = fmaxf((a[i] + P.eps) * rb, 0.f); d = fmaf(u[i], u[i], d); }
            const float rd = __fdiv_rn(1.f, d + P.eps);
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                float cc = 0.f;
#pragma unroll
                for (int i = 0; i < 8; ++i) cc = fmaf(x[i][k], u[i], cc);
                v[k] = fmaxf((cc + P.eps) * rd, 0.f);
            }
            if (rec && c.leader) {
                reinterpret_cast<float4*>(rec)[2 * t] = make_float4(u[0], u[1], u[2], u[3]);
                reinterpret_cast<float4*>(rec)[2 * t + 1] = make_float4(u[4], u[5], u[6], u[7]);
                rec[8 * P.T + t] = bb;
            }
        }

        if (P.debug & 2) {
            if (u[0] * v[0] == 123.456f) P.out[0] = 1.f;
        } else if (it.set != P.final_set) {
            // ---- early set: publish the rank-1 factors only ----
            const long long local = it.win_id - ((long long)it.set * P.B + it.b) * P.heads * G;   // head*G + window
            float4* f4 = reinterpret_cast<float4*>(P.fac + (((long long)P.fac_idx[it.set] * P.B + it.b) * P.heads * G + local) * kFacFloats);
            if (c.leader) { f4[0] = make_float4(u[0], u[1], u[2], u[3]); f4[1] = make_float4(u[4], u[5], u[6], u[7]); }
            f4[2 + c.p] = make_float4(v[0], v[1], v[2], v[3]);
            f4[2 + 64 + c.p] = make_float4(v[4], v[5], v[6], v[7]);
            pair_bar(c.barid);
            if (c.leader) signal_row(P, it.set, it.b, it.row);
        } else {
            // ---- final set: Y = (u v^T + sum over early sets of their overlapping factors) / S ----
            if (P.S > 1) {
                if (c.leader && !(P.debug & 1))
                    for (int s = 0; s < P.S; ++s)
                        if (s != P.final_set) wait_rows(P, it, s);
                pair_bar(c.barid);
            }
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int k = 0; k < 8; ++k) x[i][k] = u[i] * v[k];      // x is dead: reuse as the accumulator
            const LaneAddr la = lane_addr(P, it, c.p);
            const int q0v[2] = {c.p >> 4, (c.p + 64) >> 4};
            int i1 = it.c1 + ((c.p >> 1) & 7); if (i1 >= P.n1) i1 -= P.n1;
            for (int s = 0; s < P.S; ++s) {
                if (s == P.final_set) continue;
                const float* fs = P.fac + (((long long)P.fac_idx[s] * P.B + it.b) * P.heads + it.head) * G * kFacFloats;
                // position of this lane's columns in set s's rolled coordinates
                int rho1 = i1 + P.sh[s][1]; if (rho1 >= P.n1) rho1 -= P.n1;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    int i0 = it.c0 + q0v[h]; if (i0 >= P.n0) i0 -= P.n0;
                    int rho0 = i0 + P.sh[s][0]; if (rho0 >= P.n0) rho0 -= P.n0;
                    const int wrow = ((rho0 >> 3) * P.G1 + (rho1 >> 3)) * P.G2;
                    const int jrow = ((rho0 & 7) * 8 + (rho1 & 7)) * 8;
                    int rho2 = la.i2[h] + P.sh[s][2]; if (rho2 >= P.n2) rho2 -= P.n2;
                    if ((rho2 & 3) == 0) {
                        const float* fw = fs + (long long)(wrow + (rho2 >> 3)) * kFacFloats;
                        const float4 ua = __ldcg(reinterpret_cast<const float4*>(fw));
                        const float4 ub = __ldcg(reinterpret_cast<const float4*>(fw) + 1);
                        const float4 vv = __ldcg(reinterpret_cast<const float4*>(fw + 8 + jrow + (rho2 & 7)));
                        const float uu[8] = {ua.x, ua.y, ua.z, ua.w, ub.x, ub.y, ub.z, ub.w};
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            x[i][4 * h] = fmaf(uu[i], vv.x, x[i][4 * h]);
                            x[i][4 * h + 1] = fmaf(uu[i], vv.y, x[i][4 * h + 1]);
                            x[i][4 * h + 2] = fmaf(uu[i], vv.z, x[i][4 * h + 2]);
                            x[i][4 * h + 3] = fmaf(uu[i], vv.w, x[i][4 * h + 3]);
                        }
                    } else {
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            int r2 = rho2 + e; if (r2 >= P.n2) r2 -= P.n2;
                            const float* fw = fs + (long long)(wrow + (r2 >> 3)) * kFacFloats;
                            const float vv = __ldcg(fw + 8 + jrow + (r2 & 7));
#pragma unroll
                            for (int i = 0; i < 8; ++i) x[i][4 * h + e] = fmaf(__ldcg(fw + i), vv, x[i][4 * h + e]);
                        }
                    }
                }
            }
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int k = 0; k < 8; ++k) x[i][k] *= P.inv_S;
            store_rows_direct(P, P.out + it.chan_base, it, c.p, x, false);
        }
    }
}

// =====================================================================================================
// backward
// =====================================================================================================
__global__ void __launch_bounds__(kBwdPairs * 64, 1) swnmf_bwd_fast(const __grid_constant__ FastParams P) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    __shared__ __align__(16) float v0s[512];
    __shared__ __align__(16) float rec_all[kBwdPairs][80];
    __shared__ __align__(16) PairShared ps_all[kBwdPairs];

    const PairCtx c = pair_ctx();
    const int per_pair = 2 * kTileBytes + P.T * 2048;
    float* xin = reinterpret_cast<float*>(smem_raw + (size_t)c.pair * per_pair);
    float* gout = xin + 4096;        // dY lands here; the partial dX sum is added and leaves from here
    float* stash = gout + 4096;      // v_1 .. v_T, lane-private columns
    float* rec = rec_all[c.pair];
    PairShared& ps = ps_all[c.pair];

    for (int j = threadIdx.x; j < 512; j += blockDim.x) v0s[j] = P.v0[j];
    long long claimed = 0;
    if (c.leader) {
        mbar_init(&ps.xfull, 1);
        mbar_init(&ps.xfree, 1);
        mbar_init(&ps.gfull, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        decode(P, claim(P), &ps.info[0]);
        decode(P, claim(P), &ps.info[1]);
        prefetch_x(P, ps.info[0], xin, &ps.xfull);
        claimed = claim(P);
    }
    __syncthreads();

    uint32_t xparity = 0, fparity = 0, gparity = 0;
    int slot = 0;
    int pend_set = -1, pend_b = 0, pend_row = 0;   // leader: TMA-stored window not yet signalled
    for (int n = 0;; ++n) {
        const Info it = ps.info[n & 1];
        if (it.item >= P.total_items) break;
        const bool interior = it.flags & kInterior;
        // dY tile: gout is free once the previous dX store has been read out of it
        if (c.leader && interior) {
            bulk_wait_read_all();
            mbar_expect_tx(&ps.gfull, kTileBytes);
            tma_load_tile(gout, &P.tm_g, &ps.gfull, it.c2, it.c1, it.c0, it.head * 8, it.b);
        }
        // per-window iterate summary saved by the forward
        if (c.p < P.rec_floats / 4)
            reinterpret_cast<float4*>(rec)[c.p] = __ldcg(reinterpret_cast<const float4*>(P.saved + it.win_id * P.rec_floats) + c.p);

        float x[8][8];
        if (interior) {
            mbar_wait(&ps.xfull, xparity);
            xparity ^= 1;
            load_rows_smem(xin, c.p, x);
        } else {
            load_rows_direct(P, P.x + it.chan_base, it, c.p, x);
        }
        __syncwarp();
        if (c.wip == 1 && c.lane == 0) mbar_arrive(&ps.xfree);
        if (c.leader) {
            mbar_wait(&ps.xfree, fparity);
            prefetch_x(P, ps.info[(n + 1) & 1], xin, &ps.xfull);
        }
        fparity ^= 1;
        unsigned long long mask = ~0ULL;   // bit i*8+k: X[i][k] > 0 (ReLU adjoint, factorizer.py:44)
        if (P.relu) {
            mask = 0ULL;
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    if (x[i][k] > 0.f) mask |= 1ULL << (i * 8 + k); else x[i][k] = 0.f;
                }
        }
        pair_bar(c.barid);            // rec visible to both warps; every lane holds its copy of `it`
        if (c.leader) {
            decode(P, claimed, &ps.info[n & 1]);
            claimed = claim(P);
        }

        // ---- P1: recompute v_1..v_T from the saved u_t (one lane-local GEMV each) ----
        float4* st4 = reinterpret_cast<float4*>(stash);
        const float4* rec4 = reinterpret_cast<const float4*>(rec);
        float vT[8];
        for (int t = 0; t < P.T; ++t) {
            float u[8];
            get8(rec4, 2 * t, 2 * t + 1, u);
            float d = 0.f;
#pragma unroll
            for (int i = 0; i < 8; ++i) d = fmaf(u[i], u[i], d);
            const float rd = __fdiv_rn(1.f, d + P.eps);
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                float cc = 0.f;
#pragma unroll
                for (int i = 0; i < 8; ++i) cc = fmaf(x[i][k], u[i], cc);
                vT[k] = fmaxf((cc + P.eps) * rd, 0.f);
            }
            st4[t * 128 + c.p] = make_float4(vT[0], vT[1], vT[2], vT[3]);
            st4[t * 128 + 64 + c.p] = make_float4(vT[4], vT[5], vT[6], vT[7]);
        }
        // the previous window's dX store has long completed: tell the sets that continue its sum
        if (c.leader && pend_set >= 0) {
            bulk_wait_all();
            signal_row(P, pend_set, pend_b, pend_row);
            pend_set = -1;
        }

        // ---- P2: gu = G v_T / S, gv = G^T u_T / S  (adjoint of u v^T and of the mean over sets) ----
        float gu[8], gv[8];
        {
            float uT[8];
            get8(rec4, 2 * (P.T - 1), 2 * (P.T - 1) + 1, uT);
#pragma unroll
            for (int k = 0; k < 8; ++k) gv[k] = 0.f;
            float g[8][8];
            if (interior) {
                mbar_wait(&ps.gfull, gparity);
                gparity ^= 1;
                load_rows_smem(gout, c.p, g);
            } else {
                load_rows_direct(P, P.gy + it.chan_base, it, c.p, g);
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                float acc = 0.f;
#pragma unroll
                for (int k = 0; k < 8; ++k) { acc = fmaf(g[i][k], vT[k], acc); gv[k] = fmaf(g[i][k], uT[i], gv[k]); }
                gu[i] = acc;
            }
            float dummy = 0.f;
            pair_reduce9(gu, dummy, ps.red, slot, c.wip, c.lane, c.barid);   // also: both warps are done reading gout
#pragma unroll
            for (int i = 0; i < 8; ++i) gu[i] *= P.inv_S;
#pragma unroll
            for (int k = 0; k < 8; ++k) gv[k] *= P.inv_S;
        }
        // partial dX left by the previous set of the chain: fetch it into gout while P3 runs
        const int dep = P.dep_of[it.set];
        bool partial_in_smem = false;
        if (dep >= 0) {
            if (c.leader && !(P.debug & 1)) wait_rows(P, it, dep);
            if (interior) {
                if (c.leader) {
                    fence_proxy_async_all();
                    mbar_expect_tx(&ps.gfull, kTileBytes);
                    tma_load_tile(gout, &P.tm_out, &ps.gfull, it.c2, it.c1, it.c0, it.head * 8, it.b);
                }
                partial_in_smem = true;
            }
        }

        // ---- P3: reverse sweep; dX accumulates in registers ----
        float xb[8][8];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int k = 0; k < 8; ++k) xb[i][k] = 0.f;
        for (int t = P.T - 1; t >= P.T - P.K; --t) {
            float u[8];
            get8(rec4, 2 * t, 2 * t + 1, u);
            float d = 0.f;
#pragma unroll
            for (int i = 0; i < 8; ++i) d = fmaf(u[i], u[i], d);
            const float rd = __fdiv_rn(1.f, d + P.eps);
            const float rb = __fdiv_rn(1.f, rec[8 * P.T + t] + P.eps);
            float vt[8];
            get8(st4, t * 128 + c.p, t * 128 + 64 + c.p, vt);
            // adjoint of v_t = relu((X^T u_t + eps) / (d_t + eps))
            float cb[8], e = 0.f, w[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const float qb = vt[k] > 0.f ? gv[k] : 0.f;
                cb[k] = qb * rd;
                e = fmaf(qb, vt[k], e);
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) w[i] = 0.f;
#pragma unroll
            for (int k = 0; k < 8; ++k)
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    xb[i][k] = fmaf(u[i], cb[k], xb[i][k]);
                    w[i] = fmaf(x[i][k], cb[k], w[i]);
                }
            pair_reduce9(w, e, ps.red, slot, c.wip, c.lane, c.barid);
            const float db = -e * rd;
            // adjoint of u_t = relu((X v_{t-1} + eps) / (b_t + eps)); u_{t-1} does not feed u_t at rank 1
            float ab[8], bacc = 0.f;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                float ub = fmaf(2.f * db, u[i], w[i]);
                if (t == P.T - 1) ub += gu[i];
                const float pb = u[i] > 0.f ? ub : 0.f;
                ab[i] = pb * rb;
                bacc = fmaf(pb, u[i], bacc);
            }
            const float bbar2 = -2.f * bacc * rb;
            float vp[8];
            if (t > 0) get8(st4, (t - 1) * 128 + c.p, (t - 1) * 128 + 64 + c.p, vp);
            else get8(reinterpret_cast<const float4*>(v0s), c.p, 64 + c.p, vp);
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                float acc = bbar2 * vp[k];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    xb[i][k] = fmaf(ab[i], vp[k], xb[i][k]);
                    acc = fmaf(x[i][k], ab[i], acc);
                }
                gv[k] = acc;
            }
        }

        // ---- P4: dX (ReLU-masked), added to the chain's partial sum and stored ----
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int k = 0; k < 8; ++k)
                if (!((mask >> (i * 8 + k)) & 1ULL)) xb[i][k] = 0.f;
        if (P.debug & 2) {
            if (xb[0][0] == 123.456f) P.out[0] = 1.f;
            if (partial_in_smem) { mbar_wait(&ps.gfull, gparity); gparity ^= 1; }
        } else if (interior) {
            float4* o4 = reinterpret_cast<float4*>(gout);
            if (partial_in_smem) {
                mbar_wait(&ps.gfull, gparity);
                gparity ^= 1;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float4 a = o4[i * 128 + c.p], b = o4[i * 128 + 64 + c.p];
                    xb[i][0] += a.x; xb[i][1] += a.y; xb[i][2] += a.z; xb[i][3] += a.w;
                    xb[i][4] += b.x; xb[i][5] += b.y; xb[i][6] += b.z; xb[i][7] += b.w;
                }
            }
            // each lane rewrites exactly the chunks it read: no cross-lane hazard on gout
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                o4[i * 128 + c.p] = make_float4(xb[i][0], xb[i][1], xb[i][2], xb[i][3]);
                o4[i * 128 + 64 + c.p] = make_float4(xb[i][4], xb[i][5], xb[i][6], xb[i][7]);
            }
            fence_proxy_async_smem();
            pair_bar(c.barid);
            if (c.leader) {
                tma_store_tile(&P.tm_out, gout, it.c2, it.c1, it.c0, it.head * 8, it.b);
                bulk_commit();
                if (P.signals[it.set]) { pend_set = it.set; pend_b = it.b; pend_row = it.row; }
            }
        } else {
            store_rows_direct(P, P.out + it.chan_base, it, c.p, xb, dep >= 0);
            if (P.signals[it.set]) {
                pair_bar(c.barid);
                if (c.leader) signal_row(P, it.set, it.b, it.row);
            }
        }
    }
    if (c.leader) {
        bulk_wait_all();
        if (pend_set >= 0) signal_row(P, pend_set, pend_b, pend_row);
    }
}

// =====================================================================================================
// host side
// =====================================================================================================
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

static int make_map(CUtensorMap* m, const void* ptr, const DevGeom& G) {
    EncodeTiledFn enc = encode_fn();
    if (!enc) return fail(FZ_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
    if (reinterpret_cast<uintptr_t>(ptr) & 15) return fail(FZ_ERR_INVALID, "volume pointer %p is not 16-byte aligned", ptr);
    cuuint64_t dims[5] = {(cuuint64_t)G.n[2], (cuuint64_t)G.n[1], (cuuint64_t)G.n[0], (cuuint64_t)G.C, (cuuint64_t)G.B};
    cuuint64_t strides[4] = {(cuuint64_t)G.n[2] * 4, (cuuint64_t)G.n[2] * G.n[1] * 4, (cuuint64_t)G.vox * 4,
                             (cuuint64_t)G.vox * G.C * 4};
    cuuint32_t box[5] = {8, 8, 8, 8, 1};
    cuuint32_t es[5] = {1, 1, 1, 1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, const_cast<void*>(ptr), dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(FZ_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    return FZ_OK;
}

static int rec_floats_for(int T) { return ((9 * T + 3) / 4) * 4; }
static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

bool fast_supported(const DevGeom& G, const fz_solver& s) {
    if (s.kind != FZ_SOLVER_HALS || s.rank != 1) return false;
    if (s.num_iters < 1 || s.num_iters > kMaxT) return false;
    if (G.d != 8 || G.p[0] != 8 || G.p[1] != 8 || G.p[2] != 8) return false;
    if ((long long)G.S * G.g[0] * G.g[1] > kMaxOrder) return false;
    if ((long long)G.B * G.g[0] * G.g[1] > (1 << 24)) return false;
    if (G.mats_per_shift == 0) return false;
    return true;
}

size_t fast_saved_bytes(const DevGeom& G, const fz_solver& s) {
    if (!fast_supported(G, s)) return 0;
    return (size_t)G.S * G.mats_per_shift * rec_floats_for(s.num_iters) * sizeof(float);
}

static size_t counter_bytes(const DevGeom& G) {
    return align_up((size_t)(2 + (long long)G.S * G.B * G.g[0] * G.g[1]) * sizeof(int), 256);
}

size_t fast_workspace_bytes(const DevGeom& G, const fz_solver& s) {
    if (!fast_supported(G, s)) return 0;
    return counter_bytes(G) + (size_t)(G.S - 1) * G.mats_per_shift * kFacFloats * sizeof(float);
}

// rows of set `dep` overlapped by row `row` of set `s`; returns the largest of `when[]` over them
static long long latest_needed(const DevGeom& G, const FastParams& P, int s, int row, int dep, const long long* when) {
    const int k = row / G.g[1], g1 = row % G.g[1];
    int c0 = k * 8 - P.sh[s][0]; if (c0 < 0) c0 += G.n[0];
    int c1 = g1 * 8 - P.sh[s][1]; if (c1 < 0) c1 += G.n[1];
    const int r0 = (c0 + P.sh[dep][0]) % G.n[0], r1 = (c1 + P.sh[dep][1]) % G.n[1];
    const int r0b = (r0 + 7) % G.n[0], r1b = (r1 + 7) % G.n[1];
    const int ks[2] = {r0 / 8, r0b / 8}, gs[2] = {r1 / 8, r1b / 8};
    long long best = 0;
    for (int a = 0; a < 2; ++a)
        for (int b = 0; b < 2; ++b) {
            const long long t = when[ks[a] * G.g[1] + gs[b]];
            if (t > best) best = t;
        }
    return best;
}

// Processing order: `when[set][row]` is a virtual time; chain heads run at time = row, dependants
// `lag` rows after the last row they need.  forward: all early sets are heads, the final set depends
// on all of them.  backward: set j depends on set j-1.
static void build_order(const DevGeom& G, FastParams& P, bool forward, int pairs_in_flight) {
    const int NR = G.g[0] * G.g[1];
    static thread_local long long when[FZ_MAX_SHIFTS][kMaxOrder];
    struct Key { long long key; unsigned entry; };
    static thread_local Key keys[kMaxOrder];
    // a dependant must trail by about two generations of in-flight windows (one to finish computing,
    // one for its completion signal), measured in rows of its own set
    const long long per_row_all_sets = (long long)G.heads * G.g[2] * G.S;
    int lag = (int)((2LL * pairs_in_flight + per_row_all_sets - 1) / per_row_all_sets) + 2;
    if (const char* env = getenv("FZ_LAG_ROWS")) lag = atoi(env);
    if (forward) {
        for (int s = 0; s < G.S; ++s)
            if (s != P.final_set)
                for (int row = 0; row < NR; ++row) when[s][row] = row;
        for (int row = 0; row < NR; ++row) {
            long long t = row;
            for (int s = 0; s < G.S; ++s)
                if (s != P.final_set) {
                    const long long need = latest_needed(G, P, P.final_set, row, s, when[s]) + lag;
                    if (need > t) t = need;
                }
            when[P.final_set][row] = t;
        }
    } else {
        for (int s = 0; s < G.S; ++s)
            for (int row = 0; row < NR; ++row)
                when[s][row] = (s == 0) ? row : latest_needed(G, P, s, row, s - 1, when[s - 1]) + lag;
    }
    int n = 0;
    for (int s = 0; s < G.S; ++s)
        for (int row = 0; row < NR; ++row) {
            keys[n].key = when[s][row] * FZ_MAX_SHIFTS + s;
            keys[n].entry = ((unsigned)s << 24) | (unsigned)row;
            ++n;
        }
    for (int i = 1; i < n; ++i) {   // insertion sort: the per-set runs are already nearly ordered
        const Key t = keys[i];
        int j = i - 1;
        while (j >= 0 && keys[j].key > t.key) { keys[j + 1] = keys[j]; --j; }
        keys[j + 1] = t;
    }
    for (int i = 0; i < n; ++i) P.order[i] = keys[i].entry;
    P.entries = n;
}

static int num_sms();
static int fill_params(FastParams& P, const DevGeom& G, const fz_solver& s, int K, int relu, bool forward) {
    memset(&P, 0, sizeof(P));
    P.n0 = G.n[0]; P.n1 = G.n[1]; P.n2 = G.n[2];
    P.G0 = G.g[0]; P.G1 = G.g[1]; P.G2 = G.g[2];
    P.heads = G.heads; P.B = G.B; P.S = G.S; P.C = G.C; P.vox = G.vox;
    for (int q = 0; q < G.S; ++q)
        for (int k = 0; k < 3; ++k) {
            int v = G.sh[q][k] % G.n[k];
            if (v < 0) v += G.n[k];
            P.sh[q][k] = v;
        }
    P.NR = G.g[0] * G.g[1];
    P.TPR = G.heads * G.g[2];
    // forward: the set that writes Y should have only clean boxes -> prefer an unshifted one
    P.final_set = G.S - 1;
    for (int q = 0; q < G.S; ++q)
        if (P.sh[q][0] % 8 == 0 && P.sh[q][1] % 8 == 0 && P.sh[q][2] % 8 == 0) { P.final_set = q; break; }
    int slot = 0;
    for (int q = 0; q < FZ_MAX_SHIFTS; ++q) {
        P.fac_idx[q] = (q < G.S && q != P.final_set) ? slot++ : -1;
        P.dep_of[q] = (q > 0 && q < G.S) ? q - 1 : -1;
        P.signals[q] = forward ? (q < G.S && q != P.final_set) : (q + 1 < G.S);
    }
    build_order(G, P, forward, num_sms() * (forward ? kFwdPairs : kBwdPairs));
    P.total_items = (long long)G.B * P.entries * P.TPR;
    P.T = s.num_iters; P.K = K; P.relu = relu; P.rec_floats = rec_floats_for(s.num_iters);
    P.eps = s.eps; P.inv_S = 1.0f / (float)G.S;
    { const char* env = getenv("FZ_DEBUG_FLAGS"); P.debug = env ? atoi(env) : 0; }
    return FZ_OK;
}

static int num_sms() {
    static int sms = 0;
    if (!sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (sms <= 0) sms = 148;
    }
    return sms;
}

int fast_forward(const float* x, const float* u0, const float* v0, float* y, void* saved,
                 void* workspace, const DevGeom& G, const fz_solver& s, int relu, cudaStream_t st) {
    (void)u0;  // at rank 1 the HALS update of u does not read the previous u (matrix_factorization.py:224-227)
    if (!workspace) return fail(FZ_ERR_INVALID, "fz_swnmf_forward: workspace of %zu bytes required", fast_workspace_bytes(G, s));
    static thread_local FastParams P;
    if (int e = fill_params(P, G, s, 0, relu, true)) return e;
    if (int e = make_map(&P.tm_x, x, G)) return e;
    P.tm_g = P.tm_x;
    P.tm_out = P.tm_x;
    if (reinterpret_cast<uintptr_t>(y) & 15) return fail(FZ_ERR_INVALID, "output pointer %p is not 16-byte aligned", (void*)y);
    P.x = x; P.out = y; P.v0 = v0; P.saved = static_cast<float*>(saved);
    P.ctr = static_cast<int*>(workspace);
    P.fac = reinterpret_cast<float*>(static_cast<char*>(workspace) + counter_bytes(G));
    FZ_CUDA_CHECK(cudaMemsetAsync(workspace, 0, counter_bytes(G), st));
    const size_t smem = (size_t)kFwdPairs * kTileBytes;
    static bool attr_set = false;
    if (!attr_set) {
        FZ_CUDA_CHECK(cudaFuncSetAttribute(swnmf_fwd_fast, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set = true;
    }
    long long ctas_needed = (P.total_items + kFwdPairs - 1) / kFwdPairs;
    int grid = num_sms();
    if (ctas_needed < grid) grid = (int)ctas_needed;
    swnmf_fwd_fast<<<grid, kFwdPairs * 64, smem, st>>>(P);
    FZ_LAUNCH_CHECK();
    return FZ_OK;
}

int fast_backward(const float* x, const float* gy, const float* u0, const float* v0,
                  const void* saved, float* gx, void* workspace, const DevGeom& G,
                  const fz_solver& s, int K, int relu, cudaStream_t st) {
    (void)u0;
    if (!workspace) return fail(FZ_ERR_INVALID, "fz_swnmf_backward: workspace of %zu bytes required", fast_workspace_bytes(G, s));
    if (!saved) return fail(FZ_ERR_INVALID, "fz_swnmf_backward: the `saved` buffer written by fz_swnmf_forward is required");
    static thread_local FastParams P;
    if (int e = fill_params(P, G, s, K, relu, false)) return e;
    if (int e = make_map(&P.tm_x, x, G)) return e;
    if (int e = make_map(&P.tm_g, gy, G)) return e;
    if (int e = make_map(&P.tm_out, gx, G)) return e;
    P.x = x; P.gy = gy; P.out = gx; P.v0 = v0;
    P.saved = const_cast<float*>(static_cast<const float*>(saved));
    P.ctr = static_cast<int*>(workspace);
    FZ_CUDA_CHECK(cudaMemsetAsync(workspace, 0, counter_bytes(G), st));
    const size_t smem = (size_t)kBwdPairs * (2 * kTileBytes + (size_t)s.num_iters * 2048);
    static size_t attr_smem = 0;
    if (smem > attr_smem) {
        FZ_CUDA_CHECK(cudaFuncSetAttribute(swnmf_bwd_fast, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_smem = smem;
    }
    long long ctas_needed = (P.total_items + kBwdPairs - 1) / kBwdPairs;
    int grid = num_sms();
    if (ctas_needed < grid) grid = (int)ctas_needed;
    swnmf_bwd_fast<<<grid, kBwdPairs * 64, smem, st>>>(P);
    FZ_LAUNCH_CHECK();
    return FZ_OK;
}

}  // namespace fz
