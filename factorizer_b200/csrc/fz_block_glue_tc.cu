// Tensor-core (tcgen05 / TMEM) version of mixer_mlp_fwd for 32-channel blocks:
//     x1 = x + W_out m + b_out ;  out = x1 + W2 gelu(W1 LN(x1) + b1) + b2
// (reference factorizer/factorizer.py:53,75-76, layers/mlp.py:54-60, layers/norm.py:29-34).
//
// The three projections are GEMMs with M = voxels: a CTA of 128 threads owns 128 voxels per tile, thread = voxel.
//   * A operands (m, LN(x1), gelu(h)) live in shared memory exactly as they lie in an NCDHW tensor -- voxel-contiguous
//     = "MN-major" -- in the one layout tcgen05 accepts for 32-bit MN-major data, SWIZZLE_128B_BASE32B: atoms of
//     4 channels x 32 voxels, a channel row = 128 contiguous bytes whose 32-byte chunks are XOR-ed with (channel % 4).
//     A warp writes one such row per channel: conflict-free.
//   * B operands = the weights, (out, in) row-major = K-major, 8 x 16-byte core matrices, staged once per CTA.
//   * D lives in TMEM (128 lanes = voxels x N columns); tcgen05.ld 32x32b returns to every thread ITS voxel's row,
//     so bias / residual / LayerNorm / GELU run per thread in registers with no shuffles, and the next A operand is
//     written straight back to shared memory.
// fp32 parity rules out one TF32 pass (error 5e-3 on O(1) sums): every product is 3xTF32,
//     a b ~= a_lo b_hi + a_hi b_lo + a_hi b_hi,   x_hi = x with the low 13 mantissa bits cleared (what the tensor core
//     reads anyway), x_lo = x - x_hi (exact, fits TF32),
// measured 5e-7 relative (profiles/r01_probes.md).  Descriptor encodings were pinned by bench_probes/tcgen05_probe.cu.
#include "fz_common.cuh"

namespace fz {
namespace {

constexpr int kC = 32;
constexpr int kTM = 128;              // voxels per tile = threads per CTA
constexpr int kMaxHid = 64;

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
// instruction descriptor: D = F32, A = B = TF32, A MN-major, B K-major, M = 128
__device__ __forceinline__ uint32_t make_idesc(int n) {
    return (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(kTM >> 4) << 24);
}
__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
        :: "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// A element (voxel m of the tile, channel k): SWIZZLE_128B_BASE32B, K groups of 4 channels 2 KiB apart
__device__ __forceinline__ uint32_t a_off(int m, int k) {
    const int r = k & 3, j = m & 31;
    return (uint32_t)((k >> 2) * 2048 + (m >> 5) * 512 + r * 128 + (((j >> 3) ^ r) << 5) + (j & 7) * 4);
}
// W element (output n, input k) of an (N, KC) matrix: 8 x 4 core matrices of 128 B, K-adjacent, N groups KC*32 B apart
__device__ __forceinline__ uint32_t b_off(int n, int k, int KC) {
    return (uint32_t)((n >> 3) * (KC * 32) + (k >> 2) * 128 + (n & 7) * 16 + (k & 3) * 4);
}
__device__ __forceinline__ float tf32_hi(float v) { return __uint_as_float(__float_as_uint(v) & 0xffffe000u); }

// D[tmem_d .. +N) (+)= A[128 x KC] W[N x KC]^T, 3xTF32; issued by one thread
__device__ __forceinline__ void gemm3(uint32_t tmem_d, const unsigned char* a_hi, const unsigned char* a_lo, const unsigned char* b_hi,
                                      const unsigned char* b_lo, int KC, int N) {
    const uint32_t idesc = make_idesc(N);
    const uint32_t sbo_b = (uint32_t)KC * 32;
    for (int s = 0; s < KC / 8; ++s) {
        const uint64_t ah = make_desc(smem_u32(a_hi) + s * 4096, 512, 2048, 1);
        const uint64_t al = make_desc(smem_u32(a_lo) + s * 4096, 512, 2048, 1);
        const uint64_t bh = make_desc(smem_u32(b_hi) + s * 256, 128, sbo_b, 0);
        const uint64_t bl = make_desc(smem_u32(b_lo) + s * 256, 128, sbo_b, 0);
        mma_tf32(tmem_d, al, bh, idesc, s > 0);
        mma_tf32(tmem_d, ah, bl, idesc, 1);
        mma_tf32(tmem_d, ah, bh, idesc, 1);
    }
}

// UTCHMMA reads uniform registers: issued under `if (tid == 0)` every MMA gets an ELECT / R2UR.BROADCAST / BRA.U.ANY loop
// around it (~70 cycles each, measured); from warp-uniform control flow on warp-uniform values under elect_one() it is one
// predicated instruction (bench_probes/tcgen05_rate_probe.cu)
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ uint32_t uniform_u32(uint32_t v) { return __shfl_sync(0xffffffffu, v, 0); }

__device__ __forceinline__ void commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void wait_bar(uint64_t* bar, uint32_t parity) {
    uint32_t ok = 0;
    while (!ok) {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// 32 consecutive TMEM columns of this thread's lane
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
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
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// all threads have written an A operand (generic proxy): make it visible to the tensor core, then meet
__device__ __forceinline__ void publish_and_sync() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}

__device__ __forceinline__ float ex2_approx(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcp_approx(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

// exact-erf GELU of two values (Abramowitz-Stegun 7.1.26, |error| <= 1.5e-7; same formula as csrc/fz_block_glue.cu),
// polynomial in packed FFMA2, one MUFU.RCP + one MUFU.EX2 per value
__device__ __forceinline__ float2 gelu_exact2(float2 h) {
    const float2 z = make_float2(fabsf(h.x) * 0.70710678118654752f, fabsf(h.y) * 0.70710678118654752f);
    const float2 d = __ffma2_rn(make_float2(0.3275911f, 0.3275911f), z, make_float2(1.f, 1.f));
    const float2 t = make_float2(rcp_approx(d.x), rcp_approx(d.y));
    const float2 hh = __fmul2_rn(h, h);
    const float2 e = make_float2(ex2_approx(hh.x * -0.72134752044448170f), ex2_approx(hh.y * -0.72134752044448170f));
    float2 p = __ffma2_rn(t, make_float2(1.061405429f, 1.061405429f), make_float2(-1.453152027f, -1.453152027f));
    p = __ffma2_rn(t, p, make_float2(1.421413741f, 1.421413741f));
    p = __ffma2_rn(t, p, make_float2(-0.284496736f, -0.284496736f));
    p = __ffma2_rn(t, p, make_float2(0.254829592f, 0.254829592f));
    const float2 pt = __fmul2_rn(p, t);
    const float2 erf_abs = __ffma2_rn(make_float2(-pt.x, -pt.y), e, make_float2(1.f, 1.f));
    const float2 cdf = __ffma2_rn(make_float2(copysignf(0.5f, h.x), copysignf(0.5f, h.y)), erf_abs, make_float2(0.5f, 0.5f));
    return __fmul2_rn(h, cdf);
}

// store value v of (this thread's voxel, channel k) into the hi / lo A buffers; ab[r] = swizzled byte offset of the
// voxel inside a K group for channel k % 4 = r, so that k's own part is a compile-time immediate
#define FZ_PUT_A(k, v)                                                                          \
    do {                                                                                        \
        const float v_ = (v);                                                                   \
        const uint32_t o_ = ab[(k) & 3] + (uint32_t)((k) >> 2) * 2048u;                         \
        *reinterpret_cast<float*>(a_hi + o_) = v_;                                              \
        *reinterpret_cast<float*>(a_lo + o_) = v_ - tf32_hi(v_);                                \
    } while (0)

template <int HID>
__global__ void __launch_bounds__(kTM, 2) mixer_mlp_fwd_tc(const float* __restrict__ x, const float* __restrict__ m,
                                                           const float* __restrict__ Wout, const float* __restrict__ bout,
                                                           const float* __restrict__ gamma, const float* __restrict__ beta,
                                                           const float* __restrict__ W1, const float* __restrict__ b1,
                                                           const float* __restrict__ W2, const float* __restrict__ b2,
                                                           float* __restrict__ x1_out, float* __restrict__ out,
                                                           long long vox, int tiles_per_sample, long long total_tiles, float eps) {
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* a_hi = smem;                       // 128 voxels x up to 64 channels
    unsigned char* a_lo = a_hi + kTM * kMaxHid * 4;
    unsigned char* wo_hi = a_lo + kTM * kMaxHid * 4;  // (32, 32)
    unsigned char* wo_lo = wo_hi + kC * kC * 4;
    unsigned char* w1_hi = wo_lo + kC * kC * 4;       // (HID, 32)
    unsigned char* w1_lo = w1_hi + kMaxHid * kC * 4;
    unsigned char* w2_hi = w1_lo + kMaxHid * kC * 4;  // (32, HID)
    unsigned char* w2_lo = w2_hi + kC * kMaxHid * 4;
    float* par = reinterpret_cast<float*>(w2_lo + kC * kMaxHid * 4);     // bout | b2 | gamma | beta | b1
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5;
    const bool issuer = uniform_u32((uint32_t)warp) == 0;      // warp-uniform: warp 0 issues the MMAs

    for (int e = tid; e < kC * kC; e += kTM) {
        const int n = e / kC, k = e % kC;
        const float v = Wout[e];
        *reinterpret_cast<float*>(wo_hi + b_off(n, k, kC)) = v;
        *reinterpret_cast<float*>(wo_lo + b_off(n, k, kC)) = v - tf32_hi(v);
    }
    for (int e = tid; e < HID * kC; e += kTM) {
        {
            const int n = e / kC, k = e % kC;            // W1 (HID, 32)
            const float v = W1[e];
            *reinterpret_cast<float*>(w1_hi + b_off(n, k, kC)) = v;
            *reinterpret_cast<float*>(w1_lo + b_off(n, k, kC)) = v - tf32_hi(v);
        }
        {
            const int n = e / HID, k = e % HID;          // W2 (32, HID)
            const float v = W2[e];
            *reinterpret_cast<float*>(w2_hi + b_off(n, k, HID)) = v;
            *reinterpret_cast<float*>(w2_lo + b_off(n, k, HID)) = v - tf32_hi(v);
        }
    }
    for (int c = tid; c < kC; c += kTM) {
        par[c] = bout ? bout[c] : 0.f; par[kC + c] = b2 ? b2[c] : 0.f;
        par[2 * kC + c] = gamma ? gamma[c] : 1.f; par[3 * kC + c] = beta ? beta[c] : 0.f;
    }
    for (int j = tid; j < HID; j += kTM) par[4 * kC + j] = b1 ? b1[j] : 0.f;
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;" :: "r"(smem_u32(&tmem_base)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    publish_and_sync();
    const uint32_t tmem = uniform_u32(tmem_base);
    const uint32_t lane_addr = tmem + ((uint32_t)(warp * 32) << 16);     // this thread's TMEM lane, column 0
    // TMEM columns: [0, 32) x1 projection | [32, 32 + HID) hidden | [96, 128) output projection
    uint32_t parity = 0;
    uint32_t ab[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) ab[r] = a_off(tid, r);

    // the first tile's m and x; every later tile's are fetched while the previous tile is in its GELU phase
    float mreg[kC], xreg[kC];
    {
        const long long tile = blockIdx.x;
        if (tile < total_tiles) {
            const long long b = tile / tiles_per_sample;
            const long long v0 = (tile - b * tiles_per_sample) * kTM + tid;
            const long long base = b * kC * vox + v0;
#pragma unroll
            for (int c = 0; c < kC; ++c) mreg[c] = v0 < vox ? __ldg(m + base + c * vox) : 0.f;
#pragma unroll
            for (int c = 0; c < kC; ++c) xreg[c] = v0 < vox ? __ldg(x + base + c * vox) : 0.f;
        }
    }
    for (long long tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const long long b = tile / tiles_per_sample;
        const long long v0 = (tile - b * tiles_per_sample) * kTM + tid;
        const bool valid = v0 < vox;
        const long long base = b * kC * vox + v0;
        // ---- A <- m ----
#pragma unroll
        for (int c = 0; c < kC; ++c) FZ_PUT_A(c, mreg[c]);
        publish_and_sync();
        if (issuer) {
            if (elect_one()) { gemm3(tmem, a_hi, a_lo, wo_hi, wo_lo, kC, kC); commit(&bar); }
            __syncwarp();
        }
        float x1[kC];
#pragma unroll
        for (int c = 0; c < kC; ++c) x1[c] = xreg[c];
        wait_bar(&bar, parity); parity ^= 1;
        {
            float d[32];
            tmem_ld32(lane_addr, d);
#pragma unroll
            for (int c = 0; c < kC; ++c) x1[c] += d[c] + par[c];
        }
        if (x1_out && valid) {
#pragma unroll
            for (int c = 0; c < kC; ++c) x1_out[base + c * vox] = x1[c];
        }
        // ---- A <- LN(x1) ----
        {
            float mean = 0.f;
#pragma unroll
            for (int c = 0; c < kC; ++c) mean += x1[c];
            mean *= (1.f / kC);
            float var = 0.f;
#pragma unroll
            for (int c = 0; c < kC; ++c) { const float dlt = x1[c] - mean; var = fmaf(dlt, dlt, var); }
            const float rstd = rsqrtf(var * (1.f / kC) + eps);
#pragma unroll
            for (int c = 0; c < kC; ++c) FZ_PUT_A(c, fmaf((x1[c] - mean) * rstd, par[2 * kC + c], par[3 * kC + c]));
        }
        publish_and_sync();
        if (issuer) {
            if (elect_one()) { gemm3(tmem + 32, a_hi, a_lo, w1_hi, w1_lo, kC, HID); commit(&bar); }
            __syncwarp();
        }
        // next tile's m and x: in flight during this tile's GELU and output phases
        {
            const long long nt = tile + gridDim.x;
            if (nt < total_tiles) {
                const long long nb = nt / tiles_per_sample;
                const long long nv = (nt - nb * tiles_per_sample) * kTM + tid;
                const long long nbase = nb * kC * vox + nv;
#pragma unroll
                for (int c = 0; c < kC; ++c) mreg[c] = nv < vox ? __ldg(m + nbase + c * vox) : 0.f;
#pragma unroll
                for (int c = 0; c < kC; ++c) xreg[c] = nv < vox ? __ldg(x + nbase + c * vox) : 0.f;
            }
        }
        wait_bar(&bar, parity); parity ^= 1;
        // ---- A <- gelu(h) ----
#pragma unroll
        for (int j0 = 0; j0 < HID; j0 += 32) {
            float h[32];
            tmem_ld32(lane_addr + 32 + j0, h);
#pragma unroll
            for (int j = 0; j < 32; j += 2) {
                const float2 bj = *reinterpret_cast<const float2*>(par + 4 * kC + j0 + j);
                const float2 g = gelu_exact2(make_float2(h[j] + bj.x, h[j + 1] + bj.y));
                FZ_PUT_A(j0 + j, g.x);
                FZ_PUT_A(j0 + j + 1, g.y);
            }
        }
        publish_and_sync();
        if (issuer) {
            if (elect_one()) { gemm3(tmem + 96, a_hi, a_lo, w2_hi, w2_lo, HID, kC); commit(&bar); }
            __syncwarp();
        }
        wait_bar(&bar, parity); parity ^= 1;
        {
            float d[32];
            tmem_ld32(lane_addr + 96, d);
            if (valid) {
#pragma unroll
                for (int c = 0; c < kC; ++c) out[base + c * vox] = x1[c] + d[c] + par[kC + c];
            }
        }
        // the next tile overwrites the A region and TMEM columns [0, 32): every thread is past its TMEM loads
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;" :: "r"(tmem) : "memory");
}

// =====================================================================================================================
// linear_bwd with the weight gradient on the tensor core:
//   da = W^T dy  [through LayerNorm, plus a residual gradient]  -- per voxel, FP32 pipe (1 024 FMA per voxel)
//   dW = sum_v dy n(a)^T                                        -- tcgen05: a contraction over voxels
// Operands of the contraction are K-major = voxel-contiguous rows, no swizzle, 8 x 16-byte core matrices 144 B apart
// along K (a voxel-owning warp's stores then hit 32 distinct banks).  A = [dy ; dy_lo] (64 rows: the smallest tcgen05 M,
// and exactly the two left factors of 3xTF32), B = n(a) and then n(a)_lo: two MMAs per 8 voxels give
//   rows 0..31  = dy (b_hi + b_lo),   rows 32..63 = dy_lo (b_hi + b_lo),   dW = their sum,
// accumulated in TMEM across ALL tiles of the CTA and flushed once at the end (row r sits in lane 32 (r/16) + r%16).
// Layout facts pinned by bench_probes/tcgen05_wgrad_probe.cu.
// =====================================================================================================================
constexpr int kKP = 144;                       // pitch of the core matrices along K (bytes)
constexpr int kKS = (kTM / 4) * kKP;           // 8 rows of 128 voxels: 4608 B
__device__ __forceinline__ uint32_t k_off(int row, int v) {
    return (uint32_t)((row >> 3) * kKS + (v >> 2) * kKP + (row & 7) * 16 + (v & 3) * 4);
}
__device__ __forceinline__ void mma_tf32_kk(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    mma_tf32(tmem_d, adesc, bdesc, idesc, accumulate);
}
__device__ __forceinline__ float warp_sum_tc(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
// transposing warp reduction of 32 values: returns the warp total of element `lane`
__device__ __forceinline__ float warp_vec_sum32(float (&v)[32], int lane) {
    int off = 16;
#pragma unroll
    for (int n = 32; n > 1; n >>= 1) {
        const bool up = (lane & off) != 0;
#pragma unroll
        for (int i = 0; i < n / 2; ++i) {
            const float keep = up ? v[i + n / 2] : v[i];
            const float send = up ? v[i] : v[i + n / 2];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
        }
        off >>= 1;
    }
    return v[0];
}

template <bool LN>
__global__ void __launch_bounds__(kTM, 2) linear_bwd_tc(const float* __restrict__ dy, const float* __restrict__ a,
                                                        const float* __restrict__ gamma, const float* __restrict__ beta,
                                                        const float* __restrict__ W, const float* __restrict__ resid,
                                                        float* __restrict__ da, float* __restrict__ dW, float* __restrict__ db,
                                                        float* __restrict__ dgamma, float* __restrict__ dbeta, long long vox,
                                                        int tiles_per_sample, long long total_tiles, float eps) {
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* KA = smem;                         // [dy ; dy_lo], 64 rows
    unsigned char* KBh = KA + 8 * kKS;                // n(a), 32 rows
    unsigned char* KBl = KBh + 4 * kKS;               // n(a)_lo
    float* Ws = reinterpret_cast<float*>(KBl + 4 * kKS);   // W [o][c] as stored
    float* gs = Ws + kC * kC;
    float* bs = gs + kC;
    float* accv = bs + kC;                            // db | dgamma | dbeta
    float* scr = accv + 3 * kC;                       // [warp][3 * kC]
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < kC * kC; i += kTM) Ws[i] = W[i];
    for (int c = tid; c < kC; c += kTM) { gs[c] = gamma ? gamma[c] : 1.f; bs[c] = beta ? beta[c] : 0.f; }
    for (int c = tid; c < 3 * kC; c += kTM) accv[c] = 0.f;
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 32;" :: "r"(smem_u32(&tmem_base)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    publish_and_sync();
    const uint32_t tmem = uniform_u32(tmem_base);
    const bool issuer = uniform_u32((uint32_t)warp) == 0;
    // D = F32, A = B = TF32, both K-major, N = 32, M = 64
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(kC >> 3) << 17) | ((uint32_t)(64 >> 4) << 24);
    uint32_t parity = 0;
    bool pending = false;          // MMAs of the previous tile still read KA / KB
    uint32_t kb[8];                // this voxel's byte offset inside a row group, per row % 8
#pragma unroll
    for (int r = 0; r < 8; ++r) kb[r] = k_off(r, tid);

    for (long long tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const long long b = tile / tiles_per_sample;
        const long long v0 = (tile - b * tiles_per_sample) * kTM + tid;
        const bool valid = v0 < vox;
        const long long base = b * kC * vox + v0;
        float g[kC], av[kC];
#pragma unroll
        for (int o = 0; o < kC; ++o) g[o] = valid ? __ldg(dy + base + o * vox) : 0.f;
#pragma unroll
        for (int c = 0; c < kC; ++c) av[c] = valid ? __ldg(a + base + c * vox) : 0.f;
        if (pending) { wait_bar(&bar, parity); parity ^= 1; pending = false; }
        // ---- stage [dy ; dy_lo] ----
#pragma unroll
        for (int o = 0; o < kC; ++o) {
            const uint32_t off = kb[o & 7] + (uint32_t)(o >> 3) * kKS;
            *reinterpret_cast<float*>(KA + off) = g[o];
            *reinterpret_cast<float*>(KA + off + 4 * kKS) = g[o] - tf32_hi(g[o]);
        }
        float r_db;
        {
            float sv[kC];
#pragma unroll
            for (int o = 0; o < kC; ++o) sv[o] = g[o];
            r_db = warp_vec_sum32(sv, lane);
        }
        // ---- da = W^T dy on the FP32 pipe: packed FFMA2 on output pairs, broadcast LDS.128 of the weights ----
        float2 d2[kC / 2];
#pragma unroll
        for (int k = 0; k < kC / 2; ++k) d2[k] = make_float2(0.f, 0.f);
#pragma unroll
        for (int o = 0; o < kC; ++o) {
            const float2 go = make_float2(g[o], g[o]);
#pragma unroll
            for (int q = 0; q < kC / 4; ++q) {
                const float4 w = *reinterpret_cast<const float4*>(Ws + o * kC + 4 * q);
                d2[2 * q] = __ffma2_rn(make_float2(w.x, w.y), go, d2[2 * q]);
                d2[2 * q + 1] = __ffma2_rn(make_float2(w.z, w.w), go, d2[2 * q + 1]);
            }
        }
        // ---- n(a): stage it, and run the LayerNorm backward ----
        float r_dg = 0.f, r_dbeta = 0.f;
        if (LN) {
            float mean = 0.f;
#pragma unroll
            for (int c = 0; c < kC; ++c) mean += av[c];
            mean *= (1.f / kC);
            float var = 0.f;
#pragma unroll
            for (int c = 0; c < kC; ++c) { av[c] -= mean; var = fmaf(av[c], av[c], var); }
            const float rstd = rsqrtf(var * (1.f / kC) + eps);
            float m1 = 0.f, m2 = 0.f;
            float sg[kC], sb[kC];
#pragma unroll
            for (int c = 0; c < kC; ++c) {
                av[c] *= rstd;                                            // a_hat
                const float n2 = fmaf(av[c], gs[c], bs[c]);
                const uint32_t off = kb[c & 7] + (uint32_t)(c >> 3) * kKS;
                *reinterpret_cast<float*>(KBh + off) = n2;
                *reinterpret_cast<float*>(KBl + off) = n2 - tf32_hi(n2);
                const float d = (c & 1) ? d2[c >> 1].y : d2[c >> 1].x;
                sg[c] = d * av[c];
                sb[c] = d;
                const float t = d * gs[c];
                m1 += t;
                m2 = fmaf(t, av[c], m2);
            }
            m1 *= (1.f / kC); m2 *= (1.f / kC);
            if (valid) {
#pragma unroll
                for (int c = 0; c < kC; ++c) {
                    const float d = (c & 1) ? d2[c >> 1].y : d2[c >> 1].x;
                    float o = rstd * (d * gs[c] - m1 - av[c] * m2);
                    if (resid) o += __ldg(resid + base + c * vox);
                    da[base + c * vox] = o;
                }
            }
            r_dg = warp_vec_sum32(sg, lane);
            r_dbeta = warp_vec_sum32(sb, lane);
        } else {
#pragma unroll
            for (int c = 0; c < kC; ++c) {
                const uint32_t off = kb[c & 7] + (uint32_t)(c >> 3) * kKS;
                *reinterpret_cast<float*>(KBh + off) = av[c];
                *reinterpret_cast<float*>(KBl + off) = av[c] - tf32_hi(av[c]);
                if (valid) da[base + c * vox] = (c & 1) ? d2[c >> 1].y : d2[c >> 1].x;
            }
        }
        scr[warp * 3 * kC + lane] = r_db;
        scr[warp * 3 * kC + kC + lane] = r_dg;
        scr[warp * 3 * kC + 2 * kC + lane] = r_dbeta;
        publish_and_sync();
        if (issuer) {
            const bool first = tile == (long long)blockIdx.x;
            if (elect_one()) {
#pragma unroll
                for (int s = 0; s < kTM / 8; ++s) {
                    const uint64_t ad = make_desc(smem_u32(KA) + s * 2 * kKP, kKP, kKS, 0);
                    const uint64_t bh = make_desc(smem_u32(KBh) + s * 2 * kKP, kKP, kKS, 0);
                    const uint64_t bl = make_desc(smem_u32(KBl) + s * 2 * kKP, kKP, kKS, 0);
                    mma_tf32_kk(tmem, ad, bh, idesc, !(first && s == 0));
                    mma_tf32_kk(tmem, ad, bl, idesc, 1);
                }
                commit(&bar);
            }
            __syncwarp();
        }
        pending = true;
        if (tid < 3 * kC) {
            float t = 0.f;
#pragma unroll
            for (int w = 0; w < kTM / 32; ++w) t += scr[w * 3 * kC + tid];
            accv[tid] += t;
        }
        __syncthreads();            // scr is rewritten by the next tile
    }
    if (pending) { wait_bar(&bar, parity); parity ^= 1; }
    // ---- flush: row r of the accumulator sits in TMEM lane 32 (r / 16) + r % 16; rows r and 32 + r add up to dW[r] ----
    if (blockIdx.x < total_tiles) {
        float d[32];
        tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16), d);
        if (lane < 16) {
            const int o = (warp * 16 + lane) & 31;
#pragma unroll
            for (int c = 0; c < kC; ++c) atomicAdd(dW + o * kC + c, d[c]);
        }
    }
    for (int c = tid; c < kC; c += kTM) {
        if (db) atomicAdd(db + c, accv[c]);
        if (LN && dgamma) atomicAdd(dgamma + c, accv[kC + c]);
        if (LN && dbeta) atomicAdd(dbeta + c, accv[2 * kC + c]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 32;" :: "r"(tmem) : "memory");
}

size_t lbw_tc_smem_bytes() { return (size_t)16 * kKS + (kC * kC + 2 * kC + 3 * kC + (kTM / 32) * 3 * kC) * 4 + 1024; }

size_t tc_smem_bytes() { return (size_t)2 * kTM * kMaxHid * 4 + 2 * kC * kC * 4 + 4 * kMaxHid * kC * 4 + (4 * kC + kMaxHid) * 4 + 1024; }

int tc_sm_count() { return num_sms(); }

}  // namespace

// hidden width 32 or 64 (mlp_ratio 1 or 2 at 32 channels): TMEM columns / shared memory of this version
bool mixer_mlp_tc_supported_v1(int hidden) { return hidden == 32 || hidden == 64; }

int mixer_mlp_tc_launch_v1(const float* x, const float* m, const float* Wout, const float* bout, const float* gamma, const float* beta,
                        const float* W1, const float* b1, const float* W2, const float* b2, float* x1, float* out, long long batch,
                        int hidden, long long voxels, float eps, cudaStream_t st) {
    const size_t smem = tc_smem_bytes();
    static SmemConfig cfg32, cfg64;
    FZ_CUDA_CHECK(cfg32.ensure(mixer_mlp_fwd_tc<32>, smem));
    FZ_CUDA_CHECK(cfg64.ensure(mixer_mlp_fwd_tc<64>, smem));
    const int tps = (int)((voxels + kTM - 1) / kTM);
    const long long tiles = batch * tps;
    const long long cap = 2LL * tc_sm_count();
    const unsigned blocks = (unsigned)(tiles < cap ? tiles : cap);
    if (hidden == 64)
        mixer_mlp_fwd_tc<64><<<blocks, kTM, smem, st>>>(x, m, Wout, bout, gamma, beta, W1, b1, W2, b2, x1, out, voxels, tps, tiles, eps);
    else
        mixer_mlp_fwd_tc<32><<<blocks, kTM, smem, st>>>(x, m, Wout, bout, gamma, beta, W1, b1, W2, b2, x1, out, voxels, tps, tiles, eps);
    FZ_LAUNCH_CHECK();
    return FZ_OK;
}

int linear_bwd_tc_launch(const float* dy, const float* a, const float* gamma, const float* beta, const float* W, const float* resid,
                         float* da, float* dW, float* db, float* dgamma, float* dbeta, long long batch, long long voxels, float eps,
                         int layernorm, cudaStream_t st) {
    const size_t smem = lbw_tc_smem_bytes();
    static SmemConfig cfg_ln, cfg_plain;
    FZ_CUDA_CHECK(cfg_ln.ensure(linear_bwd_tc<true>, smem));
    FZ_CUDA_CHECK(cfg_plain.ensure(linear_bwd_tc<false>, smem));
    const int tps = (int)((voxels + kTM - 1) / kTM);
    const long long tiles = batch * tps;
    const long long cap = 2LL * tc_sm_count();
    const unsigned blocks = (unsigned)(tiles < cap ? tiles : cap);
    if (layernorm)
        linear_bwd_tc<true><<<blocks, kTM, smem, st>>>(dy, a, gamma, beta, W, resid, da, dW, db, dgamma, dbeta, voxels, tps, tiles, eps);
    else
        linear_bwd_tc<false><<<blocks, kTM, smem, st>>>(dy, a, nullptr, nullptr, W, nullptr, da, dW, db, nullptr, nullptr, voxels, tps, tiles, eps);
    FZ_LAUNCH_CHECK();
    return FZ_OK;
}

}  // namespace fz
