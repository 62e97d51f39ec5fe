// tcgen05 / TMEM building blocks shared by the tensor-core glue kernels (sm_100a): shared-memory matrix descriptors,
// kind::tf32 MMA issue, commit / mbarrier waits, TMEM loads, the 3xTF32 split, and the exact-erf GELU.
//
// Operand forms used by these kernels (all pinned on B200 by bench_probes/tcgen05_*probe.cu; kind::tf32):
//   * A in TENSOR MEMORY (per-voxel GEMMs, M = 128 voxels = TMEM lanes): column c of lane v = A(v, k = c); every thread
//     writes its own voxel's row with tcgen05.st, no shared memory, no swizzle.  The tensor core reads the top 19 bits.
//   * MN-major, SWIZZLE_128B_BASE32B (descriptor layout type 1; the only swizzle 32-bit MN-major operands accept, and
//     K-major operands reject it): element (mn, k) at
//         LBO (mn / 32) + SBO (k / 4) + 128 (k % 4) + 32 (((mn % 32) / 8) ^ (k % 4)) + 4 (mn % 8).
//     For a contraction over voxels (k = voxel) with SBO = 512 this is "voxel rows": the 32 values of a voxel are one
//     128-byte row whose 32-byte chunks are XOR-ed with (voxel % 4) -- a thread stores its voxel's row with STS.128.
//   * K-major without swizzle (layout type 0; the weights): 8 x 16-byte core matrices, LBO along K, SBO along M / N.
#pragma once
#include "fz_common.cuh"

namespace fz {
namespace tc {

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
// the same matrix `bytes` further on in shared memory (the 14-bit start field cannot overflow below 256 KiB)
__device__ __forceinline__ uint64_t desc_at(uint64_t desc, uint32_t bytes) { return desc + (uint64_t)(bytes >> 4); }

// instruction descriptor: D = F32, A = B = TF32
__device__ __forceinline__ uint32_t make_idesc(int m, int n, bool a_mn_major, bool b_mn_major) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((a_mn_major ? 1u : 0u) << 15) | ((b_mn_major ? 1u : 0u) << 16) |
           ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
        :: "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// UTCHMMA takes its operands from UNIFORM registers.  Issued under `if (threadIdx.x == 0)` the compiler wraps every MMA in
// an ELECT / R2UR.BROADCAST / BRA.U.ANY loop (~70 cycles per MMA, measured); issued from warp-uniform control flow on
// warp-uniform values (uniform_u32) under elect_one() it is a single predicated instruction.
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(pred));
    return pred != 0;
}
// tells the compiler that a value is the same in every lane of the (converged) warp
__device__ __forceinline__ uint32_t uniform_u32(uint32_t v) { return __shfl_sync(0xffffffffu, v, 0); }

__device__ __forceinline__ void commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(bar) : "memory");
}
__device__ __forceinline__ void bar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count));
}
__device__ __forceinline__ void bar_arrive(uint32_t bar) {
    asm volatile("{\n.reg .b64 st;\nmbarrier.arrive.shared::cta.b64 st, [%0];\n}" :: "r"(bar) : "memory");
}
__device__ __forceinline__ void bar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok = 0;
    while (!ok) {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    }
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// 32 / 16 consecutive TMEM columns of this thread's lane
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
__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// 3xTF32: the tensor core reads the top 19 bits of an fp32 word, so the word itself is the "hi" operand and
// lo = x - hi (exact, representable in TF32 up to its own rounding) the correction
__device__ __forceinline__ float tf32_hi(float v) { return __uint_as_float(__float_as_uint(v) & 0xffffe000u); }
__device__ __forceinline__ float tf32_lo(float v) { return v - tf32_hi(v); }

// K-major, no swizzle: element (n, k) of a matrix with KC columns (SBO = 32 KC, LBO = 128, 256 bytes per K step of 8)
__device__ __forceinline__ uint32_t kmajor_off(int n, int k, int KC) {
    return (uint32_t)((n >> 3) * (KC * 32) + (k >> 2) * 128 + (n & 7) * 16 + (k & 3) * 4);
}

__device__ __forceinline__ void mma_tf32_ta(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n"
        :: "r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// 16 / 32 consecutive TMEM columns of this thread's lane <- registers
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        :: "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
           "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 :: "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// One 32-byte chunk (8 consecutive values) of a thread's voxel row, as two STS.128 (forced: left to itself the compiler
// turns the select into predicated scalar stores, which conflict 8-way).  `chunk`: shared-memory address of the chunk;
// lanes whose voxel has bit 2 set (swap) store the upper half first, so that a quarter-warp's STS.128 covers all 32 banks.
__device__ __forceinline__ void st_row_chunk(uint32_t chunk, bool swap, const float (&a)[8]) {
    const uint32_t first = chunk + (swap ? 16u : 0u), second = chunk + (swap ? 0u : 16u);
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};"
                 :: "r"(first), "f"(swap ? a[4] : a[0]), "f"(swap ? a[5] : a[1]), "f"(swap ? a[6] : a[2]), "f"(swap ? a[7] : a[3]) : "memory");
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};"
                 :: "r"(second), "f"(swap ? a[0] : a[4]), "f"(swap ? a[1] : a[5]), "f"(swap ? a[2] : a[6]), "f"(swap ? a[3] : a[7]) : "memory");
}

// Sums of two values over the TWO threads that share a voxel (warps w and w + 4 of a 256-thread CTA: same lane, same TMEM
// lane quarter).  `slots`: 2 x 2 x 128 float2 of shared memory ([call parity][half][voxel]); `turn` counts the calls (two
// buffers: the partner is at most one exchange behind); named barrier 1 + quarter, 64 threads.
__device__ __forceinline__ float2 pair_sum2(float2 mine, float2* slots, uint32_t& turn, int hh, int vq, int v) {
    float2* s = slots + (turn & 1u) * 256;
    s[hh * 128 + v] = mine;
    asm volatile("bar.sync %0, 64;" :: "r"(1 + vq) : "memory");
    const float2 other = s[(hh ^ 1) * 128 + v];
    ++turn;
    return make_float2(mine.x + other.x, mine.y + other.y);
}

__device__ __forceinline__ float ex2_approx(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcp_approx(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

// Exact (erf) GELU through Abramowitz-Stegun 7.1.26 (|error| <= 1.5e-7; the formula of csrc/fz_block_glue.cu), two values
// at a time: Phi(h) and e = exp(-h^2 / 2)
__device__ __forceinline__ float2 gauss_cdf2(float2 h, float2& e) {
    const float2 one = make_float2(1.f, 1.f);
    const float2 z = make_float2(fabsf(h.x) * 0.70710678118654752f, fabsf(h.y) * 0.70710678118654752f);
    const float2 d = __ffma2_rn(make_float2(0.3275911f, 0.3275911f), z, one);
    const float2 t = make_float2(rcp_approx(d.x), rcp_approx(d.y));
    const float2 hh = __fmul2_rn(h, h);
    e = make_float2(ex2_approx(hh.x * -0.72134752044448170f), ex2_approx(hh.y * -0.72134752044448170f));
    float2 p = __ffma2_rn(t, make_float2(1.061405429f, 1.061405429f), make_float2(-1.453152027f, -1.453152027f));
    p = __ffma2_rn(t, p, make_float2(1.421413741f, 1.421413741f));
    p = __ffma2_rn(t, p, make_float2(-0.284496736f, -0.284496736f));
    p = __ffma2_rn(t, p, make_float2(0.254829592f, 0.254829592f));
    const float2 pt = __fmul2_rn(p, t);
    const float2 erf_abs = __ffma2_rn(make_float2(-pt.x, -pt.y), e, one);
    return __ffma2_rn(make_float2(copysignf(0.5f, h.x), copysignf(0.5f, h.y)), erf_abs, make_float2(0.5f, 0.5f));
}
// gelu(h) and its derivative Phi(h) + h phi(h)  (torch GeluBackward, approximate='none')
__device__ __forceinline__ void gelu_grad2(float2 h, float2& g, float2& gp) {
    float2 e;
    const float2 cdf = gauss_cdf2(h, e);
    g = __fmul2_rn(h, cdf);
    gp = __ffma2_rn(__fmul2_rn(h, make_float2(0.39894228040143268f, 0.39894228040143268f)), e, cdf);
}

}  // namespace tc
}  // namespace fz
