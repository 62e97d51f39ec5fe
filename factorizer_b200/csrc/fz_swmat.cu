// Standalone SWMatricize forward / inverse_forward and their adjoints: pure index math, bit-exact
// against the reference (factorizer/factorization/operations.py:266-280, 417-434).
//
// Both directions are single-pass HBM-bound copies.  The gather kernel assigns one CTA per window
// (all d rows): the column->voxel offset is computed once per column and reused for every channel
// row; writes are fully coalesced along the matrix row.  The scatter kernel assigns one thread per
// output voxel element and visits the S window sets in order, so the floating-point summation order
// is the reference's ((0.0 + inv_0) + inv_1 + ...) / S.
#include "fz_common.cuh"

namespace fz {

// y[(s*BH + b*heads + h), w, dd, j] = x[b, h*d+dd, roll_s(window w, column j)]  (optionally / S)
__global__ void __launch_bounds__(256) swmat_gather_kernel(const float* __restrict__ x,
                                                           float* __restrict__ y, DevGeom G,
                                                           int divide) {
    const long long win = blockIdx.x;  // over S * B * heads * Gwin
    const int w = (int)(win % G.G);
    long long t = win / G.G;
    const int h = (int)(t % G.heads);
    t /= G.heads;
    const int b = (int)(t % G.B);
    const int s = (int)(t / G.B);
    const float* xb = x + ((long long)b * G.C + (long long)h * G.d) * G.vox;
    float* yb = y + win * (long long)G.d * G.P;
    const float fS = (float)G.S;
    for (int j = threadIdx.x; j < G.P; j += blockDim.x) {
        const long long off = window_col_offset(G, s, w, j);
        for (int dd = 0; dd < G.d; ++dd) {
            float v = __ldg(xb + (long long)dd * G.vox + off);
            if (divide) v = __fdiv_rn(v, fS);
            yb[(long long)dd * G.P + j] = v;
        }
    }
}

// out[b, c, i] = (((0.0 +) y_0[...]) + y_1[...] + ...) (/ S)
__global__ void __launch_bounds__(256) swmat_scatter_kernel(const float* __restrict__ y,
                                                            float* __restrict__ out, DevGeom G,
                                                            int reference_inverse) {
    const long long total = (long long)G.B * G.C * G.vox;
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long matsz = (long long)G.d * G.P;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += stride) {
        long long t = idx;
        const int i2 = (int)(t % G.n[2]); t /= G.n[2];
        const int i1 = (int)(t % G.n[1]); t /= G.n[1];
        const int i0 = (int)(t % G.n[0]); t /= G.n[0];
        const int c = (int)(t % G.C);
        const int b = (int)(t / G.C);
        const int h = c / G.d, dd = c % G.d;
        float acc = 0.0f;
        for (int s = 0; s < G.S; ++s) {
            // rolled coordinate r = (i + shift) mod n; window g = r / p, in-window q = r % p
            int r0 = (i0 + G.sh[s][0]) % G.n[0]; if (r0 < 0) r0 += G.n[0];
            int r1 = (i1 + G.sh[s][1]) % G.n[1]; if (r1 < 0) r1 += G.n[1];
            int r2 = (i2 + G.sh[s][2]) % G.n[2]; if (r2 < 0) r2 += G.n[2];
            const int w = ((r0 / G.p[0]) * G.g[1] + (r1 / G.p[1])) * G.g[2] + (r2 / G.p[2]);
            const int j = ((r0 % G.p[0]) * G.p[1] + (r1 % G.p[1])) * G.p[2] + (r2 % G.p[2]);
            const long long row = ((long long)s * G.B + b) * G.heads + h;
            const float v = __ldg(y + (row * G.G + w) * matsz + (long long)dd * G.P + j);
            // operations.py:426-431: out = 0.0; out = out + inv_s  (0.0 + v keeps -0.0 -> +0.0)
            acc = (s == 0 && !reference_inverse) ? v : __fadd_rn(acc, v);
        }
        if (reference_inverse) acc = __fdiv_rn(acc, (float)G.S);  // operations.py:433
        out[idx] = acc;
    }
}

static int launch_gather(const float* x, float* y, const fz_geom* g, int divide, cudaStream_t st) {
    DevGeom G;
    if (int e = make_dev_geom(g, &G)) return e;
    if (!x || !y) return fail(FZ_ERR_INVALID, "null buffer");
    long long wins = (long long)G.S * G.mats_per_shift;
    if (wins == 0) return FZ_OK;
    if (wins > 2147483647LL) return fail(FZ_ERR_UNSUPPORTED, "too many windows (%lld)", wins);
    swmat_gather_kernel<<<(unsigned)wins, 256, 0, st>>>(x, y, G, divide);
    FZ_LAUNCH_CHECK();
    return FZ_OK;
}

static int launch_scatter(const float* y, float* out, const fz_geom* g, int ref, cudaStream_t st) {
    DevGeom G;
    if (int e = make_dev_geom(g, &G)) return e;
    if (!y || !out) return fail(FZ_ERR_INVALID, "null buffer");
    long long total = (long long)G.B * G.C * G.vox;
    if (total == 0) return FZ_OK;
    long long blocks = (total + 255) / 256;
    if (blocks > 148LL * 64) blocks = 148LL * 64;
    swmat_scatter_kernel<<<(unsigned)blocks, 256, 0, st>>>(y, out, G, ref);
    FZ_LAUNCH_CHECK();
    return FZ_OK;
}

}  // namespace fz

extern "C" {

int fz_swmat_forward(const float* x, float* y, const fz_geom* g, void* stream) {
    fz::tls().launches = 0;
    return fz::launch_gather(x, y, g, 0, (cudaStream_t)stream);
}

int fz_swmat_inverse(const float* y, float* x_out, const fz_geom* g, void* stream) {
    fz::tls().launches = 0;
    return fz::launch_scatter(y, x_out, g, 1, (cudaStream_t)stream);
}

int fz_swmat_forward_adjoint(const float* gy, float* gx, const fz_geom* g, void* stream) {
    fz::tls().launches = 0;
    return fz::launch_scatter(gy, gx, g, 0, (cudaStream_t)stream);
}

int fz_swmat_inverse_adjoint(const float* g_out, float* gy, const fz_geom* g, void* stream) {
    fz::tls().launches = 0;
    return fz::launch_gather(g_out, gy, g, 1, (cudaStream_t)stream);
}

}  // extern "C"
