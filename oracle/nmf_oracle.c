/* CPU ORACLE -- TEST INFRASTRUCTURE AND CPU BASELINE ONLY (never linked into the product).
 *
 * Plain-C restatement of the production configuration of the reference's FactMixer core
 * (factorizer/factorizer.py:41-50): SWMatricize gather (operations.py:266-272, 321-325, 417-421),
 * optional ReLU (:44), rank-1 HALS NMF unrolled T times (matrix_factorization.py:224-227 applied to
 * x and x.mT, :122-136, :514-533) and the averaged inverse scatter (operations.py:423-434), plus
 * the hand-derived adjoint for dL/dx.  Windows are independent, so the loop
 * over windows is an OpenMP parallel-for: this is the "reference CPU path with all host threads"
 * the benchmark times next to the GPU numbers (`bench.py --impl reference`, cpu_baseline.kind="port").
 *
 * Parity status: pinned by tests/test_oracle.py against the golden vectors generated from the
 * reference (tests/golden/fused.npz) and against the numpy oracle.
 *
 * Geometry: 3 spatial dims, arbitrary patch p[3] and head_dim d (matrix d x P, P = p0*p1*p2),
 * S window sets with integer shifts.  Arithmetic: float32, same operation order as the reference's
 * formulas (a = X v, b = v.v, u = relu((a+eps)/(b+eps)), c = X^T u, dd = u.u, v = relu((c+eps)/(dd+eps))).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct {
    int B, C, n[3], p[3], d, S;
    int sh[8][3];
} fzo_geom;

static inline int wrap(int i, int n) { i %= n; return i < 0 ? i + n : i; }

/* offsets (within one channel) of the P columns of window w under shift set s */
static void window_offsets(const fzo_geom* g, int s, int w, int* off) {
    const int G1 = g->n[1] / g->p[1], G2 = g->n[2] / g->p[2];
    const int g2 = w % G2, g1 = (w / G2) % G1, g0 = w / (G2 * G1);
    int j = 0;
    for (int q0 = 0; q0 < g->p[0]; ++q0) {
        const int i0 = wrap(g0 * g->p[0] + q0 - g->sh[s][0], g->n[0]);
        for (int q1 = 0; q1 < g->p[1]; ++q1) {
            const int i1 = wrap(g1 * g->p[1] + q1 - g->sh[s][1], g->n[1]);
            for (int q2 = 0; q2 < g->p[2]; ++q2) {
                const int i2 = wrap(g2 * g->p[2] + q2 - g->sh[s][2], g->n[2]);
                off[j++] = (i0 * g->n[1] + i1) * g->n[2] + i2;
            }
        }
    }
}

/* torchrun exports OMP_NUM_THREADS=1 to every rank; the benchmark's CPU legs ask for all host cores explicitly */
void fzo_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

int fzo_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* forward iterates for one matrix X (d x P, row-major): fills u_hist[T][d], v_hist[(T+1)][P] (v_hist[0] = v0),
 * b_hist[T], dd_hist[T] */
static void hals1_iterate(const float* X, int d, int P, const float* v0, int T, float eps,
                          float* u_hist, float* v_hist, float* b_hist, float* dd_hist) {
    memcpy(v_hist, v0, sizeof(float) * P);
    for (int t = 0; t < T; ++t) {
        const float* v = v_hist + (size_t)t * P;
        float* vn = v_hist + (size_t)(t + 1) * P;
        float* u = u_hist + (size_t)t * d;
        float b = 0.f;
        for (int j = 0; j < P; ++j) b += v[j] * v[j];
        float dd = 0.f;
        for (int i = 0; i < d; ++i) {
            float a = 0.f;
            const float* xr = X + (size_t)i * P;
            for (int j = 0; j < P; ++j) a += xr[j] * v[j];
            const float pre = (a + eps) / (b + eps);
            u[i] = pre > 0.f ? pre : 0.f;
            dd += u[i] * u[i];
        }
        for (int j = 0; j < P; ++j) vn[j] = 0.f;
        for (int i = 0; i < d; ++i) {
            const float* xr = X + (size_t)i * P;
            const float ui = u[i];
            for (int j = 0; j < P; ++j) vn[j] += xr[j] * ui;
        }
        for (int j = 0; j < P; ++j) {
            const float pre = (vn[j] + eps) / (dd + eps);
            vn[j] = pre > 0.f ? pre : 0.f;
        }
        b_hist[t] = b;
        dd_hist[t] = dd;
    }
}

/* y = inverse(NMF(act(matricize(x))));  x, y: (B, C, n0, n1, n2) */
void fzo_swnmf_forward(const float* x, const float* v0, float* y, const fzo_geom* g, int relu, int T, float eps) {
    const int d = g->d, P = g->p[0] * g->p[1] * g->p[2], heads = g->C / d;
    const int G = (g->n[0] / g->p[0]) * (g->n[1] / g->p[1]) * (g->n[2] / g->p[2]);
    const long long vox = (long long)g->n[0] * g->n[1] * g->n[2];
    const long long nwin = (long long)g->B * heads * G;
    for (int s = 0; s < g->S; ++s) {
#pragma omp parallel
        {
            int* off = (int*)malloc(sizeof(int) * P);
            float* X = (float*)malloc(sizeof(float) * d * P);
            float* uh = (float*)malloc(sizeof(float) * (T > 0 ? T : 1) * d);
            float* vh = (float*)malloc(sizeof(float) * (T + 1) * P);
            float* bh = (float*)malloc(sizeof(float) * (T > 0 ? T : 1) * 2);
#pragma omp for schedule(static)
            for (long long m = 0; m < nwin; ++m) {
                const int w = (int)(m % G), h = (int)((m / G) % heads), b = (int)(m / ((long long)G * heads));
                window_offsets(g, s, w, off);
                const float* xb = x + ((long long)b * g->C + (long long)h * d) * vox;
                float* yb = y + ((long long)b * g->C + (long long)h * d) * vox;
                for (int i = 0; i < d; ++i)
                    for (int j = 0; j < P; ++j) {
                        float v = xb[(long long)i * vox + off[j]];
                        X[(size_t)i * P + j] = (relu && v < 0.f) ? 0.f : v;
                    }
                hals1_iterate(X, d, P, v0, T, eps, uh, vh, bh, bh + T);
                const float* u = uh + (size_t)(T - 1) * d;
                const float* v = vh + (size_t)T * P;
                for (int i = 0; i < d; ++i)
                    for (int j = 0; j < P; ++j) {
                        float* dst = yb + (long long)i * vox + off[j];
                        float val = u[i] * v[j];
                        if (s > 0) val = *dst + val;                 /* out = out + inv_s, operations.py:431 */
                        if (s == g->S - 1) val = val / (float)g->S;  /* out / num_shifts, operations.py:433 */
                        *dst = val;
                    }
            }
            free(off); free(X); free(uh); free(vh); free(bh);
        }
    }
}

/* gx = d<gy, y>/dx  (adjoint of fzo_swnmf_forward), K = number of differentiated sweeps */
void fzo_swnmf_backward(const float* x, const float* gy, const float* v0, float* gx, const fzo_geom* g,
                        int relu, int T, int K, float eps) {
    const int d = g->d, P = g->p[0] * g->p[1] * g->p[2], heads = g->C / d;
    const int G = (g->n[0] / g->p[0]) * (g->n[1] / g->p[1]) * (g->n[2] / g->p[2]);
    const long long vox = (long long)g->n[0] * g->n[1] * g->n[2];
    const long long nwin = (long long)g->B * heads * G;
    if (K < 0 || K > T) K = T;
    for (int s = 0; s < g->S; ++s) {
#pragma omp parallel
        {
            int* off = (int*)malloc(sizeof(int) * P);
            float* X = (float*)malloc(sizeof(float) * d * P);
            float* Gm = (float*)malloc(sizeof(float) * d * P);
            float* Xb = (float*)malloc(sizeof(float) * d * P);
            float* uh = (float*)malloc(sizeof(float) * T * d);
            float* vh = (float*)malloc(sizeof(float) * (T + 1) * P);
            float* bh = (float*)malloc(sizeof(float) * T * 2);
            float* vbar = (float*)malloc(sizeof(float) * P);
            float* cbar = (float*)malloc(sizeof(float) * P);
            float* ubar = (float*)malloc(sizeof(float) * d);
            float* abar = (float*)malloc(sizeof(float) * d);
#pragma omp for schedule(static)
            for (long long m = 0; m < nwin; ++m) {
                const int w = (int)(m % G), h = (int)((m / G) % heads), b = (int)(m / ((long long)G * heads));
                window_offsets(g, s, w, off);
                const long long base = ((long long)b * g->C + (long long)h * d) * vox;
                for (int i = 0; i < d; ++i)
                    for (int j = 0; j < P; ++j) {
                        float v = x[base + (long long)i * vox + off[j]];
                        X[(size_t)i * P + j] = (relu && v < 0.f) ? 0.f : v;
                        Gm[(size_t)i * P + j] = gy[base + (long long)i * vox + off[j]] / (float)g->S;
                        Xb[(size_t)i * P + j] = 0.f;
                    }
                hals1_iterate(X, d, P, v0, T, eps, uh, vh, bh, bh + T);
                const float* uT = uh + (size_t)(T - 1) * d;
                const float* vT = vh + (size_t)T * P;
                for (int j = 0; j < P; ++j) vbar[j] = 0.f;
                for (int i = 0; i < d; ++i) {
                    float acc = 0.f;
                    for (int j = 0; j < P; ++j) {
                        acc += Gm[(size_t)i * P + j] * vT[j];
                        vbar[j] += Gm[(size_t)i * P + j] * uT[i];
                    }
                    ubar[i] = acc;
                }
                for (int t = T - 1; t >= T - K; --t) {
                    const float* u = uh + (size_t)t * d;
                    const float* vt = vh + (size_t)(t + 1) * P;
                    const float* vp = vh + (size_t)t * P;
                    const float bt = bh[t] + eps, dt = bh[T + t] + eps;
                    float e = 0.f;
                    for (int j = 0; j < P; ++j) {
                        const float qb = vt[j] > 0.f ? vbar[j] : 0.f;
                        cbar[j] = qb / dt;
                        e += qb * vt[j];
                    }
                    const float dbar = -e / dt;
                    float bacc = 0.f;
                    for (int i = 0; i < d; ++i) {
                        float wsum = 0.f;
                        const float* xr = X + (size_t)i * P;
                        float* xbr = Xb + (size_t)i * P;
                        for (int j = 0; j < P; ++j) { xbr[j] += u[i] * cbar[j]; wsum += xr[j] * cbar[j]; }
                        float ub = wsum + 2.f * dbar * u[i];
                        if (t == T - 1) ub += ubar[i];
                        const float pb = u[i] > 0.f ? ub : 0.f;
                        abar[i] = pb / bt;
                        bacc += pb * u[i];
                    }
                    const float bbar = -bacc / bt;
                    for (int j = 0; j < P; ++j) vbar[j] = 2.f * bbar * vp[j];
                    for (int i = 0; i < d; ++i) {
                        const float* xr = X + (size_t)i * P;
                        float* xbr = Xb + (size_t)i * P;
                        const float ai = abar[i];
                        for (int j = 0; j < P; ++j) { xbr[j] += ai * vp[j]; vbar[j] += xr[j] * ai; }
                    }
                }
                for (int i = 0; i < d; ++i)
                    for (int j = 0; j < P; ++j) {
                        float val = Xb[(size_t)i * P + j];
                        if (relu && !(X[(size_t)i * P + j] > 0.f)) val = 0.f;
                        float* dst = gx + base + (long long)i * vox + off[j];
                        if (s > 0) val += *dst;
                        *dst = val;
                    }
            }
            free(off); free(X); free(Gm); free(Xb); free(uh); free(vh); free(bh);
            free(vbar); free(cbar); free(ubar); free(abar);
        }
    }
}
