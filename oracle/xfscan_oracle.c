/*
 * xfscan_oracle.c -- CPU restatement of XFMamba's SS2D scan path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, the smoke() entry and the
 * cpu_baseline / --impl reference legs of bench.py may load it.  The product path
 * (xfmamba_b200/) never links, imports or falls back to anything in oracle/.
 *
 * Every function restates one reference function (paths relative to the reference tree):
 *   xfo_cross_scan        models/csm_triton.py:22-53    (cross_scan_fwd, channel-first)
 *   xfo_cross_merge       models/csm_triton.py:56-85    (cross_merge_fwd, add order of :61-62)
 *   xfo_swap_scan         models/fusion_vmamba.py:189-215 (SwappingScan_multiview.forward)
 *   xfo_selective_scan_*  models/csms6s.py:25-68        (selective_scan_torch) and, for the
 *                         gradients, the closed forms the native backward uses
 *                         (models/selective_scan/csrc/selective_scan/selective_scan_bwd_kernel.cuh:204-273)
 *   xfo_ss2d_fwd          models/fusion_vmamba.py:1145-1174 (cross_scan -> scan -> cross_merge)
 *
 * Parity pin: the .npz files under tests/golden/ were produced by importing the reference's own Python
 * (tests/golden/make_golden.py, run in the build container where /root/reference is mounted);
 * tests/test_oracle_golden.py checks every function here against them.
 *
 * Arithmetic: the forward scan is evaluated in the real type `real_t` below.  It is built
 * twice (float = the reference's fp32 arithmetic, double = tie-breaker "truth"); symbols get the
 * suffix _f32 / _f64.  Compile with -ffp-contract=off so that, like ATen on CPU, products and
 * sums are rounded separately.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifndef XFO_REAL
#define XFO_REAL float
#define XFO_SUFFIX _f32
#endif
typedef XFO_REAL real_t;
#define XFO_CAT2(a, b) a##b
#define XFO_CAT(a, b) XFO_CAT2(a, b)
#define XFO_NAME(n) XFO_CAT(n, XFO_SUFFIX)

/* ---------------------------------------------------------------------------------------------
 * Index routes (dtype-agnostic byte moves; built only once, in the float build).
 * scans: 0 = cross2d (4 routes), 1 = unidirectional (4 copies), 2 = bidirectional (fwd,fwd,rev,rev)
 * models/csm_triton.py:25-35
 * ------------------------------------------------------------------------------------------- */
#ifdef XFO_BUILD_ROUTES
/* thread count of the OpenMP loops (bench.py's CPU baseline): a launcher may have exported OMP_NUM_THREADS=1
 * (torch.distributed.run does) long before this library is loaded, so the count is set explicitly */
#ifdef _OPENMP
#include <omp.h>
int xfo_set_threads(int n) { if (n > 0) omp_set_num_threads(n); return omp_get_max_threads(); }
#else
int xfo_set_threads(int n) { (void)n; return 1; }
#endif

/* spatial offset (h*W+w) that scan position l of direction k reads */
static inline int64_t xfo_route(int k, int64_t l, int64_t H, int64_t W, int scans) {
    int64_t L = H * W;
    if (scans == 1) return l;
    if (scans == 2) return (k < 2) ? l : (L - 1 - l);
    /* scans == 0 */
    int64_t lf = (k >= 2) ? (L - 1 - l) : l;       /* k=2,3 are flips of k=0,1 */
    if ((k & 1) == 0) return lf;                    /* row-major */
    int64_t w = lf / H, h = lf % H;                 /* k odd: column-major walk (x.transpose(2,3).flatten) */
    return h * W + w;
}

/* x: (B, C, H, W) [one_by_one: (B, 4, C, H, W)] -> xs: (B, 4, C, H*W); esize = bytes per element */
void xfo_cross_scan(const void* x, void* xs, int64_t B, int64_t C, int64_t H, int64_t W,
                    int esize, int scans, int one_by_one) {
    const int64_t L = H * W;
    const char* src = (const char*)x;
    char* dst = (char*)xs;
    for (int64_t b = 0; b < B; ++b)
        for (int k = 0; k < 4; ++k)
            for (int64_t c = 0; c < C; ++c) {
                const char* s = one_by_one ? src + ((b * 4 + k) * C + c) * L * esize
                                           : src + (b * C + c) * L * esize;
                char* d = dst + ((b * 4 + k) * C + c) * L * esize;
                for (int64_t l = 0; l < L; ++l)
                    memcpy(d + l * esize, s + xfo_route(k, l, H, W, scans) * esize, (size_t)esize);
            }
}

/* out[b,0,c] = even(c) ? x2[b,c] : x[b,c];  out[b,1,c] = even(c) ? x[b,c] : x2[b,c]
 * models/fusion_vmamba.py:198-213 */
void xfo_swap_scan(const void* x, const void* x2, void* out, int64_t B, int64_t C, int64_t L, int esize) {
    const char* a = (const char*)x;
    const char* b2 = (const char*)x2;
    char* o = (char*)out;
    const size_t row = (size_t)(L * esize);
    for (int64_t b = 0; b < B; ++b)
        for (int64_t c = 0; c < C; ++c) {
            const char* ra = a + (b * C + c) * row;
            const char* rb = b2 + (b * C + c) * row;
            int even = (c % 2 == 0);
            memcpy(o + ((b * 2 + 0) * C + c) * row, even ? rb : ra, row);
            memcpy(o + ((b * 2 + 1) * C + c) * row, even ? ra : rb, row);
        }
}
int64_t xfo_route_index(int k, int64_t l, int64_t H, int64_t W, int scans) { return xfo_route(k, l, H, W, scans); }
#endif /* XFO_BUILD_ROUTES */

/* ---------------------------------------------------------------------------------------------
 * cross_merge: ys (B,4,C,L) -> y (B,C,L).  models/csm_triton.py:60-67
 *   scans 0: t = ys[0:2] + flip(ys[2:4]);  y = t0 + transpose_back(t1)
 *   scans 1: y = ys.sum(1)  (evaluated 0+1+2+3 left to right)
 *   scans 2: t = ys[0:2] + flip(ys[2:4]);  y = t0 + t1
 * one_by_one (cross_merge1b1_fwd :134-179): ys (B,4,C,L) -> (B,4,C,L), pure un-routing, no adds.
 * ------------------------------------------------------------------------------------------- */
void XFO_NAME(xfo_cross_merge)(const real_t* ys, real_t* y, int64_t B, int64_t C, int64_t H, int64_t W, int scans) {
    const int64_t L = H * W;
    for (int64_t b = 0; b < B; ++b)
        for (int64_t c = 0; c < C; ++c) {
            const real_t* y0 = ys + ((b * 4 + 0) * C + c) * L;
            const real_t* y1 = ys + ((b * 4 + 1) * C + c) * L;
            const real_t* y2 = ys + ((b * 4 + 2) * C + c) * L;
            const real_t* y3 = ys + ((b * 4 + 3) * C + c) * L;
            real_t* o = y + (b * C + c) * L;
            for (int64_t h = 0; h < H; ++h)
                for (int64_t w = 0; w < W; ++w) {
                    int64_t p = h * W + w;
                    if (scans == 0) {
                        int64_t q = w * H + h;
                        real_t t0 = y0[p] + y2[L - 1 - p];
                        real_t t1 = y1[q] + y3[L - 1 - q];
                        o[p] = t0 + t1;
                    } else if (scans == 1) {
                        o[p] = ((y0[p] + y1[p]) + y2[p]) + y3[p];
                    } else {
                        real_t t0 = y0[p] + y2[L - 1 - p];
                        real_t t1 = y1[p] + y3[L - 1 - p];
                        o[p] = t0 + t1;
                    }
                }
        }
}

/* ---------------------------------------------------------------------------------------------
 * selective scan forward.  models/csms6s.py:25-68
 *   u, delta: (Bsz, KD, L); A: (KD, N); Bm, Cm: (Bsz, K, N, L); D, delta_bias: (KD) or NULL
 *   out: (Bsz, KD, L); last_state (optional): (Bsz, KD, N)
 * softplus follows torch.nn.functional.softplus (beta=1, threshold=20): x > 20 ? x : log1p(exp(x)).
 * ------------------------------------------------------------------------------------------- */
static inline real_t xfo_softplus(real_t x) {
    if (x > (real_t)20) return x;
    return (real_t)log1p(exp((double)x)) * (real_t)1; /* evaluated in double, rounded to real_t: <=0.5ulp like ATen's vectorised log1p(exp) to ~1ulp */
}

void XFO_NAME(xfo_selective_scan_fwd)(const real_t* u, const real_t* delta, const real_t* A, const real_t* Bm,
                                      const real_t* Cm, const real_t* D, const real_t* delta_bias, int delta_softplus,
                                      real_t* out, real_t* last_state, int64_t Bsz, int64_t KD, int64_t K, int64_t N,
                                      int64_t L) {
    const int64_t Cdim = KD / K;
    /* sequences are independent: one (b, d) row per OpenMP task (used by bench.py's CPU baseline; results do not
     * depend on the thread count) */
#pragma omp parallel for collapse(2) schedule(static)
    for (int64_t b = 0; b < Bsz; ++b)
        for (int64_t d = 0; d < KD; ++d) {
            real_t hbuf[256];
            real_t* h = (N <= 256) ? hbuf : (real_t*)malloc(sizeof(real_t) * (size_t)N);
            const int64_t g = d / Cdim;                                /* :53-54  B.view(..).repeat(1,1,Cdim,1,1) */
            const real_t* ur = u + (b * KD + d) * L;
            const real_t* dr = delta + (b * KD + d) * L;
            const real_t* Bg = Bm + (b * K + g) * N * L;
            const real_t* Cg = Cm + (b * K + g) * N * L;
            real_t* orow = out + (b * KD + d) * L;
            for (int64_t n = 0; n < N; ++n) h[n] = 0;                  /* :59 */
            for (int64_t l = 0; l < L; ++l) {
                real_t dt = dr[l];
                if (delta_bias) dt = dt + delta_bias[d];               /* :47-48 */
                if (delta_softplus) dt = xfo_softplus(dt);             /* :49-50 */
                real_t y = 0;
                for (int64_t n = 0; n < N; ++n) {
                    real_t dA = (real_t)exp((double)(dt * A[d * N + n]));   /* :55 */
                    real_t dBu = (dt * Bg[n * L + l]) * ur[l];         /* :56 */
                    h[n] = dA * h[n] + dBu;                            /* :62 */
                    y = y + h[n] * Cg[n * L + l];                      /* :63 */
                }
                orow[l] = D ? (y + ur[l] * D[d]) : y;                  /* :67 */
            }
            if (last_state)
                for (int64_t n = 0; n < N; ++n) last_state[(b * KD + d) * N + n] = h[n];
            if (h != hbuf) free(h);
        }
}

/* ---------------------------------------------------------------------------------------------
 * selective scan backward (closed form of d(out)/d(inputs) for the recurrence above).
 *   dh_l = C_l*dy_l + Abar_{l+1}*dh_{l+1}
 *   du   = D*dy + sum_n dh*dt*B          ddt = sum_n dh*(B*u + A*Abar*h_{l-1})
 *   dA   = sum_{b,l} dh*dt*Abar*h_{l-1}  dB  = sum_{d in g} dh*dt*u     dC = sum_{d in g} dy*h
 *   dD   = sum_{b,l} dy*u                ddelta_raw = ddt * sigmoid(raw+bias) (raw+bias<=20), dbias = sum ddelta_raw
 * All outputs must be zero-initialised by the caller for the accumulated ones (dA,dB,dC,dD,dbias).
 * ------------------------------------------------------------------------------------------- */
void XFO_NAME(xfo_selective_scan_bwd)(const real_t* u, const real_t* delta, const real_t* A, const real_t* Bm,
                                      const real_t* Cm, const real_t* D, const real_t* delta_bias, int delta_softplus,
                                      const real_t* dout, real_t* du, real_t* ddelta, real_t* dA, real_t* dB,
                                      real_t* dC, real_t* dD, real_t* dbias, int64_t Bsz, int64_t KD, int64_t K,
                                      int64_t N, int64_t L) {
    const int64_t Cdim = KD / K;
    /* parallel over batch images: dB/dC rows belong to one image; dA/dD/dbias are summed over images -> atomics
     * (summation order over b then depends on scheduling: fine for a baseline timer, tests run single-threaded) */
#pragma omp parallel for schedule(static)
    for (int64_t b = 0; b < Bsz; ++b) {
        real_t* hs = (real_t*)malloc(sizeof(real_t) * (size_t)(N * (L + 1)));   /* hs[n][l+1] = h_l, hs[n][0] = 0 */
        real_t* dts = (real_t*)malloc(sizeof(real_t) * (size_t)L);
        real_t* dh = (real_t*)malloc(sizeof(real_t) * (size_t)N);
        for (int64_t d = 0; d < KD; ++d) {
            const int64_t g = d / Cdim;
            const real_t* ur = u + (b * KD + d) * L;
            const real_t* dr = delta + (b * KD + d) * L;
            const real_t* dyr = dout + (b * KD + d) * L;
            const real_t* Bg = Bm + (b * K + g) * N * L;
            const real_t* Cg = Cm + (b * K + g) * N * L;
            real_t* dBg = dB + (b * K + g) * N * L;
            real_t* dCg = dC + (b * K + g) * N * L;
            for (int64_t l = 0; l < L; ++l) {
                real_t dt = dr[l];
                if (delta_bias) dt = dt + delta_bias[d];
                if (delta_softplus) dt = xfo_softplus(dt);
                dts[l] = dt;
            }
            for (int64_t n = 0; n < N; ++n) {
                real_t* hn = hs + n * (L + 1);
                hn[0] = 0;
                for (int64_t l = 0; l < L; ++l)
                    hn[l + 1] = (real_t)exp((double)(dts[l] * A[d * N + n])) * hn[l] + (dts[l] * Bg[n * L + l]) * ur[l];
                dh[n] = 0;
            }
            real_t dA_acc[256];
            real_t dD_acc = 0, dbias_acc = 0;
            for (int64_t n = 0; n < N && n < 256; ++n) dA_acc[n] = 0;
            for (int64_t l = L - 1; l >= 0; --l) {
                const real_t dy = dyr[l];
                real_t du_l = D ? D[d] * dy : 0;
                real_t ddt = 0;
                for (int64_t n = 0; n < N; ++n) {
                    const real_t* hn = hs + n * (L + 1);
                    const real_t a = A[d * N + n];
                    const real_t abar = (real_t)exp((double)(dts[l] * a));
                    /* dh[n] currently holds Abar_{l+1}*dh_{l+1} */
                    const real_t dhl = Cg[n * L + l] * dy + dh[n];
                    du_l += dhl * dts[l] * Bg[n * L + l];
                    ddt += dhl * (Bg[n * L + l] * ur[l] + a * abar * hn[l]);
                    dA_acc[n] += dhl * dts[l] * abar * hn[l];
                    dBg[n * L + l] += dhl * dts[l] * ur[l];
                    dCg[n * L + l] += dy * hn[l + 1];
                    dh[n] = abar * dhl;
                }
                if (D) dD_acc += dy * ur[l];
                du[(b * KD + d) * L + l] = du_l;
                if (delta_softplus) {
                    real_t raw = dr[l] + (delta_bias ? delta_bias[d] : 0);
                    if (raw <= (real_t)20) ddt = ddt * (real_t)(1.0 / (1.0 + exp(-(double)raw)));
                }
                ddelta[(b * KD + d) * L + l] = ddt;
                if (delta_bias) dbias_acc += ddt;
            }
            for (int64_t n = 0; n < N; ++n) {
#pragma omp atomic
                dA[d * N + n] += dA_acc[n];
            }
            if (D) {
#pragma omp atomic
                dD[d] += dD_acc;
            }
            if (delta_bias) {
#pragma omp atomic
                dbias[d] += dbias_acc;
            }
        }
        free(hs);
        free(dts);
        free(dh);
    }
}
