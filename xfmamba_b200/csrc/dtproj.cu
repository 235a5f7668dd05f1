// dtproj.cu -- the low-rank delta projection of the SS2D core:
//     delta[b, k*D + d, l] = sum_r W[k, d, r] * z[b, k, r, l]          (reference: F.conv1d(dts_r, dt_projs_weight, groups=K),
//                                                                        models/fusion_vmamba.py:1155-1157; einsum at :818)
// delta is the largest stream of the scan (4 of the 6 elements per (b, d, l)), so this kernel is bound by WRITING it; cuDNN
// runs the grouped 1x1 convolution as one implicit-GEMM launch per group at ~1/4 of that bound (0.70 ms for the
// (128, 4*192, 3136) case against 0.61 ms for the scan that consumes it).  SURVEY 8(f) rank 1 proposes folding the projection
// into the scan kernel instead; with one channel per CTA (what the four shared-memory image buffers allow) every CTA would
// re-read all R rows of z, i.e. R times the L2 traffic of delta itself, so the projection stays a separate, write-bound pass.
//
// Two kernels: an fp32 FFMA2 kernel (first) and a 3xTF32 warp-MMA kernel (second, the default); plain TF32 would miss the 1e-4 parity bar.
// CTA = one (b, k), NCG groups of 8 channels x PG groups of 4 positions; z tile [R][4*PG] and W tile [R][8*NCG] in shared
// memory; a thread owns 8 channels x 4 positions (packed FFMA2 over position pairs), items flattened so that short rows
// (L = 196, 49) still fill the CTA.  The backward (dz = W^T g, dW = sum_{b,l} g z^T) is the pair of MMA kernels at the end of this file.
#include <cstdlib>
#include "xfscan_common.cuh"

namespace xfs {

constexpr int kDtThreads = 256;
constexpr int kDtMaxRank = 64;
constexpr int kDtMaxPG = 64;         // position groups (of 4) per CTA
constexpr int kDtCh = 8;             // channels per thread

struct DtGeom {
    int D, R, L, K;
    int PG, NCG, items, zpitch;     // zpitch: floats per z row in shared memory
    int PGL, NB, B;                 // short rows: a CTA covers NB images, PGL position groups each (PG = NB * PGL)
    int64_t z_sb, z_sk;             // element strides of z between batches and routes (rows are contiguous: stride L)
    bool vec;                        // 16-byte loads/stores allowed (L % 4 == 0 and aligned bases/strides)
};

// ---- forward -----------------------------------------------------------------------------------------------------
template <typename T, bool kMulti>      // kMulti: short rows, several images per CTA (see dt_geom)
__global__ void __launch_bounds__(kDtThreads)
dtproj_fwd_kernel(const T* __restrict__ z, const float* __restrict__ W, T* __restrict__ out, const DtGeom g) {
    extern __shared__ __align__(16) float smem[];
    float* sz = smem;                         // [R][zpitch]
    float* sw = smem + g.R * g.zpitch;        // [R][kDtCh * NCG][2]: every weight stored twice, a ready operand pair for FFMA2
    const int tid = threadIdx.x;
    const int bgk = blockIdx.z, bg = bgk / g.K, k = bgk - bg * g.K;
    const int b0 = bg * g.NB;                                   // first image of this CTA
    const int l0 = kMulti ? 0 : blockIdx.x * (4 * g.PG), d0 = blockIdx.y * (kDtCh * g.NCG);
    const T* __restrict__ zk = z + (int64_t)k * g.z_sk;
    const T* __restrict__ zb0 = zk + (int64_t)b0 * g.z_sb;       // single-image path: everything below is relative to this
    const int wcols = kDtCh * g.NCG;

    // z tile: R rows of up to 4*PG positions
    if (g.vec) {
        for (int i = tid; i < g.R * g.PG; i += kDtThreads) {
            const int r = i / g.PG, pg = i - r * g.PG;
            const int bl = kMulti ? pg / g.PGL : 0, l = l0 + 4 * (pg - bl * g.PGL);
            const T* __restrict__ zb = kMulti ? zk + (int64_t)(b0 + bl) * g.z_sb : zb0;
            float v[4] = {0.f, 0.f, 0.f, 0.f};
            if (l < g.L && (!kMulti || b0 + bl < g.B)) {
                if constexpr (sizeof(T) == 4) {
                    const float4 a = __ldg(reinterpret_cast<const float4*>(zb + (int64_t)r * g.L + l));
                    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
                } else {
                    const uint2 a = __ldg(reinterpret_cast<const uint2*>(zb + (int64_t)r * g.L + l));
                    const T* e = reinterpret_cast<const T*>(&a);
#pragma unroll
                    for (int q = 0; q < 4; ++q) v[q] = Elem<T>::to_f(e[q]);
                }
            }
            *reinterpret_cast<float4*>(sz + r * g.zpitch + 4 * pg) = make_float4(v[0], v[1], v[2], v[3]);
        }
    } else {
        for (int i = tid; i < g.R * 4 * g.PG; i += kDtThreads) {
            const int r = i / (4 * g.PG), c = i - r * (4 * g.PG);
            const int bl = kMulti ? c / (4 * g.PGL) : 0, l = l0 + c - bl * (4 * g.PGL);
            const T* __restrict__ zb = kMulti ? zk + (int64_t)(b0 + bl) * g.z_sb : zb0;
            sz[r * g.zpitch + c] = (l < g.L && (!kMulti || b0 + bl < g.B)) ? Elem<T>::to_f(zb[(int64_t)r * g.L + l]) : 0.0f;
        }
    }
    // W tile transposed to [r][channel], duplicated
    for (int i = tid; i < wcols * g.R; i += kDtThreads) {
        const int c = i / g.R, r = i - c * g.R, d = d0 + c;
        const float wv = (d < g.D) ? __ldg(W + ((int64_t)k * g.D + d) * g.R + r) : 0.0f;
        *reinterpret_cast<float2*>(sw + 2 * (r * wcols + c)) = make_float2(wv, wv);
    }
    __syncthreads();

    for (int item = tid; item < g.items; item += kDtThreads) {
        const int cg = item / g.PG, pg = item - cg * g.PG;
        f2 acc[kDtCh][2];
#pragma unroll
        for (int c = 0; c < kDtCh; ++c) { acc[c][0] = splat2(0.0f); acc[c][1] = splat2(0.0f); }
        const float* zq = sz + 4 * pg;
        const float* wq = sw + 2 * kDtCh * cg;
#pragma unroll 2
        for (int r = 0; r < g.R; ++r) {
            const float4 zv = *reinterpret_cast<const float4*>(zq + r * g.zpitch);
            const f2 z01 = make_float2(zv.x, zv.y), z23 = make_float2(zv.z, zv.w);
#pragma unroll
            for (int h = 0; h < kDtCh / 2; ++h) {          // one 16-byte load = two duplicated weights
                const float4 wv = *reinterpret_cast<const float4*>(wq + 2 * r * wcols + 4 * h);
                const f2 wa = make_float2(wv.x, wv.y), wb = make_float2(wv.z, wv.w);
                acc[2 * h][0] = fma2(z01, wa, acc[2 * h][0]); acc[2 * h][1] = fma2(z23, wa, acc[2 * h][1]);
                acc[2 * h + 1][0] = fma2(z01, wb, acc[2 * h + 1][0]); acc[2 * h + 1][1] = fma2(z23, wb, acc[2 * h + 1][1]);
            }
        }
        const int bl = kMulti ? pg / g.PGL : 0, l = l0 + 4 * (pg - bl * g.PGL);
        const int bk = (b0 + bl) * g.K + k;
#pragma unroll
        for (int c = 0; c < kDtCh; ++c) {
            const int d = d0 + kDtCh * cg + c;
            if (d >= g.D || l >= g.L || (kMulti && b0 + bl >= g.B)) continue;
            T* __restrict__ o = out + ((int64_t)bk * g.D + d) * g.L + l;
            const float v[4] = {acc[c][0].x, acc[c][0].y, acc[c][1].x, acc[c][1].y};
            if (g.vec) {
                if constexpr (sizeof(T) == 4) *reinterpret_cast<float4*>(o) = make_float4(v[0], v[1], v[2], v[3]);
                else {
                    uint2 a;
                    T* e = reinterpret_cast<T*>(&a);
#pragma unroll
                    for (int q = 0; q < 4; ++q) e[q] = Elem<T>::from_f(v[q]);
                    *reinterpret_cast<uint2*>(o) = a;
                }
            } else {
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    if (l + q < g.L) o[q] = Elem<T>::from_f(v[q]);
            }
        }
    }
}

// ---- forward on the tensor pipe: 3xTF32 ---------------------------------------------------------------------------------
// The FFMA2 kernel above is bound by shared-memory operand traffic (ncu: LSU wavefronts 86 %, FMA pipe 66 %) at 1.8-2.8 TB/s of
// delta for the 28x28 / 14x14 stages.  The contraction is a batched (D x R) x (R x L) product with R <= 64, so the legacy warp
// MMA (mma.sync m16n8k8, tf32 in / fp32 accumulate) does it with 1/6 of the operand loads; fp32 accuracy is kept by splitting both
// operands into a tf32 head and a tf32 tail (x = hi + lo, ~20 significant bits, split by masking; z at staging time, W at use) and issuing three MMAs per tile,
// lo*hi + hi*lo + hi*hi (the dropped lo*lo term is 2^-20 relative).  tcgen05 is not used: K = R is 8..64, the tiles are tiny and
// the kernel only has to stay under the time it takes to write delta.
// CTA = 8 warps, (k, NB images, 16*4*kMW channels, <= 256 positions); warp w owns kMW m-tiles (16 channels each) and sweeps the
// n-tiles (8 positions) in groups of 4, even groups for warps 0-3 and odd groups for warps 4-7.  W is staged in fragment order (one LDS.128 per m-tile and k-step), z as [r][column] hi / lo
// planes with a row pitch = 8 mod 32 words so that the B-fragment loads are conflict free.
constexpr int kMmaThreads = 256;
constexpr int kMmaG = 4;             // n-tiles per accumulator group
constexpr int kMmaTP = 40;           // row pitch (floats) of a warp's transposition tile: conflict free both ways

struct DtMma {
    int D, R, L, K, B;
    int kK;                          // k-steps of 8 ranks (R padded with zeros)
    int NT, NTpad, NB, cols, pitch;  // positions per image and CTA, padded to 8; images per CTA; columns = NB * NTpad
    int64_t z_sb, z_sk;
    bool vec4, vec2;                 // 16-byte z loads / 4-position delta stores allowed
};

// x = hi + lo with hi, lo valid tf32 bit patterns (low 13 mantissa bits zero).  By masking, not `cvt.rna.tf32.f32`: the conversion
// issues on the XU pipe (a quarter-rate unit shared with the transcendentals), and at two conversions per operand element it, not the
// tensor pipe, bounded these kernels (ncu: xu pipe saturated, 25 % of all instructions).  Truncation leaves |lo| < 2^-10 |x| and drops
// 2^-20 |x| from lo: ~1e-6 relative per product, two orders under the parity bar.
__device__ __forceinline__ void tf32_split(float x, float& hi, float& lo) {
    hi = __uint_as_float(__float_as_uint(x) & 0xffffe000u);
    lo = __uint_as_float(__float_as_uint(x - hi) & 0xffffe000u);
}
__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint4& a, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b0), "r"(b1));
}

template <typename T, int kMW>
__global__ void __launch_bounds__(kMmaThreads, 3)
dtproj_mma_fwd_kernel(const T* __restrict__ z, const float* __restrict__ W, T* __restrict__ out, const DtMma g) {
    extern __shared__ __align__(16) float smem[];
    constexpr int kMT = 4 * kMW;                                  // m-tiles per CTA
    float* sA = smem;                                             // [kMT][kK][32 lanes] float4, fragment order (fp32; split at use)
    float* sBh = sA + kMT * g.kK * 32 * 4;                        // [8 * kK][pitch]
    float* sBl = sBh + 8 * g.kK * g.pitch;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int q = lane >> 2, t4 = lane & 3;
    const int bgk = blockIdx.z, bg = bgk / g.K, k = bgk - bg * g.K;
    const int b0 = bg * g.NB, l0 = blockIdx.x * g.NT, d0 = blockIdx.y * (16 * kMT);
    const int R8 = 8 * g.kK;

    const T* __restrict__ zk = z + (int64_t)k * g.z_sk + (int64_t)b0 * g.z_sb;
    const bool async_z = sizeof(T) == 4 && g.vec4;
    if (async_z) {
        // fp32 z tile: every thread puts all its 16-byte pieces in flight at once (cp.async, zero fill where out of range) and splits
        // them in place below, after the W tile -- one load latency per CTA instead of one per loop trip
        const int c4 = g.cols >> 2, n4 = g.NTpad >> 2;
        for (int r = warp; r < R8; r += kMmaThreads / 32)
            for (int c = lane; c < c4; c += 32) {
                const int bl = g.NB > 1 ? c / n4 : 0, l = l0 + 4 * (c - bl * n4);
                const bool in = r < g.R && l < g.L && b0 + bl < g.B;
                const T* __restrict__ src = in ? zk + (int64_t)bl * g.z_sb + (int64_t)r * g.L + l : zk;
                const uint32_t sa = (uint32_t)__cvta_generic_to_shared(sBh + r * g.pitch + 4 * c);
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(sa), "l"(src), "r"(in ? 16 : 0));
            }
        asm volatile("cp.async.commit_group;" ::);
    }
    // W tile (fp32), one item per fragment slot: a0 (row q, col t4), a1 (q + 8, t4), a2 (q, t4 + 4), a3 (q + 8, t4 + 4)
    for (int i = tid; i < kMT * g.kK * 32; i += kMmaThreads) {
        const int mk = i >> 5, mt = mk / g.kK, ks = mk - mt * g.kK;
        const int d = d0 + mt * 16 + q, r = ks * 8 + t4;          // lane of the slot = this thread's lane (kMmaThreads % 32 == 0)
        const float* __restrict__ wr = W + ((int64_t)k * g.D + d) * g.R + r;
        const bool d_lo = d < g.D, d_hi = d + 8 < g.D, r_lo = r < g.R, r_hi = r + 4 < g.R;
        const float a0 = (d_lo && r_lo) ? __ldg(wr) : 0.0f, a1 = (d_hi && r_lo) ? __ldg(wr + 8 * g.R) : 0.0f;
        const float a2 = (d_lo && r_hi) ? __ldg(wr + 4) : 0.0f, a3 = (d_hi && r_hi) ? __ldg(wr + 8 * g.R + 4) : 0.0f;
        reinterpret_cast<float4*>(sA)[mk * 32 + lane] = make_float4(a0, a1, a2, a3);
    }
    // z tile -> [r][column] head / tail planes (warps stride the rows, lanes the columns); rows R..R8 and columns past the row end are zero
    if (async_z) {
        const int c4 = g.cols >> 2;
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        for (int r = warp; r < R8; r += kMmaThreads / 32)
            for (int c = lane; c < c4; c += 32) {
                const float4 v = *reinterpret_cast<const float4*>(sBh + r * g.pitch + 4 * c);
                float4 h, t;
                tf32_split(v.x, h.x, t.x); tf32_split(v.y, h.y, t.y); tf32_split(v.z, h.z, t.z); tf32_split(v.w, h.w, t.w);
                *reinterpret_cast<float4*>(sBh + r * g.pitch + 4 * c) = h;
                *reinterpret_cast<float4*>(sBl + r * g.pitch + 4 * c) = t;
            }
    } else if (g.vec4) {
        const int c4 = g.cols >> 2, n4 = g.NTpad >> 2;
        for (int r = warp; r < R8; r += kMmaThreads / 32)
            for (int c = lane; c < c4; c += 32) {
                const int bl = g.NB > 1 ? c / n4 : 0, l = l0 + 4 * (c - bl * n4);
                float v[4] = {0.f, 0.f, 0.f, 0.f};
                if (r < g.R && l < g.L && b0 + bl < g.B) {
                    const T* __restrict__ src = zk + (int64_t)bl * g.z_sb + (int64_t)r * g.L + l;
                    if constexpr (sizeof(T) == 4) {
                        const float4 a = __ldg(reinterpret_cast<const float4*>(src));
                        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
                    } else {
                        const uint2 a = __ldg(reinterpret_cast<const uint2*>(src));
                        const T* e = reinterpret_cast<const T*>(&a);
#pragma unroll
                        for (int j = 0; j < 4; ++j) v[j] = Elem<T>::to_f(e[j]);
                    }
                }
                float4 h, t;
                tf32_split(v[0], h.x, t.x); tf32_split(v[1], h.y, t.y); tf32_split(v[2], h.z, t.z); tf32_split(v[3], h.w, t.w);
                *reinterpret_cast<float4*>(sBh + r * g.pitch + 4 * c) = h;
                *reinterpret_cast<float4*>(sBl + r * g.pitch + 4 * c) = t;
            }
    } else {
        for (int r = warp; r < R8; r += kMmaThreads / 32)
            for (int c = lane; c < g.cols; c += 32) {
                const int bl = g.NB > 1 ? c / g.NTpad : 0, p = c - bl * g.NTpad, l = l0 + p;
                float v = 0.0f;
                if (r < g.R && l < g.L && p < g.NT && b0 + bl < g.B) v = Elem<T>::to_f(zk[(int64_t)bl * g.z_sb + (int64_t)r * g.L + l]);
                float h, t;
                tf32_split(v, h, t);
                sBh[r * g.pitch + c] = h;
                sBl[r * g.pitch + c] = t;
            }
    }
    __syncthreads();

    // warp w: m-tiles (w & 3) * kMW .., n-tile groups of parity w >> 2
    const int mw0 = (warp & 3) * kMW;
    const int ntiles = g.cols >> 3;
    const float4* __restrict__ aw = reinterpret_cast<const float4*>(sA) + (mw0 * g.kK) * 32 + lane;
    float* st = sBl + 8 * g.kK * g.pitch + warp * (16 * kMmaTP);   // this warp's 16 x 32 transposition tile
    const int seg = lane & 7, srow = lane >> 3;
    for (int n0 = (warp >> 2) * kMmaG; n0 < ntiles; n0 += 2 * kMmaG) {
        float acc[kMW][kMmaG][4];
#pragma unroll
        for (int m = 0; m < kMW; ++m)
#pragma unroll
            for (int j = 0; j < kMmaG; ++j)
#pragma unroll
                for (int e = 0; e < 4; ++e) acc[m][j][e] = 0.0f;
        const uint32_t* __restrict__ bh = reinterpret_cast<const uint32_t*>(sBh) + t4 * g.pitch + n0 * 8 + q;
        const uint32_t* __restrict__ bt = reinterpret_cast<const uint32_t*>(sBl) + t4 * g.pitch + n0 * 8 + q;
        const int nv = ntiles - n0;                               // valid tiles of this group (>= 1)
        for (int ks = 0; ks < g.kK; ++ks) {
            uint4 ah[kMW], al[kMW];
#pragma unroll
            for (int m = 0; m < kMW; ++m) {
                const float4 a = aw[(m * g.kK + ks) * 32];
                float4 hi, lo;
                tf32_split(a.x, hi.x, lo.x); tf32_split(a.y, hi.y, lo.y); tf32_split(a.z, hi.z, lo.z); tf32_split(a.w, hi.w, lo.w);
                ah[m] = make_uint4(__float_as_uint(hi.x), __float_as_uint(hi.y), __float_as_uint(hi.z), __float_as_uint(hi.w));
                al[m] = make_uint4(__float_as_uint(lo.x), __float_as_uint(lo.y), __float_as_uint(lo.z), __float_as_uint(lo.w));
            }
            uint32_t h0[kMmaG], h1[kMmaG], t0[kMmaG], t1[kMmaG];
            const int ro = ks * 8 * g.pitch;
#pragma unroll
            for (int j = 0; j < kMmaG; ++j) {
                const int cj = j < nv ? 8 * j : 0;               // the tail group re-reads tile 0 (results discarded)
                h0[j] = bh[ro + cj]; h1[j] = bh[ro + 4 * g.pitch + cj];
                if constexpr (sizeof(T) == 4) { t0[j] = bt[ro + cj]; t1[j] = bt[ro + 4 * g.pitch + cj]; }
            }
            // three passes over independent accumulators: small terms first
#pragma unroll
            for (int j = 0; j < kMmaG; ++j)
#pragma unroll
                for (int m = 0; m < kMW; ++m) mma_tf32(acc[m][j], al[m], h0[j], h1[j]);
            if constexpr (sizeof(T) == 4) {       // bf16 / f16 rows are exact tf32 values: their tail is zero, two MMAs suffice
#pragma unroll
                for (int j = 0; j < kMmaG; ++j)
#pragma unroll
                    for (int m = 0; m < kMW; ++m) mma_tf32(acc[m][j], ah[m], t0[j], t1[j]);
            }
#pragma unroll
            for (int j = 0; j < kMmaG; ++j)
#pragma unroll
                for (int m = 0; m < kMW; ++m) mma_tf32(acc[m][j], ah[m], h0[j], h1[j]);
        }
        // accumulators (c0, c1: row q, columns 2 t4, 2 t4 + 1; c2, c3: row q + 8) -> shared tile -> full 128-byte row segments
        const int cc = n0 * 8 + 4 * seg;                          // this lane's 4 columns of the group
        const int bl = g.NB > 1 ? cc / g.NTpad : 0, pp = cc - bl * g.NTpad, l = l0 + pp;
        const bool colok = cc < g.cols && pp < g.NT && l < g.L && b0 + bl < g.B;
        T* __restrict__ obase = out + (((int64_t)(b0 + bl) * g.K + k) * g.D + d0 + mw0 * 16) * g.L + l;
#pragma unroll
        for (int m = 0; m < kMW; ++m) {
            __syncwarp();
#pragma unroll
            for (int j = 0; j < kMmaG; ++j) {
                *reinterpret_cast<float2*>(st + q * kMmaTP + 8 * j + 2 * t4) = make_float2(acc[m][j][0], acc[m][j][1]);
                *reinterpret_cast<float2*>(st + (q + 8) * kMmaTP + 8 * j + 2 * t4) = make_float2(acc[m][j][2], acc[m][j][3]);
            }
            __syncwarp();
#pragma unroll
            for (int it = 0; it < 4; ++it) {
                const int row = m * 16 + srow + 4 * it;
                if (!colok || d0 + mw0 * 16 + row >= g.D) continue;
                const float4 v = *reinterpret_cast<const float4*>(st + (srow + 4 * it) * kMmaTP + 4 * seg);
                T* __restrict__ o = obase + (int64_t)row * g.L;
                if (g.vec2) {                                     // rows are 16-byte (8-byte for 16-bit types) aligned
                    if constexpr (sizeof(T) == 4) *reinterpret_cast<float4*>(o) = v;
                    else {
                        T e[4] = {Elem<T>::from_f(v.x), Elem<T>::from_f(v.y), Elem<T>::from_f(v.z), Elem<T>::from_f(v.w)};
                        *reinterpret_cast<uint2*>(o) = *reinterpret_cast<const uint2*>(e);
                    }
                } else {
                    const float e[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                    for (int c = 0; c < 4; ++c)
                        if (l + c < g.L) o[c] = Elem<T>::from_f(e[c]);
                }
            }
        }
    }
}

static DtGeom dt_geom(int64_t B, int64_t D, int64_t R, int64_t L, int64_t K, int64_t z_sb, int64_t z_sk, bool aligned) {
    DtGeom g;
    g.D = (int)D; g.R = (int)R; g.L = (int)L; g.K = (int)K; g.B = (int)B;
    const int pgs = (int)((L + 3) / 4);
    if (pgs * 2 <= kDtMaxPG) {           // short rows (L <= 128): several images share a CTA and its weight tile
        g.PGL = pgs;
        g.NB = kDtMaxPG / pgs;
        if (g.NB > B) g.NB = (int)B;
        g.PG = g.NB * g.PGL;
    } else {
        g.PG = pgs < kDtMaxPG ? pgs : kDtMaxPG;
        g.PGL = g.PG;
        g.NB = 1;
    }
    // channel groups (of 8) per CTA: as many as 8 (fewer re-reads of the z tile), chosen so that NCG * PG items fill whole
    // rounds of the 256 threads (L = 196 -> PG = 49 -> NCG = 5: 245 items in one round, 96 % of the lanes busy)
    const int cgs = (int)((D + kDtCh - 1) / kDtCh);
    int best = 1;
    double best_eff = 0.0;
    for (int ncg = 8; ncg >= 1; --ncg) {
        if (ncg > cgs && ncg > 1) continue;
        const int items = ncg * g.PG, rounds = (items + kDtThreads - 1) / kDtThreads;
        const double eff = (double)items / (rounds * kDtThreads) - (ncg < 4 ? 0.03 * (4 - ncg) : 0.0);
        if (eff > best_eff + 1e-9) { best_eff = eff; best = ncg; }
    }
    g.NCG = best;
    g.items = g.NCG * g.PG;
    g.zpitch = 4 * g.PG + 4;
    g.z_sb = z_sb; g.z_sk = z_sk;
    g.vec = aligned && (L % 4 == 0) && (z_sb % 4 == 0) && (z_sk % 4 == 0);
    return g;
}

static size_t dt_mma_smem(const DtMma& g, int mw) {
    return sizeof(float) * ((size_t)4 * mw * g.kK * 32 * 4 + (size_t)2 * 8 * g.kK * g.pitch + (size_t)(kMmaThreads / 32) * 16 * kMmaTP);
}

static int launch_dtproj_mma(const void* z, const float* W, void* out, int64_t B, int64_t K, int64_t D, int64_t R, int64_t L, int64_t z_sb,
                             int64_t z_sk, int dtype, cudaStream_t st) {
    DtMma g;
    g.D = (int)D; g.R = (int)R; g.L = (int)L; g.K = (int)K; g.B = (int)B;
    g.kK = (int)((R + 7) / 8);
    const int mw = (D % 128 == 0 || D >= 512) ? 2 : 1;             // 128 channels per CTA unless that leaves a half-empty block
    // columns per CTA: at most 256, fewer for large ranks so that two CTAs still fit an SM (~100 KB each)
    const int budget = (54 * 1024 - 2 * mw * g.kK * 1024) / (64 * g.kK);         // 3 CTAs per SM: ~74 KB each incl. the 20 KB of transposition tiles
    int maxcols = budget < 256 ? (budget / 8) * 8 : 256;
    if (maxcols < 32) maxcols = 32;
    const int L8 = (int)((L + 7) / 8) * 8;
    // a whole row that misses the 3-CTA budget but fits two CTAs per SM (R = 32, 14x14) is better kept whole: 235 -> 216 us measured
    const int budget2 = (92 * 1024 - 2 * mw * g.kK * 1024) / (64 * g.kK);
    if (L8 > maxcols && L8 <= 256 && L8 <= budget2) maxcols = L8;
    unsigned ltiles = 1;
    if (L8 <= maxcols) {                                            // whole rows: several images per CTA share the weight tile
        g.NT = (int)L; g.NTpad = L8;
        g.NB = maxcols / L8;
        if (g.NB > B) g.NB = (int)B;
        if (g.NB > 8) g.NB = 8;
    } else {
        ltiles = (unsigned)((L + maxcols - 1) / maxcols);
        g.NT = (int)(((L + ltiles - 1) / ltiles + 7) / 8) * 8;
        g.NTpad = g.NT;
        g.NB = 1;
        ltiles = (unsigned)((L + g.NT - 1) / g.NT);
    }
    g.cols = g.NB * g.NTpad;
    g.pitch = g.cols + ((8 - g.cols % 32) + 32) % 32;               // = 8 (mod 32)
    g.z_sb = z_sb; g.z_sk = z_sk;
    const int esz = dtype == XFS_F32 ? 4 : 2;
    g.vec4 = (reinterpret_cast<uintptr_t>(z) % (4 * esz) == 0) && (L % 4 == 0) && (z_sb % 4 == 0) && (z_sk % 4 == 0);
    g.vec2 = (reinterpret_cast<uintptr_t>(out) % (4 * esz) == 0) && (L % 4 == 0);
    const dim3 grid(ltiles, (unsigned)((D + 64 * mw - 1) / (64 * mw)), (unsigned)(((B + g.NB - 1) / g.NB) * K));
    if (grid.y > 65535 || grid.z > 65535) return XFS_ERR_SHAPE;
    const size_t smem = dt_mma_smem(g, mw);
#define XFS_DT_MMA(T, M)                                                                                                \
    do {                                                                                                                \
        cudaFuncSetAttribute(dtproj_mma_fwd_kernel<T, M>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);      \
        dtproj_mma_fwd_kernel<T, M><<<grid, kMmaThreads, smem, st>>>((const T*)z, W, (T*)out, g);                        \
    } while (0)
    if (dtype == XFS_F32) { if (mw == 2) XFS_DT_MMA(float, 2); else XFS_DT_MMA(float, 1); }
    else if (dtype == XFS_BF16) { if (mw == 2) XFS_DT_MMA(__nv_bfloat16, 2); else XFS_DT_MMA(__nv_bfloat16, 1); }
    else { if (mw == 2) XFS_DT_MMA(__half, 2); else XFS_DT_MMA(__half, 1); }
#undef XFS_DT_MMA
    return check_launch();
}

int launch_dtproj_fwd(const void* z, const float* W, void* out, int64_t B, int64_t K, int64_t D, int64_t R, int64_t L, int64_t z_sb,
                      int64_t z_sk, int dtype, cudaStream_t st) {
    if (R > kDtMaxRank) return XFS_ERR_UNSUPPORTED;
    // tensor-pipe kernel except for one k-step with 64-channel CTAs (R <= 8 and D not a multiple of 128: the T/S stage-1 shape), where the
    // FFMA2 kernel below is already at 4.9 TB/s of delta and the MMA kernel reaches 3.8 (B200, profiles/r02_ops_sweep.md)
    static const bool force_fma = getenv("XFS_DTPROJ_FMA") != nullptr;    // timing experiments
    const bool one_step_narrow = R <= 8 && !(D % 128 == 0 || D >= 512);
    if (!force_fma && !one_step_narrow) return launch_dtproj_mma(z, W, out, B, K, D, R, L, z_sb, z_sk, dtype, st);
    const int esz = dtype == XFS_F32 ? 4 : 2;
    const bool aligned = ((reinterpret_cast<uintptr_t>(z) | reinterpret_cast<uintptr_t>(out)) % (4 * esz)) == 0;
    const DtGeom g = dt_geom(B, D, R, L, K, z_sb, z_sk, aligned);
    const unsigned ltiles = g.NB > 1 ? 1u : (unsigned)((L + 4 * g.PG - 1) / (4 * g.PG));
    const dim3 grid(ltiles, (unsigned)((D + kDtCh * g.NCG - 1) / (kDtCh * g.NCG)), (unsigned)(((B + g.NB - 1) / g.NB) * K));
    if (grid.y > 65535 || grid.z > 65535) return XFS_ERR_SHAPE;
    const size_t smem = sizeof(float) * ((size_t)g.R * g.zpitch + (size_t)g.R * 2 * kDtCh * g.NCG);
#define XFS_DT_LAUNCH(T, M)                                                                                             \
    do {                                                                                                                \
        cudaFuncSetAttribute(dtproj_fwd_kernel<T, M>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);          \
        dtproj_fwd_kernel<T, M><<<grid, kDtThreads, smem, st>>>((const T*)z, W, (T*)out, g);                             \
    } while (0)
    const bool multi = g.NB > 1;
    if (dtype == XFS_F32) { if (multi) XFS_DT_LAUNCH(float, true); else XFS_DT_LAUNCH(float, false); }
    else if (dtype == XFS_BF16) { if (multi) XFS_DT_LAUNCH(__nv_bfloat16, true); else XFS_DT_LAUNCH(__nv_bfloat16, false); }
    else { if (multi) XFS_DT_LAUNCH(__half, true); else XFS_DT_LAUNCH(__half, false); }
#undef XFS_DT_LAUNCH
    return check_launch();
}

// ---- backward on the tensor pipe (fp32 rows): dz = W^T g and dW = sum_{b,l} g z^T, one pass over g each -----------------
// cuBLAS ran these as two batched SIMT sgemms (matmul does not use TF32) plus a reduction over the batch: 274 us per 14x14 block
// of XFMamba-B (B = 64, R = 32, D = 1024) against 31 us for reading g once.  Same 3xTF32 arithmetic as the forward; the g tiles are
// streamed with cp.async (16-byte pieces, 4-byte ones when L % 4 != 0 or a base is unaligned; zero fill out of range) through a two-slot ring, the small operand is split at staging or at use.
constexpr int kBwThreads = 256;
constexpr int kDzKC = 32;            // dz: d rows of g per ring slot (4 k-steps)
constexpr int kDwLC = 32;            // dW: columns of g / z per ring slot (4 k-steps)
constexpr int kDwPitch = 36;         //   = 4 (mod 32): conflict-free A and B fragments
constexpr int kDwMT = 128;           // dW: d rows per CTA (one m-tile per warp)

__device__ __forceinline__ void cp_async16_zfill(void* smem_dst, const void* gsrc, bool in) {
    const uint32_t sa = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(sa), "l"(gsrc), "r"(in ? 16 : 0));
}
// four consecutive floats [l, l + 4) of a global row of L floats -> 16-byte aligned shared memory, zeros where l >= L or !row_ok.
// vec: rows are 16-byte aligned and L % 4 == 0 (one 16-byte copy); otherwise four 4-byte copies with their own bounds.
__device__ __forceinline__ void cp_async_row4(float* smem_dst, const float* row, int l, int L, bool row_ok, bool vec) {
    if (vec) {
        const bool in = row_ok && l < L;
        cp_async16_zfill(smem_dst, in ? row + l : row, in);
    } else {
        const uint32_t sa = (uint32_t)__cvta_generic_to_shared(smem_dst);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const bool in = row_ok && l + i < L;
            asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(sa + 4 * i), "l"(in ? row + l + i : row), "r"(in ? 4 : 0));
        }
    }
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::); }
template <int kN>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(kN) : "memory"); }

// dz[b, k, r, l] = sum_d W[k, d, r] g[b, k, d, l]: M = r (kRT m-tiles of 16), N = l, K = d.  CTA = (b, k, NT <= 256 columns); warp w owns
// n-tiles w, w + 8, ... (kNW of them), so one W^T chunk (staged once per CTA and 32 d rows, head / tail split there) serves up to 32 n-tiles.
template <int kRT, int kNW, bool kVec>
__global__ void __launch_bounds__(kBwThreads)
dtproj_dz_kernel(const float* __restrict__ g, const float* __restrict__ W, float* __restrict__ dz, int D, int R, int L, int K, int NT,
                 int pitch) {
    constexpr bool vec = kVec;
    extern __shared__ __align__(16) float smem[];
    float* sA = smem;                                          // [m-tile][k-step][hi, lo][lane] float4, fragment order
    float* sG = smem + kRT * (kDzKC / 8) * 2 * 32 * 4;         // [2 slots][32 d rows][pitch], pitch = 8 (mod 32)
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, q = lane >> 2, t4 = lane & 3;
    const int bk = blockIdx.y, k = bk % K, l0 = blockIdx.x * NT;
    const float* __restrict__ gb = g + (int64_t)bk * D * L;
    const float* __restrict__ Wk = W + (int64_t)k * D * R;
    const int nchunks = (D + kDzKC - 1) / kDzKC, p4 = NT >> 2, ntiles = NT >> 3;
    auto issue = [&](int c) {
        float* dst = sG + (c & 1) * kDzKC * pitch;
        for (int r = warp; r < kDzKC; r += kBwThreads / 32) {
            const int d = c * kDzKC + r;
            for (int c4 = lane; c4 < p4; c4 += 32)
                cp_async_row4(dst + r * pitch + 4 * c4, gb + (int64_t)(d < D ? d : 0) * L, l0 + 4 * c4, L, d < D, vec);
        }
        cp_async_commit();
    };
    float acc[kRT][kNW][4];
#pragma unroll
    for (int m = 0; m < kRT; ++m)
#pragma unroll
        for (int j = 0; j < kNW; ++j)
#pragma unroll
            for (int e = 0; e < 4; ++e) acc[m][j][e] = 0.0f;
    issue(0);
    for (int c = 0; c < nchunks; ++c) {
        if (c + 1 < nchunks) issue(c + 1);
        // W^T chunk in A-fragment order: a0 (r = q, d = t4), a1 (q + 8, t4), a2 (q, t4 + 4), a3 (q + 8, t4 + 4)
        for (int i = tid; i < kRT * (kDzKC / 8) * 32; i += kBwThreads) {
            const int mk = i >> 5, mt = mk / (kDzKC / 8), ks = mk - mt * (kDzKC / 8);
            const int r = mt * 16 + q, d = c * kDzKC + ks * 8 + t4;
            const float* __restrict__ wp = Wk + (int64_t)d * R + r;
            const bool r0 = r < R, r1 = r + 8 < R, d0 = d < D, d1 = d + 4 < D;
            const float a0 = (r0 && d0) ? __ldg(wp) : 0.0f, a1 = (r1 && d0) ? __ldg(wp + 8) : 0.0f;
            const float a2 = (r0 && d1) ? __ldg(wp + 4 * R) : 0.0f, a3 = (r1 && d1) ? __ldg(wp + 4 * R + 8) : 0.0f;
            float4 hi, lo;
            tf32_split(a0, hi.x, lo.x); tf32_split(a1, hi.y, lo.y); tf32_split(a2, hi.z, lo.z); tf32_split(a3, hi.w, lo.w);
            float4* dst = reinterpret_cast<float4*>(sA) + (mk * 2) * 32 + lane;
            dst[0] = hi;
            dst[32] = lo;
        }
        if (c + 1 < nchunks) cp_async_wait<1>(); else cp_async_wait<0>();
        __syncthreads();
        const float* __restrict__ gt = sG + (c & 1) * kDzKC * pitch + t4 * pitch + q;
        const uint4* __restrict__ aw = reinterpret_cast<const uint4*>(sA) + lane;
#pragma unroll
        for (int ks = 0; ks < kDzKC / 8; ++ks) {
            uint4 ah[kRT], al[kRT];
#pragma unroll
            for (int m = 0; m < kRT; ++m) {
                ah[m] = aw[((m * (kDzKC / 8) + ks) * 2) * 32];
                al[m] = aw[((m * (kDzKC / 8) + ks) * 2 + 1) * 32];
            }
            float h0[kNW], h1[kNW], t0[kNW], t1[kNW];
#pragma unroll
            for (int j = 0; j < kNW; ++j) {
                const int nt = warp + 8 * j, col = nt < ntiles ? nt * 8 : 0;     // tail tiles re-read tile 0, results discarded
                tf32_split(gt[ks * 8 * pitch + col], h0[j], t0[j]);
                tf32_split(gt[(ks * 8 + 4) * pitch + col], h1[j], t1[j]);
            }
#pragma unroll
            for (int j = 0; j < kNW; ++j)
#pragma unroll
                for (int m = 0; m < kRT; ++m) mma_tf32(acc[m][j], al[m], __float_as_uint(h0[j]), __float_as_uint(h1[j]));
#pragma unroll
            for (int j = 0; j < kNW; ++j)
#pragma unroll
                for (int m = 0; m < kRT; ++m) mma_tf32(acc[m][j], ah[m], __float_as_uint(t0[j]), __float_as_uint(t1[j]));
#pragma unroll
            for (int j = 0; j < kNW; ++j)
#pragma unroll
                for (int m = 0; m < kRT; ++m) mma_tf32(acc[m][j], ah[m], __float_as_uint(h0[j]), __float_as_uint(h1[j]));
        }
        __syncthreads();
    }
#pragma unroll
    for (int j = 0; j < kNW; ++j) {
        const int nt = warp + 8 * j, l = l0 + nt * 8 + 2 * t4;          // vec: a pair is in range or not as a whole, and 8-byte aligned
        if (nt >= ntiles || l >= L) continue;
#pragma unroll
        for (int m = 0; m < kRT; ++m)
#pragma unroll
            for (int hr = 0; hr < 2; ++hr) {
                const int r = m * 16 + q + 8 * hr;
                if (r >= R) continue;
                float* __restrict__ o = dz + ((int64_t)bk * R + r) * L + l;
                if (vec) *reinterpret_cast<float2*>(o) = make_float2(acc[m][j][2 * hr], acc[m][j][2 * hr + 1]);
                else {
                    o[0] = acc[m][j][2 * hr];
                    if (l + 1 < L) o[1] = acc[m][j][2 * hr + 1];
                }
            }
    }
}

// dW[k, d, r] += sum_{b in slice} sum_l g[b, k, d, l] z[b, k, r, l]: M = d (warp w = rows 16 w ..), N = r (kRN n-tiles of 8), K = l.
// CTA = (128 d rows, k, a slice of the batch); partial sums go to dW with red.global.add (dW is zeroed by the caller).
template <int kRN, bool kVec>
__global__ void __launch_bounds__(kBwThreads)
dtproj_dw_kernel(const float* __restrict__ g, const float* __restrict__ z, float* __restrict__ dW, int B, int D, int R, int L, int K,
                 int64_t z_sb, int64_t z_sk, int nb) {
    constexpr bool vec = kVec;
    extern __shared__ __align__(16) float smem[];
    float (*sG)[kDwMT * kDwPitch] = reinterpret_cast<float (*)[kDwMT * kDwPitch]>(smem);
    float (*sZ)[kRN * 8 * kDwPitch] = reinterpret_cast<float (*)[kRN * 8 * kDwPitch]>(smem + 2 * kDwMT * kDwPitch);   // tf32 heads
    float (*sZl)[kRN * 8 * kDwPitch] = sZ + 2;                                                                      // tf32 tails
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, q = lane >> 2, t4 = lane & 3;
    const int d0 = blockIdx.x * kDwMT, k = blockIdx.y, b_lo = blockIdx.z * nb, b_hi = min(B, b_lo + nb);
    const int lchunks = (L + kDwLC - 1) / kDwLC, items = (b_hi - b_lo) * lchunks;
    auto issue = [&](int it) {
        const int bi = it / lchunks, lc = it - bi * lchunks, b = b_lo + bi, lbase = lc * kDwLC;
        const float* __restrict__ gb = g + ((int64_t)b * K + k) * D * L;
        const float* __restrict__ zb = z + (int64_t)b * z_sb + (int64_t)k * z_sk;
        float* dg = sG[it & 1];
        float* dzt = sZ[it & 1];
#pragma unroll
        for (int i = 0; i < (kDwMT * 8) / kBwThreads; ++i) {          // 128 rows x 8 pieces
            const int e = tid + i * kBwThreads, r = e >> 3, c4 = e & 7, d = d0 + r;
            cp_async_row4(dg + r * kDwPitch + 4 * c4, gb + (int64_t)(d < D ? d : 0) * L, lbase + 4 * c4, L, d < D, vec);
        }
        for (int e = tid; e < kRN * 8 * 8; e += kBwThreads) {
            const int r = e >> 3, c4 = e & 7;
            cp_async_row4(dzt + r * kDwPitch + 4 * c4, zb + (int64_t)(r < R ? r : 0) * L, lbase + 4 * c4, L, r < R, vec);
        }
        cp_async_commit();
    };
    float acc[kRN][4];
#pragma unroll
    for (int n = 0; n < kRN; ++n)
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[n][e] = 0.0f;
    if (items > 0) issue(0);
    for (int it = 0; it < items; ++it) {
        if (it + 1 < items) { issue(it + 1); cp_async_wait<1>(); } else cp_async_wait<0>();
        for (int e = tid; e < kRN * 8 * 8; e += kBwThreads) {       // the z pieces this thread copied: head / tail split, shared by all warps
            float* pz = sZ[it & 1] + (e >> 3) * kDwPitch + 4 * (e & 7);
            const float4 v = *reinterpret_cast<const float4*>(pz);
            float4 h, t;
            tf32_split(v.x, h.x, t.x); tf32_split(v.y, h.y, t.y); tf32_split(v.z, h.z, t.z); tf32_split(v.w, h.w, t.w);
            *reinterpret_cast<float4*>(pz) = h;
            *reinterpret_cast<float4*>(sZl[it & 1] + (e >> 3) * kDwPitch + 4 * (e & 7)) = t;
        }
        __syncthreads();
        const float* __restrict__ ga = sG[it & 1] + (warp * 16 + q) * kDwPitch + t4;
        const float* __restrict__ zt = sZ[it & 1] + q * kDwPitch + t4;
        const float* __restrict__ zl = sZl[it & 1] + q * kDwPitch + t4;
#pragma unroll
        for (int ks = 0; ks < kDwLC / 8; ++ks) {
            float4 hi, lo;                    // a0 (d = q, l = t4), a1 (q + 8, t4), a2 (q, t4 + 4), a3 (q + 8, t4 + 4)
            tf32_split(ga[ks * 8], hi.x, lo.x); tf32_split(ga[8 * kDwPitch + ks * 8], hi.y, lo.y);
            tf32_split(ga[ks * 8 + 4], hi.z, lo.z); tf32_split(ga[8 * kDwPitch + ks * 8 + 4], hi.w, lo.w);
            const uint4 ah = make_uint4(__float_as_uint(hi.x), __float_as_uint(hi.y), __float_as_uint(hi.z), __float_as_uint(hi.w));
            const uint4 al = make_uint4(__float_as_uint(lo.x), __float_as_uint(lo.y), __float_as_uint(lo.z), __float_as_uint(lo.w));
            float h0[kRN], h1[kRN], t0[kRN], t1[kRN];
#pragma unroll
            for (int n = 0; n < kRN; ++n) {   // b0 (l = t4, r = q), b1 (l = t4 + 4, r = q)
                h0[n] = zt[n * 8 * kDwPitch + ks * 8]; t0[n] = zl[n * 8 * kDwPitch + ks * 8];
                h1[n] = zt[n * 8 * kDwPitch + ks * 8 + 4]; t1[n] = zl[n * 8 * kDwPitch + ks * 8 + 4];
            }
#pragma unroll
            for (int n = 0; n < kRN; ++n) mma_tf32(acc[n], al, __float_as_uint(h0[n]), __float_as_uint(h1[n]));
#pragma unroll
            for (int n = 0; n < kRN; ++n) mma_tf32(acc[n], ah, __float_as_uint(t0[n]), __float_as_uint(t1[n]));
#pragma unroll
            for (int n = 0; n < kRN; ++n) mma_tf32(acc[n], ah, __float_as_uint(h0[n]), __float_as_uint(h1[n]));
        }
        __syncthreads();
    }
    // c0, c1: (d = q, r = 2 t4, 2 t4 + 1); c2, c3: d = q + 8
#pragma unroll
    for (int n = 0; n < kRN; ++n)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int d = d0 + warp * 16 + q + 8 * (e >> 1), r = n * 8 + 2 * t4 + (e & 1);
            if (d < D && r < R) atomicAdd(dW + ((int64_t)k * D + d) * R + r, acc[n][e]);
        }
}

// g: (B, K*D, L) f32 contiguous; z as in the forward; dz: (B, K, R, L) f32 contiguous or NULL; dW: (K, D, R) f32, ZEROED by the caller, or NULL
int launch_dtproj_bwd(const float* g, const float* z, const float* W, float* dz, float* dW, int64_t B, int64_t K, int64_t D, int64_t R,
                      int64_t L, int64_t z_sb, int64_t z_sk, cudaStream_t st) {
    if (R > kDtMaxRank) return XFS_ERR_UNSUPPORTED;
    const bool vec = L % 4 == 0 && z_sb % 4 == 0 && z_sk % 4 == 0 &&
                     (reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(z) | reinterpret_cast<uintptr_t>(dz)) % 16 == 0;
    if (B * K > 65535 * 32) return XFS_ERR_SHAPE;
    if (dz) {
        const int64_t L8 = (L + 7) / 8 * 8;
        const int64_t ltiles = (L8 + 255) / 256;                                  // narrower tiles to fill the GPU at small B*K measured slower (the W^T chunk is re-staged per CTA)
        const int NT = (int)(((L + ltiles - 1) / ltiles + 7) / 8 * 8);        // <= 256 columns per CTA, balanced over the tiles
        const int pitch = NT + ((8 - NT % 32) + 32) % 32;
        const int nw = (NT / 8 + 7) / 8, rt = (int)((R + 15) / 16);
        const dim3 grid((unsigned)((L + NT - 1) / NT), (unsigned)(B * K));
        const size_t smem = sizeof(float) * ((size_t)rt * (kDzKC / 8) * 2 * 32 * 4 + (size_t)2 * kDzKC * pitch);
#define XFS_DZ_V(RT, NW, V)                                                                                            \
    do {                                                                                                                \
        cudaFuncSetAttribute(dtproj_dz_kernel<RT, NW, V>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);      \
        dtproj_dz_kernel<RT, NW, V><<<grid, kBwThreads, smem, st>>>(g, W, dz, (int)D, (int)R, (int)L, (int)K, NT, pitch); \
    } while (0)
#define XFS_DZ(RT, NW) do { if (vec) XFS_DZ_V(RT, NW, true); else XFS_DZ_V(RT, NW, false); } while (0)
#define XFS_DZ_R(NW) do { if (rt == 1) XFS_DZ(1, NW); else if (rt == 2) XFS_DZ(2, NW); else if (rt == 3) XFS_DZ(3, NW); else XFS_DZ(4, NW); } while (0)
        if (nw == 1) XFS_DZ_R(1); else if (nw == 2) XFS_DZ_R(2); else if (nw == 3) XFS_DZ_R(3); else XFS_DZ_R(4);
#undef XFS_DZ_R
#undef XFS_DZ
#undef XFS_DZ_V
        const int rc = check_launch();
        if (rc) return rc;
    }
    if (dW) {
        const int64_t blocks = ((D + kDwMT - 1) / kDwMT) * K;
        int64_t slices = (148 * 4 + blocks - 1) / blocks;               // ~4 CTAs per SM in flight
        if (slices > B) slices = B;
        const int nb = (int)((B + slices - 1) / slices);
        const dim3 grid((unsigned)((D + kDwMT - 1) / kDwMT), (unsigned)K, (unsigned)((B + nb - 1) / nb));
        const int rn = (int)((R + 7) / 8);
        const size_t smem = sizeof(float) * 2 * ((size_t)kDwMT + 2 * 8 * rn) * kDwPitch;
#define XFS_DW_V(N, V)                                                                                                  \
    do {                                                                                                                \
        cudaFuncSetAttribute(dtproj_dw_kernel<N, V>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);           \
        dtproj_dw_kernel<N, V><<<grid, kBwThreads, smem, st>>>(g, z, dW, (int)B, (int)D, (int)R, (int)L, (int)K, z_sb, z_sk, nb); \
    } while (0)
#define XFS_DW(N) do { if (vec) XFS_DW_V(N, true); else XFS_DW_V(N, false); } while (0)
        switch (rn) {
            case 1: XFS_DW(1); break; case 2: XFS_DW(2); break; case 3: XFS_DW(3); break; case 4: XFS_DW(4); break;
            case 5: XFS_DW(5); break; case 6: XFS_DW(6); break; case 7: XFS_DW(7); break; default: XFS_DW(8); break;
        }
#undef XFS_DW
#undef XFS_DW_V
        return check_launch();
    }
    return XFS_OK;
}

}  // namespace xfs
