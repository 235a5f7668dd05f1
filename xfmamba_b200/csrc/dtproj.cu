// dtproj.cu -- the low-rank delta projection of the SS2D core:
//     delta[b, k*D + d, l] = sum_r W[k, d, r] * z[b, k, r, l]          (reference: F.conv1d(dts_r, dt_projs_weight, groups=K),
//                                                                        models/fusion_vmamba.py:1155-1157; einsum at :818)
// delta is the largest stream of the scan (4 of the 6 elements per (b, d, l)), so this kernel is bound by WRITING it; cuDNN
// runs the grouped 1x1 convolution as one implicit-GEMM launch per group at ~1/4 of that bound (0.70 ms for the
// (128, 4*192, 3136) case against 0.61 ms for the scan that consumes it).  SURVEY 8(f) rank 1 proposes folding the projection
// into the scan kernel instead; with one channel per CTA (what the four shared-memory image buffers allow) every CTA would
// re-read all R rows of z, i.e. R times the L2 traffic of delta itself, so the projection stays a separate, write-bound pass.
//
// fp32 FMA on purpose: TF32 tensor cores would miss the 1e-4 parity bar on delta.
// CTA = one (b, k), NCG groups of 8 channels x PG groups of 4 positions; z tile [R][4*PG] and W tile [R][8*NCG] in shared
// memory; a thread owns 8 channels x 4 positions (packed FFMA2 over position pairs), items flattened so that short rows
// (L = 196, 49) still fill the CTA.  The backward (dz = W^T g, dW = sum_b g z^T) is two plain batched GEMMs and goes to cuBLAS from proj.py.
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

int launch_dtproj_fwd(const void* z, const float* W, void* out, int64_t B, int64_t K, int64_t D, int64_t R, int64_t L, int64_t z_sb,
                      int64_t z_sk, int dtype, cudaStream_t st) {
    if (R > kDtMaxRank) return XFS_ERR_UNSUPPORTED;
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

}  // namespace xfs
