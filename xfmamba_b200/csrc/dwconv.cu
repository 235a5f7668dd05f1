// dwconv.cu -- depthwise 3x3 convolution (padding 1) fused with SiLU: the producer of the scan input x in every SS2D
// block (reference models/fusion_vmamba.py:405-413 + :1199-1200 `x = self.act(self.conv2d(x))`, same pair at :595-601 and
// :855-858 in the fusion blocks).  SURVEY section 8(f) rank 3: torch runs this as a cuDNN/ATen depthwise convolution, an
// element-wise SiLU, and in the backward a dgrad kernel, a wgrad kernel and the SiLU backward -- five passes over the
// tensor.  Here: one pass forward (read x, write y), one pass backward (read x, dy, write dx; dw/db reduced on the way).
//
// Layout: channel-first planes (B*C, H, W), contiguous.  A CTA owns P consecutive planes staged in shared memory with a
// zero halo; a warp owns a "unit" (a plane, or a slice of one when the CTA holds fewer planes than warps) so that the
// per-channel filter gradients reduce inside the warp.  A thread computes a strip of 4 rows x 1 column: lanes walk
// consecutive columns, so shared-memory reads are conflict free and global stores are whole lines.
// HBM-bound by design (8 B per element forward, 12 B backward at fp32), ~30 / ~50 issued instructions per element.
#include "xfscan_common.cuh"

namespace xfs {

constexpr int kDwThreads = 256;
constexpr int kDwWarps = kDwThreads / 32;
constexpr int kDwRows = 4;                   // rows per strip
constexpr int kDwMaxPlanes = 32;
constexpr size_t kDwSmemLimit = 200 * 1024;

struct DwGeom {
    int H, W, HW, C;
    int Hs, Wp, tile;        // staged tile: Hs rows x Wp columns (zero halo of one, rows padded to whole strips)
    int G, strips;           // row groups and strips (G * W) per plane
    int P, split, sps;       // planes per CTA, units per plane, strips per unit
    unsigned mW, mHW;        // magic multipliers: n / d == __umulhi(n, m) for the index ranges used here
    int64_t planes;          // B * C
};

__host__ inline unsigned dw_magic(unsigned d) { return (unsigned)((0x100000000ull + d - 1) / d); }

inline DwGeom dw_geom(int64_t B, int64_t C, int64_t H, int64_t W, int tiles_per_plane) {
    DwGeom g;
    g.H = (int)H; g.W = (int)W; g.HW = (int)(H * W); g.C = (int)C;
    g.G = (g.H + kDwRows - 1) / kDwRows;
    g.Hs = g.G * kDwRows + 2; g.Wp = g.W + 2; g.tile = g.Hs * g.Wp;
    g.tile = (g.tile + 3) & ~3;
    g.strips = g.G * g.W;
    g.planes = B * C;
    int P = 8192 / g.HW;
    P = P < 1 ? 1 : (P > kDwMaxPlanes ? kDwMaxPlanes : P);
    while (P > 1 && (size_t)P * g.tile * tiles_per_plane * sizeof(float) > 64 * 1024) --P;
    if (P >= kDwWarps) P -= P % kDwWarps; else while (kDwWarps % P) --P;       // whole planes per warp, or whole warps per plane
    if ((int64_t)P > g.planes) { P = (int)g.planes; if (P < kDwWarps) while (kDwWarps % P) --P; else P -= P % kDwWarps; }
    g.P = P;
    g.split = P >= kDwWarps ? 1 : kDwWarps / P;
    g.sps = (g.strips + g.split - 1) / g.split;
    g.mW = dw_magic((unsigned)g.W); g.mHW = dw_magic((unsigned)g.HW);
    return g;
}

__device__ __forceinline__ float silu_f(float v) { return v * rcp(1.0f + ex2(-v * kLog2e)); }

// zero the whole staging area, then copy the interiors of this CTA's planes (contiguous in global memory)
template <typename T>
__device__ __forceinline__ void dw_stage(const T* __restrict__ src, float* __restrict__ tile, const DwGeom& g, int np, int tid) {
    const int n = np * g.HW;
    for (int e = tid; e < n; e += kDwThreads) {
        const int p = __umulhi((unsigned)e, g.mHW), r = e - p * g.HW;
        const int h = __umulhi((unsigned)r, g.mW), w = r - h * g.W;
        tile[p * g.tile + (h + 1) * g.Wp + (w + 1)] = Elem<T>::to_f(src[e]);
    }
}

__device__ __forceinline__ void dw_zero(float* __restrict__ buf, int nfloats, int tid) {
    float4* b4 = reinterpret_cast<float4*>(buf);
    for (int i = tid; i < nfloats / 4; i += kDwThreads) b4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
}

// 6 x 3 window of a strip: rows h0-1 .. h0+4, columns w-1 .. w+1 (tile coordinates are shifted by the halo)
__device__ __forceinline__ void dw_window(const float* __restrict__ t, int h0, int w, int Wp, float (&v)[kDwRows + 2][3]) {
    const float* q = t + h0 * Wp + w;
#pragma unroll
    for (int r = 0; r < kDwRows + 2; ++r) {
#pragma unroll
        for (int c = 0; c < 3; ++c) v[r][c] = q[r * Wp + c];
    }
}

template <typename T>
__global__ void __launch_bounds__(kDwThreads)
dwconv_fwd_kernel(const T* __restrict__ x, const float* __restrict__ wgt, const float* __restrict__ bias, T* __restrict__ y,
                  const DwGeom g, const int act) {
    extern __shared__ __align__(16) float smem[];
    const int tid = threadIdx.x, lane = tid & 31, wp = tid >> 5;
    const int64_t plane0 = (int64_t)blockIdx.x * g.P;
    const int np = (int)min((int64_t)g.P, g.planes - plane0);
    dw_zero(smem, g.P * g.tile, tid);
    __syncthreads();
    dw_stage<T>(x + plane0 * g.HW, smem, g, np, tid);
    __syncthreads();

    const int units = np * g.split;
    for (int u = wp; u < units; u += kDwWarps) {
        const int p = u / g.split, part = u - p * g.split;
        const int c = (int)((plane0 + p) % g.C);
        float k[9];
#pragma unroll
        for (int i = 0; i < 9; ++i) k[i] = __ldg(wgt + c * 9 + i);
        const float bv = bias ? __ldg(bias + c) : 0.0f;
        const float* t = smem + p * g.tile;
        T* __restrict__ yo = y + (plane0 + p) * g.HW;
        const int s_end = min(g.strips, (part + 1) * g.sps);
        for (int s = part * g.sps + lane; s < s_end; s += 32) {
            const int gi = __umulhi((unsigned)s, g.mW), w = s - gi * g.W, h0 = gi * kDwRows;
            float v[kDwRows + 2][3];
            dw_window(t, h0, w, g.Wp, v);
#pragma unroll
            for (int r = 0; r < kDwRows; ++r) {
                float acc = bv;
#pragma unroll
                for (int i = 0; i < 3; ++i)
#pragma unroll
                    for (int j = 0; j < 3; ++j) acc = fmaf(v[r + i][j], k[i * 3 + j], acc);
                if (act) acc = silu_f(acc);
                if (h0 + r < g.H) yo[(h0 + r) * g.W + w] = Elem<T>::from_f(acc);
            }
        }
    }
}

// backward: dpre = dy * silu'(pre) with pre recomputed from the staged x; dx = correlation of dpre with the flipped
// filter; dw[c][i][j] = sum dpre[h][w] * x[h+i-1][w+j-1], db[c] = sum dpre -- written per plane into part (B*C, 10), summed
// over the batch by the caller (deterministic; no global atomics).
template <typename T>
__global__ void __launch_bounds__(kDwThreads)
dwconv_bwd_kernel(const T* __restrict__ x, const float* __restrict__ wgt, const float* __restrict__ bias, const T* __restrict__ dy,
                  T* __restrict__ dx, float* __restrict__ part, const DwGeom g, const int act) {
    extern __shared__ __align__(16) float smem[];
    float* tx = smem;                          // [P][tile] x with zero halo
    float* tg = smem + g.P * g.tile;           // [P][tile] dpre with zero halo
    float* acc = tg + g.P * g.tile;            // [P][10]
    const int tid = threadIdx.x, lane = tid & 31, wp = tid >> 5;
    const int64_t plane0 = (int64_t)blockIdx.x * g.P;
    const int np = (int)min((int64_t)g.P, g.planes - plane0);
    dw_zero(smem, 2 * g.P * g.tile, tid);
    for (int i = tid; i < g.P * 10; i += kDwThreads) acc[i] = 0.0f;
    __syncthreads();
    dw_stage<T>(x + plane0 * g.HW, tx, g, np, tid);
    __syncthreads();

    const int units = np * g.split;
    // ---- pass 1: dpre into shared memory
    for (int u = wp; u < units; u += kDwWarps) {
        const int p = u / g.split, part_i = u - p * g.split;
        const int c = (int)((plane0 + p) % g.C);
        float k[9];
#pragma unroll
        for (int i = 0; i < 9; ++i) k[i] = __ldg(wgt + c * 9 + i);
        const float bv = bias ? __ldg(bias + c) : 0.0f;
        const float* t = tx + p * g.tile;
        float* tgp = tg + p * g.tile;
        const T* __restrict__ dyo = dy + (plane0 + p) * g.HW;
        const int s_end = min(g.strips, (part_i + 1) * g.sps);
        for (int s = part_i * g.sps + lane; s < s_end; s += 32) {
            const int gi = __umulhi((unsigned)s, g.mW), w = s - gi * g.W, h0 = gi * kDwRows;
            float v[kDwRows + 2][3], d[kDwRows];
#pragma unroll
            for (int r = 0; r < kDwRows; ++r) d[r] = (h0 + r < g.H) ? Elem<T>::to_f(dyo[(h0 + r) * g.W + w]) : 0.0f;
            dw_window(t, h0, w, g.Wp, v);
#pragma unroll
            for (int r = 0; r < kDwRows; ++r) {
                float pre = bv;
#pragma unroll
                for (int i = 0; i < 3; ++i)
#pragma unroll
                    for (int j = 0; j < 3; ++j) pre = fmaf(v[r + i][j], k[i * 3 + j], pre);
                float gq = d[r];
                if (act) {                      // silu'(v) = s * (1 + v * (1 - s)), s = sigmoid(v)
                    const float sg = rcp(1.0f + ex2(-pre * kLog2e));
                    gq *= sg * fmaf(pre, 1.0f - sg, 1.0f);
                }
                if (h0 + r < g.H) tgp[(h0 + r + 1) * g.Wp + (w + 1)] = gq;
            }
        }
    }
    __syncthreads();
    // ---- pass 2: dx, dw, db
    for (int u = wp; u < units; u += kDwWarps) {
        const int p = u / g.split, part_i = u - p * g.split;
        const int c = (int)((plane0 + p) % g.C);
        float k[9];
#pragma unroll
        for (int i = 0; i < 9; ++i) k[i] = __ldg(wgt + c * 9 + i);
        const float* t = tx + p * g.tile;
        const float* tgp = tg + p * g.tile;
        T* __restrict__ dxo = dx + (plane0 + p) * g.HW;
        float dk[9], db = 0.0f;
#pragma unroll
        for (int i = 0; i < 9; ++i) dk[i] = 0.0f;
        const int s_end = min(g.strips, (part_i + 1) * g.sps);
        for (int s = part_i * g.sps + lane; s < s_end; s += 32) {
            const int gi = __umulhi((unsigned)s, g.mW), w = s - gi * g.W, h0 = gi * kDwRows;
            float v[kDwRows + 2][3], q[kDwRows + 2][3];
            dw_window(t, h0, w, g.Wp, v);
            dw_window(tgp, h0, w, g.Wp, q);
#pragma unroll
            for (int r = 0; r < kDwRows; ++r) {
                // dx[h][w] = sum_{i,j} dpre[h - i + 1][w - j + 1] * k[i][j]  (window row r + 2 - i, column 2 - j)
                float a = 0.0f;
#pragma unroll
                for (int i = 0; i < 3; ++i)
#pragma unroll
                    for (int j = 0; j < 3; ++j) a = fmaf(q[r + 2 - i][2 - j], k[i * 3 + j], a);
                if (h0 + r < g.H) dxo[(h0 + r) * g.W + w] = Elem<T>::from_f(a);
                const float gq = q[r + 1][1];          // dpre at (h0 + r, w); zero beyond the plane
                db += gq;
#pragma unroll
                for (int i = 0; i < 3; ++i)
#pragma unroll
                    for (int j = 0; j < 3; ++j) dk[i * 3 + j] = fmaf(gq, v[r + i][j], dk[i * 3 + j]);
            }
        }
#pragma unroll
        for (int i = 0; i < 9; ++i) dk[i] = warp_sum(dk[i]);
        db = warp_sum(db);
        if (lane == 0) {
#pragma unroll
            for (int i = 0; i < 9; ++i) atomicAdd(acc + p * 10 + i, dk[i]);     // shared memory; at most `split` adders per plane
            atomicAdd(acc + p * 10 + 9, db);
        }
    }
    __syncthreads();
    for (int i = tid; i < np * 10; i += kDwThreads) part[plane0 * 10 + i] = acc[i];
}

static size_t dw_smem(const DwGeom& g, bool backward) {
    return sizeof(float) * ((size_t)g.P * g.tile * (backward ? 2 : 1) + (backward ? g.P * 10 : 0));
}

int dwconv_supported(int64_t H, int64_t W, int backward) {
    if (H < 1 || W < 2 || H * W > (1 << 15)) return 0;      // index magic: n * d < 2^32, d >= 2
    const DwGeom g = dw_geom(1, 1, H, W, backward ? 2 : 1);
    return dw_smem(g, backward != 0) <= kDwSmemLimit;
}

template <typename T>
static int launch_dw_fwd_t(const void* x, const float* w, const float* b, void* y, const DwGeom& g, int act, cudaStream_t st) {
    const size_t smem = dw_smem(g, false);
    const unsigned grid = (unsigned)((g.planes + g.P - 1) / g.P);
    cudaFuncSetAttribute(dwconv_fwd_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    dwconv_fwd_kernel<T><<<grid, kDwThreads, smem, st>>>((const T*)x, w, b, (T*)y, g, act);
    return check_launch();
}

template <typename T>
static int launch_dw_bwd_t(const void* x, const float* w, const float* b, const void* dy, void* dx, float* part, const DwGeom& g,
                           int act, cudaStream_t st) {
    const size_t smem = dw_smem(g, true);
    const unsigned grid = (unsigned)((g.planes + g.P - 1) / g.P);
    cudaFuncSetAttribute(dwconv_bwd_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    dwconv_bwd_kernel<T><<<grid, kDwThreads, smem, st>>>((const T*)x, w, b, (const T*)dy, (T*)dx, part, g, act);
    return check_launch();
}

int launch_dwconv_fwd(const void* x, const float* w, const float* b, void* y, int64_t B, int64_t C, int64_t H, int64_t W, int dtype,
                      int act, cudaStream_t st) {
    if (!dwconv_supported(H, W, 0)) return XFS_ERR_UNSUPPORTED;
    const DwGeom g = dw_geom(B, C, H, W, 1);
    if (dtype == XFS_F32) return launch_dw_fwd_t<float>(x, w, b, y, g, act, st);
    if (dtype == XFS_BF16) return launch_dw_fwd_t<__nv_bfloat16>(x, w, b, y, g, act, st);
    return launch_dw_fwd_t<__half>(x, w, b, y, g, act, st);
}

int launch_dwconv_bwd(const void* x, const float* w, const float* b, const void* dy, void* dx, float* part, int64_t B, int64_t C,
                      int64_t H, int64_t W, int dtype, int act, cudaStream_t st) {
    if (!dwconv_supported(H, W, 1)) return XFS_ERR_UNSUPPORTED;
    const DwGeom g = dw_geom(B, C, H, W, 2);
    if (dtype == XFS_F32) return launch_dw_bwd_t<float>(x, w, b, dy, dx, part, g, act, st);
    if (dtype == XFS_BF16) return launch_dw_bwd_t<__nv_bfloat16>(x, w, b, dy, dx, part, g, act, st);
    return launch_dw_bwd_t<__half>(x, w, b, dy, dx, part, g, act, st);
}

}  // namespace xfs
