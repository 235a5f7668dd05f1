// routes.cu -- CrossScan / CrossMerge / SwappingScan as stand-alone sm_100a kernels.
//
// Replaces the Triton kernel triton_cross_scan_flex (reference models/csm_triton.py:278-400), whose transposed routes
// are scattered 4-byte accesses, and the torch index ops of SwappingScan_multiview (models/fusion_vmamba.py:198-213,
// which builds its mask on the CPU).  One CTA moves a 32x32 spatial tile of one (batch, channel) image; the
// column-major routes go through a padded shared-memory tile so that every global access is a full 128-byte line.
// HBM-bound: 5 elements per (b, c, l) (1 read + 4 writes for scan, 4 reads + 1 write for merge).
#include "xfscan_common.cuh"

namespace xfs {

constexpr int kTile = 32;
constexpr int kRows = 8;   // blockDim.y

// ---------------------------------------------------------------------------------------------------------
// cross_scan: raw-bit copy, U = uint32_t (f32) or uint16_t (bf16 / f16)
// ---------------------------------------------------------------------------------------------------------
template <typename U>
__global__ void __launch_bounds__(kTile* kRows) cross_scan_kernel(const U* __restrict__ x, U* __restrict__ xs,
                                                                    int64_t C, int H, int W, int scans, int one_by_one) {
    __shared__ U tile[kTile][kTile + 1];
    const int64_t bc = blockIdx.x;          // b*C + c (grid.x: no 65535 limit)
    const int64_t b = bc / C, c = bc % C;
    const int64_t L = (int64_t)H * W;
    const int h0 = blockIdx.y * kTile, w0 = blockIdx.z * kTile;
    const int tx = threadIdx.x, ty = threadIdx.y;
    U* __restrict__ o[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) o[k] = xs + ((b * 4 + k) * C + c) * L;

    for (int k = 0; k < 4; ++k) {
        if (k > 0 && !one_by_one) break;     // one source image unless one_by_one
        const U* __restrict__ src = one_by_one ? x + ((b * 4 + k) * C + c) * L : x + bc * L;
        if (k > 0) __syncthreads();
        // row-major routes straight from the coalesced read; stash for the column-major ones
#pragma unroll
        for (int r = ty; r < kTile; r += kRows) {
            const int h = h0 + r, w = w0 + tx;
            if (h < H && w < W) {
                const int64_t p = (int64_t)h * W + w;
                const U v = src[p];
                tile[r][tx] = v;
                if (scans == XFS_SCANS_CROSS2D) {
                    if (!one_by_one) { o[0][p] = v; o[2][L - 1 - p] = v; }
                    else if (k == 0) o[0][p] = v;
                    else if (k == 2) o[2][L - 1 - p] = v;
                } else if (scans == XFS_SCANS_UNIDI) {
                    if (!one_by_one) { o[0][p] = v; o[1][p] = v; o[2][p] = v; o[3][p] = v; }
                    else o[k][p] = v;
                } else {
                    if (!one_by_one) { o[0][p] = v; o[1][p] = v; o[2][L - 1 - p] = v; o[3][L - 1 - p] = v; }
                    else if (k < 2) o[k][p] = v;
                    else o[k][L - 1 - p] = v;
                }
            }
        }
        if (scans != XFS_SCANS_CROSS2D) continue;
        __syncthreads();
        // column-major routes: thread (tx, r) now owns (h = h0+tx, w = w0+r); consecutive tx -> consecutive q
#pragma unroll
        for (int r = ty; r < kTile; r += kRows) {
            const int h = h0 + tx, w = w0 + r;
            if (h < H && w < W) {
                const int64_t q = (int64_t)w * H + h;
                const U v = tile[tx][r];
                if (!one_by_one) { o[1][q] = v; o[3][L - 1 - q] = v; }
                else if (k == 1) o[1][q] = v;
                else if (k == 3) o[3][L - 1 - q] = v;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// cross_merge.  Every add is rounded to T, in the reference's order (models/csm_triton.py:61-62):
//   t0 = ys0 + flip(ys2);  t1 = ys1 + flip(ys3);  y = t0 + transpose_back(t1)
// ---------------------------------------------------------------------------------------------------------
template <typename T> __device__ __forceinline__ T add_rn(T a, T b) {
    return Elem<T>::from_f(Elem<T>::to_f(a) + Elem<T>::to_f(b));
}

template <typename T>
__global__ void __launch_bounds__(kTile* kRows) cross_merge_kernel(const T* __restrict__ ys, T* __restrict__ y,
                                                                     int64_t C, int H, int W, int scans, int one_by_one) {
    __shared__ T tile[kTile][kTile + 1];
    __shared__ T tile3[kTile][kTile + 1];   // only one_by_one needs the two column-major routes separately
    const int64_t bc = blockIdx.x;
    const int64_t b = bc / C, c = bc % C;
    const int64_t L = (int64_t)H * W;
    const int h0 = blockIdx.y * kTile, w0 = blockIdx.z * kTile;
    const int tx = threadIdx.x, ty = threadIdx.y;
    const T* __restrict__ s[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) s[k] = ys + ((b * 4 + k) * C + c) * L;

    if (scans == XFS_SCANS_CROSS2D) {
#pragma unroll
        for (int r = ty; r < kTile; r += kRows) {
            const int h = h0 + tx, w = w0 + r;
            if (h < H && w < W) {
                const int64_t q = (int64_t)w * H + h;
                if (!one_by_one) tile[tx][r] = add_rn(s[1][q], s[3][L - 1 - q]);
                else { tile[tx][r] = s[1][q]; tile3[tx][r] = s[3][L - 1 - q]; }
            }
        }
        __syncthreads();
    }
#pragma unroll
    for (int r = ty; r < kTile; r += kRows) {
        const int h = h0 + r, w = w0 + tx;
        if (h < H && w < W) {
            const int64_t p = (int64_t)h * W + w;
            if (one_by_one) {
                T* __restrict__ o = y + (b * 4 * C + c) * L;   // (B, 4, C, L)
                const int64_t ks = C * L;
                if (scans == XFS_SCANS_CROSS2D) {
                    o[p] = s[0][p]; o[ks + p] = tile[r][tx]; o[2 * ks + p] = s[2][L - 1 - p]; o[3 * ks + p] = tile3[r][tx];
                } else if (scans == XFS_SCANS_UNIDI) {
                    o[p] = s[0][p]; o[ks + p] = s[1][p]; o[2 * ks + p] = s[2][p]; o[3 * ks + p] = s[3][p];
                } else {
                    o[p] = s[0][p]; o[ks + p] = s[1][p]; o[2 * ks + p] = s[2][L - 1 - p]; o[3 * ks + p] = s[3][L - 1 - p];
                }
            } else if (scans == XFS_SCANS_CROSS2D) {
                y[bc * L + p] = add_rn(add_rn(s[0][p], s[2][L - 1 - p]), tile[r][tx]);
            } else if (scans == XFS_SCANS_UNIDI) {
                // torch .sum(1) over 4 values: fp32 accumulate, one rounding to T
                y[bc * L + p] = Elem<T>::from_f(((Elem<T>::to_f(s[0][p]) + Elem<T>::to_f(s[1][p])) + Elem<T>::to_f(s[2][p])) +
                                                Elem<T>::to_f(s[3][p]));
            } else {
                const T t0 = add_rn(s[0][p], s[2][L - 1 - p]);
                const T t1 = add_rn(s[1][p], s[3][L - 1 - p]);
                y[bc * L + p] = Elem<T>::from_f(Elem<T>::to_f(t0) + Elem<T>::to_f(t1));   // .sum(1) of two values
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// whole-plane variants (scans == cross2d, H*W*4 <= 48 KB): the 32x32 tiles leave 41 % of their threads
// idle at 56x56 and write the column-major routes in runs of at most 32 elements.  Here a CTA stages one (b, c) plane in
// shared memory (odd pitch: the column-major pass reads it conflict free); every global access is part of a run of L.
// ---------------------------------------------------------------------------------------------------------
constexpr int kPlaneThreads = 256;
constexpr size_t kPlaneSmem = 48 * 1024;
__host__ __device__ inline unsigned plane_magic(unsigned d) { return (unsigned)((0x100000000ull + d - 1) / d); }

template <typename U, bool kOneByOne>
__global__ void __launch_bounds__(kPlaneThreads) cross_scan_plane_kernel(const U* __restrict__ x, U* __restrict__ xs, int64_t C,
                                                                         int H, int W, unsigned mH, unsigned mW) {
    extern __shared__ __align__(16) unsigned char plane_smem[];
    U* tile = reinterpret_cast<U*>(plane_smem);
    const int pitch = W | 1, L = H * W;
    const int64_t bc = blockIdx.x, b = bc / C, c = bc % C;
    U* __restrict__ o[4];
    const U* __restrict__ src[4];           // one_by_one: route k reads its own image (B, 4, C, H, W)
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        o[k] = xs + ((b * 4 + k) * C + c) * L;
        src[k] = kOneByOne ? x + ((b * 4 + k) * C + c) * L : x + bc * L;
    }
    for (int p = threadIdx.x; p < L; p += kPlaneThreads) {
        const U v = src[0][p];
        if (!kOneByOne) {
            const int h = __umulhi((unsigned)p, mW), w = p - h * W;
            tile[h * pitch + w] = v;
            o[2][L - 1 - p] = v;
        } else {
            o[2][L - 1 - p] = src[2][p];
        }
        o[0][p] = v;
    }
#pragma unroll
    for (int k = 1; k < 4; k += 2) {        // the two column-major routes
        if (kOneByOne) {
            if (k == 3) __syncthreads();
            for (int p = threadIdx.x; p < L; p += kPlaneThreads) {
                const int h = __umulhi((unsigned)p, mW), w = p - h * W;
                tile[h * pitch + w] = src[k][p];
            }
        } else if (k == 3) {
            break;                           // same staged image: both routes were written in the k == 1 pass
        }
        __syncthreads();
        for (int q = threadIdx.x; q < L; q += kPlaneThreads) {
            const int w = __umulhi((unsigned)q, mH), h = q - w * H;
            const U v = tile[h * pitch + w];
            if (!kOneByOne) { o[1][q] = v; o[3][L - 1 - q] = v; }
            else if (k == 1) o[1][q] = v;
            else o[3][L - 1 - q] = v;
        }
    }
}

template <typename T, bool kOneByOne>
__global__ void __launch_bounds__(kPlaneThreads) cross_merge_plane_kernel(const T* __restrict__ ys, T* __restrict__ y, int64_t C,
                                                                          int H, int W, unsigned mH, unsigned mW) {
    extern __shared__ __align__(16) unsigned char plane_smem[];
    T* tile = reinterpret_cast<T*>(plane_smem);
    const int pitch = W | 1, L = H * W;
    const int64_t bc = blockIdx.x, b = bc / C, c = bc % C;
    const T* __restrict__ s[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) s[k] = ys + ((b * 4 + k) * C + c) * L;
    if (!kOneByOne) {
        for (int q = threadIdx.x; q < L; q += kPlaneThreads) {
            const int w = __umulhi((unsigned)q, mH), h = q - w * H;
            tile[h * pitch + w] = add_rn(s[1][q], s[3][L - 1 - q]);
        }
        __syncthreads();
        for (int p = threadIdx.x; p < L; p += kPlaneThreads) {
            const int h = __umulhi((unsigned)p, mW), w = p - h * W;
            y[bc * L + p] = add_rn(add_rn(s[0][p], s[2][L - 1 - p]), tile[h * pitch + w]);
        }
    } else {                                 // (B, 4, C, L) out: every route back in spatial order, no sum
        T* __restrict__ o = y + (b * 4 * C + c) * L;
        const int64_t ks = C * L;
        for (int p = threadIdx.x; p < L; p += kPlaneThreads) { o[p] = s[0][p]; o[2 * ks + p] = s[2][L - 1 - p]; }
#pragma unroll
        for (int k = 1; k < 4; k += 2) {
            if (k == 3) __syncthreads();
            for (int q = threadIdx.x; q < L; q += kPlaneThreads) {
                const int w = __umulhi((unsigned)q, mH), h = q - w * H;
                tile[h * pitch + w] = (k == 1) ? s[1][q] : s[3][L - 1 - q];
            }
            __syncthreads();
            for (int p = threadIdx.x; p < L; p += kPlaneThreads) {
                const int h = __umulhi((unsigned)p, mW), w = p - h * W;
                o[k * ks + p] = tile[h * pitch + w];
            }
        }
    }
}

static bool plane_ok(int64_t H, int64_t W, int scans, int one_by_one, size_t esize) {
    // the index magic needs n * d < 2^32 for n < H*W, d = H or W; the plane (odd pitch) must fit the default 48 KB
    (void)one_by_one;
    return scans == XFS_SCANS_CROSS2D && H >= 2 && W >= 2 && H * W <= 16384 &&
           (size_t)H * (size_t)(W | 1) * esize <= kPlaneSmem;
}

// ---------------------------------------------------------------------------------------------------------
// swap scan / merge / stack: row copies, 16 bytes per thread when rows allow
// mode 0: scan  (x, x2) -> out (B,2,C,L)  even channels exchanged
// mode 1: merge ys -> (y, y2)             plain split
// mode 2: stack (y, y2) -> ys             plain stack
// ---------------------------------------------------------------------------------------------------------
template <typename V>
__global__ void __launch_bounds__(256) swap_kernel(const V* __restrict__ a, const V* __restrict__ b2, V* __restrict__ o0,
                                                   V* __restrict__ o1, int64_t C, int64_t Lv, int mode) {
    const int64_t bc = blockIdx.x;
    const int64_t b = bc / C, c = bc % C;
    const bool even = (c % 2 == 0);
    for (int64_t i = (int64_t)blockIdx.y * blockDim.x + threadIdx.x; i < Lv; i += (int64_t)gridDim.y * blockDim.x) {
        if (mode == 0) {
            const V va = a[bc * Lv + i], vb = b2[bc * Lv + i];
            o0[((b * 2 + 0) * C + c) * Lv + i] = even ? vb : va;
            o0[((b * 2 + 1) * C + c) * Lv + i] = even ? va : vb;
        } else if (mode == 1) {
            o0[bc * Lv + i] = a[((b * 2 + 0) * C + c) * Lv + i];
            o1[bc * Lv + i] = a[((b * 2 + 1) * C + c) * Lv + i];
        } else {
            o0[((b * 2 + 0) * C + c) * Lv + i] = a[bc * Lv + i];
            o0[((b * 2 + 1) * C + c) * Lv + i] = b2[bc * Lv + i];
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------------------------------------
int launch_cross_scan(const void* x, void* xs, int64_t B, int64_t C, int64_t H, int64_t W, int dtype, int scans,
                      int one_by_one, cudaStream_t st) {
    const size_t esize = dtype == XFS_F32 ? 4 : 2;
    if (plane_ok(H, W, scans, one_by_one, esize)) {
        const size_t smem = (size_t)H * (size_t)(W | 1) * esize;
        const unsigned mH = plane_magic((unsigned)H), mW = plane_magic((unsigned)W);
#define XFS_PLANE_SCAN(U, OBO) cross_scan_plane_kernel<U, OBO><<<(unsigned)(B * C), kPlaneThreads, smem, st>>>((const U*)x, (U*)xs, C, (int)H, (int)W, mH, mW)
        if (dtype == XFS_F32) { if (one_by_one) XFS_PLANE_SCAN(uint32_t, true); else XFS_PLANE_SCAN(uint32_t, false); }
        else { if (one_by_one) XFS_PLANE_SCAN(uint16_t, true); else XFS_PLANE_SCAN(uint16_t, false); }
#undef XFS_PLANE_SCAN
        return check_launch();
    }
    dim3 grid((unsigned)(B * C), (unsigned)((H + kTile - 1) / kTile), (unsigned)((W + kTile - 1) / kTile));
    dim3 block(kTile, kRows);
    if (dtype == XFS_F32)
        cross_scan_kernel<uint32_t><<<grid, block, 0, st>>>((const uint32_t*)x, (uint32_t*)xs, C, (int)H, (int)W, scans, one_by_one);
    else
        cross_scan_kernel<uint16_t><<<grid, block, 0, st>>>((const uint16_t*)x, (uint16_t*)xs, C, (int)H, (int)W, scans, one_by_one);
    return check_launch();
}

int launch_cross_merge(const void* ys, void* y, int64_t B, int64_t C, int64_t H, int64_t W, int dtype, int scans,
                       int one_by_one, cudaStream_t st) {
    const size_t esize = dtype == XFS_F32 ? 4 : 2;
    if (plane_ok(H, W, scans, one_by_one, esize)) {
        const size_t smem = (size_t)H * (size_t)(W | 1) * esize;
        const unsigned mH = plane_magic((unsigned)H), mW = plane_magic((unsigned)W);
#define XFS_PLANE_MERGE(T, OBO) cross_merge_plane_kernel<T, OBO><<<(unsigned)(B * C), kPlaneThreads, smem, st>>>((const T*)ys, (T*)y, C, (int)H, (int)W, mH, mW)
        if (dtype == XFS_F32) { if (one_by_one) XFS_PLANE_MERGE(float, true); else XFS_PLANE_MERGE(float, false); }
        else if (dtype == XFS_BF16) { if (one_by_one) XFS_PLANE_MERGE(__nv_bfloat16, true); else XFS_PLANE_MERGE(__nv_bfloat16, false); }
        else { if (one_by_one) XFS_PLANE_MERGE(__half, true); else XFS_PLANE_MERGE(__half, false); }
#undef XFS_PLANE_MERGE
        return check_launch();
    }
    dim3 grid((unsigned)(B * C), (unsigned)((H + kTile - 1) / kTile), (unsigned)((W + kTile - 1) / kTile));
    dim3 block(kTile, kRows);
    if (dtype == XFS_F32)
        cross_merge_kernel<float><<<grid, block, 0, st>>>((const float*)ys, (float*)y, C, (int)H, (int)W, scans, one_by_one);
    else if (dtype == XFS_BF16)
        cross_merge_kernel<__nv_bfloat16><<<grid, block, 0, st>>>((const __nv_bfloat16*)ys, (__nv_bfloat16*)y, C, (int)H, (int)W, scans, one_by_one);
    else
        cross_merge_kernel<__half><<<grid, block, 0, st>>>((const __half*)ys, (__half*)y, C, (int)H, (int)W, scans, one_by_one);
    return check_launch();
}

int launch_swap(const void* a, const void* b2, void* o0, void* o1, int64_t B, int64_t C, int64_t L, int dtype, int mode,
                cudaStream_t st) {
    const int64_t esize = (dtype == XFS_F32) ? 4 : 2;
    const int64_t row_bytes = L * esize;
    auto al16 = [](const void* p) { return p == nullptr || (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
    const bool v16 = (row_bytes % 16 == 0) && al16(a) && al16(b2) && al16(o0) && al16(o1);
    dim3 grid((unsigned)(B * C), 1);
    if (v16) {
        const int64_t Lv = row_bytes / 16;
        grid.y = (unsigned)((Lv + 255) / 256 > 64 ? 64 : (Lv + 255) / 256);
        swap_kernel<uint4><<<grid, 256, 0, st>>>((const uint4*)a, (const uint4*)b2, (uint4*)o0, (uint4*)o1, C, Lv, mode);
    } else if (esize == 4) {
        grid.y = (unsigned)((L + 255) / 256 > 64 ? 64 : (L + 255) / 256);
        swap_kernel<uint32_t><<<grid, 256, 0, st>>>((const uint32_t*)a, (const uint32_t*)b2, (uint32_t*)o0, (uint32_t*)o1, C, L, mode);
    } else {
        grid.y = (unsigned)((L + 255) / 256 > 64 ? 64 : (L + 255) / 256);
        swap_kernel<uint16_t><<<grid, 256, 0, st>>>((const uint16_t*)a, (const uint16_t*)b2, (uint16_t*)o0, (uint16_t*)o1, C, L, mode);
    }
    return check_launch();
}

}  // namespace xfs
