// ss2d_tiles.cuh -- shared-memory staging of one (batch, channel) image for the fused SS2D kernels.
//
// A channel image is kept twice in shared memory, indexed by POSITION: row-major (p = h*W + w, routes 0/2) and
// column-major (q = w*H + h, routes 1/3).  Both copies use the same 128-byte XOR swizzle on 16-byte granules so that the
// scan's access pattern -- lane i reads/writes its 8 consecutive positions as two float4 -- is bank-conflict free.
// Staging moves BSxBS spatial blocks per thread (BS = 4, 2 or 1 depending on what H, W and the pointer alignment
// allow): BS row vectors are read from global memory (coalesced along w), written as BS row vectors to the row-major
// copy and, transposed in registers, as BS column vectors to the column-major copy.  The output merge does the reverse.
#pragma once

#include "xfscan_common.cuh"

namespace xfs {

__device__ __forceinline__ int swz_f4(int f) { return f ^ ((f >> 3) & 7); }
__device__ __forceinline__ int swz_pos(int p) { return (swz_f4(p >> 2) << 2) | (p & 3); }

// buffer length: L rounded up to the swizzle period (32 floats); lanes whose 8 positions start at or beyond it skip
// shared memory altogether
__host__ __device__ inline int64_t buf_len(int64_t L) { return ((L + 31) / 32) * 32; }

// lane's 8 consecutive positions starting at granule f4s (already swizzled); the partner granule is f4s ^ 1
__device__ __forceinline__ void lds8(const float* buf, int f4s, float (&v)[8]) {
    const float4 a = *reinterpret_cast<const float4*>(buf + (f4s << 2));
    const float4 b = *reinterpret_cast<const float4*>(buf + ((f4s ^ 1) << 2));
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void sts8(float* buf, int f4s, const float (&v)[8]) {
    *reinterpret_cast<float4*>(buf + (f4s << 2)) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(buf + ((f4s ^ 1) << 2)) = make_float4(v[4], v[5], v[6], v[7]);
}

// ---- BS-wide global vectors <-> floats -----------------------------------------------------------------------
template <typename T, int BS>
__device__ __forceinline__ void ldg_vec(const T* __restrict__ p, float (&v)[BS]) {
    if constexpr (BS == 1) {
        v[0] = Elem<T>::to_f(*p);
    } else if constexpr (sizeof(T) == 4) {
        if constexpr (BS == 4) { const float4 a = __ldg(reinterpret_cast<const float4*>(p)); v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; }
        else { const float2 a = __ldg(reinterpret_cast<const float2*>(p)); v[0] = a.x; v[1] = a.y; }
    } else {
        if constexpr (BS == 4) {
            const uint2 a = __ldg(reinterpret_cast<const uint2*>(p));
            const T* e = reinterpret_cast<const T*>(&a);
#pragma unroll
            for (int i = 0; i < 4; ++i) v[i] = Elem<T>::to_f(e[i]);
        } else {
            const unsigned a = __ldg(reinterpret_cast<const unsigned*>(p));
            const T* e = reinterpret_cast<const T*>(&a);
            v[0] = Elem<T>::to_f(e[0]); v[1] = Elem<T>::to_f(e[1]);
        }
    }
}

template <typename T, int BS>
__device__ __forceinline__ void stg_vec(T* __restrict__ p, const float (&v)[BS]) {
    if constexpr (BS == 1) {
        *p = Elem<T>::from_f(v[0]);
    } else if constexpr (sizeof(T) == 4) {
        if constexpr (BS == 4) *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
        else *reinterpret_cast<float2*>(p) = make_float2(v[0], v[1]);
    } else {
        if constexpr (BS == 4) {
            uint2 a;
            T* e = reinterpret_cast<T*>(&a);
#pragma unroll
            for (int i = 0; i < 4; ++i) e[i] = Elem<T>::from_f(v[i]);
            *reinterpret_cast<uint2*>(p) = a;
        } else {
            unsigned a;
            T* e = reinterpret_cast<T*>(&a);
            e[0] = Elem<T>::from_f(v[0]); e[1] = Elem<T>::from_f(v[1]);
            *reinterpret_cast<unsigned*>(p) = a;
        }
    }
}

// BS consecutive positions starting at p (p % BS == 0, so they never straddle a 16-byte granule)
template <int BS>
__device__ __forceinline__ void sts_vec(float* buf, int p, const float (&v)[BS]) {
    float* a = buf + swz_pos(p);
    if constexpr (BS == 4) *reinterpret_cast<float4*>(a) = make_float4(v[0], v[1], v[2], v[3]);
    else if constexpr (BS == 2) *reinterpret_cast<float2*>(a) = make_float2(v[0], v[1]);
    else *a = v[0];
}
template <int BS>
__device__ __forceinline__ void lds_vec(const float* buf, int p, float (&v)[BS]) {
    const float* a = buf + swz_pos(p);
    if constexpr (BS == 4) { const float4 t = *reinterpret_cast<const float4*>(a); v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w; }
    else if constexpr (BS == 2) { const float2 t = *reinterpret_cast<const float2*>(a); v[0] = t.x; v[1] = t.y; }
    else v[0] = *a;
}

// largest block size the geometry and the pointer alignment allow
template <typename T>
__device__ __forceinline__ int pick_bs(const T* base, int H, int W) {
    const uintptr_t a = reinterpret_cast<uintptr_t>(base);
    if (H % 4 == 0 && W % 4 == 0 && a % (4 * sizeof(T)) == 0) return 4;
    if (H % 2 == 0 && W % 2 == 0 && a % (2 * sizeof(T)) == 0) return 2;
    return 1;
}

// ---- global image -> (row-major, column-major) swizzled copies ------------------------------------------------
template <typename T, int BS>
__device__ __forceinline__ void stage_image_bs(const T* __restrict__ img, float* __restrict__ bN, float* __restrict__ bT,
                                               int H, int W, bool valid, int tid, int nthreads) {
    const int bw_n = W / BS, nblk = (H / BS) * bw_n;
    for (int blk = tid; blk < nblk; blk += nthreads) {
        const int bh = blk / bw_n, bw = blk - bh * bw_n;
        const int h0 = bh * BS, w0 = bw * BS;
        float r[BS][BS];
#pragma unroll
        for (int i = 0; i < BS; ++i) {
            if (valid) ldg_vec<T, BS>(img + (h0 + i) * W + w0, r[i]);
            else {
#pragma unroll
                for (int c = 0; c < BS; ++c) r[i][c] = 0.0f;
            }
        }
#pragma unroll
        for (int i = 0; i < BS; ++i) sts_vec<BS>(bN, (h0 + i) * W + w0, r[i]);
#pragma unroll
        for (int c = 0; c < BS; ++c) {
            float col[BS];
#pragma unroll
            for (int i = 0; i < BS; ++i) col[i] = r[i][c];
            sts_vec<BS>(bT, (w0 + c) * H + h0, col);
        }
    }
}

template <typename T>
__device__ __forceinline__ void stage_image(const T* __restrict__ img, float* __restrict__ bN, float* __restrict__ bT,
                                            int H, int W, int L, int Lb, bool valid, int tid, int nthreads) {
    const int bs = pick_bs(img, H, W);
    if (bs == 4) stage_image_bs<T, 4>(img, bN, bT, H, W, valid, tid, nthreads);
    else if (bs == 2) stage_image_bs<T, 2>(img, bN, bT, H, W, valid, tid, nthreads);
    else stage_image_bs<T, 1>(img, bN, bT, H, W, valid, tid, nthreads);
    for (int p = L + tid; p < Lb; p += nthreads) {      // tail of BOTH layouts: u = 0 there
        bN[swz_pos(p)] = 0.0f;
        bT[swz_pos(p)] = 0.0f;
    }
}

// ---- out[p] = aN[p] + aT[w*H + h]  (row-major copy + column-major copy, spatial order out) --------------------
template <typename TO, int BS>
__device__ __forceinline__ void merge_out_bs(TO* __restrict__ out, const float* __restrict__ aN, const float* __restrict__ aT,
                                             int H, int W, int tid, int nthreads) {
    const int bw_n = W / BS, nblk = (H / BS) * bw_n;
    for (int blk = tid; blk < nblk; blk += nthreads) {
        const int bh = blk / bw_n, bw = blk - bh * bw_n;
        const int h0 = bh * BS, w0 = bw * BS;
        float col[BS][BS];
#pragma unroll
        for (int c = 0; c < BS; ++c) lds_vec<BS>(aT, (w0 + c) * H + h0, col[c]);
#pragma unroll
        for (int i = 0; i < BS; ++i) {
            float r[BS];
            lds_vec<BS>(aN, (h0 + i) * W + w0, r);
#pragma unroll
            for (int c = 0; c < BS; ++c) r[c] += col[c][i];
            stg_vec<TO, BS>(out + (h0 + i) * W + w0, r);
        }
    }
}

template <typename TO>
__device__ __forceinline__ void merge_out(TO* __restrict__ out, const float* __restrict__ aN, const float* __restrict__ aT,
                                          int H, int W, int tid, int nthreads) {
    const int bs = pick_bs(out, H, W);
    if (bs == 4) merge_out_bs<TO, 4>(out, aN, aT, H, W, tid, nthreads);
    else if (bs == 2) merge_out_bs<TO, 2>(out, aN, aT, H, W, tid, nthreads);
    else merge_out_bs<TO, 1>(out, aN, aT, H, W, tid, nthreads);
}

}  // namespace xfs
