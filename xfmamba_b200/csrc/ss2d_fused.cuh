// ss2d_fused.cuh -- shared pieces of the fused SS2D kernels (ss2d_fwd.cu, ss2d_bwd.cu).
//
// The fused kernels replace the operator sequence of SS2Dv2.forward_corev2 (reference
// models/fusion_vmamba.py:1145,1170-1174):
//     xs = cross_scan_fn(x); ys = selective_scan_fn(xs, dts, As, Bs, Cs, Ds, delta_bias, True); y = cross_merge_fn(ys)
// which moves ~22 elements per (b, d, l) through HBM (xs written 4x and re-read, ys written 4x and re-read).  Fused, x is
// read once, delta/B/C are streamed once and the merged y is written once: 6 elements per (b, d, l) + B/C.
//
// Work decomposition (both directions)
//   CTA  = one batch image x kCh channel(s); 4 warps = the 4 CrossScan routes, all running concurrently.
//   Shared memory holds, per channel, the image in row-major order (routes 0/2) and column-major order (routes 1/3)
//   and one accumulator per orientation, all indexed by POSITION (ss2d_tiles.cuh).
//   Routes 2/3 walk the same 256-position chunks as routes 0/1 but from the far end with a reversed warp scan, so a
//   route and its flip never touch the same accumulator chunk in the same half of the walk: the first half stores,
//   then ONE 64-thread named barrier, then the second half read-modify-writes what the partner stored.  No atomics,
//   and the merge (y0 + y2) + (y1 + y3) is evaluated in the reference's order (models/csm_triton.py:61-62).
//   The S6 recurrence h_l = exp(dt_l A) h_{l-1} + dt_l B_l u_l is scanned per chunk with a warp-shuffle scan of affine
//   maps (xfscan_common.cuh); the state is carried in a register (N == 1) or in shared memory (N > 1).
//   delta/B/C of the NEXT chunk are loaded (128-bit) into a second register set while the current chunk is computed;
//   element-wise arithmetic runs on packed fp32 pairs (FFMA2/FMUL2/FADD2).
//
// Two things learned from ncu that shape the code:
//   * ptxas tracks outstanding global loads with a few COUNTING scoreboards.  If the next chunk's loads are issued
//     before the current chunk's load registers are first read, both batches can share a scoreboard and that first read
//     waits ~1000 cycles for the loads just issued.  So each chunk first reads every load register once
//     (dt -> dt + bias, B -> B*u, C -> C + 0), THEN issues the next loads.
//   * scalar edge code inlined into the hot loops made the backward 127 KB of SASS and stall on instruction fetch.
//     kFast kernels (rows 16-byte aligned, L a multiple of the vector width) contain no scalar edge code.
#pragma once

#include <cstdlib>
#include <type_traits>

#include "ss2d_tiles.cuh"

namespace xfs {

constexpr int kFusedMaxState = 64;           // states carried in smem for N > 1
constexpr size_t kSmemLimit = 227 * 1024;

__device__ __forceinline__ void pair_barrier(int pair) {   // the 2 warps of a route pair (routes k and k+2)
    asm volatile("bar.sync %0, 64;" ::"r"(pair + 1) : "memory");
}
// CTA-wide barrier 0, issued from inside the two instantiations of the route walk (per-thread arrival on sm_70+)
__device__ __forceinline__ void cta_barrier() { asm volatile("bar.sync 0;" ::: "memory"); }

__device__ __forceinline__ void pack8(const float (&v)[8], f2 (&o)[4]) {
#pragma unroll
    for (int j = 0; j < 4; ++j) o[j] = make_float2(v[2 * j], v[2 * j + 1]);
}
__device__ __forceinline__ void unpack8(const f2 (&v)[4], float (&o)[8]) {
#pragma unroll
    for (int j = 0; j < 4; ++j) { o[2 * j] = v[j].x; o[2 * j + 1] = v[j].y; }
}
// address order -> position order (a flip for routes 2/3); never writes the source, so it is pure register renaming
template <bool kRev>
__device__ __forceinline__ void to_pos(const float (&src)[8], float (&dst)[8]) {
#pragma unroll
    for (int i = 0; i < 8; ++i) dst[i] = src[kRev ? 7 - i : i];
}

// one 8-position row access of the streamed tensors; kFast: vector-only code (see load8_fast)
template <typename T, bool kFast>
__device__ __forceinline__ void row_load8(const T* __restrict__ row, int l0, int L, bool vec_ok, float (&v)[8]) {
    if constexpr (kFast) load8_fast<T>(row, l0, L, v);
    else load8<T, true>(row, l0, L, vec_ok, v);
}
template <typename T, bool kFast>
__device__ __forceinline__ void row_store8(T* __restrict__ row, int l0, int L, bool vec_ok, const float (&v)[8]) {
    if constexpr (kFast) store8_fast<T>(row, l0, L, v);
    else store8<T>(row, l0, L, vec_ok, v);
}

// ---- L2 prefetch of the streamed rows ----------------------------------------------------------------------------
// The register prefetch is one chunk deep (more would cost registers, i.e. occupancy), which leaves too few bytes in
// flight to cover DRAM latency.  ONE extra instruction per chunk pulls the delta / B / C rows of a chunk further ahead
// into L2 (lanes 0-7 / 8-15 / 16-23 take the 128-byte lines of one row each), so the register loads that follow hit L2.
constexpr int kPrefetchAhead = 2;    // chunks beyond the one being loaded into registers

template <typename T>
__device__ __forceinline__ void prefetch_chunk_l2(const T* lane_row, bool flipped, int j, int nch, int L, int lane) {
    constexpr int kLine = 128 / (int)sizeof(T);                    // elements per line
    const int start = flipped ? L - kChunk * (j + 1) : kChunk * j; // scan index of the chunk's lowest address
    const int off = start + (lane & 7) * kLine;
    if (lane_row != nullptr && j >= 0 && j < nch && (lane & 7) * kLine < kChunk && off >= 0 && off < L)
        asm volatile("prefetch.global.L2 [%0];" ::"l"(lane_row + off));
}

// ---- short sequences (L <= 64): one sequence per 8-lane group, see ss2d_small.cu ---------------------------------
constexpr int kSmallL = 64;          // positions per sequence slot
constexpr int kSmallMaxN = 16;
constexpr int kQuad = 4;             // channels per warp (one per 8-lane group)
constexpr int kQuadsPerCta = 8;      // backward: quads walked by one CTA

// inclusive scan of affine maps over the 8 lanes of a group; returns the state ENTERING this lane (sequence starts at 0)
template <bool kRev>
__device__ __forceinline__ float group_prefix(float P, float S, int j, float& seq_out) {
#pragma unroll
    for (int off = 1; off < 8; off <<= 1) {
        const float Pn = kRev ? __shfl_down_sync(kFull, P, off, 8) : __shfl_up_sync(kFull, P, off, 8);
        const float Sn = kRev ? __shfl_down_sync(kFull, S, off, 8) : __shfl_up_sync(kFull, S, off, 8);
        const bool has = kRev ? (j + off < 8) : (j >= off);
        if (has) { S = fmaf(P, Sn, S); P = P * Pn; }
    }
    seq_out = __shfl_sync(kFull, S, kRev ? 0 : 7, 8);          // state after the whole sequence (carry-in is 0)
    const float prev = kRev ? __shfl_down_sync(kFull, S, 1, 8) : __shfl_up_sync(kFull, S, 1, 8);
    const bool first = kRev ? (j == 7) : (j == 0);
    return first ? 0.0f : prev;
}

__device__ __forceinline__ float group_sum(float v) {            // sum over the 8 lanes of a group
#pragma unroll
    for (int off = 4; off > 0; off >>= 1) v += __shfl_xor_sync(kFull, v, off, 8);
    return v;
}
__device__ __forceinline__ float quad_sum(float v) {             // sum over the 4 groups of a warp (same j)
    v += __shfl_xor_sync(kFull, v, 8);
    v += __shfl_xor_sync(kFull, v, 16);
    return v;
}

// stage kQuad channel images (L <= 64 each) into row-major / column-major swizzled rows of 64 floats.  The loads of a thread
// are issued as one batch before its first shared-memory store (see staged_loop in fusion_small.cu for why).
template <typename T>
__device__ __forceinline__ void stage_quad(const T* __restrict__ base, int64_t chan_stride, int nvalid, float* bN, float* bT,
                                           int H, int W, int L, int tid) {
    constexpr int kIt = kQuad * kSmallL / 128;
    float v[kIt];
#pragma unroll
    for (int u = 0; u < kIt; ++u) {
        const int idx = tid + u * 128, ch = idx >> 6, p = idx & 63;
        v[u] = (ch < nvalid && p < L) ? Elem<T>::to_f(base[ch * chan_stride + p]) : 0.0f;
    }
#pragma unroll
    for (int u = 0; u < kIt; ++u) {
        const int idx = tid + u * 128, ch = idx >> 6, p = idx & 63;
        bN[ch * kSmallL + swz_pos(p)] = v[u];
        if (p < L) {
            const int h = p / W, w = p - h * W;
            bT[ch * kSmallL + swz_pos(w * H + h)] = v[u];
        } else {
            bT[ch * kSmallL + swz_pos(p)] = 0.0f;
        }
    }
}

inline size_t fwd_smem(int64_t L, int64_t N, int ch) {
    return sizeof(float) * (size_t)(4 * ch * buf_len(L) + (N == 1 ? 0 : 4 * ch * kFusedMaxState));
}
inline size_t bwd_smem(int64_t L, int64_t N, int ch) {
    return sizeof(float) * (size_t)(6 * ch * buf_len(L) + (N == 1 ? 0 : 2 * 4 * ch * kFusedMaxState));
}

template <typename K>
inline int set_smem(K kernel, size_t bytes) {
    return (int)cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }
__device__ __forceinline__ bool aligned16_dev(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace xfs
