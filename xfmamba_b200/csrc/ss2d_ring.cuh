// ss2d_ring.cuh -- shared pieces of the TMA-fed fused SS2D kernels (ss2d_ring_fwd.cu, ss2d_ring_bwd.cu).
//
// Same work decomposition as ss2d_fused.cuh (CTA = one (batch, channel) image, warp k = CrossScan route k, 256-position
// chunks, warp-shuffle scan of affine maps, pair protocol for the accumulators), but the streamed operands no longer
// pass through per-thread global loads:
//   * delta / B / C chunk rows (1 KB each for fp32) are copied global -> shared memory by the TMA engine
//     (cp.async.bulk, SASS UBLKCP) into a per-warp ring of slots, completion signalled on one mbarrier per slot; ONE
//     lane issues the three copies of a chunk, every lane then reads its 8 positions with two LDS.128 at constant
//     offsets.  No per-thread 64-bit address arithmetic, no range clamps, no load registers held across a chunk.
//   * the (batch, channel) images (x; dy in the backward) arrive by ONE bulk copy each, straight into the row-major
//     image buffer; the column-major copy for routes 1/3 is a shared->shared transposition.
//
// Shared-memory access pattern ("alternating halves").  A lane owns 8 consecutive positions = two 16-byte granules
// (2i, 2i+1) of a LINEAR row, which is what a bulk copy produces.  Reading granule 2i in one LDS.128 and 2i+1 in the
// next would put lanes i and i+4 on the same banks (2-way conflict).  Instead lane i reads granule 2i + s first and
// 2i + 1 - s second, s = bit 2 of the lane index: both instructions are conflict free.  The registers of the first
// instruction are the "A half" (positions p0..p0+3 if s == 0, p0+4..p0+7 if s == 1), those of the second the "B half".
// Every row (delta, B, C, u, dy, accumulators) is accessed the same way, so element-wise arithmetic never cares which
// half is which; only the sequential fold does: each half is folded on its own (two independent 4-long chains instead
// of one 8-long chain) and the two composites are combined in the lane-dependent order.
#pragma once

#include "ss2d_fused.cuh"

namespace xfs {
namespace ring {

constexpr int kRows = 3;                         // delta, B, C
constexpr int kChunkBytesF32 = kChunk * 4;       // 1 KB

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void pair_barrier_n(int pair, int nthreads) {   // all warps of a route pair (routes k and k+2)
    asm volatile("bar.sync %0, %1;" ::"r"(pair + 1), "r"(nthreads) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {}
}
// TMA bulk copy global -> shared (SASS UBLKCP); completes `bytes` on the mbarrier.  src/dst 16-byte aligned, bytes % 16 == 0
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
                 "r"(bytes), "r"(bar)
                 : "memory");
}
// TMA bulk copy shared -> global and element-wise reduction shared -> global (fp32 add at L2), bulk-group completion
__device__ __forceinline__ void bulk_s2g(void* dst, uint32_t src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_red_add_f32(float* dst, uint32_t src, uint32_t bytes) {
    asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes)
                 : "memory");
}
// asynchronous L2 prefetch of a contiguous range (one instruction, no destination); 16-byte aligned, bytes % 16 == 0
__device__ __forceinline__ void bulk_prefetch_l2(const void* src, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int kPending>
__device__ __forceinline__ void bulk_wait_read() {
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(kPending) : "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// true in exactly one lane of the (converged) warp; the compiler then issues what follows from uniform registers
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ float4 add4(float4 a, float4 b) {           // two FADD2
    const f2 lo = add2(make_float2(a.x, a.y), make_float2(b.x, b.y)), hi = add2(make_float2(a.z, a.w), make_float2(b.z, b.w));
    return make_float4(lo.x, lo.y, hi.x, hi.y);
}

__device__ __forceinline__ float4 lds128(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, float4 v) {
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// One half (4 positions) as two packed pairs, ascending POSITION order.  kFlip: the granule holds them in descending
// order (a flipped route's rows are stored in scan order) -> register renaming only.
template <bool kFlip>
__device__ __forceinline__ void half_from(const float4 g, f2& lo, f2& hi) {
    if (!kFlip) { lo = make_float2(g.x, g.y); hi = make_float2(g.z, g.w); }
    else { lo = make_float2(g.w, g.z); hi = make_float2(g.y, g.x); }
}
template <bool kFlip>
__device__ __forceinline__ float4 half_to(const f2 lo, const f2 hi) {
    return kFlip ? make_float4(hi.y, hi.x, lo.y, lo.x) : make_float4(lo.x, lo.y, hi.x, hi.y);
}

// Swizzle of the column-major image copies.  ss2d_tiles.cuh XORs granule bit 0 as well, which is exactly the alternation
// the A/B access order already provides -- combined they cancel and put lanes i and i+4 back on the same banks (measured:
// 6 instead of 4 wavefronts per LDS.128/STS.128 on those buffers).  Here only granule bits 1-2 are XORed: the lane access
// stays conflict free, the transposing 16-byte stores (granule stride H) spread over 4 bank groups (2-way).
__device__ __forceinline__ int rswz_f4(int f) { return f ^ (((f >> 3) & 3) << 1); }
__device__ __forceinline__ int rswz_pos(int p) { return (rswz_f4(p >> 2) << 2) | (p & 3); }

// per-lane byte offsets (within a 256-position chunk) of the A and B granules
struct LaneOffsets {
    uint32_t imgA, imgB;     // position-indexed image / accumulator buffers of this route's orientation
    uint32_t rowA, rowB;     // ring rows (scan order of the route: descending position for a flipped route)
    bool a_first;            // the A half precedes the B half in the route's scan order
    int posA, posB;          // first position of each half, relative to the chunk start
};

__device__ __forceinline__ LaneOffsets lane_offsets(int lane, bool rev, bool transposed) {
    LaneOffsets o;
    const int s = (lane >> 2) & 1;
    const int f = 2 * lane;                      // granule of positions p0 .. p0+3
    o.posA = 8 * lane + 4 * s;
    o.posB = 8 * lane + 4 * (1 - s);
    if (!transposed) {                           // linear (what the bulk copy of the image wrote)
        o.imgA = (uint32_t)(f + s) * 16u;
        o.imgB = (uint32_t)(f + 1 - s) * 16u;
    } else {                                     // column-major copy: XOR-swizzled granules (ss2d_tiles.cuh)
        o.imgA = (uint32_t)rswz_f4(f + s) * 16u;
        o.imgB = (uint32_t)rswz_f4(f + 1 - s) * 16u;
    }
    if (!rev) {
        o.rowA = (uint32_t)(f + s) * 16u;
        o.rowB = (uint32_t)(f + 1 - s) * 16u;
        o.a_first = (s == 0);
    } else {                                     // slot index t <-> position 255 - t: granule 63 - f holds p0+3 .. p0
        o.rowA = (uint32_t)(s ? 62 - f : 63 - f) * 16u;
        o.rowB = (uint32_t)(s ? 63 - f : 62 - f) * 16u;
        o.a_first = (s == 1);
    }
    return o;
}

// ---- column-major (swizzled) copy of a row-major LINEAR image that already sits in shared memory -----------------
__device__ __forceinline__ void transpose_image(const float* __restrict__ bN, float* __restrict__ bT, int H, int W, int L, int Lb,
                                                int tid, int nthreads) {
    if (((H | W) & 3) == 0) {
        const int bw_n = W >> 2, nblk = (H >> 2) * bw_n;
        for (int blk = tid; blk < nblk; blk += nthreads) {
            const int bh = blk / bw_n, bw = blk - bh * bw_n;
            const int h0 = bh << 2, w0 = bw << 2;
            float r[4][4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float4 t = *reinterpret_cast<const float4*>(bN + (h0 + i) * W + w0);
                r[i][0] = t.x; r[i][1] = t.y; r[i][2] = t.z; r[i][3] = t.w;
            }
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                *reinterpret_cast<float4*>(bT + rswz_pos((w0 + c) * H + h0)) = make_float4(r[0][c], r[1][c], r[2][c], r[3][c]);
            }
        }
    } else {
        for (int q = tid; q < L; q += nthreads) {
            const int w = q / H, h = q - w * H;
            bT[rswz_pos(q)] = bN[h * W + w];
        }
    }
    for (int p = L + tid; p < Lb; p += nthreads) bT[rswz_pos(p)] = 0.0f;
}

// ---- out[p] = aN[p] + aT[w*H + h]: aN linear, aT swizzled ----------------------------------------------------------
template <typename TO>
__device__ __forceinline__ void merge_out_linear(TO* __restrict__ out, const float* __restrict__ aN, const float* __restrict__ aT,
                                                 int H, int W, int L, int tid, int nthreads) {
    if (((H | W) & 3) == 0 && (reinterpret_cast<uintptr_t>(out) % (4 * sizeof(TO))) == 0) {
        const int bw_n = W >> 2, nblk = (H >> 2) * bw_n;
        for (int blk = tid; blk < nblk; blk += nthreads) {
            const int bh = blk / bw_n, bw = blk - bh * bw_n;
            const int h0 = bh << 2, w0 = bw << 2;
            float col[4][4];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const float4 t = *reinterpret_cast<const float4*>(aT + rswz_pos((w0 + c) * H + h0));
                col[c][0] = t.x; col[c][1] = t.y; col[c][2] = t.z; col[c][3] = t.w;
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float4 t = *reinterpret_cast<const float4*>(aN + (h0 + i) * W + w0);
                float r[4] = {t.x + col[0][i], t.y + col[1][i], t.z + col[2][i], t.w + col[3][i]};
                stg_vec<TO, 4>(out + (h0 + i) * W + w0, r);
            }
        }
    } else {
        for (int pp = tid; pp < L; pp += nthreads) {
            const int h = pp / W, w = pp - h * W;
            out[pp] = Elem<TO>::from_f(aN[pp] + aT[rswz_pos(w * H + h)]);
        }
    }
}

// min / max of 8 values (FMNMX3 on sm_100)
__device__ __forceinline__ float min8(const f2 (&v)[4]) {
    return fminf(fminf(fminf(v[0].x, v[0].y), fminf(v[1].x, v[1].y)), fminf(fminf(v[2].x, v[2].y), fminf(v[3].x, v[3].y)));
}
__device__ __forceinline__ float max8(const f2 (&v)[4]) {
    return fmaxf(fmaxf(fmaxf(v[0].x, v[0].y), fmaxf(v[1].x, v[1].y)), fmaxf(fmaxf(v[2].x, v[2].y), fmaxf(v[3].x, v[3].y)));
}

// exp(x) range in which softplus needs neither the small-argument series nor the x > 20 identity (see softplus_fwd):
// e >= 2^-6 keeps the relative error of lg2(1 + e) below 2e-5; e <= 2^28 means x < 19.41
constexpr float kEMin = 0.015625f;
constexpr float kEMax = 268435456.0f;

inline size_t ring_fwd_smem(int64_t L, int warps, int slots) {      // images + ring + mbarriers (image, ring, carry) + carry values
    return sizeof(float) * (size_t)(4 * buf_len(L)) + (size_t)(4 * warps * slots * kRows * kChunkBytesF32) +
           8 * (size_t)(1 + 4 * warps * slots + 8) + 8 * sizeof(float);
}
inline size_t ring_bwd_smem(int64_t L, int slots) {
    return sizeof(float) * (size_t)(6 * buf_len(L)) + (size_t)(4 * slots * kRows * kChunkBytesF32) + 8 * (size_t)(2 + 4 * slots);
}

}  // namespace ring
}  // namespace xfs
