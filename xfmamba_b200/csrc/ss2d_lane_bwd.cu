// ss2d_lane_bwd.cu -- fused SS2D backward for the shapes XFMamba's backbone spends its bytes on (N == 1, fp32 rows,
// 16-byte aligned, L % 4 == 0, more than one chunk), written around LANE-GRANULAR checkpoints.
//
// Same decomposition as ss2d_bwd.cu (CTA = one (batch, channel) image, warp k = CrossScan route k, six position-indexed
// image buffers, pair protocol for the du accumulators), but the forward kernel (ss2d_ring_fwd.cu) saves the state
// ENTERING every lane's 8 positions instead of one state per 256-position chunk (+0.5 B per (b,k,d,l) on 24 / 44).  With
// that the backward needs no forward re-scan across lanes: each lane replays its 8 positions with one FMA per position
// (h_i = a_i h_{i-1} + b_i), and only the adjoint  g_i = C_i dy_i + a_{i+1} g_{i+1}  still crosses lanes -- ONE
// warp-shuffle scan per chunk instead of two, no prefix products of the forward maps, two short dependent chains
// instead of two 8-deep affine folds.  Per (b,k,d,l): 4 MUFU (ex2, lg2, rcp, ex2), ~10 packed FFMA2/FMUL2/FADD2,
// 3 scalar chain operations, 3.75 for the adjoint scan.  Formulas as in ss2d_bwd.cu / selective_scan.cu:
//     du = D dy + g dt B      ddt = g (B u + A (h - b))      ddelta = ddt * sigmoid(delta + bias)
//     dA += g dt (h - b)      dB = g dt u                     dC = dy h          (b = dt B u, h - b = a h_prev)
// The reference computes the same quantities per thread-block chunk (selective_scan_bwd_kernel.cuh:126-274).
#include "ss2d_ring.cuh"

namespace xfs {
using namespace ring;

#ifndef XFS_LANE_DIAG
#define XFS_LANE_DIAG 0          // timing experiments only (results wrong): bit 0 = no dB / dC reductions, 1 = no ddelta stores,
                                 // 2 = no warp scan, 3 = MUFU replaced by FMA, 4 = no du accumulation, 5 = no walk at all
#endif

namespace {

__device__ __forceinline__ float dex2(float x) { return (XFS_LANE_DIAG & 8) ? fmaf(x, 0.5f, 1.0f) : ex2(x); }
__device__ __forceinline__ float dlg2(float x) { return (XFS_LANE_DIAG & 8) ? fmaf(x, 0.5f, -0.5f) : lg2(x); }
__device__ __forceinline__ float drcp(float x) { return (XFS_LANE_DIAG & 8) ? fmaf(x, -0.25f, 1.0f) : rcp(x); }

// All per-position arrays of the chunk loop are kept in MEMORY order (index 0 = lowest address of the lane's 8 elements
// of a streamed row), as four packed pairs.  Rows are stored in scan order, so memory order IS the route's forward scan
// order for all four routes: the replay chain runs over ascending indices, the adjoint over descending ones, and the
// ddelta / dB / dC results come out in store order.  Only the operands held in shared memory are position-indexed:
// for a flipped route memory element i is position 7 - i, i.e. pair q is pair 3 - q of the position-ordered granules
// with its halves swapped -- register renaming plus the HI_LO operand modifier of FFMA2 / FMUL2 / FADD2, no moves.
__device__ __forceinline__ f2 swp(const f2 v) { return make_float2(v.y, v.x); }
template <bool kRev>
__device__ __forceinline__ void reorder(const f2 (&v)[4], f2 (&o)[4]) {       // position order <-> memory order (an involution)
#pragma unroll
    for (int q = 0; q < 4; ++q) o[q] = kRev ? swp(v[3 - q]) : v[q];
}
__device__ __forceinline__ void unpack(const float4 g0, const float4 g1, f2 (&o)[4]) {
    o[0] = make_float2(g0.x, g0.y); o[1] = make_float2(g0.z, g0.w); o[2] = make_float2(g1.x, g1.y); o[3] = make_float2(g1.z, g1.w);
}

__device__ __forceinline__ float& el(f2 (&v)[4], int i) { return (i & 1) ? v[i >> 1].y : v[i >> 1].x; }

__device__ __forceinline__ float4 ldg128(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
// The streamed loads as VOLATILE asm: the compiler keeps them, in program order, in front of the warp vote that follows
// them in the chunk body, and ptxas does not sink a load below the branch that ends its basic block.
__device__ __forceinline__ float4 ldg128_pinned(const float* p) {
    float4 v;
    asm volatile("ld.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ float ldg32_pinned(const float* p) {
    float v;
    asm volatile("ld.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
    return v;
}
// 256-bit global accesses (sm_100: LDG.E.256 / STG.E.256): a lane's 8 positions = one 32-byte sector = one instruction
__device__ __forceinline__ void ldg256_pinned(const float* p, float4& a, float4& b) {
    asm volatile("ld.global.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w) : "l"(p) : "memory");
}
__device__ __forceinline__ void stg256(float* p, f2 v0, f2 v1, f2 v2, f2 v3) {
    asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "f"(v0.x), "f"(v0.y), "f"(v1.x), "f"(v1.y),
                 "f"(v2.x), "f"(v2.y), "f"(v3.x), "f"(v3.y) : "memory");
}
__device__ __forceinline__ void stg128(float* p, f2 lo, f2 hi) {
    asm volatile("st.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(lo.x), "f"(lo.y), "f"(hi.x), "f"(hi.y) : "memory");
}

// dB / dC reductions in FULL 32-byte sectors.  A lane's 8 values are one sector, but red.global.add takes 16 bytes per lane:
// issued per lane the two instructions each touch HALF of 32 sectors, and the L1 sends every half-filled sector to L2 (ncu:
// 77 M reduction sectors for 38.5 M sectors of payload, 73 % of all bytes the kernel pushes through the L1->L2 crossbar).  Lanes
// 2q and 2q+1 therefore swap one granule (4 shuffles), so that the first instruction covers the even lanes' sectors completely
// and the second the odd lanes'.  `step` = offset of lane l+1 relative to lane l (+8 / -8 elements).
template <int kStep>
__device__ __forceinline__ void red_sector_pair(float* __restrict__ row, int o, int lane, const f2 (&v)[4]) {
    const bool odd = lane & 1;
    const f2 s0 = odd ? v[0] : v[2], s1 = odd ? v[1] : v[3];          // even lanes send their high granule, odd their low one
    f2 r0, r1;
    r0.x = __shfl_xor_sync(kFull, s0.x, 1); r0.y = __shfl_xor_sync(kFull, s0.y, 1);
    r1.x = __shfl_xor_sync(kFull, s1.x, 1); r1.y = __shfl_xor_sync(kFull, s1.y, 1);
    const f2 a0 = odd ? r0 : v[0], a1 = odd ? r1 : v[1];              // even lane's sector: its low granule + (odd lane) its high one
    const f2 b0 = odd ? v[2] : r0, b1 = odd ? v[3] : r1;              // odd lane's sector: (even lane) its low granule + its high one
    red_add_v4(row + (odd ? o - kStep + 4 : o), a0.x, a0.y, a1.x, a1.y);
    red_add_v4(row + (odd ? o + 4 : o + kStep), b0.x, b0.y, b1.x, b1.y);
}

// x and dy images -> row-major and column-major swizzled copies, with ALL global loads of a thread's blocks (two 4x4 blocks
// of both images: 16 x 16 bytes) in flight before the first shared-memory store: one exposed memory round trip per CTA
// instead of four (ncu: 12 % of the kernel's stall samples sat on the first store after each batch of four loads).
__device__ __forceinline__ void stage_two_images(const float* __restrict__ x, const float* __restrict__ dy, float* xN, float* xT,
                                                 float* gN, float* gT, int H, int W, int L, int Lb, int tid) {
    // small images only (measured: 28x28 -3 %, 56x56 +2 %: there the four serial batches overlap with the other CTAs' walks)
    if (((H | W) & 3) == 0 && L <= 2048) {
        const int bw_n = W >> 2, nblk = (H >> 2) * bw_n;
        for (int base = 0; base < nblk; base += 256) {
            float4 rx[2][4], rg[2][4];
            int h0[2], w0[2];
            bool ok[2];
#pragma unroll
            for (int it = 0; it < 2; ++it) {
                const int blk = base + it * 128 + tid;
                ok[it] = blk < nblk;
                const int bh = blk / bw_n, bw = blk - bh * bw_n;
                h0[it] = bh << 2; w0[it] = bw << 2;
                if (ok[it]) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        rx[it][i] = ldg128(x + (h0[it] + i) * W + w0[it]);
                        rg[it][i] = ldg128(dy + (h0[it] + i) * W + w0[it]);
                    }
                }
            }
#pragma unroll
            for (int it = 0; it < 2; ++it) {
                if (ok[it]) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        *reinterpret_cast<float4*>(xN + swz_pos((h0[it] + i) * W + w0[it])) = rx[it][i];
                        *reinterpret_cast<float4*>(gN + swz_pos((h0[it] + i) * W + w0[it])) = rg[it][i];
                    }
                    const int q = w0[it] * H + h0[it];
                    *reinterpret_cast<float4*>(xT + swz_pos(q)) = make_float4(rx[it][0].x, rx[it][1].x, rx[it][2].x, rx[it][3].x);
                    *reinterpret_cast<float4*>(xT + swz_pos(q + H)) = make_float4(rx[it][0].y, rx[it][1].y, rx[it][2].y, rx[it][3].y);
                    *reinterpret_cast<float4*>(xT + swz_pos(q + 2 * H)) = make_float4(rx[it][0].z, rx[it][1].z, rx[it][2].z, rx[it][3].z);
                    *reinterpret_cast<float4*>(xT + swz_pos(q + 3 * H)) = make_float4(rx[it][0].w, rx[it][1].w, rx[it][2].w, rx[it][3].w);
                    *reinterpret_cast<float4*>(gT + swz_pos(q)) = make_float4(rg[it][0].x, rg[it][1].x, rg[it][2].x, rg[it][3].x);
                    *reinterpret_cast<float4*>(gT + swz_pos(q + H)) = make_float4(rg[it][0].y, rg[it][1].y, rg[it][2].y, rg[it][3].y);
                    *reinterpret_cast<float4*>(gT + swz_pos(q + 2 * H)) = make_float4(rg[it][0].z, rg[it][1].z, rg[it][2].z, rg[it][3].z);
                    *reinterpret_cast<float4*>(gT + swz_pos(q + 3 * H)) = make_float4(rg[it][0].w, rg[it][1].w, rg[it][2].w, rg[it][3].w);
                }
            }
        }
        for (int q = L + tid; q < Lb; q += 128) {
            const int sq = swz_pos(q);
            xN[sq] = 0.0f; xT[sq] = 0.0f; gN[sq] = 0.0f; gT[sq] = 0.0f;
        }
    } else {
        stage_image<float>(x, xN, xT, H, W, L, Lb, true, tid, 128);
        stage_image<float>(dy, gN, gT, H, W, L, Lb, true, tid, 128);
    }
}

// ---- streamed rows by element type.  fp32: two 16-byte granules per lane (or one 256-bit access, kV8); bf16 / f16: the lane's 8
// positions are ONE 16-byte granule (L % 8 == 0), kept as raw bits until they are consumed at the top of the next chunk.
template <typename T> struct RowRaw { float4 g0, g1; };
template <> struct RowRaw<__nv_bfloat16> { uint4 g; };
template <> struct RowRaw<__half> { uint4 g; };

__device__ __forceinline__ uint4 ldg128u_pinned(const void* p) {
    uint4 v;
    asm volatile("ld.global.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}
template <bool kV8>
__device__ __forceinline__ void row_load(RowRaw<float>& r, const float* row, unsigned o0, unsigned o1) {
    if (kV8) ldg256_pinned(row + o0, r.g0, r.g1);
    else { r.g0 = ldg128_pinned(row + o0); r.g1 = ldg128_pinned(row + o1); }
}
template <bool kV8, typename T>
__device__ __forceinline__ void row_load(RowRaw<T>& r, const T* row, unsigned o0, unsigned) { r.g = ldg128u_pinned(row + o0); }

__device__ __forceinline__ void row_unpack(const RowRaw<float>& r, f2 (&o)[4]) { unpack(r.g0, r.g1, o); }
__device__ __forceinline__ void row_unpack(const RowRaw<__nv_bfloat16>& r, f2 (&o)[4]) {
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&r.g);
#pragma unroll
    for (int i = 0; i < 4; ++i) o[i] = __bfloat1622float2(h[i]);
}
__device__ __forceinline__ void row_unpack(const RowRaw<__half>& r, f2 (&o)[4]) {
    const __half2* h = reinterpret_cast<const __half2*>(&r.g);
#pragma unroll
    for (int i = 0; i < 4; ++i) o[i] = __half22float2(h[i]);
}
// the chunk's ddelta: 8 values in memory order
template <bool kV8>
__device__ __forceinline__ void row_store(float* p, const f2 (&v)[4], bool ok0, bool ok1) {
    if (kV8) { if (ok0) stg256(p, v[0], v[1], v[2], v[3]); }
    else { if (ok0) stg128(p, v[0], v[1]); if (ok1) stg128(p + 4, v[2], v[3]); }
}
template <bool kV8>
__device__ __forceinline__ void row_store(__nv_bfloat16* p, const f2 (&v)[4], bool ok0, bool) {
    uint4 o;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __float22bfloat162_rn(v[i]);
    if (ok0) asm volatile("st.global.v4.b32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(o.x), "r"(o.y), "r"(o.z), "r"(o.w) : "memory");
}
template <bool kV8>
__device__ __forceinline__ void row_store(__half* p, const f2 (&v)[4], bool ok0, bool) {
    uint4 o;
    __half2* h = reinterpret_cast<__half2*>(&o);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __float22half2_rn(v[i]);
    if (ok0) asm volatile("st.global.v4.b32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(o.x), "r"(o.y), "r"(o.z), "r"(o.w) : "memory");
}

template <typename T>
struct LaneChunk {              // streamed operands of one chunk, as loaded (memory order)
    RowRaw<T> dt, B, C;
    float hin;
};

}  // namespace

// kV8: L % 8 == 0 and 32-byte aligned rows -- every lane's 8 positions are one aligned sector (256-bit loads and stores)
template <typename T, bool kSoftplus, bool kV8>
__global__ void __launch_bounds__(128, 3)
ss2d_lane_bwd_kernel(const xfs_ss2d_bwd_args p) {
    extern __shared__ __align__(16) float smem[];
    const int H = (int)p.H, W = (int)p.W, L = H * W;
    const int Lb = (int)buf_len(L);
    const int nch = (L + kChunk - 1) / kChunk;
    const int D = (int)p.D;
    // batch index fastest (see ss2d_bwd.cu): resident CTAs belong to different images, so their dB / dC rows differ (round 2
    // re-measured with the full-sector reductions: groups of 2 / 4 / 8 / D adjacent channels of one image first are within 1 %)
    const int b = blockIdx.x % (int)p.batch;
    const int d = blockIdx.x / (int)p.batch;
    const int tid = threadIdx.x, lane = tid & 31;
    const int k = __shfl_sync(kFull, tid >> 5, 0);                  // warp k = route k (provably warp-uniform)
    const bool transposed = k & 1;

    float* xN = smem;
    float* xT = xN + Lb;
    float* gN = xT + Lb;
    float* gT = gN + Lb;
    float* dN = gT + Lb;
    float* dT = dN + Lb;

    const int kd = k * D + d;
    const int64_t row = ((int64_t)b * 4 * D + kd) * L;
    const int64_t bc = ((int64_t)b * 4 + k) * L;
    const int rep = p.acc_replicas > 1 ? d % p.acc_replicas : 0;
    const int64_t acc = (((int64_t)rep * p.batch + b) * 4 + k) * L;
    const T* __restrict__ dt_row = reinterpret_cast<const T*>(p.delta) + row;
    T* __restrict__ ddt_row = reinterpret_cast<T*>(p.ddelta) + row;
    const T* __restrict__ B_row = reinterpret_cast<const T*>(p.Bs) + bc;
    const T* __restrict__ C_row = reinterpret_cast<const T*>(p.Cs) + bc;
    float* __restrict__ dB_row = p.dBs + acc;
    float* __restrict__ dC_row = p.dCs + acc;
    const float* __restrict__ st_row = p.states + ((int64_t)b * 4 * D + kd) * ((int64_t)nch * 32);
    // keep the seven row pointers in registers: left to itself the compiler re-derives them (64-bit multiplies) in every chunk
    asm volatile("" : "+l"(dt_row), "+l"(ddt_row), "+l"(B_row), "+l"(C_row), "+l"(dB_row), "+l"(dC_row), "+l"(st_row));

    const float bias = p.delta_bias ? p.delta_bias[kd] : 0.0f;
    const float bias_l2 = bias * kLog2e;
    const float Dd = p.Ds ? p.Ds[kd] : 0.0f;
    const float An = p.A[kd];
    const float A2 = An * kLog2e;

    // byte offsets of the lane's two granules inside a 1 KB chunk of a position-indexed (swizzled) image buffer
    const int f4l = (2 * lane) ^ ((lane >> 2) & 7);
    const uint32_t offA = (uint32_t)f4l * 16u, offB = (uint32_t)(f4l ^ 1) * 16u;
    const uint32_t xb = s32(transposed ? xT : xN), gb = s32(transposed ? gT : gN), db = s32(transposed ? dT : dN);

    // last chunk (the only one with positions >= L): which granules exist
    const int s0_last = (nch - 1) * kChunk + 8 * lane;
    const bool ok_lo = s0_last + 4 <= L, ok_hi = s0_last + 8 <= L;      // positions s0..s0+3 / s0+4..s0+7
    const bool in_buf_last = s0_last < Lb;

    const int m = nch / 2;             // routes 2/3 first touch chunks [0, m) of the du accumulators, routes 0/1 [m, nch)

    auto walk = [&](auto rev_tag) __attribute__((always_inline)) {
        constexpr bool kRev = decltype(rev_tag)::value;            // forward scan direction of the route: descending position
        // which of the lane's two granules (memory order) exist in the last chunk
        const bool okm0 = kRev ? ok_hi : ok_lo, okm1 = kRev ? ok_lo : ok_hi;

        // Both directions walk the rows towards LOWER addresses: routes 0/1 chunks nch-1 -> 0 (address = position), routes
        // 2/3 chunks 0 -> nch-1 (address = L-1 - position).  off = element offset of the granule at the lower address.
        int off = kRev ? L - 8 - 8 * lane : (nch - 1) * kChunk + 8 * lane;
        int soff = (kRev ? 0 : (nch - 1) * 32) + lane;     // checkpoint [j][lane] of the chunk being loaded
        const unsigned off_max = (unsigned)(L - ((kV8 || sizeof(T) == 2) ? 8 : 4)), soff_max = (unsigned)((nch - 1) * 32 + lane);
        // L2 prefetch ahead of the register loads: lanes 0-7 / 8-15 / 16-23 take the 128-byte lines of the dt / B / C chunk
        // rows (lanes 24-31 start so far below zero that their offset never turns non-negative)
        const T* pf_row = lane < 8 ? dt_row : (lane < 16 ? B_row : C_row);
        int pf_off = lane < 24 ? (kRev ? L - kChunk : (nch - 1) * kChunk) + (lane & 7) * 32 : -(1 << 30);      // walk step 0

        auto load = [&](LaneChunk<T>& c) __attribute__((always_inline)) {
            // Granules outside the row (last chunk; the re-load after the final chunk) are read from a valid aligned offset
            // instead -- ONE unsigned min per granule covers both ends -- and callers neutralise what they hold.
            const unsigned o0 = min((unsigned)off, off_max), o1 = min((unsigned)(off + 4), off_max);
            row_load<kV8>(c.dt, dt_row, o0, o1);
            row_load<kV8>(c.B, B_row, o0, o1);
            row_load<kV8>(c.C, C_row, o0, o1);
            c.hin = ldg32_pinned(st_row + min((unsigned)soff, soff_max));
            off -= kChunk;
            soff += kRev ? 32 : -32;
        };
        auto prefetch = [&]() __attribute__((always_inline)) {
            pf_off -= kChunk;
            if (!(XFS_LANE_DIAG & 64) && pf_off >= 0) asm volatile("prefetch.global.L2 [%0];" ::"l"(pf_row + pf_off));
        };

        LaneChunk<T> c;
        load(c);                        // in flight while the images are staged
        prefetch();                     // walk steps 1 and 2; every chunk then prefetches the step three ahead of it
        prefetch();
        {
            const T* __restrict__ x = reinterpret_cast<const T*>(p.x) + ((int64_t)b * D + d) * L;
            const float* __restrict__ dyp = reinterpret_cast<const float*>(p.dy) + ((int64_t)b * D + d) * L;
            if constexpr (std::is_same<T, float>::value) stage_two_images(x, dyp, xN, xT, gN, gT, H, W, L, Lb, tid);
            else {
                stage_image<T>(x, xN, xT, H, W, L, Lb, true, tid, 128);
                stage_image<float>(dyp, gN, gT, H, W, L, Lb, true, tid, 128);
            }
            cta_barrier();
        }

        f2 dD2 = splat2(0.0f), dbias2 = splat2(0.0f), dA2 = splat2(0.0f);
        float qcarry = 0.0f;

        auto chunk = [&](uint32_t ib, auto last_tag, const bool FIRST) __attribute__((always_inline)) {
            constexpr bool LAST = decltype(last_tag)::value;       // chunk nch-1: may hold positions >= L
            // FIRST (run time: one loop body serves both halves of the walk -- half the hot code): first touch of the pair's du accumulator
            // ---- shared-memory operands: position order in the buffers, memory order (u, dy) for the arithmetic
            f2 u[4], dy[4], dyp[4];
            {
                float4 uA = make_float4(0.f, 0.f, 0.f, 0.f), uB = uA, gA = uA, gB = uA;
                if (!LAST || in_buf_last) {
                    uA = lds128(xb + ib + offA); uB = lds128(xb + ib + offB);
                    gA = lds128(gb + ib + offA); gB = lds128(gb + ib + offB);
                }
                f2 up[4];
                unpack(uA, uB, up);
                unpack(gA, gB, dyp);
                reorder<kRev>(up, u);
                reorder<kRev>(dyp, dy);
            }
            // ---- consume the streamed registers: dt -> log2-scaled argument, C -> C dy, B -> B u and (after softplus) dt B
            f2 xr[4], Bv[4], Cv[4], xl[4], cd[4], Bu[4];
            row_unpack(c.dt, xr);
            row_unpack(c.B, Bv);
            row_unpack(c.C, Cv);
            const float hin = c.hin;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                xl[i] = kSoftplus ? fma2(xr[i], splat2(kLog2e), splat2(bias_l2)) : add2(xr[i], splat2(bias));
                cd[i] = mul2(Cv[i], dy[i]);
                Bu[i] = mul2(Bv[i], u[i]);
            }
            // The re-load of the streamed registers must be ISSUED here, most of a chunk ahead of its first use.  ptxas sinks
            // independent loads to the end of a basic block (it did: ncu showed the load latency exposed at the top of every
            // chunk), but not across a branch -- so everything that reads B / C / dt of this chunk, and the loads of the next
            // one, sit in front of the (rare) branch that repairs softplus outliers; that branch re-reads B from L2.
            f2 dt[4], sig[4], dtB[4], e2[4];
            bool odd = false;
            if (kSoftplus) {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    e2[i] = make_float2(dex2(xl[i].x), dex2(xl[i].y));
                    const f2 w = add2(e2[i], splat2(1.0f));
                    dt[i] = mul2(make_float2(dlg2(w.x), dlg2(w.y)), splat2(kLn2));
                    sig[i] = mul2(e2[i], make_float2(drcp(w.x), drcp(w.y)));     // sigmoid(x) = e / (1 + e)
                }
                // chunk-uniform test: does any element need the small-argument series or the x > 20 identity?
                odd = !(min8(e2) >= kEMin && max8(e2) <= kEMax);
            } else {
#pragma unroll
                for (int i = 0; i < 4; ++i) { dt[i] = xl[i]; sig[i] = splat2(1.0f); }
            }
            if (LAST) {     // positions >= L: identity maps (dt = 0); their u / dy are 0 in the buffers
                if (!okm0) { dt[0] = dt[1] = splat2(0.0f); }
                if (!okm1) { dt[2] = dt[3] = splat2(0.0f); }
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) dtB[i] = mul2(dt[i], Bv[i]);
            const int o = off + kChunk;            // this chunk's offset (`off` points at the next chunk until the re-load)
            load(c);
            __syncwarp();                          // a barrier ptxas does not move the (coherent) loads across
            prefetch();
            if (kSoftplus && __any_sync(kFull, odd)) {
                f2 Br[4];
                RowRaw<T> braw;
                if constexpr (std::is_same<T, float>::value) row_load<false>(braw, B_row, min((unsigned)o, (unsigned)(L - 4)), min((unsigned)(o + 4), (unsigned)(L - 4)));
                else row_load<true>(braw, B_row, min((unsigned)o, off_max), 0u);
                row_unpack(braw, Br);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const f2 x = mul2(xl[i], splat2(kLn2)), e = e2[i];       // delta + bias again
                    f2 ser = fma2(e, splat2(-0.25f), splat2(0.33333334f));
                    ser = fma2(ser, e, splat2(-0.5f));
                    ser = fma2(ser, e, splat2(1.0f));
                    ser = mul2(ser, e);
                    f2 r;
                    r.x = (e.x < kEMin) ? ser.x : dt[i].x;
                    r.y = (e.y < kEMin) ? ser.y : dt[i].y;
                    dt[i].x = (x.x > 20.0f) ? x.x : r.x;
                    dt[i].y = (x.y > 20.0f) ? x.y : r.y;
                    sig[i].x = (x.x > 20.0f) ? 1.0f : sig[i].x;
                    sig[i].y = (x.y > 20.0f) ? 1.0f : sig[i].y;
                }
                if (LAST) {
                    if (!okm0) { dt[0] = dt[1] = splat2(0.0f); }
                    if (!okm1) { dt[2] = dt[3] = splat2(0.0f); }
                }
#pragma unroll
                for (int i = 0; i < 4; ++i) dtB[i] = mul2(dt[i], Br[i]);
            }

            // ---- forward replay from the lane checkpoint (h_i = a_i h_prev + b_i) and the adjoint fold against it
            // (G_i = cd_i + a_{i+1} G_{i+1} with zero entering, Pq_i = a_{i+1} ... a_7, so that g_i = G_i + Pq_i r_in)
            f2 h[4], G[4], Pq[4], hp[4];
            float Pl, Sl;
            {
                f2 a[4], bu[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const f2 ar = mul2(dt[i], splat2(A2));
                    a[i] = make_float2(dex2(ar.x), dex2(ar.y));
                    bu[i] = mul2(dtB[i], u[i]);
                }
                float prev = hin;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    prev = fmaf(el(a, i), prev, el(bu, i));
                    el(h, i) = prev;
                }
                float Gp = el(cd, 7), Pp = 1.0f;
                el(G, 7) = Gp; el(Pq, 7) = 1.0f;
#pragma unroll
                for (int i = 6; i >= 0; --i) {
                    const float an = el(a, i + 1);
                    Gp = fmaf(an, Gp, el(cd, i));
                    Pp = (i == 6) ? an : Pp * an;
                    el(G, i) = Gp;
                    el(Pq, i) = Pp;
                }
                Pl = el(a, 0) * Pp; Sl = el(a, 0) * Gp;          // the lane's map r_in -> a_0 g_0
#pragma unroll
                for (int i = 0; i < 4; ++i) hp[i] = fma2(bu[i], splat2(-1.0f), h[i]);       // a_i h_prev
            }
            // ---- everything that does not depend on the adjoint: dC, dD, and the coefficients g will be multiplied with
            const bool st0 = !LAST || okm0, st1 = !LAST || okm1;
            f2 ts[4], dthp[4], dtu[4];
            {
                f2 dCv[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    dCv[i] = mul2(dy[i], h[i]);
                    dD2 = fma2(dy[i], u[i], dD2);
                    ts[i] = mul2(fma2(splat2(An), hp[i], Bu[i]), sig[i]);       // ddelta = g ts
                    dthp[i] = mul2(dt[i], hp[i]);                               // dA += g dthp
                    dtu[i] = mul2(dt[i], u[i]);                                 // dB = g dtu
                }
                if (!(XFS_LANE_DIAG & 1)) {
                    if (!LAST) red_sector_pair<kRev ? -8 : 8>(dC_row, o, lane, dCv);
                    else {
                        if (st0) red_add_v4(dC_row + o, dCv[0].x, dCv[0].y, dCv[1].x, dCv[1].y);
                        if (st1) red_add_v4(dC_row + o + 4, dCv[2].x, dCv[2].y, dCv[3].x, dCv[3].y);
                    }
                }
            }
            // ---- the adjoint across lanes: ONE warp scan per chunk, against the forward direction
            float q_out;
            const float r_in = (XFS_LANE_DIAG & 4) ? (q_out = fmaf(Pl, qcarry, Sl), qcarry) : warp_prefix_p<!kRev>(Pl, Sl, qcarry, lane, q_out);
            qcarry = q_out;

            f2 g[4], dd[4], dBv[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                g[i] = fma2(Pq[i], splat2(r_in), G[i]);
                dd[i] = mul2(g[i], ts[i]);
                dBv[i] = mul2(g[i], dtu[i]);
                dA2 = fma2(g[i], dthp[i], dA2);
                dbias2 = add2(dbias2, dd[i]);          // dt = 0 beyond L makes these terms exactly 0 (h_prev = 0 or g = 0 there)
            }
            if (!(XFS_LANE_DIAG & 2)) row_store<kV8>(ddt_row + o, dd, st0, st1);
            if (!(XFS_LANE_DIAG & 1)) {
                if (!LAST) red_sector_pair<kRev ? -8 : 8>(dB_row, o, lane, dBv);
                else {
                    if (st0) red_add_v4(dB_row + o, dBv[0].x, dBv[0].y, dBv[1].x, dBv[1].y);
                    if (st1) red_add_v4(dB_row + o + 4, dBv[2].x, dBv[2].y, dBv[3].x, dBv[3].y);
                }
            }
            // ---- du = D dy + g dt B into the pair's accumulator (position order: operands re-read with swapped halves)
            if (XFS_LANE_DIAG & 16) { dA2 = fma2(g[0], dtB[1], dA2); }
            else if (!LAST || in_buf_last) {
                f2 gp[4], dtBp[4], du[4];
                reorder<kRev>(g, gp);
                reorder<kRev>(dtB, dtBp);
#pragma unroll
                for (int i = 0; i < 4; ++i) du[i] = fma2(gp[i], dtBp[i], mul2(splat2(Dd), dyp[i]));
                float4 vA = make_float4(du[0].x, du[0].y, du[1].x, du[1].y), vB = make_float4(du[2].x, du[2].y, du[3].x, du[3].y);
                if (!FIRST) { vA = add4(vA, lds128(db + ib + offA)); vB = add4(vB, lds128(db + ib + offB)); }
                sts128(db + ib + offA, vA);
                sts128(db + ib + offB, vB);
            }
        };

        const std::true_type T{};
        const std::false_type F{};
        if (XFS_LANE_DIAG & 32) { pair_barrier(k & 1); }
        else if (!kRev) {                   // chunks nch-1 -> 0; first touches [m, nch)
            uint32_t ib = (uint32_t)(nch - 1) * (kChunk * 4);
            chunk(ib, T, true);
#pragma unroll 1
            for (int j = nch - 2; j >= 0; --j) {
                if (j == m - 1) pair_barrier(k & 1);
                ib -= kChunk * 4;
                chunk(ib, F, j >= m);
            }
        } else {                       // chunks 0 -> nch-1; first touches [0, m)
            uint32_t ib = 0;
#pragma unroll 1
            for (int j = 0; j < nch - 1; ++j) {
                if (j == m) pair_barrier(k & 1);
                chunk(ib, F, j < m);
                ib += kChunk * 4;
            }
            if (m >= nch - 1) pair_barrier(k & 1);
            chunk(ib, T, false);
        }

        // parameter gradients of this route
        const float vA = warp_sum(dA2.x + dA2.y), vD = warp_sum(dD2.x + dD2.y), vb = warp_sum(dbias2.x + dbias2.y);
        if (lane == 0) {
            atomicAdd(p.dA + kd, vA);
            if (p.dDs) atomicAdd(p.dDs + kd, vD);
            if (p.ddelta_bias) atomicAdd(p.ddelta_bias + kd, vb);
        }
    };
    if (k >= 2) walk(std::true_type{}); else walk(std::false_type{});
    __syncthreads();

    // dx[p] = dN[p] + dT[w*H + h]   (CrossScanF.backward = cross-merge of du, models/csm_triton.py:208-225)
    merge_out<T>(reinterpret_cast<T*>(p.dx) + ((int64_t)b * D + d) * L, dN, dT, H, W, tid, 128);
}

// ---- host side --------------------------------------------------------------------------------------------------
bool ring_enabled();

// lane-granular checkpoints are what ss2d_ring_fwd.cu (fp32) and the lane-checkpoint instantiation of ss2d_fwd.cu (bf16 / f16 rows,
// fp32 output) write and this kernel reads; both sides decide with this predicate
int ss2d_lane_states(int64_t N, int64_t H, int64_t W, int dtype, int out_dtype) {
    const int64_t L = H * W;
    if (!ring_enabled() || out_dtype != XFS_F32 || N != 1 || L <= kChunk || bwd_smem(L, 1, 1) > kSmemLimit) return 0;
    if (dtype == XFS_F32) return L % 4 == 0 && ring_fwd_smem(L, 1, 2) <= kSmemLimit;
    return L % 8 == 0;          // 16-bit rows: one 16-byte granule per lane
}

template <typename T>
static int launch_lane_bwd_t(const xfs_ss2d_bwd_args& a, cudaStream_t st) {
    const size_t smem = bwd_smem(a.H * a.W, 1, 1);
    const unsigned grid = (unsigned)(a.batch * a.D);
    auto al32 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 31) == 0; };
    const bool v8 = sizeof(T) == 4 && (a.H * a.W) % 8 == 0 && al32(a.delta) && al32(a.Bs) && al32(a.Cs) && al32(a.ddelta);
    auto go = [&](auto kern) -> int {
        if (int rc = set_smem(kern, smem)) return rc;
        kern<<<grid, 128, smem, st>>>(a);
        return 0;
    };
    int rc;
    if constexpr (sizeof(T) == 4) {
        if (a.delta_softplus) rc = v8 ? go(ss2d_lane_bwd_kernel<T, true, true>) : go(ss2d_lane_bwd_kernel<T, true, false>);
        else rc = v8 ? go(ss2d_lane_bwd_kernel<T, false, true>) : go(ss2d_lane_bwd_kernel<T, false, false>);
    } else {
        rc = a.delta_softplus ? go(ss2d_lane_bwd_kernel<T, true, false>) : go(ss2d_lane_bwd_kernel<T, false, false>);
    }
    if (rc) return rc;
    return check_launch();
}

int launch_ss2d_lane_bwd(const xfs_ss2d_bwd_args& a, cudaStream_t st) {
    if (!(aligned16(a.x) && aligned16(a.delta) && aligned16(a.Bs) && aligned16(a.Cs) && aligned16(a.dy) && aligned16(a.dx) &&
          aligned16(a.ddelta) && aligned16(a.dBs) && aligned16(a.dCs)))
        return XFS_ERR_ALIGN;
    switch (a.dtype) {
        case XFS_F32: return launch_lane_bwd_t<float>(a, st);
        case XFS_BF16: return launch_lane_bwd_t<__nv_bfloat16>(a, st);
        default: return launch_lane_bwd_t<__half>(a, st);
    }
}

}  // namespace xfs
