// ss2d_ring_fwd.cu -- fused SS2D forward, streamed operands fed by the TMA engine (design notes in ss2d_ring.cuh).
//
// Serves the shapes XFMamba's backbone spends its bytes on: N == 1, rows 16-byte aligned, L % 4 == 0, more than one
// chunk.  Everything else stays with ss2d_fwd.cu / ss2d_mid.cu / ss2d_small.cu.
#include "ss2d_ring.cuh"

namespace xfs {
using namespace ring;

// sequential fold of one half (4 positions, ascending position order in the registers) in the route's scan order.
// P/S receive the inclusive prefixes per element (composite of the half's maps up to and including that element).
template <bool kRev>
__device__ __forceinline__ void fold_half(const f2 a0, const f2 a1, const f2 b0, const f2 b1, f2& P0, f2& P1, f2& S0, f2& S1,
                                          float& Pt, float& St) {
    if (!kRev) {
        S0.x = b0.x;                   P0.x = a0.x;
        S0.y = fmaf(a0.y, S0.x, b0.y); P0.y = P0.x * a0.y;
        S1.x = fmaf(a1.x, S0.y, b1.x); P1.x = P0.y * a1.x;
        S1.y = fmaf(a1.y, S1.x, b1.y); P1.y = P1.x * a1.y;
        Pt = P1.y; St = S1.y;
    } else {
        S1.y = b1.y;                   P1.y = a1.y;
        S1.x = fmaf(a1.x, S1.y, b1.x); P1.x = P1.y * a1.x;
        S0.y = fmaf(a0.y, S1.x, b0.y); P0.y = P1.x * a0.y;
        S0.x = fmaf(a0.x, S0.y, b0.x); P0.x = P0.y * a0.x;
        Pt = P0.x; St = S0.x;
    }
}

// kDiag (timing experiments only, XFS_RING_DIAG): bit 0 = no streamed copies (arithmetic on whatever the slots hold),
// bit 1 = MUFU replaced by FMA, bit 2 = no warp scan, bit 3 = no image loads / accumulator traffic, bit 4 = scalar instead
// of packed fp32 arithmetic.  Results are wrong by construction; each bit removes one candidate bottleneck.
template <int kDiag> __device__ __forceinline__ float dex2(float x) { return (kDiag & 2) ? fmaf(x, 0.5f, 1.0f) : ex2(x); }
template <int kDiag> __device__ __forceinline__ float dlg2(float x) { return (kDiag & 2) ? fmaf(x, 0.5f, -0.5f) : lg2(x); }
template <int kDiag> __device__ __forceinline__ f2 dmul2(f2 a, f2 b) { return (kDiag & 16) ? make_float2(a.x * b.x, a.y * b.y) : mul2(a, b); }
template <int kDiag> __device__ __forceinline__ f2 dadd2(f2 a, f2 b) { return (kDiag & 16) ? make_float2(a.x + b.x, a.y + b.y) : add2(a, b); }
template <int kDiag> __device__ __forceinline__ f2 dfma2(f2 a, f2 b, f2 c) {
    return (kDiag & 16) ? make_float2(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y)) : fma2(a, b, c);
}

constexpr int kPrefetchCtas = 3 * 148;      // CTAs resident on the GPU at one time (3 per SM)

// kBig: images whose four buffers do not fit (512^2 input, stage 1: L = 16384 -> 4 x 64 KB).  The row-major image copy is
// dropped: the bulk copy of x lands in the (not yet used) yN buffer, is transposed into xT from there, and routes 0 / 2 take
// their u from a fourth row of their ring slots -- a 1 KB window of x in position order, copied chunk by chunk like delta /
// B / C.  Three 64 KB buffers + 32 KB of ring: one CTA of four warps per SM, inference only (no checkpoints written).
template <bool kSoftplus, int kDiag = 0, bool kBig = false>
__global__ void __launch_bounds__(128, kBig ? 1 : 3)
ss2d_ring_fwd_kernel(const xfs_ss2d_fwd_args p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr int kSlots = 2;
    constexpr uint32_t kSlotBytes = (kRows + (kBig ? 1 : 0)) * kChunkBytesF32;
    const int H = (int)p.H, W = (int)p.W, L = H * W;
    const int Lb = (int)buf_len(L);
    const int nch = (L + kChunk - 1) / kChunk;
    const int D = (int)p.D;
    const int b = blockIdx.x / D;
    const int d = blockIdx.x - b * D;
    const int tid = threadIdx.x, lane = tid & 31;
    const int k = __shfl_sync(kFull, tid >> 5, 0);                  // warp k runs route k (shuffle: provably warp-uniform)
    const bool transposed = k & 1;
    const bool rev = k >= 2;

    float* xT = reinterpret_cast<float*>(smem_raw) + (kBig ? 0 : Lb);
    float* yN = xT + Lb;
    float* yT = yN + Lb;
    float* xN = kBig ? yN : reinterpret_cast<float*>(smem_raw);      // kBig: staged in yN, dead once xT is built
    unsigned char* ring_mem = reinterpret_cast<unsigned char*>(yT + Lb);        // [4 warps][kSlots][3 rows][1 KB]
    uint64_t* bars = reinterpret_cast<uint64_t*>(ring_mem + 4 * kSlots * kSlotBytes);
    const uint32_t img_bar = s32(bars);
    const uint32_t my_bars = s32(bars + 1 + k * kSlots);
    const uint32_t my_ring = s32(ring_mem) + (uint32_t)k * (kSlots * kSlotBytes);

    if (tid == 0) {
        // the image first: its latency is the longest wait of the CTA
        mbar_init(img_bar, 1);
        mbar_fence_init();
        if (!(kDiag & 64)) {
            mbar_expect_tx(img_bar, (uint32_t)L * 4u);
            bulk_g2s(s32(xN), reinterpret_cast<const float*>(p.x) + ((int64_t)b * D + d) * L, (uint32_t)L * 4u, img_bar);
        }
        for (int i = 0; i < 4 * kSlots; ++i) mbar_init(s32(bars + 1 + i), 1);
        mbar_fence_init();
        // ... and pull what a CTA about one CTA-lifetime behind this one will ask for first into L2: its image and the
        // heads of its four delta rows (B / C rows are shared by all channels of an image and stay L2 resident anyway)
        const int64_t nb = (int64_t)blockIdx.x + (kBig ? 148 : kPrefetchCtas);
        if (nb < (int64_t)gridDim.x) {
            const int64_t b2 = nb / D, d2 = nb - b2 * D;
            bulk_prefetch_l2(reinterpret_cast<const float*>(p.x) + (b2 * D + d2) * L, (uint32_t)L * 4u);
            const uint32_t head = (uint32_t)min(L, 2 * kChunk) * 4u;
#pragma unroll
            for (int r = 0; r < 4; ++r)
                bulk_prefetch_l2(reinterpret_cast<const float*>(p.delta) + ((b2 * 4 + r) * D + d2) * L, head);
        }
    }
    __syncthreads();

    const int kd = k * D + d;
    // Rows are read in the route's scan order, which always ascends in memory: chunk c (c-th of the walk) covers scan
    // indices [off0 + 256 c, off0 + 256 c + 256) clipped to [0, L).  A flipped route meets the short chunk first.
    const int off0 = rev ? L - kChunk * nch : 0;
    const float* f_dt = reinterpret_cast<const float*>(p.delta) + ((int64_t)b * 4 * D + kd) * L + off0;   // next chunk to copy
    const float* f_B = reinterpret_cast<const float*>(p.Bs) + ((int64_t)b * 4 + k) * L + off0;
    const float* f_C = reinterpret_cast<const float*>(p.Cs) + ((int64_t)b * 4 + k) * L + off0;
    int f_left = nch;                  // chunks not yet requested
    int f_rem = L * 4 - (rev ? 0 : 0); // bytes of the row not yet requested (forward routes end with the short chunk)
    // kBig, routes 0 / 2: the window of x in POSITION order that belongs to the chunk being requested (a flipped route walks it downwards)
    const bool xwin = kBig && !transposed;
    const float* f_x = reinterpret_cast<const float*>(p.x) + ((int64_t)b * D + d) * L + (rev ? (nch - 1) * kChunk : 0);

    // One lane per warp feeds the warp's ring.  Every operand is warp-uniform (uniform registers); the running pointers
    // advance by one chunk per call, so a call is ~15 instructions: arm the mbarrier, three bulk copies, three adds.
    auto fill = [&](uint32_t slot) __attribute__((always_inline)) {
        if (kDiag & 1) { --f_left; return; }
        const uint32_t bytes = (uint32_t)min(f_rem, kChunkBytesF32);
        const uint32_t bar = my_bars + 8u * slot, dst = my_ring + slot * kSlotBytes;
        if (elect_one()) {
            mbar_expect_tx(bar, (xwin ? 4u : 3u) * bytes);
            bulk_g2s(dst, f_dt, bytes, bar);
            bulk_g2s(dst + kChunkBytesF32, f_B, bytes, bar);
            bulk_g2s(dst + 2 * kChunkBytesF32, f_C, bytes, bar);
            if (xwin) bulk_g2s(dst + 3 * kChunkBytesF32, f_x, bytes, bar);
        }
        f_dt += kChunk; f_B += kChunk; f_C += kChunk;
        f_x += rev ? -kChunk : kChunk;
        f_rem -= kChunkBytesF32; --f_left;
    };

    if (rev && off0 < 0) {             // the short chunk comes first: its data goes to the END of slot 0
        const uint32_t skip = (uint32_t)(-off0) * 4u, bytes = kChunkBytesF32 - skip;
        if (!(kDiag & 1) && elect_one()) {
            mbar_expect_tx(my_bars, (xwin ? 4u : 3u) * bytes);
            bulk_g2s(my_ring + skip, f_dt - off0, bytes, my_bars);
            bulk_g2s(my_ring + skip + kChunkBytesF32, f_B - off0, bytes, my_bars);
            bulk_g2s(my_ring + skip + 2 * kChunkBytesF32, f_C - off0, bytes, my_bars);
            if (xwin) bulk_g2s(my_ring + 3 * kChunkBytesF32, f_x, bytes, my_bars);      // position order: the short chunk sits at the window's start
        }
        f_dt += kChunk; f_B += kChunk; f_C += kChunk;
        f_x -= kChunk;
        f_rem -= (int)bytes; --f_left;
    } else {
        fill(0);
    }
    if (f_left > 0) fill(1);
    for (int q = L + tid; q < Lb; q += 128) xN[q] = 0.0f;       // tail of the linear copy (the bulk copy writes [0, L))
    mbar_wait(img_bar, 0);
    transpose_image(xN, xT, H, W, L, Lb, tid, 128);
    __syncthreads();

    const float bias = p.delta_bias ? p.delta_bias[kd] : 0.0f;
    const float bias_l2 = bias * kLog2e;
    const float Dd = p.Ds ? p.Ds[kd] : 0.0f;
    const float A2 = p.A[kd] * kLog2e;
    const LaneOffsets o = lane_offsets(lane, rev, transposed);
    const uint32_t xb = s32(transposed ? xT : xN), yb = s32(transposed ? yT : yN);
    // chunk index j in POSITION order: forward routes walk 0 .. nch-1, flipped routes nch-1 .. 0.  Chunks [0, m) are first
    // touched by the forward route, [m, nch) by its flip: each route first-touches during its first n_first steps.
    const int m = (nch + 1) / 2;
    const int n_first = rev ? nch - m : m;
    // LANE-GRANULAR checkpoints for ss2d_lane_bwd.cu: row (b, k*D+d) holds nch x 32 floats, entry [j][lane] = the state
    // ENTERING the 8 positions of lane `lane` of position-order chunk j, in the route's scan direction
    float* st_ptr = p.states ? p.states + ((int64_t)b * 4 * D + kd) * ((int64_t)nch * 32) + (rev ? (nch - 1) * 32 : 0) + lane : nullptr;
    const int st_inc = rev ? -32 : 32;
    float carry = 0.0f;

    auto chunk = [&](int step, uint32_t ib, auto rev_tag, auto last_tag, const bool FIRST) __attribute__((always_inline)) {
        constexpr bool R = decltype(rev_tag)::value;
        constexpr bool LAST = decltype(last_tag)::value;      // the chunk that may hold positions >= L
        // FIRST (run time: one loop body for both halves of the walk): first touch of the pair's accumulator, plain stores
        const uint32_t slot = (uint32_t)step & 1u;
        const uint32_t sb = my_ring + slot * kSlotBytes;
        bool inA = true, inB = true, okA = true, okB = true;
        if (LAST) {
            const int pA = (nch - 1) * kChunk + o.posA, pB = (nch - 1) * kChunk + o.posB;
            inA = pA < Lb; inB = pB < Lb; okA = pA < L; okB = pB < L;
        }
        if (!(kDiag & 1)) mbar_wait(my_bars + 8u * slot, ((uint32_t)step >> 1) & 1u);

        f2 dt[4], Bv[4], Cv[4], u[4];
        half_from<R>(lds128(sb + o.rowA), dt[0], dt[1]);
        half_from<R>(lds128(sb + o.rowB), dt[2], dt[3]);
        half_from<R>(lds128(sb + kChunkBytesF32 + o.rowA), Bv[0], Bv[1]);
        half_from<R>(lds128(sb + kChunkBytesF32 + o.rowB), Bv[2], Bv[3]);
        half_from<R>(lds128(sb + 2 * kChunkBytesF32 + o.rowA), Cv[0], Cv[1]);
        half_from<R>(lds128(sb + 2 * kChunkBytesF32 + o.rowB), Cv[2], Cv[3]);
        {
            float4 gA = make_float4(0.f, 0.f, 0.f, 0.f), gB = gA;
            if (kBig && !transposed) {          // u from the slot's x window (position order: the image offsets of a linear buffer)
                gA = lds128(sb + 3 * kChunkBytesF32 + o.imgA);
                gB = lds128(sb + 3 * kChunkBytesF32 + o.imgB);
                if (LAST) {                      // beyond the copied bytes the window holds stale data
                    if (!okA) gA = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (!okB) gB = make_float4(0.f, 0.f, 0.f, 0.f);
                }
            } else if (!(kDiag & 8)) {
                if (!LAST || inA) gA = lds128(xb + ib + o.imgA);
                if (!LAST || inB) gB = lds128(xb + ib + o.imgB);
            } else { gA = make_float4(dt[0].x, dt[0].y, dt[1].x, dt[1].y); gB = make_float4(dt[2].x, dt[2].y, dt[3].x, dt[3].y); }
            half_from<false>(gA, u[0], u[1]);
            half_from<false>(gB, u[2], u[3]);
        }

        // ---- dt = softplus(delta + bias)
        f2 dtp[4];
        if (kSoftplus) {
            f2 e2[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const f2 xl = dfma2<kDiag>(dt[i], splat2(kLog2e), splat2(bias_l2));
                e2[i] = make_float2(dex2<kDiag>(xl.x), dex2<kDiag>(xl.y));
                const f2 ww = dadd2<kDiag>(e2[i], splat2(1.0f));
                dtp[i] = dmul2<kDiag>(make_float2(dlg2<kDiag>(ww.x), dlg2<kDiag>(ww.y)), splat2(kLn2));
            }
            // chunk-uniform test: does any element need the small-argument series or the x > 20 identity?
            const bool odd = !(min8(e2) >= kEMin && max8(e2) <= kEMax);
            if (__any_sync(kFull, odd)) {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const f2 x = add2(dt[i], splat2(bias)), e = e2[i];
                    f2 ser = fma2(e, splat2(-0.25f), splat2(0.33333334f));
                    ser = fma2(ser, e, splat2(-0.5f));
                    ser = fma2(ser, e, splat2(1.0f));
                    ser = mul2(ser, e);
                    f2 r;
                    r.x = (e.x < kEMin) ? ser.x : dtp[i].x;
                    r.y = (e.y < kEMin) ? ser.y : dtp[i].y;
                    dtp[i].x = (x.x > 20.0f) ? x.x : r.x;
                    dtp[i].y = (x.y > 20.0f) ? x.y : r.y;
                }
            }
        } else {
#pragma unroll
            for (int i = 0; i < 4; ++i) dtp[i] = add2(dt[i], splat2(bias));
        }
        if (LAST) {     // positions >= L: identity maps (dt = 0, B = 0); whatever the slot holds there is stale
            if (!okA) { dtp[0] = dtp[1] = splat2(0.0f); Bv[0] = Bv[1] = splat2(0.0f); }
            if (!okB) { dtp[2] = dtp[3] = splat2(0.0f); Bv[2] = Bv[3] = splat2(0.0f); }
        }

        f2 a[4], bu[4], y[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const f2 ar = dmul2<kDiag>(dtp[i], splat2(A2));
            a[i] = make_float2(dex2<kDiag>(ar.x), dex2<kDiag>(ar.y));
            bu[i] = dmul2<kDiag>(dtp[i], dmul2<kDiag>(Bv[i], u[i]));
            y[i] = dmul2<kDiag>(splat2(Dd), u[i]);
        }
        f2 P[4], S[4];
        float PA, SA, PB, SB;
        fold_half<R>(a[0], a[1], bu[0], bu[1], P[0], P[1], S[0], S[1], PA, SA);
        fold_half<R>(a[2], a[3], bu[2], bu[3], P[2], P[3], S[2], S[3], PB, SB);

        // every register loaded from the slot has been consumed: refill it with the chunk two steps ahead
        __syncwarp();
        if (f_left > 0) fill(slot);

        float h_out, h_in;
        if (kDiag & 4) { h_in = carry; h_out = fmaf(PA * PB, carry, o.a_first ? fmaf(PB, SA, SB) : fmaf(PA, SB, SA)); }
        else h_in = warp_prefix_p<R>(PA * PB, o.a_first ? fmaf(PB, SA, SB) : fmaf(PA, SB, SA), carry, lane, h_out);
        carry = h_out;
        if (st_ptr) {
            *st_ptr = h_in;
            st_ptr += st_inc;
        }
        const f2 hA = splat2(o.a_first ? h_in : fmaf(PB, h_in, SB));
        const f2 hB = splat2(o.a_first ? fmaf(PA, h_in, SA) : h_in);
#pragma unroll
        for (int i = 0; i < 4; ++i) y[i] = dfma2<kDiag>(Cv[i], dfma2<kDiag>(P[i], i < 2 ? hA : hB, S[i]), y[i]);

        // ---- accumulate into the pair's buffer
        if (kDiag & 8) { carry += (y[0].x + y[0].y + y[1].x + y[1].y) + (y[2].x + y[2].y + y[3].x + y[3].y); return; }
        if (!LAST || inA) {
            float4 v = half_to<false>(y[0], y[1]);
            if (!FIRST) v = add4(v, lds128(yb + ib + o.imgA));
            sts128(yb + ib + o.imgA, v);
        }
        if (!LAST || inB) {
            float4 v = half_to<false>(y[2], y[3]);
            if (!FIRST) v = add4(v, lds128(yb + ib + o.imgB));
            sts128(yb + ib + o.imgB, v);
        }
    };

    // Walk: first-touch steps, ONE pair barrier, then read-modify-write steps.  The short chunk (j = nch - 1) is the first
    // step of a flipped route (always a first touch) and the last step of a forward route (a first touch only if nch == 1).
    const std::true_type T{};
    const std::false_type F{};
    if (kDiag & 32) {                  // no walk at all: what the prologue and the epilogue cost on their own
        pair_barrier(k & 1);
    } else if (rev) {
        // steps 0 .. nch-1 = chunks nch-1 .. 0; step 0 is the short chunk (always a first touch); first touches: steps [0, n_first)
        uint32_t ib = (uint32_t)(nch - 1) * kChunkBytesF32;
        chunk(0, ib, T, T, true);
#pragma unroll 1
        for (int step = 1; step < nch; ++step) {
            if (step == n_first) pair_barrier(k & 1);
            ib -= kChunkBytesF32;
            chunk(step, ib, T, F, step < n_first);
        }
        if (n_first >= nch) pair_barrier(k & 1);
    } else {
        // steps = chunks 0 .. nch-1; first touches [0, n_first) (n_first <= nch - 1: nch >= 2 here); the short chunk is the last step
        uint32_t ib = 0;
        int step = 0;
#pragma unroll 1
        for (; step < nch - 1; ++step) {
            if (step == n_first) pair_barrier(k & 1);
            chunk(step, ib, F, F, step < n_first);
            ib += kChunkBytesF32;
        }
        if (n_first >= nch - 1) pair_barrier(k & 1);
        chunk(step, ib, F, T, n_first >= nch);
    }
    __syncthreads();

    // merged output, spatial order: y[p] = yN[p] + yT[w*H + h]
    merge_out_linear<float>(reinterpret_cast<float*>(p.y) + ((int64_t)b * D + d) * L, yN, yT, H, W, L, tid, 128);
}

// ---- host side --------------------------------------------------------------------------------------------------
bool ring_enabled() {
    static const bool on = [] { const char* e = std::getenv("XFS_NO_RING"); return !(e && e[0] == '1'); }();
    return on;
}

int ss2d_ring_fwd_supported(const xfs_ss2d_fwd_args& a) {
    const int64_t L = a.H * a.W;
    return ring_enabled() && a.dtype == XFS_F32 && a.out_dtype == XFS_F32 && a.N == 1 && a.scans == 0 && L % 4 == 0 && L > kChunk &&
           L <= (1 << 22) && aligned16(a.x) && aligned16(a.delta) && aligned16(a.Bs) && aligned16(a.Cs) && aligned16(a.y) &&
           ring_fwd_smem(L, 1, 2) <= kSmemLimit;
}

inline size_t ring_fwd_big_smem(int64_t L) {        // xT, yN, yT + four rows per slot
    return sizeof(float) * (size_t)(3 * buf_len(L)) + (size_t)(4 * 2 * 4 * kChunkBytesF32) + 8 * (size_t)(1 + 4 * 2 + 8) + 8 * sizeof(float);
}
// shape-only predicate (xfs_ss2d_supported): forward without checkpoints of an image whose four buffers exceed shared memory
int ss2d_ring_big_shape(int64_t N, int64_t H, int64_t W, int dtype) {
    const int64_t L = H * W;
    return ring_enabled() && dtype == XFS_F32 && N == 1 && L % 4 == 0 && L > kChunk && L <= (1 << 22) &&
           ring_fwd_smem(L, 1, 2) > kSmemLimit && ring_fwd_big_smem(L) <= kSmemLimit;
}
int launch_ss2d_ring_big_fwd(const xfs_ss2d_fwd_args& a, cudaStream_t st) {
    if (!(aligned16(a.x) && aligned16(a.delta) && aligned16(a.Bs) && aligned16(a.Cs) && aligned16(a.y))) return XFS_ERR_ALIGN;
    if (a.states != nullptr || a.out_dtype != XFS_F32) return XFS_ERR_UNSUPPORTED;
    const size_t smem = ring_fwd_big_smem(a.H * a.W);
    const unsigned grid = (unsigned)(a.batch * a.D);
    if (a.delta_softplus) {
        if (int rc = set_smem(ss2d_ring_fwd_kernel<true, 0, true>, smem)) return rc;
        ss2d_ring_fwd_kernel<true, 0, true><<<grid, 128, smem, st>>>(a);
    } else {
        if (int rc = set_smem(ss2d_ring_fwd_kernel<false, 0, true>, smem)) return rc;
        ss2d_ring_fwd_kernel<false, 0, true><<<grid, 128, smem, st>>>(a);
    }
    return check_launch();
}

template <bool kSoftplus, int kDiag = 0>
static int launch_ring_fwd_k(const xfs_ss2d_fwd_args& a, cudaStream_t st) {
    const size_t smem = ring_fwd_smem(a.H * a.W, 1, 2);
    if (int rc = set_smem(ss2d_ring_fwd_kernel<kSoftplus, kDiag>, smem)) return rc;
    ss2d_ring_fwd_kernel<kSoftplus, kDiag><<<(unsigned)(a.batch * a.D), 128, smem, st>>>(a);
    return check_launch();
}

int launch_ss2d_ring_fwd(const xfs_ss2d_fwd_args& a, cudaStream_t st) {
#ifdef XFS_RING_DIAG_BUILD
    static const int diag = [] { const char* e = std::getenv("XFS_RING_DIAG"); return e ? std::atoi(e) : 0; }();
    switch (diag) {
        case 1: return launch_ring_fwd_k<true, 1>(a, st);
        case 2: return launch_ring_fwd_k<true, 2>(a, st);
        case 4: return launch_ring_fwd_k<true, 4>(a, st);
        case 8: return launch_ring_fwd_k<true, 8>(a, st);
        case 16: return launch_ring_fwd_k<true, 16>(a, st);
        case 6: return launch_ring_fwd_k<true, 6>(a, st);
        case 14: return launch_ring_fwd_k<true, 14>(a, st);
        case 3: return launch_ring_fwd_k<true, 3>(a, st);
        case 5: return launch_ring_fwd_k<true, 5>(a, st);
        case 9: return launch_ring_fwd_k<true, 9>(a, st);
        case 17: return launch_ring_fwd_k<true, 17>(a, st);
        case 15: return launch_ring_fwd_k<true, 15>(a, st);
        case 7: return launch_ring_fwd_k<true, 7>(a, st);
        case 13: return launch_ring_fwd_k<true, 13>(a, st);
        case 32: return launch_ring_fwd_k<true, 32>(a, st);
        case 33: return launch_ring_fwd_k<true, 33>(a, st);
        default: break;
    }
#endif
    return a.delta_softplus ? launch_ring_fwd_k<true>(a, st) : launch_ring_fwd_k<false>(a, st);
}

}  // namespace xfs
