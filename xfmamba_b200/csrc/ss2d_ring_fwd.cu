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

template <bool kSoftplus, int kSlots>
__global__ void __launch_bounds__(128, 3)
ss2d_ring_fwd_kernel(const xfs_ss2d_fwd_args p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int H = (int)p.H, W = (int)p.W, L = H * W;
    const int Lb = (int)buf_len(L);
    const int nch = (L + kChunk - 1) / kChunk;
    const int D = (int)p.D;
    const int b = blockIdx.x / D;
    const int d = blockIdx.x - b * D;
    const int tid = threadIdx.x, lane = tid & 31;
    const int k = __shfl_sync(kFull, tid >> 5, 0);                  // warp k runs route k (shuffle: provably warp-uniform)
    const bool transposed = k & 1;

    float* xN = reinterpret_cast<float*>(smem_raw);
    float* xT = xN + Lb;
    float* yN = xT + Lb;
    float* yT = yN + Lb;
    unsigned char* ring_mem = reinterpret_cast<unsigned char*>(yT + Lb);        // [4 warps][kSlots][3 rows][1 KB]
    uint64_t* bars = reinterpret_cast<uint64_t*>(ring_mem + 4 * kSlots * kRows * kChunkBytesF32);
    const uint32_t img_bar = s32(bars);
    const uint32_t my_bars = s32(bars + 1 + k * kSlots);
    const uint32_t my_ring = s32(ring_mem) + (uint32_t)(k * kSlots * kRows * kChunkBytesF32);

    if (tid == 0) {
        mbar_init(img_bar, 1);
        for (int i = 0; i < 4 * kSlots; ++i) mbar_init(s32(bars + 1 + i), 1);
        mbar_fence_init();
    }
    __syncthreads();

    const int kd = k * D + d;
    const float* __restrict__ dt_row = reinterpret_cast<const float*>(p.delta) + ((int64_t)b * 4 * D + kd) * L;
    const float* __restrict__ B_row = reinterpret_cast<const float*>(p.Bs) + ((int64_t)b * 4 + k) * L;
    const float* __restrict__ C_row = reinterpret_cast<const float*>(p.Cs) + ((int64_t)b * 4 + k) * L;
    const bool rev = k >= 2;

    // one lane per warp feeds the warp's ring: the three rows of the chunk walked at `step`.  Every operand is warp-uniform
    // (block index, warp index, step), so the copies are issued from uniform registers by the elected lane.
    auto fill = [&](int step, auto full_tag) __attribute__((always_inline)) {
        constexpr bool kFull256 = decltype(full_tag)::value;               // a whole 256-position chunk (all but one)
        const int j = rev ? nch - 1 - step : step;
        const int start = rev ? L - kChunk * (j + 1) : kChunk * j;        // scan index of slot element 0
        const int lo = kFull256 ? start : max(start, 0), hi = kFull256 ? start + kChunk : min(start + kChunk, L);
        const uint32_t bytes = (uint32_t)(hi - lo) * 4u, skip = (uint32_t)(lo - start) * 4u;
        const int slot = step % kSlots;
        const uint32_t bar = my_bars + 8u * slot;
        const uint32_t dst = my_ring + (uint32_t)(slot * kRows * kChunkBytesF32) + skip;
        if (elect_one()) {
            mbar_expect_tx(bar, 3u * bytes);
            bulk_g2s(dst, dt_row + lo, bytes, bar);
            bulk_g2s(dst + kChunkBytesF32, B_row + lo, bytes, bar);
            bulk_g2s(dst + 2 * kChunkBytesF32, C_row + lo, bytes, bar);
        }
    };
    const int j_part = nch - 1;                                             // the chunk that may be shorter than 256
    auto fill_any = [&](int step) __attribute__((always_inline)) {
        const int j = rev ? nch - 1 - step : step;
        if (j == j_part) fill(step, std::false_type{}); else fill(step, std::true_type{});
    };

    if (tid == 0) {
        mbar_expect_tx(img_bar, (uint32_t)L * 4u);
        bulk_g2s(s32(xN), reinterpret_cast<const float*>(p.x) + ((int64_t)b * D + d) * L, (uint32_t)L * 4u, img_bar);
    }
#pragma unroll
    for (int s = 0; s < kSlots; ++s)
        if (s < nch) fill_any(s);
    for (int q = L + tid; q < Lb; q += 128) xN[q] = 0.0f;       // tail of the linear copy (the bulk copy writes [0, L))
    mbar_wait(img_bar, 0);
    transpose_image(xN, xT, H, W, L, Lb, tid, 128);
    __syncthreads();

    const float bias = p.delta_bias ? p.delta_bias[kd] : 0.0f;
    const float bias_l2 = bias * kLog2e;
    const float Dd = p.Ds ? p.Ds[kd] : 0.0f;
    const float A2 = p.A[kd] * kLog2e;
    float* st_row = p.states ? p.states + ((int64_t)b * 4 * D + kd) * nch : nullptr;
    const LaneOffsets o = lane_offsets(lane, rev, transposed);
    const uint32_t xb = s32(transposed ? xT : xN), yb = s32(transposed ? yT : yN);
    const int m = (nch + 1) / 2;        // chunks [0, m) are first touched by the forward route, [m, nch) by its flip
    bool synced = false;
    float carry = 0.0f;

    auto chunk = [&](int step, auto rev_tag, auto last_tag) __attribute__((always_inline)) {
        constexpr bool R = decltype(rev_tag)::value;
        constexpr bool LAST = decltype(last_tag)::value;      // the chunk that may hold positions >= L
        const int j = R ? nch - 1 - step : step;
        const int slot = step % kSlots;
        const uint32_t sb = my_ring + (uint32_t)(slot * kRows * kChunkBytesF32);
        const uint32_t ib = (uint32_t)j * (uint32_t)kChunkBytesF32;
        bool inA = true, inB = true, okA = true, okB = true;
        if (LAST) {
            const int pA = j * kChunk + o.posA, pB = j * kChunk + o.posB;
            inA = pA < Lb; inB = pB < Lb; okA = pA < L; okB = pB < L;
        }
        mbar_wait(my_bars + 8u * slot, (uint32_t)(step / kSlots) & 1u);

        f2 dt[4], Bv[4], Cv[4], u[4];
        half_from<R>(lds128(sb + o.rowA), dt[0], dt[1]);
        half_from<R>(lds128(sb + o.rowB), dt[2], dt[3]);
        half_from<R>(lds128(sb + kChunkBytesF32 + o.rowA), Bv[0], Bv[1]);
        half_from<R>(lds128(sb + kChunkBytesF32 + o.rowB), Bv[2], Bv[3]);
        half_from<R>(lds128(sb + 2 * kChunkBytesF32 + o.rowA), Cv[0], Cv[1]);
        half_from<R>(lds128(sb + 2 * kChunkBytesF32 + o.rowB), Cv[2], Cv[3]);
        {
            float4 gA = make_float4(0.f, 0.f, 0.f, 0.f), gB = gA;
            if (!LAST || inA) gA = lds128(xb + ib + o.imgA);
            if (!LAST || inB) gB = lds128(xb + ib + o.imgB);
            half_from<false>(gA, u[0], u[1]);
            half_from<false>(gB, u[2], u[3]);
        }

        // ---- dt = softplus(delta + bias)
        f2 dtp[4];
        if (kSoftplus) {
            f2 e2[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                e2[i] = ex2_2(fma2(dt[i], splat2(kLog2e), splat2(bias_l2)));
                const f2 w = add2(e2[i], splat2(1.0f));
                dtp[i] = mul2(make_float2(lg2(w.x), lg2(w.y)), splat2(kLn2));
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
            a[i] = ex2_2(mul2(dtp[i], splat2(A2)));
            bu[i] = mul2(dtp[i], mul2(Bv[i], u[i]));
            y[i] = mul2(splat2(Dd), u[i]);
        }
        f2 P[4], S[4];
        float PA, SA, PB, SB;
        fold_half<R>(a[0], a[1], bu[0], bu[1], P[0], P[1], S[0], S[1], PA, SA);
        fold_half<R>(a[2], a[3], bu[2], bu[3], P[2], P[3], S[2], S[3], PB, SB);

        // every register loaded from the slot has been consumed: refill it with the chunk kSlots steps ahead
        __syncwarp();
        if (step + kSlots < nch) fill_any(step + kSlots);

        const float S_ab = fmaf(PB, SA, SB), S_ba = fmaf(PA, SB, SA);
        float h_out;
        const float h_in = warp_prefix<R>(PA * PB, o.a_first ? S_ab : S_ba, carry, lane, h_out);
        carry = h_out;
        if (st_row && lane == 0) st_row[j] = h_out;
        const f2 hA = splat2(o.a_first ? h_in : fmaf(PB, h_in, SB));
        const f2 hB = splat2(o.a_first ? fmaf(PA, h_in, SA) : h_in);
#pragma unroll
        for (int i = 0; i < 4; ++i) y[i] = fma2(Cv[i], fma2(P[i], i < 2 ? hA : hB, S[i]), y[i]);

        // ---- accumulate into the pair's buffer
        const bool first_touch = R ? (j >= m) : (j < m);
        if (!first_touch && !synced) { pair_barrier(k & 1); synced = true; }
        if (!LAST || inA) {
            float4 v = half_to<false>(y[0], y[1]);
            if (!first_touch) v = add4(v, lds128(yb + ib + o.imgA));
            sts128(yb + ib + o.imgA, v);
        }
        if (!LAST || inB) {
            float4 v = half_to<false>(y[2], y[3]);
            if (!first_touch) v = add4(v, lds128(yb + ib + o.imgB));
            sts128(yb + ib + o.imgB, v);
        }
    };

    if (rev) {
        chunk(0, std::true_type{}, std::true_type{});
#pragma unroll 1
        for (int step = 1; step < nch; ++step) chunk(step, std::true_type{}, std::false_type{});
    } else {
#pragma unroll 1
        for (int step = 0; step < nch - 1; ++step) chunk(step, std::false_type{}, std::false_type{});
        chunk(nch - 1, std::false_type{}, std::true_type{});
    }
    if (!synced) pair_barrier(k & 1);
    __syncthreads();

    // merged output, spatial order: y[p] = yN[p] + yT[w*H + h]
    merge_out_linear<float>(reinterpret_cast<float*>(p.y) + ((int64_t)b * D + d) * L, yN, yT, H, W, L, tid, 128);
}

// ---- host side --------------------------------------------------------------------------------------------------
constexpr int kRingFwdSlots = 2;

bool ring_enabled() {
    static const bool on = [] { const char* e = std::getenv("XFS_NO_RING"); return !(e && e[0] == '1'); }();
    return on;
}

int ss2d_ring_fwd_supported(const xfs_ss2d_fwd_args& a) {
    const int64_t L = a.H * a.W;
    return ring_enabled() && a.dtype == XFS_F32 && a.out_dtype == XFS_F32 && a.N == 1 && a.scans == 0 && L % 4 == 0 && L > kChunk &&
           L <= (1 << 22) && aligned16(a.x) && aligned16(a.delta) && aligned16(a.Bs) && aligned16(a.Cs) && aligned16(a.y) &&
           ring_fwd_smem(L, kRingFwdSlots) <= kSmemLimit;
}

template <bool kSoftplus>
static int launch_ring_fwd_k(const xfs_ss2d_fwd_args& a, cudaStream_t st) {
    const size_t smem = ring_fwd_smem(a.H * a.W, kRingFwdSlots);
    if (int rc = set_smem(ss2d_ring_fwd_kernel<kSoftplus, kRingFwdSlots>, smem)) return rc;
    ss2d_ring_fwd_kernel<kSoftplus, kRingFwdSlots><<<(unsigned)(a.batch * a.D), 128, smem, st>>>(a);
    return check_launch();
}

int launch_ss2d_ring_fwd(const xfs_ss2d_fwd_args& a, cudaStream_t st) {
    return a.delta_softplus ? launch_ring_fwd_k<true>(a, st) : launch_ring_fwd_k<false>(a, st);
}

}  // namespace xfs
