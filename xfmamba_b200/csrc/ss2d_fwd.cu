// ss2d_fwd.cu -- fused SS2D forward kernel (design notes in ss2d_fused.cuh)
#include "ss2d_fused.cuh"

namespace xfs {

template <int kN, int kCh>
struct FwdChunk {            // delta/B/C of one chunk of one route, in ADDRESS order, exactly as loaded
    float dt[kCh][8];
    float B[8], C[8];       // kN == 1 only
};

// kN == 1: single state carried in a register; kN == 0: runtime N (<= kFusedMaxState), states carried in smem
// kSingle: the sequence fits ONE chunk (L <= 256) -> no chunk loop, no carried state; the leaner code needs fewer
// registers, so more CTAs fit per SM (these short shapes -- XFMamba's 14x14 stage has 15 of the 21 blocks -- are latency bound)
// kLane: LANE-granular checkpoints for ss2d_lane_bwd.cu (16-bit rows; fp32 rows get them from ss2d_ring_fwd.cu): row (b, k*D+d)
// holds nch x 32 floats, entry [j][lane] = the state entering the 8 positions of lane `lane` of position-order chunk j
template <typename T, typename TO, int kN, int kCh, bool kFast, bool kSingle, bool kLane = false>
__global__ void __launch_bounds__(128, kSingle ? 6 : 4)
ss2d_fwd_kernel(const xfs_ss2d_fwd_args p) {
    extern __shared__ __align__(16) float smem[];
    const int H = (int)p.H, W = (int)p.W, L = H * W;
    const int Lb = (int)buf_len(L);
    const int nch = kSingle ? 1 : (L + kChunk - 1) / kChunk;
    const int D = (int)p.D;
    const int N = (kN == 1) ? 1 : (int)p.N;
    const int groups = (D + kCh - 1) / kCh;
    const int b = blockIdx.x / groups;
    const int d0 = (blockIdx.x - b * groups) * kCh;
    const int tid = threadIdx.x, lane = tid & 31, k = tid >> 5;     // warp k runs route k
    const bool transposed = k & 1;
    bool valid[kCh];
#pragma unroll
    for (int ch = 0; ch < kCh; ++ch) valid[ch] = (d0 + ch) < D;

    float* xN = smem;                   // [kCh][Lb]
    float* xT = xN + kCh * Lb;
    float* yN = xT + kCh * Lb;
    float* yT = yN + kCh * Lb;
    float* s_h = yT + kCh * Lb;         // [4][kCh][kFusedMaxState] (kN == 0 only)

    const float* xb = transposed ? xT : xN;
    float* yb = transposed ? yT : yN;
    const T* __restrict__ delta = reinterpret_cast<const T*>(p.delta);
    const T* __restrict__ Bk = reinterpret_cast<const T*>(p.Bs) + ((int64_t)b * 4 + k) * N * L;
    const T* __restrict__ Ck = reinterpret_cast<const T*>(p.Cs) + ((int64_t)b * 4 + k) * N * L;
    const bool vin = kFast || (row_vec_ok(delta, L) && row_vec_ok(reinterpret_cast<const T*>(p.Bs), L) &&
                               row_vec_ok(reinterpret_cast<const T*>(p.Cs), L));

    const T* dt_row[kCh];
    float* st_row[kCh];
    float bias[kCh], Dd[kCh], A2_1[kCh], carry1[kCh];
    int kd[kCh];
#pragma unroll
    for (int ch = 0; ch < kCh; ++ch) {
        kd[ch] = k * D + (valid[ch] ? d0 + ch : d0);
        dt_row[ch] = delta + ((int64_t)b * 4 * D + kd[ch]) * L;
        st_row[ch] = (p.states && valid[ch]) ? p.states + ((int64_t)b * 4 * D + kd[ch]) * nch * (kLane ? 32 : N) : nullptr;
        bias[ch] = p.delta_bias ? p.delta_bias[kd[ch]] : 0.0f;
        Dd[ch] = p.Ds ? p.Ds[kd[ch]] : 0.0f;
        A2_1[ch] = (kN == 1) ? p.A[kd[ch]] * kLog2e : 0.0f;
        carry1[ch] = 0.0f;
    }
    const f2 rt_zero2 = splat2(__int_as_float(p.scans));   // +0.0f (scans == 0), but only known at run time

    const int m = (nch + 1) / 2;        // chunks [0, m) are first touched by the forward route, [m, nch) by its flip
    bool synced = false;

    // The walk is instantiated twice (rev = false for routes 0/1, true for their flips 2/3) so that the flip is pure
    // register renaming (or an operand swizzle of the packed instructions) instead of predicated moves.
    auto walk = [&](auto rev_tag) __attribute__((always_inline)) {
        constexpr bool rev = decltype(rev_tag)::value;

        // Only the LAST chunk (spatial positions >= L) has 16-byte granules outside the rows.  Its clamped per-lane
        // offsets are computed once, so the loads of every other chunk carry no range checks (kFast); a flipped route
        // meets that chunk first -- in the peeled load below -- so its in-loop loads need no select at all.
        const int p0_last = (nch - 1) * kChunk + lane * kItems;
        const int l0_last = rev ? L - 8 - p0_last : p0_last;
        const int g0_last = (l0_last >= 0 && l0_last + 4 <= L) ? l0_last : 0;
        const int g1_last = (l0_last + 4 >= 0 && l0_last + 8 <= L) ? l0_last + 4 : 0;

        const T* pf_row = (kN != 1 || kCh != 1 || kSingle || lane >= 24) ? nullptr : (lane < 8) ? dt_row[0] : (lane < 16) ? Bk : Ck;

        auto load_chunk = [&](int step, FwdChunk<kN, kCh>& c, auto peeled_tag) __attribute__((always_inline)) {
            constexpr bool peeled = decltype(peeled_tag)::value;
            const int j = rev ? (nch - 1 - step) : step;
            if (kN == 1 && kCh == 1 && !kSingle) {
                if (peeled) {
#pragma unroll
                    for (int a = 1; a < kPrefetchAhead; ++a) prefetch_chunk_l2<T>(pf_row, rev, rev ? j - a : j + a, nch, L, lane);
                }
                prefetch_chunk_l2<T>(pf_row, rev, rev ? j - kPrefetchAhead : j + kPrefetchAhead, nch, L, lane);
            }
            const int p0 = j * kChunk + lane * kItems;
            const int l0 = rev ? L - 8 - p0 : p0;        // scan index of the lowest-address element
            if constexpr (kFast && Elem<T>::kVec == 4) {
                const bool last = rev ? peeled : (j == nch - 1);
                const int g0 = last ? g0_last : l0, g1 = last ? g1_last : l0 + 4;
#pragma unroll
                for (int ch = 0; ch < kCh; ++ch) load8_at<T>(dt_row[ch], g0, g1, c.dt[ch]);
                if (kN == 1) { load8_at<T>(Bk, g0, g1, c.B); load8_at<T>(Ck, g0, g1, c.C); }
            } else {
#pragma unroll
                for (int ch = 0; ch < kCh; ++ch) row_load8<T, kFast>(dt_row[ch], l0, L, vin, c.dt[ch]);
                if (kN == 1) {
                    row_load8<T, kFast>(Bk, l0, L, vin, c.B);
                    row_load8<T, kFast>(Ck, l0, L, vin, c.C);
                }
            }
            if (rev && peeled) {
                // positions >= L come FIRST in a flipped route and must be identity maps (dt = 0): softplus(-inf) = 0,
                // and without softplus dt = (-bias) + bias = 0.  Done here, once, instead of masking inside every chunk.
#pragma unroll
                for (int ch = 0; ch < kCh; ++ch) {
                    const float off = p.delta_softplus ? -INFINITY : -bias[ch];
#pragma unroll
                    for (int i = 0; i < 8; ++i)
                        if (l0 + i < 0) c.dt[ch][i] = off;
                }
            }
        };

        FwdChunk<kN, kCh> c0;           // ONE register set: re-loaded for the next chunk as soon as this one is consumed
        load_chunk(0, c0, std::true_type{});   // in flight while the image is staged
        {
            const T* __restrict__ x = reinterpret_cast<const T*>(p.x);
#pragma unroll
            for (int ch = 0; ch < kCh; ++ch)
                stage_image<T>(x + ((int64_t)b * D + (valid[ch] ? d0 + ch : d0)) * L, xN + ch * Lb, xT + ch * Lb, H, W, L, Lb,
                               valid[ch], tid, 128);
            if (kN == 0)
                for (int i = tid; i < 4 * kCh * kFusedMaxState; i += 128) s_h[i] = 0.0f;
            cta_barrier();              // all 4 warps arrive here, two from each instantiation of the walk
        }

        auto compute_chunk = [&](int step, FwdChunk<kN, kCh>& c, auto&& issue_next) __attribute__((always_inline)) {
            const int j = rev ? (nch - 1 - step) : step;
            const int p0 = j * kChunk + lane * kItems;
            const int l0 = rev ? L - 8 - p0 : p0;
            const int f4s = swz_f4(p0 >> 2);
            const bool in_buf = p0 < Lb;                    // lane has shared-memory backing

            // ---- read every load register once (dt -> dt + bias, B -> B*u, C -> C + 0), then re-load the SAME registers
            // with the next chunk: the loads fly during the whole computation below (scoreboard note in the header)
            f2 dt2[kCh][4], u2[kCh][4], y2[kCh][4], Bu2[kCh][4], C1[4];
#pragma unroll
            for (int ch = 0; ch < kCh; ++ch) {
                float u[8];
                if (in_buf) lds8(xb + ch * Lb, f4s, u);
                else {
#pragma unroll
                    for (int i = 0; i < 8; ++i) u[i] = 0.0f;
                }
                pack8(u, u2[ch]);
                float dtp[8];
                to_pos<rev>(c.dt[ch], dtp);
                pack8(dtp, dt2[ch]);
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) dt2[ch][jj] = add2(dt2[ch][jj], splat2(bias[ch]));
            }
            if (kN == 1) {
                float Bp[8], Cp[8];
                f2 B1[4];
                to_pos<rev>(c.B, Bp); to_pos<rev>(c.C, Cp);
                pack8(Bp, B1); pack8(Cp, C1);
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) {
                    C1[jj] = add2(C1[jj], rt_zero2);
#pragma unroll
                    for (int ch = 0; ch < kCh; ++ch) Bu2[ch][jj] = mul2(B1[jj], u2[ch][jj]);
                }
            }
            issue_next();

            // ---- dt = softplus(delta + bias); y starts as D*u
#pragma unroll
            for (int ch = 0; ch < kCh; ++ch) {
                if (p.delta_softplus) softplus8_vote(dt2[ch]);
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) y2[ch][jj] = mul2(splat2(Dd[ch]), u2[ch][jj]);
            }
            // positions >= L (last chunk only): a forward route meets them after every real position, so whatever they hold
            // never reaches a real one; a flipped route had them turned into identity maps by the peeled load.
            for (int n = 0; n < N; ++n) {
                f2 C2[4];
                if (kN == 1) {
#pragma unroll
                    for (int jj = 0; jj < 4; ++jj) C2[jj] = C1[jj];
                } else {
                    float Bv[8], Cv[8], Bp[8], Cp[8];
                    f2 B2[4];
                    row_load8<T, kFast>(Bk + n * L, l0, L, vin, Bv);
                    row_load8<T, kFast>(Ck + n * L, l0, L, vin, Cv);
                    to_pos<rev>(Bv, Bp); to_pos<rev>(Cv, Cp);
                    pack8(Bp, B2); pack8(Cp, C2);
#pragma unroll
                    for (int jj = 0; jj < 4; ++jj)
#pragma unroll
                        for (int ch = 0; ch < kCh; ++ch) Bu2[ch][jj] = mul2(B2[jj], u2[ch][jj]);
                }
#pragma unroll
                for (int ch = 0; ch < kCh; ++ch) {
                    const float A2 = (kN == 1) ? A2_1[ch] : p.A[kd[ch] * N + n] * kLog2e;
                    f2 a2[4], bu2[4], S2[4], P2[4];
#pragma unroll
                    for (int jj = 0; jj < 4; ++jj) {
                        a2[jj] = ex2_2(mul2(dt2[ch][jj], splat2(A2)));
                        bu2[jj] = mul2(dt2[ch][jj], Bu2[ch][jj]);
                    }
                    float Pr = 1.0f, Sr = 0.0f;
                    if (!rev) {
#pragma unroll
                        for (int jj = 0; jj < 4; ++jj) {
                            Sr = fmaf(a2[jj].x, Sr, bu2[jj].x); Pr *= a2[jj].x; S2[jj].x = Sr; P2[jj].x = Pr;
                            Sr = fmaf(a2[jj].y, Sr, bu2[jj].y); Pr *= a2[jj].y; S2[jj].y = Sr; P2[jj].y = Pr;
                        }
                    } else {
#pragma unroll
                        for (int jj = 3; jj >= 0; --jj) {
                            Sr = fmaf(a2[jj].y, Sr, bu2[jj].y); Pr *= a2[jj].y; S2[jj].y = Sr; P2[jj].y = Pr;
                            Sr = fmaf(a2[jj].x, Sr, bu2[jj].x); Pr *= a2[jj].x; S2[jj].x = Sr; P2[jj].x = Pr;
                        }
                    }
                    float* hs = s_h + (k * kCh + ch) * kFusedMaxState + n;
                    const float carry = (kN == 1) ? carry1[ch] : *hs;
                    float h_out;
                    const float h_in = warp_prefix_p<rev>(Pr, Sr, carry, lane, h_out);
                    const f2 hin2 = splat2(h_in);
#pragma unroll
                    for (int jj = 0; jj < 4; ++jj) y2[ch][jj] = fma2(C2[jj], fma2(P2[jj], hin2, S2[jj]), y2[ch][jj]);
                    if (kN == 1) carry1[ch] = h_out;
                    else { __syncwarp(); if (lane == 0) *hs = h_out; }
                    if (kLane) { if (st_row[ch]) st_row[ch][j * 32 + lane] = h_in; }
                    else if (st_row[ch] && lane == 0) st_row[ch][j * N + n] = h_out;
                }
            }
            // ---- accumulate into the pair's buffer
            const bool first_touch = rev ? (j >= m) : (j < m);
            if (!first_touch && !synced) { pair_barrier(k & 1); synced = true; }
            if (in_buf) {
#pragma unroll
                for (int ch = 0; ch < kCh; ++ch) {
                    float* yc = yb + ch * Lb;
                    float4 o0 = make_float4(0.f, 0.f, 0.f, 0.f), o1 = o0;
                    if (!first_touch) {
                        o0 = *reinterpret_cast<const float4*>(yc + (f4s << 2));
                        o1 = *reinterpret_cast<const float4*>(yc + ((f4s ^ 1) << 2));
                    }
                    const f2 r0 = add2(y2[ch][0], make_float2(o0.x, o0.y)), r1 = add2(y2[ch][1], make_float2(o0.z, o0.w));
                    const f2 r2 = add2(y2[ch][2], make_float2(o1.x, o1.y)), r3 = add2(y2[ch][3], make_float2(o1.z, o1.w));
                    *reinterpret_cast<float4*>(yc + (f4s << 2)) = make_float4(r0.x, r0.y, r1.x, r1.y);
                    *reinterpret_cast<float4*>(yc + ((f4s ^ 1) << 2)) = make_float4(r2.x, r2.y, r3.x, r3.y);
                }
            }
        };

#pragma unroll 1
        for (int step = 0; step < nch; ++step)
            // the re-load is UNCONDITIONAL (the last step re-reads its own chunk, an L1 hit): a conditional one keeps the old
            // values formally alive across the whole step and costs two register-to-register copies of the set per chunk
            compute_chunk(step, c0, [&]() __attribute__((always_inline)) { if (!kSingle) load_chunk(min(step + 1, nch - 1), c0, std::false_type{}); });
    };  // walk
    if (k >= 2) walk(std::true_type{}); else walk(std::false_type{});
    if (!synced) pair_barrier(k & 1);
    __syncthreads();

    // merged output, spatial order: y[p] = yN[p] + yT[w*H + h]
    TO* __restrict__ out = reinterpret_cast<TO*>(p.y);
#pragma unroll
    for (int ch = 0; ch < kCh; ++ch)
        if (valid[ch]) merge_out<TO>(out + ((int64_t)b * D + d0 + ch) * L, yN + ch * Lb, yT + ch * Lb, H, W, tid, 128);
}

// ---- host side --------------------------------------------------------------------------------------------------
constexpr int kChFwd = 1;    // measured on B200 (config-2 shape): 1 channel/CTA (4 CTAs/SM) beats 2 (B/C shared, 2 CTAs/SM)

int ss2d_supported(int64_t D, int64_t N, int64_t H, int64_t W, int dtype, int backward) {
    (void)D; (void)dtype;
    if (N < 1 || N > kFusedMaxState) return 0;
    const int64_t L = H * W;
    if (L <= 0 || L > (1 << 24)) return 0;
    return (backward ? bwd_smem(L, N, 1) : fwd_smem(L, N, kChFwd)) <= kSmemLimit;
}

template <typename T, typename TO, int kN, bool kFast, bool kSingle = false>
static int launch_fwd_k(const xfs_ss2d_fwd_args& a, cudaStream_t st) {
    const size_t smem = fwd_smem(a.H * a.W, a.N, kChFwd);
    const unsigned grid = (unsigned)(a.batch * ((a.D + kChFwd - 1) / kChFwd));
    if (int rc = set_smem(ss2d_fwd_kernel<T, TO, kN, kChFwd, kFast, kSingle>, smem)) return rc;
    ss2d_fwd_kernel<T, TO, kN, kChFwd, kFast, kSingle><<<grid, 128, smem, st>>>(a);
    return check_launch();
}

template <typename T, typename TO>
static int launch_fwd_tt(const xfs_ss2d_fwd_args& a, cudaStream_t st) {
    const int64_t L = a.H * a.W;
    // fast rows: 16-byte aligned tensors and L a multiple of the vector width -> kernels without scalar edge code.
    // Only instantiated for fp32 output (oflex, what XFMamba uses); anything else takes the general kernels.
    const bool fast = std::is_same<TO, float>::value && (L % Elem<T>::kVec == 0) && aligned16(a.delta) && aligned16(a.Bs) &&
                      aligned16(a.Cs);
    if constexpr (std::is_same<TO, float>::value) {
        if (fast && L <= kChunk)     // one chunk per sequence
            return a.N == 1 ? launch_fwd_k<T, TO, 1, true, true>(a, st) : launch_fwd_k<T, TO, 0, true, true>(a, st);
        if (fast) return a.N == 1 ? launch_fwd_k<T, TO, 1, true>(a, st) : launch_fwd_k<T, TO, 0, true>(a, st);
    }
    return a.N == 1 ? launch_fwd_k<T, TO, 1, false>(a, st) : launch_fwd_k<T, TO, 0, false>(a, st);
}

// 16-bit rows, fp32 output, N == 1, L % 8 == 0, more than one chunk: the fast kernel with lane-granular checkpoints
template <typename T>
static int launch_fwd_lane_t(const xfs_ss2d_fwd_args& a, cudaStream_t st) {
    const size_t smem = fwd_smem(a.H * a.W, 1, kChFwd);
    if (int rc = set_smem(ss2d_fwd_kernel<T, float, 1, kChFwd, true, false, true>, smem)) return rc;
    ss2d_fwd_kernel<T, float, 1, kChFwd, true, false, true><<<(unsigned)(a.batch * a.D), 128, smem, st>>>(a);
    return check_launch();
}
int launch_ss2d_fwd_lane16(const xfs_ss2d_fwd_args& a, cudaStream_t st) {
    if (!(aligned16(a.x) && aligned16(a.delta) && aligned16(a.Bs) && aligned16(a.Cs) && aligned16(a.y))) return XFS_ERR_ALIGN;
    return a.dtype == XFS_BF16 ? launch_fwd_lane_t<__nv_bfloat16>(a, st) : launch_fwd_lane_t<__half>(a, st);
}

int launch_ss2d_fwd(const xfs_ss2d_fwd_args& a, cudaStream_t st) {
    const bool o32 = a.out_dtype == XFS_F32;
    switch (a.dtype) {
        case XFS_F32: return launch_fwd_tt<float, float>(a, st);
        case XFS_BF16: return o32 ? launch_fwd_tt<__nv_bfloat16, float>(a, st) : launch_fwd_tt<__nv_bfloat16, __nv_bfloat16>(a, st);
        default: return o32 ? launch_fwd_tt<__half, float>(a, st) : launch_fwd_tt<__half, __half>(a, st);
    }
}

}  // namespace xfs
