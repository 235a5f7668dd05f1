// ss2d_bwd.cu -- fused SS2D backward kernel (design notes in ss2d_fused.cuh)
//
//   smem per channel: xN, xT (u), gN, gT (dy in both layouts), dN, dT (du accumulators, same pair protocol as y)
//   per chunk (walked in the REVERSE of the forward's order) and state n:
//     forward re-scan from the checkpointed state entering the chunk                    -> h_i
//     reverse scan of q_i = a_i (C_i dy_i + q_next)                                     -> g_i = C_i dy_i + q_next = dL/dh_i
//     du += g dt B;  ddt += g (B u + A (h - b));  dA += g dt (h - b);  dB += g dt u;  dC += dy h      (b = dt B u)
//   dBs/dCs are shared by all D channels of a route: 16-byte vector reductions (red.global.add.v4.f32) into the fp32
//   accumulators, which stay L2 resident; dA/dDs/dbias: warp-reduced, one atomic per (route, channel).
#include "ss2d_fused.cuh"

namespace xfs {

template <int kN, int kCh>
struct BwdChunk {            // ADDRESS order, exactly as loaded
    float dt[kCh][8];
    float B[8], C[8];       // kN == 1 only
    float hstart[kCh];      // kN == 1 only: checkpointed state entering the chunk
};

template <typename T, typename TDO, int kN, int kCh, bool kFast>
__global__ void __launch_bounds__(128)
ss2d_bwd_kernel(const xfs_ss2d_bwd_args p) {
    extern __shared__ __align__(16) float smem[];
    const int H = (int)p.H, W = (int)p.W, L = H * W;
    const int Lb = (int)buf_len(L), nch = (L + kChunk - 1) / kChunk;
    const int D = (int)p.D;
    const int N = (kN == 1) ? 1 : (int)p.N;
    const int groups = (D + kCh - 1) / kCh;
    const int b = blockIdx.x / groups;
    const int d0 = (blockIdx.x - b * groups) * kCh;
    const int tid = threadIdx.x, lane = tid & 31, k = tid >> 5;
    const bool transposed = k & 1;
    bool valid[kCh];
#pragma unroll
    for (int ch = 0; ch < kCh; ++ch) valid[ch] = (d0 + ch) < D;

    float* xN = smem;
    float* xT = xN + kCh * Lb;
    float* gN = xT + kCh * Lb;
    float* gT = gN + kCh * Lb;
    float* dN = gT + kCh * Lb;
    float* dT = dN + kCh * Lb;
    float* s_q = dT + kCh * Lb;                       // [4][kCh][kFusedMaxState] reverse carries (kN == 0)
    float* s_dA = s_q + 4 * kCh * kFusedMaxState;     // [4][kCh][kFusedMaxState]

    const float* xb = transposed ? xT : xN;
    const float* gb = transposed ? gT : gN;
    float* db = transposed ? dT : dN;
    const T* __restrict__ delta = reinterpret_cast<const T*>(p.delta);
    T* __restrict__ ddelta = reinterpret_cast<T*>(p.ddelta);
    const T* __restrict__ Bk = reinterpret_cast<const T*>(p.Bs) + ((int64_t)b * 4 + k) * N * L;
    const T* __restrict__ Ck = reinterpret_cast<const T*>(p.Cs) + ((int64_t)b * 4 + k) * N * L;
    float* __restrict__ dBk = p.dBs + ((int64_t)b * 4 + k) * N * L;
    float* __restrict__ dCk = p.dCs + ((int64_t)b * 4 + k) * N * L;
    const bool vin = kFast || (row_vec_ok(delta, L) && row_vec_ok(reinterpret_cast<const T*>(p.Bs), L) &&
                               row_vec_ok(reinterpret_cast<const T*>(p.Cs), L));
    const bool vout = kFast || row_vec_ok(ddelta, L);
    const bool vacc = kFast || ((L % 4 == 0) && aligned16_dev(p.dBs) && aligned16_dev(p.dCs));

    const T* dt_row[kCh];
    T* ddt_row[kCh];
    const float* st_row[kCh];
    float bias[kCh], Dd[kCh], A_1[kCh], qcarry1[kCh], dA1[kCh], dD_acc[kCh], dbias_acc[kCh];
    int kd[kCh];
#pragma unroll
    for (int ch = 0; ch < kCh; ++ch) {
        kd[ch] = k * D + (valid[ch] ? d0 + ch : d0);
        dt_row[ch] = delta + ((int64_t)b * 4 * D + kd[ch]) * L;
        ddt_row[ch] = ddelta + ((int64_t)b * 4 * D + kd[ch]) * L;
        st_row[ch] = p.states + ((int64_t)b * 4 * D + kd[ch]) * nch * N;
        bias[ch] = p.delta_bias ? p.delta_bias[kd[ch]] : 0.0f;
        Dd[ch] = p.Ds ? p.Ds[kd[ch]] : 0.0f;
        A_1[ch] = (kN == 1) ? p.A[kd[ch]] : 0.0f;
        qcarry1[ch] = 0.0f; dA1[ch] = 0.0f; dD_acc[ch] = 0.0f; dbias_acc[ch] = 0.0f;
    }
    const float rt_zero = __int_as_float(p.scans);   // +0.0f (scans == 0), but only known at run time

    // The backward of route k walks its chunks in the reverse of the forward walk: routes 0/1 go nch-1 -> 0 with a
    // reverse (lanes 31->0) adjoint scan, routes 2/3 go 0 -> nch-1 with a lanes 0->31 adjoint scan.
    const int m = nch / 2;             // routes 2/3 first touch [0, m); routes 0/1 first touch [m, nch)
    bool synced = false;

    auto walk = [&](auto rev_tag) __attribute__((always_inline)) {
        constexpr bool rev = decltype(rev_tag)::value;     // the FORWARD walk direction of this route

        auto load_chunk = [&](int step, BwdChunk<kN, kCh>& c) __attribute__((always_inline)) {
            const int j = rev ? step : (nch - 1 - step);
            const int p0 = j * kChunk + lane * kItems;
            const int l0 = rev ? L - 8 - p0 : p0;
            const int jprev = rev ? j + 1 : j - 1;           // chunk the forward walked just before this one
#pragma unroll
            for (int ch = 0; ch < kCh; ++ch) {
                row_load8<T, kFast>(dt_row[ch], l0, L, vin, c.dt[ch]);
                if (kN == 1) c.hstart[ch] = (jprev >= 0 && jprev < nch) ? st_row[ch][jprev] : 0.0f;
            }
            if (kN == 1) {
                row_load8<T, kFast>(Bk, l0, L, vin, c.B);
                row_load8<T, kFast>(Ck, l0, L, vin, c.C);
            }
        };

        BwdChunk<kN, kCh> cur, nxt;
        load_chunk(0, cur);             // in flight while the images are staged
        {
            const T* __restrict__ x = reinterpret_cast<const T*>(p.x);
            const TDO* __restrict__ dyp = reinterpret_cast<const TDO*>(p.dy);
#pragma unroll
            for (int ch = 0; ch < kCh; ++ch) {
                const int64_t row = ((int64_t)b * D + (valid[ch] ? d0 + ch : d0)) * L;
                stage_image<T>(x + row, xN + ch * Lb, xT + ch * Lb, H, W, L, Lb, valid[ch], tid, 128);
                stage_image<TDO>(dyp + row, gN + ch * Lb, gT + ch * Lb, H, W, L, Lb, valid[ch], tid, 128);
            }
            if (kN == 0)
                for (int i = tid; i < 2 * 4 * kCh * kFusedMaxState; i += 128) s_q[i] = 0.0f;
            cta_barrier();
        }

#pragma unroll 1
        for (int step = 0; step < nch; ++step) {
            BwdChunk<kN, kCh>& c = cur;
            const int j = rev ? step : (nch - 1 - step);
            const int p0 = j * kChunk + lane * kItems;
            const int l0 = rev ? L - 8 - p0 : p0;
            const int jprev = rev ? j + 1 : j - 1;
            const int f4s = swz_f4(p0 >> 2);
            const bool in_buf = p0 < Lb;
            const bool tail = p0 + 8 > L;

            // ---- read every load register once, then issue the next chunk's loads (scoreboard note in the header)
            float xraw[kCh][8], Bv[8], Cv[8], hst[kCh];
#pragma unroll
            for (int ch = 0; ch < kCh; ++ch) {
                float dtp[8];
                to_pos<rev>(c.dt[ch], dtp);
#pragma unroll
                for (int i = 0; i < 8; ++i) xraw[ch][i] = dtp[i] + bias[ch];
                hst[ch] = (kN == 1) ? c.hstart[ch] + rt_zero : 0.0f;
            }
            if (kN == 1) {
                float Bp[8], Cp[8];
                to_pos<rev>(c.B, Bp); to_pos<rev>(c.C, Cp);
#pragma unroll
                for (int i = 0; i < 8; ++i) { Bv[i] = Bp[i] + rt_zero; Cv[i] = Cp[i] + rt_zero; }
            }
            if (step + 1 < nch) load_chunk(step + 1, nxt);

            float dt[kCh][8], u[kCh][8], dy[kCh][8], sig[kCh][8], du[kCh][8], ddt[kCh][8];
#pragma unroll
            for (int ch = 0; ch < kCh; ++ch) {
                if (in_buf) { lds8(xb + ch * Lb, f4s, u[ch]); lds8(gb + ch * Lb, f4s, dy[ch]); }
                else {
#pragma unroll
                    for (int i = 0; i < 8; ++i) { u[ch][i] = 0.0f; dy[ch][i] = 0.0f; }
                }
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) {
                    const f2 xx = make_float2(xraw[ch][2 * jj], xraw[ch][2 * jj + 1]);
                    f2 e = splat2(0.0f);
                    const f2 sp = p.delta_softplus ? softplus2(xx, e) : xx;
                    dt[ch][2 * jj] = sp.x; dt[ch][2 * jj + 1] = sp.y;
                    // sigmoid(x) = e / (1 + e); x > 20 -> 1 (softplus is the identity there)
                    sig[ch][2 * jj] = p.delta_softplus ? ((xx.x > 20.0f) ? 1.0f : e.x * rcp(1.0f + e.x)) : 1.0f;
                    sig[ch][2 * jj + 1] = p.delta_softplus ? ((xx.y > 20.0f) ? 1.0f : e.y * rcp(1.0f + e.y)) : 1.0f;
                }
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    if (tail && p0 + i >= L) dt[ch][i] = 0.0f;
                    du[ch][i] = Dd[ch] * dy[ch][i];
                    ddt[ch][i] = 0.0f;
                    dD_acc[ch] = fmaf(dy[ch][i], u[ch][i], dD_acc[ch]);
                }
            }
            for (int n = 0; n < N; ++n) {
                float dBv[8], dCv[8];
                if (kN != 1) {
                    float Bl[8], Cl[8];
                    row_load8<T, kFast>(Bk + n * L, l0, L, vin, Bl);
                    row_load8<T, kFast>(Ck + n * L, l0, L, vin, Cl);
                    to_pos<rev>(Bl, Bv); to_pos<rev>(Cl, Cv);
                }
#pragma unroll
                for (int i = 0; i < 8; ++i) { dBv[i] = 0.0f; dCv[i] = 0.0f; }
#pragma unroll
                for (int ch = 0; ch < kCh; ++ch) {
                    const float An = (kN == 1) ? A_1[ch] : p.A[kd[ch] * N + n];
                    const float A2 = An * kLog2e;
                    float a[8], bu[8], S[8], P[8], Sq[8], Pq[8];
                    float Pr = 1.0f, Sr = 0.0f;
                    const float h_start = (kN == 1) ? hst[ch]
                                                    : ((jprev >= 0 && jprev < nch) ? st_row[ch][jprev * N + n] : 0.0f);
                    float unused, h_in, q_in, q_out;
                    float* qs = s_q + (k * kCh + ch) * kFusedMaxState + n;
                    const float qc = (kN == 1) ? qcarry1[ch] : *qs;
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        a[i] = ex2(dt[ch][i] * A2);
                        bu[i] = (dt[ch][i] * Bv[i]) * u[ch][i];
                    }
                    if (!rev) {
#pragma unroll
                        for (int i = 0; i < 8; ++i) { Sr = fmaf(a[i], Sr, bu[i]); Pr *= a[i]; S[i] = Sr; P[i] = Pr; }
                        h_in = warp_prefix<false>(Pr, Sr, h_start, lane, unused);
                        Pr = 1.0f; Sr = 0.0f;
#pragma unroll
                        for (int i = 7; i >= 0; --i) { Sr = a[i] * fmaf(Cv[i], dy[ch][i], Sr); Pr *= a[i]; Sq[i] = Sr; Pq[i] = Pr; }
                        q_in = warp_prefix<true>(Pr, Sr, qc, lane, q_out);
                    } else {
#pragma unroll
                        for (int i = 7; i >= 0; --i) { Sr = fmaf(a[i], Sr, bu[i]); Pr *= a[i]; S[i] = Sr; P[i] = Pr; }
                        h_in = warp_prefix<true>(Pr, Sr, h_start, lane, unused);
                        Pr = 1.0f; Sr = 0.0f;
#pragma unroll
                        for (int i = 0; i < 8; ++i) { Sr = a[i] * fmaf(Cv[i], dy[ch][i], Sr); Pr *= a[i]; Sq[i] = Sr; Pq[i] = Pr; }
                        q_in = warp_prefix<false>(Pr, Sr, qc, lane, q_out);
                    }
                    float dA_part = 0.0f;
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float h = fmaf(P[i], h_in, S[i]);
                        // q of the element that FOLLOWS i in the forward walk
                        const int inx = rev ? (i == 0 ? 0 : i - 1) : (i == 7 ? 7 : i + 1);
                        const bool edge = rev ? (i == 0) : (i == 7);
                        const float q_next = edge ? q_in : fmaf(Pq[inx], q_in, Sq[inx]);
                        const float gi = fmaf(Cv[i], dy[ch][i], q_next);
                        const float hp = h - bu[i];
                        const float gdt = gi * dt[ch][i];
                        du[ch][i] = fmaf(gdt, Bv[i], du[ch][i]);
                        ddt[ch][i] = fmaf(gi, fmaf(Bv[i], u[ch][i], An * hp), ddt[ch][i]);
                        dA_part = fmaf(gdt, hp, dA_part);
                        if (kCh == 1 || valid[ch]) {
                            dBv[i] = fmaf(gdt, u[ch][i], dBv[i]);
                            dCv[i] = fmaf(dy[ch][i], h, dCv[i]);
                        }
                    }
                    if (kN == 1) { qcarry1[ch] = q_out; dA1[ch] += dA_part; }
                    else {
                        dA_part = warp_sum(dA_part);
                        __syncwarp();
                        if (lane == 0) { *qs = q_out; s_dA[(k * kCh + ch) * kFusedMaxState + n] += dA_part; }
                    }
                }
                // dB / dC of this route at scan positions l0..l0+7 (ascending address order)
                float dBa[8], dCa[8];
                to_pos<rev>(dBv, dBa); to_pos<rev>(dCv, dCa);
                float* dBrow = dBk + n * L;
                float* dCrow = dCk + n * L;
                if (vacc) {     // L % 4 == 0: each 16-byte granule is entirely inside or outside the row
                    if (l0 >= 0 && l0 + 4 <= L) {
                        red_add_v4(dBrow + l0, dBa[0], dBa[1], dBa[2], dBa[3]);
                        red_add_v4(dCrow + l0, dCa[0], dCa[1], dCa[2], dCa[3]);
                    }
                    if (l0 + 4 >= 0 && l0 + 8 <= L) {
                        red_add_v4(dBrow + l0 + 4, dBa[4], dBa[5], dBa[6], dBa[7]);
                        red_add_v4(dCrow + l0 + 4, dCa[4], dCa[5], dCa[6], dCa[7]);
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int l = l0 + i;
                        if (l >= 0 && l < L) {
                            atomicAdd(dBrow + l, dBa[i]);
                            atomicAdd(dCrow + l, dCa[i]);
                        }
                    }
                }
            }
            // ---- ddelta (scan order of the route) and du accumulation (position order, pair protocol)
            const bool first_touch = rev ? (j < m) : (j >= m);
            if (!first_touch && !synced) { pair_barrier(k & 1); synced = true; }
#pragma unroll
            for (int ch = 0; ch < kCh; ++ch) {
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    ddt[ch][i] *= sig[ch][i];
                    dbias_acc[ch] += ddt[ch][i];            // dt = 0 beyond L makes these terms exactly 0
                }
                if (kCh == 1 || valid[ch]) {
                    float dda[8];
                    to_pos<rev>(ddt[ch], dda);
                    row_store8<T, kFast>(ddt_row[ch], l0, L, vout, dda);
                }
                if (in_buf) {
                    if (!first_touch) {
                        float o[8];
                        lds8(db + ch * Lb, f4s, o);
#pragma unroll
                        for (int i = 0; i < 8; ++i) du[ch][i] += o[i];
                    }
                    sts8(db + ch * Lb, f4s, du[ch]);
                }
            }
            cur = nxt;      // register copy (waits for the loads issued a whole chunk ago)
        }
    };  // walk
    if (k >= 2) walk(std::true_type{}); else walk(std::false_type{});
    if (!synced) pair_barrier(k & 1);

    // parameter gradients of this route
#pragma unroll
    for (int ch = 0; ch < kCh; ++ch) {
        if (kN == 1) {
            const float v = warp_sum(dA1[ch]);
            if (lane == 0 && valid[ch]) atomicAdd(p.dA + kd[ch], v);
        } else {
            __syncwarp();
            if (valid[ch])
                for (int n = lane; n < N; n += 32) atomicAdd(p.dA + kd[ch] * N + n, s_dA[(k * kCh + ch) * kFusedMaxState + n]);
        }
        const float vD = warp_sum(dD_acc[ch]), vb = warp_sum(dbias_acc[ch]);
        if (lane == 0 && valid[ch]) {
            if (p.dDs) atomicAdd(p.dDs + kd[ch], vD);
            if (p.ddelta_bias) atomicAdd(p.ddelta_bias + kd[ch], vb);
        }
    }
    __syncthreads();

    // dx[p] = dN[p] + dT[w*H + h]   (CrossScanF.backward = cross-merge of du, models/csm_triton.py:208-225)
    T* __restrict__ dx = reinterpret_cast<T*>(p.dx);
#pragma unroll
    for (int ch = 0; ch < kCh; ++ch)
        if (valid[ch]) merge_out<T>(dx + ((int64_t)b * D + d0 + ch) * L, dN + ch * Lb, dT + ch * Lb, H, W, tid, 128);
}

// ---- host side --------------------------------------------------------------------------------------------------
constexpr int kChBwd = 1;

template <typename T, typename TDO, int kN, bool kFast>
static int launch_bwd_k(const xfs_ss2d_bwd_args& a, cudaStream_t st) {
    const size_t smem = bwd_smem(a.H * a.W, a.N, kChBwd);
    const unsigned grid = (unsigned)(a.batch * ((a.D + kChBwd - 1) / kChBwd));
    if (int rc = set_smem(ss2d_bwd_kernel<T, TDO, kN, kChBwd, kFast>, smem)) return rc;
    ss2d_bwd_kernel<T, TDO, kN, kChBwd, kFast><<<grid, 128, smem, st>>>(a);
    return check_launch();
}

template <typename T, typename TDO>
static int launch_bwd_tt(const xfs_ss2d_bwd_args& a, cudaStream_t st) {
    const int64_t L = a.H * a.W;
    // fast rows (see ss2d_fused.cuh); only instantiated for fp32 upstream gradients (what oflex=True produces)
    const bool fast = std::is_same<TDO, float>::value && (L % Elem<T>::kVec == 0) && (L % 4 == 0) && aligned16(a.delta) &&
                      aligned16(a.Bs) && aligned16(a.Cs) && aligned16(a.ddelta) && aligned16(a.dBs) && aligned16(a.dCs);
    if constexpr (std::is_same<TDO, float>::value) {
        if (fast) return a.N == 1 ? launch_bwd_k<T, TDO, 1, true>(a, st) : launch_bwd_k<T, TDO, 0, true>(a, st);
    }
    return a.N == 1 ? launch_bwd_k<T, TDO, 1, false>(a, st) : launch_bwd_k<T, TDO, 0, false>(a, st);
}

int launch_ss2d_bwd(const xfs_ss2d_bwd_args& a, cudaStream_t st) {
    const bool g32 = a.dout_dtype == XFS_F32;
    switch (a.dtype) {
        case XFS_F32: return launch_bwd_tt<float, float>(a, st);
        case XFS_BF16: return g32 ? launch_bwd_tt<__nv_bfloat16, float>(a, st) : launch_bwd_tt<__nv_bfloat16, __nv_bfloat16>(a, st);
        default: return g32 ? launch_bwd_tt<__half, float>(a, st) : launch_bwd_tt<__half, __half>(a, st);
    }
}

}  // namespace xfs
