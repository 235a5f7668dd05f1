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

// kSingle: see ss2d_fwd.cu
template <typename T, typename TDO, int kN, int kCh, bool kFast, bool kSingle>
__global__ void __launch_bounds__(128, kSingle ? 5 : 3)
ss2d_bwd_kernel(const xfs_ss2d_bwd_args p) {
    extern __shared__ __align__(16) float smem[];
    const int H = (int)p.H, W = (int)p.W, L = H * W;
    const int Lb = (int)buf_len(L);
    const int nch = kSingle ? 1 : (L + kChunk - 1) / kChunk;
    const int D = (int)p.D;
    const int N = (kN == 1) ? 1 : (int)p.N;
    const int groups = (D + kCh - 1) / kCh;
    // batch index fastest: the CTAs resident at any moment then belong to as many different batch images as possible, so
    // few of them add into the same dB / dC rows at the same time (those rows are shared by all channels of one image)
    const int b = blockIdx.x % (int)p.batch;
    const int d0 = (blockIdx.x / (int)p.batch) * kCh;
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
    const int rep = p.acc_replicas > 1 ? d0 % p.acc_replicas : 0;      // accumulator replica of this channel (see xfscan.h)
    float* __restrict__ dBk = p.dBs + (((int64_t)rep * p.batch + b) * 4 + k) * N * L;
    float* __restrict__ dCk = p.dCs + (((int64_t)rep * p.batch + b) * 4 + k) * N * L;
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

        const T* pf_row = (kN != 1 || kCh != 1 || kSingle || lane >= 24) ? nullptr : (lane < 8) ? dt_row[0] : (lane < 16) ? Bk : Ck;

        // Only the LAST chunk (spatial positions >= L) has 16-byte granules outside the rows: its per-lane offsets and
        // validity flags are computed once, so every other chunk loads, stores and reduces without range checks (kFast)
        const int p0_last = (nch - 1) * kChunk + lane * kItems;
        const int l0_last = rev ? L - 8 - p0_last : p0_last;
        const bool okA_last = l0_last >= 0 && l0_last + 4 <= L, okB_last = l0_last + 4 >= 0 && l0_last + 8 <= L;
        const int g0_last = okA_last ? l0_last : 0, g1_last = okB_last ? l0_last + 4 : 0;
        constexpr bool kVec4 = kFast && Elem<T>::kVec == 4;

        auto load_chunk = [&](int step, BwdChunk<kN, kCh>& c) __attribute__((always_inline)) {
            const int j = rev ? step : (nch - 1 - step);
            if (kN == 1 && kCh == 1 && !kSingle) {
                if (step == 0) {
#pragma unroll
                    for (int a = 1; a < kPrefetchAhead; ++a) prefetch_chunk_l2<T>(pf_row, rev, rev ? j + a : j - a, nch, L, lane);
                }
                prefetch_chunk_l2<T>(pf_row, rev, rev ? j + kPrefetchAhead : j - kPrefetchAhead, nch, L, lane);
            }
            const int p0 = j * kChunk + lane * kItems;
            const int l0 = rev ? L - 8 - p0 : p0;
            const int jprev = rev ? j + 1 : j - 1;           // chunk the forward walked just before this one
            const bool last = (j == nch - 1);
            const int g0 = last ? g0_last : l0, g1 = last ? g1_last : l0 + 4;
#pragma unroll
            for (int ch = 0; ch < kCh; ++ch) {
                if constexpr (kVec4) load8_at<T>(dt_row[ch], g0, g1, c.dt[ch]);
                else row_load8<T, kFast>(dt_row[ch], l0, L, vin, c.dt[ch]);
                if (kN == 1) c.hstart[ch] = (jprev >= 0 && jprev < nch) ? st_row[ch][jprev] : 0.0f;
            }
            if (kN == 1) {
                if constexpr (kVec4) { load8_at<T>(Bk, g0, g1, c.B); load8_at<T>(Ck, g0, g1, c.C); }
                else {
                    row_load8<T, kFast>(Bk, l0, L, vin, c.B);
                    row_load8<T, kFast>(Ck, l0, L, vin, c.C);
                }
            }
        };

        BwdChunk<kN, kCh> c;            // ONE register set: re-loaded for the next chunk as soon as this one is consumed
        load_chunk(0, c);               // in flight while the images are staged
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
        f2 dD2[kCh], dbias2[kCh], dA2[kCh];
#pragma unroll
        for (int ch = 0; ch < kCh; ++ch) { dD2[ch] = splat2(0.0f); dbias2[ch] = splat2(0.0f); dA2[ch] = splat2(0.0f); }

#pragma unroll 1
        for (int step = 0; step < nch; ++step) {
            const int j = rev ? step : (nch - 1 - step);
            const int p0 = j * kChunk + lane * kItems;
            const int l0 = rev ? L - 8 - p0 : p0;
            const int jprev = rev ? j + 1 : j - 1;
            const int f4s = swz_f4(p0 >> 2);
            const bool in_buf = p0 < Lb;
            const bool tail = p0 + 8 > L;

            // ---- consume every load register (dt -> dt + bias, B -> copy and B*u, C -> C*dy), then re-load the SAME
            // registers with the next chunk: the loads fly during the whole computation below and nothing below waits on
            // a load scoreboard (see ss2d_fused.cuh)
            f2 x2[kCh][4], u2[kCh][4], dy2[kCh][4], B2[4], Bu2[kCh][4], Cdy2[kCh][4];
            float hst[kCh];
#pragma unroll
            for (int ch = 0; ch < kCh; ++ch) {
                float u[8], dy[8], dtp[8];
                if (in_buf) { lds8(xb + ch * Lb, f4s, u); lds8(gb + ch * Lb, f4s, dy); }
                else {
#pragma unroll
                    for (int i = 0; i < 8; ++i) { u[i] = 0.0f; dy[i] = 0.0f; }
                }
                pack8(u, u2[ch]); pack8(dy, dy2[ch]);
                to_pos<rev>(c.dt[ch], dtp);
                pack8(dtp, x2[ch]);
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) x2[ch][jj] = add2(x2[ch][jj], splat2(bias[ch]));
                hst[ch] = (kN == 1) ? c.hstart[ch] + rt_zero : 0.0f;
            }
            if (kN == 1) {
                float Bp[8], Cp[8];
                f2 C2[4];
                to_pos<rev>(c.B, Bp); to_pos<rev>(c.C, Cp);
                pack8(Bp, B2); pack8(Cp, C2);
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) {
#pragma unroll
                    for (int ch = 0; ch < kCh; ++ch) {
                        Bu2[ch][jj] = mul2(B2[jj], u2[ch][jj]);
                        Cdy2[ch][jj] = mul2(C2[jj], dy2[ch][jj]);
                    }
                    B2[jj] = add2(B2[jj], splat2(rt_zero));          // private copy: B is needed again for du
                }
            }
            if (!kSingle) load_chunk(min(step + 1, nch - 1), c);   // unconditional, see ss2d_fwd.cu

            f2 dt2[kCh][4], sig2[kCh][4], du2[kCh][4], ddt2[kCh][4];
#pragma unroll
            for (int ch = 0; ch < kCh; ++ch) {
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) {
                    const f2 xx = x2[ch][jj];
                    f2 e = splat2(0.0f);
                    dt2[ch][jj] = p.delta_softplus ? softplus2(xx, e) : xx;
                    // sigmoid(x) = e / (1 + e); x > 20 -> 1 (softplus is the identity there)
                    const f2 w = add2(e, splat2(1.0f));
                    f2 sg = mul2(e, make_float2(rcp(w.x), rcp(w.y)));
                    sg.x = (xx.x > 20.0f) ? 1.0f : sg.x;
                    sg.y = (xx.y > 20.0f) ? 1.0f : sg.y;
                    sig2[ch][jj] = p.delta_softplus ? sg : splat2(1.0f);
                    du2[ch][jj] = mul2(splat2(Dd[ch]), dy2[ch][jj]);
                    ddt2[ch][jj] = splat2(0.0f);
                    dD2[ch] = fma2(dy2[ch][jj], u2[ch][jj], dD2[ch]);
                }
                if (tail) {
#pragma unroll
                    for (int jj = 0; jj < 4; ++jj) {
                        if (p0 + 2 * jj >= L) dt2[ch][jj].x = 0.0f;
                        if (p0 + 2 * jj + 1 >= L) dt2[ch][jj].y = 0.0f;
                    }
                }
            }
            for (int n = 0; n < N; ++n) {
                f2 dB2[4], dC2[4];
                if (kN != 1) {
                    float Bl[8], Cl[8], Bp[8], Cp[8];
                    f2 C2[4];
                    row_load8<T, kFast>(Bk + n * L, l0, L, vin, Bl);
                    row_load8<T, kFast>(Ck + n * L, l0, L, vin, Cl);
                    to_pos<rev>(Bl, Bp); to_pos<rev>(Cl, Cp);
                    pack8(Bp, B2); pack8(Cp, C2);
#pragma unroll
                    for (int jj = 0; jj < 4; ++jj)
#pragma unroll
                        for (int ch = 0; ch < kCh; ++ch) {
                            Bu2[ch][jj] = mul2(B2[jj], u2[ch][jj]);
                            Cdy2[ch][jj] = mul2(C2[jj], dy2[ch][jj]);
                        }
                }
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) { dB2[jj] = splat2(0.0f); dC2[jj] = splat2(0.0f); }
#pragma unroll
                for (int ch = 0; ch < kCh; ++ch) {
                    const float An = (kN == 1) ? A_1[ch] : p.A[kd[ch] * N + n];
                    const float A2 = An * kLog2e;
                    f2 a2[4], bu2[4], S2[4], P2[4], Sq2[4], Pq2[4];
                    float a[8], bu[8], cd[8], S[8], P[8], Sq[8], Pq[8];
                    const float h_start = (kN == 1) ? hst[ch]
                                                    : ((jprev >= 0 && jprev < nch) ? st_row[ch][jprev * N + n] : 0.0f);
                    float q_out;
                    float* qs = s_q + (k * kCh + ch) * kFusedMaxState + n;
                    const float qc = (kN == 1) ? qcarry1[ch] : *qs;
#pragma unroll
                    for (int jj = 0; jj < 4; ++jj) {
                        a2[jj] = ex2_2(mul2(dt2[ch][jj], splat2(A2)));
                        bu2[jj] = mul2(dt2[ch][jj], Bu2[ch][jj]);
                    }
                    unpack8(a2, a); unpack8(bu2, bu); unpack8(Cdy2[ch], cd);
                    // forward re-scan (walk order) and adjoint scan (opposite order): serial folds, scalar
                    float Pr = 1.0f, Sr = 0.0f;
#pragma unroll
                    for (int ii = 0; ii < 8; ++ii) {
                        const int i = rev ? 7 - ii : ii;
                        Sr = fmaf(a[i], Sr, bu[i]); Pr *= a[i]; S[i] = Sr; P[i] = Pr;
                    }
                    float Pqr = 1.0f, Sqr = 0.0f;
#pragma unroll
                    for (int ii = 0; ii < 8; ++ii) {
                        const int i = rev ? ii : 7 - ii;
                        Sqr = a[i] * (cd[i] + Sqr); Pqr *= a[i]; Sq[i] = Sqr; Pq[i] = Pqr;
                    }
                    float h_in, q_in;
                    warp_prefix_dual<rev>(Pr, Sr, h_start, Pqr, Sqr, qc, lane, h_in, q_in, q_out);
                    pack8(S, S2); pack8(P, P2); pack8(Sq, Sq2); pack8(Pq, Pq2);
                    // element-wise part on packed pairs
                    f2 q2[4], gi2[4];
                    float q[8], gi[8];
#pragma unroll
                    for (int jj = 0; jj < 4; ++jj) q2[jj] = fma2(Pq2[jj], splat2(q_in), Sq2[jj]);      // q_i (inclusive)
                    unpack8(q2, q);
#pragma unroll
                    for (int i = 0; i < 8; ++i) {       // g_i = C_i dy_i + q of the element that FOLLOWS i in the forward walk
                        const float qn = rev ? (i == 0 ? q_in : q[i == 0 ? 0 : i - 1]) : (i == 7 ? q_in : q[i == 7 ? 7 : i + 1]);
                        gi[i] = cd[i] + qn;
                    }
                    pack8(gi, gi2);
#pragma unroll
                    for (int jj = 0; jj < 4; ++jj) {
                        const f2 h2 = fma2(P2[jj], splat2(h_in), S2[jj]);
                        const f2 hp2 = fma2(bu2[jj], splat2(-1.0f), h2);                 // a_i * h_{i-1}
                        const f2 gdt2 = mul2(gi2[jj], dt2[ch][jj]);
                        du2[ch][jj] = fma2(gdt2, B2[jj], du2[ch][jj]);
                        ddt2[ch][jj] = fma2(gi2[jj], fma2(splat2(An), hp2, Bu2[ch][jj]), ddt2[ch][jj]);
                        dA2[ch] = fma2(gdt2, hp2, dA2[ch]);
                        if (kCh == 1 || valid[ch]) {
                            dB2[jj] = fma2(gdt2, u2[ch][jj], dB2[jj]);
                            dC2[jj] = fma2(dy2[ch][jj], h2, dC2[jj]);
                        }
                    }
                    if (kN == 1) qcarry1[ch] = q_out;
                    else {
                        float dA_part = warp_sum(dA2[ch].x + dA2[ch].y);
                        dA2[ch] = splat2(0.0f);
                        __syncwarp();
                        if (lane == 0) { *qs = q_out; s_dA[(k * kCh + ch) * kFusedMaxState + n] += dA_part; }
                    }
                }
                // dB / dC of this route at scan positions l0..l0+7 (ascending address order)
                float dBv[8], dCv[8], dBa[8], dCa[8];
                unpack8(dB2, dBv); unpack8(dC2, dCv);
                to_pos<rev>(dBv, dBa); to_pos<rev>(dCv, dCa);
                float* dBrow = dBk + n * L;
                float* dCrow = dCk + n * L;
                if (vacc) {     // L % 4 == 0: each 16-byte granule is entirely inside or outside the row
                    const bool lastc = (j == nch - 1);
                    if (!lastc || okA_last) {
                        red_add_v4(dBrow + l0, dBa[0], dBa[1], dBa[2], dBa[3]);
                        red_add_v4(dCrow + l0, dCa[0], dCa[1], dCa[2], dCa[3]);
                    }
                    if (!lastc || okB_last) {
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
                float ddt[8], du[8];
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) {
                    ddt2[ch][jj] = mul2(ddt2[ch][jj], sig2[ch][jj]);
                    dbias2[ch] = add2(dbias2[ch], ddt2[ch][jj]);       // dt = 0 beyond L makes these terms exactly 0
                }
                unpack8(ddt2[ch], ddt); unpack8(du2[ch], du);
                if (kCh == 1 || valid[ch]) {
                    float dda[8];
                    to_pos<rev>(ddt, dda);
                    if constexpr (kVec4) {
                        const bool lastc = (j == nch - 1);
                        if (!lastc || okA_last)
                            stg16(ddt_row[ch] + l0, make_uint4(__float_as_uint(dda[0]), __float_as_uint(dda[1]), __float_as_uint(dda[2]), __float_as_uint(dda[3])));
                        if (!lastc || okB_last)
                            stg16(ddt_row[ch] + l0 + 4, make_uint4(__float_as_uint(dda[4]), __float_as_uint(dda[5]), __float_as_uint(dda[6]), __float_as_uint(dda[7])));
                    } else {
                        row_store8<T, kFast>(ddt_row[ch], l0, L, vout, dda);
                    }
                }
                if (in_buf) {
                    if (!first_touch) {
                        float o[8];
                        lds8(db + ch * Lb, f4s, o);
#pragma unroll
                        for (int i = 0; i < 8; ++i) du[i] += o[i];
                    }
                    sts8(db + ch * Lb, f4s, du);
                }
            }
        }
#pragma unroll
        for (int ch = 0; ch < kCh; ++ch) {
            dD_acc[ch] = dD2[ch].x + dD2[ch].y;
            dbias_acc[ch] = dbias2[ch].x + dbias2[ch].y;
            dA1[ch] = dA2[ch].x + dA2[ch].y;
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

template <typename T, typename TDO, int kN, bool kFast, bool kSingle = false>
static int launch_bwd_k(const xfs_ss2d_bwd_args& a, cudaStream_t st) {
    const size_t smem = bwd_smem(a.H * a.W, a.N, kChBwd);
    const unsigned grid = (unsigned)(a.batch * ((a.D + kChBwd - 1) / kChBwd));
    if (int rc = set_smem(ss2d_bwd_kernel<T, TDO, kN, kChBwd, kFast, kSingle>, smem)) return rc;
    ss2d_bwd_kernel<T, TDO, kN, kChBwd, kFast, kSingle><<<grid, 128, smem, st>>>(a);
    return check_launch();
}

template <typename T, typename TDO>
static int launch_bwd_tt(const xfs_ss2d_bwd_args& a, cudaStream_t st) {
    const int64_t L = a.H * a.W;
    // fast rows (see ss2d_fused.cuh); only instantiated for fp32 upstream gradients (what oflex=True produces)
    const bool fast = std::is_same<TDO, float>::value && (L % Elem<T>::kVec == 0) && (L % 4 == 0) && aligned16(a.delta) &&
                      aligned16(a.Bs) && aligned16(a.Cs) && aligned16(a.ddelta) && aligned16(a.dBs) && aligned16(a.dCs);
    if constexpr (std::is_same<TDO, float>::value) {
        if (fast && L <= kChunk)     // one chunk per sequence
            return a.N == 1 ? launch_bwd_k<T, TDO, 1, true, true>(a, st) : launch_bwd_k<T, TDO, 0, true, true>(a, st);
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
