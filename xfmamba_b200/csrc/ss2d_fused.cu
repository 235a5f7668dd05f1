// ss2d_fused.cu -- fused SS2D core for sm_100a: CrossScan gather + S6 selective scan + CrossMerge in ONE kernel
// (forward) and the whole gradient in ONE kernel (backward).
//
// Replaces the operator sequence of SS2Dv2.forward_corev2 (reference models/fusion_vmamba.py:1145,1170-1174):
//     xs = cross_scan_fn(x); ys = selective_scan_fn(xs, dts, As, Bs, Cs, Ds, delta_bias, True); y = cross_merge_fn(ys)
// which moves ~22 elements per (b, d, l) through HBM (xs written 4x and re-read, ys written 4x and re-read).  Here x is
// read once, delta/B/C are streamed once, the merged y is written once: 6 elements per (b, d, l) + B/C.
//
// Work decomposition
//   CTA  = one batch image x kCh adjacent channels; 4 warps = the 4 routes, all running concurrently.
//   The channels of a CTA share the B/C values of a route (they live in registers once per chunk).
//   Shared memory holds, per channel, the image in row-major order (xN) and column-major order (xT) and the two
//   accumulators yN (routes 0+2) and yT (routes 1+3), all indexed by POSITION (ss2d_tiles.cuh).
//   Routes 2/3 walk the same position chunks as routes 0/1 but from the far end (reverse warp scan), so a route and
//   its flip never touch the same accumulator chunk in the same half of the walk: the first half stores, then one
//   64-thread named barrier, then the second half read-modify-writes what the partner stored.  No atomics, and the
//   sum (y0 + y2) + (y1 + y3) is evaluated in the reference's order (models/csm_triton.py:61-62).
//   The S6 recurrence h_l = exp(dt_l A) h_{l-1} + dt_l B_l u_l is scanned per 256-position chunk with a warp-shuffle
//   scan of affine maps (xfscan_common.cuh), state carried in registers (N == 1) or shared memory (N > 1).
//   delta/B/C of the NEXT chunk are loaded (128-bit, register double buffer) before the current chunk is computed, and
//   the element-wise arithmetic runs on packed fp32 pairs (FFMA2/FMUL2/FADD2).
#include "ss2d_tiles.cuh"

namespace xfs {

constexpr int kChFwd = 2;                    // channels per CTA, forward
constexpr int kChBwd = 1;                    // channels per CTA, backward (register + shared-memory budget)
constexpr int kFusedMaxState = 64;           // states carried in smem for N > 1

__device__ __forceinline__ void pair_barrier(int pair) {   // the 2 warps of a route pair (routes k and k+2)
    asm volatile("bar.sync %0, 64;" ::"r"(pair + 1) : "memory");
}

__device__ __forceinline__ void pack8(const float (&v)[8], f2 (&o)[4]) {
#pragma unroll
    for (int j = 0; j < 4; ++j) o[j] = make_float2(v[2 * j], v[2 * j + 1]);
}

// =========================================================================================================
// forward
// =========================================================================================================
template <typename T, int kN, int kCh>
struct FwdChunk {            // raw inputs of one chunk of one route (position order)
    float dt[kCh][8];
    float B[8], C[8];       // kN == 1 only
};

template <typename T, typename TO, int kN, int kCh>   // kN == 1: single state in registers; kN == 0: runtime N (<= kFusedMaxState)
__global__ void __launch_bounds__(128)
ss2d_fwd_kernel(const xfs_ss2d_fwd_args p) {
    extern __shared__ __align__(16) float smem[];
    const int H = (int)p.H, W = (int)p.W, L = H * W;
    const int Lb = (int)buf_len(L), nch = Lb / kChunk;
    const int D = (int)p.D;
    const int N = (kN == 1) ? 1 : (int)p.N;
    const int groups = (D + kCh - 1) / kCh;
    const int b = blockIdx.x / groups;
    const int d0 = (blockIdx.x - b * groups) * kCh;
    const int tid = threadIdx.x, lane = tid & 31, k = tid >> 5;     // warp k runs route k
    const bool rev = k >= 2, transposed = k & 1;
    bool valid[kCh];
#pragma unroll
    for (int ch = 0; ch < kCh; ++ch) valid[ch] = (d0 + ch) < D;

    float* xN = smem;                   // [kCh][Lb]
    float* xT = xN + kCh * Lb;
    float* yN = xT + kCh * Lb;
    float* yT = yN + kCh * Lb;
    float* s_h = yT + kCh * Lb;         // [4][kCh][kFusedMaxState] (kN == 0 only)

    const T* __restrict__ x = reinterpret_cast<const T*>(p.x);
#pragma unroll
    for (int ch = 0; ch < kCh; ++ch)
        stage_image<T>(x + ((int64_t)b * D + (valid[ch] ? d0 + ch : d0)) * L, xN + ch * Lb, xT + ch * Lb, H, W, L, Lb, valid[ch], tid, 128);
    if (kN == 0)
        for (int i = tid; i < 4 * kCh * kFusedMaxState; i += 128) s_h[i] = 0.0f;
    __syncthreads();

    const float* xb = transposed ? xT : xN;
    float* yb = transposed ? yT : yN;
    const T* __restrict__ delta = reinterpret_cast<const T*>(p.delta);
    const T* __restrict__ Bk = reinterpret_cast<const T*>(p.Bs) + ((int64_t)b * 4 + k) * N * L;
    const T* __restrict__ Ck = reinterpret_cast<const T*>(p.Cs) + ((int64_t)b * 4 + k) * N * L;
    const bool vin = row_vec_ok(delta, L) && row_vec_ok(reinterpret_cast<const T*>(p.Bs), L) &&
                     row_vec_ok(reinterpret_cast<const T*>(p.Cs), L);

    const T* dt_row[kCh];
    float* st_row[kCh];
    float bias[kCh], Dd[kCh], A2_1[kCh], carry1[kCh];
    int kd[kCh];
#pragma unroll
    for (int ch = 0; ch < kCh; ++ch) {
        kd[ch] = k * D + (valid[ch] ? d0 + ch : d0);
        dt_row[ch] = delta + ((int64_t)b * 4 * D + kd[ch]) * L;
        st_row[ch] = (p.states && valid[ch]) ? p.states + ((int64_t)b * 4 * D + kd[ch]) * nch * N : nullptr;
        bias[ch] = p.delta_bias ? p.delta_bias[kd[ch]] : 0.0f;
        Dd[ch] = p.Ds ? p.Ds[kd[ch]] : 0.0f;
        A2_1[ch] = (kN == 1) ? p.A[kd[ch]] * kLog2e : 0.0f;
        carry1[ch] = 0.0f;
    }

    const int m = (nch + 1) / 2;        // chunks [0, m) are first touched by the forward route, [m, nch) by its flip
    bool synced = false;

    auto load_chunk = [&](int step, FwdChunk<T, kN, kCh>& c) {
        const int j = rev ? (nch - 1 - step) : step;
        const int p0 = j * kChunk + lane * kItems;
        const int l0 = rev ? L - 8 - p0 : p0;            // scan index of the lowest-address element
#pragma unroll
        for (int ch = 0; ch < kCh; ++ch) {
            load8(dt_row[ch], l0, L, vin, c.dt[ch]);
            if (rev) reverse8(c.dt[ch]);
        }
        if (kN == 1) {
            load8(Bk, l0, L, vin, c.B);
            load8(Ck, l0, L, vin, c.C);
            if (rev) { reverse8(c.B); reverse8(c.C); }
        }
    };

    auto compute_chunk = [&](int step, FwdChunk<T, kN, kCh>& c) {
        const int j = rev ? (nch - 1 - step) : step;
        const int p0 = j * kChunk + lane * kItems;
        const int l0 = rev ? L - 8 - p0 : p0;
        const int f4s = swz_f4(p0 >> 2);
        const bool tail = p0 + 8 > L;

        f2 dt2[kCh][4], u2[kCh][4], y2[kCh][4];
#pragma unroll
        for (int ch = 0; ch < kCh; ++ch) {
            float u[8];
            lds8(xb + ch * Lb, f4s, u);
            pack8(u, u2[ch]);
            pack8(c.dt[ch], dt2[ch]);
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
                const f2 xx = add2(dt2[ch][jj], splat2(bias[ch]));
                f2 e;
                dt2[ch][jj] = p.delta_softplus ? softplus2(xx, e) : xx;
                y2[ch][jj] = mul2(splat2(Dd[ch]), u2[ch][jj]);
            }
            if (tail) {                                   // positions >= L: identity map (dt = 0)
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) {
                    if (p0 + 2 * jj >= L) dt2[ch][jj].x = 0.0f;
                    if (p0 + 2 * jj + 1 >= L) dt2[ch][jj].y = 0.0f;
                }
            }
        }
        for (int n = 0; n < N; ++n) {
            f2 B2[4], C2[4];
            if (kN == 1) { pack8(c.B, B2); pack8(c.C, C2); }
            else {
                float Bv[8], Cv[8];
                load8(Bk + n * L, l0, L, vin, Bv);
                load8(Ck + n * L, l0, L, vin, Cv);
                if (rev) { reverse8(Bv); reverse8(Cv); }
                pack8(Bv, B2); pack8(Cv, C2);
            }
#pragma unroll
            for (int ch = 0; ch < kCh; ++ch) {
                const float A2 = (kN == 1) ? A2_1[ch] : p.A[kd[ch] * N + n] * kLog2e;
                f2 a2[4], bu2[4], S2[4], P2[4];
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) {
                    a2[jj] = ex2_2(mul2(dt2[ch][jj], splat2(A2)));
                    bu2[jj] = mul2(mul2(dt2[ch][jj], B2[jj]), u2[ch][jj]);
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
                const float h_in = rev ? warp_prefix<true>(Pr, Sr, carry, lane, h_out)
                                       : warp_prefix<false>(Pr, Sr, carry, lane, h_out);
                const f2 hin2 = splat2(h_in);
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) y2[ch][jj] = fma2(C2[jj], fma2(P2[jj], hin2, S2[jj]), y2[ch][jj]);
                if (kN == 1) carry1[ch] = h_out;
                else { __syncwarp(); if (lane == 0) *hs = h_out; }
                if (st_row[ch] && lane == 0) st_row[ch][j * N + n] = h_out;
            }
        }
        // accumulate into the pair's buffer
        const bool first_touch = rev ? (j >= m) : (j < m);
        if (!first_touch && !synced) { pair_barrier(k & 1); synced = true; }
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
    };

    {
        FwdChunk<T, kN, kCh> ca, cb;
        load_chunk(0, ca);
#pragma unroll 1
        for (int step = 0; step < nch; step += 2) {
            if (step + 1 < nch) load_chunk(step + 1, cb);
            compute_chunk(step, ca);
            if (step + 1 < nch) {
                if (step + 2 < nch) load_chunk(step + 2, ca);
                compute_chunk(step + 1, cb);
            }
        }
    }
    if (!synced) pair_barrier(k & 1);
    __syncthreads();

    // merged output, spatial order: y[p] = yN[p] + yT[w*H + h]
    TO* __restrict__ out = reinterpret_cast<TO*>(p.y);
#pragma unroll
    for (int ch = 0; ch < kCh; ++ch)
        if (valid[ch]) merge_out<TO>(out + ((int64_t)b * D + d0 + ch) * L, yN + ch * Lb, yT + ch * Lb, H, W, tid, 128);
}

// =========================================================================================================
// backward
//   smem per channel: xN, xT (u), gN, gT (dy in both layouts), dN, dT (du accumulators, same pair protocol as y)
//   per chunk (walked in the REVERSE of the forward's order) and state n:
//     forward re-scan from the checkpointed state entering the chunk, reverse scan for g = dL/dh, then the
//     closed forms listed in selective_scan.cu.  dBs/dCs: the CTA's channels are summed in registers, then one
//     fp32 atomic per (route, n, l); dA/dDs/dbias: warp-reduced, one atomic per (route, channel).
// =========================================================================================================
template <typename T, int kN, int kCh>
struct BwdChunk {
    float dt[kCh][8];
    float B[8], C[8];
    float hstart[kCh];      // kN == 1: checkpointed state entering the chunk
};

template <typename T, typename TDO, int kN, int kCh>
__global__ void __launch_bounds__(128)
ss2d_bwd_kernel(const xfs_ss2d_bwd_args p) {
    extern __shared__ __align__(16) float smem[];
    const int H = (int)p.H, W = (int)p.W, L = H * W;
    const int Lb = (int)buf_len(L), nch = Lb / kChunk;
    const int D = (int)p.D;
    const int N = (kN == 1) ? 1 : (int)p.N;
    const int groups = (D + kCh - 1) / kCh;
    const int b = blockIdx.x / groups;
    const int d0 = (blockIdx.x - b * groups) * kCh;
    const int tid = threadIdx.x, lane = tid & 31, k = tid >> 5;
    const bool rev = k >= 2, transposed = k & 1;     // `rev` is the FORWARD walk direction of this route
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
    __syncthreads();

    const float* xb = transposed ? xT : xN;
    const float* gb = transposed ? gT : gN;
    float* db = transposed ? dT : dN;
    const T* __restrict__ delta = reinterpret_cast<const T*>(p.delta);
    T* __restrict__ ddelta = reinterpret_cast<T*>(p.ddelta);
    const T* __restrict__ Bk = reinterpret_cast<const T*>(p.Bs) + ((int64_t)b * 4 + k) * N * L;
    const T* __restrict__ Ck = reinterpret_cast<const T*>(p.Cs) + ((int64_t)b * 4 + k) * N * L;
    float* __restrict__ dBk = p.dBs + ((int64_t)b * 4 + k) * N * L;
    float* __restrict__ dCk = p.dCs + ((int64_t)b * 4 + k) * N * L;
    const bool vin = row_vec_ok(delta, L) && row_vec_ok(reinterpret_cast<const T*>(p.Bs), L) &&
                     row_vec_ok(reinterpret_cast<const T*>(p.Cs), L);
    const bool vout = row_vec_ok(ddelta, L);
    const bool vacc = (L % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.dBs) & 15) == 0) && ((reinterpret_cast<uintptr_t>(p.dCs) & 15) == 0);

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

    // The backward of route k walks its chunks in the reverse of the forward walk: routes 0/1 go nch-1 -> 0 with a
    // reverse (lanes 31->0) adjoint scan, routes 2/3 go 0 -> nch-1 with a lanes 0->31 adjoint scan.
    const int m = nch / 2;             // routes 2/3 first touch [0, m); routes 0/1 first touch [m, nch)
    bool synced = false;

    auto load_chunk = [&](int step, BwdChunk<T, kN, kCh>& c) {
        const int j = rev ? step : (nch - 1 - step);
        const int p0 = j * kChunk + lane * kItems;
        const int l0 = rev ? L - 8 - p0 : p0;
        const int jprev = rev ? j + 1 : j - 1;           // chunk the forward walked just before this one
#pragma unroll
        for (int ch = 0; ch < kCh; ++ch) {
            load8(dt_row[ch], l0, L, vin, c.dt[ch]);
            if (rev) reverse8(c.dt[ch]);
            if (kN == 1) c.hstart[ch] = (jprev >= 0 && jprev < nch) ? st_row[ch][jprev] : 0.0f;
        }
        if (kN == 1) {
            load8(Bk, l0, L, vin, c.B);
            load8(Ck, l0, L, vin, c.C);
            if (rev) { reverse8(c.B); reverse8(c.C); }
        }
    };

    auto compute_chunk = [&](int step, BwdChunk<T, kN, kCh>& c) {
        const int j = rev ? step : (nch - 1 - step);
        const int p0 = j * kChunk + lane * kItems;
        const int l0 = rev ? L - 8 - p0 : p0;
        const int jprev = rev ? j + 1 : j - 1;
        const int f4s = swz_f4(p0 >> 2);
        const bool tail = p0 + 8 > L;

        float dt[kCh][8], u[kCh][8], dy[kCh][8], sig[kCh][8], du[kCh][8], ddt[kCh][8];
#pragma unroll
        for (int ch = 0; ch < kCh; ++ch) {
            lds8(xb + ch * Lb, f4s, u[ch]);
            lds8(gb + ch * Lb, f4s, dy[ch]);
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
                const f2 xx = add2(make_float2(c.dt[ch][2 * jj], c.dt[ch][2 * jj + 1]), splat2(bias[ch]));
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
            float Bv[8], Cv[8], dBv[8], dCv[8];
            if (kN == 1) {
#pragma unroll
                for (int i = 0; i < 8; ++i) { Bv[i] = c.B[i]; Cv[i] = c.C[i]; }
            } else {
                load8(Bk + n * L, l0, L, vin, Bv);
                load8(Ck + n * L, l0, L, vin, Cv);
                if (rev) { reverse8(Bv); reverse8(Cv); }
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) { dBv[i] = 0.0f; dCv[i] = 0.0f; }
#pragma unroll
            for (int ch = 0; ch < kCh; ++ch) {
                const float An = (kN == 1) ? A_1[ch] : p.A[kd[ch] * N + n];
                const float A2 = An * kLog2e;
                float a[8], bu[8], S[8], P[8], Sq[8], Pq[8];
                float Pr = 1.0f, Sr = 0.0f;
                const float h_start = (kN == 1) ? c.hstart[ch]
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
                    if (valid[ch]) {
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
            if (rev) { reverse8(dBv); reverse8(dCv); }
            float* dBrow = dBk + n * L;
            float* dCrow = dCk + n * L;
            if (vacc && l0 >= 0 && l0 + 8 <= L) {
                red_add_v4(dBrow + l0, dBv[0], dBv[1], dBv[2], dBv[3]);
                red_add_v4(dBrow + l0 + 4, dBv[4], dBv[5], dBv[6], dBv[7]);
                red_add_v4(dCrow + l0, dCv[0], dCv[1], dCv[2], dCv[3]);
                red_add_v4(dCrow + l0 + 4, dCv[4], dCv[5], dCv[6], dCv[7]);
            } else {
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int l = l0 + i;
                    if (l >= 0 && l < L) {
                        atomicAdd(dBrow + l, dBv[i]);
                        atomicAdd(dCrow + l, dCv[i]);
                    }
                }
            }
        }
        // ddelta (scan order of the route) and du accumulation (position order, pair protocol)
        const bool first_touch = rev ? (j < m) : (j >= m);
        if (!first_touch && !synced) { pair_barrier(k & 1); synced = true; }
#pragma unroll
        for (int ch = 0; ch < kCh; ++ch) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                ddt[ch][i] *= sig[ch][i];
                dbias_acc[ch] += ddt[ch][i];            // dt = 0 beyond L makes these terms exactly 0
            }
            if (valid[ch]) {
                if (rev) reverse8(ddt[ch]);
                store8(ddt_row[ch], (int64_t)l0, (int64_t)L, vout, ddt[ch]);
            }
            if (!first_touch) {
                float o[8];
                lds8(db + ch * Lb, f4s, o);
#pragma unroll
                for (int i = 0; i < 8; ++i) du[ch][i] += o[i];
            }
            sts8(db + ch * Lb, f4s, du[ch]);
        }
    };

    {
        BwdChunk<T, kN, kCh> ca, cb;
        load_chunk(0, ca);
#pragma unroll 1
        for (int step = 0; step < nch; step += 2) {
            if (step + 1 < nch) load_chunk(step + 1, cb);
            compute_chunk(step, ca);
            if (step + 1 < nch) {
                if (step + 2 < nch) load_chunk(step + 2, ca);
                compute_chunk(step + 1, cb);
            }
        }
    }
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

// =========================================================================================================
// host side
// =========================================================================================================
static size_t fwd_smem(int64_t L, int64_t N) {
    return sizeof(float) * (size_t)(4 * kChFwd * buf_len(L) + (N == 1 ? 0 : 4 * kChFwd * kFusedMaxState));
}
static size_t bwd_smem(int64_t L, int64_t N) {
    return sizeof(float) * (size_t)(6 * kChBwd * buf_len(L) + (N == 1 ? 0 : 2 * 4 * kChBwd * kFusedMaxState));
}
constexpr size_t kMaxSmem = 227 * 1024;

int ss2d_supported(int64_t D, int64_t N, int64_t H, int64_t W, int dtype, int backward) {
    (void)D; (void)dtype;
    if (N < 1 || N > kFusedMaxState) return 0;
    const int64_t L = H * W;
    if (L <= 0 || L > (1 << 24)) return 0;
    return (backward ? bwd_smem(L, N) : fwd_smem(L, N)) <= kMaxSmem;
}

template <typename K>
static int set_smem(K kernel, size_t bytes) {
    return (int)cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}

template <typename T, typename TO>
static int launch_fwd_tt(const xfs_ss2d_fwd_args& a, cudaStream_t st) {
    const size_t smem = fwd_smem(a.H * a.W, a.N);
    const unsigned grid = (unsigned)(a.batch * ((a.D + kChFwd - 1) / kChFwd));
    int rc;
    if (a.N == 1) {
        if ((rc = set_smem(ss2d_fwd_kernel<T, TO, 1, kChFwd>, smem))) return rc;
        ss2d_fwd_kernel<T, TO, 1, kChFwd><<<grid, 128, smem, st>>>(a);
    } else {
        if ((rc = set_smem(ss2d_fwd_kernel<T, TO, 0, kChFwd>, smem))) return rc;
        ss2d_fwd_kernel<T, TO, 0, kChFwd><<<grid, 128, smem, st>>>(a);
    }
    return check_launch();
}

int launch_ss2d_fwd(const xfs_ss2d_fwd_args& a, cudaStream_t st) {
    const bool o32 = a.out_dtype == XFS_F32;
    switch (a.dtype) {
        case XFS_F32: return launch_fwd_tt<float, float>(a, st);
        case XFS_BF16: return o32 ? launch_fwd_tt<__nv_bfloat16, float>(a, st) : launch_fwd_tt<__nv_bfloat16, __nv_bfloat16>(a, st);
        default: return o32 ? launch_fwd_tt<__half, float>(a, st) : launch_fwd_tt<__half, __half>(a, st);
    }
}

template <typename T, typename TDO>
static int launch_bwd_tt(const xfs_ss2d_bwd_args& a, cudaStream_t st) {
    const size_t smem = bwd_smem(a.H * a.W, a.N);
    const unsigned grid = (unsigned)(a.batch * ((a.D + kChBwd - 1) / kChBwd));
    int rc;
    if (a.N == 1) {
        if ((rc = set_smem(ss2d_bwd_kernel<T, TDO, 1, kChBwd>, smem))) return rc;
        ss2d_bwd_kernel<T, TDO, 1, kChBwd><<<grid, 128, smem, st>>>(a);
    } else {
        if ((rc = set_smem(ss2d_bwd_kernel<T, TDO, 0, kChBwd>, smem))) return rc;
        ss2d_bwd_kernel<T, TDO, 0, kChBwd><<<grid, 128, smem, st>>>(a);
    }
    return check_launch();
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
