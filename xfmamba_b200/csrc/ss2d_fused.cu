// ss2d_fused.cu -- fused SS2D core for sm_100a: CrossScan gather + S6 selective scan + CrossMerge in ONE kernel
// (forward) and the whole gradient in ONE kernel (backward).
//
// Replaces the operator sequence of SS2Dv2.forward_corev2 (reference models/fusion_vmamba.py:1145,1170-1174):
//     xs = cross_scan_fn(x); ys = selective_scan_fn(xs, dts, As, Bs, Cs, Ds, delta_bias, True); y = cross_merge_fn(ys)
// which moves ~22 elements per (b, d, l) through HBM (xs written 4x and re-read, ys written 4x and re-read).  Here x is
// read once, delta/B/C are streamed once, the merged y is written once: 6 elements per (b, d, l) + B/C.
//
// Work decomposition
//   CTA  = one batch image x kCh (2) adjacent channels; 4 warps = the 4 routes, all running concurrently.
//   The two channels of a CTA share the B/C values of a route (they live in registers once per chunk).
//   Shared memory holds, per channel, the image row in row-major order (xN) and column-major order (xT) and the two
//   accumulators yN (routes 0+2) and yT (routes 1+3), all indexed by POSITION with a 128-byte XOR swizzle so that
//   "lane reads its 8 consecutive floats as two float4" is bank-conflict free.
//   Routes 2/3 walk the same position chunks as routes 0/1 but from the far end (reverse warp scan), so a route and
//   its flip never touch the same accumulator chunk in the same half of the walk: the first half stores, then one
//   64-thread named barrier, then the second half read-modify-writes what the partner stored.  No atomics, and the
//   sum (y0 + y2) + (y1 + y3) is evaluated in the reference's order (models/csm_triton.py:61-62).
//   The S6 recurrence h_l = exp(dt_l A) h_{l-1} + dt_l B_l u_l is scanned per 256-position chunk with a warp-shuffle
//   scan of affine maps (xfscan_common.cuh), state carried in registers (N == 1) or shared memory (N > 1).
#include "xfscan_common.cuh"

namespace xfs {

constexpr int kChFwd = 2;                    // channels per CTA, forward
constexpr int kChBwd = 1;                    // channels per CTA, backward (register budget: see DESIGN.md)
constexpr int kFusedMaxState = 64;           // states carried in smem for N > 1

__host__ __device__ inline int64_t buf_len(int64_t L) { return ((L + kChunk - 1) / kChunk) * kChunk; }

// position -> float offset inside a swizzled buffer (16-byte granules XORed with bits 3..5 of the granule index)
__device__ __forceinline__ int swz_f4(int f) { return f ^ ((f >> 3) & 7); }
__device__ __forceinline__ int swz_pos(int p) { return (swz_f4(p >> 2) << 2) | (p & 3); }

__device__ __forceinline__ void lds8(const float* buf, int f4s, float (&v)[8]) {
    const float4 a = *reinterpret_cast<const float4*>(buf + (f4s << 2));
    const float4 b = *reinterpret_cast<const float4*>(buf + ((f4s ^ 1) << 2));
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void sts8(float* buf, int f4s, const float (&v)[8]) {
    *reinterpret_cast<float4*>(buf + (f4s << 2)) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(buf + ((f4s ^ 1) << 2)) = make_float4(v[4], v[5], v[6], v[7]);
}

__device__ __forceinline__ void pair_barrier(int pair) {   // the 2 warps of a route pair (routes k and k+2)
    asm volatile("bar.sync %0, 64;" ::"r"(pair + 1) : "memory");
}

// Stage one (b, d) image row into the row-major and column-major swizzled buffers; zero the tails.
template <typename T>
__device__ __forceinline__ void stage_image(const T* __restrict__ img, float* __restrict__ bN, float* __restrict__ bT,
                                            int H, int W, int L, int Lb, bool valid, int tid, int nthreads) {
    for (int p = tid; p < Lb; p += nthreads) {
        float v = 0.0f;
        if (valid && p < L) v = Elem<T>::to_f(img[p]);
        bN[swz_pos(p)] = v;
        if (p < L) {
            const int h = p / W, w = p - h * W;
            bT[swz_pos(w * H + h)] = v;
        } else {
            bT[swz_pos(p)] = 0.0f;       // positions L..Lb-1 are the tail of BOTH layouts
        }
    }
}

// =========================================================================================================
// forward
// =========================================================================================================
template <typename T, typename TO, int kN, int kCh>   // kN == 1: single state in registers; kN == 0: runtime N (<= kFusedMaxState)
__global__ void __launch_bounds__(128)
ss2d_fwd_kernel(const xfs_ss2d_fwd_args p) {
    extern __shared__ __align__(16) float smem[];
    const int H = (int)p.H, W = (int)p.W, L = H * W;
    const int Lb = (int)buf_len(L), nch = Lb / kChunk;
    const int64_t D = p.D;
    const int N = (kN == 1) ? 1 : (int)p.N;
    const int64_t pairs = (D + kCh - 1) / kCh;
    const int64_t b = blockIdx.x / pairs;
    const int64_t d0 = (blockIdx.x % pairs) * kCh;
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
        stage_image<T>(x + (b * D + (valid[ch] ? d0 + ch : d0)) * L, xN + ch * Lb, xT + ch * Lb, H, W, L, Lb, valid[ch], tid, 128);
    if (kN == 0)
        for (int i = tid; i < 4 * kCh * kFusedMaxState; i += 128) s_h[i] = 0.0f;
    __syncthreads();

    const float* xb = transposed ? xT : xN;
    float* yb = transposed ? yT : yN;
    const T* __restrict__ delta = reinterpret_cast<const T*>(p.delta);
    const T* __restrict__ Bk = reinterpret_cast<const T*>(p.Bs) + (b * 4 + k) * (int64_t)N * L;
    const T* __restrict__ Ck = reinterpret_cast<const T*>(p.Cs) + (b * 4 + k) * (int64_t)N * L;
    const bool vin = row_vec_ok(delta, L) && row_vec_ok(reinterpret_cast<const T*>(p.Bs), L) &&
                     row_vec_ok(reinterpret_cast<const T*>(p.Cs), L);

    const T* dt_row[kCh];
    float bias[kCh], Dd[kCh], A2_1[kCh], carry1[kCh];
    int64_t kd[kCh];
#pragma unroll
    for (int ch = 0; ch < kCh; ++ch) {
        kd[ch] = k * D + (valid[ch] ? d0 + ch : d0);
        dt_row[ch] = delta + (b * 4 * D + kd[ch]) * L;
        bias[ch] = p.delta_bias ? p.delta_bias[kd[ch]] : 0.0f;
        Dd[ch] = p.Ds ? p.Ds[kd[ch]] : 0.0f;
        A2_1[ch] = (kN == 1) ? p.A[kd[ch]] * kLog2e : 0.0f;
        carry1[ch] = 0.0f;
    }

    const int m = (nch + 1) / 2;        // chunks [0, m) are first touched by the forward route, [m, nch) by its flip
    bool synced = false;
    for (int step = 0; step < nch; ++step) {
        const int j = rev ? (nch - 1 - step) : step;
        const int p0 = j * kChunk + lane * kItems;
        const int64_t l0 = rev ? (int64_t)L - 8 - p0 : (int64_t)p0;     // scan index of the lowest-address element
        const int f4s = swz_f4(p0 >> 2);

        float dt[kCh][8], u[kCh][8], y[kCh][8];
#pragma unroll
        for (int ch = 0; ch < kCh; ++ch) {
            load8(dt_row[ch], l0, L, vin, dt[ch]);
            if (rev) reverse8(dt[ch]);
            lds8(xb + ch * Lb, f4s, u[ch]);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                float xx = dt[ch][i] + bias[ch], e;
                float sp = p.delta_softplus ? softplus_fwd(xx, e) : xx;
                dt[ch][i] = (p0 + i < L) ? sp : 0.0f;
                y[ch][i] = Dd[ch] * u[ch][i];
            }
        }
        for (int n = 0; n < N; ++n) {
            float Bv[8], Cv[8];
            load8(Bk + (int64_t)n * L, l0, L, vin, Bv);
            load8(Ck + (int64_t)n * L, l0, L, vin, Cv);
            if (rev) { reverse8(Bv); reverse8(Cv); }
#pragma unroll
            for (int ch = 0; ch < kCh; ++ch) {
                const float A2 = (kN == 1) ? A2_1[ch] : p.A[kd[ch] * N + n] * kLog2e;
                float S[8], P[8];
                float Pr = 1.0f, Sr = 0.0f;
                if (!rev) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float a = ex2(dt[ch][i] * A2);
                        Sr = fmaf(a, Sr, (dt[ch][i] * Bv[i]) * u[ch][i]);
                        Pr *= a;
                        S[i] = Sr; P[i] = Pr;
                    }
                } else {
#pragma unroll
                    for (int i = 7; i >= 0; --i) {
                        const float a = ex2(dt[ch][i] * A2);
                        Sr = fmaf(a, Sr, (dt[ch][i] * Bv[i]) * u[ch][i]);
                        Pr *= a;
                        S[i] = Sr; P[i] = Pr;
                    }
                }
                float* hs = s_h + (k * kCh + ch) * kFusedMaxState + n;
                const float carry = (kN == 1) ? carry1[ch] : *hs;
                float h_out;
                const float h_in = rev ? warp_prefix<true>(Pr, Sr, carry, lane, h_out)
                                       : warp_prefix<false>(Pr, Sr, carry, lane, h_out);
#pragma unroll
                for (int i = 0; i < 8; ++i) y[ch][i] = fmaf(Cv[i], fmaf(P[i], h_in, S[i]), y[ch][i]);
                if (kN == 1) carry1[ch] = h_out;
                else { __syncwarp(); if (lane == 0) *hs = h_out; }
                if (p.states && lane == 0 && valid[ch])
                    p.states[((b * 4 * D + kd[ch]) * nch + j) * N + n] = h_out;
            }
        }
        // accumulate into the pair's buffer
        const bool first_touch = rev ? (j >= m) : (j < m);
        if (!first_touch && !synced) { pair_barrier(k & 1); synced = true; }
#pragma unroll
        for (int ch = 0; ch < kCh; ++ch) {
            if (!first_touch) {
                float o[8];
                lds8(yb + ch * Lb, f4s, o);
#pragma unroll
                for (int i = 0; i < 8; ++i) y[ch][i] += o[i];
            }
            sts8(yb + ch * Lb, f4s, y[ch]);
        }
    }
    if (!synced) pair_barrier(k & 1);
    __syncthreads();

    // merged output, spatial order: y[p] = yN[p] + yT[w*H + h]
    TO* __restrict__ out = reinterpret_cast<TO*>(p.y);
#pragma unroll
    for (int ch = 0; ch < kCh; ++ch) {
        if (!valid[ch]) continue;
        TO* __restrict__ orow = out + (b * D + d0 + ch) * L;
        for (int pp = tid; pp < L; pp += 128) {
            const int h = pp / W, w = pp - h * W;
            orow[pp] = Elem<TO>::from_f(yN[ch * Lb + swz_pos(pp)] + yT[ch * Lb + swz_pos(w * H + h)]);
        }
    }
}

// =========================================================================================================
// backward
//   smem per channel: xN, xT (u), gN, gT (dy in both layouts), dN, dT (du accumulators, same pair protocol as y)
//   per chunk (walked in the REVERSE of the forward's order) and state n:
//     forward re-scan from the checkpointed state entering the chunk, reverse scan for g = dL/dh, then the
//     closed forms listed in selective_scan.cu.  dBs/dCs: the CTA's channels are summed in registers, then one
//     fp32 atomic per (route, n, l); dA/dDs/dbias: warp-reduced, one atomic per (route, channel).
// =========================================================================================================
template <typename T, typename TDO, int kN, int kCh>
__global__ void __launch_bounds__(128)
ss2d_bwd_kernel(const xfs_ss2d_bwd_args p) {
    extern __shared__ __align__(16) float smem[];
    const int H = (int)p.H, W = (int)p.W, L = H * W;
    const int Lb = (int)buf_len(L), nch = Lb / kChunk;
    const int64_t D = p.D;
    const int N = (kN == 1) ? 1 : (int)p.N;
    const int64_t pairs = (D + kCh - 1) / kCh;
    const int64_t b = blockIdx.x / pairs;
    const int64_t d0 = (blockIdx.x % pairs) * kCh;
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
        const int64_t row = (b * D + (valid[ch] ? d0 + ch : d0)) * L;
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
    const T* __restrict__ Bk = reinterpret_cast<const T*>(p.Bs) + (b * 4 + k) * (int64_t)N * L;
    const T* __restrict__ Ck = reinterpret_cast<const T*>(p.Cs) + (b * 4 + k) * (int64_t)N * L;
    float* __restrict__ dBk = p.dBs + (b * 4 + k) * (int64_t)N * L;
    float* __restrict__ dCk = p.dCs + (b * 4 + k) * (int64_t)N * L;
    const bool vin = row_vec_ok(delta, L) && row_vec_ok(reinterpret_cast<const T*>(p.Bs), L) &&
                     row_vec_ok(reinterpret_cast<const T*>(p.Cs), L);
    const bool vout = row_vec_ok(ddelta, L);

    const T* dt_row[kCh];
    T* ddt_row[kCh];
    float bias[kCh], Dd[kCh], A_1[kCh], qcarry1[kCh], dA1[kCh], dD_acc[kCh], dbias_acc[kCh];
    int64_t kd[kCh];
#pragma unroll
    for (int ch = 0; ch < kCh; ++ch) {
        kd[ch] = k * D + (valid[ch] ? d0 + ch : d0);
        dt_row[ch] = delta + (b * 4 * D + kd[ch]) * L;
        ddt_row[ch] = ddelta + (b * 4 * D + kd[ch]) * L;
        bias[ch] = p.delta_bias ? p.delta_bias[kd[ch]] : 0.0f;
        Dd[ch] = p.Ds ? p.Ds[kd[ch]] : 0.0f;
        A_1[ch] = (kN == 1) ? p.A[kd[ch]] : 0.0f;
        qcarry1[ch] = 0.0f; dA1[ch] = 0.0f; dD_acc[ch] = 0.0f; dbias_acc[ch] = 0.0f;
    }

    // The backward of route k walks its chunks in the reverse of the forward walk: routes 0/1 go nch-1 -> 0 with a
    // reverse (lanes 31->0) adjoint scan, routes 2/3 go 0 -> nch-1 with a lanes 0->31 adjoint scan.
    const int m = nch / 2;             // bwd-forward-walking routes (2/3) first touch [0, m); routes 0/1 first touch [m, nch)
    bool synced = false;
    for (int step = 0; step < nch; ++step) {
        const int j = rev ? step : (nch - 1 - step);
        const int p0 = j * kChunk + lane * kItems;
        const int64_t l0 = rev ? (int64_t)L - 8 - p0 : (int64_t)p0;
        const int f4s = swz_f4(p0 >> 2);

        float dt[kCh][8], u[kCh][8], dy[kCh][8], sig[kCh][8], du[kCh][8], ddt[kCh][8];
#pragma unroll
        for (int ch = 0; ch < kCh; ++ch) {
            load8(dt_row[ch], l0, L, vin, dt[ch]);
            if (rev) reverse8(dt[ch]);
            lds8(xb + ch * Lb, f4s, u[ch]);
            lds8(gb + ch * Lb, f4s, dy[ch]);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float xx = dt[ch][i] + bias[ch];
                float e = 0.0f;
                const float sp = p.delta_softplus ? softplus_fwd(xx, e) : xx;
                sig[ch][i] = p.delta_softplus ? ((xx > 20.0f) ? 1.0f : e * rcp(1.0f + e)) : 1.0f;
                dt[ch][i] = (p0 + i < L) ? sp : 0.0f;
                du[ch][i] = Dd[ch] * dy[ch][i];
                ddt[ch][i] = 0.0f;
                dD_acc[ch] = fmaf(dy[ch][i], u[ch][i], dD_acc[ch]);
            }
        }
        for (int n = 0; n < N; ++n) {
            float Bv[8], Cv[8], dBv[8], dCv[8];
            load8(Bk + (int64_t)n * L, l0, L, vin, Bv);
            load8(Ck + (int64_t)n * L, l0, L, vin, Cv);
            if (rev) { reverse8(Bv); reverse8(Cv); }
#pragma unroll
            for (int i = 0; i < 8; ++i) { dBv[i] = 0.0f; dCv[i] = 0.0f; }
#pragma unroll
            for (int ch = 0; ch < kCh; ++ch) {
                const float An = (kN == 1) ? A_1[ch] : p.A[kd[ch] * N + n];
                const float A2 = An * kLog2e;
                float a[8], bu[8], S[8], P[8], Sq[8], Pq[8];
                float Pr = 1.0f, Sr = 0.0f;
                // state entering this chunk in the forward walk = checkpoint of the chunk walked just before it
                const int jprev = rev ? j + 1 : j - 1;
                const float h_start = (jprev >= 0 && jprev < nch) ? p.states[((b * 4 * D + kd[ch]) * nch + jprev) * N + n] : 0.0f;
                float unused, h_in, q_in, q_out;
                float* qs = s_q + (k * kCh + ch) * kFusedMaxState + n;
                const float qc = (kN == 1) ? qcarry1[ch] : *qs;
                if (!rev) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        a[i] = ex2(dt[ch][i] * A2);
                        bu[i] = (dt[ch][i] * Bv[i]) * u[ch][i];
                        Sr = fmaf(a[i], Sr, bu[i]); Pr *= a[i]; S[i] = Sr; P[i] = Pr;
                    }
                    h_in = warp_prefix<false>(Pr, Sr, h_start, lane, unused);
                    Pr = 1.0f; Sr = 0.0f;
#pragma unroll
                    for (int i = 7; i >= 0; --i) {
                        Sr = a[i] * fmaf(Cv[i], dy[ch][i], Sr); Pr *= a[i]; Sq[i] = Sr; Pq[i] = Pr;
                    }
                    q_in = warp_prefix<true>(Pr, Sr, qc, lane, q_out);
                } else {
#pragma unroll
                    for (int i = 7; i >= 0; --i) {
                        a[i] = ex2(dt[ch][i] * A2);
                        bu[i] = (dt[ch][i] * Bv[i]) * u[ch][i];
                        Sr = fmaf(a[i], Sr, bu[i]); Pr *= a[i]; S[i] = Sr; P[i] = Pr;
                    }
                    h_in = warp_prefix<true>(Pr, Sr, h_start, lane, unused);
                    Pr = 1.0f; Sr = 0.0f;
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        Sr = a[i] * fmaf(Cv[i], dy[ch][i], Sr); Pr *= a[i]; Sq[i] = Sr; Pq[i] = Pr;
                    }
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
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int64_t l = l0 + i;
                if (l >= 0 && l < L) {
                    atomicAdd(dBk + (int64_t)n * L + l, dBv[i]);
                    atomicAdd(dCk + (int64_t)n * L + l, dCv[i]);
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
                if (p0 + i < L) dbias_acc[ch] += ddt[ch][i];
            }
            if (valid[ch]) {
                if (rev) reverse8(ddt[ch]);
                store8(ddt_row[ch], l0, L, vout, ddt[ch]);
            }
            if (!first_touch) {
                float o[8];
                lds8(db + ch * Lb, f4s, o);
#pragma unroll
                for (int i = 0; i < 8; ++i) du[ch][i] += o[i];
            }
            sts8(db + ch * Lb, f4s, du[ch]);
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
    for (int ch = 0; ch < kCh; ++ch) {
        if (!valid[ch]) continue;
        T* __restrict__ orow = dx + (b * D + d0 + ch) * L;
        for (int pp = tid; pp < L; pp += 128) {
            const int h = pp / W, w = pp - h * W;
            orow[pp] = Elem<T>::from_f(dN[ch * Lb + swz_pos(pp)] + dT[ch * Lb + swz_pos(w * H + h)]);
        }
    }
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
