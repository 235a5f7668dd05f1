// selective_scan.cu -- stand-alone S6 selective scan, forward and backward, any dstate / ngroups / seqlen.
//
// Replaces selective_scan_fwd_kernel / selective_scan_bwd_kernel of the reference's native extension
// (models/selective_scan/csrc/selective_scan/selective_scan_fwd_kernel.cuh:62-206, selective_scan_bwd_kernel.cuh:66-274)
// with the semantics of selective_scan_torch (models/csms6s.py:25-68).  Design differences: one WARP (not one
// CTA) owns a (batch, channel) sequence, so tiny-L / huge-batch*dim shapes (XFMamba's stage 3/4 and fusion blocks)
// fill the machine; the scan over L is a chunked warp-shuffle scan of affine maps with the state carried in a
// register; only h (not (a-product, h)) is checkpointed per chunk.
//
// This is the drop-in for `selective_scan_fn`; the SS2D hot path uses the fused kernels in ss2d_fused.cu instead.
#include "xfscan_common.cuh"

namespace xfs {

constexpr int kWarpsPerCta = 4;
constexpr int kMaxState = 256;   // selective_scan.cpp:199

// ---------------------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------------------
template <typename T, typename TO>
__global__ void __launch_bounds__(kWarpsPerCta * 32)
sscan_fwd_kernel(const xfs_scan_fwd_args p) {
    __shared__ float s_h[kWarpsPerCta][kMaxState];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int64_t seq = (int64_t)blockIdx.x * kWarpsPerCta + wib;      // b*dim + d
    if (seq >= p.batch * p.dim) return;
    const int64_t b = seq / p.dim, d = seq % p.dim;
    const int64_t L = p.seqlen, N = p.dstate;
    const int64_t g = d / (p.dim / p.ngroups);
    const int64_t nchunks = (L + kChunk - 1) / kChunk;

    const T* __restrict__ u_row = reinterpret_cast<const T*>(p.u) + seq * L;
    const T* __restrict__ dt_row = reinterpret_cast<const T*>(p.delta) + seq * L;
    const T* __restrict__ Bg = reinterpret_cast<const T*>(p.B) + (b * p.ngroups + g) * N * L;
    const T* __restrict__ Cg = reinterpret_cast<const T*>(p.C) + (b * p.ngroups + g) * N * L;
    TO* __restrict__ o_row = reinterpret_cast<TO*>(p.out) + seq * L;
    const bool vin = row_vec_ok(reinterpret_cast<const T*>(p.u), L) && row_vec_ok(reinterpret_cast<const T*>(p.delta), L) &&
                     row_vec_ok(reinterpret_cast<const T*>(p.B), L) && row_vec_ok(reinterpret_cast<const T*>(p.C), L);
    const bool vout = row_vec_ok(reinterpret_cast<const TO*>(p.out), L);
    const float bias = p.delta_bias ? p.delta_bias[d] : 0.0f;
    const float Dd = p.D ? p.D[d] : 0.0f;
    float* __restrict__ st = p.states ? p.states + seq * nchunks * N : nullptr;

    for (int n = lane; n < N; n += 32) s_h[wib][n] = 0.0f;
    __syncwarp();

    for (int64_t c = 0; c < nchunks; ++c) {
        const int64_t l0 = c * kChunk + lane * kItems;
        float dt[8], u[8], y[8];
        load8(dt_row, l0, L, vin, dt);
        load8(u_row, l0, L, vin, u);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            float x = dt[i] + bias, e;
            dt[i] = p.delta_softplus ? softplus_fwd(x, e) : x;
            if (l0 + i >= L) dt[i] = 0.0f;          // identity map beyond the end of the sequence
            y[i] = Dd * u[i];
        }
        for (int64_t n = 0; n < N; ++n) {
            const float A2 = p.A[d * N + n] * kLog2e;
            float Bv[8], Cv[8], S[8], P[8];
            load8(Bg + n * L, l0, L, vin, Bv);
            load8(Cg + n * L, l0, L, vin, Cv);
            float Pr = 1.0f, Sr = 0.0f;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float a = ex2(dt[i] * A2);
                const float bu = (dt[i] * Bv[i]) * u[i];
                Sr = fmaf(a, Sr, bu);
                Pr *= a;
                S[i] = Sr;
                P[i] = Pr;
            }
            float h_out;
            const float h_in = warp_prefix<false>(Pr, Sr, s_h[wib][n], lane, h_out);
#pragma unroll
            for (int i = 0; i < 8; ++i) y[i] = fmaf(Cv[i], fmaf(P[i], h_in, S[i]), y[i]);
            __syncwarp();
            if (lane == 0) {
                s_h[wib][n] = h_out;
                if (st) st[c * N + n] = h_out;
            }
        }
        __syncwarp();
        store8(o_row, l0, L, vout, y);
    }
}

// ---------------------------------------------------------------------------------------------------------
// backward.  Chunks are walked last -> first.  Per chunk and state n:
//   forward re-scan from the checkpointed state entering the chunk          -> h_i
//   reverse scan of q_i = a_i * (C_i dy_i + q_{i+1})                        -> g_i = C_i dy_i + q_{i+1}  (= dL/dh_i)
//   du += g dt B;  ddt += g (B u + A (h - b));  dA += g dt (h - b);  dB += g dt u;  dC += dy h   (b = dt B u)
// dB / dC are shared by every channel of a group -> fp32 atomics (as the reference does, bwd_kernel.cuh:221-227).
// ---------------------------------------------------------------------------------------------------------
template <typename T, typename TDO>
__global__ void __launch_bounds__(kWarpsPerCta * 32)
sscan_bwd_kernel(const xfs_scan_bwd_args p) {
    __shared__ float s_q[kWarpsPerCta][kMaxState];    // reverse carry per state
    __shared__ float s_dA[kWarpsPerCta][kMaxState];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int64_t work = (int64_t)blockIdx.x * kWarpsPerCta + wib;
    if (work >= p.batch * p.dim) return;
    // batch index fastest: the warps resident at any moment then belong to many batch images, and few of them add into
    // the same dB / dC rows (shared by all channels of one image and group) at the same time
    const int64_t b = work % p.batch, d = work / p.batch;
    const int64_t seq = b * p.dim + d;
    const int64_t L = p.seqlen, N = p.dstate;
    const int64_t g = d / (p.dim / p.ngroups);
    const int64_t nchunks = (L + kChunk - 1) / kChunk;

    const T* __restrict__ u_row = reinterpret_cast<const T*>(p.u) + seq * L;
    const T* __restrict__ dt_row = reinterpret_cast<const T*>(p.delta) + seq * L;
    const TDO* __restrict__ dy_row = reinterpret_cast<const TDO*>(p.dout) + seq * L;
    const T* __restrict__ Bg = reinterpret_cast<const T*>(p.B) + (b * p.ngroups + g) * N * L;
    const T* __restrict__ Cg = reinterpret_cast<const T*>(p.C) + (b * p.ngroups + g) * N * L;
    // every channel of a group adds into the same dB / dC rows: spread them over acc_replicas copies (xfscan.h) and use
    // 16-byte vector reductions where rows allow -- scalar atomics on shared lines were 8x the operations and serialised in L2
    const int64_t rep = p.acc_replicas > 1 ? d % p.acc_replicas : 0;
    float* __restrict__ dBg = p.dB + ((rep * p.batch + b) * p.ngroups + g) * N * L;
    float* __restrict__ dCg = p.dC + ((rep * p.batch + b) * p.ngroups + g) * N * L;
    const bool vacc = (L % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.dB) | reinterpret_cast<uintptr_t>(p.dC)) & 15) == 0;
    T* __restrict__ du_row = reinterpret_cast<T*>(p.du) + seq * L;
    T* __restrict__ ddt_row = reinterpret_cast<T*>(p.ddelta) + seq * L;
    const bool vin = row_vec_ok(reinterpret_cast<const T*>(p.u), L) && row_vec_ok(reinterpret_cast<const T*>(p.delta), L) &&
                     row_vec_ok(reinterpret_cast<const T*>(p.B), L) && row_vec_ok(reinterpret_cast<const T*>(p.C), L);
    const bool vdy = row_vec_ok(reinterpret_cast<const TDO*>(p.dout), L);
    const bool vout = row_vec_ok(reinterpret_cast<const T*>(p.du), L) && row_vec_ok(reinterpret_cast<const T*>(p.ddelta), L);
    const float bias = p.delta_bias ? p.delta_bias[d] : 0.0f;
    const float Dd = p.D ? p.D[d] : 0.0f;
    const float* __restrict__ st = p.states + seq * nchunks * N;

    for (int n = lane; n < N; n += 32) { s_q[wib][n] = 0.0f; s_dA[wib][n] = 0.0f; }
    __syncwarp();
    float dD_acc = 0.0f, dbias_acc = 0.0f;

    for (int64_t c = nchunks - 1; c >= 0; --c) {
        const int64_t l0 = c * kChunk + lane * kItems;
        float dt[8], u[8], dy[8], sig[8], du[8], ddt[8];
        load8(dt_row, l0, L, vin, dt);
        load8(u_row, l0, L, vin, u);
        load8(dy_row, l0, L, vdy, dy);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float x = dt[i] + bias;
            float e = 0.0f;
            dt[i] = p.delta_softplus ? softplus_fwd(x, e) : x;
            sig[i] = p.delta_softplus ? ((x > 20.0f) ? 1.0f : e * rcp(1.0f + e)) : 1.0f;
            if (l0 + i >= L) dt[i] = 0.0f;
            du[i] = Dd * dy[i];
            ddt[i] = 0.0f;
            dD_acc = fmaf(dy[i], u[i], dD_acc);
        }
        for (int64_t n = 0; n < N; ++n) {
            const float An = p.A[d * N + n];
            const float A2 = An * kLog2e;
            float Bv[8], Cv[8], a[8], bu[8], S[8], P[8];
            load8(Bg + n * L, l0, L, vin, Bv);
            load8(Cg + n * L, l0, L, vin, Cv);
            float Pr = 1.0f, Sr = 0.0f;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                a[i] = ex2(dt[i] * A2);
                bu[i] = (dt[i] * Bv[i]) * u[i];
                Sr = fmaf(a[i], Sr, bu[i]);
                Pr *= a[i];
                S[i] = Sr;
                P[i] = Pr;
            }
            float unused;
            const float h_start = (c > 0) ? st[(c - 1) * N + n] : 0.0f;
            const float h_in = warp_prefix<false>(Pr, Sr, h_start, lane, unused);
            // reverse scan of the maps q -> a_i*q + a_i*C_i*dy_i, walking i = 7..0 and lanes 31..0
            float Sq[8], Pq[8];
            Pr = 1.0f;
            Sr = 0.0f;
#pragma unroll
            for (int i = 7; i >= 0; --i) {
                const float cd = Cv[i] * dy[i];
                Sr = a[i] * (Sr + cd);
                Pr *= a[i];
                Sq[i] = Sr;
                Pq[i] = Pr;
            }
            float q_out;
            const float q_in = warp_prefix<true>(Pr, Sr, s_q[wib][n], lane, q_out);
            float dA_part = 0.0f, dBv[8], dCv[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float h = fmaf(P[i], h_in, S[i]);
                const float q_next = (i == 7) ? q_in : fmaf(Pq[i == 7 ? 7 : i + 1], q_in, Sq[i == 7 ? 7 : i + 1]);
                const float gi = fmaf(Cv[i], dy[i], q_next);
                const float hp = h - bu[i];                       // a_i * h_{i-1}
                du[i] = fmaf(gi * dt[i], Bv[i], du[i]);
                ddt[i] = fmaf(gi, fmaf(Bv[i], u[i], An * hp), ddt[i]);
                dA_part = fmaf(gi * dt[i], hp, dA_part);
                dBv[i] = gi * dt[i] * u[i];
                dCv[i] = dy[i] * h;
            }
            if (vacc) {         // L % 4 == 0: a 16-byte granule is entirely inside or outside the row
                if (l0 + 4 <= L) {
                    red_add_v4(dBg + n * L + l0, dBv[0], dBv[1], dBv[2], dBv[3]);
                    red_add_v4(dCg + n * L + l0, dCv[0], dCv[1], dCv[2], dCv[3]);
                }
                if (l0 + 8 <= L) {
                    red_add_v4(dBg + n * L + l0 + 4, dBv[4], dBv[5], dBv[6], dBv[7]);
                    red_add_v4(dCg + n * L + l0 + 4, dCv[4], dCv[5], dCv[6], dCv[7]);
                }
            } else {
#pragma unroll
                for (int i = 0; i < 8; ++i)
                    if (l0 + i < L) {
                        atomicAdd(dBg + n * L + l0 + i, dBv[i]);
                        atomicAdd(dCg + n * L + l0 + i, dCv[i]);
                    }
            }
            dA_part = warp_sum(dA_part);
            __syncwarp();
            if (lane == 0) {
                s_q[wib][n] = q_out;
                s_dA[wib][n] += dA_part;
            }
        }
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            ddt[i] *= sig[i];
            if (l0 + i < L) dbias_acc += ddt[i];
        }
        store8(du_row, l0, L, vout, du);
        store8(ddt_row, l0, L, vout, ddt);
    }
    __syncwarp();
    for (int n = lane; n < N; n += 32) atomicAdd(p.dA + d * N + n, s_dA[wib][n]);
    dD_acc = warp_sum(dD_acc);
    dbias_acc = warp_sum(dbias_acc);
    if (lane == 0) {
        if (p.dD) atomicAdd(p.dD + d, dD_acc);
        if (p.ddelta_bias) atomicAdd(p.ddelta_bias + d, dbias_acc);
    }
}

// ---------------------------------------------------------------------------------------------------------
// backward, N == 1, fp32 rows, L % 8 == 0, 32-byte aligned tensors: what `selective_scan_fn` sees in every VMamba-style block
// (d_state = 1; reference models/csms6s.py:71-126 -> selective_scan_bwd_kernel.cuh).  Same arithmetic as ss2d_lane_bwd.cu:
// 256-bit loads of the NEXT chunk's five rows (u, delta, dout, B, C) issued before the rare softplus-repair branch, packed
// FFMA2 / FMUL2 arithmetic, the adjoint folded as G_i = cd_i + a_{i+1} G_{i+1} with suffix products, ONE interleaved pair of
// predicate-out warp scans (forward fold from the chunk checkpoint, adjoint against it), everything that does not depend on
// the adjoint computed in front of the scan.  The general kernel above loads at the top of every chunk (latency exposed)
// and computes in scalar fp32.
// ---------------------------------------------------------------------------------------------------------
namespace {
__device__ __forceinline__ void ldg256p(const float* p, f2 (&o)[4]) {
    asm volatile("ld.global.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=f"(o[0].x), "=f"(o[0].y), "=f"(o[1].x), "=f"(o[1].y), "=f"(o[2].x), "=f"(o[2].y), "=f"(o[3].x), "=f"(o[3].y) : "l"(p) : "memory");
}
__device__ __forceinline__ void stg256p(float* p, const f2 (&v)[4]) {
    asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "f"(v[0].x), "f"(v[0].y), "f"(v[1].x), "f"(v[1].y),
                 "f"(v[2].x), "f"(v[2].y), "f"(v[3].x), "f"(v[3].y) : "memory");
}
__device__ __forceinline__ float& el8(f2 (&v)[4], int i) { return (i & 1) ? v[i >> 1].y : v[i >> 1].x; }
struct N1Chunk { f2 u[4], dt[4], dy[4], B[4], C[4]; float hstart; };
}  // namespace

template <bool kSoftplus>
__global__ void __launch_bounds__(kWarpsPerCta * 32)
sscan_n1_bwd_kernel(const xfs_scan_bwd_args p) {
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int64_t work = (int64_t)blockIdx.x * kWarpsPerCta + wib;
    if (work >= p.batch * p.dim) return;                       // whole warps only
    const int64_t b = work % p.batch, d = work / p.batch;        // batch index fastest (see sscan_bwd_kernel)
    const int64_t seq = b * p.dim + d;
    const int L = (int)p.seqlen;
    const int64_t g = d / (p.dim / p.ngroups);
    const int nch = (L + kChunk - 1) / kChunk;
    const float* __restrict__ u_row = reinterpret_cast<const float*>(p.u) + seq * L;
    const float* __restrict__ dt_row = reinterpret_cast<const float*>(p.delta) + seq * L;
    const float* __restrict__ dy_row = reinterpret_cast<const float*>(p.dout) + seq * L;
    const float* __restrict__ B_row = reinterpret_cast<const float*>(p.B) + (b * p.ngroups + g) * L;
    const float* __restrict__ C_row = reinterpret_cast<const float*>(p.C) + (b * p.ngroups + g) * L;
    const int64_t rep = p.acc_replicas > 1 ? d % p.acc_replicas : 0;
    float* __restrict__ dB_row = p.dB + ((rep * p.batch + b) * p.ngroups + g) * L;
    float* __restrict__ dC_row = p.dC + ((rep * p.batch + b) * p.ngroups + g) * L;
    float* __restrict__ du_row = reinterpret_cast<float*>(p.du) + seq * L;
    float* __restrict__ ddt_row = reinterpret_cast<float*>(p.ddelta) + seq * L;
    const float* __restrict__ st = p.states + seq * nch;
    asm volatile("" : "+l"(u_row), "+l"(dt_row), "+l"(dy_row), "+l"(B_row), "+l"(C_row));
    asm volatile("" : "+l"(dB_row), "+l"(dC_row), "+l"(du_row), "+l"(ddt_row), "+l"(st));
    const float bias = p.delta_bias ? p.delta_bias[d] : 0.0f;
    const float Dd = p.D ? p.D[d] : 0.0f;
    const float An = p.A[d], A2 = An * kLog2e;
    const unsigned off_max = (unsigned)(L - 8);

    int off = (nch - 1) * kChunk + 8 * lane;                     // this lane's 8 positions of the chunk being loaded
    int cidx = nch - 1;
    auto load = [&](N1Chunk& c) __attribute__((always_inline)) {
        const unsigned o = min((unsigned)off, off_max);            // beyond the row (last chunk) / before it (the re-load after chunk 0): clamped
        ldg256p(dt_row + o, c.dt); ldg256p(u_row + o, c.u); ldg256p(dy_row + o, c.dy);
        ldg256p(B_row + o, c.B); ldg256p(C_row + o, c.C);
        float hs = 0.0f;
        if (cidx > 0) asm volatile("ld.global.f32 %0, [%1];" : "=f"(hs) : "l"(st + (cidx - 1)) : "memory");
        c.hstart = hs;
        off -= kChunk; --cidx;
    };
    N1Chunk c;
    load(c);
    f2 dD2 = splat2(0.0f), dbias2 = splat2(0.0f), dA2 = splat2(0.0f);
    float qcarry = 0.0f;

#pragma unroll 1
    for (int ci = nch - 1; ci >= 0; --ci) {
        const int o = off + kChunk;                              // this chunk's offset
        const bool ok = o + 8 <= L;                              // L % 8 == 0: a lane's sector is inside or outside the row
        f2 u[4], dy[4], Bv[4], xl[4], cd[4], Bu[4], dt[4], sig[4], dtB[4], e2[4];
        const float hstart = c.hstart;
        bool odd = false;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            u[i] = ok ? c.u[i] : splat2(0.0f);
            dy[i] = ok ? c.dy[i] : splat2(0.0f);
            Bv[i] = c.B[i];
            xl[i] = kSoftplus ? fma2(c.dt[i], splat2(kLog2e), splat2(bias * kLog2e)) : add2(c.dt[i], splat2(bias));
            cd[i] = mul2(c.C[i], dy[i]);
            Bu[i] = mul2(Bv[i], u[i]);
        }
        if (kSoftplus) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                e2[i] = ex2_2(xl[i]);
                const f2 w = add2(e2[i], splat2(1.0f));
                dt[i] = mul2(make_float2(lg2(w.x), lg2(w.y)), splat2(kLn2));
                sig[i] = mul2(e2[i], make_float2(rcp(w.x), rcp(w.y)));
            }
            const float emin = fminf(fminf(fminf(e2[0].x, e2[0].y), fminf(e2[1].x, e2[1].y)), fminf(fminf(e2[2].x, e2[2].y), fminf(e2[3].x, e2[3].y)));
            const float emax = fmaxf(fmaxf(fmaxf(e2[0].x, e2[0].y), fmaxf(e2[1].x, e2[1].y)), fmaxf(fmaxf(e2[2].x, e2[2].y), fmaxf(e2[3].x, e2[3].y)));
            odd = !(emin >= 0.015625f && emax <= 268435456.0f);
        } else {
#pragma unroll
            for (int i = 0; i < 4; ++i) { dt[i] = xl[i]; sig[i] = splat2(1.0f); }
        }
        if (!ok) { dt[0] = dt[1] = dt[2] = dt[3] = splat2(0.0f); }
#pragma unroll
        for (int i = 0; i < 4; ++i) dtB[i] = mul2(dt[i], Bv[i]);
        load(c);                                                   // every streamed register has been read: next chunk's loads, pinned here
        __syncwarp();
        if (kSoftplus && __any_sync(kFull, odd)) {                 // rare: small-argument series / identity above 20; B re-read from L2
            f2 Br[4];
            ldg256p(B_row + min((unsigned)o, off_max), Br);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const f2 x = mul2(xl[i], splat2(kLn2)), e = e2[i];
                f2 ser = fma2(e, splat2(-0.25f), splat2(0.33333334f));
                ser = fma2(ser, e, splat2(-0.5f));
                ser = fma2(ser, e, splat2(1.0f));
                ser = mul2(ser, e);
                f2 q;
                q.x = (e.x < 0.015625f) ? ser.x : dt[i].x;
                q.y = (e.y < 0.015625f) ? ser.y : dt[i].y;
                dt[i].x = (x.x > 20.0f) ? x.x : q.x;
                dt[i].y = (x.y > 20.0f) ? x.y : q.y;
                sig[i].x = (x.x > 20.0f) ? 1.0f : sig[i].x;
                sig[i].y = (x.y > 20.0f) ? 1.0f : sig[i].y;
                if (!ok) dt[i] = splat2(0.0f);
                dtB[i] = mul2(dt[i], Br[i]);
            }
        }
        // forward fold (ascending) and adjoint fold (descending) of the lane's 8 positions
        f2 a[4], bu[4], S[4], P[4], G[4], Pq[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            a[i] = ex2_2(mul2(dt[i], splat2(A2)));
            bu[i] = mul2(dtB[i], u[i]);
        }
        float Sr = 0.0f, Pr = 1.0f;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            Sr = fmaf(el8(a, i), Sr, el8(bu, i));
            Pr = (i == 0) ? el8(a, 0) : Pr * el8(a, i);
            el8(S, i) = Sr; el8(P, i) = Pr;
        }
        float Gp = el8(cd, 7), Pp = 1.0f;
        el8(G, 7) = Gp; el8(Pq, 7) = 1.0f;
#pragma unroll
        for (int i = 6; i >= 0; --i) {
            const float an = el8(a, i + 1);
            Gp = fmaf(an, Gp, el8(cd, i));
            Pp = (i == 6) ? an : Pp * an;
            el8(G, i) = Gp; el8(Pq, i) = Pp;
        }
        // one interleaved pair of warp scans: forward (lanes ascending, from the chunk checkpoint), adjoint (descending, from qcarry)
        float Pf = Pr, Sf = Sr, Pa = el8(a, 0) * Pp, Sa = el8(a, 0) * Gp;
#pragma unroll
        for (int so = 1; so < 32; so <<= 1) {
            asm volatile("{\n\t.reg .pred p;\n\t.reg .f32 pn, sn;\n\t"
                         "shfl.sync.up.b32 pn|p, %0, %2, 0, 0xffffffff;\n\t"
                         "shfl.sync.up.b32 sn, %1, %2, 0, 0xffffffff;\n\t"
                         "@p fma.rn.ftz.f32 %1, %0, sn, %1;\n\t"
                         "@p mul.ftz.f32 %0, %0, pn;\n\t}" : "+f"(Pf), "+f"(Sf) : "r"(so));
            asm volatile("{\n\t.reg .pred p;\n\t.reg .f32 pn, sn;\n\t"
                         "shfl.sync.down.b32 pn|p, %0, %2, 0x1f, 0xffffffff;\n\t"
                         "shfl.sync.down.b32 sn, %1, %2, 0x1f, 0xffffffff;\n\t"
                         "@p fma.rn.ftz.f32 %1, %0, sn, %1;\n\t"
                         "@p mul.ftz.f32 %0, %0, pn;\n\t}" : "+f"(Pa), "+f"(Sa) : "r"(so));
        }
        const float incl_f = fmaf(Pf, hstart, Sf), incl_a = fmaf(Pa, qcarry, Sa);
        float h_in = hstart, r_in = qcarry;
        asm volatile("{\n\t.reg .pred p;\n\t.reg .f32 t;\n\tshfl.sync.up.b32 t|p, %1, 1, 0, 0xffffffff;\n\t@p mov.f32 %0, t;\n\t}" : "+f"(h_in) : "f"(incl_f));
        asm volatile("{\n\t.reg .pred p;\n\t.reg .f32 t;\n\tshfl.sync.down.b32 t|p, %1, 1, 0x1f, 0xffffffff;\n\t@p mov.f32 %0, t;\n\t}" : "+f"(r_in) : "f"(incl_a));
        qcarry = __shfl_sync(kFull, incl_a, 0);

        f2 du[4], dd[4], dBv[4], dCv[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const f2 h = fma2(P[i], splat2(h_in), S[i]);
            const f2 hp = fma2(bu[i], splat2(-1.0f), h);               // a_i h_prev
            const f2 gg = fma2(Pq[i], splat2(r_in), G[i]);
            const f2 gdt = mul2(gg, dt[i]);
            du[i] = fma2(gg, dtB[i], mul2(splat2(Dd), dy[i]));
            dd[i] = mul2(mul2(gg, fma2(splat2(An), hp, Bu[i])), sig[i]);
            dA2 = fma2(gdt, hp, dA2);
            dBv[i] = mul2(gdt, u[i]);
            dCv[i] = mul2(dy[i], h);
            dD2 = fma2(dy[i], u[i], dD2);
            dbias2 = add2(dbias2, dd[i]);
        }
        if (ok) {
            stg256p(du_row + o, du);
            stg256p(ddt_row + o, dd);
            red_add_v4(dB_row + o, dBv[0].x, dBv[0].y, dBv[1].x, dBv[1].y);
            red_add_v4(dB_row + o + 4, dBv[2].x, dBv[2].y, dBv[3].x, dBv[3].y);
            red_add_v4(dC_row + o, dCv[0].x, dCv[0].y, dCv[1].x, dCv[1].y);
            red_add_v4(dC_row + o + 4, dCv[2].x, dCv[2].y, dCv[3].x, dCv[3].y);
        }
    }
    const float vA = warp_sum(dA2.x + dA2.y), vD = warp_sum(dD2.x + dD2.y), vb = warp_sum(dbias2.x + dbias2.y);
    if (lane == 0) {
        atomicAdd(p.dA + d, vA);
        if (p.dD) atomicAdd(p.dD + d, vD);
        if (p.ddelta_bias) atomicAdd(p.ddelta_bias + d, vb);
    }
}

// forward twin of sscan_n1_bwd_kernel (d_state = 1, fp32 rows and output, L % 8 == 0, 32-byte aligned): the next chunk's four
// rows in flight (256-bit loads) while this one is computed, packed arithmetic, one predicate-out warp scan
template <bool kSoftplus>
__global__ void __launch_bounds__(kWarpsPerCta * 32)
sscan_n1_fwd_kernel(const xfs_scan_fwd_args p) {
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int64_t seq = (int64_t)blockIdx.x * kWarpsPerCta + wib;      // b*dim + d
    if (seq >= p.batch * p.dim) return;
    const int64_t b = seq / p.dim, d = seq % p.dim;
    const int L = (int)p.seqlen;
    const int64_t g = d / (p.dim / p.ngroups);
    const int nch = (L + kChunk - 1) / kChunk;
    const float* __restrict__ u_row = reinterpret_cast<const float*>(p.u) + seq * L;
    const float* __restrict__ dt_row = reinterpret_cast<const float*>(p.delta) + seq * L;
    const float* __restrict__ B_row = reinterpret_cast<const float*>(p.B) + (b * p.ngroups + g) * L;
    const float* __restrict__ C_row = reinterpret_cast<const float*>(p.C) + (b * p.ngroups + g) * L;
    float* __restrict__ o_row = reinterpret_cast<float*>(p.out) + seq * L;
    float* __restrict__ st = p.states ? p.states + seq * nch : nullptr;
    asm volatile("" : "+l"(u_row), "+l"(dt_row), "+l"(B_row), "+l"(C_row), "+l"(o_row));
    const float bias = p.delta_bias ? p.delta_bias[d] : 0.0f;
    const float Dd = p.D ? p.D[d] : 0.0f;
    const float A2 = p.A[d] * kLog2e;
    const unsigned off_max = (unsigned)(L - 8);
    int off = 8 * lane;
    f2 ldt[4], lu[4], lB[4], lC[4];
    auto load = [&]() __attribute__((always_inline)) {
        const unsigned o = min((unsigned)off, off_max);
        ldg256p(dt_row + o, ldt); ldg256p(u_row + o, lu); ldg256p(B_row + o, lB); ldg256p(C_row + o, lC);
        off += kChunk;
    };
    load();
    float carry = 0.0f;
#pragma unroll 1
    for (int c = 0; c < nch; ++c) {
        const int o = off - kChunk;
        const bool ok = o + 8 <= L;
        f2 u[4], Cv[4], Bu[4], dt[4], e2[4], xr[4];
        bool odd = false;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            u[i] = lu[i]; Cv[i] = lC[i]; xr[i] = ldt[i];
            Bu[i] = mul2(lB[i], u[i]);
        }
        load();                                  // next chunk's rows (clamped past the end), in front of the rare branch
        __syncwarp();
        if (kSoftplus) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                e2[i] = ex2_2(fma2(xr[i], splat2(kLog2e), splat2(bias * kLog2e)));
                const f2 w = add2(e2[i], splat2(1.0f));
                dt[i] = mul2(make_float2(lg2(w.x), lg2(w.y)), splat2(kLn2));
            }
            const float emin = fminf(fminf(fminf(e2[0].x, e2[0].y), fminf(e2[1].x, e2[1].y)), fminf(fminf(e2[2].x, e2[2].y), fminf(e2[3].x, e2[3].y)));
            const float emax = fmaxf(fmaxf(fmaxf(e2[0].x, e2[0].y), fmaxf(e2[1].x, e2[1].y)), fmaxf(fmaxf(e2[2].x, e2[2].y), fmaxf(e2[3].x, e2[3].y)));
            odd = !(emin >= 0.015625f && emax <= 268435456.0f);
            if (__any_sync(kFull, odd)) {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const f2 x = add2(xr[i], splat2(bias)), e = e2[i];
                    f2 ser = fma2(e, splat2(-0.25f), splat2(0.33333334f));
                    ser = fma2(ser, e, splat2(-0.5f));
                    ser = fma2(ser, e, splat2(1.0f));
                    ser = mul2(ser, e);
                    f2 q;
                    q.x = (e.x < 0.015625f) ? ser.x : dt[i].x;
                    q.y = (e.y < 0.015625f) ? ser.y : dt[i].y;
                    dt[i].x = (x.x > 20.0f) ? x.x : q.x;
                    dt[i].y = (x.y > 20.0f) ? x.y : q.y;
                }
            }
        } else {
#pragma unroll
            for (int i = 0; i < 4; ++i) dt[i] = add2(xr[i], splat2(bias));
        }
        if (!ok) { dt[0] = dt[1] = dt[2] = dt[3] = splat2(0.0f); }      // identity maps beyond the end of the sequence
        f2 a[4], bu[4], S[4], P[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            a[i] = ex2_2(mul2(dt[i], splat2(A2)));
            bu[i] = mul2(dt[i], Bu[i]);
        }
        float Sr = 0.0f, Pr = 1.0f;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            Sr = fmaf(el8(a, i), Sr, el8(bu, i));
            Pr = (i == 0) ? el8(a, 0) : Pr * el8(a, i);
            el8(S, i) = Sr; el8(P, i) = Pr;
        }
        float h_out;
        const float h_in = warp_prefix_p<false>(Pr, Sr, carry, lane, h_out);
        carry = h_out;
        if (st && lane == 0) st[c] = h_out;
        f2 y[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) y[i] = fma2(Cv[i], fma2(P[i], splat2(h_in), S[i]), mul2(splat2(Dd), u[i]));
        if (ok) stg256p(o_row + o, y);
    }
}

static bool n1_fwd_ok(const xfs_scan_fwd_args& a) {
    auto al = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 31) == 0; };
    return a.dtype == XFS_F32 && a.out_dtype == XFS_F32 && a.dstate == 1 && a.seqlen % 8 == 0 && a.seqlen <= (1 << 24) && al(a.u) &&
           al(a.delta) && al(a.B) && al(a.C) && al(a.out);
}

static bool n1_bwd_ok(const xfs_scan_bwd_args& a) {
    auto al = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 31) == 0; };
    return a.dtype == XFS_F32 && a.dout_dtype == XFS_F32 && a.dstate == 1 && a.seqlen % 8 == 0 && a.seqlen <= (1 << 24) && al(a.u) &&
           al(a.delta) && al(a.dout) && al(a.B) && al(a.C) && al(a.du) && al(a.ddelta) && al(a.dB) && al(a.dC);
}

// ---------------------------------------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------------------------------------
template <typename T>
static int launch_fwd_t(const xfs_scan_fwd_args& a, cudaStream_t st) {
    const int64_t nseq = a.batch * a.dim;
    const unsigned grid = (unsigned)((nseq + kWarpsPerCta - 1) / kWarpsPerCta);
    if (a.out_dtype == XFS_F32)
        sscan_fwd_kernel<T, float><<<grid, kWarpsPerCta * 32, 0, st>>>(a);
    else
        sscan_fwd_kernel<T, T><<<grid, kWarpsPerCta * 32, 0, st>>>(a);
    return check_launch();
}

int launch_scan_small_fwd(const xfs_scan_fwd_args& a, cudaStream_t st);
int launch_scan_small_bwd(const xfs_scan_bwd_args& a, cudaStream_t st);
bool scan_small_supported(int64_t L, int64_t N);

int launch_scan_fwd(const xfs_scan_fwd_args& a, cudaStream_t st) {
    if (scan_small_supported(a.seqlen, a.dstate)) return launch_scan_small_fwd(a, st);
    if (n1_fwd_ok(a)) {
        const unsigned grid = (unsigned)((a.batch * a.dim + kWarpsPerCta - 1) / kWarpsPerCta);
        if (a.delta_softplus) sscan_n1_fwd_kernel<true><<<grid, kWarpsPerCta * 32, 0, st>>>(a);
        else sscan_n1_fwd_kernel<false><<<grid, kWarpsPerCta * 32, 0, st>>>(a);
        return check_launch();
    }
    switch (a.dtype) {
        case XFS_F32: return launch_fwd_t<float>(a, st);
        case XFS_BF16: return launch_fwd_t<__nv_bfloat16>(a, st);
        default: return launch_fwd_t<__half>(a, st);
    }
}

template <typename T>
static int launch_bwd_t(const xfs_scan_bwd_args& a, cudaStream_t st) {
    const int64_t nseq = a.batch * a.dim;
    const unsigned grid = (unsigned)((nseq + kWarpsPerCta - 1) / kWarpsPerCta);
    if (a.dout_dtype == XFS_F32)
        sscan_bwd_kernel<T, float><<<grid, kWarpsPerCta * 32, 0, st>>>(a);
    else
        sscan_bwd_kernel<T, T><<<grid, kWarpsPerCta * 32, 0, st>>>(a);
    return check_launch();
}

int launch_scan_bwd(const xfs_scan_bwd_args& a, cudaStream_t st) {
    if (scan_small_supported(a.seqlen, a.dstate)) return launch_scan_small_bwd(a, st);
    if (n1_bwd_ok(a)) {
        const unsigned grid = (unsigned)((a.batch * a.dim + kWarpsPerCta - 1) / kWarpsPerCta);
        if (a.delta_softplus) sscan_n1_bwd_kernel<true><<<grid, kWarpsPerCta * 32, 0, st>>>(a);
        else sscan_n1_bwd_kernel<false><<<grid, kWarpsPerCta * 32, 0, st>>>(a);
        return check_launch();
    }
    switch (a.dtype) {
        case XFS_F32: return launch_bwd_t<float>(a, st);
        case XFS_BF16: return launch_bwd_t<__nv_bfloat16>(a, st);
        default: return launch_bwd_t<__half>(a, st);
    }
}

}  // namespace xfs

// =========================================================================================================
// short sequences (L <= 64, N <= 16): XFMamba's shallow-fusion scan (K = 2 groups, N = 16, L = 49, dim = 2 x 1536/2048;
// reference models/fusion_vmamba.py:831-833).  One sequence per 8-lane group, 16 rows per CTA step; in the backward a
// CTA walks up to 128 rows of ONE (batch, group) and sums their dB/dC in shared memory before a single atomic per
// (n, l) leaves the CTA -- the general kernel above (and the reference) pay one atomic per row.
// =========================================================================================================
#include "ss2d_fused.cuh"

namespace xfs {

constexpr int kRowsPerStep = 16;     // 4 warps x 4 lane groups
constexpr int kRowSteps = 8;         // backward: steps per CTA -> 128 rows

template <typename T, typename TO>
__global__ void __launch_bounds__(128)
sscan_small_fwd_kernel(const xfs_scan_fwd_args p) {
    const int lane = threadIdx.x & 31, g = lane >> 3, j = lane & 7;
    const int64_t row = (int64_t)blockIdx.x * kRowsPerStep + (threadIdx.x >> 5) * 4 + g;      // b*dim + d
    const bool valid = row < p.batch * p.dim;
    const int64_t r = valid ? row : 0;
    const int L = (int)p.seqlen, N = (int)p.dstate;
    const int64_t b = r / p.dim, d = r % p.dim, grp = d / (p.dim / p.ngroups);
    const T* __restrict__ Bg = reinterpret_cast<const T*>(p.B) + (b * p.ngroups + grp) * N * L;
    const T* __restrict__ Cg = reinterpret_cast<const T*>(p.C) + (b * p.ngroups + grp) * N * L;
    const bool vin = row_vec_ok(reinterpret_cast<const T*>(p.u), L) && row_vec_ok(reinterpret_cast<const T*>(p.delta), L) &&
                     row_vec_ok(reinterpret_cast<const T*>(p.B), L) && row_vec_ok(reinterpret_cast<const T*>(p.C), L);
    const bool vout = row_vec_ok(reinterpret_cast<const TO*>(p.out), L);
    const float bias = p.delta_bias ? p.delta_bias[d] : 0.0f;
    const float Dd = p.D ? p.D[d] : 0.0f;
    const int l0 = j * 8;
    float dt[8], u[8], y[8];
    load8<T, true>(reinterpret_cast<const T*>(p.delta) + r * L, l0, L, vin, dt);
    load8<T, true>(reinterpret_cast<const T*>(p.u) + r * L, l0, L, vin, u);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const float xx = dt[i] + bias;
        float e;
        const float sp = p.delta_softplus ? softplus_fwd(xx, e) : xx;
        const bool in = l0 + i < L;
        dt[i] = in ? sp : 0.0f;
        u[i] = in ? u[i] : 0.0f;
        y[i] = Dd * u[i];
    }
    for (int n = 0; n < N; ++n) {
        float Bv[8], Cv[8], S[8], P[8];
        load8<T, true>(Bg + n * L, l0, L, vin, Bv);
        load8<T, true>(Cg + n * L, l0, L, vin, Cv);
        const float A2 = p.A[d * N + n] * kLog2e;
        float Pr = 1.0f, Sr = 0.0f;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float a = ex2(dt[i] * A2);
            Sr = fmaf(a, Sr, (dt[i] * Bv[i]) * u[i]);
            Pr *= a;
            S[i] = Sr; P[i] = Pr;
        }
        float h_end;
        const float h_in = group_prefix<false>(Pr, Sr, j, h_end);
#pragma unroll
        for (int i = 0; i < 8; ++i) y[i] = fmaf(Cv[i], fmaf(P[i], h_in, S[i]), y[i]);
        if (p.states && valid && j == 0) p.states[r * N + n] = h_end;
    }
    if (valid) store8<TO>(reinterpret_cast<TO*>(p.out) + r * L, l0, L, vout, y);
}

template <typename T, typename TDO>
__global__ void __launch_bounds__(128)
sscan_small_bwd_kernel(const xfs_scan_bwd_args p) {
    extern __shared__ __align__(16) float sm[];
    const int L = (int)p.seqlen, N = (int)p.dstate;
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 3, j = lane & 7;
    float* mydB = sm + w * N * kSmallL;                 // [4 warps][N][64]
    float* mydC = sm + (4 + w) * N * kSmallL;
    for (int i = threadIdx.x; i < 8 * N * kSmallL; i += 128) sm[i] = 0.0f;
    __syncthreads();

    const int64_t Dg = p.dim / p.ngroups;
    const int64_t nblk = (Dg + kRowsPerStep * kRowSteps - 1) / (kRowsPerStep * kRowSteps);
    const int64_t bg = blockIdx.x / nblk;               // b*ngroups + grp
    const int64_t blk = blockIdx.x - bg * nblk;
    const int64_t b = bg / p.ngroups, grp = bg % p.ngroups;
    const T* __restrict__ Bg = reinterpret_cast<const T*>(p.B) + bg * N * L;
    const T* __restrict__ Cg = reinterpret_cast<const T*>(p.C) + bg * N * L;
    const bool vin = row_vec_ok(reinterpret_cast<const T*>(p.u), L) && row_vec_ok(reinterpret_cast<const T*>(p.delta), L) &&
                     row_vec_ok(reinterpret_cast<const T*>(p.B), L) && row_vec_ok(reinterpret_cast<const T*>(p.C), L);
    const bool vdy = row_vec_ok(reinterpret_cast<const TDO*>(p.dout), L);
    const bool vout = row_vec_ok(reinterpret_cast<const T*>(p.du), L) && row_vec_ok(reinterpret_cast<const T*>(p.ddelta), L);
    const int l0 = j * 8;

    for (int it = 0; it < kRowSteps; ++it) {
        const int64_t dg = blk * kRowsPerStep * kRowSteps + it * kRowsPerStep + w * 4 + g;      // channel inside the group
        const bool valid = dg < Dg;
        const int64_t d = grp * Dg + (valid ? dg : 0);
        const int64_t r = b * p.dim + d;
        const float bias = p.delta_bias ? p.delta_bias[d] : 0.0f;
        const float Dd = p.D ? p.D[d] : 0.0f;
        float dt[8], u[8], dy[8], sig[8], du[8], ddt[8];
        load8<T, true>(reinterpret_cast<const T*>(p.delta) + r * L, l0, L, vin, dt);
        load8<T, true>(reinterpret_cast<const T*>(p.u) + r * L, l0, L, vin, u);
        load8<TDO, true>(reinterpret_cast<const TDO*>(p.dout) + r * L, l0, L, vdy, dy);
        float dD_acc = 0.0f;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float xx = dt[i] + bias;
            float e = 0.0f;
            const float sp = p.delta_softplus ? softplus_fwd(xx, e) : xx;
            sig[i] = p.delta_softplus ? ((xx > 20.0f) ? 1.0f : e * rcp(1.0f + e)) : 1.0f;
            const bool in = (l0 + i < L) && valid;
            dt[i] = in ? sp : 0.0f;
            u[i] = in ? u[i] : 0.0f;
            dy[i] = in ? dy[i] : 0.0f;
            du[i] = Dd * dy[i];
            ddt[i] = 0.0f;
            dD_acc = fmaf(dy[i], u[i], dD_acc);
        }
        for (int n = 0; n < N; ++n) {
            float Bv[8], Cv[8], a[8], bu[8], S[8], P[8], Sq[8], Pq[8];
            load8<T, true>(Bg + n * L, l0, L, vin, Bv);
            load8<T, true>(Cg + n * L, l0, L, vin, Cv);
            const float An = p.A[d * N + n];
            const float A2 = An * kLog2e;
            float Pr = 1.0f, Sr = 0.0f;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                a[i] = ex2(dt[i] * A2);
                bu[i] = (dt[i] * Bv[i]) * u[i];
                Sr = fmaf(a[i], Sr, bu[i]); Pr *= a[i]; S[i] = Sr; P[i] = Pr;
            }
            float unused;
            const float h_in = group_prefix<false>(Pr, Sr, j, unused);
            Pr = 1.0f; Sr = 0.0f;
#pragma unroll
            for (int i = 7; i >= 0; --i) { Sr = a[i] * fmaf(Cv[i], dy[i], Sr); Pr *= a[i]; Sq[i] = Sr; Pq[i] = Pr; }
            const float q_in = group_prefix<true>(Pr, Sr, j, unused);
            float dA_part = 0.0f, dBv[8], dCv[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float h = fmaf(P[i], h_in, S[i]);
                const float q_next = (i == 7) ? q_in : fmaf(Pq[i == 7 ? 7 : i + 1], q_in, Sq[i == 7 ? 7 : i + 1]);
                const float gi = fmaf(Cv[i], dy[i], q_next);
                const float hp = h - bu[i];
                const float gdt = gi * dt[i];
                du[i] = fmaf(gdt, Bv[i], du[i]);
                ddt[i] = fmaf(gi, fmaf(Bv[i], u[i], An * hp), ddt[i]);
                dA_part = fmaf(gdt, hp, dA_part);
                dBv[i] = quad_sum(gdt * u[i]);
                dCv[i] = quad_sum(dy[i] * h);
            }
            if (g == 0) {
                float* rb = mydB + n * kSmallL + l0;
                float* rc = mydC + n * kSmallL + l0;
#pragma unroll
                for (int i = 0; i < 8; ++i) { rb[i] += dBv[i]; rc[i] += dCv[i]; }
            }
            dA_part = group_sum(dA_part);
            if (valid && j == 0) atomicAdd(p.dA + d * N + n, dA_part);
        }
        float dbias_acc = 0.0f;
#pragma unroll
        for (int i = 0; i < 8; ++i) { ddt[i] *= sig[i]; dbias_acc += ddt[i]; }
        dD_acc = group_sum(dD_acc);
        dbias_acc = group_sum(dbias_acc);
        if (valid) {
            if (j == 0) {
                if (p.dD) atomicAdd(p.dD + d, dD_acc);
                if (p.ddelta_bias) atomicAdd(p.ddelta_bias + d, dbias_acc);
            }
            store8<T>(reinterpret_cast<T*>(p.du) + r * L, l0, L, vout, du);
            store8<T>(reinterpret_cast<T*>(p.ddelta) + r * L, l0, L, vout, ddt);
        }
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < N * L; idx += 128) {
        const int n = idx / L, l = idx - n * L;
        const int o = n * kSmallL + l;
        const float vb = sm[o] + sm[N * kSmallL + o] + sm[2 * N * kSmallL + o] + sm[3 * N * kSmallL + o];
        const float vc = sm[4 * N * kSmallL + o] + sm[5 * N * kSmallL + o] + sm[6 * N * kSmallL + o] + sm[7 * N * kSmallL + o];
        atomicAdd(p.dB + bg * N * L + idx, vb);
        atomicAdd(p.dC + bg * N * L + idx, vc);
    }
}

static bool small_ok(int64_t L, int64_t N) { return L <= kSmallL && N <= kSmallMaxN; }

template <typename T>
static int launch_small_fwd(const xfs_scan_fwd_args& a, cudaStream_t st) {
    const unsigned grid = (unsigned)((a.batch * a.dim + kRowsPerStep - 1) / kRowsPerStep);
    if (a.out_dtype == XFS_F32) sscan_small_fwd_kernel<T, float><<<grid, 128, 0, st>>>(a);
    else sscan_small_fwd_kernel<T, T><<<grid, 128, 0, st>>>(a);
    return check_launch();
}

template <typename T>
static int launch_small_bwd(const xfs_scan_bwd_args& a, cudaStream_t st) {
    const int64_t Dg = a.dim / a.ngroups;
    const int64_t nblk = (Dg + kRowsPerStep * kRowSteps - 1) / (kRowsPerStep * kRowSteps);
    const unsigned grid = (unsigned)(a.batch * a.ngroups * nblk);
    const size_t smem = sizeof(float) * (size_t)(8 * a.dstate * kSmallL);
    if (a.dout_dtype == XFS_F32) {
        if (int rc = set_smem(sscan_small_bwd_kernel<T, float>, smem)) return rc;
        sscan_small_bwd_kernel<T, float><<<grid, 128, smem, st>>>(a);
    } else {
        if (int rc = set_smem(sscan_small_bwd_kernel<T, T>, smem)) return rc;
        sscan_small_bwd_kernel<T, T><<<grid, 128, smem, st>>>(a);
    }
    return check_launch();
}

int launch_scan_small_fwd(const xfs_scan_fwd_args& a, cudaStream_t st) {
    switch (a.dtype) {
        case XFS_F32: return launch_small_fwd<float>(a, st);
        case XFS_BF16: return launch_small_fwd<__nv_bfloat16>(a, st);
        default: return launch_small_fwd<__half>(a, st);
    }
}
int launch_scan_small_bwd(const xfs_scan_bwd_args& a, cudaStream_t st) {
    switch (a.dtype) {
        case XFS_F32: return launch_small_bwd<float>(a, st);
        case XFS_BF16: return launch_small_bwd<__nv_bfloat16>(a, st);
        default: return launch_small_bwd<__half>(a, st);
    }
}
bool scan_small_supported(int64_t L, int64_t N) { return small_ok(L, N); }

}  // namespace xfs
