// xfscan_common.cuh -- device helpers shared by the sm_100a scan kernels.
//
// Vocabulary (follows the reference): a *sequence* is one (batch, channel) row of length L; a *chunk* is the
// kChunk positions one warp scans per step (32 lanes x kItems consecutive positions); the *state* h is the S6
// recurrence value carried from chunk to chunk; a *route* is one of the 4 CrossScan orders.
#pragma once

#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/xfscan.h"

namespace xfs {

constexpr int kItems = 8;              // consecutive positions per lane
constexpr int kChunk = 32 * kItems;    // positions per warp step (== xfs_chunk_len())
constexpr unsigned kFull = 0xffffffffu;
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;

int check_launch();   // counts the launch (xfs_launch_count) and returns cudaGetLastError(); defined in capi.cu

// ---------------------------------------------------------------------------------------------------------
// transcendental primitives: one MUFU each
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float lg2(float x) {
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float rcp(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// softplus with torch semantics (beta=1, threshold=20; models/csms6s.py:49-50): x > 20 ? x : log1p(exp(x)).
// MUFU.LG2 has ~2^-22 ABSOLUTE error near 1, which is a large RELATIVE error of log1p(e) when e is small -- and
// small e (dt in [1e-3, 1e-1]) is exactly the regime Mamba's dt_bias init puts the model in
// (models/fusion_vmamba.py:303-310).  So below 2^-6 the alternating series is used instead (truncation < 2e-8 rel).
// `e_out` returns exp(x) for the backward's sigmoid.
__device__ __forceinline__ float softplus_fwd(float x, float& e_out) {
    const float e = ex2(x * kLog2e);
    e_out = e;
    const float big = lg2(1.0f + e) * kLn2;
    const float ser = e * fmaf(e, fmaf(e, fmaf(e, -0.25f, 0.33333334f), -0.5f), 1.0f);
    const float r = (e < 0.015625f) ? ser : big;
    return (x > 20.0f) ? x : r;
}

// ---------------------------------------------------------------------------------------------------------
// element type traits: 8 consecutive elements <-> 8 floats
// ---------------------------------------------------------------------------------------------------------
template <typename T> struct Elem;
template <> struct Elem<float> {
    static constexpr int kVec = 4;  // elements per 16-byte vector
    static __device__ __forceinline__ float to_f(float v) { return v; }
    static __device__ __forceinline__ float from_f(float v) { return v; }
};
template <> struct Elem<__nv_bfloat16> {
    static constexpr int kVec = 8;
    static __device__ __forceinline__ float to_f(__nv_bfloat16 v) { return __bfloat162float(v); }
    static __device__ __forceinline__ __nv_bfloat16 from_f(float v) { return __float2bfloat16_rn(v); }
};
template <> struct Elem<__half> {
    static constexpr int kVec = 8;
    static __device__ __forceinline__ float to_f(__half v) { return __half2float(v); }
    static __device__ __forceinline__ __half from_f(float v) { return __float2half_rn(v); }
};

// Streaming 16-byte global accesses.  Loads keep the default L1 policy on purpose: a lane's two vectors of a
// chunk share 32-byte sectors with its neighbours', the second access is an L1 hit.
__device__ __forceinline__ uint4 ldg16(const void* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }
__device__ __forceinline__ void stg16(void* p, uint4 v) { *reinterpret_cast<uint4*>(p) = v; }

// Load elements [l0, l0+8) of a row of length L into v (positions >= L or < 0 get `fill`).  `vec_ok` says the row
// base is 16-byte aligned and L is a multiple of the vector width, so any in-range aligned group may be vector loaded.
// kClampOutside: a lane whose 8 positions ALL lie outside the row reads elements 0..7 instead of getting `fill`
// (keeps it on the vector path; callers must then neutralise those positions themselves -- the fused kernels do, by
// forcing dt = 0 beyond L and never storing them).  Writing this as a third "fill" branch makes ptxas keep the
// destination arrays in local memory, hence the clamp.
template <typename T, bool kClampOutside = false>
__device__ __forceinline__ void load8(const T* __restrict__ row, int64_t l0, int64_t L, bool vec_ok, float (&v)[8],
                                      float fill = 0.0f) {
    if (kClampOutside && (l0 >= L || l0 + 8 <= 0)) l0 = 0;
    if (vec_ok && l0 >= 0 && l0 + 8 <= L) {
        if constexpr (Elem<T>::kVec == 4) {
            const uint4 a = ldg16(row + l0), b = ldg16(row + l0 + 4);
            v[0] = __uint_as_float(a.x); v[1] = __uint_as_float(a.y); v[2] = __uint_as_float(a.z); v[3] = __uint_as_float(a.w);
            v[4] = __uint_as_float(b.x); v[5] = __uint_as_float(b.y); v[6] = __uint_as_float(b.z); v[7] = __uint_as_float(b.w);
        } else {
            const uint4 a = ldg16(row + l0);
            const T* e = reinterpret_cast<const T*>(&a);
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = Elem<T>::to_f(e[i]);
        }
    } else {                                   // straddles an end of the row, or rows are not 16-byte aligned
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int64_t l = l0 + i;
            v[i] = (l >= 0 && l < L) ? Elem<T>::to_f(row[l]) : fill;
        }
    }
}

template <typename T>
__device__ __forceinline__ void store8(T* __restrict__ row, int64_t l0, int64_t L, bool vec_ok, const float (&v)[8]) {
    if (vec_ok && l0 >= 0 && l0 + 8 <= L) {
        if constexpr (Elem<T>::kVec == 4) {
            stg16(row + l0, make_uint4(__float_as_uint(v[0]), __float_as_uint(v[1]), __float_as_uint(v[2]), __float_as_uint(v[3])));
            stg16(row + l0 + 4, make_uint4(__float_as_uint(v[4]), __float_as_uint(v[5]), __float_as_uint(v[6]), __float_as_uint(v[7])));
        } else {
            uint4 a;
            T* e = reinterpret_cast<T*>(&a);
#pragma unroll
            for (int i = 0; i < 8; ++i) e[i] = Elem<T>::from_f(v[i]);
            stg16(row + l0, a);
        }
    } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int64_t l = l0 + i;
            if (l >= 0 && l < L) row[l] = Elem<T>::from_f(v[i]);
        }
    }
}

// ---- "fast rows": every row starts 16-byte aligned and L is a multiple of the 16-byte vector width, so each 16-byte
// granule of a lane's 8 positions is either entirely inside the row or entirely outside.  No scalar edge code at all
// (it would otherwise be inlined into every hot loop and blow the instruction cache).  Granules outside the row are
// read from offset 0 instead (garbage the caller neutralises, see kClampOutside above) and never written.
template <typename T>
__device__ __forceinline__ void load8_fast(const T* __restrict__ row, int l0, int L, float (&v)[8]) {
    if constexpr (Elem<T>::kVec == 4) {
        const int g0 = (l0 >= 0 && l0 + 4 <= L) ? l0 : 0, g1 = (l0 + 4 >= 0 && l0 + 8 <= L) ? l0 + 4 : 0;
        const uint4 a = ldg16(row + g0), b = ldg16(row + g1);
        v[0] = __uint_as_float(a.x); v[1] = __uint_as_float(a.y); v[2] = __uint_as_float(a.z); v[3] = __uint_as_float(a.w);
        v[4] = __uint_as_float(b.x); v[5] = __uint_as_float(b.y); v[6] = __uint_as_float(b.z); v[7] = __uint_as_float(b.w);
    } else {
        const int g0 = (l0 >= 0 && l0 + 8 <= L) ? l0 : 0;
        const uint4 a = ldg16(row + g0);
        const T* e = reinterpret_cast<const T*>(&a);
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = Elem<T>::to_f(e[i]);
    }
}

// fp32 rows, offsets of the two granules already resolved by the caller (see ss2d_fwd.cu: only one chunk per route needs clamping)
template <typename T>
__device__ __forceinline__ void load8_at(const T* __restrict__ row, int g0, int g1, float (&v)[8]) {
    static_assert(Elem<T>::kVec == 4, "fp32 rows only");
    const uint4 a = ldg16(row + g0), b = ldg16(row + g1);
    v[0] = __uint_as_float(a.x); v[1] = __uint_as_float(a.y); v[2] = __uint_as_float(a.z); v[3] = __uint_as_float(a.w);
    v[4] = __uint_as_float(b.x); v[5] = __uint_as_float(b.y); v[6] = __uint_as_float(b.z); v[7] = __uint_as_float(b.w);
}

template <typename T>
__device__ __forceinline__ void store8_fast(T* __restrict__ row, int l0, int L, const float (&v)[8]) {
    if constexpr (Elem<T>::kVec == 4) {
        if (l0 >= 0 && l0 + 4 <= L)
            stg16(row + l0, make_uint4(__float_as_uint(v[0]), __float_as_uint(v[1]), __float_as_uint(v[2]), __float_as_uint(v[3])));
        if (l0 + 4 >= 0 && l0 + 8 <= L)
            stg16(row + l0 + 4, make_uint4(__float_as_uint(v[4]), __float_as_uint(v[5]), __float_as_uint(v[6]), __float_as_uint(v[7])));
    } else {
        if (l0 >= 0 && l0 + 8 <= L) {
            uint4 a;
            T* e = reinterpret_cast<T*>(&a);
#pragma unroll
            for (int i = 0; i < 8; ++i) e[i] = Elem<T>::from_f(v[i]);
            stg16(row + l0, a);
        }
    }
}

template <typename T> __device__ __forceinline__ bool row_vec_ok(const T* base, int64_t L) {
    return (L % Elem<T>::kVec == 0) && ((reinterpret_cast<uintptr_t>(base) & 15) == 0);
}

__device__ __forceinline__ void reverse8(float (&v)[8]) {
#pragma unroll
    for (int i = 0; i < 4; ++i) { const float t = v[i]; v[i] = v[7 - i]; v[7 - i] = t; }
}

// ---------------------------------------------------------------------------------------------------------
// chunked warp-shuffle scan of the affine maps  h -> a*h + b.
//
// Each lane first folds its kItems maps sequentially into (P, S) = (prod a, value of the fold started at 0).
// warp_prefix then turns the per-lane folds into, for every lane, the state ENTERING that lane given the state
// `carry` entering the chunk, and returns the state leaving the chunk (lane-uniform).  kRev walks lanes 31 -> 0.
// 5 steps x (2 SHFL + predicated FFMA + FMUL).
// ---------------------------------------------------------------------------------------------------------
template <bool kRev>
__device__ __forceinline__ float warp_prefix(float P, float S, float carry, int lane, float& chunk_out) {
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const float Pn = kRev ? __shfl_down_sync(kFull, P, off) : __shfl_up_sync(kFull, P, off);
        const float Sn = kRev ? __shfl_down_sync(kFull, S, off) : __shfl_up_sync(kFull, S, off);
        const bool has = kRev ? (lane + off < 32) : (lane >= off);
        if (has) {             // predicated FFMA/FMUL: apply the earlier maps first, then this lane's
            S = fmaf(P, Sn, S);
            P = P * Pn;
        }
    }
    // (P, S) is now the inclusive fold up to and including this lane
    const float incl = fmaf(P, carry, S);
    chunk_out = __shfl_sync(kFull, incl, kRev ? 0 : 31);
    float prev = kRev ? __shfl_down_sync(kFull, incl, 1) : __shfl_up_sync(kFull, incl, 1);
    const bool first = kRev ? (lane == 31) : (lane == 0);
    return first ? carry : prev;
}

// Same scan; the "is there a lane `off` away" predicate of every step comes out of the shuffle itself (shfl.sync's
// predicate destination) instead of an integer compare per step.
template <bool kRev>
__device__ __forceinline__ float warp_prefix_p(float P, float S, float carry, int lane, float& chunk_out) {
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        if (kRev)
            asm volatile("{\n\t.reg .pred p;\n\t.reg .f32 pn, sn;\n\t"
                         "shfl.sync.down.b32 pn|p, %0, %2, 0x1f, 0xffffffff;\n\t"
                         "shfl.sync.down.b32 sn, %1, %2, 0x1f, 0xffffffff;\n\t"
                         "@p fma.rn.ftz.f32 %1, %0, sn, %1;\n\t"
                         "@p mul.ftz.f32 %0, %0, pn;\n\t}"
                         : "+f"(P), "+f"(S) : "r"(off));
        else
            asm volatile("{\n\t.reg .pred p;\n\t.reg .f32 pn, sn;\n\t"
                         "shfl.sync.up.b32 pn|p, %0, %2, 0, 0xffffffff;\n\t"
                         "shfl.sync.up.b32 sn, %1, %2, 0, 0xffffffff;\n\t"
                         "@p fma.rn.ftz.f32 %1, %0, sn, %1;\n\t"
                         "@p mul.ftz.f32 %0, %0, pn;\n\t}"
                         : "+f"(P), "+f"(S) : "r"(off));
    }
    const float incl = fmaf(P, carry, S);
    chunk_out = __shfl_sync(kFull, incl, kRev ? 0 : 31);
    float prev = carry;
    if (kRev)
        asm volatile("{\n\t.reg .pred p;\n\t.reg .f32 t;\n\tshfl.sync.down.b32 t|p, %1, 1, 0x1f, 0xffffffff;\n\t@p mov.f32 %0, t;\n\t}"
                     : "+f"(prev) : "f"(incl));
    else
        asm volatile("{\n\t.reg .pred p;\n\t.reg .f32 t;\n\tshfl.sync.up.b32 t|p, %1, 1, 0, 0xffffffff;\n\t@p mov.f32 %0, t;\n\t}"
                     : "+f"(prev) : "f"(incl));
    (void)lane;
    return prev;
}

// kN independent scans of the same direction, step by step interleaved (the round trips of one hide behind the others').
// carry[] enters the chunk and is replaced by the state leaving it; in[] receives the state entering each lane.
template <bool kRev, int kN>
__device__ __forceinline__ void warp_prefix_pn(float (&P)[kN], float (&S)[kN], float (&carry)[kN], float (&in)[kN]) {
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
#pragma unroll
        for (int n = 0; n < kN; ++n) {
            if (kRev)
                asm volatile("{\n\t.reg .pred p;\n\t.reg .f32 pn, sn;\n\t"
                             "shfl.sync.down.b32 pn|p, %0, %2, 0x1f, 0xffffffff;\n\t"
                             "shfl.sync.down.b32 sn, %1, %2, 0x1f, 0xffffffff;\n\t"
                             "@p fma.rn.ftz.f32 %1, %0, sn, %1;\n\t"
                             "@p mul.ftz.f32 %0, %0, pn;\n\t}"
                             : "+f"(P[n]), "+f"(S[n]) : "r"(off));
            else
                asm volatile("{\n\t.reg .pred p;\n\t.reg .f32 pn, sn;\n\t"
                             "shfl.sync.up.b32 pn|p, %0, %2, 0, 0xffffffff;\n\t"
                             "shfl.sync.up.b32 sn, %1, %2, 0, 0xffffffff;\n\t"
                             "@p fma.rn.ftz.f32 %1, %0, sn, %1;\n\t"
                             "@p mul.ftz.f32 %0, %0, pn;\n\t}"
                             : "+f"(P[n]), "+f"(S[n]) : "r"(off));
        }
    }
#pragma unroll
    for (int n = 0; n < kN; ++n) {
        const float incl = fmaf(P[n], carry[n], S[n]);
        float prev = carry[n];
        if (kRev)
            asm volatile("{\n\t.reg .pred p;\n\t.reg .f32 t;\n\tshfl.sync.down.b32 t|p, %1, 1, 0x1f, 0xffffffff;\n\t@p mov.f32 %0, t;\n\t}"
                         : "+f"(prev) : "f"(incl));
        else
            asm volatile("{\n\t.reg .pred p;\n\t.reg .f32 t;\n\tshfl.sync.up.b32 t|p, %1, 1, 0, 0xffffffff;\n\t@p mov.f32 %0, t;\n\t}"
                         : "+f"(prev) : "f"(incl));
        carry[n] = __shfl_sync(kFull, incl, kRev ? 0 : 31);
        in[n] = prev;
    }
}

// Two independent scans in one pass -- the forward re-scan along the walk direction kRev and the adjoint scan against
// it, as the backward needs them for every chunk.  Shuffles are ordered with respect to each other, so two back-to-back
// warp_prefix calls serialise their 7 round trips each; interleaved, the rounds of one hide the latency of the other.
template <bool kRev>
__device__ __forceinline__ void warp_prefix_dual(float P, float S, float carry, float Pq, float Sq, float qcarry, int lane,
                                                 float& h_in, float& q_in, float& q_out) {
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const float Pn = kRev ? __shfl_down_sync(kFull, P, off) : __shfl_up_sync(kFull, P, off);
        const float Sn = kRev ? __shfl_down_sync(kFull, S, off) : __shfl_up_sync(kFull, S, off);
        const float Pqn = kRev ? __shfl_up_sync(kFull, Pq, off) : __shfl_down_sync(kFull, Pq, off);
        const float Sqn = kRev ? __shfl_up_sync(kFull, Sq, off) : __shfl_down_sync(kFull, Sq, off);
        const bool has = kRev ? (lane + off < 32) : (lane >= off);
        const bool hasq = kRev ? (lane >= off) : (lane + off < 32);
        if (has) { S = fmaf(P, Sn, S); P = P * Pn; }
        if (hasq) { Sq = fmaf(Pq, Sqn, Sq); Pq = Pq * Pqn; }
    }
    const float incl = fmaf(P, carry, S), inclq = fmaf(Pq, qcarry, Sq);
    const float prev = kRev ? __shfl_down_sync(kFull, incl, 1) : __shfl_up_sync(kFull, incl, 1);
    const float prevq = kRev ? __shfl_up_sync(kFull, inclq, 1) : __shfl_down_sync(kFull, inclq, 1);
    q_out = __shfl_sync(kFull, inclq, kRev ? 31 : 0);
    h_in = (kRev ? (lane == 31) : (lane == 0)) ? carry : prev;
    q_in = (kRev ? (lane == 0) : (lane == 31)) ? qcarry : prevq;
}

// ---------------------------------------------------------------------------------------------------------
// packed fp32 (sm_100 FFMA2 / FMUL2 / FADD2: two fp32 lanes per issue slot on a 64-bit register pair)
// ---------------------------------------------------------------------------------------------------------
using f2 = float2;
__device__ __forceinline__ f2 splat2(float v) { return make_float2(v, v); }
__device__ __forceinline__ f2 mul2(f2 a, f2 b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ f2 add2(f2 a, f2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c) { return __ffma2_rn(a, b, c); }
__device__ __forceinline__ f2 ex2_2(f2 a) { return make_float2(ex2(a.x), ex2(a.y)); }

// softplus on a pair (same semantics as softplus_fwd); e_out = exp(x)
__device__ __forceinline__ f2 softplus2(f2 x, f2& e_out) {
    const f2 e = ex2_2(mul2(x, splat2(kLog2e)));
    e_out = e;
    const f2 w = add2(e, splat2(1.0f));
    const f2 big = mul2(make_float2(lg2(w.x), lg2(w.y)), splat2(kLn2));
    f2 ser = fma2(e, splat2(-0.25f), splat2(0.33333334f));
    ser = fma2(ser, e, splat2(-0.5f));
    ser = fma2(ser, e, splat2(1.0f));
    ser = mul2(ser, e);
    f2 r;
    r.x = (e.x < 0.015625f) ? ser.x : big.x;
    r.y = (e.y < 0.015625f) ? ser.y : big.y;
    r.x = (x.x > 20.0f) ? x.x : r.x;
    r.y = (x.y > 20.0f) ? x.y : r.y;
    return r;
}

// softplus of a lane's 8 values in place, for a fully converged warp: the fast lg2(1 + e) form for everybody, and ONE warp vote
// decides whether any element needs the small-argument series (e < 2^-6) or the x > 20 identity of softplus2 (the selects of
// softplus2 cost ~5 instructions per element on the common path)
__device__ __forceinline__ void softplus8_vote(f2 (&x)[4]) {
    f2 e[4], r[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        e[i] = ex2_2(mul2(x[i], splat2(kLog2e)));
        const f2 w = add2(e[i], splat2(1.0f));
        r[i] = mul2(make_float2(lg2(w.x), lg2(w.y)), splat2(kLn2));
    }
    const float emin = fminf(fminf(fminf(e[0].x, e[0].y), fminf(e[1].x, e[1].y)), fminf(fminf(e[2].x, e[2].y), fminf(e[3].x, e[3].y)));
    const float emax = fmaxf(fmaxf(fmaxf(e[0].x, e[0].y), fmaxf(e[1].x, e[1].y)), fmaxf(fmaxf(e[2].x, e[2].y), fmaxf(e[3].x, e[3].y)));
    if (__any_sync(kFull, !(emin >= 0.015625f && emax <= 268435456.0f))) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            f2 ser = fma2(e[i], splat2(-0.25f), splat2(0.33333334f));
            ser = fma2(ser, e[i], splat2(-0.5f));
            ser = fma2(ser, e[i], splat2(1.0f));
            ser = mul2(ser, e[i]);
            f2 q;
            q.x = (e[i].x < 0.015625f) ? ser.x : r[i].x;
            q.y = (e[i].y < 0.015625f) ? ser.y : r[i].y;
            r[i].x = (x[i].x > 20.0f) ? x[i].x : q.x;
            r[i].y = (x[i].y > 20.0f) ? x[i].y : q.y;
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) x[i] = r[i];
}

// 16-byte vector reduction into global memory (sm_90+): one L2 atomic op for 4 consecutive floats
__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// same reduction WITHOUT the compiler-level memory fence: for accumulators this kernel never reads back, so that loads of
// later, independent work may be scheduled across it (ss2d_mid.cu runs four routes as straight-line code)
__device__ __forceinline__ void red_add_v4_relaxed(float* addr, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d));
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(kFull, v, off);
    return v;
}

}  // namespace xfs
