// ss2d_small.cu -- fused SS2D forward / backward for SHORT sequences (L = H*W <= 64), any N <= 16.
//
// XFMamba's last backbone stage (7x7 tokens, D = 1536 / 2048, N = 1) and its deep-fusion block (Cross_SS2Dv5: 7x7 tokens,
// D = 1536 / 2048, N = 16, three streams; reference models/fusion_vmamba.py:446-578) have L = 49.  In the general kernels a
// warp scans 256 positions per step, so only 7 of 32 lanes would work and (batch x D) tiny CTAs would each pay the full
// prologue.  Here one sequence takes an 8-lane group (8 lanes x 8 positions = 64), a warp scans FOUR channels of its route
// at once (3-step segmented shuffle scan), a CTA (4 warps = 4 routes) handles a quad of channels, and in the backward a
// CTA walks several quads so that the dB/dC contributions of all its channels are summed in shared memory before ONE
// atomic per (route, n, l) leaves the CTA (the reference issues one atomic per channel, selective_scan_bwd_kernel.cuh:221).
// Single chunk => no carried state: the backward re-scans from h = 0 and ignores the checkpoints.
#include "ss2d_fused.cuh"

namespace xfs {

// =========================================================================================================
// forward
// =========================================================================================================
template <typename T, typename TO>
__global__ void __launch_bounds__(128)
ss2d_small_fwd_kernel(const xfs_ss2d_fwd_args p) {
    __shared__ __align__(16) float xN[kQuad * kSmallL], xT[kQuad * kSmallL], yN[kQuad * kSmallL], yT[kQuad * kSmallL];
    const int H = (int)p.H, W = (int)p.W, L = H * W, D = (int)p.D, N = (int)p.N;
    const int nquads = (D + kQuad - 1) / kQuad;
    const int b = blockIdx.x / nquads;
    const int d0 = (blockIdx.x - b * nquads) * kQuad;
    const int tid = threadIdx.x, lane = tid & 31, k = tid >> 5, g = lane >> 3, j = lane & 7;
    const int nvalid = min(kQuad, D - d0);
    const bool valid = g < nvalid;
    const int d = d0 + (valid ? g : 0);
    const bool transposed = k & 1;

    stage_quad<T>(reinterpret_cast<const T*>(p.x) + ((int64_t)b * D + d0) * L, L, nvalid, xN, xT, H, W, L, tid);
    __syncthreads();

    const int kd = k * D + d;
    const T* __restrict__ dt_row = reinterpret_cast<const T*>(p.delta) + ((int64_t)b * 4 * D + kd) * L;
    const T* __restrict__ Bk = reinterpret_cast<const T*>(p.Bs) + ((int64_t)b * 4 + k) * N * L;
    const T* __restrict__ Ck = reinterpret_cast<const T*>(p.Cs) + ((int64_t)b * 4 + k) * N * L;
    const bool vin = row_vec_ok(reinterpret_cast<const T*>(p.delta), L) && row_vec_ok(reinterpret_cast<const T*>(p.Bs), L) &&
                     row_vec_ok(reinterpret_cast<const T*>(p.Cs), L);
    const float bias = p.delta_bias ? p.delta_bias[kd] : 0.0f;
    const float Dd = p.Ds ? p.Ds[kd] : 0.0f;
    const int p0 = j * 8;
    const int f4s = swz_f4(p0 >> 2);
    float y[8];

    auto walk = [&](auto rev_tag) __attribute__((always_inline)) {
        constexpr bool rev = decltype(rev_tag)::value;
        const int l0 = rev ? L - 8 - p0 : p0;
        float dta[8], dt[8], u[8];
        load8<T, true>(dt_row, l0, L, vin, dta);
        to_pos<rev>(dta, dt);
        lds8((transposed ? xT : xN) + g * kSmallL, f4s, u);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float xx = dt[i] + bias;
            float e;
            const float sp = p.delta_softplus ? softplus_fwd(xx, e) : xx;
            dt[i] = (p0 + i < L) ? sp : 0.0f;
            y[i] = Dd * u[i];
        }
        for (int n = 0; n < N; ++n) {
            float Ba[8], Ca[8], Bv[8], Cv[8], S[8], P[8];
            load8<T, true>(Bk + n * L, l0, L, vin, Ba);
            load8<T, true>(Ck + n * L, l0, L, vin, Ca);
            to_pos<rev>(Ba, Bv); to_pos<rev>(Ca, Cv);
            const float A2 = p.A[kd * N + n] * kLog2e;
            float Pr = 1.0f, Sr = 0.0f;
#pragma unroll
            for (int ii = 0; ii < 8; ++ii) {
                const int i = rev ? 7 - ii : ii;
                const float a = ex2(dt[i] * A2);
                Sr = fmaf(a, Sr, (dt[i] * Bv[i]) * u[i]);
                Pr *= a;
                S[i] = Sr; P[i] = Pr;
            }
            float h_end;
            const float h_in = group_prefix<rev>(Pr, Sr, j, h_end);
#pragma unroll
            for (int i = 0; i < 8; ++i) y[i] = fmaf(Cv[i], fmaf(P[i], h_in, S[i]), y[i]);
            if (p.states && valid && j == 0) p.states[((int64_t)b * 4 * D + kd) * N + n] = h_end;   // one chunk per row
        }
    };
    if (k >= 2) walk(std::true_type{}); else walk(std::false_type{});

    float* yb = (transposed ? yT : yN) + g * kSmallL;
    if (k < 2) sts8(yb, f4s, y);
    __syncthreads();
    if (k >= 2) {
        float o[8];
        lds8(yb, f4s, o);
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] += y[i];            // (y_k + y_{k+2}) in the reference's order
        sts8(yb, f4s, o);
    }
    __syncthreads();

    TO* __restrict__ out = reinterpret_cast<TO*>(p.y) + ((int64_t)b * D + d0) * L;
    for (int idx = tid; idx < nvalid * L; idx += 128) {
        const int ch = idx / L, pp = idx - ch * L;
        const int h = pp / W, w = pp - h * W;
        out[idx] = Elem<TO>::from_f(yN[ch * kSmallL + swz_pos(pp)] + yT[ch * kSmallL + swz_pos(w * H + h)]);
    }
}

// =========================================================================================================
// backward
// =========================================================================================================
template <typename T, typename TDO>
__global__ void __launch_bounds__(128)
ss2d_small_bwd_kernel(const xfs_ss2d_bwd_args p) {
    extern __shared__ __align__(16) float sm[];
    const int H = (int)p.H, W = (int)p.W, L = H * W, D = (int)p.D, N = (int)p.N;
    float* xN = sm;                               // [kQuad][64] each
    float* xT = xN + kQuad * kSmallL;
    float* gN = xT + kQuad * kSmallL;
    float* gT = gN + kQuad * kSmallL;
    float* dN = gT + kQuad * kSmallL;
    float* dT = dN + kQuad * kSmallL;
    float* sdB = dT + kQuad * kSmallL;            // [4 routes][N][64]  (scan order of the route)
    float* sdC = sdB + 4 * N * kSmallL;

    const int nquads = (D + kQuad - 1) / kQuad;
    const int nblk = (nquads + kQuadsPerCta - 1) / kQuadsPerCta;
    const int b = blockIdx.x / nblk;
    const int q_begin = (blockIdx.x - b * nblk) * kQuadsPerCta;
    const int q_end = min(q_begin + kQuadsPerCta, nquads);
    const int tid = threadIdx.x, lane = tid & 31, k = tid >> 5, g = lane >> 3, j = lane & 7;
    const bool transposed = k & 1;
    const int p0 = j * 8;
    const int f4s = swz_f4(p0 >> 2);

    for (int i = tid; i < 2 * 4 * N * kSmallL; i += 128) sdB[i] = 0.0f;

    const T* __restrict__ Bk = reinterpret_cast<const T*>(p.Bs) + ((int64_t)b * 4 + k) * N * L;
    const T* __restrict__ Ck = reinterpret_cast<const T*>(p.Cs) + ((int64_t)b * 4 + k) * N * L;
    const bool vin = row_vec_ok(reinterpret_cast<const T*>(p.delta), L) && row_vec_ok(reinterpret_cast<const T*>(p.Bs), L) &&
                     row_vec_ok(reinterpret_cast<const T*>(p.Cs), L);
    const bool vout = row_vec_ok(reinterpret_cast<const T*>(p.ddelta), L);
    float* mydB = sdB + k * N * kSmallL;
    float* mydC = sdC + k * N * kSmallL;

    for (int q = q_begin; q < q_end; ++q) {
        const int d0 = q * kQuad;
        const int nvalid = min(kQuad, D - d0);
        const bool valid = g < nvalid;
        const int d = d0 + (valid ? g : 0);
        const int kd = k * D + d;
        __syncthreads();                           // previous quad's dx merge is done with the buffers
        stage_quad<T>(reinterpret_cast<const T*>(p.x) + ((int64_t)b * D + d0) * L, L, nvalid, xN, xT, H, W, L, tid);
        stage_quad<TDO>(reinterpret_cast<const TDO*>(p.dy) + ((int64_t)b * D + d0) * L, L, nvalid, gN, gT, H, W, L, tid);
        __syncthreads();

        const T* __restrict__ dt_row = reinterpret_cast<const T*>(p.delta) + ((int64_t)b * 4 * D + kd) * L;
        T* __restrict__ ddt_row = reinterpret_cast<T*>(p.ddelta) + ((int64_t)b * 4 * D + kd) * L;
        const float bias = p.delta_bias ? p.delta_bias[kd] : 0.0f;
        const float Dd = p.Ds ? p.Ds[kd] : 0.0f;
        float du[8];

        auto walk = [&](auto rev_tag) __attribute__((always_inline)) {
            constexpr bool rev = decltype(rev_tag)::value;
            const int l0 = rev ? L - 8 - p0 : p0;
            float dta[8], dt[8], u[8], dy[8], sig[8], ddt[8];
            load8<T, true>(dt_row, l0, L, vin, dta);
            to_pos<rev>(dta, dt);
            lds8((transposed ? xT : xN) + g * kSmallL, f4s, u);
            lds8((transposed ? gT : gN) + g * kSmallL, f4s, dy);
            float dD_acc = 0.0f;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float xx = dt[i] + bias;
                float e = 0.0f;
                const float sp = p.delta_softplus ? softplus_fwd(xx, e) : xx;
                sig[i] = p.delta_softplus ? ((xx > 20.0f) ? 1.0f : e * rcp(1.0f + e)) : 1.0f;
                dt[i] = (p0 + i < L) ? sp : 0.0f;
                du[i] = Dd * dy[i];
                ddt[i] = 0.0f;
                dD_acc = fmaf(dy[i], u[i], dD_acc);
            }
            for (int n = 0; n < N; ++n) {
                float Ba[8], Ca[8], Bv[8], Cv[8], a[8], bu[8], S[8], P[8], Sq[8], Pq[8];
                load8<T, true>(Bk + n * L, l0, L, vin, Ba);
                load8<T, true>(Ck + n * L, l0, L, vin, Ca);
                to_pos<rev>(Ba, Bv); to_pos<rev>(Ca, Cv);
                const float An = p.A[kd * N + n];
                const float A2 = An * kLog2e;
                float Pr = 1.0f, Sr = 0.0f;
#pragma unroll
                for (int ii = 0; ii < 8; ++ii) {
                    const int i = rev ? 7 - ii : ii;
                    a[i] = ex2(dt[i] * A2);
                    bu[i] = (dt[i] * Bv[i]) * u[i];
                    Sr = fmaf(a[i], Sr, bu[i]); Pr *= a[i]; S[i] = Sr; P[i] = Pr;
                }
                float unused;
                const float h_in = group_prefix<rev>(Pr, Sr, j, unused);
                Pr = 1.0f; Sr = 0.0f;
#pragma unroll
                for (int ii = 0; ii < 8; ++ii) {                    // adjoint walks the other way
                    const int i = rev ? ii : 7 - ii;
                    Sr = a[i] * fmaf(Cv[i], dy[i], Sr); Pr *= a[i]; Sq[i] = Sr; Pq[i] = Pr;
                }
                const float q_in = group_prefix<!rev>(Pr, Sr, j, unused);
                float dA_part = 0.0f, dBv[8], dCv[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float h = fmaf(P[i], h_in, S[i]);
                    const int inx = rev ? (i == 0 ? 0 : i - 1) : (i == 7 ? 7 : i + 1);
                    const bool edge = rev ? (i == 0) : (i == 7);
                    const float q_next = edge ? q_in : fmaf(Pq[inx], q_in, Sq[inx]);
                    const float gi = fmaf(Cv[i], dy[i], q_next);
                    const float hp = h - bu[i];
                    const float gdt = gi * dt[i];
                    du[i] = fmaf(gdt, Bv[i], du[i]);
                    ddt[i] = fmaf(gi, fmaf(Bv[i], u[i], An * hp), ddt[i]);
                    dA_part = fmaf(gdt, hp, dA_part);
                    dBv[i] = valid ? gdt * u[i] : 0.0f;
                    dCv[i] = valid ? dy[i] * h : 0.0f;
                }
                // channels of the warp summed by shuffles, then accumulated (position order) by group 0 into smem
#pragma unroll
                for (int i = 0; i < 8; ++i) { dBv[i] = quad_sum(dBv[i]); dCv[i] = quad_sum(dCv[i]); }
                if (g == 0) {
                    float* rb = mydB + n * kSmallL + p0;
                    float* rc = mydC + n * kSmallL + p0;
#pragma unroll
                    for (int i = 0; i < 8; ++i) { rb[i] += dBv[i]; rc[i] += dCv[i]; }
                }
                dA_part = group_sum(dA_part);
                if (valid && j == 0) atomicAdd(p.dA + kd * N + n, dA_part);
            }
            float dbias_acc = 0.0f;
#pragma unroll
            for (int i = 0; i < 8; ++i) { ddt[i] *= sig[i]; dbias_acc += ddt[i]; }
            dD_acc = group_sum(dD_acc);
            dbias_acc = group_sum(dbias_acc);
            if (valid && j == 0) {
                if (p.dDs) atomicAdd(p.dDs + kd, dD_acc);
                if (p.ddelta_bias) atomicAdd(p.ddelta_bias + kd, dbias_acc);
            }
            if (valid) {
                float dda[8];
                to_pos<rev>(ddt, dda);
                store8<T>(ddt_row, l0, L, vout, dda);
            }
        };
        if (k >= 2) walk(std::true_type{}); else walk(std::false_type{});

        float* dbuf = (transposed ? dT : dN) + g * kSmallL;
        if (k < 2) sts8(dbuf, f4s, du);
        __syncthreads();
        if (k >= 2) {
            float o[8];
            lds8(dbuf, f4s, o);
#pragma unroll
            for (int i = 0; i < 8; ++i) o[i] += du[i];
            sts8(dbuf, f4s, o);
        }
        __syncthreads();
        T* __restrict__ dx = reinterpret_cast<T*>(p.dx) + ((int64_t)b * D + d0) * L;
        for (int idx = tid; idx < nvalid * L; idx += 128) {
            const int ch = idx / L, pp = idx - ch * L;
            const int h = pp / W, w = pp - h * W;
            dx[idx] = Elem<T>::from_f(dN[ch * kSmallL + swz_pos(pp)] + dT[ch * kSmallL + swz_pos(w * H + h)]);
        }
    }
    __syncthreads();
    // flush dB / dC: smem rows are in POSITION order; scan index l = p (routes 0/1) or L-1-p (routes 2/3)
    for (int idx = tid; idx < 4 * N * L; idx += 128) {
        const int r = idx / (N * L), rem = idx - r * N * L, n = rem / L, l = rem - n * L;
        const int pp = (r >= 2) ? L - 1 - l : l;
        const int rep = p.acc_replicas > 1 ? (int)((blockIdx.x - b * nblk) % p.acc_replicas) : 0;   // see xfscan.h
        const int64_t off = ((((int64_t)rep * p.batch + b) * 4 + r) * N + n) * L + l;
        atomicAdd(p.dBs + off, sdB[(r * N + n) * kSmallL + pp]);
        atomicAdd(p.dCs + off, sdC[(r * N + n) * kSmallL + pp]);
    }
}

// ---- host side --------------------------------------------------------------------------------------------------
int ss2d_small_supported(int64_t N, int64_t H, int64_t W) { return H * W <= kSmallL && N >= 1 && N <= kSmallMaxN; }

template <typename T, typename TO>
static int launch_small_fwd_t(const xfs_ss2d_fwd_args& a, cudaStream_t st) {
    const unsigned grid = (unsigned)(a.batch * ((a.D + kQuad - 1) / kQuad));
    ss2d_small_fwd_kernel<T, TO><<<grid, 128, 0, st>>>(a);
    return check_launch();
}

int launch_ss2d_small_fwd(const xfs_ss2d_fwd_args& a, cudaStream_t st) {
    const bool o32 = a.out_dtype == XFS_F32;
    switch (a.dtype) {
        case XFS_F32: return launch_small_fwd_t<float, float>(a, st);
        case XFS_BF16: return o32 ? launch_small_fwd_t<__nv_bfloat16, float>(a, st) : launch_small_fwd_t<__nv_bfloat16, __nv_bfloat16>(a, st);
        default: return o32 ? launch_small_fwd_t<__half, float>(a, st) : launch_small_fwd_t<__half, __half>(a, st);
    }
}

template <typename T, typename TDO>
static int launch_small_bwd_t(const xfs_ss2d_bwd_args& a, cudaStream_t st) {
    const int64_t nquads = (a.D + kQuad - 1) / kQuad;
    const unsigned grid = (unsigned)(a.batch * ((nquads + kQuadsPerCta - 1) / kQuadsPerCta));
    const size_t smem = sizeof(float) * (size_t)(6 * kQuad * kSmallL + 2 * 4 * a.N * kSmallL);
    if (int rc = set_smem(ss2d_small_bwd_kernel<T, TDO>, smem)) return rc;
    ss2d_small_bwd_kernel<T, TDO><<<grid, 128, smem, st>>>(a);
    return check_launch();
}

int launch_ss2d_small_bwd(const xfs_ss2d_bwd_args& a, cudaStream_t st) {
    const bool g32 = a.dout_dtype == XFS_F32;
    switch (a.dtype) {
        case XFS_F32: return launch_small_bwd_t<float, float>(a, st);
        case XFS_BF16: return g32 ? launch_small_bwd_t<__nv_bfloat16, float>(a, st) : launch_small_bwd_t<__nv_bfloat16, __nv_bfloat16>(a, st);
        default: return g32 ? launch_small_bwd_t<__half, float>(a, st) : launch_small_bwd_t<__half, __half>(a, st);
    }
}

}  // namespace xfs
