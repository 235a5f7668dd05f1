// fusion_small.cu -- the cross-view fusion scans of XFMamba at their real shapes: L = H*W <= 64 positions, N <= 16 states.
//
//   xfs_cross_ss2d_x3_fwd / _bwd : the three SS2D streams of Cross_SS2Dv5.forward_corev2 (x_fuse, x, x2: one parameter
//       set, own delta / B rows, the view streams reading the fused stream's C; reference models/fusion_vmamba.py
//       :485-569) in ONE launch (grid.y = stream) instead of three.
//   xfs_swap_scan_fused_fwd / _bwd : ShallowFuse_SS2Dv4's scan (:812, :831-835) with the channel-swapping gather
//       (SwappingScan_multiview, :189-221) and the split (SwappingMerge_multiview, :224-241) folded in: u is read from x / x2 by
//       channel parity, the two output halves go straight to y / y2; the backward follows the reference AS WRITTEN
//       (gradients are handed back per half without un-swapping, :217-221).
//
// Same arithmetic as ss2d_small.cu / sscan_small_* (8-lane group = one sequence, 8 positions per lane, segmented 3-step
// shuffle scan, all N states of a channel on the same lanes).  What changes is where the bytes come from.  With L = 49 no row
// is 16-byte aligned, so the older kernels fetch B and C -- per state, per lane -- with 16 scalar range-checked loads, and
// every 8-lane group of every CTA fetches the same rows again.  Here a CTA owns one (stream, batch image[, group]) and a block
// of channels: the B / C rows of all its routes are staged ONCE, coalesced, into shared memory (position order, pitch 64,
// zero beyond L) and read back as two LDS.128 per state; delta rows take the same path.  At N = 16 these kernels are bound
// by the 17 MUFU.EX2 + 1 MUFU.LG2 per (b, k, d, l) (16 lanes/clk/SM), not by HBM -- bench.py reports them against that.
#include "ss2d_fused.cuh"

namespace xfs {

constexpr int kFsQuads = 8;          // quads (4 channels) a CTA walks: 32 channels share one staging of B / C

struct FsX3Fwd {
    const void* x[3]; const void* delta[3]; const void* Bs[3]; const void* Cs[3];
    void* y[3]; float* states[3];
    const float* A; const float* Ds; const float* bias;
    int batch, D, N, H, W, softplus;
};
struct FsX3Bwd {
    const void* x[3]; const void* delta[3]; const void* Bs[3]; const void* Cs[3]; const void* dy[3];
    void* dx[3]; void* ddelta[3]; float* dBs[3]; float* dCs[3];
    const float* A; const float* Ds; const float* bias;
    float* dA; float* dDs; float* dbias;
    int batch, D, N, H, W, softplus;
};

// element-wise staging loops: kBatch independent global loads are issued before the first shared-memory store, so that a
// thread has kBatch requests in flight instead of one (a plain load-store loop serialises on the load latency: the
// compiler may not move a global load above a store through a generic pointer)
template <int kBatch, typename LoadF, typename StoreF>
__device__ __forceinline__ void staged_loop(int begin, int end, int stride, LoadF ld, StoreF st) {
    for (int base = begin; base < end; base += stride * kBatch) {
        float v[kBatch];
#pragma unroll
        for (int u = 0; u < kBatch; ++u) {
            const int idx = base + u * stride;
            v[u] = idx < end ? ld(idx) : 0.0f;
        }
#pragma unroll
        for (int u = 0; u < kBatch; ++u) {
            const int idx = base + u * stride;
            if (idx < end) st(idx, v[u]);
        }
    }
}

// `nrows` rows of L contiguous elements each (row i at src + i * L) -> shared-memory rows of pitch 64 in POSITION order
// (row i at dst + i * 64; a flipped route walks l = L-1-p), positions >= L zeroed.  One warp per row, lanes over l: no integer
// division, and the two loads of a lane are issued before its stores.
template <typename T>
__device__ __forceinline__ void stage_rows(const T* __restrict__ src, float* __restrict__ dst, int nrows, int L, bool flipped, int tid,
                                           int nthreads) {
    const int lane = tid & 31, nwarps = nthreads >> 5;
    for (int row = tid >> 5; row < nrows; row += nwarps) {
        const T* __restrict__ rs = src + row * L;
        const float v0 = lane < L ? Elem<T>::to_f(rs[lane]) : 0.0f;
        const float v1 = lane + 32 < L ? Elem<T>::to_f(rs[lane + 32]) : 0.0f;
        float* rd = dst + row * kSmallL;
        if (!flipped) { rd[lane] = v0; rd[lane + 32] = v1; }
        else {
            if (lane < L) rd[L - 1 - lane] = v0;
            if (lane + 32 < L) rd[L - 1 - (lane + 32)] = v1;
            if (lane >= L) rd[lane] = 0.0f;
            if (lane + 32 >= L) rd[lane + 32] = 0.0f;
        }
    }
}
// the reverse: shared-memory rows (position order) -> global rows of L elements
template <typename T>
__device__ __forceinline__ void unstage_rows(T* __restrict__ dstg, const float* __restrict__ srcs, int nrows, int L, bool flipped, int tid,
                                             int nthreads) {
    const int lane = tid & 31, nwarps = nthreads >> 5;
    for (int row = tid >> 5; row < nrows; row += nwarps) {
        const float* rs = srcs + row * kSmallL;
        T* __restrict__ rd = dstg + row * L;
        if (lane < L) rd[lane] = Elem<T>::from_f(rs[flipped ? L - 1 - lane : lane]);
        if (lane + 32 < L) rd[lane + 32] = Elem<T>::from_f(rs[flipped ? L - 1 - (lane + 32) : lane + 32]);
    }
}
// position p -> swizzled offset of the same pixel in the column-major copy (computed once per CTA: 64 entries)
__device__ __forceinline__ void build_tpos(int* tpos, int H, int W, int L, int tid) {
    if (tid < kSmallL) {
        int v = swz_pos(tid);
        if (tid < L) { const int h = tid / W, w = tid - h * W; v = swz_pos(w * H + h); }
        tpos[tid] = v;
    }
}
// kQuad channel images (rows of L elements, row ch at base + ch * L) -> row-major / column-major swizzled rows of 64
template <typename T>
__device__ __forceinline__ void stage_quad_t(const T* __restrict__ base, int nvalid, float* bN, float* bT, const int* tpos, int L, int tid) {
    const int lane = tid & 31, ch = tid >> 5;          // 4 warps = 4 channels
    const T* __restrict__ rs = base + ch * L;
    const bool ok = ch < nvalid;
    const float v0 = (ok && lane < L) ? Elem<T>::to_f(rs[lane]) : 0.0f;
    const float v1 = (ok && lane + 32 < L) ? Elem<T>::to_f(rs[lane + 32]) : 0.0f;
    bN[ch * kSmallL + swz_pos(lane)] = v0;
    bN[ch * kSmallL + swz_pos(lane + 32)] = v1;
    bT[ch * kSmallL + tpos[lane]] = v0;
    bT[ch * kSmallL + tpos[lane + 32]] = v1;
}
// out[ch][p] = aN[ch][p] + aT[ch][T(p)] for the nvalid channels of a quad
template <typename TO>
__device__ __forceinline__ void merge_quad_t(TO* __restrict__ out, int nvalid, const float* aN, const float* aT, const int* tpos, int L, int tid) {
    const int lane = tid & 31, ch = tid >> 5;
    if (ch < nvalid) {
        TO* __restrict__ rd = out + ch * L;
        if (lane < L) rd[lane] = Elem<TO>::from_f(aN[ch * kSmallL + swz_pos(lane)] + aT[ch * kSmallL + tpos[lane]]);
        if (lane + 32 < L) rd[lane + 32] = Elem<TO>::from_f(aN[ch * kSmallL + swz_pos(lane + 32)] + aT[ch * kSmallL + tpos[lane + 32]]);
    }
}
__device__ __forceinline__ void lds8_lin(const float* row, int p0, float (&v)[8]) {
    const float4 a = *reinterpret_cast<const float4*>(row + p0), b = *reinterpret_cast<const float4*>(row + p0 + 4);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}

// ---- a quad's rows in REGISTERS: loaded one quad ahead (in flight while the current quad is computed), stored to shared memory at
// the top of the next step.  Warp w of the CTA owns channel w of the quad: two elements per lane and row (l = lane, lane + 32).
struct QuadRegs {
    float x[2], g[2], dt[4][2];
};
template <typename T, typename TG, bool kWithG>
__device__ __forceinline__ void quad_load(QuadRegs& r, const T* __restrict__ x, const TG* __restrict__ gsrc, const T* __restrict__ delta,
                                          int64_t img0, int64_t drow0, int64_t D, int nvalid, int L, int tid) {
    const int lane = tid & 31, ch = tid >> 5;
    const bool ok0 = ch < nvalid && lane < L, ok1 = ch < nvalid && lane + 32 < L;
    const int64_t xo = (img0 + ch) * L;
    r.x[0] = ok0 ? Elem<T>::to_f(x[xo + lane]) : 0.0f;
    r.x[1] = ok1 ? Elem<T>::to_f(x[xo + lane + 32]) : 0.0f;
    if (kWithG) {
        r.g[0] = ok0 ? Elem<TG>::to_f(gsrc[xo + lane]) : 0.0f;
        r.g[1] = ok1 ? Elem<TG>::to_f(gsrc[xo + lane + 32]) : 0.0f;
    }
#pragma unroll
    for (int rt = 0; rt < 4; ++rt) {
        const int64_t ro = (drow0 + (int64_t)rt * D + ch) * L;
        r.dt[rt][0] = ok0 ? Elem<T>::to_f(delta[ro + lane]) : 0.0f;
        r.dt[rt][1] = ok1 ? Elem<T>::to_f(delta[ro + lane + 32]) : 0.0f;
    }
}
template <bool kWithG>
__device__ __forceinline__ void quad_store(const QuadRegs& r, float* xN, float* xT, float* gN, float* gT, float* sdt, const int* tpos,
                                           int nvalid, int L, int tid) {
    const int lane = tid & 31, ch = tid >> 5;
    xN[ch * kSmallL + swz_pos(lane)] = r.x[0];
    xN[ch * kSmallL + swz_pos(lane + 32)] = r.x[1];
    xT[ch * kSmallL + tpos[lane]] = r.x[0];
    xT[ch * kSmallL + tpos[lane + 32]] = r.x[1];
    if (kWithG) {
        gN[ch * kSmallL + swz_pos(lane)] = r.g[0];
        gN[ch * kSmallL + swz_pos(lane + 32)] = r.g[1];
        gT[ch * kSmallL + tpos[lane]] = r.g[0];
        gT[ch * kSmallL + tpos[lane + 32]] = r.g[1];
    }
    if (ch < nvalid) {
#pragma unroll
        for (int rt = 0; rt < 4; ++rt) {                 // position order: a flipped route walks l = L-1-p
            float* rd = sdt + (rt * kQuad + ch) * kSmallL;
            if (rt < 2) { rd[lane] = r.dt[rt][0]; rd[lane + 32] = r.dt[rt][1]; }
            else {
                if (lane < L) rd[L - 1 - lane] = r.dt[rt][0];
                if (lane + 32 < L) rd[L - 1 - (lane + 32)] = r.dt[rt][1];
                if (lane >= L) rd[lane] = 0.0f;
                if (lane + 32 >= L) rd[lane + 32] = 0.0f;
            }
        }
    }
}

// =========================================================================================================
// three SS2D streams, forward
// =========================================================================================================
template <typename T, typename TO>
__global__ void __launch_bounds__(128)
fs_x3_fwd_kernel(const FsX3Fwd p) {
    extern __shared__ __align__(16) float sm[];
    const int H = p.H, W = p.W, L = H * W, D = p.D, N = p.N, s = blockIdx.y;
    float* xN = sm;                                   // [kQuad][64] each
    float* xT = xN + kQuad * kSmallL;
    float* yN = xT + kQuad * kSmallL;
    float* yT = yN + kQuad * kSmallL;
    float* sdt = yT + kQuad * kSmallL;                // [4 routes][kQuad][64], position order
    float* sB = sdt + 4 * kQuad * kSmallL;            // [4 routes][N][64], position order
    float* sC = sB + 4 * N * kSmallL;
    int* tpos = reinterpret_cast<int*>(sC + 4 * N * kSmallL);      // [64]

    const int nquads = (D + kQuad - 1) / kQuad;
    const int nblk = (nquads + kFsQuads - 1) / kFsQuads;
    const int b = blockIdx.x / nblk;
    const int q_begin = (blockIdx.x - b * nblk) * kFsQuads, q_end = min(q_begin + kFsQuads, nquads);
    const int tid = threadIdx.x, lane = tid & 31, k = tid >> 5, g = lane >> 3, j = lane & 7;
    const bool transposed = k & 1, rev = k >= 2;
    const int p0 = j * 8;
    const int f4s = swz_f4(p0 >> 2);

    const T* __restrict__ Bsrc = reinterpret_cast<const T*>(p.Bs[s]) + (int64_t)b * 4 * N * L;
    const T* __restrict__ Csrc = reinterpret_cast<const T*>(p.Cs[s]) + (int64_t)b * 4 * N * L;
    stage_rows<T>(Bsrc, sB, 2 * N, L, false, tid, 128);
    stage_rows<T>(Bsrc + 2 * N * L, sB + 2 * N * kSmallL, 2 * N, L, true, tid, 128);
    stage_rows<T>(Csrc, sC, 2 * N, L, false, tid, 128);
    stage_rows<T>(Csrc + 2 * N * L, sC + 2 * N * kSmallL, 2 * N, L, true, tid, 128);
    build_tpos(tpos, H, W, L, tid);
    const float* myB = sB + k * N * kSmallL;
    const float* myC = sC + k * N * kSmallL;

    const T* __restrict__ xsrc = reinterpret_cast<const T*>(p.x[s]);
    const T* __restrict__ dsrc = reinterpret_cast<const T*>(p.delta[s]);
    QuadRegs qr;
    if (q_begin < q_end)
        quad_load<T, T, false>(qr, xsrc, xsrc, dsrc, (int64_t)b * D + q_begin * kQuad, (int64_t)b * 4 * D + q_begin * kQuad, D,
                               min(kQuad, D - q_begin * kQuad), L, tid);
    for (int q = q_begin; q < q_end; ++q) {
        const int d0 = q * kQuad;
        const int nvalid = min(kQuad, D - d0);
        const bool valid = g < nvalid;
        const int d = d0 + (valid ? g : 0);
        const int kd = k * D + d;
        __syncthreads();                               // the previous quad's merge is done with the buffers
        quad_store<false>(qr, xN, xT, nullptr, nullptr, sdt, tpos, nvalid, L, tid);
        if (q + 1 < q_end)                             // the next quad's rows fly while this one is computed
            quad_load<T, T, false>(qr, xsrc, xsrc, dsrc, (int64_t)b * D + d0 + kQuad, (int64_t)b * 4 * D + d0 + kQuad, D,
                                   min(kQuad, D - d0 - kQuad), L, tid);
        __syncthreads();

        const float bias = p.bias ? p.bias[kd] : 0.0f;
        const float Dd = p.Ds ? p.Ds[kd] : 0.0f;
        float y[8];
        auto walk = [&](auto rev_tag) __attribute__((always_inline)) {
        constexpr bool R = decltype(rev_tag)::value;
        float dt[8], u[8];
        lds8_lin(sdt + (k * kQuad + g) * kSmallL, p0, dt);
        lds8((transposed ? xT : xN) + g * kSmallL, f4s, u);
        float dtu[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float xx = dt[i] + bias;
            float e;
            const float sp = p.softplus ? softplus_fwd(xx, e) : xx;
            dt[i] = (p0 + i < L) ? sp : 0.0f;
            dtu[i] = dt[i] * u[i];
            y[i] = Dd * u[i];
        }
        const float* Arow = p.A + (int64_t)kd * N;
        for (int n = 0; n < N; ++n) {
            float Bv[8], Cv[8], S[8], P[8];
            lds8_lin(myB + n * kSmallL, p0, Bv);
            lds8_lin(myC + n * kSmallL, p0, Cv);
            const float A2 = Arow[n] * kLog2e;
            float Pr = 1.0f, Sr = 0.0f;
#pragma unroll
            for (int ii = 0; ii < 8; ++ii) {
                const int i = R ? 7 - ii : ii;
                const float a = ex2(dt[i] * A2);
                Sr = fmaf(a, Sr, dtu[i] * Bv[i]);
                Pr *= a;
                S[i] = Sr; P[i] = Pr;
            }
            float h_end;
            const float h_in = group_prefix<R>(Pr, Sr, j, h_end);
#pragma unroll
            for (int i = 0; i < 8; ++i) y[i] = fmaf(Cv[i], fmaf(P[i], h_in, S[i]), y[i]);
            if (p.states[s] && valid && j == 0) p.states[s][((int64_t)b * 4 * D + kd) * N + n] = h_end;
        }
        };
        if (rev) walk(std::true_type{}); else walk(std::false_type{});
        float* yb = (transposed ? yT : yN) + g * kSmallL;
        if (k < 2) sts8(yb, f4s, y);
        __syncthreads();
        if (k >= 2) {
            float o[8];
            lds8(yb, f4s, o);
#pragma unroll
            for (int i = 0; i < 8; ++i) o[i] += y[i];
            sts8(yb, f4s, o);
        }
        __syncthreads();
        merge_quad_t<TO>(reinterpret_cast<TO*>(p.y[s]) + ((int64_t)b * D + d0) * L, nvalid, yN, yT, tpos, L, tid);
    }
}

// =========================================================================================================
// three SS2D streams, backward
// =========================================================================================================
template <typename T, typename TDO>
__global__ void __launch_bounds__(128)
fs_x3_bwd_kernel(const FsX3Bwd p) {
    extern __shared__ __align__(16) float sm[];
    const int H = p.H, W = p.W, L = H * W, D = p.D, N = p.N, s = blockIdx.y;
    float* xN = sm;
    float* xT = xN + kQuad * kSmallL;
    float* gN = xT + kQuad * kSmallL;
    float* gT = gN + kQuad * kSmallL;
    float* dN = gT + kQuad * kSmallL;
    float* dT = dN + kQuad * kSmallL;
    float* sdt = dT + kQuad * kSmallL;                // [4][kQuad][64]: delta in, ddelta out (position order)
    float* sB = sdt + 4 * kQuad * kSmallL;            // [4][N][64]
    float* sC = sB + 4 * N * kSmallL;
    float* sdB = sC + 4 * N * kSmallL;                // [4][N][64] accumulators, position order
    float* sdC = sdB + 4 * N * kSmallL;
    int* tpos = reinterpret_cast<int*>(sdC + 4 * N * kSmallL);     // [64]

    const int nquads = (D + kQuad - 1) / kQuad;
    const int nblk = (nquads + kFsQuads - 1) / kFsQuads;
    const int b = blockIdx.x / nblk;
    const int q_begin = (blockIdx.x - b * nblk) * kFsQuads, q_end = min(q_begin + kFsQuads, nquads);
    const int tid = threadIdx.x, lane = tid & 31, k = tid >> 5, g = lane >> 3, j = lane & 7;
    const bool transposed = k & 1, rev = k >= 2;
    const int p0 = j * 8;
    const int f4s = swz_f4(p0 >> 2);

    const T* __restrict__ Bsrc = reinterpret_cast<const T*>(p.Bs[s]) + (int64_t)b * 4 * N * L;
    const T* __restrict__ Csrc = reinterpret_cast<const T*>(p.Cs[s]) + (int64_t)b * 4 * N * L;
    stage_rows<T>(Bsrc, sB, 2 * N, L, false, tid, 128);
    stage_rows<T>(Bsrc + 2 * N * L, sB + 2 * N * kSmallL, 2 * N, L, true, tid, 128);
    stage_rows<T>(Csrc, sC, 2 * N, L, false, tid, 128);
    stage_rows<T>(Csrc + 2 * N * L, sC + 2 * N * kSmallL, 2 * N, L, true, tid, 128);
    build_tpos(tpos, H, W, L, tid);
    for (int i = tid; i < 2 * 4 * N * kSmallL; i += 128) sdB[i] = 0.0f;
    const float* myB = sB + k * N * kSmallL;
    const float* myC = sC + k * N * kSmallL;
    float* mydB = sdB + k * N * kSmallL;
    float* mydC = sdC + k * N * kSmallL;

    const T* __restrict__ xsrc = reinterpret_cast<const T*>(p.x[s]);
    const TDO* __restrict__ gsrc = reinterpret_cast<const TDO*>(p.dy[s]);
    const T* __restrict__ dsrc = reinterpret_cast<const T*>(p.delta[s]);
    QuadRegs qr;
    if (q_begin < q_end)
        quad_load<T, TDO, true>(qr, xsrc, gsrc, dsrc, (int64_t)b * D + q_begin * kQuad, (int64_t)b * 4 * D + q_begin * kQuad, D,
                                min(kQuad, D - q_begin * kQuad), L, tid);
    for (int q = q_begin; q < q_end; ++q) {
        const int d0 = q * kQuad;
        const int nvalid = min(kQuad, D - d0);
        const bool valid = g < nvalid;
        const int d = d0 + (valid ? g : 0);
        const int kd = k * D + d;
        __syncthreads();
        quad_store<true>(qr, xN, xT, gN, gT, sdt, tpos, nvalid, L, tid);
        if (q + 1 < q_end)                             // the next quad's rows fly while this one is computed
            quad_load<T, TDO, true>(qr, xsrc, gsrc, dsrc, (int64_t)b * D + d0 + kQuad, (int64_t)b * 4 * D + d0 + kQuad, D,
                                    min(kQuad, D - d0 - kQuad), L, tid);
        __syncthreads();

        const float bias = p.bias ? p.bias[kd] : 0.0f;
        const float Dd = p.Ds ? p.Ds[kd] : 0.0f;
        float du[8];
        auto walk = [&](auto rev_tag) __attribute__((always_inline)) {
        constexpr bool R = decltype(rev_tag)::value;
        float dt[8], u[8], dy[8], sig[8], ddt[8];
        lds8_lin(sdt + (k * kQuad + g) * kSmallL, p0, dt);
        lds8((transposed ? xT : xN) + g * kSmallL, f4s, u);
        lds8((transposed ? gT : gN) + g * kSmallL, f4s, dy);
        float dD_acc = 0.0f;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float xx = dt[i] + bias;
            float e = 0.0f;
            const float sp = p.softplus ? softplus_fwd(xx, e) : xx;
            sig[i] = p.softplus ? ((xx > 20.0f) ? 1.0f : e * rcp(1.0f + e)) : 1.0f;
            dt[i] = (p0 + i < L) ? sp : 0.0f;
            du[i] = Dd * dy[i];
            ddt[i] = 0.0f;
            dD_acc = fmaf(dy[i], u[i], dD_acc);
        }
        const float* Arow = p.A + (int64_t)kd * N;
        for (int n = 0; n < N; ++n) {
            float Bv[8], Cv[8], a[8], bu[8], S[8], P[8], Sq[8], Pq[8];
            lds8_lin(myB + n * kSmallL, p0, Bv);
            lds8_lin(myC + n * kSmallL, p0, Cv);
            const float An = Arow[n];
            const float A2 = An * kLog2e;
            float Pr = 1.0f, Sr = 0.0f;
#pragma unroll
            for (int ii = 0; ii < 8; ++ii) {
                const int i = R ? 7 - ii : ii;
                a[i] = ex2(dt[i] * A2);
                bu[i] = (dt[i] * Bv[i]) * u[i];
                Sr = fmaf(a[i], Sr, bu[i]); Pr *= a[i]; S[i] = Sr; P[i] = Pr;
            }
            float unused;
            const float h_in = group_prefix<R>(Pr, Sr, j, unused);
            Pr = 1.0f; Sr = 0.0f;
#pragma unroll
            for (int ii = 0; ii < 8; ++ii) {                    // adjoint walks the other way
                const int i = R ? ii : 7 - ii;
                Sr = a[i] * fmaf(Cv[i], dy[i], Sr); Pr *= a[i]; Sq[i] = Sr; Pq[i] = Pr;
            }
            const float q_in = group_prefix<!R>(Pr, Sr, j, unused);
            float dA_part = 0.0f, dBv[8], dCv[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float h = fmaf(P[i], h_in, S[i]);
                const int inx = R ? (i == 0 ? 0 : i - 1) : (i == 7 ? 7 : i + 1);
                const bool edge = R ? (i == 0) : (i == 7);
                const float q_next = edge ? q_in : fmaf(Pq[inx], q_in, Sq[inx]);
                const float gi = fmaf(Cv[i], dy[i], q_next);
                const float hp = h - bu[i];
                const float gdt = gi * dt[i];
                du[i] = fmaf(gdt, Bv[i], du[i]);
                ddt[i] = fmaf(gi, fmaf(Bv[i], u[i], An * hp), ddt[i]);
                dA_part = fmaf(gdt, hp, dA_part);
                dBv[i] = quad_sum(valid ? gdt * u[i] : 0.0f);
                dCv[i] = quad_sum(valid ? dy[i] * h : 0.0f);
            }
            if (g == 0) {       // one group accumulates the warp's four channels (position order)
                float4* rb = reinterpret_cast<float4*>(mydB + n * kSmallL + p0);
                float4* rc = reinterpret_cast<float4*>(mydC + n * kSmallL + p0);
                float4 b0 = rb[0], b1 = rb[1], c0 = rc[0], c1 = rc[1];
                b0.x += dBv[0]; b0.y += dBv[1]; b0.z += dBv[2]; b0.w += dBv[3]; b1.x += dBv[4]; b1.y += dBv[5]; b1.z += dBv[6]; b1.w += dBv[7];
                c0.x += dCv[0]; c0.y += dCv[1]; c0.z += dCv[2]; c0.w += dCv[3]; c1.x += dCv[4]; c1.y += dCv[5]; c1.z += dCv[6]; c1.w += dCv[7];
                rb[0] = b0; rb[1] = b1; rc[0] = c0; rc[1] = c1;
            }
            dA_part = group_sum(dA_part);
            if (valid && j == 0) atomicAdd(p.dA + (int64_t)kd * N + n, dA_part);
        }
        float dbias_acc = 0.0f;
#pragma unroll
        for (int i = 0; i < 8; ++i) { ddt[i] *= sig[i]; dbias_acc += ddt[i]; }
        dD_acc = group_sum(dD_acc);
        dbias_acc = group_sum(dbias_acc);
        if (valid && j == 0) {
            if (p.dDs) atomicAdd(p.dDs + kd, dD_acc);
            if (p.dbias) atomicAdd(p.dbias + kd, dbias_acc);
        }
        // ddelta back through the staging rows (every lane overwrites exactly what it read), then coalesced to global
        {
            float* row = sdt + (k * kQuad + g) * kSmallL + p0;
            *reinterpret_cast<float4*>(row) = make_float4(ddt[0], ddt[1], ddt[2], ddt[3]);
            *reinterpret_cast<float4*>(row + 4) = make_float4(ddt[4], ddt[5], ddt[6], ddt[7]);
        }
        };
        if (rev) walk(std::true_type{}); else walk(std::false_type{});
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
        merge_quad_t<T>(reinterpret_cast<T*>(p.dx[s]) + ((int64_t)b * D + d0) * L, nvalid, dN, dT, tpos, L, tid);
        T* __restrict__ ddst = reinterpret_cast<T*>(p.ddelta[s]) + (int64_t)b * 4 * D * L;
        for (int r = 0; r < 4; ++r)
            unstage_rows<T>(ddst + ((int64_t)r * D + d0) * L, sdt + r * kQuad * kSmallL, nvalid, L, r >= 2, tid, 128);
    }
    __syncthreads();
    // flush dB / dC: smem rows are in POSITION order; scan index l = p (routes 0/1) or L-1-p (routes 2/3)
    float* dBdst = p.dBs[s] + (int64_t)b * 4 * N * L;
    float* dCdst = p.dCs[s] + (int64_t)b * 4 * N * L;
    for (int row = tid >> 5; row < 4 * N; row += 4) {          // row = r * N + n; smem rows are in POSITION order
        const bool fl = row >= 2 * N;
        for (int l = lane; l < L; l += 32) {
            const int pp = fl ? L - 1 - l : l;
            atomicAdd(dBdst + row * L + l, sdB[row * kSmallL + pp]);
            atomicAdd(dCdst + row * L + l, sdC[row * kSmallL + pp]);
        }
    }
}

// =========================================================================================================
// shallow fusion: swap-gather + S6 (K = 2) + split
// =========================================================================================================
struct FsSwapFwd {
    const void* x; const void* x2; const void* delta; const void* Bs; const void* Cs;
    void* y; void* y2; float* states;
    const float* A; const float* Ds; const float* bias;
    int batch, D, N, L, softplus;
};
struct FsSwapBwd {
    const void* x; const void* x2; const void* delta; const void* Bs; const void* Cs; const void* dy; const void* dy2;
    void* dx; void* dx2; void* ddelta; float* dBs; float* dCs;
    const float* A; const float* Ds; const float* bias;
    float* dA; float* dDs; float* dbias;
    int batch, D, N, L, softplus;
};

constexpr int kFsRowsPerStep = 16;   // 4 warps x 4 lane groups
constexpr int kFsRowSteps = 8;       // steps per CTA -> 128 channels share one staging of B / C

// u of half k, channel d (SwappingScan_multiview.forward, models/fusion_vmamba.py:198-214): half 0 takes x2 on even channels
// and x on odd ones, half 1 the other way round
__device__ __forceinline__ bool swap_takes_x2(int k, int d) { return ((d & 1) == 0) == (k == 0); }

template <typename T>
__device__ __forceinline__ void load_row8(const T* __restrict__ row, int l0, int L, float (&v)[8]) {
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = (l0 + i < L) ? Elem<T>::to_f(row[l0 + i]) : 0.0f;
}

template <typename T, typename TO>
__global__ void __launch_bounds__(128)
fs_swap_fwd_kernel(const FsSwapFwd p) {
    extern __shared__ __align__(16) float sm[];
    const int L = p.L, N = p.N, D = p.D;
    float* sB = sm;                                   // [N][64]
    float* sC = sB + N * kSmallL;
    const int nblk = (D + kFsRowsPerStep * kFsRowSteps - 1) / (kFsRowsPerStep * kFsRowSteps);
    const int bk = blockIdx.x / nblk, blk = blockIdx.x - bk * nblk;       // bk = b * 2 + k
    const int b = bk >> 1, k = bk & 1;
    const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31, g = lane >> 3, j = lane & 7;
    stage_rows<T>(reinterpret_cast<const T*>(p.Bs) + (int64_t)bk * N * L, sB, N, L, false, tid, 128);
    stage_rows<T>(reinterpret_cast<const T*>(p.Cs) + (int64_t)bk * N * L, sC, N, L, false, tid, 128);
    __syncthreads();
    const int l0 = j * 8;
    for (int it = 0; it < kFsRowSteps; ++it) {
        const int dd = blk * kFsRowsPerStep * kFsRowSteps + it * kFsRowsPerStep + w * 4 + g;
        const bool valid = dd < D;                     // no early exit: the shuffles below are issued warp-wide
        const int d = valid ? dd : 0;
        const int kd = k * D + d;
        const T* __restrict__ urow = reinterpret_cast<const T*>(swap_takes_x2(k, d) ? p.x2 : p.x) + ((int64_t)b * D + d) * L;
        const T* __restrict__ drow = reinterpret_cast<const T*>(p.delta) + ((int64_t)b * 2 * D + kd) * L;
        const float bias = p.bias ? p.bias[kd] : 0.0f;
        const float Dd = p.Ds ? p.Ds[kd] : 0.0f;
        float dt[8], u[8], y[8], dtu[8];
        load_row8<T>(drow, l0, L, dt);
        load_row8<T>(urow, l0, L, u);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float xx = dt[i] + bias;
            float e;
            const float sp = p.softplus ? softplus_fwd(xx, e) : xx;
            dt[i] = (l0 + i < L) ? sp : 0.0f;
            dtu[i] = dt[i] * u[i];
            y[i] = Dd * u[i];
        }
        const float* Arow = p.A + (int64_t)kd * N;
        for (int n = 0; n < N; ++n) {
            float Bv[8], Cv[8], S[8], P[8];
            lds8_lin(sB + n * kSmallL, l0, Bv);
            lds8_lin(sC + n * kSmallL, l0, Cv);
            const float A2 = Arow[n] * kLog2e;
            float Pr = 1.0f, Sr = 0.0f;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float a = ex2(dt[i] * A2);
                Sr = fmaf(a, Sr, dtu[i] * Bv[i]);
                Pr *= a;
                S[i] = Sr; P[i] = Pr;
            }
            float h_end;
            const float h_in = group_prefix<false>(Pr, Sr, j, h_end);
#pragma unroll
            for (int i = 0; i < 8; ++i) y[i] = fmaf(Cv[i], fmaf(P[i], h_in, S[i]), y[i]);
            if (p.states && valid && j == 0) p.states[((int64_t)b * 2 * D + kd) * N + n] = h_end;
        }
        TO* __restrict__ orow = reinterpret_cast<TO*>(k == 0 ? p.y : p.y2) + ((int64_t)b * D + d) * L;   // SwappingMerge: half k -> view k
#pragma unroll
        for (int i = 0; i < 8; ++i)
            if (valid && l0 + i < L) orow[l0 + i] = Elem<TO>::from_f(y[i]);
    }
}

template <typename T, typename TDO>
__global__ void __launch_bounds__(128)
fs_swap_bwd_kernel(const FsSwapBwd p) {
    extern __shared__ __align__(16) float sm[];
    const int L = p.L, N = p.N, D = p.D;
    float* sB = sm;                                   // [N][64]
    float* sC = sB + N * kSmallL;
    float* sdB = sC + N * kSmallL;                    // [4 warps][N][64]
    float* sdC = sdB + 4 * N * kSmallL;
    const int nblk = (D + kFsRowsPerStep * kFsRowSteps - 1) / (kFsRowsPerStep * kFsRowSteps);
    const int bk = blockIdx.x / nblk, blk = blockIdx.x - bk * nblk;
    const int b = bk >> 1, k = bk & 1;
    const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31, g = lane >> 3, j = lane & 7;
    stage_rows<T>(reinterpret_cast<const T*>(p.Bs) + (int64_t)bk * N * L, sB, N, L, false, tid, 128);
    stage_rows<T>(reinterpret_cast<const T*>(p.Cs) + (int64_t)bk * N * L, sC, N, L, false, tid, 128);
    for (int i = tid; i < 8 * N * kSmallL; i += 128) sdB[i] = 0.0f;
    __syncthreads();
    float* mydB = sdB + w * N * kSmallL;
    float* mydC = sdC + w * N * kSmallL;
    const int l0 = j * 8;
    for (int it = 0; it < kFsRowSteps; ++it) {
        const int dd = blk * kFsRowsPerStep * kFsRowSteps + it * kFsRowsPerStep + w * 4 + g;
        const bool valid = dd < D;                     // (the quad_sum below spans the warp: invalid groups contribute zeros)
        const int d = valid ? dd : 0;
        const int kd = k * D + d;
        const T* __restrict__ urow = reinterpret_cast<const T*>(swap_takes_x2(k, d) ? p.x2 : p.x) + ((int64_t)b * D + d) * L;
        const T* __restrict__ drow = reinterpret_cast<const T*>(p.delta) + ((int64_t)b * 2 * D + kd) * L;
        const TDO* __restrict__ gyrow = reinterpret_cast<const TDO*>(k == 0 ? p.dy : p.dy2) + ((int64_t)b * D + d) * L;
        const float bias = p.bias ? p.bias[kd] : 0.0f;
        const float Dd = p.Ds ? p.Ds[kd] : 0.0f;
        float dt[8], u[8], dy[8], sig[8], du[8], ddt[8];
        load_row8<T>(drow, l0, valid ? L : 0, dt);
        load_row8<T>(urow, l0, valid ? L : 0, u);
        load_row8<TDO>(gyrow, l0, valid ? L : 0, dy);
        float dD_acc = 0.0f;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float xx = dt[i] + bias;
            float e = 0.0f;
            const float sp = p.softplus ? softplus_fwd(xx, e) : xx;
            sig[i] = p.softplus ? ((xx > 20.0f) ? 1.0f : e * rcp(1.0f + e)) : 1.0f;
            dt[i] = (valid && l0 + i < L) ? sp : 0.0f;
            du[i] = Dd * dy[i];
            ddt[i] = 0.0f;
            dD_acc = fmaf(dy[i], u[i], dD_acc);
        }
        const float* Arow = p.A + (int64_t)kd * N;
        for (int n = 0; n < N; ++n) {
            float Bv[8], Cv[8], a[8], bu[8], S[8], P[8], Sq[8], Pq[8];
            lds8_lin(sB + n * kSmallL, l0, Bv);
            lds8_lin(sC + n * kSmallL, l0, Cv);
            const float An = Arow[n];
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
                float4* rb = reinterpret_cast<float4*>(mydB + n * kSmallL + l0);
                float4* rc = reinterpret_cast<float4*>(mydC + n * kSmallL + l0);
                float4 b0 = rb[0], b1 = rb[1], c0 = rc[0], c1 = rc[1];
                b0.x += dBv[0]; b0.y += dBv[1]; b0.z += dBv[2]; b0.w += dBv[3]; b1.x += dBv[4]; b1.y += dBv[5]; b1.z += dBv[6]; b1.w += dBv[7];
                c0.x += dCv[0]; c0.y += dCv[1]; c0.z += dCv[2]; c0.w += dCv[3]; c1.x += dCv[4]; c1.y += dCv[5]; c1.z += dCv[6]; c1.w += dCv[7];
                rb[0] = b0; rb[1] = b1; rc[0] = c0; rc[1] = c1;
            }
            dA_part = group_sum(dA_part);
            if (valid && j == 0) atomicAdd(p.dA + (int64_t)kd * N + n, dA_part);
        }
        float dbias_acc = 0.0f;
#pragma unroll
        for (int i = 0; i < 8; ++i) { ddt[i] *= sig[i]; dbias_acc += ddt[i]; }
        dD_acc = group_sum(dD_acc);
        dbias_acc = group_sum(dbias_acc);
        if (valid) {
            if (j == 0) {
                if (p.dDs) atomicAdd(p.dDs + kd, dD_acc);
                if (p.dbias) atomicAdd(p.dbias + kd, dbias_acc);
            }
            // SwappingScan_multiview.backward AS WRITTEN (models/fusion_vmamba.py:217-221): half 0 -> dx, half 1 -> dx2, no un-swap
            T* __restrict__ durow = reinterpret_cast<T*>(k == 0 ? p.dx : p.dx2) + ((int64_t)b * D + d) * L;
            T* __restrict__ ddrow = reinterpret_cast<T*>(p.ddelta) + ((int64_t)b * 2 * D + kd) * L;
#pragma unroll
            for (int i = 0; i < 8; ++i)
                if (l0 + i < L) { durow[l0 + i] = Elem<T>::from_f(du[i]); ddrow[l0 + i] = Elem<T>::from_f(ddt[i]); }
        }
    }
    __syncthreads();
    for (int idx = tid; idx < N * L; idx += 128) {
        const int n = idx / L, l = idx - n * L;
        const int o = n * kSmallL + l;
        const float vb = sdB[o] + sdB[N * kSmallL + o] + sdB[2 * N * kSmallL + o] + sdB[3 * N * kSmallL + o];
        const float vc = sdC[o] + sdC[N * kSmallL + o] + sdC[2 * N * kSmallL + o] + sdC[3 * N * kSmallL + o];
        atomicAdd(p.dBs + (int64_t)bk * N * L + idx, vb);
        atomicAdd(p.dCs + (int64_t)bk * N * L + idx, vc);
    }
}

// ---- host side --------------------------------------------------------------------------------------------------
int fusion_small_supported(int64_t N, int64_t L) { return L >= 1 && L <= kSmallL && N >= 1 && N <= kSmallMaxN; }

template <typename T, typename TO>
static int launch_x3_fwd_t(const FsX3Fwd& a, int nstreams, cudaStream_t st) {
    const int nquads = (a.D + kQuad - 1) / kQuad;
    const dim3 grid((unsigned)(a.batch * ((nquads + kFsQuads - 1) / kFsQuads)), (unsigned)nstreams);
    const size_t smem = sizeof(float) * (size_t)(4 * kQuad * kSmallL + 4 * kQuad * kSmallL + 2 * 4 * a.N * kSmallL + kSmallL);
    if (int rc = set_smem(fs_x3_fwd_kernel<T, TO>, smem)) return rc;
    fs_x3_fwd_kernel<T, TO><<<grid, 128, smem, st>>>(a);
    return check_launch();
}
template <typename T, typename TDO>
static int launch_x3_bwd_t(const FsX3Bwd& a, int nstreams, cudaStream_t st) {
    const int nquads = (a.D + kQuad - 1) / kQuad;
    const dim3 grid((unsigned)(a.batch * ((nquads + kFsQuads - 1) / kFsQuads)), (unsigned)nstreams);
    const size_t smem = sizeof(float) * (size_t)(6 * kQuad * kSmallL + 4 * kQuad * kSmallL + 4 * 4 * a.N * kSmallL + kSmallL);
    if (int rc = set_smem(fs_x3_bwd_kernel<T, TDO>, smem)) return rc;
    fs_x3_bwd_kernel<T, TDO><<<grid, 128, smem, st>>>(a);
    return check_launch();
}

int launch_cross_ss2d_x3_fwd(const xfs_cross_ss2d_x3_fwd_args& a, cudaStream_t st) {
    FsX3Fwd k{};
    for (int s = 0; s < 3; ++s) {
        k.x[s] = a.x[s]; k.delta[s] = a.delta[s]; k.Bs[s] = a.Bs[s]; k.Cs[s] = a.Cs[s]; k.y[s] = a.y[s]; k.states[s] = a.states[s];
    }
    k.A = a.A; k.Ds = a.Ds; k.bias = a.delta_bias;
    k.batch = (int)a.batch; k.D = (int)a.D; k.N = (int)a.N; k.H = (int)a.H; k.W = (int)a.W; k.softplus = a.delta_softplus;
    const bool o32 = a.out_dtype == XFS_F32;
    const int ns = a.nstreams;
    switch (a.dtype) {
        case XFS_F32: return launch_x3_fwd_t<float, float>(k, ns, st);
        case XFS_BF16: return o32 ? launch_x3_fwd_t<__nv_bfloat16, float>(k, ns, st) : launch_x3_fwd_t<__nv_bfloat16, __nv_bfloat16>(k, ns, st);
        default: return o32 ? launch_x3_fwd_t<__half, float>(k, ns, st) : launch_x3_fwd_t<__half, __half>(k, ns, st);
    }
}

int launch_cross_ss2d_x3_bwd(const xfs_cross_ss2d_x3_bwd_args& a, cudaStream_t st) {
    FsX3Bwd k{};
    for (int s = 0; s < 3; ++s) {
        k.x[s] = a.x[s]; k.delta[s] = a.delta[s]; k.Bs[s] = a.Bs[s]; k.Cs[s] = a.Cs[s]; k.dy[s] = a.dy[s];
        k.dx[s] = a.dx[s]; k.ddelta[s] = a.ddelta[s]; k.dBs[s] = a.dBs[s]; k.dCs[s] = a.dCs[s];
    }
    k.A = a.A; k.Ds = a.Ds; k.bias = a.delta_bias; k.dA = a.dA; k.dDs = a.dDs; k.dbias = a.ddelta_bias;
    k.batch = (int)a.batch; k.D = (int)a.D; k.N = (int)a.N; k.H = (int)a.H; k.W = (int)a.W; k.softplus = a.delta_softplus;
    const bool g32 = a.dout_dtype == XFS_F32;
    const int ns = a.nstreams;
    switch (a.dtype) {
        case XFS_F32: return launch_x3_bwd_t<float, float>(k, ns, st);
        case XFS_BF16: return g32 ? launch_x3_bwd_t<__nv_bfloat16, float>(k, ns, st) : launch_x3_bwd_t<__nv_bfloat16, __nv_bfloat16>(k, ns, st);
        default: return g32 ? launch_x3_bwd_t<__half, float>(k, ns, st) : launch_x3_bwd_t<__half, __half>(k, ns, st);
    }
}

template <typename T, typename TO>
static int launch_swap_fwd_t(const FsSwapFwd& a, cudaStream_t st) {
    const int nblk = (a.D + kFsRowsPerStep * kFsRowSteps - 1) / (kFsRowsPerStep * kFsRowSteps);
    const size_t smem = sizeof(float) * (size_t)(2 * a.N * kSmallL);
    if (int rc = set_smem(fs_swap_fwd_kernel<T, TO>, smem)) return rc;
    fs_swap_fwd_kernel<T, TO><<<(unsigned)(a.batch * 2 * nblk), 128, smem, st>>>(a);
    return check_launch();
}
template <typename T, typename TDO>
static int launch_swap_bwd_t(const FsSwapBwd& a, cudaStream_t st) {
    const int nblk = (a.D + kFsRowsPerStep * kFsRowSteps - 1) / (kFsRowsPerStep * kFsRowSteps);
    const size_t smem = sizeof(float) * (size_t)(10 * a.N * kSmallL);
    if (int rc = set_smem(fs_swap_bwd_kernel<T, TDO>, smem)) return rc;
    fs_swap_bwd_kernel<T, TDO><<<(unsigned)(a.batch * 2 * nblk), 128, smem, st>>>(a);
    return check_launch();
}

int launch_swap_scan_fused_fwd(const xfs_swap_scan_fused_fwd_args& a, cudaStream_t st) {
    FsSwapFwd k{a.x, a.x2, a.delta, a.Bs, a.Cs, a.y, a.y2, a.states, a.A, a.Ds, a.delta_bias,
                (int)a.batch, (int)a.D, (int)a.N, (int)a.L, (int)a.delta_softplus};
    const bool o32 = a.out_dtype == XFS_F32;
    switch (a.dtype) {
        case XFS_F32: return launch_swap_fwd_t<float, float>(k, st);
        case XFS_BF16: return o32 ? launch_swap_fwd_t<__nv_bfloat16, float>(k, st) : launch_swap_fwd_t<__nv_bfloat16, __nv_bfloat16>(k, st);
        default: return o32 ? launch_swap_fwd_t<__half, float>(k, st) : launch_swap_fwd_t<__half, __half>(k, st);
    }
}

int launch_swap_scan_fused_bwd(const xfs_swap_scan_fused_bwd_args& a, cudaStream_t st) {
    FsSwapBwd k{a.x, a.x2, a.delta, a.Bs, a.Cs, a.dy, a.dy2, a.dx, a.dx2, a.ddelta, a.dBs, a.dCs, a.A, a.Ds, a.delta_bias,
                a.dA, a.dDs, a.ddelta_bias, (int)a.batch, (int)a.D, (int)a.N, (int)a.L, (int)a.delta_softplus};
    const bool g32 = a.dout_dtype == XFS_F32;
    switch (a.dtype) {
        case XFS_F32: return launch_swap_bwd_t<float, float>(k, st);
        case XFS_BF16: return g32 ? launch_swap_bwd_t<__nv_bfloat16, float>(k, st) : launch_swap_bwd_t<__nv_bfloat16, __nv_bfloat16>(k, st);
        default: return g32 ? launch_swap_bwd_t<__half, float>(k, st) : launch_swap_bwd_t<__half, __half>(k, st);
    }
}

}  // namespace xfs
