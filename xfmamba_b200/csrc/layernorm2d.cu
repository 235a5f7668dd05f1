// layernorm2d.cu -- channel-first LayerNorm over C for (B, C, H*W) tensors: the consumer of the merged scan output
// (out_norm = LayerNorm2d(d_inner), reference models/fusion_vmamba.py:52-57, 1183-1188) and every other LayerNorm2d of
// the channel-first backbone.  The reference permutes to channel-last, which makes torch copy the tensor twice around the
// normalisation; here the tensor is read where it lies: a CTA owns 32 consecutive positions (one 128-byte line per
// channel row), 8 warps stride over the channels, statistics are combined through shared memory.
// HBM-bound: forward 2 reads (second one from L2) + 1 write; backward x, dy read twice (second from L2) + 1 write.
#include "xfscan_common.cuh"

namespace xfs {

constexpr int kLnWarps = 8;
constexpr int kLnPos = 32;

template <typename T>
__global__ void __launch_bounds__(kLnWarps * 32)
ln2d_fwd_kernel(const T* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias, T* __restrict__ y,
                float* __restrict__ mean_out, float* __restrict__ rstd_out, int C, int HW, float eps) {
    __shared__ float s_a[kLnWarps][kLnPos], s_b[kLnWarps][kLnPos];
    const int tiles = (HW + kLnPos - 1) / kLnPos;
    const int b = blockIdx.x / tiles, t = blockIdx.x - b * tiles;
    const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
    const int pos = t * kLnPos + lane;
    const bool ok = pos < HW;
    const T* __restrict__ xb = x + (int64_t)b * C * HW + pos;
    // shifted sums (shift = first channel) keep E[x^2] - E[x]^2 well conditioned
    const float shift = ok ? Elem<T>::to_f(xb[0]) : 0.0f;
    float s1 = 0.0f, s2 = 0.0f;
    if (ok)
        for (int c = wp; c < C; c += kLnWarps) {
            const float v = Elem<T>::to_f(xb[(int64_t)c * HW]) - shift;
            s1 += v;
            s2 = fmaf(v, v, s2);
        }
    s_a[wp][lane] = s1;
    s_b[wp][lane] = s2;
    __syncthreads();
    s1 = 0.0f; s2 = 0.0f;
#pragma unroll
    for (int i = 0; i < kLnWarps; ++i) { s1 += s_a[i][lane]; s2 += s_b[i][lane]; }
    const float inv = 1.0f / (float)C;
    const float m = s1 * inv;
    const float var = fmaxf(fmaf(-m, m, s2 * inv), 0.0f);
    const float rstd = rsqrtf(var + eps);
    const float mean = m + shift;
    if (ok) {
        if (wp == 0 && mean_out) { mean_out[(int64_t)b * HW + pos] = mean; rstd_out[(int64_t)b * HW + pos] = rstd; }
        T* __restrict__ yb = y + (int64_t)b * C * HW + pos;
        for (int c = wp; c < C; c += kLnWarps) {
            const float v = (Elem<T>::to_f(xb[(int64_t)c * HW]) - mean) * rstd;
            yb[(int64_t)c * HW] = Elem<T>::from_f(fmaf(v, w ? w[c] : 1.0f, bias ? bias[c] : 0.0f));
        }
    }
}

template <typename T>
__global__ void __launch_bounds__(kLnWarps * 32)
ln2d_bwd_kernel(const T* __restrict__ x, const T* __restrict__ dy, const float* __restrict__ w, const float* __restrict__ mean_in,
                const float* __restrict__ rstd_in, T* __restrict__ dx, float* __restrict__ dw, float* __restrict__ db, int C,
                int HW) {
    __shared__ float s_a[kLnWarps][kLnPos], s_b[kLnWarps][kLnPos];
    const int tiles = (HW + kLnPos - 1) / kLnPos;
    const int b = blockIdx.x / tiles, t = blockIdx.x - b * tiles;
    const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
    const int pos = t * kLnPos + lane;
    const bool ok = pos < HW;
    const int64_t base = (int64_t)b * C * HW + pos;
    const float mean = ok ? mean_in[(int64_t)b * HW + pos] : 0.0f;
    const float rstd = ok ? rstd_in[(int64_t)b * HW + pos] : 0.0f;
    float s1 = 0.0f, s2 = 0.0f;
    for (int c = wp; c < C; c += kLnWarps) {
        float g = 0.0f, xh = 0.0f;
        if (ok) {
            g = Elem<T>::to_f(dy[base + (int64_t)c * HW]);
            xh = (Elem<T>::to_f(x[base + (int64_t)c * HW]) - mean) * rstd;
        }
        const float gw = g * (w ? w[c] : 1.0f);
        s1 += gw;
        s2 = fmaf(gw, xh, s2);
        // parameter gradients: reduce this channel over the 32 positions of the tile
        float pw = g * xh, pb = g;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            pw += __shfl_xor_sync(kFull, pw, off);
            pb += __shfl_xor_sync(kFull, pb, off);
        }
        if (lane == 0) {
            if (dw) atomicAdd(dw + c, pw);
            if (db) atomicAdd(db + c, pb);
        }
    }
    s_a[wp][lane] = s1;
    s_b[wp][lane] = s2;
    __syncthreads();
    s1 = 0.0f; s2 = 0.0f;
#pragma unroll
    for (int i = 0; i < kLnWarps; ++i) { s1 += s_a[i][lane]; s2 += s_b[i][lane]; }
    const float inv = 1.0f / (float)C;
    s1 *= inv; s2 *= inv;
    if (ok)
        for (int c = wp; c < C; c += kLnWarps) {
            const float g = Elem<T>::to_f(dy[base + (int64_t)c * HW]);
            const float xh = (Elem<T>::to_f(x[base + (int64_t)c * HW]) - mean) * rstd;
            dx[base + (int64_t)c * HW] = Elem<T>::from_f(rstd * (g * (w ? w[c] : 1.0f) - s1 - xh * s2));
        }
}

int launch_ln2d_fwd(const void* x, const float* w, const float* b, void* y, float* mean, float* rstd, int64_t B, int64_t C,
                    int64_t HW, float eps, int dtype, cudaStream_t st) {
    const unsigned grid = (unsigned)(B * ((HW + kLnPos - 1) / kLnPos));
    if (dtype == XFS_F32) ln2d_fwd_kernel<float><<<grid, kLnWarps * 32, 0, st>>>((const float*)x, w, b, (float*)y, mean, rstd, (int)C, (int)HW, eps);
    else if (dtype == XFS_BF16) ln2d_fwd_kernel<__nv_bfloat16><<<grid, kLnWarps * 32, 0, st>>>((const __nv_bfloat16*)x, w, b, (__nv_bfloat16*)y, mean, rstd, (int)C, (int)HW, eps);
    else ln2d_fwd_kernel<__half><<<grid, kLnWarps * 32, 0, st>>>((const __half*)x, w, b, (__half*)y, mean, rstd, (int)C, (int)HW, eps);
    return check_launch();
}

int launch_ln2d_bwd(const void* x, const void* dy, const float* w, const float* mean, const float* rstd, void* dx, float* dw,
                    float* db, int64_t B, int64_t C, int64_t HW, int dtype, cudaStream_t st) {
    const unsigned grid = (unsigned)(B * ((HW + kLnPos - 1) / kLnPos));
    if (dtype == XFS_F32) ln2d_bwd_kernel<float><<<grid, kLnWarps * 32, 0, st>>>((const float*)x, (const float*)dy, w, mean, rstd, (float*)dx, dw, db, (int)C, (int)HW);
    else if (dtype == XFS_BF16) ln2d_bwd_kernel<__nv_bfloat16><<<grid, kLnWarps * 32, 0, st>>>((const __nv_bfloat16*)x, (const __nv_bfloat16*)dy, w, mean, rstd, (__nv_bfloat16*)dx, dw, db, (int)C, (int)HW);
    else ln2d_bwd_kernel<__half><<<grid, kLnWarps * 32, 0, st>>>((const __half*)x, (const __half*)dy, w, mean, rstd, (__half*)dx, dw, db, (int)C, (int)HW);
    return check_launch();
}

}  // namespace xfs
