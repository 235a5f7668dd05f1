/*
 * xfscan.h -- C ABI of libxfscan.so: B200 (sm_100a) kernels for XFMamba's SS2D scan path.
 *
 * This is the drop-in boundary.  The reference reaches its native scan through a pybind11 module whose
 * entry points are
 *     fwd(u, delta, A, B, C, D?, delta_bias?, delta_softplus, nrows[, oflex]) -> [out, x]
 *     bwd(u, delta, A, B, C, D?, delta_bias?, dout, x?, delta_softplus, nrows) -> [du, ddelta, dA, dB, dC, dD, ddelta_bias]
 * (reference: models/selective_scan/csrc/selective_scan/selective_scan.cpp:165-172, 251-260, 364-367; call sites
 * models/csms6s.py:81-85, 97-108), and its cross-scan/merge through Triton launches
 * (models/csm_triton.py:403-497) and torch index ops (models/fusion_vmamba.py:189-241).
 * Each function below replaces one of those; the comment above it cites the interface it replaces.
 *
 * Conventions
 *   - plain C, no torch / ATen types; every pointer is a DEVICE pointer owned by the caller
 *   - the library never allocates, never synchronises, keeps no global state; work is enqueued on `stream`
 *   - all tensors are dense row-major ("contiguous") with the shapes written at each function
 *   - index arithmetic is 64-bit (the reference's is uint32: selective_scan.h:27)
 *   - return value: 0 = success; >0 = cudaError_t of the launch; <0 = argument error (xfs_error_string)
 *   - re-entrant and CUDA-graph capturable (no allocation / sync / host state)
 */
#ifndef XFSCAN_H_
#define XFSCAN_H_

#include <stdint.h>

#if defined(__GNUC__)
#define XFS_API __attribute__((visibility("default")))
#else
#define XFS_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

typedef void* xfs_stream_t; /* cudaStream_t */

/* element types of u / delta / B / C / x (A, D, delta_bias, chunk states and dB/dC accumulators are always f32;
 * same dtype rules as selective_scan.cpp:175-180) */
enum { XFS_F32 = 0, XFS_BF16 = 1, XFS_F16 = 2 };

/* scan route sets (models/csm_triton.py:25-35): 0 = cross2d (4 routes), 1 = unidirectional, 2 = bidirectional */
enum { XFS_SCANS_CROSS2D = 0, XFS_SCANS_UNIDI = 1, XFS_SCANS_BIDI = 2 };

enum {
    XFS_OK = 0,
    XFS_ERR_NULL = -1,        /* required pointer is NULL */
    XFS_ERR_SHAPE = -2,       /* non-positive size, dim % ngroups != 0, dstate > 256, ... */
    XFS_ERR_DTYPE = -3,       /* unknown dtype enum / unsupported combination */
    XFS_ERR_ALIGN = -4,       /* base pointer not 16-byte aligned */
    XFS_ERR_UNSUPPORTED = -5, /* valid request this build has no kernel for (e.g. fused path with L too large) */
    XFS_ERR_ARCH = -6         /* device is not sm_100 */
};

XFS_API int xfs_version(void);                    /* ABI version, bumps on incompatible change */
XFS_API const char* xfs_error_string(int code);   /* static string for negative codes; cudaGetErrorString for positive */
XFS_API int xfs_device_ok(int device);            /* 0 if `device` is compute capability 10.x, XFS_ERR_ARCH otherwise */

/* length of one scan chunk: the forward kernels store the recurrent state at the end of every chunk so that the
 * backward can restart from it (same role as `x` of shape (B, dim, n_chunks, 2N) in selective_scan.cpp:225-228,
 * but only h is kept: (B, dim, n_chunks, N) f32). */
XFS_API int64_t xfs_chunk_len(void);
XFS_API int64_t xfs_num_chunks(int64_t seqlen);

/* -------------------------------------------------------------------------------------------------------------
 * CrossScan / CrossMerge, channel-first.  Replaces CrossScanF / CrossScanTritonF and CrossMergeF /
 * CrossMergeTritonF (models/csm_triton.py:182-273, 403-497).
 *   cross_scan : x  (B, C, H, W)  [one_by_one: (B, 4, C, H, W)]  ->  xs (B, 4, C, H*W)      pure permutation
 *   cross_merge: ys (B, 4, C, H*W) -> y (B, C, H*W), y = (ys0 + flip(ys2)) + T(ys1 + flip(ys3))
 *                [one_by_one: -> (B, 4, C, H*W), un-routing only]            add order of csm_triton.py:61-62
 * The backward of each is the other (csm_triton.py:208-225, 249-273).
 * ----------------------------------------------------------------------------------------------------------- */
XFS_API int xfs_cross_scan(const void* x, void* xs, int64_t B, int64_t C, int64_t H, int64_t W, int dtype, int scans,
                   int one_by_one, xfs_stream_t stream);
XFS_API int xfs_cross_merge(const void* ys, void* y, int64_t B, int64_t C, int64_t H, int64_t W, int dtype, int scans,
                    int one_by_one, xfs_stream_t stream);

/* -------------------------------------------------------------------------------------------------------------
 * SwappingScan_multiview / SwappingMerge_multiview (models/fusion_vmamba.py:189-241).
 *   swap_scan : x, x2 (B, C, L) -> out (B, 2, C, L); out[:,0,c] = even(c) ? x2[:,c] : x[:,c]; out[:,1,c] = the other
 *   swap_merge: ys (B, 2, C, L) -> y, y2 (B, C, L)  (contiguous split)
 *   swap_stack: y, y2 (B, C, L) -> ys (B, 2, C, L)  (backward of swap_merge as written, :234-241)
 * ----------------------------------------------------------------------------------------------------------- */
XFS_API int xfs_swap_scan(const void* x, const void* x2, void* out, int64_t B, int64_t C, int64_t L, int dtype,
                  xfs_stream_t stream);
XFS_API int xfs_swap_merge(const void* ys, void* y, void* y2, int64_t B, int64_t C, int64_t L, int dtype, xfs_stream_t stream);
XFS_API int xfs_swap_stack(const void* y, const void* y2, void* ys, int64_t B, int64_t C, int64_t L, int dtype,
                   xfs_stream_t stream);

/* -------------------------------------------------------------------------------------------------------------
 * Selective scan (S6).  Replaces selective_scan_cuda_oflex.fwd / .bwd (ABI quoted at the top; semantics of
 * models/csms6s.py:25-68).
 *   u, delta : (batch, dim, seqlen)  dtype                 A : (dim, dstate) f32
 *   B, C     : (batch, ngroups, dstate, seqlen) dtype      D, delta_bias : (dim) f32 or NULL
 *   out      : (batch, dim, seqlen)  out_dtype (f32 when oflex, else dtype)
 *   states   : (batch, dim, xfs_num_chunks(seqlen), dstate) f32, written by fwd, read by bwd; may be NULL in fwd
 *              (inference) -- bwd then recomputes them itself into `states` which must still be provided
 * bwd outputs: du, ddelta (batch, dim, seqlen) dtype;  dA (dim, dstate), dD, ddelta_bias (dim) f32, ACCUMULATED
 *   into (caller zero-fills, as selective_scan.cpp:331-337 does);  dB, dC (batch, ngroups, dstate, seqlen) f32,
 *   accumulated into (caller zero-fills and casts to dtype afterwards, selective_scan.cpp:332-333, 360).
 * ----------------------------------------------------------------------------------------------------------- */
typedef struct {
    const void* u;
    const void* delta;
    const float* A;
    const void* B;
    const void* C;
    const float* D;          /* nullable */
    const float* delta_bias; /* nullable */
    void* out;
    float* states;           /* nullable */
    int64_t batch, dim, dstate, seqlen, ngroups;
    int32_t dtype, out_dtype, delta_softplus, reserved;
} xfs_scan_fwd_args;

typedef struct {
    const void* u;
    const void* delta;
    const float* A;
    const void* B;
    const void* C;
    const float* D;          /* nullable */
    const float* delta_bias; /* nullable */
    const void* dout;        /* (batch, dim, seqlen) dout_dtype */
    const float* states;     /* from fwd */
    void* du;
    void* ddelta;
    float* dA;
    float* dB;
    float* dC;
    float* dD;               /* nullable iff D is */
    float* ddelta_bias;      /* nullable iff delta_bias is */
    int64_t batch, dim, dstate, seqlen, ngroups;
    int32_t dtype, dout_dtype, delta_softplus;
    int32_t acc_replicas;    /* R > 1: dB / dC are (R, batch, ngroups, dstate, seqlen), channel d adds into copy d % R and the
                                caller sums over R (see xfs_ss2d_bwd_args.acc_replicas); 0 or 1: plain */
} xfs_scan_bwd_args;

XFS_API int xfs_selective_scan_fwd(const xfs_scan_fwd_args* a, xfs_stream_t stream);
XFS_API int xfs_selective_scan_bwd(const xfs_scan_bwd_args* a, xfs_stream_t stream);

/* -------------------------------------------------------------------------------------------------------------
 * Fused SS2D core: y = cross_merge(selective_scan(cross_scan(x), delta, A, Bs, Cs, Ds, delta_bias)) in ONE kernel
 * (forward) and its gradient in ONE kernel (backward).  Replaces the three-operator sequence of
 * SS2Dv2.forward_corev2 (models/fusion_vmamba.py:1145, 1170-1174; sibling models/vmamba.py:603, 628-632) and of
 * each of Cross_SS2Dv5's three streams (models/fusion_vmamba.py:485, 508-512).
 *   x      : (batch, D, H, W) dtype                        delta : (batch, 4*D, H*W) dtype, in scan order of each route
 *   A      : (4*D, N) f32                                  Bs,Cs : (batch, 4, N, H*W) dtype, in scan order
 *   Ds, delta_bias : (4*D) f32 or NULL                     y     : (batch, D, H*W) out_dtype, spatial order
 *   states : (batch, 4*D, xfs_ss2d_states_len(N, H, W, dtype, out_dtype)) f32 (nullable in fwd): checkpoints of the
 *            recurrence written by fwd and read by bwd.  One state per 256-position chunk and n (xfs_num_chunks(H*W) * N
 *            floats per row) in general; one state per LANE and chunk (32 per chunk) on the path that serves the
 *            backbone shapes (f32 rows, N == 1, H*W % 4 == 0, H*W > 256) -- there every tensor must be 16-byte
 *            aligned (XFS_ERR_ALIGN otherwise).  The layout is private to the fwd / bwd pair.
 * bwd: dy (batch, D, H*W) dout_dtype -> dx (batch, D, H, W) dtype, ddelta (batch, 4*D, H*W) dtype, and the
 *   accumulated f32 dA, dBs, dCs, dDs, ddelta_bias as for xfs_selective_scan_bwd.
 * Returns XFS_ERR_UNSUPPORTED when the per-channel working set does not fit in shared memory
 * (xfs_ss2d_supported tells in advance); callers then compose the three stand-alone operators.
 * ----------------------------------------------------------------------------------------------------------- */
typedef struct {
    const void* x;
    const void* delta;
    const float* A;
    const void* Bs;
    const void* Cs;
    const float* Ds;         /* nullable */
    const float* delta_bias; /* nullable */
    void* y;
    float* states;           /* nullable */
    int64_t batch, D, N, H, W;
    int32_t dtype, out_dtype, delta_softplus, scans;
} xfs_ss2d_fwd_args;

typedef struct {
    const void* x;
    const void* delta;
    const float* A;
    const void* Bs;
    const void* Cs;
    const float* Ds;
    const float* delta_bias;
    const void* dy;
    const float* states;
    void* dx;
    void* ddelta;
    float* dA;
    float* dBs;
    float* dCs;
    float* dDs;
    float* ddelta_bias;
    int64_t batch, D, N, H, W;
    int32_t dtype, dout_dtype, delta_softplus, scans;
    /* dBs / dCs receive one contribution per channel and position: all D channels of a batch image add into the same few
     * L2 lines at the same time, which serialises (measured: 18 % of the config-2 backward, 60 % at 14x14 with D = 1024).
     * With acc_replicas = R > 1 the accumulators are (R, B, 4, N, L) and channel d adds into replica d % R; the caller sums
     * over R afterwards (a few MB).  0 or 1: plain (B, 4, N, L). */
    int32_t acc_replicas, reserved;
} xfs_ss2d_bwd_args;

XFS_API int xfs_ss2d_supported(int64_t D, int64_t N, int64_t H, int64_t W, int dtype, int backward);
XFS_API int64_t xfs_ss2d_states_len(int64_t N, int64_t H, int64_t W, int dtype, int out_dtype);   /* floats per (batch, 4*D) row */
XFS_API int xfs_ss2d_fwd(const xfs_ss2d_fwd_args* a, xfs_stream_t stream);
XFS_API int xfs_ss2d_bwd(const xfs_ss2d_bwd_args* a, xfs_stream_t stream);

/* -------------------------------------------------------------------------------------------------------------
 * Deep fusion: the three SS2D streams of Cross_SS2Dv5.forward_corev2 (models/fusion_vmamba.py:446-578) in ONE launch.
 * Stream order is free; the reference runs x_fuse, x, x2 (:483-512, :517-543, :548-574).  All streams share A, Ds,
 * delta_bias (one parameter set, :470-478); every stream has its own x, delta, Bs; Cs[s] may alias -- the two view
 * streams read the fused stream's C (:536-538, :567-569), so callers pass the same pointer three times.  Shapes per
 * stream as xfs_ss2d_fwd / xfs_ss2d_bwd; H*W <= 64 and N <= 16 (the fusion blocks run on the last backbone stage: 7x7
 * tokens, N = 16); other shapes return XFS_ERR_UNSUPPORTED and callers use xfs_ss2d_fwd per stream.
 * bwd: dBs[s] / dCs[s] (batch, 4, N, L) f32 are ACCUMULATED into (zero-filled by the caller; aliased dCs sum the
 * three streams); dA, dDs, ddelta_bias accumulate over the streams as they do over the batch.
 * ----------------------------------------------------------------------------------------------------------- */
typedef struct {
    const void* x[3];
    const void* delta[3];
    const void* Bs[3];
    const void* Cs[3];
    void* y[3];
    float* states[3];        /* each nullable */
    const float* A;
    const float* Ds;         /* nullable */
    const float* delta_bias; /* nullable */
    int64_t batch, D, N, H, W;
    int32_t dtype, out_dtype, delta_softplus, nstreams;   /* nstreams in 1..3 */
} xfs_cross_ss2d_x3_fwd_args;

typedef struct {
    const void* x[3];
    const void* delta[3];
    const void* Bs[3];
    const void* Cs[3];
    const void* dy[3];
    void* dx[3];
    void* ddelta[3];
    float* dBs[3];
    float* dCs[3];
    const float* A;
    const float* Ds;
    const float* delta_bias;
    float* dA;
    float* dDs;
    float* ddelta_bias;
    int64_t batch, D, N, H, W;
    int32_t dtype, dout_dtype, delta_softplus, nstreams;
} xfs_cross_ss2d_x3_bwd_args;

XFS_API int xfs_cross_ss2d_x3_supported(int64_t N, int64_t H, int64_t W);
XFS_API int xfs_cross_ss2d_x3_fwd(const xfs_cross_ss2d_x3_fwd_args* a, xfs_stream_t stream);
XFS_API int xfs_cross_ss2d_x3_bwd(const xfs_cross_ss2d_x3_bwd_args* a, xfs_stream_t stream);

/* -------------------------------------------------------------------------------------------------------------
 * Shallow fusion: SwappingScan_multiview + selective scan (K = 2 halves) + SwappingMerge_multiview of
 * ShallowFuse_SS2Dv4.forward_corev2 (models/fusion_vmamba.py:812, 831-835; :189-241) in ONE kernel: the scan input
 * of half k, channel c is x2[b, c] when (c even) == (k == 0), else x[b, c] (:198-214); output half 0 -> y, half 1 -> y2.
 *   x, x2 : (batch, D, L) dtype            delta : (batch, 2*D, L) dtype           A : (2*D, N) f32
 *   Bs, Cs: (batch, 2, N, L) dtype         Ds, delta_bias : (2*D) f32 or NULL      y, y2 : (batch, D, L) out_dtype
 *   states: (batch, 2*D, 1, N) f32, nullable.    L <= 64, N <= 16 (else XFS_ERR_UNSUPPORTED: compose
 *   xfs_swap_scan + xfs_selective_scan_fwd + xfs_swap_merge).
 * bwd follows the reference AS WRITTEN: dy / dy2 are the gradients of halves 0 / 1 (SwappingMerge.backward stacks them,
 * :234-241) and the scan-input gradient of half 0 goes to dx, of half 1 to dx2 WITHOUT un-swapping the even channels
 * (SwappingScan.backward, :217-221).  dA, dBs, dCs, dDs, ddelta_bias are accumulated into (caller zero-fills).
 * ----------------------------------------------------------------------------------------------------------- */
typedef struct {
    const void* x;
    const void* x2;
    const void* delta;
    const float* A;
    const void* Bs;
    const void* Cs;
    const float* Ds;         /* nullable */
    const float* delta_bias; /* nullable */
    void* y;
    void* y2;
    float* states;           /* nullable */
    int64_t batch, D, N, L;
    int32_t dtype, out_dtype, delta_softplus, reserved;
} xfs_swap_scan_fused_fwd_args;

typedef struct {
    const void* x;
    const void* x2;
    const void* delta;
    const float* A;
    const void* Bs;
    const void* Cs;
    const float* Ds;
    const float* delta_bias;
    const void* dy;
    const void* dy2;
    void* dx;
    void* dx2;
    void* ddelta;
    float* dA;
    float* dBs;
    float* dCs;
    float* dDs;
    float* ddelta_bias;
    int64_t batch, D, N, L;
    int32_t dtype, dout_dtype, delta_softplus, reserved;
} xfs_swap_scan_fused_bwd_args;

XFS_API int xfs_swap_scan_fused_supported(int64_t N, int64_t L);
XFS_API int xfs_swap_scan_fused_fwd(const xfs_swap_scan_fused_fwd_args* a, xfs_stream_t stream);
XFS_API int xfs_swap_scan_fused_bwd(const xfs_swap_scan_fused_bwd_args* a, xfs_stream_t stream);

/* -------------------------------------------------------------------------------------------------------------
 * LayerNorm2d: LayerNorm over the channel dimension of a channel-first tensor, the consumer of the merged scan output
 * (reference LayerNorm2d, models/fusion_vmamba.py:52-57; out_norm at :1183-1188).  x, y, dy, dx: (B, C, HW) dtype;
 * weight, bias: (C) f32 or NULL; mean, rstd: (B, HW) f32 (written by fwd when non-NULL, required by bwd);
 * dweight, dbias: (C) f32, ACCUMULATED into (caller zero-fills), nullable.
 * ----------------------------------------------------------------------------------------------------------- */
XFS_API int xfs_layernorm2d_fwd(const void* x, const float* weight, const float* bias, void* y, float* mean, float* rstd,
                                int64_t B, int64_t C, int64_t HW, float eps, int dtype, xfs_stream_t stream);
XFS_API int xfs_layernorm2d_bwd(const void* x, const void* dy, const float* weight, const float* mean, const float* rstd,
                                void* dx, float* dweight, float* dbias, int64_t B, int64_t C, int64_t HW, int dtype,
                                xfs_stream_t stream);

/* -------------------------------------------------------------------------------------------------------------
 * Depthwise 3x3 convolution (stride 1, zero padding 1) fused with SiLU: the producer of the scan input in every SS2D
 * block (reference nn.Conv2d(groups=d_inner) + act, models/fusion_vmamba.py:405-413,1199-1200; :595-601; :855-858).
 * x, y, dy, dx: (B, C, H, W) dtype, contiguous; weight: (C, 1, 3, 3) f32; bias: (C) f32 or NULL; act: 1 = SiLU, 0 = none.
 * bwd writes dx and per-plane partial sums part: (B*C, 10) f32 = 9 filter taps + bias; the caller sums over B
 * (deterministic, no global atomics).  xfs_dwconv3x3_supported: 0 when a plane does not fit shared memory.
 * ----------------------------------------------------------------------------------------------------------- */
XFS_API int xfs_dwconv3x3_supported(int64_t H, int64_t W, int backward);
XFS_API int xfs_dwconv3x3_fwd(const void* x, const float* weight, const float* bias, void* y, int64_t B, int64_t C, int64_t H,
                              int64_t W, int dtype, int act, xfs_stream_t stream);
XFS_API int xfs_dwconv3x3_bwd(const void* x, const float* weight, const float* bias, const void* dy, void* dx, float* part,
                              int64_t B, int64_t C, int64_t H, int64_t W, int dtype, int act, xfs_stream_t stream);

/* -------------------------------------------------------------------------------------------------------------
 * Low-rank delta projection of the SS2D core (reference F.conv1d(dts_r, dt_projs_weight, groups=K),
 * models/fusion_vmamba.py:1155-1157; einsum at :818):  delta[b, k*D + d, l] = sum_r W[k, d, r] * z[b, k, r, l].
 * z: rows of L contiguous elements at z + b*z_batch_stride + k*z_route_stride + r*L (a slice of the x_proj output is
 * accepted as it lies); W: (K, D, R) f32, R <= 64; delta: (B, K*D, L) dtype, contiguous.
 * ----------------------------------------------------------------------------------------------------------- */
XFS_API int xfs_dt_proj_fwd(const void* z, const float* W, void* delta, int64_t B, int64_t K, int64_t D, int64_t R, int64_t L,
                            int64_t z_batch_stride, int64_t z_route_stride, int dtype, xfs_stream_t stream);

/* Backward of the projection, one pass over g each: dz[b,k,r,l] = sum_d W[k,d,r] g[b,k*D+d,l] and
 * dW[k,d,r] += sum_{b,l} g[b,k*D+d,l] z[b,k,r,l] (the autograd of the grouped conv1d above; the reference leaves it to cuDNN).
 * g: (B, K*D, L) contiguous; z as in the forward; dz: (B, K, R, L) f32 contiguous, may be NULL; dW: (K, D, R) f32, accumulated
 * into (zero it first), may be NULL.  f32 rows only (xfs_dt_proj_bwd_supported; 16-bit rows: XFS_ERR_UNSUPPORTED and the caller uses
 * its library GEMMs); rows that are not 16-byte aligned (L % 4 != 0, odd strides) are copied in 4-byte pieces. */
XFS_API int xfs_dt_proj_bwd_supported(int64_t R, int64_t L, int64_t z_batch_stride, int64_t z_route_stride, int dtype);
XFS_API int xfs_dt_proj_bwd(const void* g, const void* z, const float* W, void* dz, float* dW, int64_t B, int64_t K, int64_t D, int64_t R,
                            int64_t L, int64_t z_batch_stride, int64_t z_route_stride, int dtype, xfs_stream_t stream);

/* number of kernels this library has launched since load (process-wide, relaxed atomic): lets bench.py report
 * `gpu_launches` from a count instead of a guess */
XFS_API int64_t xfs_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* XFSCAN_H_ */
