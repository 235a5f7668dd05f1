"""Cross-view fusion scans and the fused SS2D core.

* ``SwappingScan_multiview`` / ``SwappingMerge_multiview`` -- drop-ins for ``models/fusion_vmamba.py:189-241`` (the
  reference builds a boolean mask on the CPU and runs ~8 index kernels; here it is one copy kernel).  The backward
  passes follow the reference AS WRITTEN: ``SwappingScan_multiview.backward`` returns ``ys[:,0], ys[:,1]`` without
  un-swapping the even channels (``:217-221``), so gradients of even channels reach the *other* view.  Parity with
  the reference is the contract; ``exact_adjoint=True`` on ``swapping_scan`` gives the mathematically exact adjoint.
* ``ss2d_scan`` -- NEW fused entry point: ``cross_merge(selective_scan(cross_scan(x), ...))`` of
  ``SS2Dv2.forward_corev2`` (``models/fusion_vmamba.py:1145,1170-1174``) as ONE forward kernel and ONE backward kernel
  (csrc/ss2d_fused.cu).  When the per-channel working set does not fit shared memory it composes the three
  stand-alone CUDA operators instead (still no CPU path).
"""
from __future__ import annotations

import os

import torch

from . import _lib
from .csm import cross_merge_raw, cross_scan_raw
from .csms6s import _check_scan_args, selective_scan_bwd_raw, selective_scan_fwd_raw

__all__ = ["SwappingScan_multiview", "SwappingMerge_multiview", "swapping_scan", "swapping_merge", "ss2d_scan",
           "SS2DScanFn", "ss2d_fused_supported", "ss2d_fwd_raw", "ss2d_bwd_raw", "ss2d_acc_replicas",
           "cross_ss2d_x3", "cross_ss2d_x3_supported", "swap_scan_fused", "swap_scan_fused_supported"]


def _swap_scan_raw(x, x2):
    dev = _lib.require_cuda(x, x2)
    if x.shape != x2.shape or x.dtype != x2.dtype:
        raise RuntimeError(f"SwappingScan: the two views must match; got {tuple(x.shape)}/{x.dtype} vs {tuple(x2.shape)}/{x2.dtype}")
    B, C = x.shape[:2]
    L = x[0, 0].numel()
    x, x2 = x.contiguous(), x2.contiguous()
    out = torch.empty((B, 2, C, L), dtype=x.dtype, device=dev)
    if out.numel():
        with torch.cuda.device(dev):
            rc = _lib.lib().xfs_swap_scan(_lib.ptr(x), _lib.ptr(x2), _lib.ptr(out), B, C, L, _lib.dtype_code(x), _lib.stream(dev))
        _lib.check(rc, "swap_scan")
    return out


def _swap_merge_raw(ys):
    dev = _lib.require_cuda(ys)
    B, K, C, L = ys.shape
    if K != 2:
        raise RuntimeError(f"SwappingMerge expects (B, 2, C, L); got {tuple(ys.shape)}")
    ys = ys.contiguous()
    y = torch.empty((B, C, L), dtype=ys.dtype, device=dev)
    y2 = torch.empty_like(y)
    if y.numel():
        with torch.cuda.device(dev):
            rc = _lib.lib().xfs_swap_merge(_lib.ptr(ys), _lib.ptr(y), _lib.ptr(y2), B, C, L, _lib.dtype_code(ys), _lib.stream(dev))
        _lib.check(rc, "swap_merge")
    return y, y2


def _swap_stack_raw(y, y2):
    dev = _lib.require_cuda(y, y2)
    B, C, L = y.shape
    y, y2 = y.contiguous(), y2.contiguous()
    ys = torch.empty((B, 2, C, L), dtype=y.dtype, device=dev)
    if ys.numel():
        with torch.cuda.device(dev):
            rc = _lib.lib().xfs_swap_stack(_lib.ptr(y), _lib.ptr(y2), _lib.ptr(ys), B, C, L, _lib.dtype_code(y), _lib.stream(dev))
        _lib.check(rc, "swap_stack")
    return ys


class SwappingScan_multiview(torch.autograd.Function):
    """models/fusion_vmamba.py:189-221"""

    @staticmethod
    def forward(ctx, x: torch.Tensor, x2: torch.Tensor, exact_adjoint: bool = False):
        B, C, H, W = x.shape
        ctx.shape = (B, C, H, W)
        ctx.exact_adjoint = exact_adjoint
        return _swap_scan_raw(x, x2)

    @staticmethod
    def backward(ctx, ys: torch.Tensor):
        B, C, H, W = ctx.shape
        if ctx.exact_adjoint:
            # the swap is an involution on (view, even channel): its adjoint is the same swap of the two grad halves
            g = _swap_scan_raw(ys[:, 0], ys[:, 1])
            return g[:, 0].reshape(B, C, H, W), g[:, 1].reshape(B, C, H, W), None
        # reference behaviour: plain split (models/fusion_vmamba.py:217-221)
        return ys[:, 0].reshape(B, -1, H, W), ys[:, 1].reshape(B, -1, H, W), None


class SwappingMerge_multiview(torch.autograd.Function):
    """models/fusion_vmamba.py:224-241"""

    @staticmethod
    def forward(ctx, ys: torch.Tensor):
        return _swap_merge_raw(ys)

    @staticmethod
    def backward(ctx, x: torch.Tensor, x2: torch.Tensor):
        return _swap_stack_raw(x, x2)


def swapping_scan(x, x2, exact_adjoint=False):
    return SwappingScan_multiview.apply(x, x2, exact_adjoint)


def swapping_merge(ys):
    return SwappingMerge_multiview.apply(ys)


# ------------------------------------------------------------------------------------------------------------------
# fused SS2D core
# ------------------------------------------------------------------------------------------------------------------
def ss2d_fused_supported(D: int, N: int, H: int, W: int, dtype: torch.dtype, backward: bool = False) -> bool:
    return bool(_lib.lib().xfs_ss2d_supported(D, N, H, W, _lib._DTYPES[dtype], int(backward)))


def ss2d_fwd_raw(x, delta, A, Bs, Cs, Ds, delta_bias, delta_softplus=True, out_dtype=torch.float32, need_states=True,
                 y=None, states=None):
    """One launch of the fused forward kernel (C ABI ``xfs_ss2d_fwd``) on contiguous, already validated tensors.
    ``y`` / ``states`` may be passed to reuse buffers (CUDA-graph capture, benchmarks)."""
    Bsz, D, H, W = x.shape
    L, N, dev = H * W, Bs.shape[2], x.device
    if y is None:
        y = torch.empty((Bsz, D, L), dtype=out_dtype, device=dev)
    if states is None and need_states:
        states = torch.empty((Bsz, 4 * D, _lib.ss2d_states_len(N, H, W, x.dtype, y.dtype)), dtype=torch.float32, device=dev)
    if y.numel():
        args = _lib.Ss2dFwdArgs(_lib.ptr(x), _lib.ptr(delta), _lib.ptr(A), _lib.ptr(Bs), _lib.ptr(Cs), _lib.ptr(Ds),
                                _lib.ptr(delta_bias), _lib.ptr(y), _lib.ptr(states), Bsz, D, N, H, W,
                                _lib.dtype_code(x), _lib.dtype_code(y), int(bool(delta_softplus)), 0)
        with torch.cuda.device(dev):
            rc = _lib.lib().xfs_ss2d_fwd(args, _lib.stream(dev))
        _lib.check(rc, "ss2d_fwd")
    return y, states


def ss2d_acc_replicas(D, L, batch=1):
    """Replicas of the dBs/dCs accumulators (``xfs_ss2d_bwd_args.acc_replicas``).  Every channel of a batch image adds into the
    same L2 lines; the backward kernels already walk the batch index fastest, so with a large batch the CTAs resident at one
    time belong to different images and no replica is needed (measured: R = 1 is fastest from batch 64 on).  Small batches
    bring the channels of one image back together, and spreading them over R copies removes the serialisation."""
    env = os.environ.get("XFS_ACC_REPLICAS")
    if env:
        return max(1, min(int(env), int(D), 64))
    if L <= 64 or batch >= 32:
        return 1                      # L <= 64: the short-sequence kernel sums 32 channels per CTA in shared memory
    R = 4 if L >= 2048 else (8 if D < 1024 else 16)
    if batch >= 8:
        R //= 2
    return max(1, min(R, int(D)))


def ss2d_bwd_raw(x, delta, A, Bs, Cs, Ds, delta_bias, dy, states, delta_softplus=True, out=None, zero=True, reduce=True):
    """One launch of the fused backward kernel (C ABI ``xfs_ss2d_bwd``).  Returns (dx, ddelta, dA, dBs, dCs, dDs,
    ddelta_bias) with dBs/dCs fp32.  ``out`` may carry the 7 buffers for reuse; its dBs/dCs entries may be replicated
    accumulators (R, B, 4, N, L) (see ``ss2d_acc_replicas``), which are summed over R here (``reduce=False``: returned as they
    are, for callers that time the kernel alone).  The accumulated buffers are
    zero-filled here (stream-ordered memsets) unless ``zero=False`` (caller already did), as the reference host code does
    (selective_scan.cpp:331-337)."""
    Bsz, D, H, W = x.shape
    N, dev = Bs.shape[2], x.device
    if out is None:
        R = ss2d_acc_replicas(D, H * W, Bsz)
        acc_shape = tuple(Bs.shape) if R == 1 else (R,) + tuple(Bs.shape)
        out = (torch.empty_like(x), torch.empty_like(delta), torch.empty_like(A),
               torch.empty(acc_shape, dtype=torch.float32, device=dev), torch.empty(acc_shape, dtype=torch.float32, device=dev),
               None if Ds is None else torch.empty_like(Ds), None if delta_bias is None else torch.empty_like(delta_bias))
    dx, ddelta, dA, dBs, dCs, dDs, dbias = out
    R = dBs.shape[0] if dBs.dim() == 5 else 1
    if dCs.dim() != dBs.dim() or (dCs.dim() == 5 and dCs.shape[0] != R):
        raise RuntimeError("ss2d_bwd_raw: dBs and dCs accumulators must have the same number of replicas")
    if zero:
        for acc in (dA, dBs, dCs, dDs, dbias):
            if acc is not None:
                acc.zero_()
    if x.numel():
        args = _lib.Ss2dBwdArgs(_lib.ptr(x), _lib.ptr(delta), _lib.ptr(A), _lib.ptr(Bs), _lib.ptr(Cs), _lib.ptr(Ds),
                                _lib.ptr(delta_bias), _lib.ptr(dy), _lib.ptr(states), _lib.ptr(dx), _lib.ptr(ddelta),
                                _lib.ptr(dA), _lib.ptr(dBs), _lib.ptr(dCs), _lib.ptr(dDs), _lib.ptr(dbias),
                                Bsz, D, N, H, W, _lib.dtype_code(x), _lib.dtype_code(dy), int(bool(delta_softplus)), 0, R, 0)
        with torch.cuda.device(dev):
            rc = _lib.lib().xfs_ss2d_bwd(args, _lib.stream(dev))
        _lib.check(rc, "ss2d_bwd")
    if dBs.dim() == 5 and reduce:
        dBs, dCs = dBs.sum(0), dCs.sum(0)
    return dx, ddelta, dA, dBs, dCs, dDs, dbias


class SS2DScanFn(torch.autograd.Function):
    """y = cross_merge(selective_scan(cross_scan(x), delta, A, Bs, Cs, Ds, delta_bias)), all four routes in one kernel."""

    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda")
    def forward(ctx, x, delta, A, Bs, Cs, Ds, delta_bias, delta_softplus, oflex, need_grad=None):
        if x.dim() != 4:
            raise RuntimeError(f"ss2d_scan: x must be (B, D, H, W); got {tuple(x.shape)}")
        Bsz, D, H, W = x.shape
        L = H * W
        delta = delta.reshape(Bsz, 4 * D, L)
        if x.dtype != delta.dtype:
            raise RuntimeError(f"ss2d_scan: x ({x.dtype}) and delta ({delta.dtype}) must share one dtype")
        dev, Bs, Cs, _, _, _, G, N = _check_scan_args(delta, delta, A, Bs, Cs, Ds, delta_bias)
        _lib.require_cuda(x)
        if G != 4:
            raise RuntimeError(f"ss2d_scan: Bs/Cs must be (B, 4, N, L); got {tuple(Bs.shape)}")
        need = any(ctx.needs_input_grad) if need_grad is None else bool(need_grad)     # see _lib.grad_needed
        fused = ss2d_fused_supported(D, N, H, W, x.dtype, False) and (not need or ss2d_fused_supported(D, N, H, W, x.dtype, True))
        x, delta, A, Bs, Cs = (t.contiguous() for t in (x, delta, A, Bs, Cs))
        Ds = None if Ds is None else Ds.contiguous()
        delta_bias = None if delta_bias is None else delta_bias.contiguous()
        out_dtype = torch.float32 if oflex else x.dtype
        if fused:
            y, states = ss2d_fwd_raw(x, delta, A, Bs, Cs, Ds, delta_bias, delta_softplus, out_dtype, need)
        else:
            xs = cross_scan_raw(x).view(Bsz, 4 * D, L)
            ys, states, _ = selective_scan_fwd_raw(xs, delta, A, Bs, Cs, Ds, delta_bias, delta_softplus, oflex, need_states=need)
            y = cross_merge_raw(ys.view(Bsz, 4, D, L), H, W)
        if need:
            ctx.fused, ctx.delta_softplus = fused, delta_softplus
            ctx.has_D, ctx.has_bias = Ds is not None, delta_bias is not None
            ctx.save_for_backward(*[t for t in (x, delta, A, Bs, Cs, Ds, delta_bias) if t is not None], states)
        return y

    @staticmethod
    @torch.amp.custom_bwd(device_type="cuda")
    def backward(ctx, dy):
        saved = list(ctx.saved_tensors)
        states = saved.pop()
        x, delta, A, Bs, Cs = saved[:5]
        rest = saved[5:]
        Ds = rest.pop(0) if ctx.has_D else None
        delta_bias = rest.pop(0) if ctx.has_bias else None
        Bsz, D, H, W = x.shape
        L = H * W
        N = Bs.shape[2]
        dev = x.device
        dy = dy.contiguous()
        if dy.dtype not in (torch.float32, x.dtype):
            dy = dy.to(x.dtype)
        if ctx.fused:
            dx, ddelta, dA, dBs, dCs, dDs, dbias = ss2d_bwd_raw(x, delta, A, Bs, Cs, Ds, delta_bias, dy, states, ctx.delta_softplus)
            dBs, dCs = dBs.to(Bs.dtype), dCs.to(Cs.dtype)
        else:
            xs = cross_scan_raw(x).view(Bsz, 4 * D, L)
            dys = cross_scan_raw(dy.view(Bsz, D, H, W)).view(Bsz, 4 * D, L)
            du, ddelta, dA, dBs, dCs, dDs, dbias = selective_scan_bwd_raw(xs, delta, A, Bs, Cs, Ds, delta_bias, dys, states,
                                                                          ctx.delta_softplus)
            dx = cross_merge_raw(du.view(Bsz, 4, D, L), H, W).view(Bsz, D, H, W)
        return dx, ddelta, dA, dBs, dCs, dDs, dbias, None, None, None


def ss2d_scan(x, delta, A, Bs, Cs, Ds=None, delta_bias=None, delta_softplus=True, oflex=True):
    """Fused SS2D core.

    x (B, D, H, W); delta (B, 4*D, H*W) or (B, 4, D, H*W) -- the dt_proj output, in the scan order of each route;
    A (4*D, N) f32; Bs, Cs (B, 4, N, H*W) in scan order; Ds, delta_bias (4*D) f32.  Returns y (B, D, H*W) in spatial
    order: f32 when ``oflex`` else x.dtype -- exactly ``cross_merge_fn(selective_scan_fn(cross_scan_fn(x).view(B,-1,L),
    delta, A, Bs, Cs, Ds, delta_bias, delta_softplus, oflex).view(B,4,-1,H,W))``.
    """
    if x.dim() == 4 and delta.dim() == 4:          # (B, 4, D, L): flatten OUTSIDE the Function so that autograd un-flattens ddelta
        delta = delta.reshape(x.shape[0], 4 * x.shape[1], -1)
    need = _lib.grad_needed(x, delta, A, Bs, Cs, Ds, delta_bias)
    return SS2DScanFn.apply(x, delta, A, Bs, Cs, Ds, delta_bias, delta_softplus, oflex, need)


# ---------------------------------------------------------------------------------------------------------------------
# fusion blocks at their real shapes (L <= 64, N <= 16): single-launch kernels of csrc/fusion_small.cu
# ---------------------------------------------------------------------------------------------------------------------
def cross_ss2d_x3_supported(N, H, W) -> bool:
    return bool(_lib.lib().xfs_cross_ss2d_x3_supported(int(N), int(H), int(W)))


def swap_scan_fused_supported(N, L) -> bool:
    return bool(_lib.lib().xfs_swap_scan_fused_supported(int(N), int(L)))


class CrossSS2Dx3Fn(torch.autograd.Function):
    """The three SS2D streams of Cross_SS2Dv5.forward_corev2 (models/fusion_vmamba.py:485-569) in one forward and one
    backward launch.  Inputs: 3 x (x, delta, Bs), ONE Cs (the fused stream's, shared by all three, :536-538/:567-569), and the
    shared A, Ds, delta_bias.  Returns the three merged outputs (B, D, L) fp32."""

    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda")
    def forward(ctx, x0, x1, x2, d0, d1, d2, B0, B1, B2, Cs, A, Ds, delta_bias, need_grad):
        xs, ds, Bs = [x0, x1, x2], [d0, d1, d2], [B0, B1, B2]
        dev = _lib.require_cuda(*xs, *ds, *Bs, Cs, A, Ds, delta_bias)
        Bsz, D, H, W = x0.shape
        L, N = H * W, Cs.shape[2]
        xs = [t.contiguous() for t in xs]
        ds = [t.reshape(Bsz, 4 * D, L).contiguous() for t in ds]
        Bs = [t.contiguous() for t in Bs]
        Cs, A = Cs.contiguous(), A.float().contiguous()
        Ds = None if Ds is None else Ds.float().contiguous()
        delta_bias = None if delta_bias is None else delta_bias.float().contiguous()
        ys = [torch.empty((Bsz, D, L), dtype=torch.float32, device=dev) for _ in range(3)]
        args = _lib.X3FwdArgs(_lib.p3(xs), _lib.p3(ds), _lib.p3(Bs), _lib.p3([Cs, Cs, Cs]), _lib.p3(ys), _lib.p3([None] * 3),
                              _lib.ptr(A), _lib.ptr(Ds), _lib.ptr(delta_bias), Bsz, D, N, H, W,
                              _lib.dtype_code(x0), _lib.F32, 1, 3)
        with torch.cuda.device(dev):
            rc = _lib.lib().xfs_cross_ss2d_x3_fwd(args, _lib.stream(dev))
        _lib.check(rc, "cross_ss2d_x3_fwd")
        if need_grad:
            ctx.has_D, ctx.has_bias = Ds is not None, delta_bias is not None
            ctx.save_for_backward(*xs, *ds, *Bs, Cs, A, *[t for t in (Ds, delta_bias) if t is not None])
        return tuple(ys)

    @staticmethod
    @torch.amp.custom_bwd(device_type="cuda")
    def backward(ctx, g0, g1, g2):
        sv = list(ctx.saved_tensors)
        xs, ds, Bs, Cs, A = sv[0:3], sv[3:6], sv[6:9], sv[9], sv[10]
        rest = sv[11:]
        Ds = rest.pop(0) if ctx.has_D else None
        delta_bias = rest.pop(0) if ctx.has_bias else None
        dev = xs[0].device
        Bsz, D, H, W = xs[0].shape
        L, N = H * W, Cs.shape[2]
        gs = [(torch.zeros((Bsz, D, L), dtype=torch.float32, device=dev) if g is None else g.contiguous().float()) for g in (g0, g1, g2)]
        dxs = [torch.empty_like(t) for t in xs]
        dds = [torch.empty_like(t) for t in ds]
        dBs = [torch.zeros(Bs[0].shape, dtype=torch.float32, device=dev) for _ in range(3)]
        dCs = torch.zeros(Cs.shape, dtype=torch.float32, device=dev)
        dA = torch.zeros_like(A)
        dDs = None if Ds is None else torch.zeros_like(Ds)
        dbias = None if delta_bias is None else torch.zeros_like(delta_bias)
        args = _lib.X3BwdArgs(_lib.p3(xs), _lib.p3(ds), _lib.p3(Bs), _lib.p3([Cs, Cs, Cs]), _lib.p3(gs), _lib.p3(dxs), _lib.p3(dds),
                              _lib.p3(dBs), _lib.p3([dCs, dCs, dCs]), _lib.ptr(A), _lib.ptr(Ds), _lib.ptr(delta_bias),
                              _lib.ptr(dA), _lib.ptr(dDs), _lib.ptr(dbias), Bsz, D, N, H, W, _lib.dtype_code(xs[0]), _lib.F32, 1, 3)
        with torch.cuda.device(dev):
            rc = _lib.lib().xfs_cross_ss2d_x3_bwd(args, _lib.stream(dev))
        _lib.check(rc, "cross_ss2d_x3_bwd")
        dt = Bs[0].dtype
        return (*dxs, *dds, *[t.to(dt) for t in dBs], dCs.to(Cs.dtype), dA, dDs, dbias, None)


def cross_ss2d_x3(xs, deltas, Bs, Cs, A, Ds=None, delta_bias=None):
    """three SS2D streams sharing one parameter set and one Cs: (y0, y1, y2), each (B, D, H*W) fp32 (delta_softplus, oflex)"""
    need = _lib.grad_needed(*xs, *deltas, *Bs, Cs, A, Ds, delta_bias)
    return CrossSS2Dx3Fn.apply(*xs, *deltas, *Bs, Cs, A, Ds, delta_bias, need)


class SwapScanFusedFn(torch.autograd.Function):
    """SwappingScan_multiview + selective scan (K = 2) + SwappingMerge_multiview (models/fusion_vmamba.py:812, 831-835) in one
    forward and one backward launch; the backward follows the reference as written (no un-swap, :217-221)."""

    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda")
    def forward(ctx, x, x2, delta, A, Bs, Cs, Ds, delta_bias, need_grad):
        dev = _lib.require_cuda(x, x2, delta, A, Bs, Cs, Ds, delta_bias)
        Bsz, D = x.shape[:2]
        L, N = x[0, 0].numel(), Bs.shape[2]
        x, x2, Bs, Cs = x.contiguous(), x2.contiguous(), Bs.contiguous(), Cs.contiguous()
        delta = delta.reshape(Bsz, 2 * D, L).contiguous()
        A = A.float().contiguous()
        Ds = None if Ds is None else Ds.float().contiguous()
        delta_bias = None if delta_bias is None else delta_bias.float().contiguous()
        y = torch.empty((Bsz, D, L), dtype=torch.float32, device=dev)
        y2 = torch.empty_like(y)
        args = _lib.SwapFusedFwdArgs(_lib.ptr(x), _lib.ptr(x2), _lib.ptr(delta), _lib.ptr(A), _lib.ptr(Bs), _lib.ptr(Cs), _lib.ptr(Ds),
                                     _lib.ptr(delta_bias), _lib.ptr(y), _lib.ptr(y2), None, Bsz, D, N, L,
                                     _lib.dtype_code(x), _lib.F32, 1, 0)
        with torch.cuda.device(dev):
            rc = _lib.lib().xfs_swap_scan_fused_fwd(args, _lib.stream(dev))
        _lib.check(rc, "swap_scan_fused_fwd")
        if need_grad:
            ctx.has_D, ctx.has_bias = Ds is not None, delta_bias is not None
            ctx.save_for_backward(x, x2, delta, A, Bs, Cs, *[t for t in (Ds, delta_bias) if t is not None])
        return y, y2

    @staticmethod
    @torch.amp.custom_bwd(device_type="cuda")
    def backward(ctx, gy, gy2):
        sv = list(ctx.saved_tensors)
        x, x2, delta, A, Bs, Cs = sv[:6]
        rest = sv[6:]
        Ds = rest.pop(0) if ctx.has_D else None
        delta_bias = rest.pop(0) if ctx.has_bias else None
        dev = x.device
        Bsz, D = x.shape[:2]
        L, N = x[0, 0].numel(), Bs.shape[2]
        zero = lambda: torch.zeros((Bsz, D, L), dtype=torch.float32, device=dev)
        gy = zero() if gy is None else gy.contiguous().float()
        gy2 = zero() if gy2 is None else gy2.contiguous().float()
        dx, dx2, ddelta = torch.empty_like(x), torch.empty_like(x2), torch.empty_like(delta)
        dA = torch.zeros_like(A)
        dBs = torch.zeros(Bs.shape, dtype=torch.float32, device=dev)
        dCs = torch.zeros(Cs.shape, dtype=torch.float32, device=dev)
        dDs = None if Ds is None else torch.zeros_like(Ds)
        dbias = None if delta_bias is None else torch.zeros_like(delta_bias)
        args = _lib.SwapFusedBwdArgs(_lib.ptr(x), _lib.ptr(x2), _lib.ptr(delta), _lib.ptr(A), _lib.ptr(Bs), _lib.ptr(Cs), _lib.ptr(Ds),
                                     _lib.ptr(delta_bias), _lib.ptr(gy), _lib.ptr(gy2), _lib.ptr(dx), _lib.ptr(dx2), _lib.ptr(ddelta),
                                     _lib.ptr(dA), _lib.ptr(dBs), _lib.ptr(dCs), _lib.ptr(dDs), _lib.ptr(dbias), Bsz, D, N, L,
                                     _lib.dtype_code(x), _lib.F32, 1, 0)
        with torch.cuda.device(dev):
            rc = _lib.lib().xfs_swap_scan_fused_bwd(args, _lib.stream(dev))
        _lib.check(rc, "swap_scan_fused_bwd")
        return dx, dx2, ddelta, dA, dBs.to(Bs.dtype), dCs.to(Cs.dtype), dDs, dbias, None


def swap_scan_fused(x, x2, delta, A, Bs, Cs, Ds=None, delta_bias=None):
    """(y, y2), each (B, D, L) fp32 = SwappingMerge(selective_scan(SwappingScan(x, x2), delta, A, Bs, Cs, Ds, delta_bias, True))"""
    need = _lib.grad_needed(x, x2, delta, A, Bs, Cs, Ds, delta_bias)
    return SwapScanFusedFn.apply(x, x2, delta, A, Bs, Cs, Ds, delta_bias, need)
