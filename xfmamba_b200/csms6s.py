"""S6 selective scan on sm_100a -- drop-in for the reference's ``models/csms6s.py``.

Operator surface kept verbatim (``models/csms6s.py:71-126``):

    selective_scan_fn(u, delta, A, B, C, D=None, delta_bias=None, delta_softplus=True, oflex=True, backend=None)
    SelectiveScanCuda.apply(u, delta, A, B, C, D, delta_bias, delta_softplus, oflex, backend)

``backend`` in (None, "oflex", "mamba", "core") all select the one sm_100a implementation (csrc/selective_scan.cu via
the C ABI of include/xfscan.h).  ``backend="torch"`` -- the reference's pure-PyTorch ``selective_scan_torch`` -- is
deliberately NOT provided: this package has no CPU / PyTorch fallback (the restatement lives in ``oracle/`` as test
infrastructure only).  Argument checks follow the native extension's TORCH_CHECKs
(``models/selective_scan/csrc/selective_scan/selective_scan.cpp:173-223``) and raise ``RuntimeError`` like them.
"""
from __future__ import annotations

import torch

from . import _lib
from .csm import CrossMerge, CrossMergeF, CrossScan, CrossScanF, cross_merge_fn, cross_scan_fn  # noqa: F401  (north-star names)

__all__ = ["selective_scan_fn", "SelectiveScanCuda", "selective_scan_fwd_raw", "selective_scan_bwd_raw",
           "CrossScan", "CrossMerge", "cross_scan_fn", "cross_merge_fn"]

_BACKENDS = (None, "oflex", "mamba", "core", "xfscan")


def _check_scan_args(u, delta, A, B, C, D, delta_bias):
    dev = _lib.require_cuda(u, delta, A, B, C, D, delta_bias)
    if u.dim() != 3:
        raise RuntimeError(f"selective_scan: u must be (batch, dim, seqlen); got {tuple(u.shape)}")
    batch, dim, L = u.shape
    if delta.shape != u.shape:
        raise RuntimeError(f"selective_scan: delta {tuple(delta.shape)} must match u {tuple(u.shape)}")
    if B.dim() == 3:
        B = B.unsqueeze(1)
    if C.dim() == 3:
        C = C.unsqueeze(1)
    if B.dim() != 4 or C.shape != B.shape or B.shape[0] != batch or B.shape[3] != L:
        raise RuntimeError(f"selective_scan: B, C must be (batch, ngroups, dstate, seqlen); got {tuple(B.shape)}, {tuple(C.shape)}")
    G, N = B.shape[1], B.shape[2]
    if dim % G != 0:
        raise RuntimeError("selective_scan: dims should be dividable by n_groups")          # selective_scan.cpp:198
    if N > 256:
        raise RuntimeError("selective_scan: only supports state dimension <= 256")          # selective_scan.cpp:199
    if A.shape != (dim, N):
        raise RuntimeError(f"selective_scan: A must be (dim, dstate) = ({dim}, {N}); got {tuple(A.shape)}")
    if not (delta.dtype == u.dtype and B.dtype == u.dtype and C.dtype == u.dtype):
        raise RuntimeError("selective_scan: u, delta, B, C must share one dtype "
                           f"(got {u.dtype}, {delta.dtype}, {B.dtype}, {C.dtype})")             # selective_scan.cpp:175-180
    _lib.dtype_code(u)
    for name, t in (("A", A), ("D", D), ("delta_bias", delta_bias)):
        if t is not None and t.dtype != torch.float32:
            raise RuntimeError(f"selective_scan: {name} must be float32 (got {t.dtype})")     # selective_scan.cpp:176,211,219
    for name, t in (("D", D), ("delta_bias", delta_bias)):
        if t is not None and t.shape != (dim,):
            raise RuntimeError(f"selective_scan: {name} must have shape ({dim},); got {tuple(t.shape)}")
    return dev, B, C, batch, dim, L, G, N


def selective_scan_fwd_raw(u, delta, A, B, C, D, delta_bias, delta_softplus, oflex=True, need_states=True):
    """mirrors ``ext.fwd(u, delta, A, B, C, D, delta_bias, delta_softplus, nrows, oflex) -> [out, x]`` (selective_scan.cpp:165-172)"""
    dev, B, C, batch, dim, L, G, N = _check_scan_args(u, delta, A, B, C, D, delta_bias)
    u, delta, A, B, C = (t.contiguous() for t in (u, delta, A, B, C))
    D = None if D is None else D.contiguous()
    delta_bias = None if delta_bias is None else delta_bias.contiguous()
    out = torch.empty((batch, dim, L), dtype=torch.float32 if oflex else u.dtype, device=dev)
    states = torch.empty((batch, dim, _lib.num_chunks(L), N), dtype=torch.float32, device=dev) if need_states else None
    if out.numel() > 0:
        args = _lib.ScanFwdArgs(_lib.ptr(u), _lib.ptr(delta), _lib.ptr(A), _lib.ptr(B), _lib.ptr(C), _lib.ptr(D),
                                _lib.ptr(delta_bias), _lib.ptr(out), _lib.ptr(states), batch, dim, N, L, G,
                                _lib.dtype_code(u), _lib.dtype_code(out), int(bool(delta_softplus)), 0)
        with torch.cuda.device(dev):
            rc = _lib.lib().xfs_selective_scan_fwd(args, _lib.stream(dev))
        _lib.check(rc, "selective_scan_fwd")
    return out, states, (u, delta, A, B, C, D, delta_bias)


def _acc_replicas(channels_per_group, L, batch):
    """copies of the dB/dC accumulators (``xfs_scan_bwd_args.acc_replicas``); short rows go through the kernel that already
    sums 128 rows per CTA in shared memory, and with a large batch the batch-fastest walk of the kernel keeps the channels
    of one image apart (see fusion_ops.ss2d_acc_replicas)"""
    if L <= 64 or channels_per_group < 8 or batch >= 32:
        return 1
    return min(4 if L >= 2048 else 8, channels_per_group)


def selective_scan_bwd_raw(u, delta, A, B, C, D, delta_bias, dout, states, delta_softplus):
    """mirrors ``ext.bwd(...) -> [du, ddelta, dA, dB, dC, dD, ddelta_bias]`` (selective_scan.cpp:251-260, 329-360)"""
    dev = u.device
    batch, dim, L = u.shape
    G, N = B.shape[1], B.shape[2]
    if dout.stride(-1) != 1 or not dout.is_contiguous():
        dout = dout.contiguous()
    if dout.dtype not in (torch.float32, u.dtype):
        dout = dout.to(u.dtype)
    du = torch.empty_like(u)
    ddelta = torch.empty_like(delta)
    dA = torch.zeros_like(A)
    # dB / dC accumulators in R copies (channel d -> copy d % R): all dim/G channels of a group add into the same L2 lines
    R = _acc_replicas(dim // G, L, batch)
    acc_shape = tuple(B.shape) if R == 1 else (R,) + tuple(B.shape)
    dB = torch.zeros(acc_shape, dtype=torch.float32, device=dev)
    dC = torch.zeros(acc_shape, dtype=torch.float32, device=dev)
    dD = None if D is None else torch.zeros_like(D)
    dbias = None if delta_bias is None else torch.zeros_like(delta_bias)
    if u.numel() > 0:
        args = _lib.ScanBwdArgs(_lib.ptr(u), _lib.ptr(delta), _lib.ptr(A), _lib.ptr(B), _lib.ptr(C), _lib.ptr(D),
                                _lib.ptr(delta_bias), _lib.ptr(dout), _lib.ptr(states), _lib.ptr(du), _lib.ptr(ddelta),
                                _lib.ptr(dA), _lib.ptr(dB), _lib.ptr(dC), _lib.ptr(dD), _lib.ptr(dbias),
                                batch, dim, N, L, G, _lib.dtype_code(u), _lib.dtype_code(dout), int(bool(delta_softplus)), R)
        with torch.cuda.device(dev):
            rc = _lib.lib().xfs_selective_scan_bwd(args, _lib.stream(dev))
        _lib.check(rc, "selective_scan_bwd")
    if R > 1:
        dB, dC = dB.sum(0), dC.sum(0)
    return du, ddelta, dA, dB.to(B.dtype), dC.to(C.dtype), dD, dbias


class SelectiveScanCuda(torch.autograd.Function):
    """models/csms6s.py:71-109"""

    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda")
    def forward(ctx, u, delta, A, B, C, D=None, delta_bias=None, delta_softplus=False, oflex=True, backend=None, need_grad=None):
        if backend not in _BACKENDS:
            raise RuntimeError(f"selective_scan: backend {backend!r} is not provided by xfmamba_b200 "
                               "(one sm_100a implementation; the torch path is test-only, see oracle/)")
        ctx.delta_softplus = delta_softplus
        ctx.b3, ctx.c3 = B.dim() == 3, C.dim() == 3
        need = any(ctx.needs_input_grad) if need_grad is None else bool(need_grad)     # see _lib.grad_needed
        out, states, saved = selective_scan_fwd_raw(u, delta, A, B, C, D, delta_bias, delta_softplus, oflex, need_states=need)
        if need:
            ctx.has_D, ctx.has_bias = D is not None, delta_bias is not None
            tensors = [t for t in saved if t is not None] + [states]
            ctx.save_for_backward(*tensors)
        return out

    @staticmethod
    @torch.amp.custom_bwd(device_type="cuda")
    def backward(ctx, dout, *args):
        saved = list(ctx.saved_tensors)
        states = saved.pop()
        u, delta, A, B, C = saved[:5]
        rest = saved[5:]
        D = rest.pop(0) if ctx.has_D else None
        delta_bias = rest.pop(0) if ctx.has_bias else None
        du, ddelta, dA, dB, dC, dD, dbias = selective_scan_bwd_raw(u, delta, A, B, C, D, delta_bias, dout, states,
                                                                   ctx.delta_softplus)
        if ctx.b3:
            dB = dB.squeeze(1)
        if ctx.c3:
            dC = dC.squeeze(1)
        return du, ddelta, dA, dB, dC, dD, dbias, None, None, None, None


def selective_scan_fn(
    u: torch.Tensor,           # (B, K * C, L)
    delta: torch.Tensor,       # (B, K * C, L)
    A: torch.Tensor,           # (K * C, N)
    B: torch.Tensor,           # (B, K, N, L)
    C: torch.Tensor,           # (B, K, N, L)
    D: torch.Tensor = None,    # (K * C)
    delta_bias: torch.Tensor = None,   # (K * C)
    delta_softplus=True,
    oflex=True,
    backend=None,
):
    """models/csms6s.py:112-126 (called positionally by the forward_corev2 closures, e.g. models/fusion_vmamba.py:1064-1065)"""
    if backend == "torch":
        raise RuntimeError("selective_scan_fn(backend='torch'): the PyTorch path is not part of xfmamba_b200 "
                           "(no CPU fallback); use the reference's selective_scan_torch or oracle/ in tests")
    need = _lib.grad_needed(u, delta, A, B, C, D, delta_bias)
    return SelectiveScanCuda.apply(u, delta, A, B, C, D, delta_bias, delta_softplus, oflex, backend, need)
