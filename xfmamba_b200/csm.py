"""CrossScan / CrossMerge on sm_100a -- drop-in for the reference's ``models/csm_triton.py``.

Same operator surface as the reference (``models/csm_triton.py:182-273, 501-517``):

    cross_scan_fn(x, in_channel_first=True, out_channel_first=True, one_by_one=False, scans=0, force_torch=False)
    cross_merge_fn(y, in_channel_first=True, out_channel_first=True, one_by_one=False, scans=0, force_torch=False)
    CrossScanF.apply(x, in_channel_first, out_channel_first, one_by_one, scans)
    CrossMergeF.apply(ys, in_channel_first, out_channel_first, one_by_one, scans)

but one hand-written CUDA kernel each (csrc/routes.cu) instead of Triton / torch index ops.  ``force_torch`` is
accepted and ignored (there is one implementation).  The kernels work on channel-first tensors; channel-last layouts
(not used by XFMamba, which is ``channel_first=True`` throughout: models/fusion_vmamba.py:1658) are served by permuting
around the same kernels.  The merge reproduces the torch implementation's add order ``(y0+flip(y2)) + T(y1+flip(y3))``
(``models/csm_triton.py:61-62``) bit for bit.
"""
from __future__ import annotations

import torch

from . import _lib

__all__ = ["cross_scan_fn", "cross_merge_fn", "CrossScanF", "CrossMergeF", "CrossScan", "CrossMerge",
           "cross_scan_raw", "cross_merge_raw"]


def cross_scan_raw(x: torch.Tensor, scans: int = 0, one_by_one: bool = False) -> torch.Tensor:
    """x (B,C,H,W) [one_by_one: (B,4,C,H,W)] contiguous -> (B,4,C,H*W)"""
    dev = _lib.require_cuda(x)
    x = x.contiguous()
    if one_by_one:
        B, K, C, H, W = x.shape
        if K != 4:
            raise RuntimeError(f"cross_scan one_by_one expects 4 inputs on dim 1, got {K}")
    else:
        B, C, H, W = x.shape
    out = torch.empty((B, 4, C, H * W), dtype=x.dtype, device=dev)
    if out.numel() == 0:
        return out
    with torch.cuda.device(dev):
        rc = _lib.lib().xfs_cross_scan(_lib.ptr(x), _lib.ptr(out), B, C, H, W, _lib.dtype_code(x), int(scans),
                                       int(one_by_one), _lib.stream(dev))
    _lib.check(rc, "cross_scan")
    return out


def cross_merge_raw(ys: torch.Tensor, H: int, W: int, scans: int = 0, one_by_one: bool = False) -> torch.Tensor:
    """ys (B,4,C,L) contiguous -> (B,C,L) [one_by_one: (B,4,C,L)]"""
    dev = _lib.require_cuda(ys)
    ys = ys.contiguous()
    B, K, C, L = ys.shape
    if K != 4 or L != H * W:
        raise RuntimeError(f"cross_merge expects (B,4,C,H*W); got {tuple(ys.shape)} for H={H}, W={W}")
    out = torch.empty((B, 4, C, L) if one_by_one else (B, C, L), dtype=ys.dtype, device=dev)
    if out.numel() == 0:
        return out
    with torch.cuda.device(dev):
        rc = _lib.lib().xfs_cross_merge(_lib.ptr(ys), _lib.ptr(out), B, C, H, W, _lib.dtype_code(ys), int(scans),
                                        int(one_by_one), _lib.stream(dev))
    _lib.check(rc, "cross_merge")
    return out


def _scan_any_layout(x, in_channel_first, out_channel_first, one_by_one, scans):
    """layouts of models/csm_triton.py:22-53 / 88-131"""
    if not in_channel_first:
        x = x.permute(0, 3, 4, 1, 2) if one_by_one else x.permute(0, 3, 1, 2)     # (B,H,W,[4,]C) -> (B,[4,]C,H,W)
    y = cross_scan_raw(x, scans, one_by_one)                                       # (B,4,C,L)
    if not out_channel_first:
        y = y.permute(0, 3, 1, 2).contiguous()                                     # (B,L,4,C)
    return y


def _merge_any_layout(ys, in_channel_first, out_channel_first, one_by_one, scans, H, W):
    """layouts of models/csm_triton.py:56-85 / 134-179; ys is (B,4,C,H,W) or (B,H,W,4,C)"""
    if out_channel_first:
        B, K, C = ys.shape[:3]
        ys = ys.reshape(B, K, C, H * W)
    else:
        B, K, C = ys.shape[0], ys.shape[3], ys.shape[4]
        ys = ys.reshape(B, H * W, K, C).permute(0, 2, 3, 1)                        # -> (B,4,C,L)
    y = cross_merge_raw(ys, H, W, scans, one_by_one)                               # (B,C,L) | (B,4,C,L)
    if not in_channel_first:
        y = (y.permute(0, 3, 1, 2) if one_by_one else y.permute(0, 2, 1)).contiguous()   # (B,L,4,C) | (B,L,C)
    return y


class CrossScanF(torch.autograd.Function):
    """models/csm_triton.py:182-225"""

    @staticmethod
    def forward(ctx, x: torch.Tensor, in_channel_first=True, out_channel_first=True, one_by_one=False, scans=0):
        ctx.in_channel_first = in_channel_first
        ctx.out_channel_first = out_channel_first
        ctx.one_by_one = one_by_one
        ctx.scans = scans
        if one_by_one:
            if in_channel_first:
                B, K, C, H, W = x.shape
            else:
                B, H, W, K, C = x.shape
        else:
            if in_channel_first:
                B, C, H, W = x.shape
            else:
                B, H, W, C = x.shape
        ctx.shape = (B, C, H, W)
        return _scan_any_layout(x, in_channel_first, out_channel_first, one_by_one, scans)

    @staticmethod
    def backward(ctx, ys: torch.Tensor):
        B, C, H, W = ctx.shape
        ys = ys.reshape(B, -1, C, H, W) if ctx.out_channel_first else ys.reshape(B, H, W, -1, C)
        y = _merge_any_layout(ys, ctx.in_channel_first, ctx.out_channel_first, ctx.one_by_one, ctx.scans, H, W)
        if ctx.one_by_one:
            y = y.view(B, 4, -1, H, W) if ctx.in_channel_first else y.view(B, H, W, 4, -1)
        else:
            y = y.view(B, -1, H, W) if ctx.in_channel_first else y.view(B, H, W, -1)
        return y, None, None, None, None


class CrossMergeF(torch.autograd.Function):
    """models/csm_triton.py:228-273"""

    @staticmethod
    def forward(ctx, ys: torch.Tensor, in_channel_first=True, out_channel_first=True, one_by_one=False, scans=0):
        ctx.in_channel_first = in_channel_first
        ctx.out_channel_first = out_channel_first
        ctx.one_by_one = one_by_one
        ctx.scans = scans
        if out_channel_first:
            B, K, C, H, W = ys.shape
        else:
            B, H, W, K, C = ys.shape
        ctx.shape = (B, C, H, W)
        return _merge_any_layout(ys, in_channel_first, out_channel_first, one_by_one, scans, H, W)

    @staticmethod
    def backward(ctx, x: torch.Tensor):
        B, C, H, W = ctx.shape
        if not ctx.one_by_one:
            x = x.reshape(B, C, H, W) if ctx.in_channel_first else x.reshape(B, H, W, C)
        else:
            x = x.reshape(B, 4, C, H, W) if ctx.in_channel_first else x.reshape(B, H, W, 4, C)
        x = _scan_any_layout(x, ctx.in_channel_first, ctx.out_channel_first, ctx.one_by_one, ctx.scans)
        x = x.view(B, 4, C, H, W) if ctx.out_channel_first else x.view(B, H, W, 4, C)
        return x, None, None, None, None


# upstream VMamba (and BASELINE.json's wording) call these CrossScan / CrossMerge
CrossScan = CrossScanF
CrossMerge = CrossMergeF


def cross_scan_fn(x: torch.Tensor, in_channel_first=True, out_channel_first=True, one_by_one=False, scans=0,
                  force_torch=False):
    """models/csm_triton.py:501-507.  x: (B,C,H,W) | (B,H,W,C) | (B,4,C,H,W) | (B,H,W,4,C) -> (B,4,C,L) | (B,L,4,C)"""
    return CrossScanF.apply(x, in_channel_first, out_channel_first, one_by_one, scans)


def cross_merge_fn(y: torch.Tensor, in_channel_first=True, out_channel_first=True, one_by_one=False, scans=0,
                   force_torch=False):
    """models/csm_triton.py:511-517.  y: (B,4,C,H,W) | (B,H,W,4,C) -> (B,C,L) | (B,L,C) | (B,4,C,L) | (B,L,4,C)"""
    return CrossMergeF.apply(y, in_channel_first, out_channel_first, one_by_one, scans)
