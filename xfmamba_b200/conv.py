"""Depthwise 3x3 convolution + SiLU on sm_100a: the producer of the scan input in every SS2D block
(``self.act(self.conv2d(x))`` with ``nn.Conv2d(d_inner, d_inner, 3, padding=1, groups=d_inner)``, reference
``models/fusion_vmamba.py:405-413,1199-1200``; fusion blocks ``:595-601``, ``:855-858``).  One kernel forward, one
backward (dx, dweight, dbias), instead of torch's convolution + activation + dgrad + wgrad + activation-backward passes.
Same result as ``F.silu(F.conv2d(x, weight, bias, padding=1, groups=C))``.  CUDA only."""
from __future__ import annotations

import torch
import torch.nn.functional as F

from . import _lib

__all__ = ["dwconv3x3_silu", "DwConvSiLUFn"]


class DwConvSiLUFn(torch.autograd.Function):
    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda")
    def forward(ctx, x, weight, bias, act):
        dev = _lib.require_cuda(x, weight, bias)
        B, C, H, W = x.shape
        x = x.contiguous()
        w = weight.float().contiguous()
        b = None if bias is None else bias.float().contiguous()
        y = torch.empty_like(x)
        if x.numel():
            with torch.cuda.device(dev):
                rc = _lib.lib().xfs_dwconv3x3_fwd(_lib.ptr(x), _lib.ptr(w), _lib.ptr(b), _lib.ptr(y), B, C, H, W, _lib.dtype_code(x),
                                                  int(act), _lib.stream(dev))
            _lib.check(rc, "dwconv3x3_fwd")
        if any(ctx.needs_input_grad):
            ctx.save_for_backward(x, w, b)
            ctx.act, ctx.has_b = bool(act), bias is not None
            ctx.wdtype, ctx.wshape = weight.dtype, weight.shape
        return y

    @staticmethod
    @torch.amp.custom_bwd(device_type="cuda")
    def backward(ctx, dy):
        x, w, b = ctx.saved_tensors
        dev = x.device
        B, C, H, W = x.shape
        dy = dy.contiguous().to(x.dtype)
        dx = torch.empty_like(x)
        part = torch.empty((B, C, 10), dtype=torch.float32, device=dev)
        if x.numel():
            with torch.cuda.device(dev):
                rc = _lib.lib().xfs_dwconv3x3_bwd(_lib.ptr(x), _lib.ptr(w), _lib.ptr(b), _lib.ptr(dy), _lib.ptr(dx), _lib.ptr(part),
                                                  B, C, H, W, _lib.dtype_code(x), int(ctx.act), _lib.stream(dev))
            _lib.check(rc, "dwconv3x3_bwd")
        else:
            part.zero_()
        tot = part.sum(0)                                  # (C, 10): 9 taps + bias
        dw = tot[:, :9].reshape(ctx.wshape).to(ctx.wdtype)
        db = tot[:, 9].to(ctx.wdtype) if ctx.has_b else None
        return dx, dw, db, None


def dwconv3x3_silu(x, weight, bias=None, act=True):
    """x: (B, C, H, W); weight: (C, 1, 3, 3); bias: (C) or None.  Planes too large for shared memory (beyond ~150 x 150 in
    the backward) are composed from torch's CUDA convolution instead -- still on the GPU, there is no CPU path."""
    _lib.require_cuda(x, weight, bias)
    if x.dim() != 4 or tuple(weight.shape) != (x.shape[1], 1, 3, 3):
        raise RuntimeError(f"dwconv3x3_silu expects x (B, C, H, W) and weight (C, 1, 3, 3); got {tuple(x.shape)}, {tuple(weight.shape)}")
    H, W = x.shape[2:]
    if not _lib.lib().xfs_dwconv3x3_supported(H, W, 1):
        y = F.conv2d(x, weight.to(x.dtype), None if bias is None else bias.to(x.dtype), padding=1, groups=x.shape[1])
        return F.silu(y) if act else y
    return DwConvSiLUFn.apply(x, weight, bias, act)
