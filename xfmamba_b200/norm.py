"""LayerNorm2d on sm_100a: LayerNorm over C of a channel-first (B, C, H, W) tensor without the permute/copy round trip
of the reference's ``LayerNorm2d`` (``models/fusion_vmamba.py:52-57``).  Same parameters (``weight``, ``bias``, ``eps``),
same result as ``F.layer_norm(x.permute(0,2,3,1), (C,), w, b, eps).permute(0,3,1,2)``.  CUDA only."""
from __future__ import annotations

import torch

from . import _lib

__all__ = ["layer_norm_2d", "LayerNorm2dFn"]


class LayerNorm2dFn(torch.autograd.Function):
    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda")
    def forward(ctx, x, weight, bias, eps, need_grad=None):
        dev = _lib.require_cuda(x, weight, bias)
        if x.dim() < 3:
            raise RuntimeError(f"layer_norm_2d expects (B, C, ...) with at least one spatial dim; got {tuple(x.shape)}")
        B, C = x.shape[:2]
        HW = x[0, 0].numel()
        x = x.contiguous()
        w = None if weight is None else weight.float().contiguous()
        b = None if bias is None else bias.float().contiguous()
        y = torch.empty_like(x)
        need = any(ctx.needs_input_grad) if need_grad is None else bool(need_grad)     # see _lib.grad_needed
        mean = torch.empty((B, HW), dtype=torch.float32, device=dev) if need else None
        rstd = torch.empty((B, HW), dtype=torch.float32, device=dev) if need else None
        if x.numel():
            with torch.cuda.device(dev):
                rc = _lib.lib().xfs_layernorm2d_fwd(_lib.ptr(x), _lib.ptr(w), _lib.ptr(b), _lib.ptr(y), _lib.ptr(mean), _lib.ptr(rstd),
                                                    B, C, HW, float(eps), _lib.dtype_code(x), _lib.stream(dev))
            _lib.check(rc, "layernorm2d_fwd")
        if need:
            ctx.save_for_backward(x, w, mean, rstd)
            ctx.has_w, ctx.has_b = weight is not None, bias is not None
            ctx.wdtype = None if weight is None else weight.dtype
        return y

    @staticmethod
    @torch.amp.custom_bwd(device_type="cuda")
    def backward(ctx, dy):
        x, w, mean, rstd = ctx.saved_tensors
        dev = x.device
        B, C = x.shape[:2]
        HW = x[0, 0].numel()
        dy = dy.contiguous().to(x.dtype)
        dx = torch.empty_like(x)
        dw = torch.zeros(C, dtype=torch.float32, device=dev) if ctx.has_w else None
        db = torch.zeros(C, dtype=torch.float32, device=dev) if ctx.has_b else None
        if x.numel():
            with torch.cuda.device(dev):
                rc = _lib.lib().xfs_layernorm2d_bwd(_lib.ptr(x), _lib.ptr(dy), _lib.ptr(w), _lib.ptr(mean), _lib.ptr(rstd), _lib.ptr(dx),
                                                    _lib.ptr(dw), _lib.ptr(db), B, C, HW, _lib.dtype_code(x), _lib.stream(dev))
            _lib.check(rc, "layernorm2d_bwd")
        if dw is not None and ctx.wdtype is not None:
            dw = dw.to(ctx.wdtype)
        if db is not None and ctx.wdtype is not None:
            db = db.to(ctx.wdtype)
        return dx, dw, db, None, None


def layer_norm_2d(x, weight=None, bias=None, eps=1e-5):
    return LayerNorm2dFn.apply(x, weight, bias, eps, _lib.grad_needed(x, weight, bias))
