"""Low-rank delta projection of the SS2D core on sm_100a (reference ``F.conv1d(dts_r, dt_projs_weight, groups=K)``,
``models/fusion_vmamba.py:1155-1157``; ``torch.einsum("b k r l, k d r -> b k d l")`` at ``:818``).  delta is the largest
stream of the scan, so the projection is bound by writing it; one CUDA launch at that bound replaces cuDNN's
launch-per-group implicit GEMM.  ``z`` may be the ``torch.split`` slice of the x_proj output (no copy is made).  CUDA only."""
from __future__ import annotations

import torch

from . import _lib

__all__ = ["dt_proj", "DtProjFn"]


def _rows_contiguous(z):
    B, K, R, L = z.shape
    return z.stride(3) == 1 and (z.stride(2) == L or R == 1)


class DtProjFn(torch.autograd.Function):
    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda")
    def forward(ctx, z, weight):
        dev = _lib.require_cuda(z, weight)
        B, K, R, L = z.shape
        D = weight.shape[1]
        if not _rows_contiguous(z):
            z = z.contiguous()
        w = weight.float().contiguous()
        out = torch.empty((B, K * D, L), dtype=z.dtype, device=dev)
        if out.numel():
            with torch.cuda.device(dev):
                rc = _lib.lib().xfs_dt_proj_fwd(_lib.ptr(z), _lib.ptr(w), _lib.ptr(out), B, K, D, R, L, z.stride(0), z.stride(1),
                                                _lib.dtype_code(z), _lib.stream(dev))
            _lib.check(rc, "dt_proj_fwd")
        if any(ctx.needs_input_grad):
            ctx.save_for_backward(z, w)
            ctx.wdtype = weight.dtype
        return out

    @staticmethod
    @torch.amp.custom_bwd(device_type="cuda")
    def backward(ctx, g):
        z, w = ctx.saved_tensors
        B, K, R, L = z.shape
        D = w.shape[1]
        need_z, need_w = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        g = g.contiguous()
        L_ = _lib.lib()
        if (g.dtype == torch.float32 and z.dtype == torch.float32
                and L_.xfs_dt_proj_bwd_supported(R, L, z.stride(0), z.stride(1), _lib.dtype_code(z))):
            # hand-written: one pass over g for each of dz = W^T g and dW = sum_{b,l} g z^T (3xTF32 warp MMA, csrc/dtproj.cu)
            dev = g.device
            dz = torch.empty((B, K, R, L), dtype=torch.float32, device=dev) if need_z else None
            dw = torch.zeros((K, D, R), dtype=torch.float32, device=dev) if need_w else None
            if need_z or need_w:
                with torch.cuda.device(dev):
                    rc = L_.xfs_dt_proj_bwd(_lib.ptr(g), _lib.ptr(z), _lib.ptr(w), _lib.ptr(dz) if need_z else None,
                                            _lib.ptr(dw) if need_w else None, B, K, D, R, L, z.stride(0), z.stride(1),
                                            _lib.dtype_code(z), _lib.stream(dev))
                _lib.check(rc, "dt_proj_bwd")
            return dz, (dw.to(ctx.wdtype) if need_w else None)
        # 16-bit rows: two plain batched GEMMs (library work): dz = W^T g over D, dW = sum_b g z^T over L then the batch
        g4 = g.to(z.dtype).view(B, K, D, L)
        dz = torch.matmul(w.transpose(1, 2).unsqueeze(0).to(g4.dtype), g4) if need_z else None
        dw = torch.matmul(g4, z.transpose(2, 3)).float().sum(0).to(ctx.wdtype) if need_w else None
        return dz, dw


def dt_proj(z, weight):
    """z: (B, K, R, L); weight: (K, D, R) -> delta (B, K*D, L) in z's dtype"""
    if z.dim() != 4 or weight.dim() != 3 or weight.shape[0] != z.shape[1] or weight.shape[2] != z.shape[2]:
        raise RuntimeError(f"dt_proj expects z (B, K, R, L) and weight (K, D, R); got {tuple(z.shape)}, {tuple(weight.shape)}")
    if z.shape[2] > 64:
        raise RuntimeError("dt_proj: dt_rank > 64 is not supported")
    return DtProjFn.apply(z, weight)
