"""The reference's own GPU path for the SS2D core, for a same-box context number and a GPU-side parity check.

TEST / BENCH INFRASTRUCTURE ONLY (imported by tests/ and bench.py's ``ref_gpu`` leg; nothing in xfmamba_b200/ uses it).

What runs: ``selective_scan_cuda_core`` -- the reference's CUDA extension (models/selective_scan/csrc/selective_scan/*), compiled
unmodified for sm_100 by ``oracle/build_ref.py`` into ``oracle/_ref/`` -- called exactly as the reference calls it
(models/csms6s.py:83 ``fwd(u, delta, A, B, C, D, delta_bias, delta_softplus, 1)``, :101 ``bwd(...)``), between a torch restatement
of the reference's CrossScan / CrossMerge (models/csm_triton.py:22-30, 56-62, the ``force_torch`` path; its Triton variant lives in
/root/reference, which does not travel to the GPU box).  Four HBM passes (scan, scan kernel, merge) against the one fused launch.
"""
from __future__ import annotations

import sys
from pathlib import Path

REF_DIR = Path(__file__).resolve().parent / "_ref"


def available() -> bool:
    return (REF_DIR / "selective_scan_cuda_core.so").exists()


def load():
    import torch  # noqa: F401  (the extension links against libtorch)
    if str(REF_DIR) not in sys.path:
        sys.path.insert(0, str(REF_DIR))
    import selective_scan_cuda_core
    return selective_scan_cuda_core


def cross_scan(x):
    """(B, D, H, W) -> (B, 4, D, L): row-major, column-major and their flips (csm_triton.py:22-30)"""
    B, D, H, W = x.shape
    y = x.new_empty(B, 4, D, H * W)
    y[:, 0] = x.flatten(2)
    y[:, 1] = x.transpose(2, 3).flatten(2)
    y[:, 2:4] = y[:, 0:2].flip(-1)
    return y


def cross_merge(ys, H, W):
    """(B, 4, D, L) -> (B, D, L) (csm_triton.py:56-62)"""
    B, K, D, L = ys.shape
    y = ys[:, 0:2] + ys[:, 2:4].flip(-1)
    return y[:, 0] + y[:, 1].reshape(B, D, W, H).transpose(2, 3).contiguous().view(B, D, L)


def ss2d_fwd(m, x, delta, A, Bs, Cs, Ds, bias):
    B, D, H, W = x.shape
    xs = cross_scan(x).view(B, 4 * D, H * W)
    out, st, *_ = m.fwd(xs, delta, A, Bs, Cs, Ds, bias, True, 1)
    return cross_merge(out.view(B, 4, D, H * W), H, W), (xs, st)


def ss2d_bwd(m, saved, dy, x_shape, delta, A, Bs, Cs, Ds, bias):
    B, D, H, W = x_shape
    xs, st = saved
    dys = cross_scan(dy.view(B, D, H, W)).view(B, 4 * D, H * W)          # CrossMerge backward = cross-scan of dy
    du, ddelta, dA, dB, dC, dD, dbias, *_ = m.bwd(xs, delta, A, Bs, Cs, Ds, bias, dys, st, True, 1)
    dx = cross_merge(du.view(B, 4, D, H * W), H, W).view(B, D, H, W)      # CrossScan backward = merge of du
    return dx, ddelta, dA, dB, dC, dD, dbias


def time_config(d, H, W, iters=10, warmup=3):
    """d: the bench's device tensors (x (B, D, H, W) fp32, delta, A, Bs, Cs, Ds, delta_bias).  Returns a dict of ms per step:
    whole forward / backward of the reference path and the share of its CUDA scan kernel alone, plus the last results."""
    import torch
    m = load()
    B, D = d["x"].shape[0], d["x"].shape[1]
    L = H * W
    x = d["x"].view(B, D, H, W)
    delta, A, Bs, Cs, Ds, bias = d["delta"], d["A"], d["Bs"], d["Cs"], d["Ds"], d["delta_bias"]
    dy = torch.randn(B, D, L, device=x.device)
    ev = lambda: torch.cuda.Event(enable_timing=True)
    t = dict(fwd_ms=0.0, bwd_ms=0.0, scan_fwd_ms=0.0, scan_bwd_ms=0.0)
    y = grads = None
    for i in range(warmup + iters):
        e = [ev() for _ in range(7)]
        e[0].record()
        xs = cross_scan(x).view(B, 4 * D, L)
        e[1].record()
        out, st, *_ = m.fwd(xs, delta, A, Bs, Cs, Ds, bias, True, 1)
        e[2].record()
        y = cross_merge(out.view(B, 4, D, L), H, W)
        e[3].record()
        dys = cross_scan(dy.view(B, D, H, W)).view(B, 4 * D, L)
        e[4].record()
        du, ddelta, dA, dB, dC, dD, dbias, *_ = m.bwd(xs, delta, A, Bs, Cs, Ds, bias, dys, st, True, 1)
        e[5].record()
        dx = cross_merge(du.view(B, 4, D, L), H, W).view(B, D, H, W)
        e[6].record()
        torch.cuda.synchronize()
        grads = (dx, ddelta, dA, dB, dC, dD, dbias)
        if i >= warmup:
            t["fwd_ms"] += e[0].elapsed_time(e[3]) / iters
            t["bwd_ms"] += e[3].elapsed_time(e[6]) / iters
            t["scan_fwd_ms"] += e[1].elapsed_time(e[2]) / iters
            t["scan_bwd_ms"] += e[4].elapsed_time(e[5]) / iters
        del xs, out, st, dys, du
    return t, y, dy, grads
