"""ctypes front end of ``libxfscan_oracle.so`` plus numpy restatements of the index routes.

All functions take/return numpy arrays.  ``real`` selects the arithmetic of the scan:
``"f32"`` mirrors ``selective_scan_torch`` (models/csms6s.py:52 casts everything to float32),
``"f64"`` is the tie-breaker used to judge which of two fp32 answers is closer to the truth.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from pathlib import Path

import numpy as np

__all__ = [
    "build", "lib", "set_threads", "route_table", "cross_scan", "cross_merge", "cross_merge_1b1", "swap_scan", "swap_merge",
    "selective_scan_fwd", "selective_scan_bwd", "ss2d_fwd", "ss2d_bwd", "bf16_round", "np_cross_scan", "layernorm2d",
    "dwconv3x3_silu", "dwconv3x3_silu_bwd", "dt_proj", "dt_proj_bwd",
]

_HERE = Path(__file__).resolve().parent
_SO = _HERE / "libxfscan_oracle.so"
_lib = None


def build(force: bool = False) -> Path:
    """Compile the C oracle with gcc (``make -C oracle``)."""
    src = _HERE / "xfscan_oracle.c"
    if force or (not _SO.exists()) or _SO.stat().st_mtime < src.stat().st_mtime:
        subprocess.run(["make", "-C", str(_HERE), "-s", "-B"], check=True, capture_output=True)
    return _SO


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(str(_SO))
        _lib.xfo_route_index.restype = ctypes.c_int64
        _lib.xfo_route_index.argtypes = [ctypes.c_int, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64, ctypes.c_int]
    return _lib


def set_threads(n: int = 0) -> int:
    """OpenMP threads of the oracle's loops (0: leave as is); returns the count in effect (1 without OpenMP)."""
    L = lib()
    L.xfo_set_threads.restype = ctypes.c_int
    L.xfo_set_threads.argtypes = [ctypes.c_int]
    return int(L.xfo_set_threads(int(n)))


def _p(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def _i64(v):
    return ctypes.c_int64(int(v))


# ------------------------------------------------------------------------------------------------
# routes (models/csm_triton.py:22-53)
# ------------------------------------------------------------------------------------------------
def route_table(H: int, W: int, scans: int = 0) -> np.ndarray:
    """(4, L) int64: spatial offset h*W+w read by scan position l of direction k (pure numpy)."""
    L = H * W
    nat = np.arange(L, dtype=np.int64)
    if scans == 1:
        return np.stack([nat, nat, nat, nat])
    if scans == 2:
        return np.stack([nat, nat, nat[::-1], nat[::-1]])
    tr = nat.reshape(H, W).T.reshape(-1)          # x.transpose(2, 3).flatten(2, 3)
    return np.stack([nat, tr, nat[::-1], tr[::-1]])


def np_cross_scan(x: np.ndarray, scans: int = 0) -> np.ndarray:
    """numpy restatement (fancy indexing) -- independent of the C code, used to cross-check it."""
    B, C, H, W = x.shape
    r = route_table(H, W, scans)
    return np.ascontiguousarray(x.reshape(B, 1, C, H * W)[:, 0][:, None, :, :].repeat(4, 1)[
        np.arange(B)[:, None, None, None], np.arange(4)[None, :, None, None],
        np.arange(C)[None, None, :, None], r[None, :, None, :]])


def cross_scan(x: np.ndarray, scans: int = 0, one_by_one: bool = False) -> np.ndarray:
    x = np.ascontiguousarray(x)
    if one_by_one:
        B, K, C, H, W = x.shape
        assert K == 4
    else:
        B, C, H, W = x.shape
    out = np.empty((B, 4, C, H * W), dtype=x.dtype)
    lib().xfo_cross_scan(_p(x), _p(out), _i64(B), _i64(C), _i64(H), _i64(W), ctypes.c_int(x.dtype.itemsize),
                         ctypes.c_int(scans), ctypes.c_int(int(one_by_one)))
    return out


def cross_merge(ys: np.ndarray, H: int, W: int, scans: int = 0) -> np.ndarray:
    """ys (B,4,C,L) float32/float64 -> (B,C,L); add order of models/csm_triton.py:61-62."""
    ys = np.ascontiguousarray(ys)
    B, K, C, L = ys.shape
    assert K == 4 and L == H * W
    out = np.empty((B, C, L), dtype=ys.dtype)
    fn = {np.dtype(np.float32): lib().xfo_cross_merge_f32, np.dtype(np.float64): lib().xfo_cross_merge_f64}[ys.dtype]
    fn(_p(ys), _p(out), _i64(B), _i64(C), _i64(H), _i64(W), ctypes.c_int(scans))
    return out


def cross_merge_1b1(ys: np.ndarray, H: int, W: int, scans: int = 0) -> np.ndarray:
    """cross_merge1b1_fwd (models/csm_triton.py:134-179): inverse routing of each direction, no adds."""
    B, K, C, L = ys.shape
    r = route_table(H, W, scans)
    out = np.empty_like(ys)
    for k in range(4):
        out[:, k][..., r[k]] = ys[:, k]
    return out


def swap_scan(x: np.ndarray, x2: np.ndarray) -> np.ndarray:
    x = np.ascontiguousarray(x)
    x2 = np.ascontiguousarray(x2)
    B, C = x.shape[:2]
    L = int(np.prod(x.shape[2:]))
    out = np.empty((B, 2, C, L), dtype=x.dtype)
    lib().xfo_swap_scan(_p(x), _p(x2), _p(out), _i64(B), _i64(C), _i64(L), ctypes.c_int(x.dtype.itemsize))
    return out


def swap_merge(ys: np.ndarray):
    """SwappingMerge_multiview.forward (models/fusion_vmamba.py:226-232): plain split."""
    return np.ascontiguousarray(ys[:, 0]), np.ascontiguousarray(ys[:, 1])


# ------------------------------------------------------------------------------------------------
# selective scan (models/csms6s.py:25-68)
# ------------------------------------------------------------------------------------------------
def _real(real):
    return (np.float32, "_f32") if real == "f32" else (np.float64, "_f64")


def _c(a, dt):
    return None if a is None else np.ascontiguousarray(np.asarray(a, dtype=dt))


def selective_scan_fwd(u, delta, A, B, C, D=None, delta_bias=None, delta_softplus=True, real="f32",
                       return_last_state=False):
    dt, sfx = _real(real)
    u, delta, A, B, C, D, delta_bias = (_c(t, dt) for t in (u, delta, A, B, C, D, delta_bias))
    Bsz, KD, L = u.shape
    _, K, N, _ = B.shape
    assert delta.shape == u.shape and A.shape == (KD, N) and C.shape == B.shape and KD % K == 0
    out = np.empty_like(u)
    last = np.empty((Bsz, KD, N), dtype=dt) if return_last_state else None
    getattr(lib(), "xfo_selective_scan_fwd" + sfx)(
        _p(u), _p(delta), _p(A), _p(B), _p(C), _p(D), _p(delta_bias), ctypes.c_int(int(delta_softplus)),
        _p(out), _p(last), _i64(Bsz), _i64(KD), _i64(K), _i64(N), _i64(L))
    return (out, last) if return_last_state else out


def selective_scan_bwd(u, delta, A, B, C, D, delta_bias, dout, delta_softplus=True, real="f64"):
    """returns (du, ddelta, dA, dB, dC, dD, ddelta_bias); dD / ddelta_bias are None when the input was None."""
    dt, sfx = _real(real)
    u, delta, A, B, C, D, delta_bias, dout = (_c(t, dt) for t in (u, delta, A, B, C, D, delta_bias, dout))
    Bsz, KD, L = u.shape
    _, K, N, _ = B.shape
    du, ddelta = np.empty_like(u), np.empty_like(u)
    dA, dB, dC = np.zeros_like(A), np.zeros_like(B), np.zeros_like(C)
    dD = np.zeros((KD,), dtype=dt) if D is not None else None
    dbias = np.zeros((KD,), dtype=dt) if delta_bias is not None else None
    getattr(lib(), "xfo_selective_scan_bwd" + sfx)(
        _p(u), _p(delta), _p(A), _p(B), _p(C), _p(D), _p(delta_bias), ctypes.c_int(int(delta_softplus)), _p(dout),
        _p(du), _p(ddelta), _p(dA), _p(dB), _p(dC), _p(dD), _p(dbias),
        _i64(Bsz), _i64(KD), _i64(K), _i64(N), _i64(L))
    return du, ddelta, dA, dB, dC, dD, dbias


# ------------------------------------------------------------------------------------------------
# SS2D core composition (models/fusion_vmamba.py:1145-1174): cross_scan -> scan -> cross_merge
# ------------------------------------------------------------------------------------------------
def ss2d_fwd(x, delta, A, Bs, Cs, Ds=None, delta_bias=None, delta_softplus=True, real="f32", scans=0):
    """x (B,D,H,W); delta (B,4D,L); A (4D,N); Bs,Cs (B,4,N,L) -> y (B,D,L)"""
    dt, _ = _real(real)
    x = _c(x, dt)
    Bsz, Dm, H, W = x.shape
    xs = cross_scan(x, scans).reshape(Bsz, 4 * Dm, H * W)
    ys = selective_scan_fwd(xs, delta, A, Bs, Cs, Ds, delta_bias, delta_softplus, real)
    return cross_merge(ys.reshape(Bsz, 4, Dm, H * W), H, W, scans)


def ss2d_bwd(x, delta, A, Bs, Cs, Ds, delta_bias, dy, delta_softplus=True, real="f64", scans=0):
    """gradients of ss2d_fwd wrt (x, delta, A, Bs, Cs, Ds, delta_bias) for upstream dy (B,D,L).

    CrossMergeF.backward is a cross-scan of dy (models/csm_triton.py:249-273) and CrossScanF.backward a
    cross-merge of the per-direction du (models/csm_triton.py:208-225)."""
    dt, _ = _real(real)
    x, dy = _c(x, dt), _c(dy, dt)
    Bsz, Dm, H, W = x.shape
    L = H * W
    xs = cross_scan(x, scans).reshape(Bsz, 4 * Dm, L)
    dys = cross_scan(dy.reshape(Bsz, Dm, H, W), scans).reshape(Bsz, 4 * Dm, L)
    du, ddelta, dA, dB, dC, dD, dbias = selective_scan_bwd(xs, delta, A, Bs, Cs, Ds, delta_bias, dys, delta_softplus, real)
    dx = cross_merge(du.reshape(Bsz, 4, Dm, L), H, W, scans).reshape(Bsz, Dm, H, W)
    return dx, ddelta, dA, dB, dC, dD, dbias


# ------------------------------------------------------------------------------------------------
def layernorm2d(x, weight=None, bias=None, eps=1e-5):
    """LayerNorm2d (models/fusion_vmamba.py:52-57): layer_norm over C of a channel-first (B, C, ...) array, in float64"""
    x64 = np.asarray(x, dtype=np.float64)
    mean = x64.mean(axis=1, keepdims=True)
    var = x64.var(axis=1, keepdims=True)
    y = (x64 - mean) / np.sqrt(var + eps)
    shp = (1, -1) + (1,) * (x64.ndim - 2)
    if weight is not None:
        y = y * np.asarray(weight, dtype=np.float64).reshape(shp)
    if bias is not None:
        y = y + np.asarray(bias, dtype=np.float64).reshape(shp)
    return y


def _dw_shift(xp, i, j, H, W):
    return xp[:, :, i:i + H, j:j + W]


def dwconv3x3_silu(x, weight, bias=None, act=True):
    """``act(conv2d(x))`` with a depthwise 3x3, padding-1 convolution (nn.Conv2d(groups=d_inner) + SiLU,
    models/fusion_vmamba.py:405-413,1199-1200), in float64.  x (B, C, H, W); weight (C, 1, 3, 3); bias (C) or None."""
    x64 = np.asarray(x, dtype=np.float64)
    w = np.asarray(weight, dtype=np.float64).reshape(-1, 3, 3)
    B, C, H, W = x64.shape
    xp = np.pad(x64, ((0, 0), (0, 0), (1, 1), (1, 1)))
    pre = np.zeros_like(x64)
    for i in range(3):
        for j in range(3):
            pre += _dw_shift(xp, i, j, H, W) * w[None, :, i, j, None, None]
    if bias is not None:
        pre += np.asarray(bias, dtype=np.float64)[None, :, None, None]
    return pre / (1.0 + np.exp(-pre)) if act else pre


def dwconv3x3_silu_bwd(x, weight, bias, dy, act=True):
    """gradients of dwconv3x3_silu: returns (dx, dweight (C,1,3,3), dbias (C)) in float64"""
    x64 = np.asarray(x, dtype=np.float64)
    w = np.asarray(weight, dtype=np.float64).reshape(-1, 3, 3)
    B, C, H, W = x64.shape
    pre = dwconv3x3_silu(x64, w, bias, act=False)
    g = np.asarray(dy, dtype=np.float64)
    if act:
        s = 1.0 / (1.0 + np.exp(-pre))
        g = g * s * (1.0 + pre * (1.0 - s))
    xp = np.pad(x64, ((0, 0), (0, 0), (1, 1), (1, 1)))
    gp = np.pad(g, ((0, 0), (0, 0), (1, 1), (1, 1)))
    dx = np.zeros_like(x64)
    dw = np.zeros((C, 3, 3))
    for i in range(3):
        for j in range(3):
            dx += _dw_shift(gp, 2 - i, 2 - j, H, W) * w[None, :, i, j, None, None]
            dw[:, i, j] = (g * _dw_shift(xp, i, j, H, W)).sum(axis=(0, 2, 3))
    return dx, dw.reshape(C, 1, 3, 3), g.sum(axis=(0, 2, 3))


def dt_proj(z, weight):
    """delta[b, k*D + d, l] = sum_r W[k, d, r] z[b, k, r, l]  (F.conv1d(dts_r, dt_projs_weight, groups=K),
    models/fusion_vmamba.py:1155-1157), float64.  z (B, K, R, L); weight (K, D, R) -> (B, K*D, L)"""
    z64, w64 = np.asarray(z, dtype=np.float64), np.asarray(weight, dtype=np.float64)
    B, K, R, L = z64.shape
    return np.einsum("bkrl,kdr->bkdl", z64, w64).reshape(B, -1, L)


def dt_proj_bwd(z, weight, g):
    """gradients of dt_proj: (dz (B, K, R, L), dweight (K, D, R)), float64"""
    z64, w64 = np.asarray(z, dtype=np.float64), np.asarray(weight, dtype=np.float64)
    B, K, R, L = z64.shape
    g4 = np.asarray(g, dtype=np.float64).reshape(B, K, -1, L)
    return np.einsum("bkdl,kdr->bkrl", g4, w64), np.einsum("bkdl,bkrl->kdr", g4, z64)


def bf16_round(a: np.ndarray) -> np.ndarray:
    """round-to-nearest-even float32 -> bfloat16 -> float32 (what torch's .to(bfloat16).float() does)."""
    a = np.ascontiguousarray(a, dtype=np.float32)
    bits = a.view(np.uint32).astype(np.uint64)
    rounded = ((bits + 0x7FFF + ((bits >> 16) & 1)) & 0xFFFF0000).astype(np.uint32)
    out = rounded.view(np.float32).copy()
    nan = np.isnan(a)
    out[nan] = a[nan]
    return out
