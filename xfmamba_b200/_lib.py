"""ctypes binding of ``libxfscan.so`` (C ABI in ``include/xfscan.h``).

PyTorch is only the allocator / stream provider here: every call passes raw device pointers and the current CUDA
stream.  There is NO CPU path and NO fallback: if the shared library is missing or the tensors are not on a CUDA
device the call raises.
"""
from __future__ import annotations

import ctypes
import os
from pathlib import Path

import torch

_PKG = Path(__file__).resolve().parent
_SO = Path(os.environ["XFMAMBA_B200_LIB"]) if os.environ.get("XFMAMBA_B200_LIB") else _PKG / "libxfscan.so"
_lib = None

F32, BF16, F16 = 0, 1, 2
_DTYPES = {torch.float32: F32, torch.bfloat16: BF16, torch.float16: F16}

c_i64, c_i32, c_vp = ctypes.c_int64, ctypes.c_int32, ctypes.c_void_p


class ScanFwdArgs(ctypes.Structure):
    _fields_ = [(n, c_vp) for n in ("u", "delta", "A", "B", "C", "D", "delta_bias", "out", "states")] + \
               [(n, c_i64) for n in ("batch", "dim", "dstate", "seqlen", "ngroups")] + \
               [(n, c_i32) for n in ("dtype", "out_dtype", "delta_softplus", "reserved")]


class ScanBwdArgs(ctypes.Structure):
    _fields_ = [(n, c_vp) for n in ("u", "delta", "A", "B", "C", "D", "delta_bias", "dout", "states", "du", "ddelta",
                                    "dA", "dB", "dC", "dD", "ddelta_bias")] + \
               [(n, c_i64) for n in ("batch", "dim", "dstate", "seqlen", "ngroups")] + \
               [(n, c_i32) for n in ("dtype", "dout_dtype", "delta_softplus", "acc_replicas")]


class Ss2dFwdArgs(ctypes.Structure):
    _fields_ = [(n, c_vp) for n in ("x", "delta", "A", "Bs", "Cs", "Ds", "delta_bias", "y", "states")] + \
               [(n, c_i64) for n in ("batch", "D", "N", "H", "W")] + \
               [(n, c_i32) for n in ("dtype", "out_dtype", "delta_softplus", "scans")]


class Ss2dBwdArgs(ctypes.Structure):
    _fields_ = [(n, c_vp) for n in ("x", "delta", "A", "Bs", "Cs", "Ds", "delta_bias", "dy", "states", "dx", "ddelta",
                                    "dA", "dBs", "dCs", "dDs", "ddelta_bias")] + \
               [(n, c_i64) for n in ("batch", "D", "N", "H", "W")] + \
               [(n, c_i32) for n in ("dtype", "dout_dtype", "delta_softplus", "scans", "acc_replicas", "reserved")]


_P3 = c_vp * 3


class X3FwdArgs(ctypes.Structure):
    _fields_ = [(n, _P3) for n in ("x", "delta", "Bs", "Cs", "y", "states")] + [(n, c_vp) for n in ("A", "Ds", "delta_bias")] + \
               [(n, c_i64) for n in ("batch", "D", "N", "H", "W")] + \
               [(n, c_i32) for n in ("dtype", "out_dtype", "delta_softplus", "nstreams")]


class X3BwdArgs(ctypes.Structure):
    _fields_ = [(n, _P3) for n in ("x", "delta", "Bs", "Cs", "dy", "dx", "ddelta", "dBs", "dCs")] + \
               [(n, c_vp) for n in ("A", "Ds", "delta_bias", "dA", "dDs", "ddelta_bias")] + \
               [(n, c_i64) for n in ("batch", "D", "N", "H", "W")] + \
               [(n, c_i32) for n in ("dtype", "dout_dtype", "delta_softplus", "nstreams")]


class SwapFusedFwdArgs(ctypes.Structure):
    _fields_ = [(n, c_vp) for n in ("x", "x2", "delta", "A", "Bs", "Cs", "Ds", "delta_bias", "y", "y2", "states")] + \
               [(n, c_i64) for n in ("batch", "D", "N", "L")] + \
               [(n, c_i32) for n in ("dtype", "out_dtype", "delta_softplus", "reserved")]


class SwapFusedBwdArgs(ctypes.Structure):
    _fields_ = [(n, c_vp) for n in ("x", "x2", "delta", "A", "Bs", "Cs", "Ds", "delta_bias", "dy", "dy2", "dx", "dx2", "ddelta",
                                    "dA", "dBs", "dCs", "dDs", "ddelta_bias")] + \
               [(n, c_i64) for n in ("batch", "D", "N", "L")] + \
               [(n, c_i32) for n in ("dtype", "dout_dtype", "delta_softplus", "reserved")]


def p3(tensors):
    """three device pointers (missing / None entries stay NULL)"""
    arr = _P3()
    for i, t in enumerate(tensors):
        arr[i] = None if t is None else t.data_ptr()
    return arr


# every symbol include/xfscan.h declares (tests/test_cabi.py checks the .so exports all of them)
SYMBOLS = [
    "xfs_version", "xfs_error_string", "xfs_device_ok", "xfs_chunk_len", "xfs_num_chunks", "xfs_launch_count",
    "xfs_cross_scan", "xfs_cross_merge", "xfs_swap_scan", "xfs_swap_merge", "xfs_swap_stack",
    "xfs_selective_scan_fwd", "xfs_selective_scan_bwd", "xfs_ss2d_supported", "xfs_ss2d_states_len", "xfs_ss2d_fwd", "xfs_ss2d_bwd",
    "xfs_layernorm2d_fwd", "xfs_layernorm2d_bwd",
    "xfs_dwconv3x3_supported", "xfs_dwconv3x3_fwd", "xfs_dwconv3x3_bwd", "xfs_dt_proj_fwd", "xfs_dt_proj_bwd", "xfs_dt_proj_bwd_supported",
    "xfs_cross_ss2d_x3_supported", "xfs_cross_ss2d_x3_fwd", "xfs_cross_ss2d_x3_bwd",
    "xfs_swap_scan_fused_supported", "xfs_swap_scan_fused_fwd", "xfs_swap_scan_fused_bwd",
]


def library_path() -> Path:
    return _SO


def lib() -> ctypes.CDLL:
    """Loads the library; raises (never falls back) when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not _SO.exists():
        raise ImportError(
            f"{_SO} is missing: build it with `python -m xfmamba_b200.build` (needs nvcc, targets sm_100a). "
            "xfmamba_b200 has no CPU or PyTorch fallback.")
    L = ctypes.CDLL(str(_SO))
    L.xfs_version.restype = ctypes.c_int
    L.xfs_error_string.restype = ctypes.c_char_p
    L.xfs_error_string.argtypes = [ctypes.c_int]
    L.xfs_device_ok.argtypes = [ctypes.c_int]
    L.xfs_chunk_len.restype = c_i64
    L.xfs_num_chunks.restype = c_i64
    L.xfs_num_chunks.argtypes = [c_i64]
    L.xfs_launch_count.restype = c_i64
    for name in ("xfs_cross_scan", "xfs_cross_merge"):
        getattr(L, name).argtypes = [c_vp, c_vp, c_i64, c_i64, c_i64, c_i64, ctypes.c_int, ctypes.c_int, ctypes.c_int, c_vp]
    L.xfs_swap_scan.argtypes = [c_vp, c_vp, c_vp, c_i64, c_i64, c_i64, ctypes.c_int, c_vp]
    L.xfs_swap_merge.argtypes = [c_vp, c_vp, c_vp, c_i64, c_i64, c_i64, ctypes.c_int, c_vp]
    L.xfs_swap_stack.argtypes = [c_vp, c_vp, c_vp, c_i64, c_i64, c_i64, ctypes.c_int, c_vp]
    L.xfs_selective_scan_fwd.argtypes = [ctypes.POINTER(ScanFwdArgs), c_vp]
    L.xfs_selective_scan_bwd.argtypes = [ctypes.POINTER(ScanBwdArgs), c_vp]
    L.xfs_ss2d_supported.argtypes = [c_i64, c_i64, c_i64, c_i64, ctypes.c_int, ctypes.c_int]
    L.xfs_ss2d_states_len.restype = c_i64
    L.xfs_ss2d_states_len.argtypes = [c_i64, c_i64, c_i64, ctypes.c_int, ctypes.c_int]
    L.xfs_ss2d_fwd.argtypes = [ctypes.POINTER(Ss2dFwdArgs), c_vp]
    L.xfs_ss2d_bwd.argtypes = [ctypes.POINTER(Ss2dBwdArgs), c_vp]
    L.xfs_layernorm2d_fwd.argtypes = [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_i64, c_i64, c_i64, ctypes.c_float, ctypes.c_int, c_vp]
    L.xfs_layernorm2d_bwd.argtypes = [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_i64, c_i64, c_i64, ctypes.c_int, c_vp]
    L.xfs_dwconv3x3_supported.argtypes = [c_i64, c_i64, ctypes.c_int]
    L.xfs_dwconv3x3_fwd.argtypes = [c_vp, c_vp, c_vp, c_vp, c_i64, c_i64, c_i64, c_i64, ctypes.c_int, ctypes.c_int, c_vp]
    L.xfs_dwconv3x3_bwd.argtypes = [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_i64, c_i64, c_i64, c_i64, ctypes.c_int, ctypes.c_int, c_vp]
    L.xfs_dt_proj_fwd.argtypes = [c_vp, c_vp, c_vp, c_i64, c_i64, c_i64, c_i64, c_i64, c_i64, c_i64, ctypes.c_int, c_vp]
    L.xfs_dt_proj_bwd.argtypes = [c_vp, c_vp, c_vp, c_vp, c_vp, c_i64, c_i64, c_i64, c_i64, c_i64, c_i64, c_i64, ctypes.c_int, c_vp]
    L.xfs_dt_proj_bwd_supported.argtypes = [c_i64, c_i64, c_i64, c_i64, ctypes.c_int]
    L.xfs_cross_ss2d_x3_supported.argtypes = [c_i64, c_i64, c_i64]
    L.xfs_cross_ss2d_x3_fwd.argtypes = [ctypes.POINTER(X3FwdArgs), c_vp]
    L.xfs_cross_ss2d_x3_bwd.argtypes = [ctypes.POINTER(X3BwdArgs), c_vp]
    L.xfs_swap_scan_fused_supported.argtypes = [c_i64, c_i64]
    L.xfs_swap_scan_fused_fwd.argtypes = [ctypes.POINTER(SwapFusedFwdArgs), c_vp]
    L.xfs_swap_scan_fused_bwd.argtypes = [ctypes.POINTER(SwapFusedBwdArgs), c_vp]
    _lib = L
    return L


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = lib().xfs_error_string(rc).decode()
        raise RuntimeError(f"{what}: {msg} (code {rc})")


def dtype_code(t: torch.Tensor) -> int:
    try:
        return _DTYPES[t.dtype]
    except KeyError:
        raise RuntimeError(f"xfmamba_b200: unsupported dtype {t.dtype}; expected float32, bfloat16 or float16") from None


def require_cuda(*tensors: torch.Tensor) -> torch.device:
    dev = None
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError("xfmamba_b200 operators run on CUDA tensors only (sm_100a kernels, no CPU fallback); "
                               f"got a tensor on {t.device}")
        if dev is None:
            dev = t.device
        elif t.device != dev:
            raise RuntimeError(f"xfmamba_b200: tensors on different devices ({dev} vs {t.device})")
    return dev


def ptr(t):
    return None if t is None else c_vp(t.data_ptr())


def stream(dev: torch.device):
    return c_vp(torch.cuda.current_stream(dev).cuda_stream)


def num_chunks(L: int) -> int:
    return int(lib().xfs_num_chunks(int(L)))


def ss2d_states_len(N: int, H: int, W: int, dtype, out_dtype) -> int:
    """floats per (batch, 4*D) row of the fused SS2D checkpoints (layout private to the fwd / bwd kernel pair)"""
    return int(lib().xfs_ss2d_states_len(int(N), int(H), int(W), _DTYPES[dtype], _DTYPES[out_dtype]))


def launch_count() -> int:
    return int(lib().xfs_launch_count())


def grad_needed(*tensors) -> bool:
    """Evaluated by the operator wrappers BEFORE ``Function.apply``: inside ``forward`` grad mode is always off and
    ``ctx.needs_input_grad`` only mirrors ``requires_grad`` of the inputs, which is True for parameters under ``torch.no_grad()``
    as well -- inference would then write and keep the chunk states / statistics that only a backward pass reads."""
    import torch
    return torch.is_grad_enabled() and any(isinstance(t, torch.Tensor) and t.requires_grad for t in tensors)
