"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol include/xfscan.h declares,
its argument structs have the layout the ctypes mirror assumes, and argument errors come back as negative codes
(no GPU is touched: every check below returns before a CUDA call)."""
import ctypes
import os
import re
import subprocess
import sys
import tempfile

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "xfscan.h")


@pytest.fixture(scope="module")
def L():
    from xfmamba_b200 import build, _lib
    build.build()           # nvcc cross-compiles for sm_100a without a GPU
    return _lib.lib()


def declared_symbols():
    src = open(HEADER).read()
    return sorted(set(re.findall(r"XFS_API[^;(]*?\b(xfs_\w+)\s*\(", src)))


def test_exports_every_declared_symbol(L):
    from xfmamba_b200 import _lib
    names = declared_symbols()
    assert len(names) >= 16
    assert sorted(_lib.SYMBOLS) == names, "ctypes mirror and header disagree"
    for n in names:
        assert hasattr(L, n), f"libxfscan.so does not export {n}"


def test_struct_layout_matches_header(L):
    from xfmamba_b200 import _lib
    prog = r'''
#include <stdio.h>
#include <stddef.h>
#include "xfscan.h"
int main(void) {
  printf("%zu %zu %zu %zu\n", sizeof(xfs_scan_fwd_args), sizeof(xfs_scan_bwd_args), sizeof(xfs_ss2d_fwd_args), sizeof(xfs_ss2d_bwd_args));
  printf("%zu %zu %zu %zu\n", offsetof(xfs_scan_fwd_args, batch), offsetof(xfs_scan_bwd_args, batch), offsetof(xfs_ss2d_fwd_args, batch), offsetof(xfs_ss2d_bwd_args, batch));
  printf("%zu %zu %zu %zu\n", offsetof(xfs_scan_fwd_args, dtype), offsetof(xfs_scan_bwd_args, dtype), offsetof(xfs_ss2d_fwd_args, dtype), offsetof(xfs_ss2d_bwd_args, dtype));
  return 0; }'''
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "t.c")
        open(c, "w").write(prog)
        exe = os.path.join(d, "t")
        subprocess.run(["gcc", "-std=c11", "-I", os.path.join(ROOT, "include"), c, "-o", exe], check=True)
        rows = [list(map(int, l.split())) for l in subprocess.run([exe], capture_output=True, text=True, check=True).stdout.splitlines()]
    structs = [_lib.ScanFwdArgs, _lib.ScanBwdArgs, _lib.Ss2dFwdArgs, _lib.Ss2dBwdArgs]
    assert rows[0] == [ctypes.sizeof(s) for s in structs]
    assert rows[1] == [s.batch.offset for s in structs]
    assert rows[2] == [s.dtype.offset for s in structs]


def test_scalars_and_error_strings(L):
    assert L.xfs_version() == 1
    assert L.xfs_chunk_len() == 256
    assert [L.xfs_num_chunks(n) for n in (1, 256, 257, 3136, 16384)] == [1, 1, 2, 13, 64]
    assert L.xfs_error_string(0) == b"ok"
    for code in range(-6, 0):
        assert L.xfs_error_string(code).startswith(b"xfscan:")
    assert L.xfs_launch_count() >= 0


def test_ss2d_checkpoint_row_length(L):
    """xfs_ss2d_states_len: floats per (batch, 4*D) row of the checkpoints the fused forward writes.  One state per chunk
    and n in general; one per LANE and chunk (32 per chunk) on the backbone path (f32, N == 1, L % 4 == 0, L > 256,
    working set within shared memory), whose backward replays 8 positions per lane from them."""
    F32, BF16 = 0, 1
    f = L.xfs_ss2d_states_len
    assert f(1, 56, 56, F32, F32) == 13 * 32            # config 2: lane checkpoints
    assert f(1, 28, 28, F32, F32) == 4 * 32
    assert f(1, 56, 56, BF16, F32) == 13 * 32           # 16-bit rows, fp32 output, L % 8 == 0: lane checkpoints too
    assert f(1, 56, 56, BF16, BF16) == 13               # 16-bit output: chunk checkpoints (generic kernels)
    assert f(1, 18, 18, BF16, F32) == 2                 # L % 8 != 0
    assert f(16, 56, 56, F32, F32) == 13 * 16           # N > 1
    assert f(1, 14, 14, F32, F32) == 1                  # one chunk
    assert f(1, 7, 7, F32, F32) == 1
    assert f(1, 57, 57, F32, F32) == 13                 # L % 4 != 0: generic kernels
    assert f(1, 128, 128, F32, F32) == 64               # beyond the fused working set: chunk checkpoints (stand-alone scan)
    assert f(0, 1, 1, F32, F32) == 0


def test_argument_errors_are_negative_codes(L):
    from xfmamba_b200 import _lib
    vp = ctypes.c_void_p
    one = vp(16)
    assert L.xfs_cross_scan(None, one, 1, 1, 1, 1, 0, 0, 0, None) == -1
    assert L.xfs_cross_scan(one, one, 0, 1, 1, 1, 0, 0, 0, None) == -2
    assert L.xfs_cross_scan(one, one, 1, 1, 1, 1, 7, 0, 0, None) == -3
    assert L.xfs_cross_merge(one, one, 1, 1, 1, 1, 0, 5, 0, None) == -5
    assert L.xfs_swap_scan(one, None, one, 1, 1, 1, 0, None) == -1
    a = _lib.ScanFwdArgs(one, one, one, one, one, None, None, one, None, 2, 6, 4, 8, 4, 0, 0, 1, 0)   # dim % ngroups != 0
    assert L.xfs_selective_scan_fwd(a, None) == -2
    a = _lib.ScanFwdArgs(one, one, one, one, one, None, None, one, None, 2, 6, 300, 8, 1, 0, 0, 1, 0)  # dstate > 256
    assert L.xfs_selective_scan_fwd(a, None) == -2
    a = _lib.ScanFwdArgs(one, one, one, one, one, None, None, one, None, 2, 6, 4, 8, 1, 1, 2, 1, 0)    # bf16 in, f16 out
    assert L.xfs_selective_scan_fwd(a, None) == -3
    f = _lib.Ss2dFwdArgs(one, one, one, one, one, None, None, one, None, 1, 4, 1, 400, 400, 0, 0, 1, 0)  # L too large for smem
    assert L.xfs_ss2d_fwd(f, None) == -5
    assert L.xfs_ss2d_supported(192, 1, 56, 56, 0, 0) == 1
    assert L.xfs_ss2d_supported(192, 1, 400, 400, 0, 0) == 0
    assert L.xfs_ss2d_supported(192, 100, 7, 7, 0, 0) == 0


def test_python_surface_has_no_cpu_fallback():
    import xfmamba_b200 as xf
    x = torch.randn(1, 2, 3, 4)
    with pytest.raises(RuntimeError, match="CUDA"):
        xf.cross_scan_fn(x)
    with pytest.raises(RuntimeError, match="CUDA"):
        xf.cross_merge_fn(torch.randn(1, 4, 2, 3, 4))
    u = torch.randn(1, 4, 8)
    with pytest.raises(RuntimeError, match="CUDA"):
        xf.selective_scan_fn(u, u, torch.randn(4, 2), torch.randn(1, 1, 2, 8), torch.randn(1, 1, 2, 8))
    with pytest.raises(RuntimeError, match="torch"):
        xf.selective_scan_fn(u, u, torch.randn(4, 2), torch.randn(1, 1, 2, 8), torch.randn(1, 1, 2, 8), backend="torch")
    with pytest.raises(RuntimeError, match="CUDA"):
        xf.SwappingScan_multiview.apply(x, x)
    with pytest.raises(RuntimeError, match="CUDA"):
        xf.ss2d_scan(x, torch.randn(1, 8, 12), torch.randn(8, 1), torch.randn(1, 4, 1, 12), torch.randn(1, 4, 1, 12))


def test_operator_signatures_match_reference():
    """positional call shapes used by forward_corev2 (models/fusion_vmamba.py:1064-1065, 1145, 1174)"""
    import inspect
    import xfmamba_b200 as xf
    assert list(inspect.signature(xf.selective_scan_fn).parameters) == [
        "u", "delta", "A", "B", "C", "D", "delta_bias", "delta_softplus", "oflex", "backend"]
    for fn in (xf.cross_scan_fn, xf.cross_merge_fn):
        assert list(inspect.signature(fn).parameters)[1:] == [
            "in_channel_first", "out_channel_first", "one_by_one", "scans", "force_torch"]
    from xfmamba_b200 import csms6s, csm_triton
    assert csms6s.CrossScan is xf.CrossScanF and csms6s.CrossMerge is xf.CrossMergeF
    assert csm_triton.cross_scan_fn is xf.cross_scan_fn


def test_accumulator_replica_heuristics(monkeypatch):
    """host logic only: replicas are never more than channels, 1 for large batches and short rows, overridable"""
    from xfmamba_b200 import csms6s, fusion_ops
    monkeypatch.delenv("XFS_ACC_REPLICAS", raising=False)
    assert fusion_ops.ss2d_acc_replicas(192, 3136, 64) == 1
    assert fusion_ops.ss2d_acc_replicas(192, 3136, 2) == 4
    assert fusion_ops.ss2d_acc_replicas(1024, 196, 2) == 16
    assert fusion_ops.ss2d_acc_replicas(1024, 196, 16) == 8
    assert fusion_ops.ss2d_acc_replicas(2048, 49, 2) == 1
    assert fusion_ops.ss2d_acc_replicas(3, 196, 1) == 3
    monkeypatch.setenv("XFS_ACC_REPLICAS", "5")
    assert fusion_ops.ss2d_acc_replicas(192, 3136, 64) == 5
    assert csms6s._acc_replicas(192, 3136, 64) == 1
    assert csms6s._acc_replicas(192, 3136, 4) == 4
    assert csms6s._acc_replicas(4, 3136, 4) == 1
