"""xfmamba_b200.patch against the real reference modules (only where the reference tree is mounted, i.e. the build
container; skipped on the GPU box).  CPU-only: checks that install() re-points the five operator names and that the fused
core replacement reads the reference module's own parameters -- by running it with oracle-backed ops and comparing with the
reference's forward_corev2."""
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
import _refload  # noqa: E402
from conftest import rel_err

pytestmark = pytest.mark.skipif(not _refload.available(), reason="reference tree not mounted")


def test_install_repoints_operator_names():
    import xfmamba_b200 as xf
    import xfmamba_b200.patch as xfpatch
    ref = _refload.load()
    fv = ref.fusion_vmamba
    saved = {n: getattr(fv, n) for n in xfpatch._NAMES}
    try:
        xfpatch.install(fv)
        assert fv.cross_scan_fn is xf.cross_scan_fn and fv.selective_scan_fn is xf.selective_scan_fn
        assert fv.SwappingScan_multiview is xf.SwappingScan_multiview
        m = fv.SS2Dv2(d_model=8, d_state=1, forward_type="v05_noz", channel_first=True)
        with pytest.raises(RuntimeError, match="CUDA"):      # reference module now reaches the CUDA-only operators
            m.forward_core(torch.randn(1, 16, 4, 4))
    finally:
        for n, v in saved.items():
            setattr(fv, n, v)


def test_fused_core_replacement_matches_reference_core(monkeypatch):
    import xfmamba_b200.model as M
    import xfmamba_b200.patch as xfpatch
    from test_model_host import OracleOps
    ref = _refload.load()
    fv = ref.fusion_vmamba
    torch.manual_seed(0)
    m = fv.SS2Dv2(d_model=8, d_state=1, ssm_ratio=2.0, forward_type="v05_noz", channel_first=True)
    x = torch.randn(2, m.d_inner, 6, 5)
    with torch.no_grad():
        want = m.forward_core(x)
        for name in ("ss2d_scan", "cross_scan_fn", "layer_norm_2d", "dt_proj"):
            monkeypatch.setattr(M.OPS, name, getattr(OracleOps, name))
        got = xfpatch._fused_ss2d_core(m, x)
    assert rel_err(got.numpy(), want.numpy()) < 1e-4


def test_install_fused_only_reroutes_cross2d_and_uninstall_restores():
    import functools
    import xfmamba_b200.patch as xfpatch
    ref = _refload.load()
    fv = ref.fusion_vmamba
    before = {n: getattr(fv, n) for n in xfpatch._NAMES}
    init0, cross0, shallow0 = fv.SS2Dv2.__init__, fv.Cross_SS2Dv5.forward_corev2, fv.ShallowFuse_SS2Dv4.forward_corev2
    try:
        xfpatch.install(fv, fused=True)
        xfpatch.install(fv, fused=True)                       # idempotent: the ORIGINALS stay saved
        m = fv.SS2Dv2(d_model=8, d_state=1, forward_type="v05_noz", channel_first=True)
        assert isinstance(m.forward_core, functools.partial) and m.forward_core.func.__func__ is xfpatch._fused_ss2d_core
        assert m.forward_core.keywords == dict(force_fp32=False, no_einsum=True)
        bidi = fv.SS2Dv2(d_model=8, d_state=1, forward_type="v052d_noz", channel_first=True)
        assert bidi.forward_core.keywords["scan_mode"] == "bidi"
        assert not xfpatch._reroutable(bidi, bidi.forward_core.keywords)      # stays on the reference core
        assert fv.Cross_SS2Dv5.forward_corev2 is xfpatch._fused_cross_core
        assert fv.Cross_SS2Dv5._xfs_reference_corev2 is cross0
    finally:
        xfpatch.uninstall(fv)
    assert fv.SS2Dv2.__init__ is init0 and fv.Cross_SS2Dv5.forward_corev2 is cross0
    assert fv.ShallowFuse_SS2Dv4.forward_corev2 is shallow0
    assert all(getattr(fv, n) is v for n, v in before.items())
    assert not hasattr(fv.Cross_SS2Dv5, "_xfs_reference_corev2")
