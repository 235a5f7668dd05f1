"""Drop the sm_100a operators into the reference's modules without editing them.

    import xfmamba_b200.patch as xfpatch
    import models.fusion_vmamba as fv
    xfpatch.install(fv)                 # cross_scan_fn / cross_merge_fn / selective_scan_fn / SwappingScan / SwappingMerge
    xfpatch.install(fv, fused=True)     # additionally SS2Dv2.forward_core and Cross_SS2Dv5.forward_corev2 use ss2d_scan

`fused=True` replaces the body of the reference's scan cores (models/fusion_vmamba.py:446-578, 1035-1188) by calls into
xfmamba_b200.model.ss2d_core / cross_fuse_core, which read the SAME module parameters (x_proj_weight, dt_projs_weight,
dt_projs_bias, A_logs, Ds, out_norm); everything above the core (in_proj, conv, gating, out_proj) stays reference code.
Only the cross2d scan mode of the XFMamba configuration ("v05_noz", channel_first) is rerouted.
"""
from __future__ import annotations

import types

from . import SwappingMerge_multiview, SwappingScan_multiview, cross_merge_fn, cross_scan_fn, selective_scan_fn

_NAMES = dict(cross_scan_fn=cross_scan_fn, cross_merge_fn=cross_merge_fn, selective_scan_fn=selective_scan_fn,
              SwappingScan_multiview=SwappingScan_multiview, SwappingMerge_multiview=SwappingMerge_multiview)


def _fused_ss2d_core(self, x, **kwargs):
    from .model import ss2d_core
    B, D, H, W = x.shape
    y = ss2d_core(x, self.x_proj_weight, self.dt_projs_weight, self.dt_projs_bias, self.A_logs, self.Ds).view(B, D, H, W)
    if not self.channel_first:
        y = y.permute(0, 2, 3, 1)
    return self.out_norm(y).to(x.dtype)


def _fused_cross_core(self, x=None, x2=None, x_fuse=None, **kwargs):
    from .model import cross_fuse_core
    B, D, H, W = x.shape
    ys = cross_fuse_core(x, x2, x_fuse, self.x_proj_weight, self.dt_projs_weight, self.dt_projs_bias, self.A_logs, self.Ds)
    outs = []
    for y, src in zip(ys, (x, x2, x_fuse)):
        y = y.view(B, D, H, W)
        if not self.channel_first:
            y = y.permute(0, 2, 3, 1)
        outs.append(self.out_norm(y).to(src.dtype))
    return tuple(outs)


def install(*modules: types.ModuleType, fused: bool = False) -> None:
    for mod in modules:
        for name, obj in _NAMES.items():
            if hasattr(mod, name):
                setattr(mod, name, obj)
        if fused:
            if hasattr(mod, "SS2Dv2"):
                cls = mod.SS2Dv2
                orig_init = cls.__init__

                def patched_init(self, *a, __orig=orig_init, **k):
                    __orig(self, *a, **k)
                    self.forward_core = types.MethodType(_fused_ss2d_core, self)
                cls.__init__ = patched_init
            if hasattr(mod, "Cross_SS2Dv5"):
                mod.Cross_SS2Dv5.forward_corev2 = _fused_cross_core
