"""Drop the sm_100a operators into the reference's modules without editing them.

    import xfmamba_b200.patch as xfpatch
    import models.fusion_vmamba as fv
    xfpatch.install(fv)                 # cross_scan_fn / cross_merge_fn / selective_scan_fn / SwappingScan / SwappingMerge
    xfpatch.install(fv, fused=True)     # additionally the scan cores call the fused kernels
    xfpatch.uninstall(fv)               # restores every attribute install() replaced

`fused=True` replaces the body of the reference's scan cores (models/fusion_vmamba.py:446-578, 777-845, 1035-1188) by calls
into xfmamba_b200.model.ss2d_core / cross_fuse_core / shallow_fuse_core, which read the SAME module parameters
(x_proj_weight, dt_projs_weight, dt_projs_bias, A_logs, Ds, out_norm); everything above the core (in_proj, conv, gating,
out_proj) stays reference code.

Only what the fused kernels compute is rerouted: scan_mode "cross2d", no x_proj_bias, ssoflex (fp32 scan output) --
which is what every XFMamba configuration uses ("v05_noz": `partial(forward_corev2, force_fp32=False, no_einsum=True)`,
models/fusion_vmamba.py:976).  Any other combination (v051d unidi, v052d bidi, v052dc cascade2d, an `x_proj_bias`
attribute, ssoflex=False, a to_dt_softmax variant, ...) keeps the reference's own core, which after install() still runs
on the sm_100a stand-alone operators.  `force_fp32`, `no_einsum`, `selective_scan_backend`, `scan_force_torch`, `nrows` and
`backnrows` do not change the function being computed on this path (the fused kernels always compute in fp32 registers)
and are accepted and ignored.
"""
from __future__ import annotations

import functools
import types

from . import SwappingMerge_multiview, SwappingScan_multiview, cross_merge_fn, cross_scan_fn, selective_scan_fn

_NAMES = dict(cross_scan_fn=cross_scan_fn, cross_merge_fn=cross_merge_fn, selective_scan_fn=selective_scan_fn,
              SwappingScan_multiview=SwappingScan_multiview, SwappingMerge_multiview=SwappingMerge_multiview)

_IGNORED = {"force_fp32", "no_einsum", "selective_scan_backend", "scan_force_torch", "nrows", "backnrows", "cascade2d"}
_SAVED = {}          # module -> {attribute name: original object}; classes are keyed by (module, "Class.attr")


def _reroutable(self, kwargs) -> bool:
    """True iff the call asks for exactly what the fused kernels compute (see the module docstring)."""
    scan_mode = kwargs.get("scan_mode", "cross2d")
    if scan_mode not in ("cross2d", 0):
        return False
    if getattr(self, "x_proj_bias", None) is not None:
        return False
    if not kwargs.get("ssoflex", True) or not kwargs.get("delta_softplus", True):
        return False
    if kwargs.get("to_dt_softmax", False) or kwargs.get("x_proj_weight", None) is not None:
        return False
    return all(k in _IGNORED or k in ("scan_mode", "ssoflex", "delta_softplus", "to_dt_softmax", "x_proj_weight") for k in kwargs)


def _finish(self, y, like):
    """(B, D, L) fp32 -> the reference's post-processing of the merged scan output (out_norm, layout, dtype)"""
    B, D, H, W = like.shape
    y = y.view(B, D, H, W)
    if not self.channel_first:
        y = y.permute(0, 2, 3, 1)
    return self.out_norm(y).to(like.dtype)


def _fused_ss2d_core(self, x, **kwargs):
    from .model import ss2d_core
    if not _reroutable(self, kwargs):
        return type(self).forward_corev2(self, x, **kwargs)
    return _finish(self, ss2d_core(x, self.x_proj_weight, self.dt_projs_weight, self.dt_projs_bias, self.A_logs, self.Ds), x)


def _fused_cross_core(self, x=None, x2=None, x_fuse=None, **kwargs):
    from .model import cross_fuse_core
    if not _reroutable(self, kwargs):
        return self._xfs_reference_corev2(x, x2, x_fuse, **kwargs)
    ys = cross_fuse_core(x, x2, x_fuse, self.x_proj_weight, self.dt_projs_weight, self.dt_projs_bias, self.A_logs, self.Ds)
    return tuple(_finish(self, y, src) for y, src in zip(ys, (x, x2, x_fuse)))


def _fused_shallow_core(self, x=None, x2=None, **kwargs):
    from .model import shallow_fuse_core
    if not _reroutable(self, kwargs):
        return self._xfs_reference_corev2(x, x2, **kwargs)
    ys = shallow_fuse_core(x, x2, self.x_proj_weight, self.dt_projs_weight, self.dt_projs_bias, self.A_logs, self.Ds)
    return tuple(_finish(self, y, src) for y, src in zip(ys, (x, x2)))


def _set(mod, owner, name, value):
    key = name if owner is mod else f"{owner.__name__}.{name}"
    saved = _SAVED.setdefault(mod, {})
    if key not in saved:                     # keep the ORIGINAL across repeated install() calls
        saved[key] = (owner, name, owner.__dict__.get(name) if isinstance(owner, type) else getattr(owner, name))
    setattr(owner, name, value)


def install(*modules: types.ModuleType, fused: bool = False) -> None:
    for mod in modules:
        for name, obj in _NAMES.items():
            if hasattr(mod, name):
                _set(mod, mod, name, obj)
        if not fused:
            continue
        if hasattr(mod, "SS2Dv2"):
            cls = mod.SS2Dv2
            orig_init = cls.__init__

            @functools.wraps(orig_init)
            def patched_init(self, *a, __orig=orig_init, **k):
                __orig(self, *a, **k)
                core = getattr(self, "forward_core", None)
                # only the `partial(self.forward_corev2, ...)` cores (models/fusion_vmamba.py:971-986) are candidates; the
                # keywords bound by the partial travel with it, so _fused_ss2d_core can tell cross2d from the other modes
                if isinstance(core, functools.partial) and getattr(core.func, "__name__", "") == "forward_corev2":
                    self.forward_core = functools.partial(types.MethodType(_fused_ss2d_core, self), *core.args, **core.keywords)
            _set(mod, cls, "__init__", patched_init)
        for cname, repl in (("Cross_SS2Dv5", _fused_cross_core), ("ShallowFuse_SS2Dv4", _fused_shallow_core)):
            if hasattr(mod, cname):
                cls = getattr(mod, cname)
                if "_xfs_reference_corev2" not in cls.__dict__:
                    cls._xfs_reference_corev2 = cls.__dict__["forward_corev2"]
                _set(mod, cls, "forward_corev2", repl)


def uninstall(*modules: types.ModuleType) -> None:
    """Restores everything install() replaced in `modules` (instances built while SS2Dv2.__init__ was patched keep their
    rerouted forward_core; build models after uninstall() to get pure reference modules)."""
    for mod in modules:
        for owner, name, orig in _SAVED.pop(mod, {}).values():
            if orig is None and isinstance(owner, type):
                delattr(owner, name)
            else:
                setattr(owner, name, orig)
        for cname in ("Cross_SS2Dv5", "ShallowFuse_SS2Dv4"):
            cls = getattr(mod, cname, None)
            if cls is not None and "_xfs_reference_corev2" in cls.__dict__:
                delattr(cls, "_xfs_reference_corev2")
