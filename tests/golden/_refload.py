"""Loads the UNMODIFIED reference Python (read-only at /root/reference or $XFMAMBA_REF) for fixture generation.

Only usable in the build container -- the GPU box has no reference tree, which is why the outputs are
committed as .npz fixtures.  Shims (none of them touch reference files):
  * timm / fvcore / torchinfo are not installed -> tiny stub modules with the few names the reference imports
  * models/csm_triton.py:506,516 wraps the call in ``torch.cuda.device(x.device)`` which raises for CPU
    tensors on torch>=2.x -> ``cross_scan_fn/cross_merge_fn`` are re-pointed at ``CrossScanF/CrossMergeF.apply``
    (the torch implementation named as the oracle) inside the loaded modules.
"""
import importlib.util
import os
import sys
import types
import warnings

import torch

REF_ROOT = os.environ.get("XFMAMBA_REF", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "models", "csms6s.py"))


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def _install_stubs():
    if "timm" not in sys.modules:
        class DropPath(torch.nn.Module):
            def __init__(self, drop_prob=0.0, scale_by_keep=True):
                super().__init__()
                self.drop_prob = drop_prob

            def forward(self, x):
                if self.drop_prob == 0.0 or not self.training:
                    return x
                raise NotImplementedError("stub DropPath only supports eval / p=0")

        def trunc_normal_(t, mean=0.0, std=1.0, a=-2.0, b=2.0):
            return torch.nn.init.trunc_normal_(t, mean=mean, std=std, a=a, b=b)

        timm = _stub("timm")
        models = _stub("timm.models")
        layers = _stub("timm.models.layers", DropPath=DropPath, trunc_normal_=trunc_normal_)
        timm.models = models
        models.layers = layers
    if "fvcore" not in sys.modules:
        fv = _stub("fvcore")
        nnm = _stub("fvcore.nn", FlopCountAnalysis=object, flop_count_str=lambda *a, **k: "",
                    flop_count=lambda *a, **k: ({}, {}), parameter_count=lambda *a, **k: {})
        fv.nn = nnm
    if "torchinfo" not in sys.modules:
        _stub("torchinfo", summary=lambda *a, **k: None)


def _load(modname, relpath):
    path = os.path.join(REF_ROOT, relpath)
    spec = importlib.util.spec_from_file_location(modname, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[modname] = mod
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        spec.loader.exec_module(mod)
    return mod


_cache = {}


def load():
    """returns a namespace with the reference modules: csms6s, csm_triton, fusion_vmamba, net_fusionmamba"""
    if _cache:
        return types.SimpleNamespace(**_cache)
    assert available(), f"reference tree not found at {REF_ROOT}"
    _install_stubs()
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        csms6s = _load("ref_csms6s", "models/csms6s.py")
        csm = _load("ref_csm_triton", "models/csm_triton.py")
        # fusion_vmamba does relative imports (.csm_triton, .csms6s, .mamba2...) -> needs a package context
        pkg = types.ModuleType("refmodels")
        pkg.__path__ = [os.path.join(REF_ROOT, "models")]
        sys.modules["refmodels"] = pkg
        sys.modules["refmodels.csms6s"] = csms6s
        sys.modules["refmodels.csm_triton"] = csm
        try:
            fusion = _load("refmodels.fusion_vmamba", "models/fusion_vmamba.py")
        except Exception as e:  # pragma: no cover - diagnostic
            raise RuntimeError(f"cannot import reference fusion_vmamba: {e!r}")

    def cs(x, in_channel_first=True, out_channel_first=True, one_by_one=False, scans=0, force_torch=False):
        return csm.CrossScanF.apply(x, in_channel_first, out_channel_first, one_by_one, scans)

    def cm(y, in_channel_first=True, out_channel_first=True, one_by_one=False, scans=0, force_torch=False):
        return csm.CrossMergeF.apply(y, in_channel_first, out_channel_first, one_by_one, scans)

    fusion.cross_scan_fn = cs
    fusion.cross_merge_fn = cm
    _cache.update(csms6s=csms6s, csm_triton=csm, fusion_vmamba=fusion, cross_scan_fn=cs, cross_merge_fn=cm)
    return types.SimpleNamespace(**_cache)


def load_net():
    ns = load()
    if "net" not in _cache:
        # net_fusionmamba.py does `from models.fusion_vmamba import ...`
        models_pkg = types.ModuleType("models")
        models_pkg.__path__ = [os.path.join(REF_ROOT, "models")]
        models_pkg.build_model = lambda *a, **k: (_ for _ in ()).throw(NotImplementedError("stub"))
        sys.modules["models"] = models_pkg
        sys.modules["models.fusion_vmamba"] = ns.fusion_vmamba
        import contextlib
        import io
        with contextlib.redirect_stdout(io.StringIO()):
            _cache["net"] = _load("ref_net_fusionmamba", "net_fusionmamba.py")
    return _cache["net"]
