"""Alias module: lets ``from xfmamba_b200.csm_triton import cross_scan_fn, cross_merge_fn`` replace the reference's
``from .csm_triton import ...`` (models/fusion_vmamba.py:23-25) unchanged.  No Triton is involved."""
from .csm import *  # noqa: F401,F403
from .csm import CrossMerge, CrossMergeF, CrossScan, CrossScanF, cross_merge_fn, cross_scan_fn  # noqa: F401
