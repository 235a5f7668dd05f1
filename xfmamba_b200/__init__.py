"""xfmamba_b200 -- B200 (sm_100a) kernels for XFMamba's SS2D scan path behind the reference's operator surface.

    from xfmamba_b200 import selective_scan_fn, cross_scan_fn, cross_merge_fn          # drop-ins (models/csms6s.py, csm_triton.py)
    from xfmamba_b200 import SwappingScan_multiview, SwappingMerge_multiview           # drop-ins (models/fusion_vmamba.py:189-241)
    from xfmamba_b200 import ss2d_scan                                                 # fused scan+S6+merge (new)

All operators are CUDA-only; there is no CPU or PyTorch fallback (see oracle/ for the test-side restatement).
"""
from .csm import CrossMerge, CrossMergeF, CrossScan, CrossScanF, cross_merge_fn, cross_scan_fn
from .csms6s import SelectiveScanCuda, selective_scan_fn
from .fusion_ops import (SS2DScanFn, SwappingMerge_multiview, SwappingScan_multiview, ss2d_fused_supported, ss2d_scan,
                         swapping_merge, swapping_scan)

__all__ = [
    "selective_scan_fn", "SelectiveScanCuda", "cross_scan_fn", "cross_merge_fn", "CrossScanF", "CrossMergeF",
    "CrossScan", "CrossMerge", "SwappingScan_multiview", "SwappingMerge_multiview", "swapping_scan", "swapping_merge",
    "ss2d_scan", "SS2DScanFn", "ss2d_fused_supported",
]
__version__ = "0.1.0"
