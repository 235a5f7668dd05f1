"""oracle/_ref: the reference's own CUDA selective scan, built for sm_100 by oracle/build_ref.py (bench / GPU-test side only)."""
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]


def test_reference_cuda_core_loads_and_exports_fwd_bwd():
    from oracle import ref_gpu
    if not ref_gpu.available():
        pytest.skip("oracle/_ref/selective_scan_cuda_core.so not built (python oracle/build_ref.py needs /root/reference)")
    m = ref_gpu.load()
    assert callable(m.fwd) and callable(m.bwd)
    # an sm_100 cubin is inside (the reference's own setup.py stops at sm_90)
    out = subprocess.run(["cuobjdump", "--list-elf", str(ref_gpu.REF_DIR / "selective_scan_cuda_core.so")], capture_output=True, text=True)
    if out.returncode == 0:
        assert "sm_100" in out.stdout


def test_nothing_in_the_product_imports_the_oracle_or_the_reference_build():
    """the product path must not route through oracle/ (CPU restatement) or oracle/_ref (the reference's kernels)"""
    for py in (ROOT / "xfmamba_b200").rglob("*.py"):
        src = py.read_text()
        assert "import oracle" not in src and "from oracle" not in src and "oracle/_ref" not in src and "selective_scan_cuda" not in src, py


def test_cross_scan_merge_restatement_matches_oracle_routes():
    """oracle/ref_gpu.py's torch CrossScan / CrossMerge against the pinned oracle routes (CPU tensors)"""
    import numpy as np
    import torch
    import oracle
    from oracle import ref_gpu
    rng = np.random.default_rng(5)
    x = rng.standard_normal((2, 3, 5, 7)).astype(np.float32)
    xs = ref_gpu.cross_scan(torch.from_numpy(x)).numpy()
    assert np.array_equal(xs, oracle.cross_scan(x))
    ys = rng.standard_normal((2, 4, 3, 35)).astype(np.float32)
    y = ref_gpu.cross_merge(torch.from_numpy(ys), 5, 7).numpy()
    assert np.array_equal(y, oracle.cross_merge(ys, 5, 7))          # same add order (csm_triton.py:61-62): bit exact
