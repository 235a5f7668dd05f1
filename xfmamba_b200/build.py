"""Builds ``xfmamba_b200/libxfscan.so`` in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m xfmamba_b200.build [--force] [--verbose]

The library is a plain C-ABI shared object (include/xfscan.h); it does not link against torch.
"""
from __future__ import annotations

import hashlib
import os
import shlex
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
SO = PKG / "libxfscan.so"
OBJ = PKG / "build"
SOURCES = ["routes.cu", "selective_scan.cu", "ss2d_fwd.cu", "ss2d_bwd.cu", "ss2d_ring_fwd.cu", "ss2d_lane_bwd.cu", "fusion_small.cu", "ss2d_small.cu", "ss2d_mid.cu", "layernorm2d.cu", "dwconv.cu", "dtproj.cu", "capi.cu"]
HEADERS = [CSRC / "xfscan_common.cuh", CSRC / "ss2d_tiles.cuh", CSRC / "ss2d_fused.cuh", CSRC / "ss2d_ring.cuh", PKG.parent / "include" / "xfscan.h"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
    "-Xptxas", "-v",
] + shlex.split(os.environ.get("XFS_NVCC_EXTRA", ""))      # e.g. -DXFS_... for timing experiments (tools/)


# --use_fast_math (ftz, approximate division / sqrt / exp) only for the scan translation units, whose transcendental code is
# written with explicit approx PTX anyway and is held to the oracle by the parity tests.  The routes (documented as bit exact,
# including fp32 denormals), LayerNorm2d (rstd) and the depthwise convolution (SiLU / sigmoid backward) use IEEE arithmetic.
FAST_MATH = {"selective_scan.cu", "ss2d_fwd.cu", "ss2d_bwd.cu", "ss2d_ring_fwd.cu", "ss2d_lane_bwd.cu", "ss2d_small.cu", "ss2d_mid.cu",
             "fusion_small.cu"}


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found (needed to build libxfscan.so for sm_100a)")


def _digest() -> str:
    h = hashlib.sha256()
    for p in [CSRC / s for s in SOURCES] + HEADERS:
        h.update(p.read_bytes())
    h.update((" ".join(NVCC_FLAGS) + "|" + " ".join(sorted(FAST_MATH))).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> Path:
    stamp = OBJ / "digest.txt"
    digest = _digest()
    if not force and SO.exists() and stamp.exists() and stamp.read_text() == digest:
        return SO
    OBJ.mkdir(exist_ok=True)
    nvcc = _nvcc()

    def compile_one(src: str):
        obj = OBJ / (src[:-3] + ".o")
        cmd = [nvcc, *NVCC_FLAGS, *(["--use_fast_math"] if src in FAST_MATH else []), "-c", str(CSRC / src), "-o", str(obj)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        (OBJ / (src[:-3] + ".ptxas.log")).write_text(r.stderr)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{r.stderr[-4000:]}")
        if verbose:
            print(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    link = [nvcc, "-shared", "-o", str(SO), *map(str, objs), "-gencode", "arch=compute_100a,code=sm_100a",
            "-Xcompiler", "-fPIC", "-lcudart"]
    r = subprocess.run(link, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stderr[-4000:]}")
    stamp.write_text(digest)
    return SO


if __name__ == "__main__":
    so = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(so)
