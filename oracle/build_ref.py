"""Builds the reference's OWN selective-scan CUDA extension for sm_100 into oracle/_ref/ (bench-side GPU context number only).

    python oracle/build_ref.py          # needs /root/reference (build container); the .so travels to the GPU box

The sources are compiled where they lie under /root/reference/models/selective_scan/csrc/selective_scan (nothing is copied into
this repository); only the built ``selective_scan_cuda_core*.so`` lands in ``oracle/_ref/`` (git-ignored).  It is the mamba/VMamba
selective-scan kernel the reference calls at models/csms6s.py:83,101 when its extension is installed.  ``bench.py`` times it as
``ref_gpu`` next to the fused kernels; nothing in ``xfmamba_b200/`` imports it.  Plain nvcc/g++ commands, not the reference's setup.py.
"""
from __future__ import annotations

import shutil
import subprocess
import sys
import sysconfig
from pathlib import Path

HERE = Path(__file__).resolve().parent
OUT = HERE / "_ref"
SRC = Path("/root/reference/models/selective_scan/csrc/selective_scan")
NAME = "selective_scan_cuda_core"
SOURCES = ["selective_scan.cpp", "selective_scan_core.cu", "selective_scan_core_fwd2.cu", "selective_scan_core_fwd3.cu",
           "selective_scan_core_fwd4.cu"]


def build(verbose: bool = False) -> Path | None:
    so = OUT / f"{NAME}.so"
    if not SRC.exists():
        return so if so.exists() else None
    if so.exists() and all(so.stat().st_mtime >= (SRC / s).stat().st_mtime for s in SOURCES):
        return so
    import torch
    from torch.utils import cpp_extension as ce
    OUT.mkdir(exist_ok=True)
    obj = OUT / "obj"
    obj.mkdir(exist_ok=True)
    inc = [f"-I{p}" for p in ce.include_paths("cuda")] + [f"-I{sysconfig.get_paths()['include']}", f"-I{SRC}"]
    defs = [f"-DTORCH_EXTENSION_NAME={NAME}", "-DTORCH_API_INCLUDE_EXTENSION_H", f"-D_GLIBCXX_USE_CXX11_ABI={int(torch._C._GLIBCXX_USE_CXX11_ABI)}"]
    nvcc_flags = ["-O3", "-std=c++17", "-U__CUDA_NO_HALF_OPERATORS__", "-U__CUDA_NO_HALF_CONVERSIONS__", "-U__CUDA_NO_BFLOAT16_OPERATORS__",
                  "-U__CUDA_NO_BFLOAT16_CONVERSIONS__", "-U__CUDA_NO_BFLOAT162_OPERATORS__", "-U__CUDA_NO_BFLOAT162_CONVERSIONS__",
                  "--expt-relaxed-constexpr", "--expt-extended-lambda", "--use_fast_math", "-lineinfo",
                  "-gencode", "arch=compute_100,code=sm_100", "-Xcompiler", "-fPIC"]      # the reference's own flags, arch -> sm_100
    procs, objs = [], []
    for s in SOURCES:
        o = obj / (s.rsplit(".", 1)[0] + ".o")
        objs.append(o)
        if s.endswith(".cu"):
            cmd = ["nvcc", *nvcc_flags, *defs, *inc, "-c", str(SRC / s), "-o", str(o)]
        else:
            cmd = ["g++", "-O3", "-std=c++17", "-fPIC", *defs, *inc, "-c", str(SRC / s), "-o", str(o)]
        procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for s, p in procs:
        log, _ = p.communicate()
        if verbose or p.returncode:
            print(f"--- {s}\n{log[-3000:]}")
        if p.returncode:
            raise RuntimeError(f"reference build failed on {s}")
    libs = [f"-L{p}" for p in ce.library_paths("cuda")]
    link = ["g++", "-shared", "-o", str(so), *map(str, objs), *libs, "-lc10", "-lc10_cuda", "-ltorch_cpu", "-ltorch_cuda", "-ltorch",
            "-ltorch_python", "-lcudart"]
    r = subprocess.run(link, capture_output=True, text=True)
    if r.returncode:
        raise RuntimeError("reference link failed:\n" + r.stderr[-3000:])
    shutil.rmtree(obj, ignore_errors=True)           # only the .so travels to the GPU box
    return so


if __name__ == "__main__":
    print(build(verbose="--verbose" in sys.argv))
