"""In-tree build of libsas_b200.so (hand-written CUDA for sm_100a behind the C ABI of include/sas.h)."""
from __future__ import annotations

import os
import shutil
import subprocess
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
LIB = PKG / "libsas_b200.so"
SOURCES = ["sas_api.cu", "sas_project.cu", "sas_bin.cu", "sas_work.cu", "sas_sort.cu", "sas_binsort.cu", "sas_composite.cu", "sas_pose.cu"]
# -fmad=false: no implicit FMA contraction anywhere (numerics contract, csrc/sas_common.cuh);
# fused multiply-adds exist only where fmaf() is written.
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-fmad=false",
              "-Xcompiler", "-fPIC,-ffp-contract=off", "-shared", "-Xptxas", "-v"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found; libsas_b200.so cannot be built (there is no CPU fallback)")


def needs_build() -> bool:
    if not LIB.exists():
        return True
    t = LIB.stat().st_mtime
    deps = [CSRC / s for s in SOURCES] + [CSRC / "sas_common.cuh", PKG.parent / "include" / "sas.h"]
    return any(d.stat().st_mtime > t for d in deps)


def build_library(force: bool = False, verbose: bool = False) -> Path:
    if not force and not needs_build():
        return LIB
    cmd = [_nvcc(), *NVCC_FLAGS, "-ccbin", "/usr/bin/g++", "-o", str(LIB), *[str(CSRC / s) for s in SOURCES]]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    (PKG / "build_ptxas.log").write_text(res.stdout + res.stderr)
    if verbose:
        print(res.stdout + res.stderr)
    return LIB


if __name__ == "__main__":
    build_library(force=True, verbose=True)
