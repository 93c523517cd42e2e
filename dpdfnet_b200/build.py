"""In-tree build of the CUDA engine (``libdpdfnet_b200.so``) for sm_100a.

    python -m dpdfnet_b200.build          # compile if sources are newer than the library

nvcc cross-compiles without a GPU; the resulting ``.so`` is git-ignored but travels to the GPU box.
"""
from __future__ import annotations

import concurrent.futures
import os
import shutil
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
LIB = PKG / "lib" / "libdpdfnet_b200.so"
SOURCES = ["api.cu", "k_frontend.cu", "k_conv.cu", "k_dprnn.cu", "k_dprnn_tc.cu", "k_dprnn_intra_tc.cu", "k_conv_tc.cu", "k_conv_tma.cu", "k_gru_tc.cu", "k_dft_tc.cu", "k_dense.cu", "k_resample.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC"]


def _nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found: the engine has no pure-Python / CPU fallback")
    return exe


def needs_build() -> bool:
    if not LIB.exists():
        return True
    t = LIB.stat().st_mtime
    deps = list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.h")) + \
        [PKG.parent / "include" / "dpdfnet_b200.h"]
    return any(p.stat().st_mtime > t for p in deps)


def build_library(force: bool = False, verbose: bool = False) -> Path:
    if not force and not needs_build():
        return LIB
    nvcc = _nvcc()
    obj_dir = PKG.parent / "build"
    obj_dir.mkdir(exist_ok=True)
    LIB.parent.mkdir(exist_ok=True)

    def compile_one(src: str) -> Path:
        obj = obj_dir / (src.replace(".cu", ".o"))
        cmd = [nvcc, *NVCC_FLAGS, "-c", str(CSRC / src), "-o", str(obj)]
        if verbose:
            print(" ".join(cmd), flush=True)
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        return obj

    with concurrent.futures.ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    cmd = [nvcc, "-shared", "-o", str(LIB), *map(str, objs)]       # static cudart: no runtime-version skew with torch
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose=True))
