"""Builds the CUDA library in-tree (akuaengine_b200/libakua_pbf.so) for sm_100a with nvcc.

The library is the product: hand-written CUDA kernels + the C++ host solver + the C ABI of include/akua_pbf.h.
nvcc cross-compiles without a GPU, so this also runs in the CPU-only dev container.
"""
from __future__ import annotations

import os
import shutil
import subprocess
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
REPO_DIR = PKG_DIR.parent
CSRC = PKG_DIR / "csrc"
LIB_PATH = PKG_DIR / "libakua_pbf.so"
SOURCES = [CSRC / "pbf_solver.cu"]
HEADERS = [CSRC / "pbf_kernels.cuh", CSRC / "list_build.cuh", CSRC / "pbf_params.h", CSRC / "radix_sort.cuh", CSRC / "slab_kernels.cuh", CSRC / "wall_model.cuh", CSRC / "pbf_slab.inl",
           REPO_DIR / "include" / "akua_pbf.h"]

NVCC_FLAGS = [
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "-Xcompiler", "-fPIC", "-shared",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found; cannot build libakua_pbf.so")


def is_stale() -> bool:
    if not LIB_PATH.exists():
        return True
    t = LIB_PATH.stat().st_mtime
    return any(p.stat().st_mtime > t for p in SOURCES + HEADERS)


def build_library(force: bool = False, verbose: bool = False) -> Path:
    if not force and not is_stale():
        return LIB_PATH
    cmd = [_nvcc(), *NVCC_FLAGS, "-o", str(LIB_PATH), *map(str, SOURCES), "-ldl"]
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    return LIB_PATH


if __name__ == "__main__":
    import sys
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
