"""In-tree nvcc build of the C-ABI library (sm_100a only).

The shared object lands in ``femflow_b200/_lib/`` so that it travels to the GPU
box with the repository snapshot (it is git-ignored, not gpurun-ignored).
"""
from __future__ import annotations

import os
import shutil
import subprocess

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
LIBDIR = os.path.join(PKG, "_lib")
LIBNAME = "libfemflow_mpm.so"
LIBPATH = os.path.join(LIBDIR, LIBNAME)

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xptxas=-v",
    "-Xcompiler", "-fPIC", "-shared",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; the femflow_b200 CUDA library cannot be built")


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _stale() -> bool:
    if not os.path.exists(LIBPATH):
        return True
    t = os.path.getmtime(LIBPATH)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    deps.append(os.path.join(PKG, "..", "include", "femflow_mpm.h"))
    return any(os.path.getmtime(d) > t for d in deps)


def build_library(force: bool = False, verbose: bool = False) -> str:
    """Compile csrc/*.cu into femflow_b200/_lib/libfemflow_mpm.so."""
    if not force and not _stale():
        return LIBPATH
    os.makedirs(LIBDIR, exist_ok=True)
    cmd = [_nvcc(), *NVCC_FLAGS, "-o", LIBPATH, *sources()]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or proc.returncode != 0:
        print(" ".join(cmd))
        print(proc.stdout)
        print(proc.stderr)
    if proc.returncode != 0:
        raise RuntimeError("nvcc failed building libfemflow_mpm.so")
    with open(os.path.join(LIBDIR, "ptxas.log"), "w") as f:
        f.write(proc.stderr)
    return LIBPATH


if __name__ == "__main__":
    print(build_library(force=True, verbose=True))
