"""Builds scgaussian_b200/libscgr.so (the C-ABI CUDA library, include/scgr.h) in-tree with nvcc
for sm_100a.  Called by __graft_entry__.build(); safe to call repeatedly (rebuilds only when a
source is newer than the library)."""
from __future__ import annotations

import os
import shutil
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LIB = os.path.join(_HERE, "libscgr.so")
SOURCES = ["capi.cu", "preprocess.cu", "binning.cu", "render.cu", "loss.cu", "collective.cu", "knn.cu", "model.cu", "prior.cu"]
HEADERS = [os.path.join(CSRC, "common.cuh"), os.path.join(_HERE, "..", "include", "scgr.h")]

NVCC_FLAGS = [
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "-Xcompiler", "-fPIC", "-shared",
]


def _nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: cannot build libscgr.so")


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES] + HEADERS
    return any(os.path.getmtime(d) > t for d in deps)


def build_library(force: bool = False, verbose: bool = False) -> str:
    if force or needs_build():
        cmd = [_nvcc(), *NVCC_FLAGS, "-o", LIB, *[os.path.join(CSRC, s) for s in SOURCES]]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
            print(" ".join(cmd))
        env = dict(os.environ)
        env.pop("CC", None)   # the image's $CC wrapper lacks the OpenMP spec; nvcc wants plain gcc
        env.pop("CXX", None)
        subprocess.check_call(cmd, env=env)
    return LIB


if __name__ == "__main__":
    print(build_library(force=True, verbose=True))
