"""Builds libjagua_cuda.so (CUDA kernels + C ABI + host CDE mirror) in-tree with nvcc for sm_100a.

Bit-exactness flags (see csrc/predicates.cuh): no FMA contraction on either side, IEEE division and
square root, no flush-to-zero.
"""
from __future__ import annotations

import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
# JG_LIB_NAME (+ JG_EXTRA_NVCC): a second build of the library under another name, for same-box A/B runs of kernel variants
LIB_NAME = os.environ.get("JG_LIB_NAME", "libjagua_cuda.so")
LIB = os.path.join(HERE, LIB_NAME)
SOURCES = ["kernels.cu", "detect.cu", "search.cu", "engine.cu", "pool.cu", "cde_host.cpp", "cde_image.cpp", "item_host.cpp"]
HEADERS = ["kernels.cuh", "detect_grid.cuh", "collect_grid.cuh", "launch.hpp", "predicates.cuh", "layout_soa.h", "cde_host.hpp",
           os.path.join("..", "..", "include", "jagua_cuda.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    *os.environ.get("JG_EXTRA_NVCC", "").split(), "-fmad=false", "-prec-div=true", "-prec-sqrt=true", "-ftz=false",
    "-Xcompiler", "-fPIC,-ffp-contract=off,-fno-fast-math,-Wall,-Wno-unknown-pragmas",
]


def nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: the CUDA engine cannot be built")


def is_fresh() -> bool:
    if not os.path.exists(LIB):
        return False
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return all(os.path.getmtime(d) <= t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and is_fresh():
        return LIB
    objdir = os.path.join(HERE, "build" if LIB_NAME == "libjagua_cuda.so" else "build_" + os.path.splitext(LIB_NAME)[0])
    os.makedirs(objdir, exist_ok=True)
    objs = []
    for src in SOURCES:
        obj = os.path.join(objdir, os.path.splitext(src)[0] + ".o")
        cmd = [nvcc(), *NVCC_FLAGS, "-c", os.path.join(CSRC, src), "-o", obj]
        if verbose and src.endswith(".cu"):
            cmd[1:1] = ["-Xptxas", "-v"]
        subprocess.run(cmd, check=True)
        objs.append(obj)
    subprocess.run([nvcc(), "-shared", "-o", LIB, *objs, "-lpthread"], check=True)
    return LIB


if __name__ == "__main__":
    import sys
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
