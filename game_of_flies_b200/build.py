"""Build game_of_flies_b200/csrc/libgof_b200.so for sm_100a with nvcc (in-tree).

nvcc cross-compiles without a GPU, so this runs in the CPU-only build container
and the resulting .so travels to the GPU box with the repo snapshot.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(CSRC, "libgof_b200.so")
SOURCES = [os.path.join(CSRC, "gof_b200.cu")]
import glob  # noqa: E402

HEADERS = sorted(glob.glob(os.path.join(CSRC, "*.cuh")))  # every header gof_b200.cu can include
HEADERS.append(os.path.join(os.path.dirname(HERE), "include", "gof_b200.h"))

def nvcc_path() -> str:
    return shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"


def stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(f) > t for f in SOURCES + HEADERS + [os.path.abspath(__file__)])


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not stale():
        return LIB
    flags = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
             "-shared", "-Xcompiler", "-fPIC"]
    if verbose:
        flags += ["-Xptxas", "-v"]
    # write next to the target and rename: a snapshot of the tree (gpurun) never sees half a library
    tmp = LIB + f".tmp{os.getpid()}"
    cmd = [nvcc_path(), *flags, *SOURCES, "-o", tmp]
    try:
        subprocess.run(cmd, check=True)
        os.replace(tmp, LIB)
    finally:
        if os.path.exists(tmp):
            os.remove(tmp)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
