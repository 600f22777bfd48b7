"""Build the C oracle (oracle/gof_oracle.c) into oracle/libgof_oracle.so.

TEST INFRASTRUCTURE.  The shared object is git-ignored (*.so) but travels to
the GPU box with the gpurun snapshot.  No reference sources are compiled here:
the reference is Python (numba) and cannot be built into oracle/_ref.
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "gof_oracle.c")
LIB = os.path.join(HERE, "libgof_oracle.so")

CFLAGS = ["-O2", "-fPIC", "-shared", "-fopenmp", "-ffp-contract=off", "-fno-fast-math",
          "-Wall", "-Wextra", "-std=c11"]


def build(force: bool = False) -> str:
    if (not force and os.path.exists(LIB)
            and os.path.getmtime(LIB) >= os.path.getmtime(SRC)):
        return LIB
    cmd = ["gcc", *CFLAGS, SRC, "-o", LIB, "-lm"]
    subprocess.run(cmd, check=True)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
