"""Stage the UNMODIFIED reference hot-path modules under baseline/_ref/ (git-ignored).

The reference has no setup.py / pyproject.toml, so there is nothing to pip-install: its
"install" is a copy of the five modules `simulation.py` imports.  baseline/_ref/ is listed
in .gitignore (the sources never enter this repository's history) but NOT in .gpurunignore,
so the copy travels to the GPU box, where /root/reference does not exist, and
baseline/run_reference_gpu.py can time the reference's own numba-CUDA kernels on the same
B200 as the engine.  Run in the build container:

    python baseline/install_ref.py
"""
from __future__ import annotations

import os
import shutil
import sys

SRC = "/root/reference/main"
DST = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "main")
MODULES = ("simulation.py", "integrator.py", "acceleration_updater.py", "state_parameters.py",
           "force_profiles.py")


def install() -> str | None:
    if not os.path.isdir(SRC):
        return None
    os.makedirs(DST, exist_ok=True)
    for m in MODULES:
        shutil.copyfile(os.path.join(SRC, m), os.path.join(DST, m))
    return DST


if __name__ == "__main__":
    out = install()
    print(out if out else f"{SRC} not present: nothing staged", file=sys.stderr if not out else sys.stdout)
