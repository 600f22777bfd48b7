"""Time the UNMODIFIED reference (numba-CUDA) on the GPU of this box, at its own sizes,
next to this engine on the same systems.

    python baseline/run_reference_gpu.py [--steps 500] [--out gpurun_out/reference_numba.json]

The reference modules are the copy staged by baseline/install_ref.py (baseline/_ref/main).
Two systems, both straight from the reference's hive.py:
  * benchmark mode, `hive.py -be -nt 1000 1000`: Simulation([1000, 1000], [0], limits=(100, 100, 0),
    seed=434), `update()` once, then `bench_run(500)` (hive.py:88-96, simulation.py:361-379)
  * the default run mode: Simulation([2000]*4, [2000]*3, limits=(100, 100, 100), seed=434)
    (hive.py:97-103), same calls
The reference times with time.perf_counter() around the loop without a device synchronise;
here a `cuda.synchronize()` closes the timed region so that queued kernels are counted.
matplotlib (imported by force_profiles.py for plotting only) is stubbed when absent.

The engine's numbers are taken on the same Simulation arguments through
game_of_flies_b200.simulation.Simulation.bench_run's loop (core_step x steps, synchronised).
Nothing here is part of the product path or of the tests.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import time
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_MAIN = os.path.join(ROOT, "baseline", "_ref", "main")

SYSTEMS = {
    "hive_benchmark_2d_n2000": dict(clus=[1000, 1000], boid=[0], limits=(100, 100, 0)),
    "hive_default_3d_mixed_n14000": dict(clus=[2000] * 4, boid=[2000] * 3, limits=(100, 100, 100)),
}
# --large: BASELINE.json configs[1] (Clusters 3D, 5 species, N = 1M at the density of the 3D run above)
# through the reference's own constructor and kernels.  NOTE: on B200 the reference's kernels FAULT at this
# size (CUDA error 700, illegal address, in the first core_step: see profiles/r02_reference_numba.json), which
# also poisons the CUDA context of the process; kept for the record, do not run it with anything else.
LARGE = {
    "baseline_config1_clusters3d_n1m": dict(clus=[200_000] * 5, boid=[0], limits=(414.913,) * 3, steps=10),
}


def clocks():
    try:
        out = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,clocks.max.sm,clocks_event_reasons.active",
                              "--format=csv,noheader", "-i", "0"], capture_output=True, text=True, timeout=20).stdout
        return out.strip()
    except Exception as e:  # pragma: no cover
        return f"unavailable: {e}"


def time_reference(steps: int) -> dict:
    import numpy as np
    if not os.path.isdir(REF_MAIN):
        return {"unavailable": "baseline/_ref/main missing: run baseline/install_ref.py in the build container"}
    if "matplotlib" not in sys.modules:
        try:
            import matplotlib  # noqa: F401
        except Exception:
            mpl = types.ModuleType("matplotlib")
            plt = types.ModuleType("matplotlib.pyplot")
            mpl.pyplot = plt
            sys.modules["matplotlib"] = mpl
            sys.modules["matplotlib.pyplot"] = plt
    sys.path.insert(0, REF_MAIN)
    out = {}
    try:
        import numba
        from numba import cuda
        import simulation as ref_simulation
        out["numba"] = numba.__version__
        try:
            name = cuda.get_current_device().name
            out["device"] = name.decode() if isinstance(name, bytes) else str(name)
        except AttributeError:  # numba's CUDA simulator
            out["device"] = "cudasim"
    except Exception as e:
        return {"unavailable": f"numba-CUDA import failed: {type(e).__name__}: {e}"}
    for name, cfg in SYSTEMS.items():
        try:
            np.random.seed(12345)
            sim = ref_simulation.Simulation(np.array(cfg["clus"]), np.array(cfg["boid"]), limits=cfg["limits"], seed=434)
            t0 = time.perf_counter()
            sim.update()  # hive.py's dummy call: compiles every kernel
            cuda.synchronize()
            compile_s = time.perf_counter() - t0
            steps = cfg.get("steps", steps)
            for _ in range(min(10, steps)):
                sim.core_step()
            cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(steps):  # bench_run's loop (simulation.py:369-377, record=None)
                sim.core_step()
            cuda.synchronize()
            dt = time.perf_counter() - t0
            n = int(sim.num_particles)
            out[name] = {"n": n, "steps": steps, "seconds": dt, "steps_per_s": steps / dt,
                         "particle_updates_per_s": n * steps / dt, "first_update_s": compile_s}
        except Exception as e:
            out[name] = {"error": f"{type(e).__name__}: {e}"}
    sys.path.remove(REF_MAIN)
    for m in ("simulation", "integrator", "acceleration_updater", "state_parameters", "force_profiles"):
        sys.modules.pop(m, None)
    return out


def time_engine(steps: int) -> dict:
    import numpy as np
    import torch
    sys.path.insert(0, ROOT)
    from game_of_flies_b200.simulation import Simulation
    out = {}
    for name, cfg in SYSTEMS.items():
        np.random.seed(12345)
        sim = Simulation(np.array(cfg["clus"]), np.array(cfg["boid"]), limits=cfg["limits"], seed=434, device="cuda:0")
        sim.update()
        steps = max(cfg.get("steps", steps), 50) if "steps" in cfg else steps
        for _ in range(10):
            sim.core_step()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(steps):
            sim.core_step()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        n = int(sim.num_particles)
        out[name] = {"n": n, "steps": steps, "seconds": dt, "steps_per_s": steps / dt,
                     "particle_updates_per_s": n * steps / dt}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=500)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "reference_numba.json"))
    ap.add_argument("--large", action="store_true", help="also BASELINE.json configs[1] at N = 1M")
    args = ap.parse_args()
    if args.large:
        SYSTEMS.update(LARGE)
    res = {"clocks_before": clocks()}
    res["reference_numba_cuda"] = time_reference(args.steps)
    res["clocks_between"] = clocks()
    res["engine"] = time_engine(args.steps)
    res["clocks_after"] = clocks()
    ratio = {}
    for name in SYSTEMS:
        r, e = res["reference_numba_cuda"].get(name, {}), res["engine"].get(name, {})
        if "steps_per_s" in r and "steps_per_s" in e:
            ratio[name] = e["steps_per_s"] / r["steps_per_s"]
    res["engine_over_reference"] = ratio
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w") as f:
        json.dump(res, f, indent=1)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
