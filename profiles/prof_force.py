"""A short run of one workload for ncu (plain kernel launches, no CUDA graph) that also prints the
pair statistics of the last step:

    python profiles/prof_force.py [workload] [particles] [steps] [evolve]
    ncu --set full --clock-control none --import-source on -k regex:k_force -s 3 -c 2 -o gpurun_out/prof \
        python profiles/prof_force.py clusters3d 1000000 6
"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from game_of_flies_b200.engine import ParticleEngine  # noqa: E402

workload = sys.argv[1] if len(sys.argv) > 1 else "clusters3d"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 6
evolve = int(sys.argv[4]) if len(sys.argv) > 4 else 0
hs = bench.build_workload(workload, n, 434, None)
init = hs["init"]
eng = ParticleEngine(n, hs["parameter_matrix"], hs["limits"], float(hs["r_max"]))
eng.upload([init[k] for k in ("pos_x", "pos_y", "pos_z")], hs["particle_type_index_array"],
           vel=[init[k] for k in ("vel_x", "vel_y", "vel_z")], acc=[init[k] for k in ("acc_x", "acc_y", "acc_z")])
if evolve:
    eng.step(evolve)
eng.set_graph(False)
eng.timing(True)
torch.cuda.synchronize()
t0 = time.perf_counter()
eng.step(steps)
torch.cuda.synchronize()
ms, k = eng.force_time()
st = eng.pair_stats()
print(f"{workload} N={n}: {steps} steps {1e3 * (time.perf_counter() - t0) / steps:.3f} ms/step, force kernel "
      f"{ms / max(k, 1):.3f} ms; per particle: candidates {st['candidates'] / n:.1f}, evaluated "
      f"{st['evaluated'] / n:.1f}, in range {st['in_range'] / n:.1f}")
eng.close()
