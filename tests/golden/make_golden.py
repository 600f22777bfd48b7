"""Generate the golden vectors under tests/golden/ from the UNMODIFIED reference.

Run in the build container only (it needs /root/reference, which does not
exist on the GPU box):

    python tests/golden/make_golden.py [case ...]

The reference's hot path is numba-CUDA; here it is executed by numba's CUDA
simulator (NUMBA_ENABLE_CUDASIM=1), importing main/simulation.py,
main/integrator.py and main/acceleration_updater.py as they are.  Two things
are controlled from outside, neither touching reference source:

* matplotlib (imported by main/force_profiles.py for plotting only) is not in
  this image, so an empty stub module is registered before the import;
* every case is launched as ONE CUDA block (threads >= N).  The reference's
  kernels zero their accumulators and then use a block-level syncthreads as
  if it were a grid-wide barrier (acceleration_updater.py:253-266,
  integrator.py:20-30), so a multi-block launch is racy on a GPU and plainly
  wrong in the simulator, which runs blocks one after the other.  With one
  block the kernels compute what they were written to compute.  For
  N <= 512 this IS Simulation's own configuration (threads = 512).

Each .npz holds the initial state as built by Simulation.__init__ plus, for
every recorded step, the state before and after Simulation.core_step().
"""
from __future__ import annotations

import os
import sys
import types

os.environ["NUMBA_ENABLE_CUDASIM"] = "1"

import numpy as np  # noqa: E402

REF_MAIN = "/root/reference/main"
OUT_DIR = os.path.dirname(os.path.abspath(__file__))

# name: (clus_types, boid_types, seed, limits, steps, timestep)
CASES = {
    # BASELINE.json configs[0]: Clusters 2D, 3 species, N=2,000, periodic box
    "clusters2d_3sp_n2000": ([700, 700, 600], [0], 4, (100, 100, 0), 1, None),
    "clusters2d_3sp_n300": ([100, 100, 100], [0], 434, (100, 100, 0), 6, None),
    "clusters2d_fixed_dt": ([120, 90], [0], 35, (60, 80, 0), 4, 0.02),
    "clusters3d_5sp_n400": ([80, 80, 80, 80, 80], [0], 10, (40, 40, 40), 4, None),
    "clusters3d_aniso": ([150, 150, 100], [0], 954, (40, 55, 70), 3, None),
    "boids3d_3sp_n450": ([0], [150, 150, 150], 100, (30, 30, 30), 6, None),
    "boids2d_2sp_n300": ([0], [150, 150], 50, (40, 40, 0), 6, None),
    "mixed3d_4sp_n480": ([120, 120], [120, 120], 69, (45, 45, 45), 5, None),
    "mixed3d_7sp_n490": ([70] * 4, [70] * 3, 434, (50, 50, 50), 3, None),
}


def _import_reference():
    if "matplotlib" not in sys.modules:
        mpl = types.ModuleType("matplotlib")
        plt = types.ModuleType("matplotlib.pyplot")
        mpl.pyplot = plt
        sys.modules["matplotlib"] = mpl
        sys.modules["matplotlib.pyplot"] = plt
    sys.path.insert(0, REF_MAIN)
    import simulation as ref_simulation  # noqa: E402
    return ref_simulation


def _host(a):
    return np.array(a.copy_to_host())


def _snapshot(sim):
    return dict(
        pos=np.stack([_host(sim.d_pos_x), _host(sim.d_pos_y), _host(sim.d_pos_z)]),
        vel=np.stack([_host(sim.d_vel_x), _host(sim.d_vel_y), _host(sim.d_vel_z)]),
        acc=np.stack([_host(sim.d_acc_x), _host(sim.d_acc_y), _host(sim.d_acc_z)]),
    )


def run_case(name, ref_simulation):
    clus, boid, seed, limits, steps, timestep = CASES[name]
    np.random.seed(12345)  # Simulation draws np.random.random() for mixed kinds after seeding
    sim = ref_simulation.Simulation(np.array(clus), np.array(boid), seed=seed, limits=limits)
    sim.timestep = timestep
    n = int(sim.num_particles)
    threads = 512
    while threads < n:
        threads *= 2
    sim.threads = threads
    sim.blocks = 1
    out = dict(
        clus_types=np.array(clus), boid_types=np.array(boid), seed=seed,
        limits=np.array(limits, dtype=np.float64), steps=steps,
        timestep=np.nan if timestep is None else timestep,
        num_particles=n, num_types=int(sim.num_types),
        parameter_matrix=np.array(sim.parameter_matrix), r_max=float(sim.r_max),
        particle_type_index_array=np.array(sim.particle_type_index_array),
        num_bin_x=sim.num_bin_x, num_bin_y=sim.num_bin_y, num_bin_z=sim.num_bin_z,
        bin_size_x=sim.bin_size_x, bin_size_y=sim.bin_size_y, bin_size_z=sim.bin_size_z,
        num_bins=int(sim.num_bins), bin_neighbours=np.array(sim.bin_neighbours),
        threads=threads, blocks=1,
    )
    for s in range(steps):
        before = _snapshot(sim)
        sim.core_step()
        after = _snapshot(sim)
        for k, v in before.items():
            out[f"s{s}_in_{k}"] = v
        for k, v in after.items():
            out[f"s{s}_out_{k}"] = v
        out[f"s{s}_particle_bins"] = _host(sim.d_particle_bins)
        out[f"s{s}_bin_counts"] = _host(sim.d_particle_bin_counts)
        out[f"s{s}_bin_starts"] = _host(sim.d_particle_bin_starts)
        out[f"s{s}_bin_offsets"] = _host(sim.d_bin_offsets)
        out[f"s{s}_particle_indices"] = _host(sim.d_particle_indices)
        out[f"s{s}_boid_counts"] = _host(sim.d_boid_counts)
        out[f"s{s}_boid_acc"] = np.stack([_host(sim.d_boid_acc_x), _host(sim.d_boid_acc_y),
                                          _host(sim.d_boid_acc_z)])
        out[f"s{s}_boid_vel"] = np.stack([_host(sim.d_boid_vel_x), _host(sim.d_boid_vel_y),
                                          _host(sim.d_boid_vel_z)])
        if timestep is None:
            # integrate() reads the reduced maximum from the last element (integrator.py:183)
            out[f"s{s}_max_sq_speed"] = float(_host(sim.d_sq_speed)[-1])
        print(f"  {name}: step {s} done", flush=True)
    path = os.path.join(OUT_DIR, name + ".npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path} ({os.path.getsize(path) / 1024:.0f} KiB)")


if __name__ == "__main__":
    ref = _import_reference()
    names = sys.argv[1:] or list(CASES)
    for nm in names:
        run_case(nm, ref)
