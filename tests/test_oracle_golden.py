"""Pin the C oracle against the golden vectors the UNMODIFIED reference
produced under numba's CUDA simulator (tests/golden/make_golden.py).

Integer outputs (bins, counts, starts, the set of indices per bin, boid
neighbour counts) must be identical.  Float64 outputs differ from the
reference only by the order in which the simulator's threads issued their
atomic adds: ~1e-13 relative for the float64 accumulators, ~1e-6 for
anything that went through the reference's float32 boid accumulators.
"""
import os

import numpy as np
import pytest

from conftest import golden_files
from oracle import oracle as orc

FILES = golden_files()
IDS = [os.path.basename(f)[:-4] for f in FILES]


def _load(path):
    with np.load(path) as z:
        return {k: z[k] for k in z.files}


def test_golden_present():
    assert len(FILES) >= 6, "golden vectors missing: run tests/golden/make_golden.py"


@pytest.mark.parametrize("path", FILES, ids=IDS)
def test_grid_and_bin_neighbours(path):
    g = _load(path)
    grid = orc.grid_from_limits(g["limits"], float(g["r_max"]))
    for k in ("num_bin_x", "num_bin_y", "num_bin_z", "num_bins"):
        assert grid[k] == int(g[k]), k
    for k in ("bin_size_x", "bin_size_y", "bin_size_z"):
        assert grid[k] == float(g[k]), k
    nb = np.zeros_like(g["bin_neighbours"])
    orc.set_bin_neighbours(grid["num_bin_x"], grid["num_bin_y"], grid["num_bin_z"], nb)
    assert np.array_equal(nb, g["bin_neighbours"])


@pytest.mark.parametrize("path", FILES, ids=IDS)
def test_core_step_matches_reference(path):
    g = _load(path)
    ts = None if np.isnan(g["timestep"]) else float(g["timestep"])
    has_boids = bool(np.any(g["parameter_matrix"][-1] == 1))
    n = int(g["num_particles"])
    for s in range(int(g["steps"])):
        sim = orc.OracleSimulation(g[f"s{s}_in_pos"], g[f"s{s}_in_vel"], g[f"s{s}_in_acc"],
                                   g["particle_type_index_array"], g["parameter_matrix"],
                                   g["limits"], float(g["r_max"]), timestep=ts)
        used = sim.core_step()
        # --- binning: bit exact -------------------------------------------------
        assert np.array_equal(sim.particle_bins, g[f"s{s}_particle_bins"])
        assert np.array_equal(sim.particle_bin_counts, g[f"s{s}_bin_counts"])
        assert np.array_equal(sim.particle_bin_starts, g[f"s{s}_bin_starts"])
        ref_idx = g[f"s{s}_particle_indices"]
        ref_off = g[f"s{s}_bin_offsets"]
        cells = sim.grid["num_cells"]
        for b in range(cells):
            a, c = sim.particle_bin_starts[b], sim.particle_bin_counts[b]
            # the simulator hands out atomic offsets in thread-scheduling order;
            # the oracle in particle order.  Same set, oracle sorted.
            assert np.array_equal(np.sort(ref_idx[a:a + c]), sim.particle_indices[a:a + c])
        assert np.array_equal(ref_idx[sim.particle_bin_starts[sim.particle_bins] + ref_off],
                              np.arange(n))
        # --- timestep -----------------------------------------------------------
        if ts is None:
            m = float(g[f"s{s}_max_sq_speed"])
            t = np.sqrt(m)
            expect = 0.4 / t if t > 1e-6 else 0.1
            assert used == pytest.approx(expect, rel=1e-15)
        else:
            assert used == ts
        # --- boid accumulators ----------------------------------------------------
        assert np.array_equal(sim.boid_counts, g[f"s{s}_boid_counts"])
        if has_boids:
            for k in range(3):
                np.testing.assert_allclose(sim.boid[k], g[f"s{s}_boid_acc"][k], rtol=2e-5,
                                           atol=1e-3)
                np.testing.assert_allclose(sim.boid[3 + k], g[f"s{s}_boid_vel"][k], rtol=2e-5,
                                           atol=1e-3)
        # --- state after the step ---------------------------------------------------
        rtol, atol = (1e-5, 1e-5) if has_boids else (1e-10, 1e-11)
        got_acc = np.stack([sim.acc_x, sim.acc_y, sim.acc_z])
        got_vel = np.stack([sim.vel_x, sim.vel_y, sim.vel_z])
        got_pos = np.stack([sim.pos_x, sim.pos_y, sim.pos_z])
        np.testing.assert_allclose(got_acc, g[f"s{s}_out_acc"], rtol=rtol, atol=atol)
        np.testing.assert_allclose(got_vel, g[f"s{s}_out_vel"], rtol=rtol, atol=atol)
        np.testing.assert_allclose(got_pos, g[f"s{s}_out_pos"], rtol=rtol, atol=atol)


@pytest.mark.parametrize("path", FILES, ids=IDS)
def test_trajectory_from_initial_state(path):
    """Chain all recorded steps from the first state: the oracle must follow
    the reference trajectory, not just single steps."""
    g = _load(path)
    ts = None if np.isnan(g["timestep"]) else float(g["timestep"])
    has_boids = bool(np.any(g["parameter_matrix"][-1] == 1))
    sim = orc.OracleSimulation(g["s0_in_pos"], g["s0_in_vel"], g["s0_in_acc"],
                               g["particle_type_index_array"], g["parameter_matrix"],
                               g["limits"], float(g["r_max"]), timestep=ts)
    steps = int(g["steps"])
    for s in range(steps):
        sim.core_step()
    got_pos = np.stack([sim.pos_x, sim.pos_y, sim.pos_z])
    tol = 1e-3 if has_boids else 1e-8
    np.testing.assert_allclose(got_pos, g[f"s{steps - 1}_out_pos"], rtol=0, atol=tol)


def test_clusters_profile_matches_host_profile():
    """acceleration_updater.clusters == force_profiles.general_force_function
    inside [0, r_max) (force_profiles.py:135-150)."""
    r_min, r_max, f_min, f_max = 5.5, 11.0, -8.0, 4.0
    for d in np.linspace(0.0, r_max, 200, endpoint=False):
        if d < r_min:
            e = -f_min / r_min * d + f_min
        elif d < (r_min + r_max) / 2:
            e = 2 * f_max / (r_max - r_min) * (d - r_min)
        else:
            e = 2 * f_max / (r_max - r_min) * (r_max - d)
        assert orc.clusters(d, r_min, r_max, f_min, f_max) == e
    # past r_max the device function keeps the falling line (no clamp)
    assert orc.clusters(12.0, r_min, r_max, f_min, f_max) == 2 * f_max / (r_max - r_min) * (r_max - 12.0)
