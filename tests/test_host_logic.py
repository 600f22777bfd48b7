"""Host-side logic against the reference's own outputs (golden vectors), no GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import ROOT, golden_files
from game_of_flies_b200 import _lib
from game_of_flies_b200.simulation import host_state
from game_of_flies_b200 import force_profiles
from oracle import oracle as orc

FILES = golden_files()
IDS = [os.path.basename(f)[:-4] for f in FILES]


@pytest.mark.parametrize("path", FILES, ids=IDS)
def test_host_state_matches_reference_constructor(path):
    """Same seed -> same positions, species kinds, scaled parameters, r_max and bins as
    reference Simulation.__init__ (main/simulation.py:12-139)."""
    g = np.load(path)
    np.random.seed(12345)  # make_golden.py seeds the generator the same way before constructing
    hs = host_state(g["clus_types"], g["boid_types"], seed=int(g["seed"]), limits=tuple(g["limits"]))
    assert np.array_equal(hs["parameter_matrix"], g["parameter_matrix"])
    assert float(hs["r_max"]) == float(g["r_max"])
    assert np.array_equal(hs["particle_type_index_array"], g["particle_type_index_array"])
    init = hs["init"]
    assert np.array_equal(np.stack([init["pos_x"], init["pos_y"], init["pos_z"]]), g["s0_in_pos"])
    assert int(hs["num_particles"]) == int(g["num_particles"])
    for k in ("num_bin_x", "num_bin_y", "num_bin_z", "num_bins"):
        assert int(hs[k]) == int(g[k]), k
    for k in ("bin_size_x", "bin_size_y", "bin_size_z"):
        assert float(hs[k]) == float(g[k]), k


def test_library_exports_every_declared_symbol():
    """The C-ABI library loads on a CPU-only host and exports everything include/gof_b200.h declares."""
    lib = _lib.load()
    header = open(os.path.join(ROOT, "include", "gof_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(gof_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 25
    for name in declared:
        assert hasattr(lib, name), f"{name} declared but not exported"
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    assert lib.gof_abi_version() == _lib.GOF_ABI_VERSION


@pytest.mark.parametrize("path", FILES, ids=IDS)
def test_grid_and_half_stencil_host_functions(path):
    g = np.load(path)
    grid = _lib.make_grid(g["limits"], float(g["r_max"]))
    assert (grid.num_bin_x, grid.num_bin_y, grid.num_bin_z) == (int(g["num_bin_x"]), int(g["num_bin_y"]), int(g["num_bin_z"]))
    assert (grid.bin_size_x, grid.bin_size_y, grid.bin_size_z) == (float(g["bin_size_x"]), float(g["bin_size_y"]), float(g["bin_size_z"]))
    nb = np.zeros_like(g["bin_neighbours"])
    from game_of_flies_b200.acceleration_updater import set_bin_neighbours
    set_bin_neighbours(grid.num_bin_x, grid.num_bin_y, grid.num_bin_z, nb)
    assert np.array_equal(nb, g["bin_neighbours"])


def test_grid_rejects_fewer_than_three_bins():
    with pytest.raises(_lib.GofError) as e:
        _lib.make_grid((30.0, 30.0, 0.0), 12.0)
    assert e.value.code == _lib.E_GRID


def test_force_profiles_match_oracle_device_function():
    args = (5.5, 11.0, -8.0, 4.0)
    d = np.linspace(0.0, 13.0, 300)
    want = np.array([orc.clusters(x, *args) for x in d])
    np.testing.assert_allclose(force_profiles.clusters_device(d, args), want, rtol=0, atol=0)
    # the host profile of force_profiles.py clamps to zero outside [0, r_max)
    f = force_profiles.wrap_clusters_force_distance(*args)
    assert f(12.0) == 0 and f(-1.0) == 0
    assert f(3.0) == orc.clusters(3.0, *args)
    m = force_profiles.all_force_functions("cluster_distance_input", *[[args, args], [args, args]])
    assert len(m) == 2 and len(m[0]) == 2 and m[1][0](6.0) == force_profiles.general_force_function(0, [6.0], args)


def test_compute_entry_points_fail_loudly_without_gpu():
    """No CPU fallback: on a host without CUDA the product path raises."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from game_of_flies_b200.simulation import Simulation
    with pytest.raises(RuntimeError):
        Simulation(np.array([50, 50]), np.array([0]), seed=1, limits=(100, 100, 0))


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` runs the oracle port on the host cores (no GPU involved) and
    prints one JSON line with the keys the driver reads."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0", "--reference-particles", "4000"],
                         capture_output=True, text=True, timeout=300, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "particle-updates/sec"
    assert line["value"] > 0 and line["higher_is_better"] is True
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert line["e2e"]["value"] == line["value"]
