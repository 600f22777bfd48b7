"""The CUDA slab path on ONE GPU: several SlabEngine instances in one process exchange
migration records and halo layers -- through the peer-memory transport (P2PRunner: the same
kernels that store into a neighbour on another GPU, here into another arena of this GPU) and
through LocalComm (device copies in place of NCCL send/recv) -- and must reproduce the
single-engine result bit for bit: same cells, same order inside a cell, same summation order."""
import numpy as np
import pytest

from oracle import oracle as orc
from parity_util import random_state

pytestmark = pytest.mark.gpu


def run_single(st, steps, timestep, wrap_mode):
    from game_of_flies_b200.engine import ParticleEngine
    e = ParticleEngine(st["n"], st["parameter_matrix"], st["limits"], st["r_max"], wrap_mode=wrap_mode)
    e.upload([st["pos"][k].astype(np.float32) for k in range(3)], st["types"],
             vel=[st["vel"][k].astype(np.float32) for k in range(3)],
             acc=[st["acc"][k].astype(np.float32) for k in range(3)])
    e.step(steps, timestep)
    out = e.download(pos=True, vel=True, acc=True)
    e.close()
    return out


def run_slabs(st, world_slabs, steps, timestep, wrap_mode, overlap=True, transport="copies"):
    from game_of_flies_b200.slab import (LocalComm, P2PRunner, SlabEngine, SlabRunner, assign_to_slabs,
                                         split_layers)
    grid = orc.grid_from_limits(st["limits"], st["r_max"])
    layers = split_layers(grid["num_bin_z"], world_slabs)
    owner = assign_to_slabs(st["pos"][2], grid["bin_size_z"], grid["num_bin_z"], layers)
    slabs = {}
    for s, (lo, hi) in enumerate(layers):
        m = owner == s
        e = SlabEngine(st["parameter_matrix"], st["limits"], st["r_max"], lo, hi, capacity=int(m.sum() * 1.5) + 4096,
                       wrap_mode=wrap_mode)
        e.upload(st["pos"][:, m], st["types"][m], np.nonzero(m)[0], vel=st["vel"][:, m], acc=st["acc"][:, m])
        slabs[s] = e
    if transport == "p2p":
        runner = P2PRunner(slabs, world_slabs, group_ready=False)
    else:
        runner = SlabRunner(slabs, world_slabs, LocalComm(world_slabs), overlap=overlap)
    # in two calls: the second continues from device-resident counts
    runner.step(steps // 2, timestep)
    runner.step(steps - steps // 2, timestep)
    out = runner.gather_state(st["n"])
    moved = int((out["owner"] != owner).sum())
    runner.close()
    for e in slabs.values():
        e.close()
    return out, moved


@pytest.mark.parametrize("kinds,counts,limits,speed", [
    ("clusters", [6000] * 5, (120, 120, 240), 4.0),
    ("boids", [8000] * 3, (60, 60, 150), 9.0),
    ("mixed", [3000] * 6, (110, 110, 200), 8.0),
])
@pytest.mark.parametrize("world_slabs", [2, 4])
@pytest.mark.parametrize("timestep", [None, 0.03])
def test_slabs_reproduce_single_engine_bit_for_bit(kinds, counts, limits, speed, world_slabs, timestep):
    st = random_state(counts, limits, seed=29, kinds=kinds, speed=speed)
    steps = 8
    for wrap_mode in (orc.WRAP_REFERENCE, orc.WRAP_MINIMAGE):
        want = run_single(st, steps, timestep, wrap_mode)
        # with and without running the halo traffic under the interior forces
        got, moved = run_slabs(st, world_slabs, steps, timestep, wrap_mode, overlap=(wrap_mode == orc.WRAP_REFERENCE))
        assert (got["owner"] >= 0).all(), "every particle must be owned by exactly one slab"
        assert moved > 0, "the run should exercise migration between slabs"
        for k in ("pos", "vel", "acc"):
            assert np.array_equal(got[k], want[k]), f"{k} differs from the single-engine result"
        # the peer-memory transport: stores into the neighbour's arena, flag waits, device-side sizes
        got, moved = run_slabs(st, world_slabs, steps, timestep, wrap_mode, transport="p2p")
        assert (got["owner"] >= 0).all() and moved > 0
        for k in ("pos", "vel", "acc"):
            assert np.array_equal(got[k], want[k]), f"p2p transport: {k} differs from the single-engine result"


def test_p2p_slab_errors_surface_on_the_host():
    """Device-side failures of the peer-memory transport raise at the next host read: a migration
    buffer that overflows (tiny migrate_capacity, fast particles) must not pass silently."""
    from game_of_flies_b200 import _lib
    from game_of_flies_b200.slab import P2PRunner, SlabEngine, assign_to_slabs, split_layers
    st = random_state([6000] * 3, (100, 100, 200), seed=5, kinds="clusters", speed=15.0)
    grid = orc.grid_from_limits(st["limits"], st["r_max"])
    layers = split_layers(grid["num_bin_z"], 2)
    owner = assign_to_slabs(st["pos"][2], grid["bin_size_z"], grid["num_bin_z"], layers)
    slabs = {}
    for s, (lo, hi) in enumerate(layers):
        m = owner == s
        e = SlabEngine(st["parameter_matrix"], st["limits"], st["r_max"], lo, hi, capacity=int(m.sum() * 1.5) + 4096,
                       migrate_capacity=2)
        e.upload(st["pos"][:, m], st["types"][m], np.nonzero(m)[0], vel=st["vel"][:, m], acc=st["acc"][:, m])
        slabs[s] = e
    runner = P2PRunner(slabs, 2, group_ready=False)
    runner.step(30, 0.05)
    with pytest.raises(_lib.GofError, match="migration buffer overflow"):
        for e in slabs.values():
            e.count()
    runner.close()
    for e in slabs.values():
        e.close()


def test_migration_overflow_raises_on_the_sequenced_transport():
    """The same overflow on the Python-sequenced transport (send/recv, here device copies): the
    sticky flag travels in the gathered words and the size check raises before the halo exchange."""
    from game_of_flies_b200 import _lib
    from game_of_flies_b200.slab import LocalComm, SlabEngine, SlabRunner, assign_to_slabs, split_layers
    st = random_state([6000] * 3, (100, 100, 200), seed=5, kinds="clusters", speed=15.0)
    grid = orc.grid_from_limits(st["limits"], st["r_max"])
    layers = split_layers(grid["num_bin_z"], 2)
    owner = assign_to_slabs(st["pos"][2], grid["bin_size_z"], grid["num_bin_z"], layers)
    slabs = {}
    for s, (lo, hi) in enumerate(layers):
        m = owner == s
        e = SlabEngine(st["parameter_matrix"], st["limits"], st["r_max"], lo, hi, capacity=int(m.sum() * 1.5) + 4096,
                       migrate_capacity=2)
        e.upload(st["pos"][:, m], st["types"][m], np.nonzero(m)[0], vel=st["vel"][:, m], acc=st["acc"][:, m])
        slabs[s] = e
    runner = SlabRunner(slabs, 2, LocalComm(2))
    with pytest.raises(_lib.GofError, match="migration buffer overflow"):
        runner.step(30, 0.05)
    for e in slabs.values():
        e.close()


def test_engine_entry_points_reject_slab_engines():
    """gof_engine_step / accelerate / upload on an engine made by gof_slab_create must fail instead
    of treating the local slab grid as a whole periodic box."""
    import ctypes as C
    import torch
    from game_of_flies_b200 import _lib
    from game_of_flies_b200.slab import SlabEngine
    st = random_state([2000] * 2, (80, 80, 160), seed=3, kinds="clusters")
    e = SlabEngine(st["parameter_matrix"], st["limits"], st["r_max"], 0, 4, capacity=8192)
    lib = _lib.load()
    stream = torch.cuda.current_stream().cuda_stream
    assert lib.gof_engine_step(e._handle, 1, C.c_float(-1.0), 0, stream) == _lib.E_STATE
    assert lib.gof_engine_accelerate(e._handle, 0, stream) == _lib.E_STATE
    # the caller's current device is left alone by every entry point
    assert torch.cuda.current_device() == e.index
    e.close()


@pytest.mark.parametrize("transport", ["p2p", "copies"])
def test_a_slab_that_starts_empty(transport):
    """All particles start in the lower half of the box: the upper slabs begin EMPTY, receive their
    first particles through migration, and boundary layers / halos of zero particles travel."""
    st = random_state([5000] * 3, (100, 100, 260), seed=9, kinds="clusters", speed=12.0)
    st["pos"][2] *= 0.45
    st["pos"] = st["pos"].astype(np.float32).astype(np.float64)
    steps = 12
    want = run_single(st, steps, None, orc.WRAP_MINIMAGE)
    got, moved = run_slabs(st, 4, steps, None, orc.WRAP_MINIMAGE, transport=transport)
    assert (got["owner"] >= 0).all()
    for k in ("pos", "vel", "acc"):
        assert np.array_equal(got[k], want[k]), f"{transport}: {k} differs from the single-engine result"
