"""Host-side slab decomposition logic on the CPU: two gloo ranks, one slab each, the CUDA
engine replaced by tests/fake_slab.py; the result must follow the single-domain oracle."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from conftest import ROOT
from game_of_flies_b200.slab import DistComm, LocalComm, SlabRunner, assign_to_slabs, split_layers
from oracle import oracle as orc
from parity_util import random_state

sys.path.insert(0, os.path.join(ROOT, "tests"))
from fake_slab import FakeSlabEngine  # noqa: E402

STEPS = 6


def make_state():
    st = random_state([90, 90, 80], (56.0, 58.0, 110.0), seed=17, kinds="clusters", speed=6.0)
    return st


def oracle_result(st, timestep):
    o = orc.OracleSimulation(st["pos"], st["vel"], st["acc"], st["types"], st["parameter_matrix"], st["limits"],
                             st["r_max"], timestep=timestep, wrap_mode=orc.WRAP_MINIMAGE)
    for _ in range(STEPS):
        o.core_step()
    return np.stack([o.pos_x, o.pos_y, o.pos_z]), np.stack([o.vel_x, o.vel_y, o.vel_z])


def build_slabs(st, world_slabs, local):
    grid = orc.grid_from_limits(st["limits"], st["r_max"])
    layers = split_layers(grid["num_bin_z"], world_slabs)
    owner = assign_to_slabs(st["pos"][2], grid["bin_size_z"], grid["num_bin_z"], layers)
    slabs = {}
    for s in local:
        e = FakeSlabEngine(st["parameter_matrix"], st["limits"], st["r_max"], layers[s][0], layers[s][1])
        m = owner == s
        e.upload(st["pos"][:, m], st["types"][m], np.nonzero(m)[0], vel=st["vel"][:, m], acc=st["acc"][:, m])
        slabs[s] = e
    return slabs


def check(runner, st, timestep, only_local=False):
    want_pos, want_vel = oracle_result(st, timestep)
    got = runner.gather_state(st["n"])
    mine = got["owner"] >= 0
    if not only_local:
        assert mine.all()
    np.testing.assert_allclose(got["pos"][:, mine], want_pos[:, mine], rtol=0, atol=2e-5)
    np.testing.assert_allclose(got["vel"][:, mine], want_vel[:, mine], rtol=0, atol=2e-4)
    return int(mine.sum())


def test_split_layers():
    assert split_layers(8, 2) == [(0, 4), (4, 8)]
    assert split_layers(9, 4) == [(0, 2), (2, 4), (4, 6), (6, 9)]
    with pytest.raises(ValueError):
        split_layers(5, 4)
    with pytest.raises(ValueError):
        split_layers(8, 1)


@pytest.mark.parametrize("world_slabs", [2, 3, 4])
@pytest.mark.parametrize("timestep", [None, 0.05])
def test_slabs_in_one_process_follow_the_oracle(world_slabs, timestep):
    """LocalComm: the runner's bookkeeping (who sends what to whom, where halos land)."""
    st = make_state()
    runner = SlabRunner(build_slabs(st, world_slabs, range(world_slabs)), world_slabs, LocalComm(world_slabs))
    runner.step(STEPS, timestep)
    check(runner, st, timestep)
    # particles did cross slab boundaries during the run (migration was exercised)
    grid = orc.grid_from_limits(st["limits"], st["r_max"])
    layers = split_layers(grid["num_bin_z"], world_slabs)
    before = assign_to_slabs(st["pos"][2], grid["bin_size_z"], grid["num_bin_z"], layers)
    after = runner.gather_state(st["n"])["owner"]
    assert (before != after).sum() > 0


def _worker(rank, world, port, timestep, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        st = make_state()
        runner = SlabRunner(build_slabs(st, world, [rank]), world, DistComm(world))
        runner.step(STEPS, timestep)
        n_mine = check(runner, st, timestep, only_local=True)
        total = torch.tensor([n_mine])
        dist.all_reduce(total)
        assert int(total.item()) == st["n"], "every particle must be owned by exactly one rank"
        ret[rank] = n_mine
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("timestep", [None, 0.05])
def test_two_gloo_ranks_follow_the_oracle(timestep):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ret = mp.Manager().dict()
    mp.spawn(_worker, args=(2, port, timestep, ret), nprocs=2, join=True)
    assert len(ret) == 2 and sum(ret.values()) == make_state()["n"]


def test_word_checks_raise_on_every_slabs_failure():
    """After the size all-gather every rank inspects EVERY slab's words: an error flag, a live
    count over capacity or a boundary layer the RECEIVER cannot hold raises the same error
    everywhere (a rank that only checked its own slab would leave its peers in send/recv)."""
    from game_of_flies_b200 import _lib
    st = make_state()
    slabs = build_slabs(st, 3, range(3))
    for s, e in slabs.items():
        e.capacity, e.halo_capacity, e.migrate_capacity = 1000, 50 + 10 * s, 64
    runner = SlabRunner(slabs, 3, LocalComm(3))
    assert runner.caps == {0: [1000, 50, 64, 0], 1: [1000, 60, 64, 0], 2: [1000, 70, 64, 0]}
    ok = {0: [100, 40, 40, 0], 1: [100, 40, 40, 0], 2: [100, 40, 40, 0]}
    runner._check_words(ok)
    with pytest.raises(_lib.GofError, match="slab 1: migration buffer overflow"):
        runner._check_words({**ok, 1: [100, 40, 40, 2]})
    with pytest.raises(_lib.GofError, match="slab 2: particle capacity exceeded"):
        runner._check_words({**ok, 2: [1001, 40, 40, 0]})
    # slab 1's bottom layer goes to slab 0 (halo_capacity 50): 55 does not fit although slab 1 could hold it
    with pytest.raises(_lib.GofError, match="slab 1: bottom layer of 55 particles does not fit slab 0"):
        runner._check_words({**ok, 1: [100, 55, 40, 0]})
    slabs[2].migrate_capacity = 128
    with pytest.raises(ValueError, match="migrate_capacity must be the same"):
        SlabRunner(slabs, 3, LocalComm(3))


def _worker_error(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch.distributed as dist
    from game_of_flies_b200 import _lib
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        st = make_state()
        slabs = build_slabs(st, world, [rank])
        for e in slabs.values():
            e.capacity, e.migrate_capacity = 10_000, 64
            e.halo_capacity = 10_000 if rank == 0 else 3   # rank 1 cannot hold rank 0's boundary layers
        runner = SlabRunner(slabs, world, DistComm(world))
        try:
            runner.step(1, 0.05)
            ret[rank] = "no error"
        except _lib.GofError as exc:
            ret[rank] = str(exc)
    finally:
        dist.destroy_process_group()


def test_two_gloo_ranks_raise_the_same_error():
    """The failure belongs to rank 0's layers and rank 1's capacity: both ranks must raise, and
    neither may be left waiting in an exchange."""
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ret = mp.Manager().dict()
    mp.spawn(_worker_error, args=(2, port, ret), nprocs=2, join=True)
    assert len(ret) == 2 and ret[0] == ret[1] and "does not fit slab 1's halo_capacity 3" in ret[0], dict(ret)
