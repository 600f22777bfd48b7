"""Multi-GPU check of the slab path, both transports (needs >= 2 GPUs; launch with torchrun):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \\
        --master-port 29511 tests/run_slab_nccl.py

Every rank drives one slab; rank 0 also runs the whole box on a single engine and the
gathered multi-GPU state must equal it bit for bit.
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import oracle as orc  # noqa: E402
from parity_util import random_state  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    from game_of_flies_b200.engine import ParticleEngine
    from game_of_flies_b200.slab import SlabEngine, assign_to_slabs, make_runner, nccl_options, split_layers
    dist.init_process_group("nccl", device_id=device, pg_options=nccl_options())

    ok = True
    for transport, kinds, counts, limits, speed, timestep in [
        ("p2p", "clusters", [20000] * 5, (160, 160, 60 * world), 4.0, None),
        ("p2p", "mixed", [6000] * 6, (110, 110, 55 * world), 8.0, 0.03),
        ("p2p", "boids", [20000] * 3, (120, 120, 50 * world), 9.0, None),
        ("nccl", "clusters", [20000] * 5, (160, 160, 60 * world), 4.0, None),
        ("nccl", "mixed", [6000] * 6, (110, 110, 55 * world), 8.0, 0.03),
    ]:
        st = random_state(counts, limits, seed=31, kinds=kinds, speed=speed)
        grid = orc.grid_from_limits(st["limits"], st["r_max"])
        layers = split_layers(grid["num_bin_z"], world)
        owner = assign_to_slabs(st["pos"][2], grid["bin_size_z"], grid["num_bin_z"], layers)
        m = owner == rank
        e = SlabEngine(st["parameter_matrix"], st["limits"], st["r_max"], layers[rank][0], layers[rank][1],
                       capacity=int(m.sum() * 1.5) + 4096, device=device)
        e.upload(st["pos"][:, m], st["types"][m], np.nonzero(m)[0], vel=st["vel"][:, m], acc=st["acc"][:, m])
        runner = make_runner(e, world, device, transport=transport)
        steps = 10
        runner.step(steps // 2, timestep)
        runner.step(steps - steps // 2, timestep)
        got = runner.gather_state(st["n"])
        mine = got["owner"] >= 0
        cnt = torch.tensor([int(mine.sum())], device=device)
        dist.all_reduce(cnt)
        single = ParticleEngine(st["n"], st["parameter_matrix"], st["limits"], st["r_max"], device=device)
        single.upload([st["pos"][k].astype(np.float32) for k in range(3)], st["types"],
                      vel=[st["vel"][k].astype(np.float32) for k in range(3)],
                      acc=[st["acc"][k].astype(np.float32) for k in range(3)])
        single.step(steps, timestep)
        want = single.download(pos=True, vel=True, acc=True)
        same = all(np.array_equal(got[k][:, mine], want[k][:, mine]) for k in ("pos", "vel", "acc"))
        good = torch.tensor([1 if same else 0], device=device)
        dist.all_reduce(good, op=dist.ReduceOp.MIN)
        if rank == 0:
            print(f"{transport} {kinds}: world={world} particles={st['n']} owned_total={int(cnt.item())} "
                  f"bit_identical={bool(good.item())}", flush=True)
        ok = ok and bool(good.item()) and int(cnt.item()) == st["n"]
        runner.close()
        e.close(); single.close()
    dist.barrier()
    dist.destroy_process_group()
    if not ok:
        sys.exit(1)


if __name__ == "__main__":
    main()
