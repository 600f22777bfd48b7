#!/usr/bin/env python
"""bench.py -- particle-updates/sec of the Game_of_Flies hot path on B200.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --steps K --warmup W    # the reference algorithm on host cores

One "step" is one Simulation.core_step (bin -> neighbour forces -> integrate)
over the synthetic workload named in `config.workload`.  At N=1 that is
BASELINE.json configs[1]: Clusters 3D, 5 species, N = 1,000,000, periodic box,
at the number density of the reference's own 3D run (hive.py: 14,000 particles
in 100^3 = 0.014 / unit^3), uniform random start, adaptive timestep.  At N>1 the
box of the one-GPU workload is stacked N times along z and cut into N z-slabs
(weak scaling, one slab per GPU).

Both arms (`--impl b200`, `--impl reference`) build the SAME initial state with
`global_state()`: the reference arm runs the oracle port (oracle/gof_oracle.c,
the reference's algorithm restated in C: half stencil + atomics, float64,
OpenMP over particles like numba-parallel) on ALL host cores this process may
use (`os.sched_getaffinity`, whatever OMP_NUM_THREADS says: torchrun sets it to
1) at the same N as the GPU arm, unless that would not fit the time budget.

Prints ONE JSON line (rank 0).  See the module constants for the algorithmic
byte count behind `roofline`.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

REFERENCE_DENSITY_3D = 14000 / 100.0 ** 3   # reference main/hive.py:71-75
REFERENCE_DENSITY_2D = 2000 / 100.0 ** 2    # BASELINE.json configs[0]

# Algorithmic HBM bytes per particle of the force+integrate kernel (DESIGN.md "k_force"):
# read posw 16 + veli 16 + cell id 4; write posw 16 + veli 16 + acc 16.  Neighbour
# positions are re-read from L1/L2, not HBM, and the cell tables add < 0.3 B/particle.
FORCE_BYTES_PER_PARTICLE = 84

WORKLOADS = {
    # name: (kind, species, dims)
    "clusters3d": ("clusters", 5, 3),
    "boids3d": ("boids", 3, 3),
    "mixed3d": ("mixed", 8, 3),
    "clusters2d": ("clusters", 3, 2),
}


def host_threads() -> int:
    """Cores this process may run on.  Not omp_get_max_threads(): torch.distributed.run exports
    OMP_NUM_THREADS=1 to its workers, which would time the host baseline on one core."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:  # pragma: no cover
        return max(1, os.cpu_count() or 1)


def build_workload(name: str, n_particles: int, seed: int, density: float | None):
    """Initial state through the package's own host constructor logic (same code path as
    Simulation.__init__), on a cubic / square box sized for the reference density."""
    from game_of_flies_b200.simulation import host_state
    kind, species, dims = WORKLOADS[name]
    per = n_particles // species
    counts = np.array([per] * species)
    counts[0] += n_particles - per * species
    if dims == 3:
        rho = density or REFERENCE_DENSITY_3D
        L = float((n_particles / rho) ** (1.0 / 3.0))
        limits = (L, L, L)
    else:
        rho = density or REFERENCE_DENSITY_2D
        L = float((n_particles / rho) ** 0.5)
        limits = (L, L, 0)
    np.random.seed(seed)
    if kind == "clusters":
        hs = host_state(counts, np.array([0]), seed=seed, limits=limits)
    elif kind == "boids":
        hs = host_state(np.array([0]), counts, seed=seed, limits=limits)
    else:
        h = species // 2
        hs = host_state(counts[:h], counts[h:], seed=seed, limits=limits)
    hs["density"] = rho
    hs["box"] = limits
    return hs


def slab_particles(args, side: float, bin_size_z: float, lo: int, hi: int, rank: int, n: int, species: int):
    """Rank `rank`'s particles of the multi-GPU workload: n particles uniform in its slab of
    z-layers [lo, hi), types round-robin blocks, global ids rank * n + i."""
    rng = np.random.default_rng(args.seed + rank)
    pos = rng.random((3, n))
    pos[0] *= side
    pos[1] *= side
    pos[2] = (lo + pos[2] * (hi - lo)) * bin_size_z
    pos = pos.astype(np.float32)
    types = np.repeat(np.arange(species, dtype=np.int32), -(-n // species))[:n]
    ids = np.arange(n, dtype=np.int32) + rank * n
    return pos, types, ids


def global_state(args, world: int, n: int):
    """The whole job's species, box and (lazily) particles, identical for both arms.

    world == 1: the workload as Simulation.__init__ would build it.  world > 1: the one-GPU box
    stacked `world` times along z, cut into `world` slabs of whole bin layers."""
    from game_of_flies_b200 import _lib
    from game_of_flies_b200.slab import split_layers
    if world == 1:
        hs = build_workload(args.workload, n, args.seed, args.density)
        hs["layers"] = None
        return hs
    kind, species, dims = WORKLOADS[args.workload]
    if dims != 3:
        raise SystemExit("the slab decomposition needs a 3D workload")
    hs = build_workload(args.workload, 1000, args.seed, args.density)  # species parameters only
    rho = hs["density"]
    side = float((n / rho) ** (1.0 / 3.0))
    limits = np.array([side, side, side * world])
    hs["limits"], hs["box"] = limits, tuple(limits)
    grid = _lib.make_grid(limits, float(hs["r_max"]))
    hs["num_bin_x"], hs["num_bin_y"], hs["num_bin_z"] = grid.num_bin_x, grid.num_bin_y, grid.num_bin_z
    hs["grid"] = grid
    hs["side"] = side
    hs["layers"] = split_layers(int(grid.num_bin_z), world)
    hs["species"] = species
    return hs


class ClockSampler:
    """nvidia-smi clocks and throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_indices):
        """gpu_indices: the GPUs of this job (rank 0 samples all of them with one nvidia-smi)."""
        self.gpu_indices = list(gpu_indices)
        self.rows = []
        self.proc = None
        self.thread = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "--id=" + ",".join(str(i) for i in self.gpu_indices), f"--query-gpu={self.QUERY}",
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._read, daemon=True)
        self.thread.start()

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for row in self.rows:
            parts = [p.strip() for p in row.split(",")]
            if len(parts) < 8:
                continue
            try:
                sm.append(float(parts[1]))
                mx.append(float(parts[2]))
            except ValueError:
                continue
            for nm, val in zip(names, parts[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_min_mhz": float(min(sm)) if sm else None,
                "sm_max_mhz": float(max(mx)) if mx else None,
                "samples": len(sm), "gpus": len(self.gpu_indices), "reasons": sorted(reasons)}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def recorded_capture(workload: str, n: int):
    """The committed ncu capture of the force kernel that matches this workload, if any."""
    path = os.path.join(ROOT, "profiles", "force_kernel_traffic.json")
    if not os.path.exists(path):
        return None
    try:
        with open(path) as f:
            rec = json.load(f)
        for r in rec.get("captures", []):
            if r.get("workload") == workload and int(r.get("particles", -1)) == n:
                return r
    except Exception:
        return None
    return None


def oracle_sim_from(hs, n_threads, pos=None, types=None):
    from oracle import oracle as orc
    f = lambda *a: np.stack(a).astype(np.float32).astype(np.float64)
    if pos is None:
        init = hs["init"]
        p = f(init["pos_x"], init["pos_y"], init["pos_z"])
        v = f(init["vel_x"], init["vel_y"], init["vel_z"])
        a = f(init["acc_x"], init["acc_y"], init["acc_z"])
        types = hs["particle_type_index_array"]
    else:
        p = np.ascontiguousarray(pos, dtype=np.float64)
        v = np.zeros_like(p)
        a = np.zeros_like(p)
    return orc.OracleSimulation(p, v, a, types, hs["parameter_matrix"], hs["limits"], float(hs["r_max"]),
                                timestep=None, n_threads=n_threads)


def cpu_baseline(workload: str, seed: int, density, budget_s: float = 20.0):
    """The oracle port on this host's cores, on a bounded sample of the workload: the same species,
    parameters, density and box shape at a particle count that one step finishes in seconds.
    particle-updates/s is intensive at fixed density."""
    threads = host_threads()
    n_sample = 250_000
    hs = build_workload(workload, n_sample, seed, density)
    o = oracle_sim_from(hs, threads)
    o.core_step()  # warm-up (page faults, OpenMP pool)
    t0 = time.perf_counter()
    steps = 0
    while True:
        o.core_step()
        steps += 1
        el = time.perf_counter() - t0
        if el > budget_s or steps >= 400:
            break
    return {"value": n_sample * steps / el, "unit": "particle-updates/s", "cores": threads,
            "kind": "port",
            "sample": f"{steps} steps of {workload} at N={n_sample} (same density {hs['density']:.4g}, "
                      f"species and parameters), oracle/gof_oracle.c with {threads} OpenMP threads, {el:.1f} s"}


def run_reference_arm(args):
    """--impl reference: the reference's algorithm on the host cores (oracle port: the reference
    itself is numba-CUDA Python and cannot be compiled into oracle/_ref; its own GPU kernels are
    timed on the same box by baseline/run_reference_gpu.py, see profiles/).  Same initial state as
    the GPU arm (`global_state`), all host threads; under torchrun rank 0 alone runs it."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = max(int(os.environ.get("WORLD_SIZE", str(args.gpus))), 1)
    threads = host_threads()
    n_total = world * args.particles
    budget_s = args.reference_seconds
    # Same N as the GPU arm unless (warmup + steps) of it cannot fit the budget: a short probe at
    # N = 100k measures this host's rate (particle-updates/s is intensive at fixed density).
    probe = build_workload(args.workload, 100_000, args.seed, args.density)
    o = oracle_sim_from(probe, threads)
    o.core_step()
    t0 = time.perf_counter()
    o.core_step()
    rate = 100_000 / (time.perf_counter() - t0)
    fits = n_total * (args.steps + args.warmup) / rate <= budget_s
    if args.reference_particles:
        n_run, same = int(args.reference_particles), int(args.reference_particles) == n_total
    elif fits:
        n_run, same = n_total, True
    else:
        n_run = int(max(100_000, budget_s * rate / (args.steps + args.warmup)))
        same = False
    if same and world > 1:
        hs = global_state(args, world, args.particles)
        parts = [slab_particles(args, hs["side"], float(hs["grid"].bin_size_z), lo, hi, r, args.particles, hs["species"])
                 for r, (lo, hi) in enumerate(hs["layers"])]
        pos = np.concatenate([p[0] for p in parts], axis=1)
        types = np.concatenate([p[1] for p in parts])
        o = oracle_sim_from(hs, threads, pos=pos, types=types)
    else:
        hs = build_workload(args.workload, n_run, args.seed, args.density)
        o = oracle_sim_from(hs, threads)
    for _ in range(args.warmup):
        o.core_step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        o.core_step()
    el = time.perf_counter() - t0
    value = n_run * args.steps / el
    sample = (f"{args.steps} steps at N={n_run}" + (" (the GPU arm's N and initial state)" if same else
              f" (bounded sample: the GPU arm's N={n_total} would need more than {budget_s:.0f} s; same "
              "density, species and parameters)") + f", {threads} OpenMP threads")
    line = {
        "impl": "reference", "reference_kind": "oracle port of the reference algorithm on host cores (C, OpenMP)",
        "metric": "particle-updates/sec", "value": value, "unit": "particle-updates/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * el / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, hs, n_run, world=world if same else 1,
                                  note=None if same else "bounded sample of the workload on host cores"),
        "same_config": same,
        "cpu_baseline": {"value": value, "unit": "particle-updates/s", "cores": threads, "kind": "port",
                         "sample": sample},
        "e2e": {"value": value, "unit": "particle-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(args, hs, n_total, world=1, note=None):
    kind, species, dims = WORKLOADS[args.workload]
    cfg = {
        "workload": f"{kind.capitalize()} {dims}D, {species} species, N={n_total}, periodic box "
                    f"{tuple(round(float(x), 2) for x in hs['box'])}, density {hs['density']:.4g}, "
                    f"r_max {float(hs['r_max']):.3f}, uniform random start, "
                    + ("adaptive timestep" if args.timestep is None else f"fixed timestep {args.timestep}"),
        "particles_per_gpu": int(n_total // max(world, 1)),
        "bins": [int(hs["num_bin_x"]), int(hs["num_bin_y"]), int(hs["num_bin_z"])],
        "wrap_mode": getattr(args, "wrap", "reference"),
        "cache": ("L2 flushed between timed steps (256 MiB overwrite outside the per-step event pairs)"
                  if getattr(args, "flush_l2", True) else "no L2 flush: per-step working set ~100 MB at N = 1M"),
    }
    if note:
        cfg["note"] = note
    return cfg


def slab_parity_check(args, world, rank, device, wrap_mode, make_runner):
    """Multi-process self-check before any timing: a 100k-particle box of the same species runs 6
    adaptive steps on the slab path (this job's transport, one slab per rank) and on ONE engine on
    rank 0; positions, velocities and accelerations must agree bit for bit, particle by particle."""
    import torch
    import torch.distributed as dist
    from game_of_flies_b200 import _lib
    from game_of_flies_b200.engine import ParticleEngine
    from game_of_flies_b200.slab import SlabEngine, assign_to_slabs, split_layers
    kind, species, dims = WORKLOADS[args.workload]
    hs = build_workload(args.workload, 1000, args.seed, args.density)
    r_max, rho = float(hs["r_max"]), float(hs["density"])
    n_total, steps = 100_000, 6
    side = 8.5 * r_max
    lz = max(n_total / (rho * side * side), (2 * world + 0.5) * r_max)
    limits = np.array([side, side, lz])
    grid = _lib.make_grid(limits, r_max)
    layers = split_layers(int(grid.num_bin_z), world)
    rng = np.random.default_rng(args.seed + 1000)
    pos = (rng.random((3, n_total)) * limits[:, None]).astype(np.float32)
    vel = rng.normal(0.0, 3.0, (3, n_total)).astype(np.float32)
    acc = np.zeros((3, n_total), dtype=np.float32)
    types = (np.arange(n_total) % species).astype(np.int32)
    owner = assign_to_slabs(pos[2], float(grid.bin_size_z), int(grid.num_bin_z), layers)
    m = owner == rank
    lo, hi = layers[rank]
    eng = SlabEngine(hs["parameter_matrix"], limits, r_max, lo, hi, capacity=int(m.sum() * 1.5) + 8192,
                     device=device, wrap_mode=wrap_mode)
    eng.upload(pos[:, m], types[m], np.nonzero(m)[0].astype(np.int32), vel=vel[:, m], acc=acc[:, m])
    runner = make_runner(eng)
    runner.step(steps, None)
    d = eng.download()
    bits = torch.zeros((9, n_total), dtype=torch.int32, device=device)
    own = torch.zeros(n_total, dtype=torch.int32, device=device)
    ids = torch.from_numpy(d["ids"].astype(np.int64)).to(device)
    stack = np.concatenate([d["pos"], d["vel"], d["acc"]]).view(np.int32)
    bits[:, ids] = torch.from_numpy(stack).to(device)
    own[ids] = 1
    dist.all_reduce(bits)
    dist.all_reduce(own)
    runner.close()
    eng.close()
    res = None
    if rank == 0:
        single = ParticleEngine(n_total, hs["parameter_matrix"], limits, r_max, device=device, wrap_mode=wrap_mode)
        single.upload([pos[k] for k in range(3)], types, vel=[vel[k] for k in range(3)], acc=[acc[k] for k in range(3)])
        single.step(steps, None)
        w = single.download(pos=True, vel=True, acc=True)
        single.close()
        want = np.concatenate([w["pos"], w["vel"], w["acc"]]).view(np.int32)
        got = bits.cpu().numpy()
        owned_once = bool((own.cpu().numpy() == 1).all())
        moved = int((assign_to_slabs(w["pos"][2], float(grid.bin_size_z), int(grid.num_bin_z), layers) != owner).sum())
        res = {"bit_identical": bool(owned_once and np.array_equal(got, want)), "n": n_total, "steps": steps,
               "slabs": world, "every_particle_owned_once": owned_once, "migrated": moved,
               "against": "one engine on rank 0, same initial state"}
    flag = torch.tensor([1 if (res is None or res["bit_identical"]) else 0], device=device)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="clusters3d", choices=sorted(WORKLOADS))
    ap.add_argument("--particles", type=int, default=1_000_000, help="particles per GPU")
    ap.add_argument("--reference-particles", type=int, default=0,
                    help="force the reference arm's N (default: the GPU arm's N when it fits --reference-seconds)")
    ap.add_argument("--reference-seconds", type=float, default=150.0)
    ap.add_argument("--density", type=float, default=None)
    ap.add_argument("--seed", type=int, default=434)
    ap.add_argument("--e2e-steps", type=int, default=10)
    ap.add_argument("--e2e-systems", type=int, default=3,
                    help="one GPU: independent systems driven from the host in turn in the end-to-end leg "
                         "(1: a single system, every step waits for its own round trip over PCIe)")
    ap.add_argument("--evolve", type=int, default=300,
                    help="steps to evolve before the steady-state timing (0: skip it)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity-check", action="store_true")
    ap.add_argument("--no-north-star", action="store_true",
                    help="skip the extra timed configurations of the default run (BASELINE.json configs[2..4], "
                         "and at 8 GPUs the north_star configuration, 3D Clusters N=16M)")
    ap.add_argument("--no-flush-l2", dest="flush_l2", action="store_false",
                    help="do not overwrite a 256 MiB buffer between timed steps")
    ap.add_argument("--wrap", default="reference", choices=["reference", "minimage"],
                    help="periodic wrap semantics: the reference's (z-wrap quirk included) or minimum image")
    ap.add_argument("--transport", default="auto", choices=["auto", "p2p", "nccl"],
                    help="multi-GPU halo / migration transport: peer-memory stores or NCCL send/recv")
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    ap.add_argument("--timestep", type=float, default=None,
                    help="fixed timestep (default: the reference's adaptive one, 0.4 / max speed)")
    ap.add_argument("--force-split", type=int, default=-1, choices=(-1, 1, 3, 9),
                    help="warps per home particle in the small-system force kernel (-1: automatic)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    if args.impl == "reference":
        run_reference_arm(args)
        return

    import torch
    import torch.distributed as dist
    from game_of_flies_b200 import _lib
    from game_of_flies_b200.engine import ParticleEngine

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        from game_of_flies_b200.slab import nccl_options
        dist.init_process_group("nccl", device_id=device, pg_options=nccl_options())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(device)

    def max_over_ranks(x: float) -> float:
        t = torch.tensor([x], dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    lib = _lib.load()
    _lib.check(lib.gof_set_force_split(args.force_split))
    wrap_mode = _lib.WRAP_REFERENCE if args.wrap == "reference" else _lib.WRAP_MINIMAGE
    pin = lambda a, dt=np.float32: torch.from_numpy(np.ascontiguousarray(a, dtype=dt)).pin_memory()
    parity = None

    def make_job(n, want_e2e=True):
        """Engine, step(k) and e2e_step() for `n` particles per GPU (pinned host buffers only when
        the end-to-end leg will run)."""
        hs = global_state(args, world, n)
        if world == 1:
            # ---- one GPU: the resident engine on the whole periodic box ---------------------
            init = hs["init"]
            h_pos = [pin(init[k]) for k in ("pos_x", "pos_y", "pos_z")]
            h_vel = [pin(init[k]) for k in ("vel_x", "vel_y", "vel_z")]
            h_acc = [pin(init[k]) for k in ("acc_x", "acc_y", "acc_z")]
            h_types = pin(hs["particle_type_index_array"], np.int32)
            bufs = {"in": (h_pos, h_vel, h_acc),
                    "out": tuple([torch.empty(n, dtype=torch.float32).pin_memory() for _ in range(3)]
                                 for _ in range(3)) if want_e2e else None}
            eng = ParticleEngine(n, hs["parameter_matrix"], hs["limits"], float(hs["r_max"]), device=device,
                                 wrap_mode=wrap_mode)
            eng.upload(h_pos, h_types, vel=h_vel, acc=h_acc)
            step = lambda k: eng.step(k, args.timestep)

            def e2e_step():
                p, v, a = bufs["in"]
                eng.upload(p, h_types, vel=v, acc=a)
                eng.step(1, args.timestep)
                op, ov, oa = bufs["out"]
                eng.download_into(pos=op, vel=ov, acc=oa)
                torch.cuda.current_stream(device).synchronize()
                # the next step of a host-driven loop starts from what came back
                bufs["in"], bufs["out"] = bufs["out"], bufs["in"]

            def e2e_pipelined(k):
                """The same loop through gof_engine_step_host: every step uploads the nine state
                arrays and the types from pinned host memory, steps, and downloads the nine arrays
                into the other buffer set, which is the next step's input.  The copies of
                consecutive steps overlap (both directions of the link busy); one wait at the end."""
                for _ in range(k):
                    (p, v, a), (op, ov, oa) = bufs["in"], bufs["out"]
                    eng.step_host([*p, *v, *a], h_types, [*op, *ov, *oa], args.timestep)
                    bufs["in"], bufs["out"] = bufs["out"], bufs["in"]
                eng.host_wait()
                torch.cuda.current_stream(device).synchronize()
            e2e_step.pipelined = e2e_pipelined

            systems = []  # [engine, stream, [in set, out set]] per interleaved system

            def e2e_interleaved(k, count):
                """`count` independent systems driven from the host in turn (the batch path of the
                reference renders many simulations): each keeps its own host-driven loop -- every
                step uploads that system's state from pinned host memory, steps, downloads it into
                the buffers its next step starts from (gof_engine_step_host) -- on its own stream, so
                that the copies of one system run under the kernels of another.  One host wait per
                system at the end."""
                if not systems:
                    for s in range(count):
                        if s == 0:
                            e, sets = eng, [bufs["in"], bufs["out"]]
                        else:
                            e = ParticleEngine(n, hs["parameter_matrix"], hs["limits"], float(hs["r_max"]),
                                               device=device, wrap_mode=wrap_mode)
                            sets = [tuple([t.clone().pin_memory() for t in grp] for grp in bufs["in"]),
                                    tuple([torch.empty(n, dtype=torch.float32).pin_memory() for _ in range(3)]
                                          for _ in range(3))]
                        systems.append([e, torch.cuda.Stream(device), sets])
                    torch.cuda.synchronize(device)
                for _ in range(k):
                    for sy in systems:
                        e, st, sets = sy
                        (p, v, a), (op, ov, oa) = sets
                        with torch.cuda.stream(st):
                            e.step_host([*p, *v, *a], h_types, [*op, *ov, *oa], args.timestep)
                        sy[2] = [sets[1], sets[0]]
                for e, st, _ in systems:
                    e.host_wait()
                    st.synchronize()

            def close_systems():
                for e, _, _ in systems[1:]:
                    e.close()
                systems.clear()
            e2e_step.interleaved = e2e_interleaved
            e2e_step.close = close_systems
            return hs, eng, None, step, e2e_step
        # ---- several GPUs: z-slab decomposition, one slab per rank, weak scaling ---------------
        from game_of_flies_b200.slab import SlabEngine, make_runner
        lo, hi = hs["layers"][rank]
        pos, types, ids = slab_particles(args, hs["side"], float(hs["grid"].bin_size_z), lo, hi, rank, n,
                                         hs["species"])
        cap = int(1.3 * n) + 8192
        eng = SlabEngine(hs["parameter_matrix"], hs["limits"], float(hs["r_max"]), lo, hi, capacity=cap,
                         device=device, wrap_mode=wrap_mode)
        zeros = np.zeros((3, n), dtype=np.float32)
        eng.upload(pos, types, ids, vel=zeros, acc=zeros)
        runner = make_runner(eng, world, device, transport=args.transport)
        step = lambda k: runner.step(k, args.timestep)
        # end to end: pinned, preallocated host buffers of `capacity` entries, ping-pong like the
        # one-GPU path; the particle count of a slab changes from step to step (migration)
        host, cur = [], [0]
        if want_e2e:
            mk = lambda dt=torch.float32: torch.empty(cap, dtype=dt).pin_memory()
            host = [dict(pos=[mk() for _ in range(3)], vel=[mk() for _ in range(3)], acc=[mk() for _ in range(3)],
                         types=mk(torch.int32), ids=mk(torch.int32), n=0) for _ in range(2)]
            for k in range(3):
                host[0]["pos"][k][:n] = torch.from_numpy(pos[k])
                host[0]["vel"][k][:n] = 0
                host[0]["acc"][k][:n] = 0
            host[0]["types"][:n] = torch.from_numpy(types)
            host[0]["ids"][:n] = torch.from_numpy(ids)
            host[0]["n"] = n

        def e2e_step():
            a, b = host[cur[0]], host[cur[0] ^ 1]
            eng.upload(a["pos"], a["types"], a["ids"], vel=a["vel"], acc=a["acc"], n=a["n"])
            runner.step(1, args.timestep)
            b["n"] = eng.download_into(b["pos"], b["vel"], b["acc"], b["ids"], b["types"])
            cur[0] ^= 1
        return hs, eng, runner, step, e2e_step

    if world > 1 and not args.no_parity_check:
        from game_of_flies_b200.slab import make_runner
        parity = slab_parity_check(args, world, rank, device, wrap_mode,
                                   lambda e: make_runner(e, world, device, transport=args.transport))

    def timed_run(n, steps, warmup, e2e_steps, evolve):
        hs, eng, runner, step, e2e_step = make_job(n, want_e2e=e2e_steps > 0)
        # clocks of every GPU of the job, sampled by rank 0 from the warm-up on (nvidia-smi needs a
        # moment to start; the timed region of a short run is over before its first sample)
        sampler = ClockSampler(range(world)) if rank == 0 else None
        if sampler:
            sampler.start()
        step(warmup)
        barrier()
        if sampler:
            sampler.rows.clear()  # (what follows is the timed region and, if it is short, more of the same load)
        eng.timing(True)
        lib.gof_reset_launch_count()
        # Every timed step starts with a cold L2: a 256 MiB buffer (2x the 126 MB L2) is overwritten
        # between steps, outside the per-step event pairs, because the per-step working set at
        # N = 1M (~100 MB) would otherwise stay L2-resident from one step to the next.
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=device) if args.flush_l2 else None

        def timed_steps(count):
            starts = [torch.cuda.Event(enable_timing=True) for _ in range(count)]
            stops = [torch.cuda.Event(enable_timing=True) for _ in range(count)]
            barrier()
            for k in range(count):
                if flush is not None:
                    flush.fill_(k & 0xFF)
                starts[k].record()
                step(1)
                stops[k].record()
            barrier()
            return max_over_ranks(float(sum(a.elapsed_time(b) for a, b in zip(starts, stops))))

        ms_total = timed_steps(steps)
        launches = int(lib.gof_launch_count())
        force_ms, force_launches = eng.force_time()
        phases = eng.phase_times() if hasattr(eng, "phase_times") else None
        phases_all = None
        if world > 1 and phases and phases.get("steps"):
            # every rank's phases (a step ends when the slowest slab is done: the next timestep needs
            # every slab's maximum speed)
            mine = torch.tensor([phases["prefix"], phases["interior_forces"], phases["halo_and_boundary_forces"],
                                 phases["step"]], dtype=torch.float64, device=device)
            allp = [torch.zeros_like(mine) for _ in range(world)]
            dist.all_gather(allp, mine)
            phases_all = [[round(float(v), 4) for v in t.tolist()] for t in allp]
        eng.timing(False)
        # a timed region of a few tens of milliseconds ends before nvidia-smi's next sample: keep the
        # same load on (untimed) until there are at least three samples per GPU, at most ~1.5 s
        t_load = time.perf_counter()
        while True:
            enough = torch.tensor([1 if (sampler is None or len(sampler.rows) >= 3 * world or
                                         time.perf_counter() - t_load > 1.5) else 0], device=device)
            if world > 1:
                dist.broadcast(enough, 0)
            if int(enough.item()):
                break
            step(10)
            torch.cuda.synchronize(device)
        clocks = sampler.stop() if sampler else None
        stats = eng.pair_stats() if hasattr(eng, "pair_stats") else None
        out = {"hs": hs, "ms_total": ms_total, "launches": launches, "force_ms": force_ms, "phases": phases,
               "phases_all": phases_all,
               "force_launches": force_launches, "clocks": clocks, "stats": stats,
               "value": world * n * steps / (ms_total * 1e-3), "ms_per_step": ms_total / steps}
        # ---- end to end through the public API with host buffers --------------------------
        # every step: host state -> device, one core_step, full state back to the host
        if e2e_steps > 0:
            e2e_step()  # (first call touches the pinned buffers)
            barrier()
            t0 = time.perf_counter()
            for _ in range(e2e_steps):
                e2e_step()
            barrier()
            e2e_s = max_over_ranks(time.perf_counter() - t0)
            out["e2e_value"] = world * n * e2e_steps / e2e_s
            if hasattr(e2e_step, "pipelined"):  # (one GPU) the host-driven loop with overlapped transfers
                e2e_step.pipelined(2)
                barrier()
                t0 = time.perf_counter()
                e2e_step.pipelined(e2e_steps)
                barrier()
                out["e2e_sync_value"] = out["e2e_value"]
                out["e2e_value"] = world * n * e2e_steps / (time.perf_counter() - t0)
                if args.e2e_systems > 1:  # several independent systems, each in its own host-driven loop
                    e2e_step.interleaved(2, args.e2e_systems)
                    barrier()
                    t0 = time.perf_counter()
                    e2e_step.interleaved(e2e_steps, args.e2e_systems)
                    barrier()
                    out["e2e_interleaved_value"] = world * n * e2e_steps * args.e2e_systems / (time.perf_counter() - t0)
                    e2e_step.close()
        # ---- steady state: the same system after it has clumped ---------------------------
        if evolve > 0:
            # (the end-to-end steps above left the state `e2e_steps` further on; the count below is
            # what matters: clusters have formed after ~150 steps)
            step(evolve)
            ms_ss = timed_steps(steps)
            st2 = eng.pair_stats() if hasattr(eng, "pair_stats") else None
            out["steady"] = {"evolve_steps": evolve, "value": world * n * steps / (ms_ss * 1e-3),
                             "ms_per_step": ms_ss / steps, "unit": "particle-updates/s",
                             "what": f"the same run timed again after {evolve} more steps (particles clumped)"}
            if st2:
                out["steady"].update({"pairs_in_range_per_particle": st2["in_range"] / max(st2["n"], 1),
                                      "pairs_evaluated_per_particle": st2["evaluated"] / max(st2["n"], 1)})
        if runner is not None and rank == 0 and getattr(runner, "phase_ms", None):
            pm = runner.phase_ms
            print(f"[slab timing] per step: prefix (hash, migration, sort) {pm['prefix'] / max(pm['steps'], 1):.4f} ms, "
                  f"forces + halo {pm['forces'] / max(pm['steps'], 1):.4f} ms", file=sys.stderr, flush=True)
        out["transport"] = getattr(runner, "transport", None)
        if runner is not None:
            runner.close()
        eng.close()
        return out

    n = args.particles
    res = timed_run(n, args.steps, args.warmup, args.e2e_steps, args.evolve if world == 1 else 0)
    north = None
    others = {}
    default_run = args.workload == "clusters3d" and n == 1_000_000 and not args.no_north_star

    def extra(workload, n_per_gpu, steps, what):
        """Another BASELINE.json configuration as a short second timed run of the same job."""
        keep = args.workload
        args.workload = workload
        try:
            r = timed_run(n_per_gpu, steps, 3, 0, 0)
        finally:
            args.workload = keep
        rec = {"workload": what, "value": r["value"], "unit": "particle-updates/s", "ms_per_step": r["ms_per_step"],
               "force_ms_per_step": r["force_ms"] / steps, "steps": steps, "warmup": 3, "n_gpus": world,
               "particles_per_gpu": n_per_gpu, "clocks": r["clocks"]}
        if r["stats"]:
            rec["pairs_in_range_per_particle"] = r["stats"]["in_range"] / max(r["stats"]["n"], 1)
        return rec

    if default_run and world == 8:
        # BASELINE.json's target configuration, as a second timed run of the same job
        north = extra("clusters3d", 2_000_000, args.steps, "Clusters 3D, 5 species, N=16000000 on 8 GPUs (2M per GPU)")
        north["target"] = 1.0e10
        others["configs3_mixed3d_16m"] = extra("mixed3d", 2_000_000, 20,
                                               "BASELINE configs[3]: mixed Clusters+Boids 3D, 8 species, N=16M over 8 GPUs")
    if default_run and world == 1:
        others["configs2_boids3d_4m"] = extra("boids3d", 4_000_000, 30, "BASELINE configs[2]: Boids 3D, 3 species, N=4M, one GPU")
        others["mixed3d_2m"] = extra("mixed3d", 2_000_000, 20,
                                     "mixed Clusters+Boids 3D, 8 species, N=2M, one GPU (configs[3]'s per-GPU share)")
    if default_run:
        others["configs4_clusters3d_16m_per_gpu"] = extra(
            "clusters3d", 16_000_000, 10, f"BASELINE configs[4]: Clusters 3D weak scaling, 16M per GPU, N={16 * world}M")

    if rank == 0:
        hs = res["hs"]
        peak, peak_src = measured_peaks()
        steps = args.steps
        # force-kernel time PER STEP: one launch on one GPU; interior + boundary launches in slab mode
        # (they overlap on two streams, so their sum can exceed the step's wall time)
        force_ms_step = res["force_ms"] / steps
        achieved = FORCE_BYTES_PER_PARTICLE * n / (force_ms_step * 1e-3) / 1e9
        cap = recorded_capture(args.workload, n) if world == 1 else None
        stats = res["stats"]
        roofline = {
            "bound": "hbm",
            "kernel": "gof::k_force<%s, integrate> (neighbour force + integrator)"
                      % {"clusters": "Clusters", "boids": "Boids", "mixed": "Mixed"}[WORKLOADS[args.workload][0]],
            "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "peak_source": peak_src,
            "traffic": float(cap["dram_bytes_per_launch"]) if cap else None,
            "traffic_source": "committed ncu capture of this workload (profiles/force_kernel_traffic.json), "
                              "not measured in this run" if cap else None,
            "algorithmic_bytes_per_step": FORCE_BYTES_PER_PARTICLE * n,
            "force_ms_per_step": force_ms_step, "launches_timed": res["force_launches"],
            "launches_per_step": res["force_launches"] / steps,
            "share_of_step": res["force_ms"] / res["ms_total"] if res["ms_total"] > 0 else None,
        }
        if cap:
            roofline["recorded_issue_active_pct"] = cap.get("issue_active_pct")
            roofline["recorded_active_lanes"] = cap.get("active_lanes_per_instruction")
        if stats:
            # why the HBM fraction is what it is: the kernel is instruction-issue bound on the
            # neighbour search.  Per particle: stencil candidates (population of its 9 / 27 cells),
            # candidates the kernel evaluates after culling, and pairs really in range.
            roofline.update({
                "candidates_per_particle": stats["candidates"] / stats["n"],
                "pairs_evaluated_per_particle": stats["evaluated"] / stats["n"],
                "pairs_in_range_per_particle": stats["in_range"] / stats["n"],
                "evaluated_per_in_range": stats["evaluated"] / max(stats["in_range"], 1),
                "pairs_evaluated_per_s": stats["evaluated"] / (force_ms_step * 1e-3),
            })
        h2d = n * (9 * 4 + 4) + (n * 4 if world > 1 else 0)
        d2h = n * 9 * 4 + (n * 8 if world > 1 else 0)
        line = {
            "metric": "particle-updates/sec", "value": res["value"], "unit": "particle-updates/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": res["ms_per_step"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": workload_config(args, hs, world * n, world=world,
                                      note=None if world == 1 else
                                      f"one periodic box cut into {world} z-slabs (one per GPU); transport "
                                      f"{res['transport']}"),
            "clocks": res["clocks"],
            "e2e": {"value": res.get("e2e_value"), "unit": "particle-updates/s", "h2d_bytes_per_step": h2d * world,
                    "d2h_bytes_per_step": d2h * world, "steps": args.e2e_steps,
                    "what": ("every step: upload(pinned host pos/vel/acc/types) + one core_step + download(pinned "
                             "host pos/vel/acc) through the C ABI call gof_engine_step_host; a step's input is the "
                             "previous step's output buffers, the copies of consecutive steps overlap (full duplex), "
                             "one host wait at the end of the timed steps") if res.get("e2e_sync_value") else
                            ("upload(pinned host pos/vel/acc/types) + one core_step + download(pinned host "
                             "pos/vel/acc) through the C ABI, synchronised every step"),
                    "synchronised_every_step": res.get("e2e_sync_value") or res.get("e2e_value")},
            "gpu_launches": res["launches"],
            "roofline": roofline,
        }
        if res.get("e2e_interleaved_value"):
            # not the headline `e2e.value` (one system, every step waits for its own round trip over
            # PCIe): what the link and the GPU sustain when the host drives several systems in turn
            line["e2e"]["interleaved_systems"] = {
                "systems": args.e2e_systems, "value": res["e2e_interleaved_value"], "unit": "particle-updates/s",
                "what": (f"{args.e2e_systems} independent systems of N particles driven from the host in turn, each "
                         "in its own loop of gof_engine_step_host calls on its own stream (every step of every "
                         "system uploads its state from pinned host memory and downloads it into the buffers its "
                         "next step starts from): the copies of one system run under the kernels of another")}
        if res.get("phases") and res["phases"]["steps"]:
            line["slab_phases_ms_per_step"] = res["phases"]  # (rank 0's slab; peer-memory transport)
            if res.get("phases_all"):
                line["slab_phases_all_ranks"] = {"columns": ["prefix", "interior_forces", "halo_and_boundary_forces", "step"],
                                                 "ms_per_step": res["phases_all"]}
        if "steady" in res:
            line["steady_state"] = res["steady"]
        if parity is not None:
            line["parity_check"] = parity
        if north is not None:
            line["north_star_16m"] = north
        if others:
            line["other_configs"] = others
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(args.workload, args.seed, args.density, args.cpu_seconds)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
