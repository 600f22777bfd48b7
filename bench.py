#!/usr/bin/env python
"""bench.py -- particle-updates/sec of the Game_of_Flies hot path on B200.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --steps K --warmup W    # the reference algorithm on host cores

One "step" is one Simulation.core_step (bin -> neighbour forces -> integrate)
over the synthetic workload named in `config.workload`.  At N=1 that is
BASELINE.json configs[1]: Clusters 3D, 5 species, N = 1,000,000, periodic box,
at the number density of the reference's own 3D run (hive.py: 14,000 particles
in 100^3 = 0.014 / unit^3), uniform random start, adaptive timestep.

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


class ClockSampler:
    """nvidia-smi clocks and throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu_index = gpu_index
        self.rows = []
        self.proc = None
        self.thread = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.gpu_index}", f"--query-gpu={self.QUERY}",
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
                "sm_max_mhz": float(max(mx)) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


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


def recorded_traffic(workload: str, n: int):
    """dram bytes per force-kernel launch from the committed ncu capture, if one matches."""
    r = recorded_capture(workload, n)
    return float(r["dram_bytes_per_launch"]) if r else None


def oracle_sim_from(hs, n_threads):
    from oracle import oracle as orc
    init = hs["init"]
    f = lambda *a: np.stack(a).astype(np.float32).astype(np.float64)
    return orc.OracleSimulation(f(init["pos_x"], init["pos_y"], init["pos_z"]),
                                f(init["vel_x"], init["vel_y"], init["vel_z"]),
                                f(init["acc_x"], init["acc_y"], init["acc_z"]),
                                hs["particle_type_index_array"], hs["parameter_matrix"],
                                hs["limits"], float(hs["r_max"]), timestep=None, n_threads=n_threads)


def cpu_baseline(workload: str, seed: int, density, budget_s: float = 20.0):
    """The oracle port (reference algorithm: half stencil + atomics, float64, OpenMP over
    particles like numba-parallel) on this host's cores, on a bounded sample of the workload:
    the same species, parameters, density and box shape at a particle count that one step
    finishes in seconds.  particle-updates/s is intensive at fixed density."""
    from oracle import oracle as orc
    threads = orc.max_threads()
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
    """--impl reference: the reference's algorithm on the host cores (oracle port; the
    reference itself is numba-CUDA Python and cannot be compiled into oracle/_ref)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle as orc
    threads = orc.max_threads()
    n_sample = args.reference_particles
    hs = build_workload(args.workload, n_sample, args.seed, args.density)
    o = oracle_sim_from(hs, threads)
    for _ in range(args.warmup):
        o.core_step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        o.core_step()
    el = time.perf_counter() - t0
    value = n_sample * args.steps / el
    line = {
        "impl": "reference",
        "metric": "particle-updates/sec", "value": value, "unit": "particle-updates/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * el / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, hs, n_sample, note="bounded sample of the workload on host cores"),
        "cpu_baseline": {"value": value, "unit": "particle-updates/s", "cores": threads, "kind": "port",
                         "sample": f"{args.steps} steps at N={n_sample}, same density/species/parameters as the "
                                   f"GPU arm's workload, {threads} OpenMP threads"},
        "e2e": {"value": value, "unit": "particle-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(args, hs, n_total, note=None):
    kind, species, dims = WORKLOADS[args.workload]
    cfg = {
        "workload": f"{kind.capitalize()} {dims}D, {species} species, N={n_total}, periodic box "
                    f"{tuple(round(float(x), 2) for x in hs['box'])}, density {hs['density']:.4g}, "
                    f"r_max {float(hs['r_max']):.3f}, uniform random start, "
                    + ("adaptive timestep" if args.timestep is None else f"fixed timestep {args.timestep}"),
        "particles_per_gpu": int(n_total // max(int(os.environ.get("WORLD_SIZE", "1")), 1)),
        "bins": [int(hs["num_bin_x"]), int(hs["num_bin_y"]), int(hs["num_bin_z"])],
        "wrap_mode": getattr(args, "wrap", "reference"),
        "cache": ("L2 flushed between timed steps (256 MiB overwrite outside the per-step event pairs)"
                  if getattr(args, "flush_l2", True) else "no L2 flush: per-step working set ~100 MB at N = 1M"),
    }
    if note:
        cfg["note"] = note
    return cfg


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="clusters3d", choices=sorted(WORKLOADS))
    ap.add_argument("--particles", type=int, default=1_000_000, help="particles per GPU")
    ap.add_argument("--reference-particles", type=int, default=100_000)
    ap.add_argument("--density", type=float, default=None)
    ap.add_argument("--seed", type=int, default=434)
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-flush-l2", dest="flush_l2", action="store_false",
                    help="do not overwrite a 256 MiB buffer between timed steps")
    ap.add_argument("--wrap", default="reference", choices=["reference", "minimage"],
                    help="periodic wrap semantics: the reference's (z-wrap quirk included) or minimum image")
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

    lib = _lib.load()
    _lib.check(lib.gof_set_force_split(args.force_split))
    n = args.particles
    cand = None
    wrap_mode = _lib.WRAP_REFERENCE if args.wrap == "reference" else _lib.WRAP_MINIMAGE
    if world == 1:
        # ---- one GPU: the resident engine on the whole periodic box -------------------------
        hs = build_workload(args.workload, n, args.seed, args.density)
        init = hs["init"]
        pin = lambda a, dt=np.float32: torch.from_numpy(np.ascontiguousarray(a, dtype=dt)).pin_memory()
        h_pos = [pin(init[k]) for k in ("pos_x", "pos_y", "pos_z")]
        h_vel = [pin(init[k]) for k in ("vel_x", "vel_y", "vel_z")]
        h_acc = [pin(init[k]) for k in ("acc_x", "acc_y", "acc_z")]
        h_types = pin(hs["particle_type_index_array"], np.int32)
        o_pos = [torch.empty(n, dtype=torch.float32).pin_memory() for _ in range(3)]
        o_vel = [torch.empty(n, dtype=torch.float32).pin_memory() for _ in range(3)]
        o_acc = [torch.empty(n, dtype=torch.float32).pin_memory() for _ in range(3)]
        eng = ParticleEngine(n, hs["parameter_matrix"], hs["limits"], float(hs["r_max"]), device=device,
                             wrap_mode=wrap_mode)
        eng.upload(h_pos, h_types, vel=h_vel, acc=h_acc)
        step = lambda k: eng.step(k, args.timestep)

        def e2e_step():
            nonlocal h_pos, h_vel, h_acc, o_pos, o_vel, o_acc
            eng.upload(h_pos, h_types, vel=h_vel, acc=h_acc)
            eng.step(1, args.timestep)
            eng.download_into(pos=o_pos, vel=o_vel, acc=o_acc)
            torch.cuda.current_stream(device).synchronize()
            # the next step of a host-driven loop starts from what came back
            h_pos, o_pos = o_pos, h_pos
            h_vel, o_vel = o_vel, h_vel
            h_acc, o_acc = o_acc, h_acc
    else:
        # ---- several GPUs: z-slab decomposition, one slab per rank, weak scaling -----------------
        # the box of the one-GPU workload stacked `world` times along z; every rank starts with
        # `n` particles uniform in its own slab
        from game_of_flies_b200.slab import DistComm, SlabEngine, SlabRunner, split_layers
        hs = build_workload(args.workload, 1000, args.seed, args.density)  # species parameters only
        kind, species, dims = WORKLOADS[args.workload]
        if dims != 3:
            raise SystemExit("the slab decomposition needs a 3D workload")
        rho = hs["density"]
        side = float((n / rho) ** (1.0 / 3.0))
        limits = np.array([side, side, side * world])
        hs["limits"], hs["box"] = limits, tuple(limits)
        grid = _lib.make_grid(limits, float(hs["r_max"]))
        hs["num_bin_x"], hs["num_bin_y"], hs["num_bin_z"] = grid.num_bin_x, grid.num_bin_y, grid.num_bin_z
        layers = split_layers(int(grid.num_bin_z), world)
        lo, hi = layers[rank]
        rng = np.random.default_rng(args.seed + rank)
        pos = rng.random((3, n))
        pos[0] *= side
        pos[1] *= side
        pos[2] = (lo + pos[2] * (hi - lo)) * grid.bin_size_z
        pos = pos.astype(np.float32)
        types = np.repeat(np.arange(species, dtype=np.int32), -(-n // species))[:n]
        ids = np.arange(n, dtype=np.int32) + rank * n
        zeros = np.zeros((3, n), dtype=np.float32)
        eng = SlabEngine(hs["parameter_matrix"], limits, float(hs["r_max"]), lo, hi, capacity=int(1.3 * n) + 8192,
                         device=device, wrap_mode=wrap_mode)
        eng.upload(pos, types, ids, vel=zeros, acc=zeros)
        runner = SlabRunner({rank: eng}, world, DistComm(world, device=device))
        step = lambda k: runner.step(k, args.timestep)
        state = {"pos": pos, "vel": zeros, "acc": zeros, "types": types, "ids": ids}

        def e2e_step():
            eng.upload(state["pos"], state["types"], state["ids"], vel=state["vel"], acc=state["acc"])
            runner.step(1, args.timestep)
            d = eng.download()
            state["pos"], state["vel"], state["acc"], state["ids"] = d["pos"], d["vel"], d["acc"], d["ids"]
            # types travel with the particles: recover them from the ids (species are id ranges per rank)
            state["types"] = types[(d["ids"] % n)]

    # ---- device-resident throughput ------------------------------------------------------------
    step(args.warmup)
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    eng.timing(True)
    lib.gof_reset_launch_count()
    # Every timed step starts with a cold L2: a 256 MiB buffer (2x the 126 MB L2) is overwritten
    # between steps, outside the per-step event pairs, because the per-step working set at
    # N = 1M (~100 MB) would otherwise stay L2-resident from one step to the next.
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=device) if args.flush_l2 else None
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    stops = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    barrier()
    for k in range(args.steps):
        if flush is not None:
            flush.fill_(k & 0xFF)
        starts[k].record()
        step(1)
        stops[k].record()
    barrier()
    launches = int(lib.gof_launch_count())
    ms_total = float(sum(a.elapsed_time(b) for a, b in zip(starts, stops)))
    force_ms, force_launches = eng.force_time()
    eng.timing(False)
    clocks = sampler.stop()
    if world == 1:
        cand = eng.candidate_pairs()
    t = torch.tensor([ms_total], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    ms_per_step = ms_total / args.steps
    value = world * n * args.steps / (ms_total * 1e-3)

    # ---- end to end through the public API with host buffers ------------------------------
    # every step: host state -> device, one core_step, full state back to the host
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.e2e_steps):
        e2e_step()
    barrier()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_s = float(t.item())
    e2e_value = world * n * args.e2e_steps / e2e_s
    h2d = n * (9 * 4 + 4) + (n * 4 if world > 1 else 0)
    d2h = n * 9 * 4 + (n * 4 if world > 1 else 0)

    if world > 1 and rank == 0 and getattr(runner, "phase_ms", None):
        pm = runner.phase_ms
        print(f"[slab timing] per step: prefix (hash, migration, sort) {pm['prefix'] / max(pm['steps'], 1):.4f} ms, "
              f"forces + halo {pm['forces'] / max(pm['steps'], 1):.4f} ms", file=sys.stderr, flush=True)
    if rank == 0:
        peak, peak_src = measured_peaks()
        force_avg_ms = force_ms / max(force_launches, 1)
        achieved = FORCE_BYTES_PER_PARTICLE * n / (force_avg_ms * 1e-3) / 1e9
        roofline = {
            "bound": "hbm",
            "kernel": "gof::k_force<%s, integrate> (neighbour force + integrator)"
                      % {"clusters": "Clusters", "boids": "Boids", "mixed": "Mixed"}[WORKLOADS[args.workload][0]],
            "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "peak_source": peak_src,
            "traffic": recorded_traffic(args.workload, n),
            # the kernel's real limit, from the same ncu capture (profiles/): issue slots in use
            "issue_active_pct": (recorded_capture(args.workload, n) or {}).get("issue_active_pct"),
            "active_lanes": (recorded_capture(args.workload, n) or {}).get("active_lanes_per_instruction"),
            "algorithmic_bytes_per_launch": FORCE_BYTES_PER_PARTICLE * n,
            "avg_launch_ms": force_avg_ms, "launches_timed": force_launches,
            "share_of_step": force_ms / ms_total if ms_total > 0 else None,
            # why the HBM fraction is what it is: the kernel is instruction-issue bound on the
            # neighbour search (stencil candidates = population of a particle's 9 / 27 cells; the
            # Clusters kernel culls about 45 % of them per warp before evaluating)
            "candidate_pairs_per_particle": (cand / n) if cand is not None else None,
            "stencil_candidates_per_s": (cand / (force_avg_ms * 1e-3)) if cand is not None else None,
        }
        line = {
            "metric": "particle-updates/sec", "value": value, "unit": "particle-updates/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": workload_config(args, hs, world * n,
                                      note=None if world == 1 else
                                      f"one periodic box cut into {world} z-slabs (one per GPU); halo layers and "
                                      "migrating particles move with NCCL send/recv, message sizes with two small "
                                      "all-gathers per step, the adaptive timestep with a one-word MAX all-reduce"),
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "particle-updates/s", "h2d_bytes_per_step": h2d * world,
                    "d2h_bytes_per_step": d2h * world, "steps": args.e2e_steps,
                    "what": "upload(host pos/vel/acc/types) + one core_step + download(host pos/vel/acc) through "
                            "the C ABI, synchronised every step"},
            "gpu_launches": launches,
            "roofline": roofline,
        }
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(args.workload, args.seed, args.density, args.cpu_seconds)
        print(json.dumps(line), flush=True)
    eng.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
