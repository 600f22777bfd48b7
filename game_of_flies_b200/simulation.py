"""Simulation: drop-in for reference main/simulation.py's class of the same name.

Same constructor, attributes and methods (core_step, update,
update_parameter_matrix, record, blind_run, bench_run), so hive.py, vizman.py,
ui_container.py and npy2vid.py work against it unchanged.  What changes is
underneath: the state lives on the B200 as cell-sorted float4 arrays inside a
ParticleEngine and one core_step is one enqueue of the fused CUDA pipeline
instead of eight numba kernels and a host round trip for the timestep.

`host_state` is the host-only half of the reference constructor
(simulation.py:12-139: species kinds, parameter scaling, r_max, bins); it is a
pure numpy function so that it can be tested without a GPU.
"""
from __future__ import annotations

import os
import time

import numpy as np

from . import _lib
from .state_parameters import initialise


def host_state(clus_types, boid_types, seed=None, limits=(100, 100, 0)) -> dict:
    """Everything Simulation.__init__ computes on the host (reference lines 12-139).

    Draws from numpy's global generator in the reference's order: `initialise`
    seeds and draws positions and parameters, then one np.random.random() per
    mixed (Clusters x Boids) species pair decides that pair's kind.
    """
    clus_types = np.asarray(clus_types)
    boid_types = np.asarray(boid_types)
    if len(clus_types) == 1 and clus_types[0] == 0:
        clus_types = np.array([])
    if np.sum(clus_types) == 0:
        all_types = np.array(boid_types, dtype=np.int32)
    elif np.sum(boid_types) == 0:
        all_types = np.array(clus_types, dtype=np.int32)
    else:
        all_types = np.array(np.append(clus_types, boid_types), dtype=np.int32)

    init = initialise(all_types, seed=seed, limits=limits)
    pm = init["parameter_matrix"]
    num_types = int(all_types.size)
    n_clus = int(clus_types.size)

    # interaction kind per ordered pair: 0 = Clusters, 1 = Boids; mixed pairs are a coin flip
    kind = pm[-1]
    kind[:, :] = 0
    for i in range(num_types):
        for j in range(num_types):
            both_boid = i >= n_clus and j >= n_clus
            both_clus = i < n_clus and j < n_clus
            if both_boid:
                kind[i, j] = 1
            elif both_clus:
                kind[i, j] = 0
            elif np.random.random() > 0.5 and i <= j:
                kind[i, j] = 1
            if i > j:
                kind[i, j] = kind[j, i]

    boid = kind == 1
    clus = kind == 0
    # Boids planes: r_max, separation, alignment, cohesion
    for plane, (scale, shift) in enumerate(((2, 12), (4, 2), (2, 2), (0.1, -0.05))):
        pm[plane][boid] *= scale
        pm[plane][boid] += shift
    for t in range(num_types):
        if kind[t, t] == 1:
            pm[3, t, t] = abs(pm[3, t, t])
    # Clusters planes: r_min, r_max, f_min, f_max
    for plane, (scale, shift) in enumerate(((3, 5), (5, 8), (3, -10), (12, -6))):
        pm[plane][clus] *= scale
        pm[plane][clus] += shift

    r_max = np.max(pm[1, :, :])  # plane 1 whatever the kind, as in the reference (line 98)
    lim = np.array(init["limits"])
    nbx = int(np.floor(lim[0] / r_max))
    nby = int(np.floor(lim[1] / r_max))
    nbz = int(np.floor(lim[2] / r_max))
    bsx = lim[0] / nbx
    bsy = lim[1] / nby
    if lim[2] != 0:
        bsz = lim[2] / nbz
        cells, width = nbx * nby * nbz, 14
    else:
        bsz, nbz = 0, 0
        cells, width = nbx * nby, 5
    num_bins = 1
    while num_bins < cells:
        num_bins *= 2

    return dict(
        init=init, all_types=all_types, num_types=num_types,
        num_particles=np.sum(init["n_type_array"]),
        particle_type_index_array=np.array(init["particle_type_indx_array"], dtype="int32"),
        parameter_matrix=pm, r_max=r_max, limits=lim,
        num_bin_x=nbx, num_bin_y=nby, num_bin_z=nbz,
        bin_size_x=bsx, bin_size_y=bsy, bin_size_z=bsz,
        num_cells=cells, num_bins=num_bins, stencil_width=width,
    )


class Simulation:
    def __init__(self, clus_types, boid_types, seed=None, limits=(100, 100, 0), device=None,
                 wrap_mode: int = _lib.WRAP_REFERENCE) -> None:
        # FOR NUMPY ARRAY EXPORTING IN BLIND RUN (same attributes as the reference)
        self.record_path = None
        self.export_set = False
        self.seed = seed
        self.timestep = None  # None = adaptive (integrator.py:174-188)

        hs = host_state(clus_types, boid_types, seed=seed, limits=limits)
        self.init = hs["init"]
        for k in ("pos_x", "pos_y", "pos_z", "vel_x", "vel_y", "vel_z", "acc_x", "acc_y", "acc_z"):
            setattr(self, k, self.init[k])
        self.limits = hs["limits"]
        self.num_particles = hs["num_particles"]
        self.particle_type_index_array = hs["particle_type_index_array"]
        self.num_types = hs["num_types"]
        self.parameter_matrix = hs["parameter_matrix"]
        self.r_max = hs["r_max"]
        for k in ("num_bin_x", "num_bin_y", "num_bin_z", "bin_size_x", "bin_size_y", "bin_size_z",
                  "num_bins"):
            setattr(self, k, hs[k])
        # kept for callers that inspect it; the CUDA kernel derives the stencil itself
        from .acceleration_updater import set_bin_neighbours
        self.bin_neighbours = np.zeros((hs["num_cells"], hs["stencil_width"]), dtype=np.int32)
        set_bin_neighbours(self.num_bin_x, self.num_bin_y, self.num_bin_z, self.bin_neighbours)
        self.threads = 512  # reference launch configuration, unused here
        self.blocks = int(np.ceil(self.num_particles / self.threads))

        self.output = np.zeros((self.num_particles, 3))

        # ---- device side ---------------------------------------------------------------
        import torch
        from .engine import ParticleEngine
        n = int(self.num_particles)
        self.engine = ParticleEngine(n, self.parameter_matrix, self.limits, float(self.r_max),
                                     device=device, wrap_mode=wrap_mode)
        self.engine.upload((self.pos_x, self.pos_y, self.pos_z), self.particle_type_index_array,
                           vel=(self.vel_x, self.vel_y, self.vel_z),
                           acc=(self.acc_x, self.acc_y, self.acc_z))
        # snapshots stream to pinned host memory on a side stream (viewer / recorder): two pinned
        # frames, so that frame k is copied out (and written to disk) while step k+1 computes
        self._xyz_pinned = [torch.empty((n, 3), dtype=torch.float32, pin_memory=True) for _ in range(2)]
        self._snap_stream = torch.cuda.Stream(device=self.engine.index)
        self._staged = torch.cuda.Event()      # un-permute kernel of the latest frame done
        self._copied = [torch.cuda.Event(), torch.cuda.Event()]  # frame landed in pinned[b]
        self._frame_buf = 0
        self._copy_pending = False
        self._torch = torch

    # ------------------------------------------------------------------------------------
    def core_step(self):
        """setup_bins + integrate (reference simulation.py:193-247) as one enqueue."""
        self.engine.step(1, self.timestep)

    def update_parameter_matrix(self, i, j, params):
        self.parameter_matrix[:5, i, j] = params
        self.engine.set_parameters(self.parameter_matrix)

    def _snapshot_begin(self) -> int:
        """Enqueue the snapshot of the current state; returns the pinned buffer it will land in.
        Compute stream: un-permute kernel into the device staging frame (must run before the next
        step rewrites the state).  Side stream: the device-to-host copy, overlapping that step."""
        torch = self._torch
        main = torch.cuda.current_stream(self.engine.index)
        b = self._frame_buf
        self._frame_buf ^= 1
        if self._copy_pending:
            main.wait_event(self._copied[b ^ 1])  # the staging frame is free once the last copy left it
        self.engine.snapshot_stage()
        self._staged.record(main)
        with torch.cuda.stream(self._snap_stream):
            self._snap_stream.wait_event(self._staged)
            self.engine.snapshot_copy(self._xyz_pinned[b])
            self._copied[b].record(self._snap_stream)
        self._copy_pending = True
        return b

    def _snapshot_end(self, b: int):
        """Wait for frame `b` and publish it as `output` (float64 [N][3], particle order)."""
        self._copied[b].synchronize()
        xyz = self._xyz_pinned[b].numpy()
        self.output[:, :] = xyz
        self.pos_x[:] = xyz[:, 0]
        self.pos_y[:] = xyz[:, 1]
        self.pos_z[:] = xyz[:, 2]

    def update(self):
        self.core_step()
        self._snapshot_end(self._snapshot_begin())

    def synchronize(self):
        self._torch.cuda.synchronize(self.engine.index)

    def record(self):
        if not self.export_set:
            os.mkdir(self.record_path)
            os.mkdir(os.path.join(self.record_path, "frames"))
            np.savez(
                os.path.join(self.record_path, "state.npz"),
                type_array=self.particle_type_index_array,
                limits=self.limits,
                parameter_matrix=self.parameter_matrix,
                num_particles=self.num_particles,
                seed=self.seed,
            )
            self.frame = 0
            self.export_set = True
        np.save(os.path.join(self.record_path, "frames", f"{self.frame:05}"), self.output)

    def _run(self, n_steps, record, progress):
        if record is not None:
            self.record_path = record
            self.record()
        steps = range(n_steps)
        if progress:
            from tqdm import tqdm
            steps = tqdm(steps)
        start = time.perf_counter()
        pending = None
        for _ in steps:
            if record is not None:
                # pipelined update() + record(): frame k-1 is published and written while the
                # GPU computes step k (same files, same order as the reference's loop)
                self.core_step()
                b = self._snapshot_begin()
                if pending is not None:
                    self._snapshot_end(pending)
                    self.record()
                    self.frame += 1
                pending = b
            else:
                self.core_step()
        if pending is not None:
            self._snapshot_end(pending)
            self.record()
            self.frame += 1
        self.synchronize()
        return time.perf_counter() - start

    def blind_run(self, n_steps, record=None):
        elapsed = self._run(n_steps, record, progress=True)
        print(f"Total time in ms: {1e3 * elapsed}")

    def bench_run(self, n_steps, record=None):
        print(self._run(n_steps, record, progress=False))
