"""ParticleEngine: the resident, cell-sorted device state and its fused step.

Thin wrapper over the gof_engine_* entry points of include/gof_b200.h.  torch
provides the device memory (one arena tensor) and the stream; all compute is
in libgof_b200.so.  This is the path Simulation.core_step uses
(reference main/simulation.py:193-247).
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from .device import _require_cuda


def _any_ptr(x) -> int:
    """Pointer of a host numpy array, a (pinned) CPU tensor or a CUDA tensor; 0 for None."""
    if x is None:
        return 0
    if isinstance(x, np.ndarray):
        assert x.flags.c_contiguous
        return x.ctypes.data
    if isinstance(x, torch.Tensor):
        assert x.is_contiguous()
        return x.data_ptr()
    if hasattr(x, "ptr"):
        return x.ptr
    raise TypeError(type(x))


class ParticleEngine:
    def __init__(self, num_particles: int, parameter_matrix: np.ndarray, limits, r_max: float,
                 device=None, wrap_mode: int = _lib.WRAP_REFERENCE):
        self.lib = _lib.load()
        self.device = _require_cuda(device)
        self.index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self.n = int(num_particles)
        self.wrap_mode = int(wrap_mode)
        pm = np.ascontiguousarray(parameter_matrix, dtype=np.float64)
        assert pm.ndim == 3 and pm.shape[0] == 5 and pm.shape[1] == pm.shape[2]
        self.num_types = int(pm.shape[1])
        self.grid = _lib.make_grid(limits, r_max)
        self._handle = C.c_void_p()
        _lib.check(self.lib.gof_engine_create(C.byref(self._handle), self.index, self.n, self.num_types,
                                              C.byref(self.grid), pm.ctypes.data))
        nbytes = int(self.lib.gof_engine_arena_bytes(self._handle))
        with torch.cuda.device(self.index):
            self.arena = torch.empty(nbytes + 256, dtype=torch.uint8, device=self.device)
        base = self.arena.data_ptr()
        self._arena_ptr = (base + 255) // 256 * 256
        _lib.check(self.lib.gof_engine_bind_arena(self._handle, self._arena_ptr, nbytes))
        self.arena_bytes = nbytes

    # -- helpers ---------------------------------------------------------------------------
    def _stream(self) -> int:
        return torch.cuda.current_stream(self.index).cuda_stream

    def close(self):
        if getattr(self, "_handle", None) is not None and self._handle.value:
            torch.cuda.synchronize(self.index)
            self.lib.gof_engine_destroy(self._handle)
            self._handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- state ------------------------------------------------------------------------------
    def set_parameters(self, parameter_matrix: np.ndarray):
        pm = np.ascontiguousarray(parameter_matrix, dtype=np.float64)
        _lib.check(self.lib.gof_engine_set_parameters(self._handle, pm.ctypes.data, self._stream()))

    def upload(self, pos, types, vel=None, acc=None):
        """pos/vel/acc: three fp32 arrays each (numpy, pinned CPU tensors or CUDA tensors)."""
        def f32(a):
            if a is None:
                return None
            if isinstance(a, np.ndarray):
                return np.ascontiguousarray(a, dtype=np.float32)
            return a
        p = [f32(a) for a in pos]
        v = [f32(a) for a in vel] if vel is not None else [None] * 3
        ac = [f32(a) for a in acc] if acc is not None else [None] * 3
        t = np.ascontiguousarray(types, dtype=np.int32) if isinstance(types, np.ndarray) else types
        keep = (p, v, ac, t)  # keep the converted host arrays alive until the copies are enqueued
        _lib.check(self.lib.gof_engine_upload(self._handle, *[_any_ptr(a) for a in p],
                                              *[_any_ptr(a) for a in v], *[_any_ptr(a) for a in ac],
                                              _any_ptr(t), self._stream()))
        # pageable sources are staged before the call returns; pinned ones must outlive the copy
        if any(isinstance(a, torch.Tensor) and not a.is_cuda for grp in (p, v, ac) for a in grp if a is not None):
            torch.cuda.current_stream(self.index).synchronize()
        del keep

    def step(self, n_steps: int = 1, timestep=None):
        ts = -1.0 if timestep is None else float(timestep)
        _lib.check(self.lib.gof_engine_step(self._handle, int(n_steps), ts, self.wrap_mode, self._stream()))

    def accelerate(self):
        _lib.check(self.lib.gof_engine_accelerate(self._handle, self.wrap_mode, self._stream()))

    def download(self, pos=True, vel=False, acc=False):
        """Returns dict of (3, N) fp32 numpy arrays in particle order (synchronises)."""
        out = {}
        ptrs = []
        for name, want in (("pos", pos), ("vel", vel), ("acc", acc)):
            if want:
                out[name] = np.empty((3, self.n), dtype=np.float32)
                ptrs += [out[name][k].ctypes.data for k in range(3)]
            else:
                ptrs += [0, 0, 0]
        _lib.check(self.lib.gof_engine_download(self._handle, *ptrs, self._stream()))
        torch.cuda.current_stream(self.index).synchronize()
        return out

    def download_into(self, pos=None, vel=None, acc=None):
        """Asynchronous variant: each of pos/vel/acc is None or three buffers (pinned tensors / CUDA tensors)."""
        ptrs = []
        for grp in (pos, vel, acc):
            ptrs += [_any_ptr(a) for a in grp] if grp is not None else [0, 0, 0]
        _lib.check(self.lib.gof_engine_download(self._handle, *ptrs, self._stream()))

    def step_host(self, state_in, types, state_out, timestep=None):
        """One step of a host-driven loop (gof_engine_step_host): `state_in` / `state_out` are nine
        fp32 buffers each (pos xyz, vel xyz, acc xyz; pinned CPU tensors, numpy arrays or CUDA
        tensors).  Uploads and step on the current stream, downloads on the engine's own stream;
        consecutive calls that use two sets of buffers alternately overlap their transfers.
        Returns at once: host_wait() before reading `state_out`."""
        assert len(state_in) == 9 and len(state_out) == 9
        arr = C.c_void_p * 9
        ins = arr(*[_any_ptr(a) for a in state_in])
        outs = arr(*[_any_ptr(a) for a in state_out])
        ts = -1.0 if timestep is None else float(timestep)
        _lib.check(self.lib.gof_engine_step_host(self._handle, ins, _any_ptr(types), outs, ts, self.wrap_mode,
                                                 self._stream()))

    def host_wait(self):
        """Block until the downloads of the last step_host are complete."""
        _lib.check(self.lib.gof_engine_host_wait(self._handle))

    def snapshot_into(self, xyz):
        """Positions as [N][3] fp32 into a pinned tensor / numpy array / CUDA tensor (async)."""
        _lib.check(self.lib.gof_engine_snapshot(self._handle, _any_ptr(xyz), self._stream()))

    def snapshot_stage(self):
        """Un-permute the positions into the device staging buffer (current stream)."""
        _lib.check(self.lib.gof_engine_snapshot_stage(self._handle, self._stream()))

    def snapshot_copy(self, xyz):
        """Copy the staged frame out (current stream: make it a side stream)."""
        _lib.check(self.lib.gof_engine_snapshot_copy(self._handle, _any_ptr(xyz), self._stream()))

    def read_bins(self):
        """particle_bins, bin_counts, bin_starts, particle_indices of the last step (reference layout)."""
        cells = int(self.grid.num_cells)
        pb = np.empty(self.n, dtype=np.int32)
        pi = np.empty(self.n, dtype=np.int32)
        cnt = np.empty(cells, dtype=np.int32)
        st = np.empty(cells, dtype=np.int32)
        _lib.check(self.lib.gof_engine_read_bins(self._handle, pb.ctypes.data, cnt.ctypes.data,
                                                 st.ctypes.data, pi.ctypes.data, self._stream()))
        torch.cuda.current_stream(self.index).synchronize()
        return pb, cnt, st, pi

    def last_timestep(self) -> float:
        out = C.c_float()
        _lib.check(self.lib.gof_engine_last_timestep(self._handle, C.byref(out), self._stream()))
        return float(out.value)

    def set_graph(self, enable: bool):
        """One CUDA graph launch per step (default) or individual kernel launches; same results."""
        _lib.check(self.lib.gof_engine_set_graph(self._handle, 1 if enable else 0))

    def timing(self, enable: bool):
        """Bracket every force-kernel launch with CUDA events (bench.py roofline)."""
        _lib.check(self.lib.gof_engine_timing(self._handle, 1 if enable else 0))

    def force_time(self):
        """(total milliseconds, launches) of the force kernel since timing(True)."""
        ms, k = C.c_float(), C.c_int()
        _lib.check(self.lib.gof_engine_force_time(self._handle, C.byref(ms), C.byref(k)))
        return float(ms.value), int(k.value)

    def pair_stats(self) -> dict:
        """Per-launch totals of the neighbour search on the last step's snapshot (diagnostics for
        bench.py): stencil candidates, force evaluations after culling, pairs really in range."""
        out = (C.c_ulonglong * 3)()
        _lib.check(self.lib.gof_engine_pair_stats(self._handle, self.wrap_mode, out, self._stream()))
        return {"n": self.n, "candidates": int(out[0]), "evaluated": int(out[1]), "in_range": int(out[2])}

    def candidate_pairs(self) -> int:
        """Distance tests the force kernel performs for the current binning: for every particle
        the population of its 9 / 27 stencil cells (host arithmetic on the cell counts)."""
        _, cnt, _, _ = self.read_bins()
        g = self.grid
        nz = max(int(g.num_bin_z), 1)
        c = cnt.astype(np.int64).reshape(nz, int(g.num_bin_y), int(g.num_bin_x))
        nb = np.zeros_like(c)
        for dz in ((-1, 0, 1) if g.num_bin_z > 0 else (0,)):
            for dy in (-1, 0, 1):
                for dx in (-1, 0, 1):
                    nb += np.roll(c, (dz, dy, dx), axis=(0, 1, 2))
        return int((c * nb).sum())
