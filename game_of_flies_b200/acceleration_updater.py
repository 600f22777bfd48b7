"""Launch sites of reference main/acceleration_updater.py with the same names
and argument order, bound to the sm_100a kernels through the C ABI.

The reference indexes its kernels as `kernel[blocks, threads](...)`; here they
are plain functions (grid sizing is the library's business), so call sites drop
the `[blocks, threads]` subscript and pass DeviceArray buffers from
game_of_flies_b200.device.to_device.  The three binning kernels of the
reference (bin_particles, cumsum, set_indices) are one enqueue
(`integrator.setup_bins`); they are not exposed separately because their
intermediate state (arrival-order atomic offsets) is not deterministic.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from .device import DeviceArray, as_ptr, current_stream_ptr

_workspaces: dict = {}


def _workspace(device: torch.device, nbytes: int) -> torch.Tensor:
    key = (device.type, device.index)
    ws = _workspaces.get(key)
    if ws is None or ws.numel() < nbytes + 256:
        ws = torch.empty(nbytes + 256, dtype=torch.uint8, device=device)
        _workspaces[key] = ws
    return ws


def _aligned(t: torch.Tensor) -> int:
    return (t.data_ptr() + 255) // 256 * 256


def set_bin_neighbours(num_bin_x: int, num_bin_y: int, num_bin_z: int, bin_neighbours: np.ndarray):
    """acceleration_updater.py:140-215: fill the half-stencil table (host array, int32)."""
    assert bin_neighbours.dtype == np.int32 and bin_neighbours.flags.c_contiguous
    _lib.check(_lib.load().gof_set_bin_neighbours(int(num_bin_x), int(num_bin_y), int(num_bin_z),
                                                  bin_neighbours.ctypes.data))


def setup_bins_device(pos_x, pos_y, pos_z, grid: _lib.Grid, num_bins, num_particles, particle_bins,
                      particle_bin_counts, bin_offsets, particle_bin_starts, particle_indices):
    lib = _lib.load()
    dev = pos_x.tensor.device
    need = int(lib.gof_workspace_bytes(int(num_particles), int(num_bins), 1))
    ws = _workspace(dev, need)
    _lib.check(lib.gof_setup_bins(as_ptr(pos_x), as_ptr(pos_y), as_ptr(pos_z), int(num_particles),
                                  C.byref(grid), int(num_bins), as_ptr(particle_bins),
                                  as_ptr(particle_bin_counts), as_ptr(bin_offsets),
                                  as_ptr(particle_bin_starts), as_ptr(particle_indices),
                                  _aligned(ws), need, current_stream_ptr(dev)))


def accelerator(pos_x, pos_y, pos_z, vel_x, vel_y, vel_z, limitx, limity, limitz, r_max,
                num_particles, parameter_matrix, particle_type_index_array, acc_x, acc_y, acc_z,
                num_types, boid_acc_x, boid_acc_y, boid_acc_z, boid_vel_x, boid_vel_y, boid_vel_z,
                boid_counts, bin_neighbours, particle_bins, particle_indices, bin_starts, bin_counts,
                wrap_mode: int = _lib.WRAP_REFERENCE):
    """acceleration_updater.py:219-458.

    `parameter_matrix` is the HOST float64 (5, T, T) matrix (the reference passes
    its device copy; the library rebuilds its packed species table from the host
    matrix on every call, which is what makes update_parameter_matrix trivial).
    `bin_neighbours` is accepted and ignored: the kernel derives the stencil.
    """
    lib = _lib.load()
    dev = pos_x.tensor.device
    pm = np.ascontiguousarray(parameter_matrix, dtype=np.float64)
    grid = _lib.make_grid((limitx, limity, limitz), r_max)
    need = int(lib.gof_workspace_bytes(int(num_particles), int(grid.num_cells), int(num_types)))
    ws = _workspace(dev, need)
    _lib.check(lib.gof_accelerator(
        as_ptr(pos_x), as_ptr(pos_y), as_ptr(pos_z), as_ptr(vel_x), as_ptr(vel_y), as_ptr(vel_z),
        int(num_particles), C.byref(grid), pm.ctypes.data, int(num_types),
        as_ptr(particle_type_index_array), as_ptr(acc_x), as_ptr(acc_y), as_ptr(acc_z),
        as_ptr(boid_acc_x), as_ptr(boid_acc_y), as_ptr(boid_acc_z), as_ptr(boid_vel_x),
        as_ptr(boid_vel_y), as_ptr(boid_vel_z), as_ptr(boid_counts), as_ptr(particle_bins),
        as_ptr(particle_indices), as_ptr(bin_starts), as_ptr(bin_counts), int(wrap_mode),
        _aligned(ws), need, current_stream_ptr(dev)))
