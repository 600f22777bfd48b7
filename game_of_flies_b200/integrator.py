"""Mirror of reference main/integrator.py: `integrate` and `setup_bins` with the
reference's argument lists (minus nothing: `blocks`/`threads` are accepted and
ignored), plus the element-wise kernels they sequence.  Buffers are
DeviceArray objects; everything runs on the current torch CUDA stream through
the C ABI.  This is the reference-layout (SoA, particle-order) API; the
resident engine (engine.py) fuses the same work into far fewer passes.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib
from .acceleration_updater import accelerator, setup_bins_device
from .device import as_ptr, current_stream_ptr, device_array, to_device

fac = 0.95      # integrator.py:13 (compiled into the kernels as gof::kDamping)
max_spd = 20    # integrator.py:14 (gof::kMaxSpeed)

_self_kind_cache: dict = {}


def _self_kind(parameter_matrix, device):
    """parameter_matrix[-1, t, t] as a device int32 array (cached per matrix content)."""
    pm = np.asarray(parameter_matrix)
    diag = np.ascontiguousarray(np.diagonal(pm[-1]) != 0, dtype=np.int32)
    key = (device.type, device.index, diag.tobytes())
    arr = _self_kind_cache.get(key)
    if arr is None:
        arr = to_device(diag, device)
        _self_kind_cache[key] = arr
    return arr


def step1(vel_x, vel_y, vel_z, acc_x, acc_y, acc_z, num_particles, timestep=0.02):
    """integrator.py:33-55."""
    _lib.check(_lib.load().gof_step1(as_ptr(vel_x), as_ptr(vel_y), as_ptr(vel_z), as_ptr(acc_x),
                                     as_ptr(acc_y), as_ptr(acc_z), int(num_particles), float(timestep),
                                     current_stream_ptr(vel_x.tensor.device)))


def step2(pos_x, pos_y, pos_z, vel_x, vel_y, vel_z, acc_x, acc_y, acc_z, particle_tia, parameter_matrix,
          num_particles, timestep=0.02):
    """integrator.py:58-88 (parameter_matrix: host matrix, only its kind diagonal is used)."""
    dev = pos_x.tensor.device
    sk = _self_kind(parameter_matrix, dev)
    _lib.check(_lib.load().gof_step2(as_ptr(pos_x), as_ptr(pos_y), as_ptr(pos_z), as_ptr(vel_x),
                                     as_ptr(vel_y), as_ptr(vel_z), as_ptr(acc_x), as_ptr(acc_y),
                                     as_ptr(acc_z), as_ptr(particle_tia), as_ptr(sk), int(num_particles),
                                     float(timestep), current_stream_ptr(dev)))


def boundary_condition(pos_x, pos_y, pos_z, limitx, limity, limitz, num_particles):
    """integrator.py:91-114."""
    _lib.check(_lib.load().gof_boundary_condition(as_ptr(pos_x), as_ptr(pos_y), as_ptr(pos_z),
                                                  float(limitx), float(limity), float(limitz),
                                                  int(num_particles),
                                                  current_stream_ptr(pos_x.tensor.device)))


def max_sq_speed(vel_x, vel_y, vel_z, num_particles, out=None):
    """set_sq_speed + max_reduction (integrator.py:20-30, 117-129) -> one-element device array."""
    dev = vel_x.tensor.device
    if out is None:
        out = device_array(1, torch.float32, dev)
    _lib.check(_lib.load().gof_max_sq_speed(as_ptr(vel_x), as_ptr(vel_y), as_ptr(vel_z),
                                            int(num_particles), as_ptr(out), current_stream_ptr(dev)))
    return out


def adaptive_timestep(max_sq: float) -> float:
    """integrator.py:183-188, on the host in float64 like the reference."""
    t = float(np.sqrt(max_sq))
    return 0.4 / t if t > 1e-6 else 0.1


def integrate(pos_x, pos_y, pos_z, vel_x, vel_y, vel_z, limits, r_max, num_particles, parameter_matrix,
              particle_type_index_array, acc_x, acc_y, acc_z, num_types, boid_acc_x, boid_acc_y,
              boid_acc_z, boid_vel_x, boid_vel_y, boid_vel_z, boid_counts, sq_speed, bin_neighbours,
              particle_bins, particle_indices, bin_starts, bin_counts, blocks=None, threads=None,
              timestep=None, wrap_mode: int = _lib.WRAP_REFERENCE):
    """integrator.py:141-244: adaptive timestep, step1, accelerator, step2, boundary_condition.

    `sq_speed` may be None or a DeviceArray; its LAST element receives the maximum
    squared speed, as in the reference (integrator.py:183).  Returns the timestep used.
    """
    if timestep is None:
        m = max_sq_speed(vel_x, vel_y, vel_z, num_particles)
        value = float(m.copy_to_host()[0])
        if sq_speed is not None:
            sq_speed.tensor[-1] = value
        timestep = adaptive_timestep(value)
    step1(vel_x, vel_y, vel_z, acc_x, acc_y, acc_z, num_particles, timestep)
    accelerator(pos_x, pos_y, pos_z, vel_x, vel_y, vel_z, limits[0], limits[1], limits[2], r_max,
                num_particles, parameter_matrix, particle_type_index_array, acc_x, acc_y, acc_z,
                num_types, boid_acc_x, boid_acc_y, boid_acc_z, boid_vel_x, boid_vel_y, boid_vel_z,
                boid_counts, bin_neighbours, particle_bins, particle_indices, bin_starts, bin_counts,
                wrap_mode=wrap_mode)
    step2(pos_x, pos_y, pos_z, vel_x, vel_y, vel_z, acc_x, acc_y, acc_z, particle_type_index_array,
          parameter_matrix, num_particles, timestep)
    boundary_condition(pos_x, pos_y, pos_z, limits[0], limits[1], limits[2], num_particles)
    return timestep


def setup_bins(pos_x, pos_y, pos_z, num_bin_x, num_bin_y, num_bin_z, bin_size_x, bin_size_y, bin_size_z,
               num_bins, num_particles, particle_bins, particle_bin_counts, bin_offsets,
               particle_bin_starts, particle_indices, blocks=None, threads=None):
    """integrator.py:247-290: bin_particles + cumsum + set_indices in one enqueue."""
    g = _lib.Grid()
    g.num_bin_x, g.num_bin_y, g.num_bin_z = int(num_bin_x), int(num_bin_y), int(num_bin_z)
    g.num_cells = int(num_bin_x) * int(num_bin_y) * max(int(num_bin_z), 1)
    g.bin_size_x, g.bin_size_y, g.bin_size_z = float(bin_size_x), float(bin_size_y), float(bin_size_z)
    g.limit_x = float(bin_size_x) * int(num_bin_x)
    g.limit_y = float(bin_size_y) * int(num_bin_y)
    g.limit_z = float(bin_size_z) * int(num_bin_z)
    g.r_max = min(float(bin_size_x), float(bin_size_y))  # not used by the binning kernels
    setup_bins_device(pos_x, pos_y, pos_z, g, num_bins, num_particles, particle_bins, particle_bin_counts,
                      bin_offsets, particle_bin_starts, particle_indices)
