"""ctypes/numpy front end of the C oracle (oracle/gof_oracle.c).

TEST INFRASTRUCTURE, NOT PRODUCT CODE: only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs may import this module.
Parity status: PINNED against golden vectors produced by the unmodified
reference under numba's CUDA simulator (tests/golden/make_golden.py).

Function names and argument order follow the reference kernels they restate
(main/acceleration_updater.py, main/integrator.py); launch-configuration
arguments (blocks, threads) do not exist here.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import build as _build

WRAP_REFERENCE = 0  # literal z wrap of acceleration_updater.py:325-328
WRAP_MINIMAGE = 1   # z treated like x and y

_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        path = _build.LIB
        if not os.path.exists(path) or os.path.getmtime(path) < os.path.getmtime(_build.SRC):
            _build.build()
        _lib = C.CDLL(path)
        _lib.gof_oracle_clusters.restype = C.c_double
        _lib.gof_oracle_clusters.argtypes = [C.c_double] * 5
        _lib.gof_oracle_timestep.restype = C.c_double
        _lib.gof_oracle_core_step.restype = C.c_double
        _lib.gof_oracle_max_threads.restype = C.c_int
    return _lib


def _p(a: np.ndarray, dtype) -> C.c_void_p:
    assert isinstance(a, np.ndarray) and a.dtype == dtype and a.flags.c_contiguous, (
        f"need C-contiguous {dtype}, got {getattr(a, 'dtype', type(a))}")
    return C.c_void_p(a.ctypes.data)


def _f64(a):
    return _p(a, np.float64)


def _f32(a):
    return _p(a, np.float32)


def _i32(a):
    return _p(a, np.int32)


def max_threads() -> int:
    return int(lib().gof_oracle_max_threads())


def clusters(dist, r_min, r_max, f_min, f_max) -> float:
    """acceleration_updater.py:10-20."""
    return float(lib().gof_oracle_clusters(dist, r_min, r_max, f_min, f_max))


def grid_from_limits(limits, r_max):
    """Bin geometry exactly as Simulation.__init__ derives it
    (main/simulation.py:103-139): floor(limits / r_max) bins per axis, the
    count array padded to the next power of two."""
    limits = np.asarray(limits, dtype=np.float64)
    nbx = int(np.floor(limits[0] / r_max))
    nby = int(np.floor(limits[1] / r_max))
    bsx = limits[0] / nbx
    bsy = limits[1] / nby
    if limits[2] != 0:
        nbz = int(np.floor(limits[2] / r_max))
        bsz = limits[2] / nbz
        cells = nbx * nby * nbz
        width = 14
    else:
        nbz = 0
        bsz = 0.0
        cells = nbx * nby
        width = 5
    num_bins = 1
    while num_bins < cells:
        num_bins *= 2
    return dict(num_bin_x=nbx, num_bin_y=nby, num_bin_z=nbz, bin_size_x=float(bsx),
                bin_size_y=float(bsy), bin_size_z=float(bsz), num_cells=cells,
                num_bins=num_bins, stencil=width)


def set_bin_neighbours(num_bin_x, num_bin_y, num_bin_z, bin_neighbours):
    """acceleration_updater.py:140-215."""
    lib().gof_oracle_set_bin_neighbours(int(num_bin_x), int(num_bin_y), int(num_bin_z),
                                        _i32(bin_neighbours))


def bin_particles(pos_x, pos_y, pos_z, num_bin_x, num_bin_y, num_bin_z, bin_size_x, bin_size_y,
                  bin_size_z, num_particles, particle_bins, particle_bin_counts, bin_offsets,
                  num_bins):
    """acceleration_updater.py:38-76."""
    lib().gof_oracle_bin_particles(
        _f64(pos_x), _f64(pos_y), _f64(pos_z), int(num_bin_x), int(num_bin_y), int(num_bin_z),
        C.c_double(bin_size_x), C.c_double(bin_size_y), C.c_double(bin_size_z),
        int(num_particles), _i32(particle_bins), _i32(particle_bin_counts), _i32(bin_offsets),
        int(num_bins))


def cumsum(values, result, d_max=None):
    """acceleration_updater.py:81-121 (d_max is the reference's log2 argument, unused)."""
    lib().gof_oracle_cumsum(_i32(values), _i32(result), int(values.size))


def set_indices(particle_bins, particle_bin_starts, bin_offsets, particle_indices,
                num_particles):
    """acceleration_updater.py:125-135."""
    lib().gof_oracle_set_indices(_i32(particle_bins), _i32(particle_bin_starts),
                                 _i32(bin_offsets), _i32(particle_indices), int(num_particles))


def setup_bins(pos_x, pos_y, pos_z, num_bin_x, num_bin_y, num_bin_z, bin_size_x, bin_size_y,
               bin_size_z, num_bins, num_particles, particle_bins, particle_bin_counts,
               bin_offsets, particle_bin_starts, particle_indices):
    """integrator.py:247-290."""
    bin_particles(pos_x, pos_y, pos_z, num_bin_x, num_bin_y, num_bin_z, bin_size_x, bin_size_y,
                  bin_size_z, num_particles, particle_bins, particle_bin_counts, bin_offsets,
                  num_bins)
    cumsum(particle_bin_counts, particle_bin_starts)
    set_indices(particle_bins, particle_bin_starts, bin_offsets, particle_indices, num_particles)


def accelerator(pos_x, pos_y, pos_z, vel_x, vel_y, vel_z, limitx, limity, limitz, r_max,
                num_particles, parameter_matrix, particle_type_index_array, acc_x, acc_y, acc_z,
                num_types, boid_acc_x, boid_acc_y, boid_acc_z, boid_vel_x, boid_vel_y,
                boid_vel_z, boid_counts, bin_neighbours, particle_bins, particle_indices,
                bin_starts, bin_counts, wrap_mode=WRAP_REFERENCE, n_threads=1):
    """acceleration_updater.py:219-458."""
    lib().gof_oracle_accelerator(
        _f64(pos_x), _f64(pos_y), _f64(pos_z), _f64(vel_x), _f64(vel_y), _f64(vel_z),
        C.c_double(limitx), C.c_double(limity), C.c_double(limitz), C.c_double(r_max),
        int(num_particles), _f64(parameter_matrix), _i32(particle_type_index_array),
        _f64(acc_x), _f64(acc_y), _f64(acc_z), int(num_types),
        _f32(boid_acc_x), _f32(boid_acc_y), _f32(boid_acc_z),
        _f32(boid_vel_x), _f32(boid_vel_y), _f32(boid_vel_z), _i32(boid_counts),
        _i32(bin_neighbours), _i32(particle_bins), _i32(particle_indices), _i32(bin_starts),
        _i32(bin_counts), int(wrap_mode), int(n_threads))


def timestep(vel_x, vel_y, vel_z, num_particles) -> float:
    """integrator.py:174-188 (the adaptive rule, host side in float64)."""
    return float(lib().gof_oracle_timestep(_f64(vel_x), _f64(vel_y), _f64(vel_z),
                                           int(num_particles)))


def step1(vel_x, vel_y, vel_z, acc_x, acc_y, acc_z, num_particles, timestep=0.02):
    """integrator.py:33-55."""
    lib().gof_oracle_step1(_f64(vel_x), _f64(vel_y), _f64(vel_z), _f64(acc_x), _f64(acc_y),
                           _f64(acc_z), int(num_particles), C.c_double(timestep))


def step2(pos_x, pos_y, pos_z, vel_x, vel_y, vel_z, acc_x, acc_y, acc_z, particle_tia,
          parameter_matrix, num_particles, timestep=0.02):
    """integrator.py:58-88."""
    num_types = parameter_matrix.shape[1]
    lib().gof_oracle_step2(_f64(pos_x), _f64(pos_y), _f64(pos_z), _f64(vel_x), _f64(vel_y),
                           _f64(vel_z), _f64(acc_x), _f64(acc_y), _f64(acc_z),
                           _i32(particle_tia), _f64(parameter_matrix), int(num_types),
                           int(num_particles), C.c_double(timestep))


def boundary_condition(pos_x, pos_y, pos_z, limitx, limity, limitz, num_particles):
    """integrator.py:91-114."""
    lib().gof_oracle_boundary_condition(_f64(pos_x), _f64(pos_y), _f64(pos_z),
                                        C.c_double(limitx), C.c_double(limity),
                                        C.c_double(limitz), int(num_particles))


class OracleSimulation:
    """The per-step state of main/simulation.py's Simulation (lines 35-190)
    held as float64 numpy arrays, advanced by the C oracle.  It takes the
    already-initialised state (positions, types, parameter matrix, r_max)
    so that it does not depend on any host code under test."""

    def __init__(self, pos, vel, acc, types, parameter_matrix, limits, r_max, timestep=None,
                 wrap_mode=WRAP_REFERENCE, n_threads=1):
        f = lambda a: np.ascontiguousarray(np.asarray(a, dtype=np.float64)).copy()
        self.pos_x, self.pos_y, self.pos_z = (f(pos[k]) for k in range(3))
        self.vel_x, self.vel_y, self.vel_z = (f(vel[k]) for k in range(3))
        self.acc_x, self.acc_y, self.acc_z = (f(acc[k]) for k in range(3))
        self.num_particles = int(self.pos_x.size)
        self.particle_type_index_array = np.ascontiguousarray(types, dtype=np.int32).copy()
        self.parameter_matrix = np.ascontiguousarray(parameter_matrix, dtype=np.float64).copy()
        self.num_types = int(self.parameter_matrix.shape[1])
        self.limits = np.asarray(limits, dtype=np.float64).copy()
        self.r_max = float(r_max)
        self.timestep = timestep
        self.wrap_mode = wrap_mode
        self.n_threads = n_threads
        g = grid_from_limits(self.limits, self.r_max)
        self.grid = g
        self.bin_neighbours = np.zeros((g["num_cells"], g["stencil"]), dtype=np.int32)
        set_bin_neighbours(g["num_bin_x"], g["num_bin_y"], g["num_bin_z"], self.bin_neighbours)
        n, t = self.num_particles, self.num_types
        self.particle_bin_counts = np.zeros(g["num_bins"], dtype=np.int32)
        self.particle_bin_starts = np.zeros(g["num_bins"], dtype=np.int32)
        self.particle_bins = np.zeros(n, dtype=np.int32)
        self.particle_indices = np.zeros(n, dtype=np.int32)
        self.bin_offsets = np.zeros(n, dtype=np.int32)
        self.boid = [np.zeros(n * t, dtype=np.float32) for _ in range(6)]
        self.boid_counts = np.zeros(n * t, dtype=np.int32)
        self.last_timestep = None

    def setup_bins(self):
        g = self.grid
        setup_bins(self.pos_x, self.pos_y, self.pos_z, g["num_bin_x"], g["num_bin_y"],
                   g["num_bin_z"], g["bin_size_x"], g["bin_size_y"], g["bin_size_z"],
                   g["num_bins"], self.num_particles, self.particle_bins,
                   self.particle_bin_counts, self.bin_offsets, self.particle_bin_starts,
                   self.particle_indices)

    def accelerate(self):
        """setup_bins + accelerator only (no integration): fills acc_*."""
        self.setup_bins()
        accelerator(self.pos_x, self.pos_y, self.pos_z, self.vel_x, self.vel_y, self.vel_z,
                    self.limits[0], self.limits[1], self.limits[2], self.r_max,
                    self.num_particles, self.parameter_matrix, self.particle_type_index_array,
                    self.acc_x, self.acc_y, self.acc_z, self.num_types, *self.boid,
                    self.boid_counts, self.bin_neighbours, self.particle_bins,
                    self.particle_indices, self.particle_bin_starts, self.particle_bin_counts,
                    wrap_mode=self.wrap_mode, n_threads=self.n_threads)

    def core_step(self):
        """simulation.py:193-247 in one C call."""
        g = self.grid
        ts = -1.0 if self.timestep is None else float(self.timestep)
        self.last_timestep = float(lib().gof_oracle_core_step(
            _f64(self.pos_x), _f64(self.pos_y), _f64(self.pos_z),
            _f64(self.vel_x), _f64(self.vel_y), _f64(self.vel_z),
            _f64(self.acc_x), _f64(self.acc_y), _f64(self.acc_z),
            C.c_double(self.limits[0]), C.c_double(self.limits[1]), C.c_double(self.limits[2]),
            C.c_double(self.r_max), int(self.num_particles), _f64(self.parameter_matrix),
            _i32(self.particle_type_index_array), int(self.num_types),
            int(g["num_bin_x"]), int(g["num_bin_y"]), int(g["num_bin_z"]),
            C.c_double(g["bin_size_x"]), C.c_double(g["bin_size_y"]), C.c_double(g["bin_size_z"]),
            int(g["num_bins"]), _i32(self.bin_neighbours), _i32(self.particle_bins),
            _i32(self.particle_bin_counts), _i32(self.bin_offsets), _i32(self.particle_bin_starts),
            _i32(self.particle_indices), *[_f32(b) for b in self.boid], _i32(self.boid_counts),
            C.c_double(ts), int(self.wrap_mode), int(self.n_threads)))
        return self.last_timestep
