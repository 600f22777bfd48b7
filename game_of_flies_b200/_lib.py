"""ctypes binding of the C ABI declared in include/gof_b200.h.

The shared library is built in-tree by game_of_flies_b200/build.py (nvcc,
sm_100a).  There is no fallback: if the library is missing the import of any
compute module fails with instructions, and if it is present but no CUDA
device is, every compute entry point raises GofError with the CUDA message.
"""
from __future__ import annotations

import ctypes as C
import os

from . import build as _build

GOF_ABI_VERSION = 1
WRAP_REFERENCE = 0
WRAP_MINIMAGE = 1
MAX_TYPES = 16

E_INVALID, E_STATE, E_NOMEM, E_GRID = -1, -2, -3, -4


class GofError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"gof_b200 error {code}: {message}")
        self.code = code


class Grid(C.Structure):
    """struct gof_grid (include/gof_b200.h)."""
    _fields_ = [
        ("num_bin_x", C.c_int32), ("num_bin_y", C.c_int32), ("num_bin_z", C.c_int32),
        ("num_cells", C.c_int32),
        ("bin_size_x", C.c_double), ("bin_size_y", C.c_double), ("bin_size_z", C.c_double),
        ("limit_x", C.c_double), ("limit_y", C.c_double), ("limit_z", C.c_double),
        ("r_max", C.c_double),
    ]


_vp = C.c_void_p
_f = C.c_float
_i = C.c_int
_sz = C.c_size_t

# name -> (restype, argtypes); every symbol include/gof_b200.h declares
SIGNATURES = {
    "gof_abi_version": (_i, []),
    "gof_last_error": (C.c_char_p, []),
    "gof_device_count": (_i, []),
    "gof_launch_count": (C.c_uint64, []),
    "gof_reset_launch_count": (None, []),
    "gof_grid_init": (_i, [C.POINTER(Grid), C.c_double, C.c_double, C.c_double, C.c_double]),
    "gof_set_bin_neighbours": (_i, [_i, _i, _i, _vp]),
    "gof_workspace_bytes": (_sz, [_i, _i, _i]),
    "gof_setup_bins": (_i, [_vp, _vp, _vp, _i, C.POINTER(Grid), _i, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "gof_accelerator": (_i, [_vp] * 6 + [_i, C.POINTER(Grid), _vp, _i, _vp] + [_vp] * 3 + [_vp] * 7
                        + [_vp] * 4 + [_i, _vp, _sz, _vp]),
    "gof_step1": (_i, [_vp] * 6 + [_i, _f, _vp]),
    "gof_step2": (_i, [_vp] * 9 + [_vp, _vp, _i, _f, _vp]),
    "gof_boundary_condition": (_i, [_vp] * 3 + [_f, _f, _f, _i, _vp]),
    "gof_max_sq_speed": (_i, [_vp] * 3 + [_i, _vp, _vp]),
    "gof_engine_create": (_i, [C.POINTER(_vp), _i, _i, _i, C.POINTER(Grid), _vp]),
    "gof_engine_destroy": (None, [_vp]),
    "gof_engine_arena_bytes": (_sz, [_vp]),
    "gof_engine_bind_arena": (_i, [_vp, _vp, _sz]),
    "gof_engine_set_parameters": (_i, [_vp, _vp, _vp]),
    "gof_engine_upload": (_i, [_vp] + [_vp] * 10 + [_vp]),
    "gof_engine_step": (_i, [_vp, _i, _f, _i, _vp]),
    "gof_engine_accelerate": (_i, [_vp, _i, _vp]),
    "gof_engine_download": (_i, [_vp] + [_vp] * 9 + [_vp]),
    "gof_engine_step_host": (_i, [_vp, _vp, _vp, _vp, _f, _i, _vp]),
    "gof_engine_host_wait": (_i, [_vp]),
    "gof_engine_snapshot": (_i, [_vp, _vp, _vp]),
    "gof_engine_snapshot_stage": (_i, [_vp, _vp]),
    "gof_engine_snapshot_copy": (_i, [_vp, _vp, _vp]),
    "gof_engine_read_bins": (_i, [_vp] + [_vp] * 4 + [_vp]),
    "gof_set_force_split": (_i, [_i]),
    "gof_engine_set_graph": (_i, [_vp, _i]),
    "gof_engine_timing": (_i, [_vp, _i]),
    "gof_engine_force_time": (_i, [_vp, C.POINTER(_f), C.POINTER(_i)]),
    "gof_slab_window": (_i, [_vp, _vp, _i]),
    "gof_slab_connect": (_i, [_vp, _i, _i, _vp, _vp]),
    "gof_slab_step_p2p": (_i, [_vp, _i, _i, C.c_float, _i, _vp, _vp]),
    "gof_slab_check": (_i, [_vp, _vp, _vp]),
    "gof_slab_phase_times": (_i, [_vp, _vp, _vp]),
    "gof_ipc_export": (_i, [_vp, _vp, _vp]),
    "gof_ipc_open": (_i, [_vp, _vp]),
    "gof_ipc_close": (_i, [_vp]),
    "gof_engine_pair_stats": (_i, [_vp, _i, _vp, _vp]),
    "gof_engine_last_timestep": (_i, [_vp, C.POINTER(_f), _vp]),
    "gof_slab_create": (_i, [C.POINTER(_vp), _i, _i, _i, C.POINTER(Grid), _vp, _i, _i, _i, _i]),
    "gof_slab_upload": (_i, [_vp, _i] + [_vp] * 11 + [_vp]),
    "gof_slab_pointers": (_i, [_vp, C.POINTER(_vp), _i]),
    "gof_slab_phase_hash_async": (_i, [_vp, _vp]),
    "gof_slab_phase_sort_async": (_i, [_vp, _f, _vp]),
    "gof_slab_commit": (_i, [_vp, _i, _i, _i, _i]),
    "gof_slab_phase_force": (_i, [_vp, _i, _i, _i, _vp]),
    "gof_slab_phase_force_interior": (_i, [_vp, _i, _vp]),
    "gof_slab_phase_force_boundary": (_i, [_vp, _i, _i, _i, _vp]),
    "gof_slab_parity": (_i, [_vp]),
    "gof_slab_count": (_i, [_vp]),
    "gof_slab_download": (_i, [_vp] + [_vp] * 11 + [_vp]),
}

_lib = None


def library_path() -> str:
    """The in-tree library; GOF_B200_LIB selects another build of it (kernel tuning sweeps)."""
    return os.environ.get("GOF_B200_LIB") or _build.LIB


def load() -> C.CDLL:
    """Load libgof_b200.so (building it first if nvcc is here and it is stale)."""
    global _lib
    if _lib is not None:
        return _lib
    path = library_path()
    if path == _build.LIB and _build.stale():
        try:
            _build.build()
        except FileNotFoundError as exc:  # no nvcc on this host
            if not os.path.exists(path):
                raise ImportError(
                    f"{path} is missing and could not be built ({exc}). Run "
                    "`python game_of_flies_b200/build.py` where nvcc is available; "
                    "this package has no CPU fallback.") from exc
            import warnings
            warnings.warn(f"{path} is OLDER than its sources and nvcc is not available to rebuild it: "
                          "loading the stale library", RuntimeWarning, stacklevel=2)
        except Exception as exc:  # nvcc is here and the sources do not compile: never run old kernels silently
            raise ImportError(f"{path} is stale and rebuilding it failed: {exc}") from exc
    lib = C.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here = header/library mismatch
        fn.restype = res
        fn.argtypes = args
    if lib.gof_abi_version() != GOF_ABI_VERSION:
        raise ImportError(f"{path}: ABI version {lib.gof_abi_version()} != {GOF_ABI_VERSION}")
    _lib = lib
    return lib


def check(rc: int) -> None:
    if rc != 0:
        raise GofError(rc, load().gof_last_error().decode("utf-8", "replace"))


def make_grid(limits, r_max) -> Grid:
    """gof_grid_init: the bin geometry of reference main/simulation.py:103-139."""
    g = Grid()
    check(load().gof_grid_init(C.byref(g), float(limits[0]), float(limits[1]), float(limits[2]),
                               float(r_max)))
    return g
