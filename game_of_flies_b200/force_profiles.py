"""Host-side force profiles (mirror of reference main/force_profiles.py).

The reference keeps numba-jitted scalar copies of the Clusters profile for
plotting and for its jit-vs-list benchmark; the simulation itself evaluates
the profile inside the CUDA force kernel (csrc/gof_force.cuh).  These are the
same functions, vectorised over numpy arrays, for UI code and tests.
"""
from __future__ import annotations

import numpy as np


def _clusters(dist, r_min, r_max, f_min, f_max, clamp: bool):
    d = np.asarray(dist, dtype=np.float64)
    inner = -f_min / r_min * d + f_min
    slope = 2 * f_max / (r_max - r_min)
    rising = slope * (d - r_min)
    falling = slope * (r_max - d)
    out = np.where(d < r_min, inner, np.where(d < (r_min + r_max) / 2, rising, falling))
    if clamp:
        out = np.where((d >= 0) & (d < r_max), out, 0.0)
    return out if out.ndim else float(out)


def wrap_clusters_force_distance(*args):
    """force_profiles.py:80-103: closure over (r_min, r_max, f_min, f_max), zero outside [0, r_max)."""
    r_min, r_max, f_min, f_max = args

    def clusters_force_distance(dist):
        return _clusters(dist, r_min, r_max, f_min, f_max, clamp=True)

    return clusters_force_distance


def all_force_functions(profile_name: str, *params):
    """force_profiles.py:106-130: matrix of profile functions, one per species pair."""
    if profile_name != "cluster_distance_input":
        return []
    return [[wrap_clusters_force_distance(*pair) for pair in row] for row in params]


def general_force_function(profile_type: int, input_vect, args):
    """force_profiles.py:133-154."""
    if profile_type == 0:
        return _clusters(input_vect[0], args[0], args[1], args[2], args[3], clamp=True)
    raise NotImplementedError("Position Input not implemented yet")


def clusters_device(dist, args):
    """What the force kernel evaluates (acceleration_updater.py:10-20): no clamp past r_max."""
    return _clusters(dist, args[0], args[1], args[2], args[3], clamp=False)
