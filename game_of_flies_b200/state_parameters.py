"""Initial state and random species parameters.

Host-side mirror of reference main/state_parameters.py: `initialise` (lines
39-96) and `rand_param_matrix` (lines 5-36) with the same names, arguments,
returned dictionary keys and -- because the viewer, the recorder and users'
seeds depend on it -- the same sequence of draws from numpy's global generator,
so a seed gives the same simulation as in the reference.
"""
from __future__ import annotations

import numpy as np


def rand_param_matrix(n_type, force_type: str = "Clusters", max_param: int = 4, seed=None):
    """(max_param + 1, T, T) matrix: four parameter planes plus the interaction kind.

    Planes 0 and 1 (r_min, r_max) come from one sorted pool so that every
    r_min is below every r_max; each plane is then shuffled on its own.
    Plane 2 (f_min) is negated.  Returns the matrix and max of plane 3.
    """
    t = len(n_type)
    tt = t * t
    if seed is not None:
        np.random.seed(seed)
    pool = np.random.rand(tt * max_param)
    pool[: 2 * tt].sort()
    np.random.shuffle(pool[:tt])
    np.random.shuffle(pool[tt: 2 * tt])
    matrix = np.zeros((max_param + 1, t, t))
    matrix[:max_param] = pool.reshape(max_param, t, t)
    matrix[2] *= -1.0
    # plane -1 stays 0 ("Clusters"); Simulation assigns the kinds afterwards
    return matrix, np.max(matrix[3])


def initialise(n_type, seed=None, limits: tuple = (100, 100, 0)):
    """Uniform random positions in the box, zero velocity and acceleration."""
    counts = np.asarray(n_type)
    total = int(np.sum(counts))
    if seed is not None:
        np.random.seed(seed)
    u = np.random.rand(3, total)
    pos = np.empty((3, total))
    for axis in range(3):
        pos[axis] = limits[axis] * u[axis]
    vel = np.zeros((3, total))
    acc = np.zeros((3, total))
    # species index per particle: counts[0] zeros, then counts[1] ones, ...
    type_index = np.repeat(np.arange(len(counts), dtype=np.float64), counts.astype(np.int64))
    param_matrix, max_rmax = rand_param_matrix(n_type, seed=seed)
    return {
        "pos_x": pos[0], "pos_y": pos[1], "pos_z": pos[2],
        "vel_x": vel[0], "vel_y": vel[1], "vel_z": vel[2],
        "acc_x": acc[0], "acc_y": acc[1], "acc_z": acc[2],
        "limits": limits,
        "n_type_array": n_type,
        "particle_type_indx_array": type_index,
        "max_particle_per_type": 10,
        "parameter_matrix": param_matrix,
        "max_rmax": max_rmax,
    }


if __name__ == "__main__":
    print(initialise([10, 6, 7]))
