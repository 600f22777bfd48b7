"""Shared helpers of the GPU parity tests (the oracle is the checker here)."""
import numpy as np

from oracle import oracle as orc


def f32_state(arr):
    """Round a float64 state to the fp32 values the device holds, as float64 for the oracle."""
    return np.ascontiguousarray(np.asarray(arr, dtype=np.float32).astype(np.float64))


def random_state(n_per_type, limits, seed, kinds="clusters", speed=0.0):
    """A Simulation-like state without going through the constructor (explicit parameters)."""
    rng = np.random.default_rng(seed)
    counts = np.asarray(n_per_type)
    T = len(counts)
    n = int(counts.sum())
    lim = np.asarray(limits, dtype=np.float64)
    pos = rng.random((3, n)) * lim[:, None]
    if lim[2] == 0:
        pos[2] = 0.0
    vel = rng.normal(0.0, speed, (3, n)) if speed > 0 else np.zeros((3, n))
    if lim[2] == 0:
        vel[2] = 0.0
    acc = np.zeros((3, n))
    types = np.repeat(np.arange(T, dtype=np.int32), counts)
    pm = np.zeros((5, T, T))
    if kinds == "clusters":
        kind = np.zeros((T, T))
    elif kinds == "boids":
        kind = np.ones((T, T))
    else:  # mixed: first half clusters, second half boids, cross pairs symmetric random
        h = T // 2
        kind = np.zeros((T, T))
        kind[h:, h:] = 1
        for i in range(h):
            for j in range(h, T):
                kind[i, j] = kind[j, i] = float(rng.random() > 0.5)
    # parameter planes drawn like the reference's rand_param_matrix (state_parameters.py:5-36):
    # planes 0 and 1 come from ONE sorted pool, so every r_min lies below every r_max
    pool = rng.random(4 * T * T)
    pool[: 2 * T * T].sort()
    rng.shuffle(pool[: T * T])
    rng.shuffle(pool[T * T: 2 * T * T])
    u = pool.reshape(4, T, T)
    clus = kind == 0
    # scaling of Simulation.__init__ (simulation.py:70-96)
    pm[0] = np.where(clus, 5 + 3 * u[0], 12 + 2 * u[0])
    pm[1] = np.where(clus, 8 + 5 * u[1], 2 + 4 * u[1])
    pm[2] = np.where(clus, -10 + 3 * u[2], 2 + 2 * u[2])
    pm[3] = np.where(clus, -6 + 12 * u[3], -0.05 + 0.1 * u[3])
    pm[4] = kind
    r_max = float(np.max(pm[1]))
    return dict(pos=f32_state(pos), vel=f32_state(vel), acc=f32_state(acc), types=types,
                parameter_matrix=pm, limits=lim, r_max=r_max, n=n, T=T)


def oracle_sim(st, timestep=None, wrap_mode=orc.WRAP_REFERENCE, n_threads=1):
    return orc.OracleSimulation(st["pos"], st["vel"], st["acc"], st["types"], st["parameter_matrix"],
                                st["limits"], st["r_max"], timestep=timestep, wrap_mode=wrap_mode,
                                n_threads=n_threads)


def acc_errors(got, want):
    """Per-particle error of an acceleration field relative to the typical magnitude.

    got/want: (3, N).  Returns (max relative error, rms relative error) where each
    particle's error vector is measured against max(|a_i|, median |a|): the single-step
    bar of BASELINE.json is 1e-5 relative, and fp32 summation of ~100 signed terms
    cannot promise digits of a sum that cancels to far below its terms.
    """
    got = np.asarray(got, dtype=np.float64)
    want = np.asarray(want, dtype=np.float64)
    err = np.linalg.norm(got - want, axis=0)
    mag = np.linalg.norm(want, axis=0)
    floor = np.median(mag) if mag.size else 1.0
    rel = err / np.maximum(mag, max(floor, 1e-30))
    return float(rel.max(initial=0.0)), float(np.sqrt(np.mean(rel ** 2))) if rel.size else 0.0


def acc_error_unfloored(got, want):
    """max over particles of |got_i - want_i| / |want_i| with NO floor on the denominator (reported
    next to acc_errors: it is what BASELINE.json's "1e-5 relative" reads literally; a particle whose
    ~100 signed terms cancel to far below their size can exceed it without any term being wrong)."""
    got = np.asarray(got, dtype=np.float64)
    want = np.asarray(want, dtype=np.float64)
    err = np.linalg.norm(got - want, axis=0)
    mag = np.linalg.norm(want, axis=0)
    ok = mag > 0
    return float((err[ok] / mag[ok]).max(initial=0.0))
