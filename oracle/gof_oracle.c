/*
 * gof_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A plain-C, float64 restatement of the Game_of_Flies per-step hot path
 * (binning -> neighbour forces -> integrate), written to be the CHECKER for
 * the sm_100a CUDA path in game_of_flies_b200/csrc and the CPU baseline leg
 * of bench.py.  Nothing under game_of_flies_b200/ may import, link or call
 * this file; only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs do.
 *
 * Parity status: PINNED.  tests/golden/make_golden.py runs the unmodified
 * reference kernels (main/acceleration_updater.py, main/integrator.py) under
 * numba's CUDA simulator in single-block launches and stores their inputs
 * and outputs under tests/golden/ (.npz files); tests/test_oracle_golden.py checks
 * every function below against those vectors.
 *
 * Each function cites the reference code it restates (path:line relative to
 * the reference checkout).  Where the reference kernel is racy across CUDA
 * blocks (it zeroes accumulators and then relies on a block-level
 * syncthreads as if it were a grid barrier) this file implements the
 * single-block semantics, i.e. "zero everything, then accumulate", which is
 * also what the reference computes whenever the grid is one block.
 *
 * Arithmetic: float64 everywhere the reference uses float64 (positions,
 * velocities, accelerations, parameter matrix), float32 for the boid
 * accumulators (main/simulation.py:151-159 allocates them as float32) and
 * int32 for every index array.  Compile with -ffp-contract=off so that no
 * FMA contraction changes the rounding relative to numba's float64 code.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define GOF_ORACLE_WRAP_REFERENCE 0 /* bug-for-bug z wrap (acceleration_updater.py:325-328) */
#define GOF_ORACLE_WRAP_MINIMAGE 1  /* the evident intent: z treated like x and y */

/* Python/numba float floor division  a // b  (numba lowers it with the same
 * divmod algorithm as CPython's float_divmod / numpy's npy_divmod). */
static double py_floordiv(double a, double b)
{
    double mod = fmod(a, b);
    double div = (a - mod) / b;
    if (mod != 0.0) {
        if ((b < 0.0) != (mod < 0.0)) {
            mod += b;
            div -= 1.0;
        }
    }
    if (div != 0.0) {
        double fl = floor(div);
        if (div - fl > 0.5)
            fl += 1.0;
        return fl;
    }
    return copysign(0.0, a / b);
}

int gof_oracle_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* acceleration_updater.py:38-76  bin_particles.
 * Serial semantics of the atomic: offsets are handed out in particle order. */
void gof_oracle_bin_particles(const double *pos_x, const double *pos_y, const double *pos_z,
                              int num_bin_x, int num_bin_y, int num_bin_z,
                              double bin_size_x, double bin_size_y, double bin_size_z,
                              int num_particles, int32_t *particle_bins,
                              int32_t *particle_bin_counts, int32_t *bin_offsets, int num_bins)
{
    (void)num_bin_z;
    for (int b = 0; b < num_bins; ++b)
        particle_bin_counts[b] = 0;
    if (bin_size_z > 1) {
        for (int i = 0; i < num_particles; ++i) {
            double v = py_floordiv(pos_x[i], bin_size_x)
                     + num_bin_x * py_floordiv(pos_y[i], bin_size_y)
                     + (double)(num_bin_x * num_bin_y) * py_floordiv(pos_z[i], bin_size_z);
            particle_bins[i] = (int32_t)v;
            bin_offsets[i] = particle_bin_counts[particle_bins[i]]++;
        }
    } else {
        for (int i = 0; i < num_particles; ++i) {
            double v = py_floordiv(pos_x[i], bin_size_x)
                     + num_bin_x * py_floordiv(pos_y[i], bin_size_y);
            particle_bins[i] = (int32_t)v;
            bin_offsets[i] = particle_bin_counts[particle_bins[i]]++;
        }
    }
}

/* acceleration_updater.py:81-121  cumsum: a Blelloch exclusive scan over the
 * (power-of-two padded) bin counts.  Result: result[k] = sum(values[:k]). */
void gof_oracle_cumsum(const int32_t *values, int32_t *result, int n)
{
    int32_t run = 0;
    for (int k = 0; k < n; ++k) {
        result[k] = run;
        run += values[k];
    }
}

/* acceleration_updater.py:125-135  set_indices. */
void gof_oracle_set_indices(const int32_t *particle_bins, const int32_t *particle_bin_starts,
                            const int32_t *bin_offsets, int32_t *particle_indices,
                            int num_particles)
{
    for (int i = 0; i < num_particles; ++i)
        particle_indices[particle_bin_starts[particle_bins[i]] + bin_offsets[i]] = i;
}

/* acceleration_updater.py:140-215  set_bin_neighbours: the half stencil
 * (5 bins in 2D, 14 in 3D) with periodic wrap of the bin coordinates.
 * bin_neighbours is row-major [num_cells][5 or 14]. */
void gof_oracle_set_bin_neighbours(int num_bin_x, int num_bin_y, int num_bin_z,
                                   int32_t *bin_neighbours)
{
    const int width = (num_bin_z < 1) ? 5 : 14;
    for (int i = 0; i < num_bin_x; ++i) {
        int ip1 = i + 1, im1 = i - 1;
        if (i == 0)
            im1 = num_bin_x - 1;
        else if (i == num_bin_x - 1)
            ip1 = 0;
        for (int j = 0; j < num_bin_y; ++j) {
            int jp1 = j + 1, jm1 = j - 1;
            if (j == 0)
                jm1 = num_bin_y - 1;
            else if (j == num_bin_y - 1)
                jp1 = 0;
            if (num_bin_z < 1) {
                int32_t *row = bin_neighbours + (size_t)(i + j * num_bin_x) * width;
                row[0] = i + j * num_bin_x;
                row[1] = ip1 + j * num_bin_x;
                row[2] = im1 + jp1 * num_bin_x;
                row[3] = i + jp1 * num_bin_x;
                row[4] = ip1 + jp1 * num_bin_x;
            } else {
                const int nxy = num_bin_x * num_bin_y;
                for (int k = 0; k < num_bin_z; ++k) {
                    int kp1 = k + 1;
                    if (k == num_bin_z - 1)
                        kp1 = 0;
                    int32_t *row = bin_neighbours + (size_t)(i + j * num_bin_x + k * nxy) * width;
                    row[0] = i + j * num_bin_x + k * nxy;
                    row[1] = ip1 + j * num_bin_x + k * nxy;
                    row[2] = im1 + jp1 * num_bin_x + k * nxy;
                    row[3] = i + jp1 * num_bin_x + k * nxy;
                    row[4] = ip1 + jp1 * num_bin_x + k * nxy;
                    row[5] = ip1 + j * num_bin_x + kp1 * nxy;
                    row[6] = im1 + j * num_bin_x + kp1 * nxy;
                    row[7] = i + jp1 * num_bin_x + kp1 * nxy;
                    row[8] = i + jm1 * num_bin_x + kp1 * nxy;
                    row[9] = ip1 + jp1 * num_bin_x + kp1 * nxy;
                    row[10] = ip1 + jm1 * num_bin_x + kp1 * nxy;
                    row[11] = im1 + jp1 * num_bin_x + kp1 * nxy;
                    row[12] = im1 + jm1 * num_bin_x + kp1 * nxy;
                    row[13] = i + j * num_bin_x + kp1 * nxy;
                }
            }
        }
    }
}

/* acceleration_updater.py:10-20  clusters(dist, args): r_min, r_max, f_min, f_max.
 * No clamp beyond r_max: the device function lets the third branch go
 * negative past args[1]. */
static double clusters_force(double dist, const double *pm, int stride, int p1, int p2, int T)
{
    const double a0 = pm[0 * stride + p1 * T + p2];
    const double a1 = pm[1 * stride + p1 * T + p2];
    const double a2 = pm[2 * stride + p1 * T + p2];
    const double a3 = pm[3 * stride + p1 * T + p2];
    if (dist < a0)
        return -a2 / a0 * dist + a2;
    else if (dist < (a0 + a1) / 2)
        return 2 * a3 / (a1 - a0) * (dist - a0);
    else
        return 2 * a3 / (a1 - a0) * (a1 - dist);
}

/* Host-visible copy of the device function, for the force_profiles tests. */
double gof_oracle_clusters(double dist, double r_min, double r_max, double f_min, double f_max)
{
    const double pm[4] = {r_min, r_max, f_min, f_max};
    return clusters_force(dist, pm, 1, 0, 0, 1);
}

static inline void atomic_add_f64(double *p, double v)
{
#pragma omp atomic
    *p += v;
}
static inline void atomic_add_f32(float *p, float v)
{
#pragma omp atomic
    *p += v;
}
static inline void atomic_add_i32(int32_t *p, int32_t v)
{
#pragma omp atomic
    *p += v;
}

/* acceleration_updater.py:219-458  accelerator.
 * parameter_matrix is [5][T][T] float64; boid_* are [T][N] float32 (index
 * p*N + i); bin_neighbours is [cells][5|14].  wrap_mode selects the literal
 * z-wrap of line 327 (which tests pos_y[i]) or the min-image intent.
 * n_threads <= 1 runs the serial, deterministic order used by the checker;
 * > 1 runs the particle loop under OpenMP with atomics (the numba-parallel
 * style CPU baseline). */
void gof_oracle_accelerator(const double *pos_x, const double *pos_y, const double *pos_z,
                            const double *vel_x, const double *vel_y, const double *vel_z,
                            double limitx, double limity, double limitz, double r_max,
                            int num_particles, const double *parameter_matrix,
                            const int32_t *particle_type_index_array,
                            double *acc_x, double *acc_y, double *acc_z, int num_types,
                            float *boid_acc_x, float *boid_acc_y, float *boid_acc_z,
                            float *boid_vel_x, float *boid_vel_y, float *boid_vel_z,
                            int32_t *boid_counts, const int32_t *bin_neighbours,
                            const int32_t *particle_bins, const int32_t *particle_indices,
                            const int32_t *bin_starts, const int32_t *bin_counts,
                            int wrap_mode, int n_threads)
{
    const int N = num_particles;
    const int T = num_types;
    const int TT = T * T;
    const double *pm = parameter_matrix;
    const double *kind = pm + 4 * TT; /* parameter_matrix[-1] */
    const int num_neighbours = (limitz == 0) ? 5 : 14;
    if (n_threads < 1)
        n_threads = 1;

    /* lines 253-264: zero the accumulators */
    memset(acc_x, 0, sizeof(double) * N);
    memset(acc_y, 0, sizeof(double) * N);
    memset(acc_z, 0, sizeof(double) * N);
    memset(boid_acc_x, 0, sizeof(float) * (size_t)N * T);
    memset(boid_acc_y, 0, sizeof(float) * (size_t)N * T);
    memset(boid_acc_z, 0, sizeof(float) * (size_t)N * T);
    memset(boid_vel_x, 0, sizeof(float) * (size_t)N * T);
    memset(boid_vel_y, 0, sizeof(float) * (size_t)N * T);
    memset(boid_vel_z, 0, sizeof(float) * (size_t)N * T);
    memset(boid_counts, 0, sizeof(int32_t) * (size_t)N * T);

#pragma omp parallel for schedule(dynamic, 64) num_threads(n_threads) if (n_threads > 1)
    for (int i = 0; i < N; ++i) {
        const int p1 = particle_type_index_array[i];
        const int my_bin = particle_bins[i];
        for (int b = 0; b < num_neighbours; ++b) {
            const int bin2 = bin_neighbours[(size_t)my_bin * num_neighbours + b];
            const int cnt = bin_counts[bin2];
            for (int p = 0; p < cnt; ++p) {
                const int j = particle_indices[bin_starts[bin2] + p];
                if (i == j)
                    continue;
                const int p2 = particle_type_index_array[j];
                const double xi = pos_x[i], yi = pos_y[i], zi = pos_z[i];
                double pos_x_1 = xi, pos_y_1 = yi, pos_z_1 = zi;
                double pos_x_2 = pos_x[j], pos_y_2 = pos_y[j], pos_z_2 = pos_z[j];

                /* lines 304-315: particle i as seen from j */
                if (pos_x_2 < r_max && pos_x_1 > limitx - r_max)
                    pos_x_1 -= limitx;
                else if (pos_x_1 < r_max && pos_x_2 > limitx - r_max)
                    pos_x_1 += limitx;
                if (pos_y_2 < r_max && pos_y_1 > limity - r_max)
                    pos_y_1 -= limity;
                else if (pos_y_1 < r_max && pos_y_2 > limity - r_max)
                    pos_y_1 += limity;
                if (pos_z_2 < r_max && pos_z_1 > limitz - r_max)
                    pos_z_1 -= limitz;
                else if (pos_z_1 < r_max && pos_z_2 > limitz - r_max)
                    pos_z_1 += limitz;

                /* lines 317-328: particle j as seen from i */
                if (xi < r_max && pos_x_2 > limitx - r_max)
                    pos_x_2 -= limitx;
                else if (pos_x_2 < r_max && xi > limitx - r_max)
                    pos_x_2 += limitx;
                if (yi < r_max && pos_y_2 > limity - r_max)
                    pos_y_2 -= limity;
                else if (pos_y_2 < r_max && yi > limity - r_max)
                    pos_y_2 += limity;
                if (zi < r_max && pos_z_2 > limitz - r_max)
                    pos_z_2 -= limitz;
                else if (pos_z_2 < r_max &&
                         ((wrap_mode == GOF_ORACLE_WRAP_REFERENCE) ? yi : zi) > limitz - r_max)
                    pos_z_2 += limitz; /* line 327 tests pos_y[i] */

                const double dx = xi - pos_x_2, dy = yi - pos_y_2, dz = zi - pos_z_2;
                double dist = dx * dx + dy * dy + dz * dz;
                if (!((1e-10 < dist) && (dist < r_max * r_max)))
                    continue;
                dist = sqrt(dist);

                const double k12 = kind[p1 * T + p2];
                if (k12 == 0) {
                    const double acc = clusters_force(dist, pm, TT, p1, p2, T);
                    atomic_add_f64(&acc_x[i], -acc * dx / dist);
                    atomic_add_f64(&acc_y[i], -acc * dy / dist);
                    atomic_add_f64(&acc_z[i], -acc * dz / dist);
                } else if (k12 == 1) {
                    /* boids(): separation only, inside args[0] / 5 */
                    const double a0 = pm[0 * TT + p1 * T + p2];
                    const double a1 = pm[1 * TT + p1 * T + p2];
                    if (dist < a0 / 5) {
                        atomic_add_f64(&acc_x[i], -(a1 * (pos_x_2 - xi) / dist));
                        atomic_add_f64(&acc_y[i], -(a1 * (pos_y_2 - yi) / dist));
                        atomic_add_f64(&acc_z[i], -(a1 * (pos_z_2 - zi) / dist));
                    }
                    const size_t s = (size_t)p2 * N + i;
                    atomic_add_i32(&boid_counts[s], 1);
                    atomic_add_f32(&boid_acc_x[s], (float)pos_x_2);
                    atomic_add_f32(&boid_acc_y[s], (float)pos_y_2);
                    atomic_add_f32(&boid_acc_z[s], (float)pos_z_2);
                    atomic_add_f32(&boid_vel_x[s], (float)vel_x[j]);
                    atomic_add_f32(&boid_vel_y[s], (float)vel_y[j]);
                    atomic_add_f32(&boid_vel_z[s], (float)vel_z[j]);
                }

                if (bin2 != my_bin) {
                    /* lines 377-420: Newton's-third-law style update of j */
                    const double k21 = kind[p2 * T + p1];
                    if (k21 == 0) {
                        const double acc = clusters_force(dist, pm, TT, p2, p1, T);
                        atomic_add_f64(&acc_x[j], acc * dx / dist);
                        atomic_add_f64(&acc_y[j], acc * dy / dist);
                        atomic_add_f64(&acc_z[j], acc * dz / dist);
                    } else if (k21 == 1) {
                        const double a0 = pm[0 * TT + p2 * T + p1];
                        const double a1 = pm[1 * TT + p2 * T + p1];
                        if (dist < a0 / 5) {
                            atomic_add_f64(&acc_x[j], -(a1 * (-pos_x_2 + xi) / dist));
                            atomic_add_f64(&acc_y[j], -(a1 * (-pos_y_2 + yi) / dist));
                            atomic_add_f64(&acc_z[j], -(a1 * (-pos_z_2 + zi) / dist));
                        }
                        const size_t s = (size_t)p1 * N + j;
                        atomic_add_i32(&boid_counts[s], 1);
                        atomic_add_f32(&boid_acc_x[s], (float)pos_x_1);
                        atomic_add_f32(&boid_acc_y[s], (float)pos_y_1);
                        atomic_add_f32(&boid_acc_z[s], (float)pos_z_1);
                        atomic_add_f32(&boid_vel_x[s], (float)vel_x[i]);
                        atomic_add_f32(&boid_vel_y[s], (float)vel_y[i]);
                        atomic_add_f32(&boid_vel_z[s], (float)vel_z[i]);
                    }
                }
            }
        }
    }

    /* lines 422-458: cohesion and alignment from the per-type means */
#pragma omp parallel for schedule(static) num_threads(n_threads) if (n_threads > 1)
    for (int i = 0; i < N; ++i) {
        const int p1 = particle_type_index_array[i];
        for (int j = 0; j < T; ++j) {
            const size_t s = (size_t)j * N + i;
            const int32_t c = boid_counts[s];
            if (kind[p1 * T + j] == 1 && c > 0) {
                const double cohesion = pm[3 * TT + p1 * T + j];
                const double alignment = pm[2 * TT + p1 * T + j];
                const double ax = (double)boid_vel_x[s] / c;
                const double ay = (double)boid_vel_y[s] / c;
                const double az = (double)boid_vel_z[s] / c;
                double spd = sqrt(vel_x[i] * vel_x[i] + vel_y[i] * vel_y[i] + vel_z[i] * vel_z[i]);
                if (spd > 10)
                    spd = (ax * vel_x[i] + ay * vel_y[i] + az * vel_z[i]) / spd;
                else
                    spd = 0.0;
                acc_x[i] += cohesion * (((double)boid_acc_x[s] / c) - pos_x[i])
                          + alignment * (ax - spd * vel_x[i]);
                acc_y[i] += cohesion * (((double)boid_acc_y[s] / c) - pos_y[i])
                          + alignment * (ay - spd * vel_y[i]);
                acc_z[i] += cohesion * (((double)boid_acc_z[s] / c) - pos_z[i])
                          + alignment * (az - spd * vel_z[i]);
            }
        }
    }
}

/* integrator.py:117-129 + 174-188: adaptive timestep from the maximum speed
 * (the reference reduces squared speeds on the device and finishes on the
 * host in float64). */
double gof_oracle_timestep(const double *vel_x, const double *vel_y, const double *vel_z,
                           int num_particles)
{
    double m = 0.0;
    for (int i = 0; i < num_particles; ++i) {
        double s = vel_x[i] * vel_x[i] + vel_y[i] * vel_y[i] + vel_z[i] * vel_z[i];
        if (s > m)
            m = s;
    }
    double t = sqrt(m);
    if (t > 1e-6)
        return 0.4 / t;
    return 0.1;
}

/* integrator.py:33-55  step1: half kick, then clamp the speed to max_spd = 20. */
void gof_oracle_step1(double *vel_x, double *vel_y, double *vel_z, const double *acc_x,
                      const double *acc_y, const double *acc_z, int num_particles,
                      double timestep)
{
    const double max_spd = 20;
    for (int i = 0; i < num_particles; ++i) {
        vel_x[i] += acc_x[i] * 0.5 * timestep;
        vel_y[i] += acc_y[i] * 0.5 * timestep;
        vel_z[i] += acc_z[i] * 0.5 * timestep;
        double spd = vel_x[i] * vel_x[i] + vel_y[i] * vel_y[i] + vel_z[i] * vel_z[i];
        if (spd > max_spd * max_spd) {
            spd = sqrt(spd);
            vel_x[i] *= max_spd / spd;
            vel_y[i] *= max_spd / spd;
            vel_z[i] *= max_spd / spd;
        }
    }
}

/* integrator.py:58-88  step2: drift with the new acceleration, second half
 * kick, and the 0.95 damping for particles whose self-interaction is of the
 * Clusters kind. */
void gof_oracle_step2(double *pos_x, double *pos_y, double *pos_z, double *vel_x, double *vel_y,
                      double *vel_z, const double *acc_x, const double *acc_y,
                      const double *acc_z, const int32_t *particle_tia,
                      const double *parameter_matrix, int num_types, int num_particles,
                      double timestep)
{
    const double fac = 0.95;
    const int T = num_types;
    const double *kind = parameter_matrix + 4 * T * T;
    for (int i = 0; i < num_particles; ++i) {
        pos_x[i] += (vel_x[i] + 0.5 * acc_x[i] * timestep) * timestep;
        pos_y[i] += (vel_y[i] + 0.5 * acc_y[i] * timestep) * timestep;
        pos_z[i] += (vel_z[i] + 0.5 * acc_z[i] * timestep) * timestep;
        vel_x[i] += acc_x[i] * 0.5 * timestep;
        vel_y[i] += acc_y[i] * 0.5 * timestep;
        vel_z[i] += acc_z[i] * 0.5 * timestep;
        const int p = particle_tia[i];
        if (kind[p * T + p] == 0) {
            vel_x[i] *= fac;
            vel_y[i] *= fac;
            vel_z[i] *= fac;
        }
    }
}

/* integrator.py:91-114  boundary_condition: one periodic image, strict
 * comparisons (a particle exactly on `limit` stays there). */
void gof_oracle_boundary_condition(double *pos_x, double *pos_y, double *pos_z, double limitx,
                                   double limity, double limitz, int num_particles)
{
    for (int i = 0; i < num_particles; ++i) {
        if (pos_x[i] < 0)
            pos_x[i] += limitx;
        else if (pos_x[i] > limitx)
            pos_x[i] -= limitx;
        if (pos_y[i] < 0)
            pos_y[i] += limity;
        else if (pos_y[i] > limity)
            pos_y[i] -= limity;
        if (pos_z[i] < 0)
            pos_z[i] += limitz;
        else if (pos_z[i] > limitz)
            pos_z[i] -= limitz;
    }
}

/* One whole Simulation.core_step (simulation.py:193-247): setup_bins
 * (integrator.py:247-290) followed by integrate (integrator.py:141-244).
 * timestep < 0 selects the adaptive rule.  Scratch arrays are caller
 * provided so that the timed CPU baseline does not measure malloc.
 * Returns the timestep that was used. */
double gof_oracle_core_step(double *pos_x, double *pos_y, double *pos_z, double *vel_x,
                            double *vel_y, double *vel_z, double *acc_x, double *acc_y,
                            double *acc_z, double limitx, double limity, double limitz,
                            double r_max, int num_particles, const double *parameter_matrix,
                            const int32_t *particle_tia, int num_types, int num_bin_x,
                            int num_bin_y, int num_bin_z, double bin_size_x, double bin_size_y,
                            double bin_size_z, int num_bins, const int32_t *bin_neighbours,
                            int32_t *particle_bins, int32_t *bin_counts, int32_t *bin_offsets,
                            int32_t *bin_starts, int32_t *particle_indices, float *boid_acc_x,
                            float *boid_acc_y, float *boid_acc_z, float *boid_vel_x,
                            float *boid_vel_y, float *boid_vel_z, int32_t *boid_counts,
                            double timestep, int wrap_mode, int n_threads)
{
    gof_oracle_bin_particles(pos_x, pos_y, pos_z, num_bin_x, num_bin_y, num_bin_z, bin_size_x,
                             bin_size_y, bin_size_z, num_particles, particle_bins, bin_counts,
                             bin_offsets, num_bins);
    gof_oracle_cumsum(bin_counts, bin_starts, num_bins);
    gof_oracle_set_indices(particle_bins, bin_starts, bin_offsets, particle_indices,
                           num_particles);
    if (timestep < 0)
        timestep = gof_oracle_timestep(vel_x, vel_y, vel_z, num_particles);
    gof_oracle_step1(vel_x, vel_y, vel_z, acc_x, acc_y, acc_z, num_particles, timestep);
    gof_oracle_accelerator(pos_x, pos_y, pos_z, vel_x, vel_y, vel_z, limitx, limity, limitz,
                           r_max, num_particles, parameter_matrix, particle_tia, acc_x, acc_y,
                           acc_z, num_types, boid_acc_x, boid_acc_y, boid_acc_z, boid_vel_x,
                           boid_vel_y, boid_vel_z, boid_counts, bin_neighbours, particle_bins,
                           particle_indices, bin_starts, bin_counts, wrap_mode, n_threads);
    gof_oracle_step2(pos_x, pos_y, pos_z, vel_x, vel_y, vel_z, acc_x, acc_y, acc_z, particle_tia,
                     parameter_matrix, num_types, num_particles, timestep);
    gof_oracle_boundary_condition(pos_x, pos_y, pos_z, limitx, limity, limitz, num_particles);
    return timestep;
}
