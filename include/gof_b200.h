/*
 * gof_b200.h -- C ABI of the B200-native Game_of_Flies particle engine.
 *
 * This is the drop-in boundary for the reference's per-step hot path
 * (Simulation.core_step -> integrator.setup_bins -> integrator.integrate,
 * reference main/simulation.py:193-247).  Everything here is `extern "C"`,
 * plain pointers and sizes: no torch, numba or numpy types.  The host side
 * (game_of_flies_b200/*.py) binds it with ctypes; INTEGRATION.md shows the
 * stub a maintainer of the reference would add.
 *
 * Conventions
 *  - Every function returns 0 on success, a negative GOF_E_* code for an
 *    argument/state error, or a positive cudaError_t; gof_last_error()
 *    returns a message for the calling thread's last failure.
 *  - `stream` is a cudaStream_t passed as void* (NULL = legacy default
 *    stream).  All work is enqueued on it; nothing synchronises unless the
 *    function says so.
 *  - "SoA" arrays are fp32/int32 arrays in PARTICLE order (index i is the
 *    reference's particle i).  Pointers named d_* must be device pointers.
 *    Pointers named *_any may be host (pinned or pageable) or device
 *    pointers: they are only touched by cudaMemcpyAsync(cudaMemcpyDefault).
 *  - The library never allocates device memory: the caller (PyTorch on the
 *    Python side) owns every buffer and passes arenas/workspaces in.
 *  - There is no CPU fallback anywhere: without a CUDA device every compute
 *    entry point fails with the CUDA error.
 */
#ifndef GOF_B200_H
#define GOF_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GOF_ABI_VERSION 1

/* error codes (negative; positive values are cudaError_t) */
#define GOF_E_INVALID (-1)   /* bad argument */
#define GOF_E_STATE (-2)     /* call order (e.g. step before bind/upload) */
#define GOF_E_NOMEM (-3)     /* arena / workspace too small */
#define GOF_E_GRID (-4)      /* fewer than 3 bins along an axis (reference assumes r_max < L/3) */

/* periodic-wrap behaviour of the neighbour force kernel */
#define GOF_WRAP_REFERENCE 0 /* bug-for-bug: reference acceleration_updater.py:325-328 tests pos_y[i] in the z wrap */
#define GOF_WRAP_MINIMAGE 1  /* z handled like x and y */

#define GOF_MAX_TYPES 16

/* Bin geometry exactly as Simulation.__init__ derives it
 * (reference main/simulation.py:103-139).  num_bin_z == 0 selects 2D. */
typedef struct gof_grid {
    int32_t num_bin_x, num_bin_y, num_bin_z;
    int32_t num_cells;  /* num_bin_x * num_bin_y * max(num_bin_z, 1) */
    double bin_size_x, bin_size_y, bin_size_z; /* limits / num_bin, float64 like the reference */
    double limit_x, limit_y, limit_z;
    double r_max;
} gof_grid;

int gof_abi_version(void);
const char *gof_last_error(void);
/* Number of CUDA devices visible to the library (0 on a CPU-only host). */
int gof_device_count(void);
/* Kernels launched by this library since load / since the last reset (for bench.py's gpu_launches). */
uint64_t gof_launch_count(void);
void gof_reset_launch_count(void);

/* Fill a gof_grid from limits and r_max; returns GOF_E_GRID if an axis has < 3 bins
 * (mirrors simulation.py:103-139; host only, no CUDA). */
int gof_grid_init(gof_grid *grid, double limit_x, double limit_y, double limit_z, double r_max);

/* Host-only: the half-stencil neighbour table of acceleration_updater.py:140-215,
 * row-major [num_cells][5 (2D) or 14 (3D)].  Kept for callers that display or
 * test it; the CUDA force kernel derives the full stencil on the fly. */
int gof_set_bin_neighbours(int num_bin_x, int num_bin_y, int num_bin_z, int32_t *bin_neighbours);

/* ------------------------------------------------------------------------
 * Kernel-level entry points on SoA arrays in particle order.  One per
 * reference launch site, same argument meaning, fp32 state.
 * ---------------------------------------------------------------------- */

/* Scratch bytes needed by gof_setup_bins / gof_accelerator for n particles. */
size_t gof_workspace_bytes(int num_particles, int num_bins, int num_types);

/* integrator.setup_bins (reference main/integrator.py:247-290): bin_particles
 * + cumsum + set_indices.  num_bins is the padded length of the count/start
 * arrays (>= grid->num_cells; the reference pads to a power of two).
 * Outputs are deterministic: bin_offsets is the rank of the particle among
 * the particles of its bin in particle order (what the reference's atomic
 * yields when threads arrive in order), particle_indices is stable. */
int gof_setup_bins(const float *d_pos_x, const float *d_pos_y, const float *d_pos_z,
                   int num_particles, const gof_grid *grid, int num_bins,
                   int32_t *d_particle_bins, int32_t *d_particle_bin_counts,
                   int32_t *d_bin_offsets, int32_t *d_particle_bin_starts,
                   int32_t *d_particle_indices, void *d_workspace, size_t workspace_bytes,
                   void *stream);

/* acceleration_updater.accelerator (reference main/acceleration_updater.py:219-458).
 * parameter_matrix_host is the reference's float64 [5][T][T] matrix on the HOST.
 * d_boid_* may all be NULL; when given they receive the reference's per-type
 * neighbour sums, [T][N] indexed p*N + i (counts are exact, sums fp32). */
int gof_accelerator(const float *d_pos_x, const float *d_pos_y, const float *d_pos_z,
                    const float *d_vel_x, const float *d_vel_y, const float *d_vel_z,
                    int num_particles, const gof_grid *grid,
                    const double *parameter_matrix_host, int num_types,
                    const int32_t *d_particle_type_index_array,
                    float *d_acc_x, float *d_acc_y, float *d_acc_z,
                    float *d_boid_acc_x, float *d_boid_acc_y, float *d_boid_acc_z,
                    float *d_boid_vel_x, float *d_boid_vel_y, float *d_boid_vel_z,
                    int32_t *d_boid_counts,
                    const int32_t *d_particle_bins, const int32_t *d_particle_indices,
                    const int32_t *d_bin_starts, const int32_t *d_bin_counts,
                    int wrap_mode, void *d_workspace, size_t workspace_bytes, void *stream);

/* integrator.step1 (main/integrator.py:33-55). */
int gof_step1(float *d_vel_x, float *d_vel_y, float *d_vel_z, const float *d_acc_x,
              const float *d_acc_y, const float *d_acc_z, int num_particles, float timestep,
              void *stream);

/* integrator.step2 (main/integrator.py:58-88).  d_self_kind[t] is
 * parameter_matrix[-1, t, t] as int32 (0 = Clusters: apply the 0.95 damping). */
int gof_step2(float *d_pos_x, float *d_pos_y, float *d_pos_z, float *d_vel_x, float *d_vel_y,
              float *d_vel_z, const float *d_acc_x, const float *d_acc_y, const float *d_acc_z,
              const int32_t *d_particle_tia, const int32_t *d_self_kind, int num_particles,
              float timestep, void *stream);

/* integrator.boundary_condition (main/integrator.py:91-114). */
int gof_boundary_condition(float *d_pos_x, float *d_pos_y, float *d_pos_z, float limit_x,
                           float limit_y, float limit_z, int num_particles, void *stream);

/* set_sq_speed + max_reduction (main/integrator.py:20-30, 117-129): writes
 * max_i |v_i|^2 to *d_out (one float). */
int gof_max_sq_speed(const float *d_vel_x, const float *d_vel_y, const float *d_vel_z,
                     int num_particles, float *d_out, void *stream);

/* ------------------------------------------------------------------------
 * The resident engine: Simulation.core_step as one enqueue.  State lives on
 * the device as cell-sorted float4 arrays and never leaves between steps.
 * ---------------------------------------------------------------------- */
typedef struct gof_engine gof_engine;

/* parameter_matrix_host: float64 [5][T][T]; particle types are uploaded with the state. */
int gof_engine_create(gof_engine **out, int device, int num_particles, int num_types,
                      const gof_grid *grid, const double *parameter_matrix_host);
void gof_engine_destroy(gof_engine *e);

/* Device bytes the engine needs; the caller allocates them (256-byte aligned) and binds. */
size_t gof_engine_arena_bytes(const gof_engine *e);
int gof_engine_bind_arena(gof_engine *e, void *d_arena, size_t bytes);

/* Simulation.update_parameter_matrix (main/simulation.py:291-293): replace the whole matrix. */
int gof_engine_set_parameters(gof_engine *e, const double *parameter_matrix_host, void *stream);

/* Upload the state (cuda.to_device calls of main/simulation.py:161-173).  Any of
 * vel_* / acc_* may be NULL (zeros). */
int gof_engine_upload(gof_engine *e, const float *pos_x_any, const float *pos_y_any,
                      const float *pos_z_any, const float *vel_x_any, const float *vel_y_any,
                      const float *vel_z_any, const float *acc_x_any, const float *acc_y_any,
                      const float *acc_z_any, const int32_t *types_any, void *stream);

/* n_steps x Simulation.core_step (main/simulation.py:193-247).  timestep < 0 selects
 * the adaptive rule of integrator.py:174-188 (0.4 / max speed), evaluated on the device. */
int gof_engine_step(gof_engine *e, int n_steps, float timestep, int wrap_mode, void *stream);

/* Run setup_bins + accelerator on the resident state without integrating:
 * afterwards gof_engine_download's acc_* hold the accelerations of the current state. */
int gof_engine_accelerate(gof_engine *e, int wrap_mode, void *stream);

/* Copy the state back in particle order; any pointer may be NULL. */
int gof_engine_download(gof_engine *e, float *pos_x_any, float *pos_y_any, float *pos_z_any,
                        float *vel_x_any, float *vel_y_any, float *vel_z_any, float *acc_x_any,
                        float *acc_y_any, float *acc_z_any, void *stream);

/* One step of a HOST-driven loop: upload the state from nine host (or device) arrays, one
 * core_step, download the new state into nine arrays -- what a caller does who keeps the state
 * on the host between steps, as main/simulation.py does between `cuda.to_device`
 * (simulation.py:161-173) and `copy_to_host` (:295-304).  The uploads and the step run on
 * `stream`, the downloads on a stream of the engine's own, so that in a loop whose next input is
 * this call's output (two sets of host buffers used alternately) array k of the next call goes up
 * while array k + 1 of this call still comes down: both directions of the link are busy.  The
 * dependencies are tracked by host pointer: an input array that the previous call is still
 * writing waits for exactly that copy.  The call returns at once; gof_engine_host_wait blocks
 * until the outputs of the last call are complete.  pos/vel/acc xyz in the order of
 * gof_engine_upload; all nine are required both ways.  Not for slab engines. */
int gof_engine_step_host(gof_engine *e, const float *const *in9_any, const int32_t *types_any,
                         float *const *out9_any, float timestep, int wrap_mode, void *stream);
int gof_engine_host_wait(gof_engine *e);

/* Simulation.update's host copy (main/simulation.py:295-304): positions as [N][3]
 * fp32 in particle order.  xyz_any is typically pinned host memory; the copy is
 * asynchronous on `stream`. */
int gof_engine_snapshot(gof_engine *e, float *xyz_any, void *stream);

/* gof_engine_snapshot in two halves for double buffering: `stage` un-permutes into the engine's
 * device staging buffer (enqueue it on the compute stream, before the next step), `copy` moves
 * the staged frame out (enqueue it on a side stream that waited for the stage).  The next stage
 * must wait for the previous copy. */
int gof_engine_snapshot_stage(gof_engine *e, void *stream);
int gof_engine_snapshot_copy(gof_engine *e, float *xyz_any, void *stream);

/* The binning of the LAST step/upload in the reference's layout, particle order:
 * particle_bins[N], bin_counts[num_cells], bin_starts[num_cells], particle_indices[N].
 * Any pointer may be NULL. */
int gof_engine_read_bins(gof_engine *e, int32_t *particle_bins_any, int32_t *bin_counts_any,
                         int32_t *bin_starts_any, int32_t *particle_indices_any, void *stream);

/* gof_engine_step launches every step as one CUDA graph (captured on first use, per timestep
 * value and wrap mode); enable = 0 goes back to individual kernel launches.  Same results. */
int gof_engine_set_graph(gof_engine *e, int enable);

/* Small-system tuning knob of the force kernel (process-wide).  For Clusters-kind systems the
 * 9 (3 in 2D) stencil rows of a home particle can be walked by 3 or 9 warps instead of one
 * thread; -1 (default) picks by system size.  Results are bit-identical whatever the value. */
int gof_set_force_split(int split);

/* Per-kernel timing for bench.py's roofline: when enabled, gof_engine_step brackets every
 * force-kernel launch with CUDA events on `stream` (up to 4096 launches per enable).
 * gof_engine_force_time synchronises on the recorded events and returns the summed
 * device time in milliseconds and the number of launches it covers. */
int gof_engine_timing(gof_engine *e, int enable);
int gof_engine_force_time(gof_engine *e, float *total_ms, int *launches);

/* Diagnostics behind bench.py's roofline block (never part of a step): for the cell-ordered
 * snapshot the last step / accelerate read, out3 = {stencil candidates, force evaluations the
 * kernel spends after culling, pairs closer than r_max}, each summed over the particles (the
 * owned particles in slab mode; halo cells count as candidates).  Synchronises `stream`. */
int gof_engine_pair_stats(gof_engine *e, int wrap_mode, unsigned long long *out3, void *stream);

/* The timestep the last step used (synchronises `stream`). */
int gof_engine_last_timestep(gof_engine *e, float *out, void *stream);

/* ------------------------------------------------------------------------
 * Slab decomposition (multi-GPU).  The periodic box is cut along z into slabs of
 * whole bin layers; one engine (one GPU, one process) owns the layers
 * [layer_lo, layer_hi) plus one halo layer on either side.  The host drives three
 * phases per step and moves the buffers between neighbouring slabs with NCCL
 * send/recv (game_of_flies_b200/slab.py); every phase that produces sizes the host
 * needs synchronises `stream` and returns them.
 *
 *   1. gof_slab_phase_hash_async : bin resident particles, pack the ones that left the slab into the
 *                                  two send buffers (header word = record count); the buffers travel
 *                                  whole to the two neighbours, no size ever reaches the host
 *   2. gof_slab_phase_sort_async : append arrivals (counts read from the received headers), counting
 *                                  sort, step1 kick; leaves {live, bottom layer, top layer, error} in
 *                                  device words, which the host all-gathers and hands back with
 *                                  gof_slab_commit -> exchange boundary layers (positions, velocities
 *                                  if needed, per-cell counts)
 *   3. gof_slab_phase_force_interior (needs no halo, overlaps the exchange) +
 *      gof_slab_phase_force_boundary, or gof_slab_phase_force for everything at once
 *
 * With an adaptive timestep the maximum squared speed (pointer 7 or 8 of gof_slab_pointers,
 * chosen by gof_slab_parity) is all-reduced (MAX, as int32 bit patterns) between phases 1 and 2.
 * One host synchronisation per step (the all-gather of phase 2's words).  Results are
 * bit-identical to the single-GPU engine.
 * ---------------------------------------------------------------------- */
int gof_slab_create(gof_engine **out, int device, int capacity, int num_types, const gof_grid *grid,
                    const double *parameter_matrix_host, int layer_lo, int layer_hi,
                    int migrate_capacity, int halo_capacity);
/* Upload this slab's particles with their global ids. */
int gof_slab_upload(gof_engine *e, int num_particles, const float *pos_x_any, const float *pos_y_any,
                    const float *pos_z_any, const float *vel_x_any, const float *vel_y_any,
                    const float *vel_z_any, const float *acc_x_any, const float *acc_y_any,
                    const float *acc_z_any, const int32_t *types_any, const int32_t *ids_any, void *stream);
/* ptrs[0..10]: migration send down / send up / recv from below / recv from above (one float4
 * header + migrate_capacity records of three float4; always sent whole), snapshot positions, snapshot velocities (float4 each, cell order), cell counts
 * (int32), the two max-squared-speed slots (uint32 bit patterns), cell starts (int32), the 8 counter words (int32). */
int gof_slab_pointers(gof_engine *e, void **ptrs, int n_ptrs);
int gof_slab_phase_hash_async(gof_engine *e, void *stream);
int gof_slab_phase_sort_async(gof_engine *e, float timestep, void *stream);
/* words of pointer 10: [2] sticky error flag, [4..7] = live / bottom layer / top layer / error
 * after phase 2. */
int gof_slab_commit(gof_engine *e, int n_live, int lo_pop, int hi_pop, int error);
int gof_slab_phase_force(gof_engine *e, int n_halo_lo, int n_halo_hi, int wrap_mode, void *stream);
/* Phase 3 split for overlap: the interior layers need no halo and can be enqueued right after
 * gof_slab_phase_sort_async (their range is read from the device words), while sizes are gathered
 * and halo layers move on another stream; the two boundary layers follow once the halo is in. */
int gof_slab_phase_force_interior(gof_engine *e, int wrap_mode, void *stream);
int gof_slab_phase_force_boundary(gof_engine *e, int n_halo_lo, int n_halo_hi, int wrap_mode, void *stream);
int gof_slab_parity(const gof_engine *e);
int gof_slab_count(const gof_engine *e);
int gof_slab_download(gof_engine *e, float *pos_x_any, float *pos_y_any, float *pos_z_any,
                      float *vel_x_any, float *vel_y_any, float *vel_z_any, float *acc_x_any,
                      float *acc_y_any, float *acc_z_any, int32_t *ids_any, int32_t *types_any,
                      void *stream);


/* ------------------------------------------------------------------------
 * Slab decomposition over PEER MEMORY (the default transport on one node).  The slabs of a
 * job map each other's arenas (CUDA IPC between processes; plain pointers inside one) and a
 * step is enqueued without any host synchronisation or library collective:
 *   leavers and boundary layers are STORED into the neighbours' memory by kernels, tagged flag
 *   words announce them, one-block kernels on the consumer's stream wait for the flags, the
 *   maximum speed of the adaptive timestep goes to every slab the same way, and the sizes of a
 *   step (live particles, boundary layers, halo layers) never leave the device.
 * Results are bit-identical to the single-GPU engine and to the send/recv transport above.
 *
 *   gof_slab_window   12 words describing what peers write into: byte offsets inside the arena
 *                     of {records from below, records from above, cell-ordered positions,
 *                     velocities, cell counts, flag block}, then capacity, halo_capacity, local
 *                     layers, migrate_capacity, cells per layer, offset of the halo count rows
 *   gof_slab_connect  bases[t] = slab t's arena base as mapped in THIS process (this slab's own
 *                     base at index `rank`), infos = the `world` windows back to back.  All slabs
 *                     must have connected (and synchronised their devices) before any of them steps.
 *   gof_slab_step_p2p n_steps x core_step over this process's slabs (normally one); `side_stream`
 *                     carries the halo traffic and the boundary layers under the interior forces.
 *   gof_slab_check    synchronise `stream`, raise the slab's sticky device-side error if any
 *                     (capacity, buffer overflow, a neighbour that stopped answering for 4 s) and
 *                     return the live particle count.
 * ---------------------------------------------------------------------- */
int gof_slab_window(gof_engine *e, int64_t *info, int n_info);
int gof_slab_connect(gof_engine *e, int rank, int world, void *const *bases, const int64_t *infos);
int gof_slab_step_p2p(gof_engine *const *slabs, int n_slabs, int n_steps, float timestep, int wrap_mode,
                      void *main_stream, void *side_stream);
int gof_slab_check(gof_engine *e, int *n_live, void *stream);
/* Phase times of the peer-memory steps since gof_engine_timing(e, 1) (first slab of each call), summed
 * over *steps steps, in ms: {start -> sorted, sorted -> interior forces done, sorted -> boundary forces
 * done (side stream), start -> end}.  Synchronises on the recorded events. */
int gof_slab_phase_times(gof_engine *e, float *out4, int *steps);

/* CUDA IPC plumbing for gof_slab_connect between processes: the 64-byte handle of the device
 * allocation that holds d_ptr and d_ptr's offset inside it; open / close a peer's handle. */
int gof_ipc_export(const void *d_ptr, void *handle64, int64_t *offset);
int gof_ipc_open(const void *handle64, void **d_base);
int gof_ipc_close(void *d_base);

#ifdef __cplusplus
}
#endif
#endif /* GOF_B200_H */
