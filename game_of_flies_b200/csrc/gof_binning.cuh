// gof_binning.cuh -- cell hash, warp-aggregated counts, exclusive scan and the
// stable counting-sort scatter that reorders particles by cell.
//
// Replaces the reference's bin_particles / cumsum / set_indices
// (main/acceleration_updater.py:38-135).  The reference never moves particle
// data: it keeps an index permutation.  Here the permutation is applied to the
// float4 state every step so that the force kernel streams contiguous cells.
//
// Order inside a cell: by particle id (the reference's particle index).  That
// is what the reference's atomic hands out when threads arrive in index order
// and what its serial restatement (oracle/gof_oracle.c) produces, so
// particle_indices is bit-exact and deterministic.  The GPU gets there in two
// moves: an atomic (arrival-order) scatter of ids into the cell's segment, then
// a rank pass that counts, for every slot, the smaller ids in its segment.
#pragma once
#include "gof_kernels.cuh"

namespace gof {

constexpr int kBinThreads = 256;
constexpr int kScanThreads = 1024;
constexpr int kScanItems = 4;
constexpr int kScanTile = kScanThreads * kScanItems;

// Slab-mode send buffers for particles that left the owned layers (start-of-step
// migration; see game_of_flies_b200/slab.py).  A buffer is one float4 header whose first
// word is the record count, followed by records of three float4: {x, y, z, type},
// {vx, vy, vz, id}, {ax, ay, az, 0}.  Buffers always travel whole (fixed size), so neither
// side needs the count on the host.  All pointers are null on a single GPU.
struct MigrateOut {
    float4* down;    // header + records for the slab below
    float4* up;      // header + records for the slab above
    int* error;      // sticky error flag (1: moved more than a layer, 2: buffer overflow)
    int capacity;    // records per buffer
};

// Cell of every particle + per-cell population.  One atomic per distinct cell per
// warp (__match_any_sync groups lanes that hit the same cell; after the first
// step the arrays are already almost sorted, so a warp touches one or two cells).
//   SOA = true : reference-layout inputs x[], y[], z[] in particle order (gof_setup_bins)
//   SOA = false: the engine's float4 {x, y, z, type}; type < 0 marks a dead slot.
// `off` receives the arrival-order offset inside the cell (the reference's
// bin_offsets before it is made deterministic by the rank pass).
template <bool SOA>
__global__ void __launch_bounds__(kBinThreads)
k_hash_count(const float4* __restrict__ posw, const float4* __restrict__ veli,
             const float4* __restrict__ acc, const float* __restrict__ px,
             const float* __restrict__ py, const float* __restrict__ pz, int n,
             const int* __restrict__ n_limit, const __grid_constant__ Geometry g,
             int* __restrict__ cell_of, int* __restrict__ off, int* __restrict__ count,
             MigrateOut mig)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (n_limit) n = min(n, *n_limit);  // slab mode: the real count lives on the device
    int cell = -1;
    if (i < n) {
        float x, y, z;
        bool dead = false;
        if (SOA) {
            x = px[i]; y = py[i]; z = pz[i];
        } else {
            const float4 p = posw[i];
            x = p.x; y = p.y; z = p.z;
            dead = __float_as_int(p.w) < 0;
        }
        if (dead) {
            cell = g.num_cells;  // sentinel bin: sorts behind every live particle
        } else {
            const int cx = bin_coord(x, g.bsx, g.nbx);
            const int cy = bin_coord(y, g.bsy, g.nby);
            int lz = 0;
            if (g.is3d) {
                const int cz = bin_coord(z, g.bsz, g.nbz);
                lz = cz - g.layer_base;
                if (!g.index_wrap_z) {  // slab: local layer of a global layer, periodic
                    if (lz >= g.nbz) lz -= g.nbz;
                    if (lz < 0) lz += g.nbz;
                }
            }
            cell = cx + g.nbx * (cy + g.nby * lz);
            if (!SOA && !g.index_wrap_z && g.is3d) {
                // slab mode: a particle in a halo layer has left; hand it to the neighbour
                const bool down = (lz == 0), up = (lz == g.nlz - 1);
                if (down || up || lz >= g.nlz) {
                    if (lz >= g.nlz || mig.down == nullptr) {
                        // more than one layer away, or an arrival that is not ours: cannot happen
                        // while a step moves a particle by far less than a bin (speed is capped)
                        if (mig.error) atomicExch(mig.error, 1);
                    } else {
                        float4* buf = down ? mig.down : mig.up;
                        const int k = atomicAdd(reinterpret_cast<int*>(buf), 1);
                        if (k < mig.capacity) {
                            float4* rec = buf + 1 + 3 * (size_t)k;
                            rec[0] = posw[i];
                            rec[1] = veli[i];
                            rec[2] = acc[i];
                        } else {
                            atomicExch(mig.error, 2);
                        }
                    }
                    cell = g.num_cells;
                }
            }
        }
        cell_of[i] = cell;
    }
    // warp-aggregated atomic: lanes with the same cell elect a leader
    const unsigned active = __ballot_sync(0xffffffffu, cell >= 0);
    if (cell >= 0) {
        const unsigned peers = __match_any_sync(active, cell);
        const int leader = __ffs(peers) - 1;
        const int rank = __popc(peers & ((1u << lane_id()) - 1u));
        int base = 0;
        if (lane_id() == leader) base = atomicAdd(&count[cell], __popc(peers));
        base = __shfl_sync(peers, base, leader);
        off[i] = base + rank;
    }
}

// ---- exclusive scan of the cell counts (block-level scans, reduce-then-scan) ----

__device__ __forceinline__ int warp_inclusive_scan(int v)
{
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, v, d);
        if (lane_id() >= d) v += t;
    }
    return v;
}

// Block-wide exclusive scan of one int per thread (kScanThreads threads); returns the
// exclusive prefix and, through `total`, the block sum.
__device__ __forceinline__ int block_exclusive_scan(int v, int& total)
{
    __shared__ int warp_sums[32];
    const int inc = warp_inclusive_scan(v);
    const int w = threadIdx.x >> 5;
    if (lane_id() == 31) warp_sums[w] = inc;
    __syncthreads();
    if (w == 0) {
        const int nw = blockDim.x >> 5;
        int s = lane_id() < nw ? warp_sums[lane_id()] : 0;
        s = warp_inclusive_scan(s);
        warp_sums[lane_id()] = s;
    }
    __syncthreads();
    const int warp_prefix = w ? warp_sums[w - 1] : 0;
    total = warp_sums[31];
    __syncthreads();
    return warp_prefix + inc - v;
}

// Phase 1: sum of every tile of kScanTile counts.
__global__ void __launch_bounds__(kScanThreads)
k_scan_tile_sums(const int* __restrict__ in, int m, int* __restrict__ tile_sums)
{
    const int base = blockIdx.x * kScanTile + threadIdx.x * kScanItems;
    int s = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; ++k)
        if (base + k < m) s += in[base + k];
    int total;
    block_exclusive_scan(s, total);
    if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}

// Phase 2: one block turns the tile sums into exclusive tile prefixes (loops if needed).
__global__ void __launch_bounds__(kScanThreads)
k_scan_tile_prefix(int* __restrict__ tile_sums, int tiles)
{
    int carry = 0;
    for (int base = 0; base < tiles; base += kScanThreads) {
        const int idx = base + threadIdx.x;
        const int v = idx < tiles ? tile_sums[idx] : 0;
        int total;
        const int ex = block_exclusive_scan(v, total);
        if (idx < tiles) tile_sums[idx] = carry + ex;
        carry += total;
    }
}

// Phase 3: scan inside each tile and add the tile prefix.  With a single tile
// (up to 4096 cells: every configuration the reference UI can build) phases 1
// and 2 are skipped and this is the whole scan.
__global__ void __launch_bounds__(kScanThreads)
k_scan_apply(const int* __restrict__ in, int m, const int* __restrict__ tile_prefix,
             int* __restrict__ out)
{
    const int base = blockIdx.x * kScanTile + threadIdx.x * kScanItems;
    int v[kScanItems];
    int s = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
        v[k] = (base + k < m) ? in[base + k] : 0;
        s += v[k];
    }
    int total;
    int ex = block_exclusive_scan(s, total) + (tile_prefix ? tile_prefix[blockIdx.x] : 0);
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
        if (base + k < m) out[base + k] = ex;
        ex += v[k];
    }
}

// Arrival-order scatter of the ids into the cell segments.
// The engine also scatters the sort key of the order INSIDE a cell: the bit pattern of x (x >= 0,
// so the patterns order like the values); see k_rank_gather_kick.
// Sort key of the order INSIDE a cell: the bit pattern of x (x >= 0, so the patterns order like
// the values); ties are broken by id.  (Ordering a mixed system's cells by (type, x) -- so that the
// whole-warp walk changes its partner type a few times per cell instead of per candidate -- was
// measured and dropped: 32 consecutive particles then span two whole cells in x and the box of the
// warp culls nothing, 753 instead of ~610 evaluations per particle.)
__device__ __forceinline__ unsigned cell_order_key(const float4 p)
{
    return __float_as_uint(fabsf(p.x));
}

template <bool SOA>
__global__ void __launch_bounds__(kBinThreads)
k_scatter_ids(const float4* __restrict__ veli, int n, const int* __restrict__ n_limit,
              const int* __restrict__ cell_of, const int* __restrict__ off,
              const int* __restrict__ start, int* __restrict__ slot_id,
              const float4* __restrict__ posw, unsigned* __restrict__ slot_key,
              int* __restrict__ slab_words = nullptr, const int* __restrict__ slab_error = nullptr,
              int num_cells = 0, int per_layer = 0, int top = 0)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (slab_words && i == 0) {
        // slab mode, peer-memory transport: the scan is done, so the sizes of this step are known:
        // {live particles, population of the bottom owned layer, of the top owned layer, error flag}
        slab_words[0] = start[num_cells];
        slab_words[1] = start[2 * per_layer] - start[per_layer];
        slab_words[2] = start[(top + 1) * per_layer] - start[top * per_layer];
        slab_words[3] = *slab_error;
    }
    if (n_limit) n = min(n, *n_limit);
    if (i >= n) return;
    const int dst = start[cell_of[i]] + off[i];
    slot_id[dst] = SOA ? i : __float_as_int(veli[i].w);
    if (slot_key) slot_key[dst] = cell_order_key(posw[i]);
}

// Rank pass for the reference-layout API: final position of every particle in its
// cell = number of smaller ids in the segment.  Writes the reference's outputs.
__global__ void __launch_bounds__(kBinThreads)
k_rank_indices(int n, const int* __restrict__ cell_of, const int* __restrict__ start,
               const int* __restrict__ count, const int* __restrict__ slot_id,
               int* __restrict__ bin_offsets, int* __restrict__ particle_indices)
{
    const int id = blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= n) return;
    const int c = cell_of[id];
    const int b = start[c], e = b + count[c];
    int rank = 0;
    for (int t = b; t < e; ++t) rank += (slot_id[t] < id);
    bin_offsets[id] = rank;
    particle_indices[b + rank] = id;
}

// Rank pass of the engine, fused with the gather that physically reorders the
// state and with integrator.step1 (main/integrator.py:33-55: half kick with the
// previous acceleration, then the speed clamp).  step1 has to be finished for
// every particle before the force kernel runs because Boids alignment reads
// the neighbours' kicked velocities, and this pass rewrites the velocities anyway.
//   in : posw/veli/acc in last step's order (+ appended arrivals)
//   out: posw/veli in this step's cell order, cell_sorted
// `dt_in` < 0: read the adaptive timestep from scalars (see k_prepare_dt).
struct StepScalars {
    unsigned max_sq_speed_bits[2];  // ping-pong: fp32 bit pattern of max |v|^2 (non-negative => uint order)
    float dt;                       // timestep of the step being enqueued
    int error;                      // sticky device-side error flags
};

constexpr float kDummyPos = 1e18f;  // a padding entry of a candidate list sits here: (1e18)^2 * 3 still fits fp32, never in range

// One thread per SOURCE particle (rest order): its rank in its cell is its slot in the cell's
// segment.  The order inside a cell is by x, ties by id (`slot_key` given), or by id alone.  By x
// because the force kernel walks the candidates of a whole warp at once and culls them against
// the box of the warp's 32 home particles: 32 consecutive particles nearly always straddle two
// cells, and sorted by x they still span only about one cell width.  The order is a function of
// the positions and ids alone, so it is the same on every decomposition.  The reference's
// id-ordered `particle_indices` is re-derived on export (k_export_bins).  `pstart` is the table the segments are laid out with: the
// true starts, or the even-padded starts (`padded`), in which case the last particle of an odd
// cell also writes the dummy that fills the segment (the packed two-homes-per-lane force kernel
// needs the two particles of a lane to share a cell).
__global__ void __launch_bounds__(kBinThreads)
k_rank_gather_kick(int n, const int* __restrict__ n_limit, const int* __restrict__ cell_of,
                   const int* __restrict__ pstart,
                   const int* __restrict__ count, const int* __restrict__ slot_id,
                   const unsigned* __restrict__ slot_key,
                   int sentinel_cell, const float4* __restrict__ posw_in,
                   const float4* __restrict__ veli_in, const float4* __restrict__ acc_in,
                   float4* __restrict__ posw_out, float4* __restrict__ veli_out,
                   int* __restrict__ cell_sorted,
                   const StepScalars* __restrict__ scalars, int apply_kick)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (n_limit) n = min(n, *n_limit);
    if (i >= n) return;
    const int c = cell_of[i];
    if (c == sentinel_cell) return;  // dead slot: dropped by the sort
    float4 v = veli_in[i];
    const int id = __float_as_int(v.w);
    const int b = pstart[c], cnt = count[c], e = b + cnt;
    const float4 p = posw_in[i];
    int rank = 0;
    if (slot_key) {
        // branch-free, four entries in flight (the short-circuit form stalled on its branches)
        const unsigned key = cell_order_key(p);
        int t = b;
        for (; t + 4 <= e; t += 4) {
            const unsigned k0 = slot_key[t], k1 = slot_key[t + 1], k2 = slot_key[t + 2], k3 = slot_key[t + 3];
            const int i0 = slot_id[t], i1 = slot_id[t + 1], i2 = slot_id[t + 2], i3 = slot_id[t + 3];
            rank += (int)((k0 < key) | ((k0 == key) & (i0 < id))) + (int)((k1 < key) | ((k1 == key) & (i1 < id))) +
                    (int)((k2 < key) | ((k2 == key) & (i2 < id))) + (int)((k3 < key) | ((k3 == key) & (i3 < id)));
        }
        for (; t < e; ++t) {
            const unsigned kt = slot_key[t];
            rank += (int)((kt < key) | ((kt == key) & (slot_id[t] < id)));
        }
    } else {
        for (int t = b; t < e; ++t) rank += (slot_id[t] < id);
    }
    const int dst = b + rank;

    if (apply_kick) {
        const float4 a = acc_in[i];
        const float h = 0.5f * scalars->dt;
        v.x = fmaf(a.x, h, v.x);
        v.y = fmaf(a.y, h, v.y);
        v.z = fmaf(a.z, h, v.z);
        const float sq = v.x * v.x + v.y * v.y + v.z * v.z;
        if (sq > kMaxSpeed * kMaxSpeed) {
            const float k = kMaxSpeed / sqrtf(sq);
            v.x *= k; v.y *= k; v.z *= k;
        }
    }
    posw_out[dst] = p;
    veli_out[dst] = v;
    cell_sorted[dst] = c;
}

}  // namespace gof

namespace gof {

// Slab mode: unpack the migration records received from the two neighbours behind the
// resident particles (they are binned with the rest right after).  The counts are the
// header words of the received buffers; words[0] = arrivals, words[1] = slots in use.
__global__ void __launch_bounds__(kBinThreads)
k_append_arrivals(const float4* __restrict__ from_down, const float4* __restrict__ from_up, int capacity,
                  int base, float4* __restrict__ posw, float4* __restrict__ veli,
                  float4* __restrict__ acc, int* __restrict__ words)
{
    const int n_down = min(__float_as_int(from_down[0].x), capacity);
    const int n_up = min(__float_as_int(from_up[0].x), capacity);
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k == 0) {
        words[0] = n_down + n_up;
        words[1] = base + n_down + n_up;
    }
    if (k >= n_down + n_up) return;
    const float4* rec = k < n_down ? from_down + 1 + 3 * (size_t)k : from_up + 1 + 3 * (size_t)(k - n_down);
    posw[base + k] = rec[0];
    veli[base + k] = rec[1];
    acc[base + k] = rec[2];
}

// Slab mode: first index of every cell of the two halo layers.  The halo particles were
// received behind the owned ones (lower halo at `base_lo`, upper at `base_hi`), cell by
// cell in the sender's order, together with the per-cell counts.  One block per layer.
__global__ void __launch_bounds__(kScanThreads)
k_halo_starts(const int* __restrict__ count, int* __restrict__ start, int cells_per_layer,
              int first_cell_hi, int base_lo, int base_hi)
{
    const int first = blockIdx.x ? first_cell_hi : 0;
    int carry = blockIdx.x ? base_hi : base_lo;
    for (int b = 0; b < cells_per_layer; b += kScanThreads) {
        const int idx = b + threadIdx.x;
        const int v = idx < cells_per_layer ? count[first + idx] : 0;
        int total;
        const int ex = block_exclusive_scan(v, total);
        if (idx < cells_per_layer) start[first + idx] = carry + ex;
        carry += total;
    }
}

}  // namespace gof
