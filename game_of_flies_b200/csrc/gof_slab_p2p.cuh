// gof_slab_p2p.cuh -- the slab decomposition's exchange steps as kernels that STORE INTO THE
// NEIGHBOUR'S MEMORY over NVLink (peer mappings of the neighbours' arenas), with stream-ordered
// waits on tagged flag words instead of host synchronisation or library collectives.
//
// Per step a slab
//   1. packs the particles that left its layers (k_hash_count) and pushes the two record
//      buffers into the neighbours' receive buffers                         (k_push_migration)
//   2. waits for the two buffers addressed to it, appends and bins them      (k_arrivals_p2p)
//   3. sorts; pushes its bottom / top layer -- contiguous slices of the sorted arrays -- and
//      their per-cell counts into the neighbours' halo regions              (k_push_halo)
//   4. waits for its two halo layers, then runs the boundary layers' forces (k_wait_tags)
//   5. publishes its maximum squared speed to EVERY slab                    (k_publish_speed)
// and the next step's arrivals kernel waits for all of them                 (k_arrivals_p2p).
//
// A flag is one 64-bit word {step tag << 32 | payload} written with a single store after a
// system-scope fence: a reader that sees the tag sees the data.  Buffers are single (the speed
// slots double, by step parity): the neighbour handshakes of a step order every overwrite behind
// the last read of the previous step (DESIGN.md section 5 walks through the chains).  A wait gives
// up after ~4 s and raises the slab's sticky error flag, so a dead peer ends a run with an error
// instead of a hang.
//
// The reference has no multi-GPU path; this serves BASELINE.json's slab configurations.
#pragma once
#include "gof_binning.cuh"

namespace gof {

constexpr int kMaxWorld = 16;          // slabs a job can have (speed slots per parity)
constexpr int kPushThreads = 512;
constexpr int kHaloPushBlocks = 24;    // per direction
constexpr int kErrPeerTimeout = 3;     // sticky error: a neighbour's flag did not arrive
constexpr int kErrHaloOverflow = 4;    // sticky error: a boundary layer larger than the neighbour's halo region

// What a slab writes into ONE neighbour (device pointers valid in this process: the neighbour's
// arena mapped with CUDA IPC, or plain pointers when both slabs live in one process).
struct PeerSide {
    float4* mig_recv;                  // its receive buffer for records coming from this slab
    float4* halo_posw;                 // its halo region for the layer this slab hands over
    float4* halo_veli;
    int* halo_count;                   // its count row of that halo layer
    unsigned long long* mig_flag;      // {tag, -}
    unsigned long long* halo_flag;     // {tag, particles}
    int halo_capacity;
};

__device__ __forceinline__ unsigned long long make_flag(unsigned tag, unsigned payload)
{
    return ((unsigned long long)tag << 32) | payload;
}

__device__ __forceinline__ void store_flag(unsigned long long* p, unsigned long long v)
{
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

__device__ __forceinline__ unsigned long long load_flag(const unsigned long long* p)
{
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ unsigned long long global_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

// Spin until the flag carries `tag`; false after ~4 s.
__device__ __forceinline__ bool wait_flag(const unsigned long long* p, unsigned tag, unsigned long long* out)
{
    unsigned long long v = load_flag(p);
    if ((unsigned)(v >> 32) != tag) {
        const unsigned long long t0 = global_ns();
        unsigned backoff = 32;
        for (;;) {
            v = load_flag(p);
            if ((unsigned)(v >> 32) == tag) break;
            __nanosleep(backoff);
            if (backoff < 1024) backoff <<= 1;
            if (global_ns() - t0 > 4000000000ull) return false;
        }
    }
    *out = v;
    return true;
}

// Two blocks: block 0 pushes the buffer for the slab below, block 1 the one for the slab above
// (header word = record count, then the records: three float4 each).
__global__ void __launch_bounds__(kPushThreads)
k_push_migration(const float4* __restrict__ send_down, const float4* __restrict__ send_up,
                 const PeerSide below, const PeerSide above, int capacity, unsigned tag)
{
    const float4* src = blockIdx.x == 0 ? send_down : send_up;
    const PeerSide& to = blockIdx.x == 0 ? below : above;
    const int n = min(__float_as_int(src[0].x), capacity);
    const int words = 1 + 3 * n;
    for (int k = threadIdx.x; k < words; k += blockDim.x) to.mig_recv[k] = src[k];
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) store_flag(to.mig_flag, make_flag(tag, (unsigned)n));
}

// One thread per flag this slab waits for (up to four).  `payload_out[t]` receives the flag's low word.
__global__ void k_wait_tags(const unsigned long long* f0, const unsigned long long* f1, unsigned tag,
                            int* payload_out, int* error)
{
    const unsigned long long* f = threadIdx.x == 0 ? f0 : f1;
    if (threadIdx.x > 1 || f == nullptr) return;
    unsigned long long v = 0;
    if (!wait_flag(f, tag, &v)) {
        atomicExch(error, kErrPeerTimeout);
        return;
    }
    const int payload = (int)(unsigned)(v & 0xffffffffu);
    if (payload_out) payload_out[threadIdx.x] = payload;
    if (payload < 0) atomicExch(error, kErrHaloOverflow);  // the sender could not fit its layer
}

// The bottom owned layer goes to the slab below (as its upper halo), the top owned layer to the
// slab above (as its lower halo): slices [0, lo) and [live - hi, live) of the sorted arrays, plus
// the count rows of local layers 1 and nlz - 2.  words = {live, lo, hi} (k_slab_counts).
// Blocks [0, kHaloPushBlocks) serve the slab below, the rest the slab above; the last block of a
// direction to finish publishes the flag.  `done` = two zeroed ints (left zero again).
__global__ void __launch_bounds__(kPushThreads)
k_push_halo(const float4* __restrict__ posw, const float4* __restrict__ veli, const int* __restrict__ count,
            const int* __restrict__ words, int per_layer, int top_layer, int with_velocities,
            const PeerSide below, const PeerSide above, unsigned tag, int* __restrict__ done)
{
    const int dir = blockIdx.x / kHaloPushBlocks, b = blockIdx.x - dir * kHaloPushBlocks;
    const PeerSide& to = dir == 0 ? below : above;
    const int live = words[0], lo = words[1], hi = words[2];
    const int n = dir == 0 ? lo : hi;
    const int first = dir == 0 ? 0 : live - hi;
    const int layer = dir == 0 ? 1 : top_layer;
    const bool fits = n <= to.halo_capacity;
    if (fits) {
        const int stride = kHaloPushBlocks * blockDim.x;
        for (int k = b * blockDim.x + threadIdx.x; k < n; k += stride) {
            to.halo_posw[k] = posw[first + k];
            if (with_velocities) to.halo_veli[k] = veli[first + k];
        }
        for (int k = b * blockDim.x + threadIdx.x; k < per_layer; k += stride)
            to.halo_count[k] = count[layer * per_layer + k];
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const int prev = atomicAdd(done + dir, 1);
        if (prev == kHaloPushBlocks - 1) {
            done[dir] = 0;
            __threadfence_system();
            store_flag(to.halo_flag, make_flag(tag, fits ? (unsigned)n : 0xffffffffu));
        }
    }
}

// After a step's forces: this slab's maximum squared speed goes to every slab's slot array of the
// parity that describes the new velocities (slots[t] points into slab t's memory).
__global__ void k_publish_speed(const StepScalars* __restrict__ s, int parity,
                                unsigned long long* const* __restrict__ slots, int world, unsigned tag)
{
    const int t = threadIdx.x;
    if (t < world) store_flag(slots[t], make_flag(tag, s->max_sq_speed_bits[parity]));
}

// Phase 2's front end in ONE launch: every block waits for the two record buffers addressed to this
// slab, block 0 also settles the timestep (it waits for every slab's published maximum speed: the
// one global dependency of a step), then thread k appends arrival k behind the resident
// particles and bins it.  Also clears the headers of this
// slab's send buffers for the next step: their push is behind us on the stream.
//   words: [0] arrivals, [1] slots in use, [2] sticky error, [4] live particles before this step
__global__ void __launch_bounds__(kBinThreads)
k_arrivals_p2p(const unsigned long long* __restrict__ f_below, const unsigned long long* __restrict__ f_above,
               unsigned tag, const float4* __restrict__ from_down, const float4* __restrict__ from_up,
               int mig_capacity, float4* __restrict__ posw, float4* __restrict__ veli, float4* __restrict__ acc,
               int* __restrict__ words, int slot_capacity, const __grid_constant__ Geometry g,
               int* __restrict__ cell_of, int* __restrict__ off, int* __restrict__ count, StepScalars* s,
               int parity, float fixed_dt, const unsigned long long* __restrict__ speed_in, int world,
               float4* __restrict__ send_down, float4* __restrict__ send_up)
{
    if (threadIdx.x < 2) {
        unsigned long long v = 0;
        if (!wait_flag(threadIdx.x ? f_above : f_below, tag, &v)) atomicExch(words + 2, kErrPeerTimeout);
    }
    if (blockIdx.x == 0 && threadIdx.x >= 32 && threadIdx.x < 64) {
        // the timestep of this step: the maximum over every slab's published word
        const int ln = (int)threadIdx.x - 32;
        unsigned bits = 0u;
        bool ok = true;
        if (fixed_dt < 0.0f && ln < world) {
            unsigned long long v = 0;
            ok = wait_flag(speed_in + ln, tag, &v);
            bits = (unsigned)(v & 0xffffffffu);
        }
        bits = __reduce_max_sync(0xffffffffu, bits);
        ok = __all_sync(0xffffffffu, ok);
        if (ln == 0) {
            float dt = fixed_dt;
            if (fixed_dt < 0.0f) {
                if (!ok) atomicExch(words + 2, kErrPeerTimeout);
                const double m = sqrt((double)__uint_as_float(bits));
                dt = (m > 1e-6) ? (float)(0.4 / m) : 0.1f;
                s->max_sq_speed_bits[parity] = bits;
            }
            s->dt = dt;
            s->max_sq_speed_bits[parity ^ 1] = 0u;
        }
    }
    __syncthreads();  // the flags have been seen: the buffers are complete
    const int base = words[4];
    int n_down = min(__float_as_int(from_down[0].x), mig_capacity);
    int n_up = min(__float_as_int(from_up[0].x), mig_capacity);
    if (base + n_down + n_up > slot_capacity) {
        if (blockIdx.x == 0 && threadIdx.x == 0) atomicExch(words + 2, 5);
        n_down = n_up = 0;
    }
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k == 0) {
        words[0] = n_down + n_up;
        words[1] = base + n_down + n_up;
        send_down[0] = make_float4(0.f, 0.f, 0.f, 0.f);  // (record counters of the next step's leavers)
        send_up[0] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    if (k >= n_down + n_up) return;
    const float4* rec = k < n_down ? from_down + 1 + 3 * (size_t)k : from_up + 1 + 3 * (size_t)(k - n_down);
    const float4 p = rec[0];
    posw[base + k] = p;
    veli[base + k] = rec[1];
    acc[base + k] = rec[2];
    // its cell in the local grid; an arrival that is not in an owned layer cannot happen
    const int cx = bin_coord(p.x, g.bsx, g.nbx), cy = bin_coord(p.y, g.bsy, g.nby);
    int lz = bin_coord(p.z, g.bsz, g.nbz) - g.layer_base;
    if (lz >= g.nbz) lz -= g.nbz;
    if (lz < 0) lz += g.nbz;
    int cell = cx + g.nbx * (cy + g.nby * lz);
    if (lz <= 0 || lz >= g.nlz - 1) {
        atomicExch(words + 2, 1);
        cell = g.num_cells;
    }
    cell_of[base + k] = cell;
    off[base + k] = atomicAdd(&count[cell], 1);
}

// Halo cell table with the halo regions at FIXED slots (capacity, capacity + halo_capacity), so
// nothing about the halo sizes is needed on the host.  The neighbours deliver the per-cell counts
// of their layers into `count_in` (two rows of cells_per_layer) -- not into the count table
// itself, which this slab is still scanning when a fast neighbour's push arrives -- and this
// kernel copies them into the rows of the two halo layers next to the cell starts.
__global__ void __launch_bounds__(kScanThreads)
k_halo_starts_p2p(const int* __restrict__ count_in, int* __restrict__ count, int* __restrict__ start,
                  int cells_per_layer, int first_cell_hi, int base_lo, int base_hi)
{
    const int first = blockIdx.x ? first_cell_hi : 0;
    const int* in = count_in + (blockIdx.x ? cells_per_layer : 0);
    int carry = blockIdx.x ? base_hi : base_lo;
    for (int b = 0; b < cells_per_layer; b += kScanThreads) {
        const int idx = b + threadIdx.x;
        const int v = idx < cells_per_layer ? in[idx] : 0;
        int total;
        const int ex = block_exclusive_scan(v, total);
        if (idx < cells_per_layer) {
            count[first + idx] = v;
            start[first + idx] = carry + ex;
        }
        carry += total;
    }
}

}  // namespace gof
