// gof_force.cuh -- the neighbour force kernel with the integrator fused into its
// epilogue.
//
// Replaces the reference's accelerator kernel (main/acceleration_updater.py:219-458)
// and, in INTEGRATE mode, step2 + boundary_condition + set_sq_speed/max_reduction
// (main/integrator.py:58-129).
//
// The reference walks HALF the stencil (5 / 14 bins) and pushes the reaction
// onto the neighbour with atomics.  Here every particle GATHERS over the full
// 9 / 27 cell stencil: no atomics on the accelerations, a fixed summation
// order (rows in (dz, dy) order, cells by x, particles by (x, id)) and therefore
// bit-reproducible results.  Because the particles are sorted by cell, the
// three x-adjacent cells of a stencil row are ONE contiguous run of the float4
// position array.  How a warp goes through the runs depends on the kind of system:
//   Clusters kind  the 32 home particles of a warp walk the union of their runs as one,
//                  candidates culled against the box of the 32 (walk_culled); warps that
//                  straddle two rows of cells fall back to one walk per lane (run_pairs)
//   Boids kind     per lane; pairs in range are queued and the interactions done with every
//                  lane on its own entries (pair_note / queue_drain)
//   mixed          per lane, interaction behind a branch (pair_branchy)
// plus the literal paths of the reference's z-wrap quirk (quirk_walk, quirk_home) for the few
// home cells that need them, and the split variant for small systems (SPLIT > 1).
// A particle's result never depends on which of these paths its warp takes.
#pragma once
#include <climits>
#include "gof_binning.cuh"

namespace gof {

#ifndef GOF_FORCE_THREADS
#define GOF_FORCE_THREADS 128
#endif
constexpr int kForceThreads = GOF_FORCE_THREADS;  // home particles per block (SPLIT == 1)
constexpr int kBoidSlots = 8;  // per partner type, two float4: {count, sum of relative positions}, {sum of velocities, -}

// The per-type sums of a thread live in shared memory as float4 pairs: vector v (0 / 1) of partner
// type t sits (t * 2 + v) * kForceThreads float4 behind the thread's first one, so that a warp's
// 16-byte accesses are contiguous (conflict-free) and one pair's update is two loads + two stores.
__device__ __forceinline__ float4* boid_vec(float* boid_acc, int t, int v)
{
    return reinterpret_cast<float4*>(boid_acc) + (t * 2 + v) * kForceThreads;
}

// which kinds of species pairs a launch has to handle (decides the code of the hot loop)
enum ForceKind : int {
    kKindClusters = 0,  // every pair is of the Clusters kind: branch-free masked evaluation
    kKindMixed = 1,     // both kinds: test, interaction behind a branch
    kKindBoids = 2,     // every pair is of the Boids kind: test + note, interactions deferred
};

enum ForceMode : int {
    kModeIntegrate = 0,  // engine: fused step2 + boundary condition + max |v|^2, outputs in cell order
    kModeScatter = 1,    // reference-layout API: accelerations (and boid sums) scattered to particle order
};

struct ForceArgs {
    const float4* posw;       // {x, y, z, type bits}, cell order (owned + appended halo)
    const float4* veli;       // {vx, vy, vz, id bits}, cell order, already kicked by step1
    const int* cell_sorted;   // cell of every home particle
    const int* start;         // first index of every cell in the compact (rest) order
    const int* count;         // population of every cell
    const int* pstart;        // first slot of every cell in the snapshot the kernel reads
                              // (== start, or the even-padded table of the packed kernel)
    int home_begin, home_count;
    // Optional second range (slab mode: the two boundary layers in one launch): threads
    // [0, home_count) cover home_begin.., threads [home_count, home_count + home_count2) cover
    // home_begin2...
    int home_begin2, home_count2;
    // Slab mode, interior launch: the range is not known on the host yet; it is derived from
    // the device words {live, bottom layer population, top layer population} (null otherwise).
    const int* interior_words;
    // Slab mode, boundary launch with the sizes still on the device (peer-memory transport): the two
    // ranges are [0, lo) and [live - hi, live) from the same words; home_count is an upper bound
    // of lo + hi for the grid.
    const int* boundary_words;
    const PairParam* table;   // [T*T]
    const unsigned* kind_rows;  // [T] bit j of row i set => (i, j) is of Boids kind
    const double* sep_exact;  // [T*T] args[0] / 5 in float64 (guard band of the separation radius)
    const int* self_kind;     // [T] parameter_matrix[-1, t, t]
    int T;
    int wrap_mode;
    // kModeIntegrate
    float4* posw_out; float4* veli_out; float4* acc_out;
    StepScalars* scalars; int parity_out;
    int integrate;            // 0: leave positions and velocities as they are, only store the acceleration
    // kModeScatter (indexed by particle id)
    float* acc_x; float* acc_y; float* acc_z;
    int* boid_counts; float* boid_acc_x; float* boid_acc_y; float* boid_acc_z;
    float* boid_vel_x; float* boid_vel_y; float* boid_vel_z;
    int n_total;
    int slot_limit;           // upper bound of the indices into posw / veli (Boids-kind queue entries)
    int wave_sms;             // SMs of the device (placement of the literal z-wrap blocks in the first wave)
    int split_homes;          // host side only: the population that decides the row split (0: this launch's)
};

// Per-thread state of the particle being updated.
struct Home {
    float x, y, z;          // raw position
    float vx, vy, vz;       // kicked velocity
    float ax, ay, az;       // accumulated acceleration
    int type;
    unsigned kinds;         // kind_rows[type]
};

// Exact (float64) re-evaluation of the squared distance, in the operation order of
// the reference (`pos_2 -= limit` first, then `pos[i] - pos_2`, then dx*dx + dy*dy + dz*dz
// without contraction).  Only reached inside the guard band around r_max^2 (and on the
// edge of a Boids separation radius).  Scalars by value: taking the address of the
// neighbour's float4 would force it through local memory in the hot loop.
__device__ __noinline__ double exact_dist2(const Geometry& g, float xi, float yi, float zi,
                                           float xj, float yj, float zj, int shifts)
{
    const int sx = (shifts & 3) - 1, sy = ((shifts >> 2) & 3) - 1, sz = ((shifts >> 4) & 3) - 1;
    const double x2 = sx ? __dsub_rn((double)xj, sx * g.Lxd) : (double)xj;
    const double y2 = sy ? __dsub_rn((double)yj, sy * g.Lyd) : (double)yj;
    const double z2 = sz ? __dsub_rn((double)zj, sz * g.Lzd) : (double)zj;
    const double dx = __dsub_rn((double)xi, x2), dy = __dsub_rn((double)yi, y2),
                 dz = __dsub_rn((double)zi, z2);
    return __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
}

__device__ __forceinline__ int pack_shifts(int sx, int sy, int sz)
{
    return (sx + 1) | ((sy + 1) << 2) | ((sz + 1) << 4);
}

// d2 > 1e-10 is guaranteed by the caller, so the flush-to-zero approximate form is safe
// (a single MUFU.RSQ; plain rsqrtf adds denormal scaling around it).  Max error 2 ulp.
__device__ __forceinline__ float fast_rsqrt(float v)
{
    float r;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
    return r;
}

// profile(dist) / dist of a Clusters-kind pair, formed without dist itself: inside r_min
// f_min / d - k0, beyond it f_max / d - s1 |1 - mid / d| (the same piecewise-linear profile as the
// reference's, acceleration_updater.py:369-383, divided through by d; one multiply-add less per pair
// than profile(d2 * inv) * inv and no less accurate: the rounding of rsqrt enters one term instead
// of two).  `qa` = {r_min^2, k0, f_min, mid}, `qb` = {s1, f_max}, inv = 1 / d.  EVERY Clusters-kind
// path evaluates a pair with this function (hot loops, literal z-wrap, cold repairs), so that what
// a repair takes back is bit for bit what the hot loop added.
__device__ __forceinline__ float clusters_scale(const float4 qa, const float2 qb, float d2, float inv)
{
    const float s_in = fmaf(qa.z, inv, -qa.y);
    const float u = fmaf(-qa.w, inv, 1.0f);
    const float v = qb.y * inv;
    const float s_tri = fmaf(-qb.x, fabsf(u), v);
    return d2 < qa.x ? s_in : s_tri;
}

// One accepted pair: add what j does to the home particle.
//   (dx, dy, dz) = home - neighbour (after the periodic shift), d2 = |d|^2
//   (rx, ry, rz) = neighbour - home as the reference accumulates it for cohesion
//   exact_d2()   = the float64 squared distance, evaluated only on the edge of r_sep
// boids(): separation inside r_sep, and the per-type sums for cohesion / alignment
template <typename ExactD2>
__device__ __forceinline__ void interact_boids(Home& h, const float4 qa, float* __restrict__ boid_acc,
                                               const ForceArgs& a, int T, int j, int tj, float dx, float dy,
                                               float dz, float rx, float ry, float rz, float inv, float dist,
                                               ExactD2 exact_d2, const float4* vj_known = nullptr)
{
    bool sep = dist < qa.x;
    if (fabsf(dist - qa.x) < 4e-6f * qa.x) {
        // on the edge of the separation radius the decision is taken in float64
        sep = sqrt(exact_d2()) < a.sep_exact[h.type * T + tj];
    }
    if (sep) {
        const float s = qa.y * inv;
        h.ax = fmaf(s, dx, h.ax);
        h.ay = fmaf(s, dy, h.ay);
        h.az = fmaf(s, dz, h.az);
    }
    const float4 vj = vj_known ? *vj_known : ldg4(a.veli + j);
    float4* s0 = boid_vec(boid_acc, tj, 0);
    float4* s1 = boid_vec(boid_acc, tj, 1);
    float4 c = *s0, v = *s1;
    c.x += 1.0f; c.y += rx; c.z += ry; c.w += rz;
    v.x += vj.x; v.y += vj.y; v.z += vj.z;
    *s0 = c;
    *s1 = v;
}

template <bool BOIDS, typename ExactD2>
__device__ __forceinline__ void interact(Home& h, const PairParam* __restrict__ tabrow,
                                         float* __restrict__ boid_acc, const ForceArgs& a,
                                         int T, int j, int tj, float dx, float dy, float dz,
                                         float rx, float ry, float rz, float d2, ExactD2 exact_d2,
                                         const float4* vj_known = nullptr)
{
    const float inv = fast_rsqrt(d2);
    const float dist = d2 * inv;
    const float4 qa = tabrow[tj].a;
    if (!BOIDS || !((h.kinds >> tj) & 1u)) {
        // clusters(): piecewise linear profile, force along the separation vector (the same
        // arithmetic as the unified pair of the whole-warp walk, pair_mixed)
        const float4 qb = tabrow[tj].b;
        const float s = clusters_scale(make_float4(qa.x * qa.x, qa.y, qa.z, qa.w), make_float2(qb.x, qb.y), d2, inv);
        h.ax = fmaf(-s, dx, h.ax);
        h.ay = fmaf(-s, dy, h.ay);
        h.az = fmaf(-s, dz, h.az);
    } else {
        interact_boids(h, qa, boid_acc, a, T, j, tj, dx, dy, dz, rx, ry, rz, inv, dist, exact_d2, vj_known);
    }
}

// The periodic image of the home particle that a run of neighbours sees: home + s * L
// per axis, s in {-1, 0, 1}.  Subtracting L from a coordinate near L is exact; ADDING L to
// a small coordinate rounds away its low bits, the same way for every pair of the run, so
// a particle next to the lower wall would collect a systematic error.  `lo` holds what the
// rounded sum lost (two-sum); runs that shifted upwards add it back per pair (LOWFIX).
struct Image {
    float x, y, z;     // rounded home + s * L
    float lx, ly, lz;  // exact remainder of that sum
    int shifts;        // packed (sx, sy, sz)
    bool up;           // some `lo` part is non-zero: the run needs the LOWFIX variant
};

__device__ __forceinline__ void shifted(float v, int s, float L, float L_lo, float& hi, float& lo)
{
    const float t = s * L;
    hi = v + t;
    const float bb = hi - v;
    lo = (v - (hi - bb)) + (t - bb) + s * L_lo;  // + the part of the float64 limit that (float)L lost
}

__device__ __forceinline__ Image make_image(const Home& h, const Geometry& g, int sx, int sy, int sz)
{
    Image im;
    if ((sx | sy | sz) == 0) {  // no periodic image (most runs): the two-sums below would yield exactly this
        im.x = h.x; im.y = h.y; im.z = h.z;
        im.lx = im.ly = im.lz = 0.0f;
        im.shifts = pack_shifts(0, 0, 0);
        im.up = false;
        return im;
    }
    shifted(h.x, sx, g.Lx, g.Lx_lo, im.x, im.lx);
    shifted(h.y, sy, g.Ly, g.Ly_lo, im.y, im.ly);
    shifted(h.z, sz, g.Lz, g.Lz_lo, im.z, im.lz);
    im.shifts = pack_shifts(sx, sy, sz);
    im.up = (im.lx != 0.0f) | (im.ly != 0.0f) | (im.lz != 0.0f);  // some remainder to add per pair
    return im;
}

// Shared-memory address of the home type's row in the parameter table of the hot Clusters
// loop.  A row is 16 entries A = {r_min, k0, f_min, mid} (16 B each) followed, 256 bytes on, by
// 16 entries B = {s1, f_max} in 16-byte slots: the pair (i, j) costs ONE address computation
// (row + 16 * tj), the second load uses an immediate offset.  Lanes of a warp have different
// home types but (mostly) the same neighbour type, so they read up to T different rows at
// once; the row stride of 528 bytes (= 4 banks mod 32) puts the rows of any <= 8 home types on
// different banks: one wavefront per load instead of the 3+ that 32-byte AoS rows cost.
struct RowAddr { unsigned a; };
constexpr unsigned kRowBytes = 528u;   // 16 * 16 (A) + 16 * 16 (B) + 16 (bank skew)
constexpr unsigned kRowB = 256u;       // offset of the B entries inside a row

__device__ __forceinline__ float4 lds4(unsigned saddr)
{
    float4 v;
    asm("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(saddr));
    return v;
}
__device__ __forceinline__ float2 lds2(unsigned saddr)
{
    float2 v;
    asm("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(saddr));
    return v;
}

constexpr float kTinyD2 = 1e-30f;  // keeps rsqrt finite for the self pair (whose separation is exactly 0)

// squared distance of the hot Clusters loop: |d|^2 + 1e-30, the offset folded into the first
// multiply-add (it is below half an ulp of any d2 > 1e-22 and makes the self pair's d2 exactly
// kTinyD2)
__device__ __forceinline__ float dist2_tiny(float dx, float dy, float dz)
{
    return fmaf(dz, dz, fmaf(dy, dy, fmaf(dx, dx, kTinyD2)));
}

// Cold path of the Clusters loop: a pair inside the guard band of r_max^2, or closer
// than the reference's 1e-10 lower bound.  The hot loop added the pair's force iff its
// fp32 d2 was below r2_lo; here the reference's float64 decision is taken and the
// difference returned.  Recomputes the fp32 separation exactly as the hot loop does (the
// `lo` parts are zero for images that were not shifted upwards).
__device__ __noinline__ float3 repair_pair_clusters(const Geometry& g, float xi, float yi, float zi,
                                                    float imx, float imy, float imz, float lx,
                                                    float ly, float lz, int shifts, float xj,
                                                    float yj, float zj, const PairParam* q)
{
    const float dx = (imx - xj) + lx, dy = (imy - yj) + ly, dz = (imz - zj) + lz;
    const float d2 = dist2_tiny(dx, dy, dz);
    const bool added = d2 < g.r2_lo;
    const double e = exact_dist2(g, xi, yi, zi, xj, yj, zj, shifts);
    const bool wanted = (1e-10 < e) && (e < g.r2d);
    if (added == wanted) return make_float3(0.f, 0.f, 0.f);
    float4 qa = q->a;
    const float4 qb = q->b;
    qa.x = qa.x * qa.x;  // r_min^2, as k_force's prologue stores it for the hot loop
    float s = clusters_scale(qa, make_float2(qb.x, qb.y), d2, fast_rsqrt(d2));
    if (added) s = -s;  // take it back
    return make_float3(-s * dx, -s * dy, -s * dz);
}

// Branch-free Clusters pair.  With ~830 candidates and ~130 neighbours per particle at least one
// lane of a warp accepts almost every candidate, so a branch around the force evaluation never
// skips it; it only adds divergence bookkeeping and serialises the loop.  The evaluation is
// therefore done for every candidate and its three multiply-adds predicated on d2 < r2_lo
// (the self pair adds s * 0 = 0 with s finite); returns d2 so that the caller can spot the rare
// pairs that need the cold path.
template <bool LOWFIX>
__device__ __forceinline__ float pair_clusters(Home& h, RowAddr tabrow_s, const Geometry& g,
                                               const Image& im, const float4 pj)
{
    float dx = im.x - pj.x, dy = im.y - pj.y, dz = im.z - pj.z;
    if (LOWFIX) { dx += im.lx; dy += im.ly; dz += im.lz; }
    const float d2 = dist2_tiny(dx, dy, dz);
    const unsigned at = tabrow_s.a + ((unsigned)__float_as_int(pj.w) << 4);
    const float4 qa = lds4(at);
    const float2 qb = lds2(at + kRowB);
    const float s = clusters_scale(qa, qb, d2, fast_rsqrt(d2));
    if (d2 < g.r2_lo) {
        h.ax = fmaf(-s, dx, h.ax);
        h.ay = fmaf(-s, dy, h.ay);
        h.az = fmaf(-s, dz, h.az);
    }
    return d2;
}

// needs the cold path: in the guard band, or at/below the reference's 1e-10 bound -- except at
// distance exactly zero (the particle itself, met once by every particle): the hot loop added
// s * 0 = 0 there, which is what the reference's lower bound yields, so nothing is to be repaired
// (the cold path is ~100 instructions with one lane active: 10 % of the kernel when it was taken
// for every self pair).  d2 carries the kTinyD2 offset of dist2_tiny: exactly kTinyD2 for the
// self pair.
__device__ __forceinline__ bool needs_repair(const Geometry& g, float d2)
{
    return (!(d2 < g.r2_lo) && d2 < g.r2_hi) || (d2 <= 1e-10f && d2 > kTinyD2);
}

// The corrections of a run are summed apart from its main chain of fused multiply-adds and added
// at its end: a particle's result must not depend on how the candidates were grouped in fours,
// which differs between the per-lane walk and the culled walk of a whole warp.
__device__ __forceinline__ void repair(float3& rep, const Home& h, const PairParam* __restrict__ tabrow,
                                       const Geometry& g, const Image& im, const float4 pj)
{
    const float3 f = repair_pair_clusters(g, h.x, h.y, h.z, im.x, im.y, im.z, im.lx, im.ly, im.lz,
                                          im.shifts, pj.x, pj.y, pj.z, tabrow + __float_as_int(pj.w));
    rep.x += f.x; rep.y += f.y; rep.z += f.z;
}

// ---- systems with Boids-kind pairs: deferred interactions ----------------------------------
// A Boids-kind pair that is in range costs ~40 instructions (its velocity load and seven
// shared-memory sums), and behind a branch in the candidate loop it runs with the few lanes that
// accepted this neighbour (measured: 15 of 32 lanes active on average).  Instead the candidate
// loop only NOTES a pair that is in range -- the neighbour's index goes to a per-lane queue in
// shared memory, one predicated store -- and the queue is drained with every lane working on
// ITS OWN entries.  A lane's entries are in run order, so its sums are formed in the same order
// whenever the drains happen.
#ifndef GOF_QUEUE_DEPTH
#define GOF_QUEUE_DEPTH 32
#endif
constexpr int kQueueDepth = GOF_QUEUE_DEPTH;  // entries per lane; drained before a lane could overflow

constexpr unsigned kQueueStride = kForceThreads * 4u;  // bytes between a lane's consecutive entries
struct BoidQueue {
    unsigned base;   // shared-memory address of this lane's entry 0 (entries kQueueStride bytes apart)
    unsigned next;   // ... of its first free entry: a push is one predicated store and one predicated add
    float fx, fy, fz;  // force of the drained pairs so far, summed in queue order
};

__device__ __forceinline__ int queue_count(const BoidQueue& q) { return (int)((q.next - q.base) / kQueueStride); }
#ifndef GOF_BOIDS_GROUP
#define GOF_BOIDS_GROUP 10
#endif
constexpr int kBoidsGroup = GOF_BOIDS_GROUP;  // candidates tested per round of the Boids walk (run_pairs)
static_assert(kBoidsGroup >= 1 && kBoidsGroup < kQueueDepth, "a round's notes must fit the queue");
// (a warp vote on this triggers a drain: room for one more round of notes per lane is always kept)
__device__ __forceinline__ bool queue_nearly_full(const BoidQueue& q)
{
    return q.next - q.base > (unsigned)(kQueueDepth - kBoidsGroup) * kQueueStride;
}

// An entry is the neighbour's slot and the periodic image of the run it was met in:
// slot | packed shifts << 26 (at most 2^26 slots per engine while a species pair is Boids-kind).
constexpr int kQueueSlotBits = 26;

__device__ __forceinline__ void queue_push(BoidQueue& q, unsigned entry, bool hit)
{
    if (hit) {
        asm volatile("st.shared.b32 [%0], %1;" ::"r"(q.next), "r"(entry));
        q.next += kQueueStride;
    }
}

// Systems with Boids-kind pairs: the hot loop is nothing but the test and the note.
template <bool LOWFIX>
__device__ __forceinline__ void pair_note(const Geometry& g, const Image& im, const float4 pj, unsigned entry,
                                          BoidQueue& q, bool live = true)
{
    float dx = im.x - pj.x, dy = im.y - pj.y, dz = im.z - pj.z;
    if (LOWFIX) { dx += im.lx; dy += im.ly; dz += im.lz; }
    const float d2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
    queue_push(q, entry, live && d2 < g.r2_hi && d2 > 1e-10f);
}

// Out of line (one copy: the hot loops stay small).  The entries of a lane are handled one after
// the other and add to ONE force accumulator in queue order, so the sum does not depend on where
// the drains fall (a warp vote triggers them: warp-mates must not matter).  The queue outlives
// the runs -- per run the lanes of a warp collect very different numbers of pairs (a particle in
// the upper corner of its cell meets its neighbours in the upper rows), over the whole stencil
// about the same -- so every entry rebuilds its periodic image from the packed shifts.  The `lo`
// parts are zero for images that were not shifted upwards: adding them changes nothing.
__device__ __noinline__ float3 queue_drain(float3 f, unsigned qbase, int count, float hx, float hy, float hz,
                                           int type, unsigned kinds,
                                           const PairParam* __restrict__ tabrow, float* __restrict__ boid_acc,
                                           const ForceArgs& a, const Geometry& g)
{
    Home h;
    h.x = hx; h.y = hy; h.z = hz;
    h.vx = h.vy = h.vz = 0.0f;  // (not used by the pair interactions)
    h.ax = f.x; h.ay = f.y; h.az = f.z;
    h.type = type;
    h.kinds = kinds;
    for (int k = 0; k < count; ++k) {
        unsigned entry;
        asm volatile("ld.shared.b32 %0, [%1];" : "=r"(entry) : "r"(qbase + (unsigned)k * (kForceThreads * 4u)));
        const int j = (int)(entry & ((1u << kQueueSlotBits) - 1u));
        const int shifts = (int)(entry >> kQueueSlotBits);
        const float4 pj = ldg4(a.posw + j);
        // the velocity is requested together with the position (a queued pair all but always
        // interacts): one round trip to L2 per entry instead of two dependent ones
        const float4 vj = ldg4(a.veli + j);
        float dx = hx - pj.x, dy = hy - pj.y, dz = hz - pj.z;
        if (shifts != pack_shifts(0, 0, 0)) {  // met across a periodic boundary (home cells on the box faces)
            const Image im = make_image(h, g, (shifts & 3) - 1, ((shifts >> 2) & 3) - 1, ((shifts >> 4) & 3) - 1);
            dx = (im.x - pj.x) + im.lx; dy = (im.y - pj.y) + im.ly; dz = (im.z - pj.z) + im.lz;
        }
        const float d2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
        bool ok = true;
        if (d2 > g.r2_lo) {
            const double e = exact_dist2(g, hx, hy, hz, pj.x, pj.y, pj.z, shifts);
            ok = (1e-10 < e) && (e < g.r2d);
        }
        if (ok)
            interact<true>(h, tabrow, boid_acc, a, a.T, j, __float_as_int(pj.w), dx, dy, dz, -dx, -dy, -dz, d2,
                           [&]() { return exact_dist2(g, hx, hy, hz, pj.x, pj.y, pj.z, shifts); }, &vj);
    }
    return make_float3(h.ax, h.ay, h.az);
}

__device__ __forceinline__ void drain(BoidQueue& q, const Home& h, const PairParam* __restrict__ tabrow,
                                      float* __restrict__ boid_acc, const ForceArgs& a, const Geometry& g)
{
    const float3 f = queue_drain(make_float3(q.fx, q.fy, q.fz), q.base, queue_count(q), h.x, h.y, h.z,
                                 h.type, h.kinds, tabrow, boid_acc, a, g);
    q.fx = f.x; q.fy = f.y; q.fz = f.z;
    q.next = q.base;
}

// Mixed systems (both kinds of pairs): distance test, then the interaction behind a branch.
// (Queueing was measured here too: with ~130 pairs in range per particle the lanes of a warp fill
// their queues at very different times -- a particle low in its cell meets its neighbours in the
// first rows of the stencil walk -- and the drains ran 13 lanes wide: 3.0 ms against 2.6 ms.)
template <bool LOWFIX>
__device__ __forceinline__ void pair_branchy(Home& h, const PairParam* __restrict__ tabrow,
                                             float* __restrict__ boid_acc, const ForceArgs& a,
                                             const Geometry& g, const Image& im, int j, const float4 pj)
{
    float dx = im.x - pj.x, dy = im.y - pj.y, dz = im.z - pj.z;
    if (LOWFIX) { dx += im.lx; dy += im.ly; dz += im.lz; }
    const float d2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
    if (d2 < g.r2_hi && d2 > 1e-10f) {
        bool ok = true;
        if (d2 > g.r2_lo) {
            const double e = exact_dist2(g, h.x, h.y, h.z, pj.x, pj.y, pj.z, im.shifts);
            ok = (1e-10 < e) && (e < g.r2d);
        }
        if (ok)
            interact<true>(h, tabrow, boid_acc, a, a.T, j, __float_as_int(pj.w), dx, dy, dz, -dx, -dy, -dz, d2,
                           [&]() { return exact_dist2(g, h.x, h.y, h.z, pj.x, pj.y, pj.z, im.shifts); });
    }
}

// A contiguous run [s, e) of neighbour particles seen through one periodic image.
// Four independent 16-byte loads are issued before the four tests that consume them, so
// the L1 latency of a neighbour is hidden behind the arithmetic of the previous ones.
template <int KIND, bool LOWFIX>
__device__ __forceinline__ void run_pairs(Home& h, RowAddr tabrow_s,
                                          const PairParam* __restrict__ tabrow,
                                          float* __restrict__ boid_acc, const ForceArgs& a,
                                          const Geometry& g, const Image& im, int s, int e, BoidQueue& q)
{
    // the run sums into its own accumulator (shorter partial sums: less fp32 rounding)
    const float ax0 = h.ax, ay0 = h.ay, az0 = h.az;
    h.ax = h.ay = h.az = 0.0f;
    float3 rep = make_float3(0.0f, 0.0f, 0.0f);
    const unsigned tag = (unsigned)im.shifts << kQueueSlotBits;
    int j = s;
    if (KIND == kKindBoids) {
        // Masked groups of kBoidsGroup candidates: every load of a group is in flight before the
        // first test, and the last group of a run re-reads the run's last candidate and notes
        // nothing for the slots past the end.  At three particles per cell a run is ~9 candidates:
        // one round trip to L2 instead of the three of "two groups of four and a tail, one load at
        // a time" (N = 4M: 1.26 -> 1.11 ms, flocked 2.71 -> 2.41 ms; groups of 4 / 6 / 8 / 12: 1.15 /
        // 1.11 / 1.12 / 1.21 ms).  The kernel is bound by these round trips, not by instructions:
        // a push of 2.75 instead of 7 instructions changed nothing.
        constexpr int G = kBoidsGroup;
        for (; j < e; j += G) {
            float4 p[G];
#pragma unroll
            for (int k = 0; k < G; ++k) p[k] = ldg4(a.posw + min(j + k, e - 1));
            const unsigned tj = tag + (unsigned)j;
#pragma unroll
            for (int k = 0; k < G; ++k) pair_note<LOWFIX>(g, im, p[k], tj + (unsigned)k, q, j + k < e);
            // lanes drain together, each its own entries
            if (__any_sync(__activemask(), queue_nearly_full(q))) drain(q, h, tabrow, boid_acc, a, g);
        }
        h.ax += ax0; h.ay += ay0; h.az += az0;
        return;
    }
    for (; j + 4 <= e; j += 4) {
        // four independent 16-byte loads issued ahead of the arithmetic that consumes them
        const float4 p0 = ldg4(a.posw + j), p1 = ldg4(a.posw + j + 1), p2 = ldg4(a.posw + j + 2),
                     p3 = ldg4(a.posw + j + 3);
        if (KIND == kKindMixed) {
            pair_branchy<LOWFIX>(h, tabrow, boid_acc, a, g, im, j, p0);
            pair_branchy<LOWFIX>(h, tabrow, boid_acc, a, g, im, j + 1, p1);
            pair_branchy<LOWFIX>(h, tabrow, boid_acc, a, g, im, j + 2, p2);
            pair_branchy<LOWFIX>(h, tabrow, boid_acc, a, g, im, j + 3, p3);
        } else {
            // four independent, branch-free evaluations the scheduler can interleave; the rare
            // pairs in the guard band (~1 in 1e5) or below 1e-10 (the self pair) are repaired
            // after the group, found with one vote over the four squared distances
            const float d0 = pair_clusters<LOWFIX>(h, tabrow_s, g, im, p0);
            const float d1 = pair_clusters<LOWFIX>(h, tabrow_s, g, im, p1);
            const float d2 = pair_clusters<LOWFIX>(h, tabrow_s, g, im, p2);
            const float d3 = pair_clusters<LOWFIX>(h, tabrow_s, g, im, p3);
            const bool b0 = !(d0 < g.r2_lo) && d0 < g.r2_hi, b1 = !(d1 < g.r2_lo) && d1 < g.r2_hi,
                       b2 = !(d2 < g.r2_lo) && d2 < g.r2_hi, b3 = !(d3 < g.r2_lo) && d3 < g.r2_hi;
            const bool tiny = fminf(fminf(d0, d1), fminf(d2, d3)) <= 1e-10f;
            if (b0 | b1 | b2 | b3 | tiny) {
                if (needs_repair(g, d0)) repair(rep, h, tabrow, g, im, p0);
                if (needs_repair(g, d1)) repair(rep, h, tabrow, g, im, p1);
                if (needs_repair(g, d2)) repair(rep, h, tabrow, g, im, p2);
                if (needs_repair(g, d3)) repair(rep, h, tabrow, g, im, p3);
            }
        }
    }
    for (; j < e; ++j) {  // at most three
        const float4 pj = ldg4(a.posw + j);
        if (KIND == kKindMixed) pair_branchy<LOWFIX>(h, tabrow, boid_acc, a, g, im, j, pj);
        else if (needs_repair(g, pair_clusters<LOWFIX>(h, tabrow_s, g, im, pj))) repair(rep, h, tabrow, g, im, pj);
    }
    if (KIND == kKindClusters) { h.ax += rep.x; h.ay += rep.y; h.az += rep.z; }
    h.ax += ax0; h.ay += ay0; h.az += az0;
}

// (the literal z-wrap pair of the Clusters kind and its cold path are defined further down)
__device__ __forceinline__ float pair_clusters_quirk(Home& h, RowAddr tabrow_s, const Geometry& g,
                                                     const Image& im, bool direct, const float4 pj);
__device__ __forceinline__ void repair_quirk(float3& rep, const Home& h, const PairParam* __restrict__ tabrow,
                                             const Geometry& g, const Image& im, int sx, int sy, bool direct,
                                             const float4 pj);
__device__ __forceinline__ void quirk_dz(const Geometry& g, bool direct, float zi, float yi, float zj,
                                         float yj, float& dz, float& relz);
__device__ __noinline__ float3 repair_quirk_clusters(const Geometry& g, bool direct, float hx, float hy,
                                                     float hz, float pjx, float pjy, float pjz, int sx,
                                                     int sy, float dx, float dy, float dz, float d2,
                                                     const PairParam* q);

// ---- Clusters kind: one walk for the whole warp, culled against the box of its home particles --
// When the 32 home particles of a warp sit in one row of cells (same y and z cell), the union of
// their runs is itself one contiguous run per stencil row, and the warp can walk it as one: every
// lane meets a superset of its own candidates in the same order, and the extra ones are farther
// than one bin (>= r_max) from it, i.e. add exactly zero.  What this buys is warp-uniform control
// flow, which makes culling possible: about half of the candidates are out of range for EVERY
// lane.  Per chunk of 64 candidates each lane tests TWO against the box of the warp's home
// particles, the votes are collected with a ballot, and only the surviving candidates are
// evaluated, four at a time as in the per-lane walk.  A skipped candidate would have added exactly zero, so a particle's result
// does not depend on which walk its warp takes.
struct Box { float lox, loy, loz, hix, hiy, hiz; };

// order-preserving map of a float onto unsigned (any sign: a caller may upload positions that
// are outside the box, the boundary condition only brings them back after the step)
__device__ __forceinline__ unsigned float_key(float v)
{
    const unsigned b = __float_as_uint(v);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float key_float(unsigned k)
{
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

__device__ __forceinline__ Box warp_box(const Home& h)
{
    const unsigned ux = float_key(h.x), uy = float_key(h.y), uz = float_key(h.z);
    Box b;
    b.lox = key_float(__reduce_min_sync(0xffffffffu, ux));
    b.loy = key_float(__reduce_min_sync(0xffffffffu, uy));
    b.loz = key_float(__reduce_min_sync(0xffffffffu, uz));
    b.hix = key_float(__reduce_max_sync(0xffffffffu, ux));
    b.hiy = key_float(__reduce_max_sync(0xffffffffu, uy));
    b.hiz = key_float(__reduce_max_sync(0xffffffffu, uz));
    return b;
}

constexpr int kListLen = 128;  // mixed kind: survivors a warp can hold (a ring, a power of two)
// Clusters kind: the survivor list is linear.  A run appends chunk after chunk behind the < 4
// entries that still wait for a full group; only when the next chunk of 64 might not fit are
// those moved to the front (runs of more than ~3 chunks: crowded cells).
constexpr int kWalkCap = 192;

__device__ __forceinline__ void sts4(unsigned saddr, const float4 v)
{
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
                 : "memory");
}
__device__ __forceinline__ float4 lds4v(unsigned saddr)
{
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "r"(saddr)
                 : "memory");
    return v;
}

// Four survivors from the warp's list (one broadcast 16-byte shared-memory load each), every lane,
// no branch.  A group in which this lane has a pair for the cold path (guard band of r_max^2, or
// closer than the reference's 1e-10 bound -- the lane's own particle is one, once per stencil) is
// only NOTED: mark = mark * 4096 + address.  After the chunk, mark == 0 says nothing to do, a
// value below 2^18 (the shared-memory window) is the address of the one noted group, anything
// larger means several (addresses grow along the loop and are >= 64, so this holds whatever
// overflows).  A branch to the cold code per group, taken once in ~100 groups, cost 2-3 % of the
// loop's instructions and kept the compiler from scheduling across it.
// LITERAL: the home cells are in the bottom / top layer in GOF_WRAP_REFERENCE mode; the pair is
// evaluated by the literal z-wrap rules, in the role (`direct` or not, see quirk_home) that follows
// from the survivor's slot `at_j` against the threshold `own` of the lane.
template <bool LITERAL, bool LOWFIX>
__device__ __forceinline__ void walk_group(Home& h, unsigned& mark, RowAddr tabrow_s, const Geometry& g,
                                           const Image& im, unsigned at, unsigned at_j, int own)
{
    const float4 p0 = lds4v(at), p1 = lds4v(at + 16u), p2 = lds4v(at + 32u), p3 = lds4v(at + 48u);
    float d0, d1, d2, d3;
    if (LITERAL) {
        const float4 jj = lds4v(at_j);
        d0 = pair_clusters_quirk(h, tabrow_s, g, im, __float_as_int(jj.x) >= own, p0);
        d1 = pair_clusters_quirk(h, tabrow_s, g, im, __float_as_int(jj.y) >= own, p1);
        d2 = pair_clusters_quirk(h, tabrow_s, g, im, __float_as_int(jj.z) >= own, p2);
        d3 = pair_clusters_quirk(h, tabrow_s, g, im, __float_as_int(jj.w) >= own, p3);
    } else {
        d0 = pair_clusters<LOWFIX>(h, tabrow_s, g, im, p0);
        d1 = pair_clusters<LOWFIX>(h, tabrow_s, g, im, p1);
        d2 = pair_clusters<LOWFIX>(h, tabrow_s, g, im, p2);
        d3 = pair_clusters<LOWFIX>(h, tabrow_s, g, im, p3);
    }
    const unsigned noted = mark * 4096u + at;
    const bool ok = (d0 < g.r2_lo || !(d0 < g.r2_hi)) && (d1 < g.r2_lo || !(d1 < g.r2_hi)) &&
                    (d2 < g.r2_lo || !(d2 < g.r2_hi)) && (d3 < g.r2_lo || !(d3 < g.r2_hi)) &&
                    fminf(fminf(d0, d1), fminf(d2, d3)) > 1e-10f;
    mark = ok ? mark : noted;
}

// The cold path of a chunk: entries [at, end) of the list once more, in order, for the lanes that
// flagged one of them.  The squared distance is formed exactly as the hot loop formed it (the `lo`
// parts of an image that was not shifted upwards are zero: adding them changes nothing).
template <bool LITERAL>
__device__ __noinline__ float3 repair_entries(float3 rep, float hx, float hy, float hz, float imx, float imy,
                                              float imz, float lx, float ly, float lz, int shifts,
                                              const PairParam* __restrict__ tabrow, const Geometry& g,
                                              unsigned at, unsigned end, unsigned at_j, int own, int sx, int sy)
{
    for (; at != end; at += 16u, at_j += 4u) {
        const float4 pj = lds4v(at);
        const PairParam* q = tabrow + __float_as_int(pj.w);
        float3 f;
        if (LITERAL) {
            int j;
            asm volatile("ld.shared.b32 %0, [%1];" : "=r"(j) : "r"(at_j) : "memory");
            const bool direct = j >= own;
            float dz, relz;
            quirk_dz(g, direct, hz, hy, pj.z, pj.y, dz, relz);
            const float dx = (imx - pj.x) + lx, dy = (imy - pj.y) + ly;
            const float d2 = dist2_tiny(dx, dy, dz);
            if (!needs_repair(g, d2)) continue;
            f = repair_quirk_clusters(g, direct, hx, hy, hz, pj.x, pj.y, pj.z, sx, sy, dx, dy, dz, d2, q);
        } else {
            const float dx = (imx - pj.x) + lx, dy = (imy - pj.y) + ly, dz = (imz - pj.z) + lz;
            if (!needs_repair(g, dist2_tiny(dx, dy, dz))) continue;
            f = repair_pair_clusters(g, hx, hy, hz, imx, imy, imz, lx, ly, lz, shifts, pj.x, pj.y, pj.z, q);
        }
        rep.x += f.x; rep.y += f.y; rep.z += f.z;
    }
    return rep;
}

// `lst`: shared-memory address of the warp's survivor list (kWalkCap float4).  Per chunk of 64
// candidates every lane loads two and tests them against the box; the survivors are appended in
// order (ballot + prefix count) and the list is consumed four at a time with a running address,
// the < 4 left-overs waiting where they are for the next chunk.  The run's last group is padded
// with a particle parked far outside the box.
//
// LITERAL: `im` carries no z shift; the z separation of a pair is whatever the literal rules make
// it, so a candidate survives if it is in reach of the box as it is or one box length up or down.
// `lst_j`: the survivors' slots (LITERAL only).
template <bool LITERAL, bool LOWFIX>
__device__ __forceinline__ void walk_culled(Home& h, RowAddr tabrow_s, const PairParam* __restrict__ tabrow,
                                            const ForceArgs& a, const Geometry& g, const Image& im,
                                            const Box& b, int sx, int sy, int sz, int s, int e, unsigned lst,
                                            unsigned lst_j, int own)
{
    // the run sums into its own accumulator (shorter partial sums: less fp32 rounding)
    const float ax0 = h.ax, ay0 = h.ay, az0 = h.az;
    h.ax = h.ay = h.az = 0.0f;
    float3 rep = make_float3(0.0f, 0.0f, 0.0f);
    // the box as this run sees it (the periodic shift is the same for all lanes)
    const float tx = sx * g.Lx, ty = sy * g.Ly, tz = sz * g.Lz;
    const float lox = b.lox + tx - g.cull_slack, hix = b.hix + tx + g.cull_slack;
    const float loy = b.loy + ty - g.cull_slack, hiy = b.hiy + ty + g.cull_slack;
    const float loz = b.loz + tz - g.cull_slack, hiz = b.hiz + tz + g.cull_slack;
    const unsigned lane = (unsigned)lane_id(), below = (1u << lane) - 1u;
    unsigned head = 0u, tail = 0u;  // entries [head, tail) wait; warp-uniform; head is a multiple of four
    // chunks of 64 candidates, two per lane: half as many ballot / append / syncwarp rounds
    for (int base = s; base < e; base += 64) {
        // (no branch around the loads and the tests: a slot past the run's end re-reads its last
        // candidate and is dropped by the comparison of its index)
        const int j0 = base + (int)lane, j1 = j0 + 32;
        const float4 p0 = ldg4(a.posw + min(j0, e - 1)), p1 = ldg4(a.posw + min(j1, e - 1));
        auto reach = [&](const float4 p) {
            const float ex = fmaxf(fmaxf(lox - p.x, p.x - hix), 0.0f);
            const float ey = fmaxf(fmaxf(loy - p.y, p.y - hiy), 0.0f);
            float ez = fmaxf(fmaxf(loz - p.z, p.z - hiz), 0.0f);
            if (LITERAL) {
                const float up = fmaxf(fmaxf((loz + g.Lz) - p.z, p.z - (hiz + g.Lz)), 0.0f);
                const float dn = fmaxf(fmaxf((loz - g.Lz) - p.z, p.z - (hiz - g.Lz)), 0.0f);
                ez = fminf(ez, fminf(up, dn));
            }
            return fmaf(ez, ez, fmaf(ey, ey, ex * ex)) < g.cull_r2;
        };
        const bool keep0 = reach(p0) && j0 < e, keep1 = reach(p1) && j1 < e;
        if (tail + 64u > (unsigned)kWalkCap) {  // (warp-uniform, rare) make room: the waiting entries go to the front
            float4 w = p0;
            int wj = 0;
            const bool mine = lane < tail - head;
            if (mine) {
                w = lds4v(lst + ((head + lane) << 4));
                if (LITERAL) asm volatile("ld.shared.b32 %0, [%1];" : "=r"(wj) : "r"(lst_j + ((head + lane) << 2)) : "memory");
            }
            __syncwarp();
            if (mine) {
                sts4(lst + (lane << 4), w);
                if (LITERAL) asm volatile("st.shared.b32 [%0], %1;" ::"r"(lst_j + (lane << 2)), "r"(wj) : "memory");
            }
            tail -= head;
            head = 0u;
        }
        const unsigned m0 = __ballot_sync(0xffffffffu, keep0), m1 = __ballot_sync(0xffffffffu, keep1);
        if (keep0) {
            const unsigned at = tail + __popc(m0 & below);
            sts4(lst + (at << 4), p0);
            if (LITERAL) asm volatile("st.shared.b32 [%0], %1;" ::"r"(lst_j + (at << 2)), "r"(j0) : "memory");
        }
        if (keep1) {
            const unsigned at = tail + __popc(m0) + __popc(m1 & below);
            sts4(lst + (at << 4), p1);
            if (LITERAL) asm volatile("st.shared.b32 [%0], %1;" ::"r"(lst_j + (at << 2)), "r"(j1) : "memory");
        }
        tail += __popc(m0) + __popc(m1);
        __syncwarp();
        const unsigned stop = tail & ~3u;
        if (stop != head) {
            const unsigned first = lst + (head << 4), end = lst + (stop << 4);
            unsigned mark = 0u, at = first, at_j = lst_j + (head << 2);
            do {
                walk_group<LITERAL, LOWFIX>(h, mark, tabrow_s, g, im, at, at_j, own);
                at += 64u;
                at_j += 16u;
            } while (at != end);
            if (__any_sync(0xffffffffu, mark != 0u)) {
                if (mark != 0u) {  // one noted group: that one; several: the whole chunk once more
                    const unsigned f = mark < (1u << 18) ? mark : first, t = mark < (1u << 18) ? mark + 64u : end;
                    rep = repair_entries<LITERAL>(rep, h.x, h.y, h.z, im.x, im.y, im.z, im.lx, im.ly, im.lz, im.shifts,
                                                  tabrow, g, f, t, lst_j + ((f - lst) >> 2), own, sx, sy);
                }
                __syncwarp();
            }
            head = stop;
        }
    }
    if (tail != head) {
        if (lane < 4u - (tail - head)) {
            const unsigned at = tail + lane;
            sts4(lst + (at << 4), make_float4(kDummyPos, kDummyPos, kDummyPos, 0.0f));
            if (LITERAL) asm volatile("st.shared.b32 [%0], %1;" ::"r"(lst_j + (at << 2)), "r"(0) : "memory");
        }
        __syncwarp();
        const unsigned first = lst + (head << 4);
        unsigned mark = 0u;
        walk_group<LITERAL, LOWFIX>(h, mark, tabrow_s, g, im, first, lst_j + (head << 2), own);
        if (mark != 0u)
            rep = repair_entries<LITERAL>(rep, h.x, h.y, h.z, im.x, im.y, im.z, im.lx, im.ly, im.lz, im.shifts, tabrow, g,
                                          first, first + 64u, lst_j + (head << 2), own, sx, sy);
    }
    __syncwarp();  // the list is read before the next run appends to it
    h.ax += rep.x; h.ay += rep.y; h.az += rep.z;
    h.ax += ax0; h.ay += ay0; h.az += az0;
}

__device__ __forceinline__ void walk_image(Home& h, RowAddr tabrow_s, const PairParam* __restrict__ tabrow,
                                           const ForceArgs& a, const Geometry& g, const Box& b,
                                           int sx, int sy, int sz, int s, int e, unsigned lst)
{
    const Image im = make_image(h, g, sx, sy, sz);
    if (__any_sync(0xffffffffu, im.up))
        walk_culled<false, true>(h, tabrow_s, tabrow, a, g, im, b, sx, sy, sz, s, e, lst, 0u, 0);
    else
        walk_culled<false, false>(h, tabrow_s, tabrow, a, g, im, b, sx, sy, sz, s, e, lst, 0u, 0);
}

// ---- mixed systems (Clusters- and Boids-kind pairs): the whole-warp walk, one unified pair ------
// The per-lane walk (pair_branchy) runs the interaction behind a branch with the few lanes that
// have this candidate in range (15 of 32 on average), and a Boids-kind pair costs seven
// shared-memory read-modify-writes.  When the 32 home particles of a warp share a row of cells the
// warp walks the union of their runs as one, like the Clusters kind (walk_culled): candidates are
// culled against the box of the 32, the survivors' positions AND velocities go to two lists, and
// every lane evaluates every survivor BRANCH-FREE: both force laws computed, one selected by the
// kind of the species pair and the cut-off, the Boids sums multiplied by a 0 / 1 mask.  All lanes
// see the same candidate at the same time, so its type is warp-uniform: the seven running sums of
// the partner type the walk is in stay in registers and are exchanged with their shared-memory
// slot (two 16-byte loads, two stores) only when the type changes.  The sums still grow pair by
// pair in candidate order
// (slot -> registers -> slot moves nothing), and a masked-out pair adds exactly zero, so a
// particle's result equals that of the per-lane walk bit for bit.  The float64 decisions (guard
// band of r_max^2, edge of the separation radius) are rare: a block of four survivors in which
// some lane needs one is redone by the per-lane code, pair by pair.
struct BoidRun {
    float c, rx, ry, rz, vx, vy, vz;  // count, sum of relative positions, sum of velocities
    int t;                             // partner type they belong to (-1: nothing in registers)
};

__device__ __forceinline__ void boidrun_store(BoidRun& r, float* __restrict__ boid_acc)
{
    if (r.t < 0) return;
    *boid_vec(boid_acc, r.t, 0) = make_float4(r.c, r.rx, r.ry, r.rz);
    *boid_vec(boid_acc, r.t, 1) = make_float4(r.vx, r.vy, r.vz, 0.0f);
    r.t = -1;
}

__device__ __forceinline__ void boidrun_load(BoidRun& r, const float* __restrict__ boid_acc, int t)
{
    const float4 c = *boid_vec(const_cast<float*>(boid_acc), t, 0), v = *boid_vec(const_cast<float*>(boid_acc), t, 1);
    r.c = c.x; r.rx = c.y; r.ry = c.z; r.rz = c.w;
    r.vx = v.x; r.vy = v.y; r.vz = v.z;
    r.t = t;
}

// One survivor, every lane, no branch.  Returns true when this lane needs a float64 decision for
// it (the caller then redoes the whole block with the per-lane code; nothing of this call counts).
struct MixedPair { float s, dx, dy, dz; bool in, sums, exact; };

// The parameter rows of a mixed system are laid out so that ONE formula serves both kinds
// (k_force's prologue): a Clusters-kind pair stores A = {r_min^2, k0, f_min, mid}, B = {s1, f_max}
// like the Clusters kernel, a Boids-kind pair A = {r_sep^2, 0, -separation, 0}, B = {0, 0} --
// clusters_scale then yields -separation / d inside r_sep and 0 outside, i.e. exactly the Boids
// separation term once negated like every Clusters force.  Whether the pair also counts for the
// lane's per-type sums is bit t of the home species' kinds row (a register), not another table
// word: the walk is bound by shared-memory wavefronts, and 24 bytes of parameters per lane and
// candidate are six of them where 32 were eight.
template <bool LOWFIX>
__device__ __forceinline__ MixedPair pair_mixed(unsigned kinds, RowAddr tabrow_s, const Geometry& g, const Image& im,
                                                const float4 pj)
{
    MixedPair r;
    float dx = im.x - pj.x, dy = im.y - pj.y, dz = im.z - pj.z;
    if (LOWFIX) { dx += im.lx; dy += im.ly; dz += im.lz; }
    const float d2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
    const unsigned at = tabrow_s.a + ((unsigned)__float_as_int(pj.w) << 4);
    const float4 qa = lds4(at);
    const float2 qb = lds2(at + kRowB);
    const bool below = d2 < g.r2_lo;
    r.in = below && d2 > 1e-10f;
    r.sums = r.in && ((kinds >> __float_as_int(pj.w)) & 1u);
    r.s = clusters_scale(qa, qb, d2, fast_rsqrt(d2));
    r.dx = dx; r.dy = dy; r.dz = dz;
    // float64 decisions: the guard band of r_max^2, and -- in squared distances, a little wider than
    // the per-lane code's |dist - r_sep| < 4e-6 r_sep -- the edge of the radius in qa.x (the
    // separation radius of a Boids-kind pair; for a Clusters-kind pair it is r_min, where nothing
    // has to be decided, but a flag there is as rare and costs nothing else)
    r.exact = (!below && d2 < g.r2_hi) || fabsf(d2 - qa.x) < 1e-5f * qa.x;
    return r;
}

// (predicated, not masked: a pair out of range costs the issue slots of its ten additions but no
// select in front of them; what it would have added is exactly zero either way)
__device__ __forceinline__ void apply_mixed(Home& h, BoidRun& run, const MixedPair& r, const float4 vj)
{
    if (r.in) {
        h.ax = fmaf(-r.s, r.dx, h.ax);
        h.ay = fmaf(-r.s, r.dy, h.ay);
        h.az = fmaf(-r.s, r.dz, h.az);
    }
    if (r.sums) {
        run.c += 1.0f;
        run.rx -= r.dx; run.ry -= r.dy; run.rz -= r.dz;
        run.vx += vj.x; run.vy += vj.y; run.vz += vj.z;
    }
}

// four survivors from the warp's lists (positions at `at`, velocities at `at_v`)
template <bool LOWFIX>
__device__ __forceinline__ void walk_group_mixed(Home& h, BoidRun& run, RowAddr tabrow_s,
                                                 const PairParam* __restrict__ tabrow, float* __restrict__ boid_acc,
                                                 const ForceArgs& a, const Geometry& g, const Image& im, unsigned at,
                                                 unsigned at_v)
{
    const float4 p0 = lds4v(at), p1 = lds4v(at + 16u), p2 = lds4v(at + 32u), p3 = lds4v(at + 48u);
    const MixedPair r0 = pair_mixed<LOWFIX>(h.kinds, tabrow_s, g, im, p0), r1 = pair_mixed<LOWFIX>(h.kinds, tabrow_s, g, im, p1),
                    r2 = pair_mixed<LOWFIX>(h.kinds, tabrow_s, g, im, p2), r3 = pair_mixed<LOWFIX>(h.kinds, tabrow_s, g, im, p3);
    if (!__any_sync(0xffffffffu, r0.exact | r1.exact | r2.exact | r3.exact)) {
        const float4 v0 = lds4v(at_v), v1 = lds4v(at_v + 16u), v2 = lds4v(at_v + 32u), v3 = lds4v(at_v + 48u);
        // the candidate's type is the same in every lane: a warp-uniform exchange of the running sums
        int t = __float_as_int(p0.w);
        if (t != run.t) { boidrun_store(run, boid_acc); boidrun_load(run, boid_acc, t); }
        apply_mixed(h, run, r0, v0);
        t = __float_as_int(p1.w);
        if (t != run.t) { boidrun_store(run, boid_acc); boidrun_load(run, boid_acc, t); }
        apply_mixed(h, run, r1, v1);
        t = __float_as_int(p2.w);
        if (t != run.t) { boidrun_store(run, boid_acc); boidrun_load(run, boid_acc, t); }
        apply_mixed(h, run, r2, v2);
        t = __float_as_int(p3.w);
        if (t != run.t) { boidrun_store(run, boid_acc); boidrun_load(run, boid_acc, t); }
        apply_mixed(h, run, r3, v3);
        return;
    }
    // some lane needs a float64 decision in this block: the per-lane code for all four, on the
    // shared-memory slots (out of line: ~1 block in 1000)
    boidrun_store(run, boid_acc);
#pragma unroll 1
    for (int k = 0; k < 4; ++k) {
        const float4 pj = lds4v(at + 16u * k), vj = lds4v(at_v + 16u * k);
        float dx = im.x - pj.x, dy = im.y - pj.y, dz = im.z - pj.z;
        if (LOWFIX) { dx += im.lx; dy += im.ly; dz += im.lz; }
        const float d2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
        if (d2 < g.r2_hi && d2 > 1e-10f) {
            bool ok = true;
            if (d2 > g.r2_lo) {
                const double e = exact_dist2(g, h.x, h.y, h.z, pj.x, pj.y, pj.z, im.shifts);
                ok = (1e-10 < e) && (e < g.r2d);
            }
            if (ok)
                interact<true>(h, tabrow, boid_acc, a, a.T, 0, __float_as_int(pj.w), dx, dy, dz, -dx, -dy, -dz, d2,
                               [&]() { return exact_dist2(g, h.x, h.y, h.z, pj.x, pj.y, pj.z, im.shifts); }, &vj);
        }
    }
}

// like walk_culled (not LITERAL): `lst` / `lst_v` = the warp's survivor lists (kListLen float4 each,
// the velocities kListLen * 16 bytes behind the positions; linear, the < 4 waiting entries moved
// to the front when the next chunk might not fit)
template <bool LOWFIX>
__device__ __forceinline__ void walk_culled_mixed(Home& h, BoidRun& run, RowAddr tabrow_s,
                                                  const PairParam* __restrict__ tabrow, float* __restrict__ boid_acc,
                                                  const ForceArgs& a, const Geometry& g, const Image& im, const Box& b,
                                                  int sx, int sy, int sz, int s, int e, unsigned lst, unsigned lst_v)
{
    // the run sums into its own accumulator, like run_pairs
    const float ax0 = h.ax, ay0 = h.ay, az0 = h.az;
    h.ax = h.ay = h.az = 0.0f;
    const float tx = sx * g.Lx, ty = sy * g.Ly, tz = sz * g.Lz;
    const float lox = b.lox + tx - g.cull_slack, hix = b.hix + tx + g.cull_slack;
    const float loy = b.loy + ty - g.cull_slack, hiy = b.hiy + ty + g.cull_slack;
    const float loz = b.loz + tz - g.cull_slack, hiz = b.hiz + tz + g.cull_slack;
    const unsigned lane = (unsigned)lane_id(), below = (1u << lane) - 1u;
    const unsigned dv = lst_v - lst;
    unsigned head = 0u, tail = 0u;  // entries [head, tail) wait; warp-uniform; head is a multiple of four
    for (int base = s; base < e; base += 64) {
        const int j0 = base + (int)lane, j1 = j0 + 32;
        const float4 p0 = ldg4(a.posw + min(j0, e - 1)), p1 = ldg4(a.posw + min(j1, e - 1));
        auto reach = [&](const float4 p) {
            const float ex = fmaxf(fmaxf(lox - p.x, p.x - hix), 0.0f);
            const float ey = fmaxf(fmaxf(loy - p.y, p.y - hiy), 0.0f);
            const float ez = fmaxf(fmaxf(loz - p.z, p.z - hiz), 0.0f);
            return fmaf(ez, ez, fmaf(ey, ey, ex * ex)) < g.cull_r2;
        };
        const bool keep0 = reach(p0) && j0 < e, keep1 = reach(p1) && j1 < e;
        if (tail + 64u > (unsigned)kListLen) {  // (warp-uniform, every third chunk of a run) make room
            float4 w = p0, wv = p0;
            const bool mine = lane < tail - head;
            if (mine) {
                w = lds4v(lst + ((head + lane) << 4));
                wv = lds4v(lst_v + ((head + lane) << 4));
            }
            __syncwarp();
            if (mine) {
                sts4(lst + (lane << 4), w);
                sts4(lst_v + (lane << 4), wv);
            }
            tail -= head;
            head = 0u;
        }
        const unsigned m0 = __ballot_sync(0xffffffffu, keep0), m1 = __ballot_sync(0xffffffffu, keep1);
        if (keep0) {
            const unsigned at = lst + ((tail + __popc(m0 & below)) << 4);
            sts4(at, p0);
            sts4(at + dv, ldg4(a.veli + j0));
        }
        if (keep1) {
            const unsigned at = lst + ((tail + __popc(m0) + __popc(m1 & below)) << 4);
            sts4(at, p1);
            sts4(at + dv, ldg4(a.veli + j1));
        }
        tail += __popc(m0) + __popc(m1);
        __syncwarp();
        const unsigned stop = tail & ~3u;
        if (stop != head) {
            unsigned at = lst + (head << 4);
            const unsigned end = lst + (stop << 4);
            do {
                walk_group_mixed<LOWFIX>(h, run, tabrow_s, tabrow, boid_acc, a, g, im, at, at + dv);
                at += 64u;
            } while (at != end);
            head = stop;
        }
    }
    if (tail != head) {
        if (lane < 4u - (tail - head)) {
            // (type 0 and zero velocity: the parked particle is out of everybody's range)
            const unsigned at = lst + ((tail + lane) << 4);
            sts4(at, make_float4(kDummyPos, kDummyPos, kDummyPos, 0.0f));
            sts4(at + dv, make_float4(0.f, 0.f, 0.f, 0.f));
        }
        __syncwarp();
        const unsigned at = lst + (head << 4);
        walk_group_mixed<LOWFIX>(h, run, tabrow_s, tabrow, boid_acc, a, g, im, at, at + dv);
    }
    __syncwarp();  // the lists are read before the next run appends to them
    h.ax += ax0; h.ay += ay0; h.az += az0;
}

__device__ __forceinline__ void walk_image_mixed(Home& h, BoidRun& run, RowAddr tabrow_s,
                                                 const PairParam* __restrict__ tabrow, float* __restrict__ boid_acc,
                                                 const ForceArgs& a, const Geometry& g, const Box& b, int sx, int sy,
                                                 int sz, int s, int e, unsigned lst, unsigned lst_v)
{
    const Image im = make_image(h, g, sx, sy, sz);
    if (__any_sync(0xffffffffu, im.up))
        walk_culled_mixed<true>(h, run, tabrow_s, tabrow, boid_acc, a, g, im, b, sx, sy, sz, s, e, lst, lst_v);
    else
        walk_culled_mixed<false>(h, run, tabrow_s, tabrow, boid_acc, a, g, im, b, sx, sy, sz, s, e, lst, lst_v);
}

// Choice of the LOWFIX variant for the lanes that are converged here (every lane of a code
// path walks the same sequence of runs, possibly empty).  The variant does not change a
// lane's result when its own `lo` parts are zero, so the vote only saves instructions.
template <int KIND>
__device__ __forceinline__ void run_image(Home& h, RowAddr tabrow_s,
                                          const PairParam* __restrict__ tabrow,
                                          float* __restrict__ boid_acc, const ForceArgs& a,
                                          const Geometry& g, int sx, int sy, int sz, int s, int e, BoidQueue& q)
{
    const Image im = make_image(h, g, sx, sy, sz);
    if (__any_sync(__activemask(), im.up && e > s))
        run_pairs<KIND, true>(h, tabrow_s, tabrow, boid_acc, a, g, im, s, e, q);
    else
        run_pairs<KIND, false>(h, tabrow_s, tabrow, boid_acc, a, g, im, s, e, q);
}

// The reference's z wrap (acceleration_updater.py:325-328) tests pos_y[i] where it
// means pos_z[i]; which particle plays "i" depends on which of the two bins lists
// the other in the HALF stencil.  In GOF_WRAP_REFERENCE mode warps whose home cells
// touch the bottom or top layer therefore take this literal per-pair path, cell by
// cell (no x merging), with the comparisons in float64 like the reference.
//   direct : the home particle is the reference's thread i, the neighbour its j
//   !direct: the neighbour is thread i and updated the home particle as its j
// z separation (home - neighbour) and the neighbour's relative z the reference accumulates
// for cohesion, by the literal rules above, in fp32: the comparisons use the pre-rounded
// thresholds (exactly the reference's float64 decisions) and every shift subtracts Lz from
// the LARGER coordinate (exact), never adds it to the smaller one.
__device__ __forceinline__ void quirk_dz(const Geometry& g, bool direct, float zi, float yi, float zj,
                                         float yj, float& dz, float& relz)
{
    const bool i_low = zi < g.zq_r, j_low = zj < g.zq_r, i_top = zi > g.zq_top, j_top = zj > g.zq_top;
    const float zi_m = (zi - g.Lz) - g.Lz_lo, zj_m = (zj - g.Lz) - g.Lz_lo;  // one box length down
    if (direct) {
        if (i_low && j_top) dz = zi - zj_m;                   // z2 = zj - Lz
        else if (j_low && yi > g.zq_top) dz = zi_m - zj;     // z2 = zj + Lz   (the y test is the quirk)
        else dz = zi - zj;
        relz = -dz;
    } else {
        if (j_low && i_top) dz = zi_m - zj;                   // the neighbour moved the home: z2 = zi - Lz
        else if (i_low && yj > g.zq_top) dz = zi - zj_m;     // z2 = zi + Lz
        else dz = zi - zj;
        if (i_low && j_top) relz = zj_m - zi;                 // ... and min-imaged ITS z towards the home
        else if (j_low && i_top) relz = zj - zi_m;
        else relz = zj - zi;
    }
}

// float64 squared distance of a pair on the literal path (guard band / edge of r_sep)
__device__ __noinline__ double quirk_exact_d2(const Geometry& g, bool direct, float hx, float hy, float hz,
                                              float pjx, float pjy, float pjz, int sx, int sy)
{
    const double R = g.rmaxd, Lz = g.Lzd, top = g.Lzd - g.rmaxd;
    const double zi = hz, zj = pjz;
    double dzd;
    if (direct) {
        double z2 = zj;
        if (zi < R && z2 > top) z2 -= Lz;
        else if (z2 < R && (double)hy > top) z2 += Lz;
        dzd = zi - z2;
    } else {
        double z2 = zi;
        if (zj < R && z2 > top) z2 -= Lz;
        else if (z2 < R && (double)pjy > top) z2 += Lz;
        dzd = z2 - zj;
    }
    const double x2 = sx ? __dsub_rn((double)pjx, sx * g.Lxd) : (double)pjx;
    const double y2 = sy ? __dsub_rn((double)pjy, sy * g.Lyd) : (double)pjy;
    const double ddx = __dsub_rn((double)hx, x2), ddy = __dsub_rn((double)hy, y2);
    return __dadd_rn(__dadd_rn(__dmul_rn(ddx, ddx), __dmul_rn(ddy, ddy)), __dmul_rn(dzd, dzd));
}

// Cold path of the literal Clusters loop (cf. repair_pair_clusters).
__device__ __noinline__ float3 repair_quirk_clusters(const Geometry& g, bool direct, float hx, float hy,
                                                     float hz, float pjx, float pjy, float pjz, int sx,
                                                     int sy, float dx, float dy, float dz, float d2,
                                                     const PairParam* q)
{
    const bool added = d2 < g.r2_lo;
    const double e = quirk_exact_d2(g, direct, hx, hy, hz, pjx, pjy, pjz, sx, sy);
    const bool wanted = (1e-10 < e) && (e < g.r2d);
    if (added == wanted) return make_float3(0.f, 0.f, 0.f);
    float4 qa = q->a;
    const float4 qb = q->b;
    qa.x = qa.x * qa.x;
    float s = clusters_scale(qa, make_float2(qb.x, qb.y), d2, fast_rsqrt(d2));
    if (added) s = -s;
    return make_float3(-s * dx, -s * dy, -s * dz);
}

// Literal pair of the Clusters kind, branch-free like pair_clusters: force for every candidate,
// masked by the cut-off; returns d2 so that the caller can spot the pairs that need the cold path.
__device__ __forceinline__ float pair_clusters_quirk(Home& h, RowAddr tabrow_s, const Geometry& g,
                                                     const Image& im, bool direct, const float4 pj)
{
    float dz, relz;
    quirk_dz(g, direct, h.z, h.y, pj.z, pj.y, dz, relz);
    const float dx = (im.x - pj.x) + im.lx, dy = (im.y - pj.y) + im.ly;
    const float d2 = dist2_tiny(dx, dy, dz);
    const unsigned at = tabrow_s.a + ((unsigned)__float_as_int(pj.w) << 4);
    const float4 qa = lds4(at);
    const float2 qb = lds2(at + kRowB);
    const float s = clusters_scale(qa, qb, d2, fast_rsqrt(d2));
    if (d2 < g.r2_lo) {
        h.ax = fmaf(-s, dx, h.ax);
        h.ay = fmaf(-s, dy, h.ay);
        h.az = fmaf(-s, dz, h.az);
    }
    return d2;
}

// its cold path; like `repair`, the corrections of a run are summed apart (`rep`)
__device__ __forceinline__ void repair_quirk(float3& rep, const Home& h, const PairParam* __restrict__ tabrow,
                                             const Geometry& g, const Image& im, int sx, int sy, bool direct,
                                             const float4 pj)
{
    float dz, relz;
    quirk_dz(g, direct, h.z, h.y, pj.z, pj.y, dz, relz);
    const float dx = (im.x - pj.x) + im.lx, dy = (im.y - pj.y) + im.ly;
    const float d2 = dist2_tiny(dx, dy, dz);
    const float3 f = repair_quirk_clusters(g, direct, h.x, h.y, h.z, pj.x, pj.y, pj.z, sx, sy, dx, dy, dz, d2,
                                           tabrow + __float_as_int(pj.w));
    rep.x += f.x; rep.y += f.y; rep.z += f.z;
}

template <bool BOIDS>
__device__ __forceinline__ void quirk_pair(Home& h, float3& rep, RowAddr tabrow_s,
                                           const PairParam* __restrict__ tabrow,
                                           float* __restrict__ boid_acc, const ForceArgs& a,
                                           const Geometry& g, const Image& im, int sx, int sy, bool direct,
                                           int j, const float4 pj)
{
    if (!BOIDS) {
        if (needs_repair(g, pair_clusters_quirk(h, tabrow_s, g, im, direct, pj)))
            repair_quirk(rep, h, tabrow, g, im, sx, sy, direct, pj);
        return;
    }
    float dz, relz;
    quirk_dz(g, direct, h.z, h.y, pj.z, pj.y, dz, relz);
    const float dx = (im.x - pj.x) + im.lx, dy = (im.y - pj.y) + im.ly;
    const float d2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
    if (d2 < g.r2_hi && d2 > 1e-10f) {
        const float hx = h.x, hy = h.y, hz = h.z;
        auto exact = [&]() { return quirk_exact_d2(g, direct, hx, hy, hz, pj.x, pj.y, pj.z, sx, sy); };
        bool ok = true;
        if (d2 > g.r2_lo) {
            const double d2d = exact();
            ok = (1e-10 < d2d) && (d2d < g.r2d);
        }
        if (ok)
            interact<BOIDS>(h, tabrow, boid_acc, a, a.T, j, __float_as_int(pj.w), dx, dy, dz, -dx, -dy, relz,
                            d2, exact);
    }
}

template <bool BOIDS>
__device__ __forceinline__ void run_pairs_zquirk(Home& h, float3& rep, RowAddr tabrow_s,
                                              const PairParam* __restrict__ tabrow,
                                              float* __restrict__ boid_acc, const ForceArgs& a,
                                              const Geometry& g, int s, int e, int sx, int sy,
                                              bool direct)
{
    const Image im = make_image(h, g, sx, sy, 0);
    int j = s;
    for (; j + 4 <= e; j += 4) {
        const float4 p0 = ldg4(a.posw + j), p1 = ldg4(a.posw + j + 1), p2 = ldg4(a.posw + j + 2),
                     p3 = ldg4(a.posw + j + 3);
        quirk_pair<BOIDS>(h, rep, tabrow_s, tabrow, boid_acc, a, g, im, sx, sy, direct, j, p0);
        quirk_pair<BOIDS>(h, rep, tabrow_s, tabrow, boid_acc, a, g, im, sx, sy, direct, j + 1, p1);
        quirk_pair<BOIDS>(h, rep, tabrow_s, tabrow, boid_acc, a, g, im, sx, sy, direct, j + 2, p2);
        quirk_pair<BOIDS>(h, rep, tabrow_s, tabrow, boid_acc, a, g, im, sx, sy, direct, j + 3, p3);
    }
    for (; j < e; ++j)
        quirk_pair<BOIDS>(h, rep, tabrow_s, tabrow, boid_acc, a, g, im, sx, sy, direct, j, ldg4(a.posw + j));
}

// Split kernel: run q's partial sum of this warp's home particle, [run][component][home lane].
constexpr int kSplitHomes = 32;
constexpr int kSplitRuns = 18;  // 9 stencil rows x (un-shifted x cells, cell across the x boundary)
__device__ __forceinline__ void store_partial(float* __restrict__ part, int q, const Home& h)
{
    part[(q * 3 + 0) * kSplitHomes] = h.ax;
    part[(q * 3 + 1) * kSplitHomes] = h.ay;
    part[(q * 3 + 2) * kSplitHomes] = h.az;
}

// The whole stencil of ONE home particle on the literal path, out of line: the fast path of the
// kernel keeps its own register allocation, and an edge-layer lane pays one call.
// Same run structure and partial sums as the fast path (per row: the un-shifted x cells in
// ascending order, then the cell across the x boundary), so that a particle's result does not
// depend on which path its warp-mates take.
//
// SPLIT > 1 (small systems, see k_force): only the stencil rows of this warp's share are walked and
// every run's partial sum goes to `part` (shared memory) instead of being folded here.
template <bool BOIDS, int SPLIT>
__device__ __noinline__ float3 quirk_home(float hx, float hy, float hz, int type, unsigned kinds,
                                          RowAddr tabrow_s, const PairParam* __restrict__ tabrow,
                                          float* __restrict__ boid_acc, const ForceArgs& a,
                                          const Geometry& g, int cx, int cy, int lz, int gz, bool valid,
                                          int share, int rows_per_share, float* __restrict__ part)
{
    Home h;
    h.x = hx; h.y = hy; h.z = hz;
    h.vx = h.vy = h.vz = 0.0f;
    h.ax = h.ay = h.az = 0.0f;
    h.type = type;
    h.kinds = kinds;
    const int dz_lo = g.is3d ? -1 : 0, dz_hi = g.is3d ? 1 : 0;
    int row = 0;
    for (int dzi = dz_lo; dzi <= dz_hi; ++dzi) {
        int nl = lz + dzi;
        const int ngz = gz + dzi;
        if (g.index_wrap_z) {
            if (ngz < 0) nl += g.nbz;
            else if (ngz >= g.nbz) nl -= g.nbz;
        }
        for (int dyi = -1; dyi <= 1; ++dyi, ++row) {
            if (SPLIT > 1 && row / rows_per_share != share) continue;
            int ny = cy + dyi, sy = 0;
            if (ny < 0) { ny += g.nby; sy = 1; }
            else if (ny >= g.nby) { ny -= g.nby; sy = -1; }
            const int rb = (nl * g.nby + ny) * g.nbx;
            const int lo = cx > 0 ? cx - 1 : 0;
            const int hi = cx < g.nbx - 1 ? cx + 1 : g.nbx - 1;
            float ax0 = h.ax, ay0 = h.ay, az0 = h.az;
            h.ax = h.ay = h.az = 0.0f;
            float3 rep = make_float3(0.0f, 0.0f, 0.0f);  // guard-band corrections of the run (Clusters kind)
            // Half stencil of set_bin_neighbours (acceleration_updater.py:140-215): the home particle
            // is the reference's thread i ("direct") for the upper layer, for the row above in its own
            // layer, and for its own cell and the next one in x; elsewhere the neighbour is.  Only the
            // home row (dz = dy = 0) mixes the two, so every other row is one contiguous run.
            if (dzi == 0 && dyi == 0) {
                if (lo < cx) {
                    const int s = a.pstart[rb + lo];
                    const int e = valid ? s + a.count[rb + lo] : s;
                    run_pairs_zquirk<BOIDS>(h, rep, tabrow_s, tabrow, boid_acc, a, g, s, e, 0, sy, false);
                }
                const int s = a.pstart[rb + cx];
                const int e = valid ? a.pstart[rb + hi] + a.count[rb + hi] : s;
                run_pairs_zquirk<BOIDS>(h, rep, tabrow_s, tabrow, boid_acc, a, g, s, e, 0, sy, true);
            } else {
                const bool direct = dzi == 1 || (dzi == 0 && dyi == 1);
                const int s = a.pstart[rb + lo];
                const int e = valid ? a.pstart[rb + hi] + a.count[rb + hi] : s;
                run_pairs_zquirk<BOIDS>(h, rep, tabrow_s, tabrow, boid_acc, a, g, s, e, 0, sy, direct);
            }
            if (!BOIDS) { h.ax += rep.x; h.ay += rep.y; h.az += rep.z; }
            if (SPLIT > 1) store_partial(part, 2 * row, h);
            h.ax += ax0; h.ay += ay0; h.az += az0;
            ax0 = h.ax; ay0 = h.ay; az0 = h.az;
            h.ax = h.ay = h.az = 0.0f;
            rep = make_float3(0.0f, 0.0f, 0.0f);
            if (cx == 0 || cx == g.nbx - 1) {
                const int wc = cx == 0 ? g.nbx - 1 : 0;
                const int dxi = cx == 0 ? -1 : 1;
                const bool direct = dzi == 1 || (dzi == 0 && (dyi == 1 || (dyi == 0 && dxi >= 0)));
                const int s = a.pstart[rb + wc];
                const int e = valid ? s + a.count[rb + wc] : s;
                run_pairs_zquirk<BOIDS>(h, rep, tabrow_s, tabrow, boid_acc, a, g, s, e, cx == 0 ? 1 : -1, sy, direct);
            }
            if (!BOIDS) { h.ax += rep.x; h.ay += rep.y; h.az += rep.z; }
            if (SPLIT > 1) store_partial(part, 2 * row + 1, h);
            h.ax += ax0; h.ay += ay0; h.az += az0;
        }
    }
    return make_float3(h.ax, h.ay, h.az);
}

// The whole-warp walk of a warp whose home cells are in the bottom / top layer in
// GOF_WRAP_REFERENCE mode (Clusters kind): rows and runs as in k_force's walk, pairs by the
// literal rules.  The role of a pair -- is the home particle the reference's thread i
// ("direct")? -- follows from the stencil row, and in the home row from the candidate's cell
// against the lane's own (see quirk_home): direct <=> the candidate's slot >= `own`.  Out of line:
// the common walk keeps its register allocation.
template <int SPLIT>
__device__ __noinline__ float3 quirk_walk(float hx, float hy, float hz, int type, RowAddr tabrow_s,
                                          const PairParam* __restrict__ tabrow, const ForceArgs& a,
                                          const Geometry& g, float lox, float loy, float loz, float hix,
                                          float hiy, float hiz, int cx, int cy, int lz, int gz, int cx_min,
                                          int cx_max, unsigned lst, unsigned lst_j, int share, int rows_per_share,
                                          float* __restrict__ part)
{
    Home h;
    h.x = hx; h.y = hy; h.z = hz;
    h.vx = h.vy = h.vz = 0.0f;
    h.ax = h.ay = h.az = 0.0f;
    h.type = type;
    h.kinds = 0u;
    Box box;
    box.lox = lox; box.loy = loy; box.loz = loz; box.hix = hix; box.hiy = hiy; box.hiz = hiz;
    const int all = INT_MIN, none = INT_MAX;
    int row = 0;
    for (int dzi = -1; dzi <= 1; ++dzi) {
        int nl = lz + dzi;
        const int ngz = gz + dzi;
        if (g.index_wrap_z) {
            if (ngz < 0) nl += g.nbz;
            else if (ngz >= g.nbz) nl -= g.nbz;
        }
        for (int dyi = -1; dyi <= 1; ++dyi, ++row) {
            if (SPLIT > 1 && row / rows_per_share != share) continue;
            int ny = cy + dyi, sy = 0;
            if (ny < 0) { ny += g.nby; sy = 1; }
            else if (ny >= g.nby) { ny -= g.nby; sy = -1; }
            const int rb = (nl * g.nby + ny) * g.nbx;
            const int lo = cx_min > 0 ? cx_min - 1 : 0;
            const int hi = cx_max < g.nbx - 1 ? cx_max + 1 : g.nbx - 1;
            const bool row_direct = dzi == 1 || (dzi == 0 && dyi == 1);
            const bool home_row = dzi == 0 && dyi == 0;
            {
                const Image im = make_image(h, g, 0, sy, 0);
                walk_culled<true, true>(h, tabrow_s, tabrow, a, g, im, box, 0, sy, 0, a.pstart[rb + lo],
                                        a.pstart[rb + hi] + a.count[rb + hi], lst, lst_j,
                                        home_row ? a.pstart[rb + cx] : (row_direct ? all : none));
            }
            if (SPLIT > 1) {
                store_partial(part, 2 * row, h);
                h.ax = h.ay = h.az = 0.0f;
            }
            const float ax0 = h.ax, ay0 = h.ay, az0 = h.az;
            h.ax = h.ay = h.az = 0.0f;
            for (int side = 0; side < 2; ++side) {
                // side 0: the home's x - 1 neighbour across the boundary; side 1: its x + 1 neighbour
                if (side == 0 ? cx_min != 0 : cx_max != g.nbx - 1) continue;
                const int wc = side == 0 ? g.nbx - 1 : 0, sx = side == 0 ? 1 : -1;
                const bool direct = row_direct || (home_row && side == 1);
                const Image im = make_image(h, g, sx, sy, 0);
                const int s = a.pstart[rb + wc];
                walk_culled<true, true>(h, tabrow_s, tabrow, a, g, im, box, sx, sy, 0, s, s + a.count[rb + wc], lst,
                                        lst_j, direct ? all : none);
            }
            if (SPLIT > 1) {
                store_partial(part, 2 * row + 1, h);
                h.ax = h.ay = h.az = 0.0f;
            } else {
                h.ax += ax0; h.ay += ay0; h.az += az0;
            }
        }
    }
    return make_float3(h.ax, h.ay, h.az);
}

//
// SPLIT > 1 is the small-system variant (Clusters kind only).  With one thread per particle a
// system of a few thousand particles cannot fill 148 SMs, and a step costs the latency of one
// thread walking ~1000 candidates.  Here a block is 32 home particles x SPLIT warps; warp w walks
// only the stencil rows of share w (one row, or one dz plane) and leaves every run's partial sum
// in shared memory; warp 0 then folds the 18 partials in the order the one-thread kernel adds
// them, so the result is bit-identical to SPLIT == 1 whatever the split.
#ifndef GOF_FORCE_MINBLOCKS
#define GOF_FORCE_MINBLOCKS 4
#endif
template <int KIND, int MODE, int SPLIT>
#ifndef GOF_MIXED_MINBLOCKS
#define GOF_MIXED_MINBLOCKS 4
#endif
#ifndef GOF_BOIDS_MINBLOCKS
#define GOF_BOIDS_MINBLOCKS 6
#endif
__global__ void __launch_bounds__(SPLIT == 1 ? kForceThreads : kSplitHomes * SPLIT,
                                  SPLIT == 1 ? (KIND == kKindBoids ? GOF_BOIDS_MINBLOCKS
                                                : (KIND == kKindMixed ? GOF_MIXED_MINBLOCKS : GOF_FORCE_MINBLOCKS))
                                             : (SPLIT == 3 ? 6 : 2))
k_force(const __grid_constant__ Geometry g, const __grid_constant__ ForceArgs a)
{
    constexpr bool BOIDS = KIND != kKindClusters;
    static_assert(SPLIT == 1 || !BOIDS, "the split variant covers Clusters-kind systems only");
    constexpr int kHomes = SPLIT == 1 ? kForceThreads : kSplitHomes;  // home particles per block
    const int hl = SPLIT == 1 ? (int)threadIdx.x : (int)(threadIdx.x & 31u);
    const int share = SPLIT == 1 ? 0 : (int)(threadIdx.x >> 5);
    const int rows_per_share = SPLIT == 1 ? 9 : (g.is3d ? 9 : 3) / SPLIT;
    extern __shared__ float4 smem_raw[];
    PairParam* tab = reinterpret_cast<PairParam*>(smem_raw);
    const int T = a.T;
    char* tabAB = reinterpret_cast<char*>(tab + T * T);  // T rows of kRowBytes (see RowAddr)
    char* behind = tabAB + T * kRowBytes;
    float* boid_acc = reinterpret_cast<float*>(behind) + threadIdx.x * 4;  // this thread's first float4 (boid_vec)
    float* part = reinterpret_cast<float*>(behind) + hl;  // SPLIT > 1 only
    // Clusters kind, behind `part`: every warp's list of surviving candidates (walk_culled) and,
    // behind all the lists, those of their slots (literal z-wrap walk only)
    const unsigned wmem = (unsigned)__cvta_generic_to_shared(behind) +
                          (SPLIT > 1 ? kSplitRuns * 3 * kSplitHomes * 4u : 0u);
    const unsigned lst = wmem + (threadIdx.x >> 5) * (kWalkCap * 16u);
    const unsigned lst_j = wmem + (blockDim.x >> 5) * (kWalkCap * 16u) + (threadIdx.x >> 5) * (kWalkCap * 4u);
    // mixed kind: every warp's lists of surviving positions and velocities, behind the per-type sums
    const unsigned mmem = (unsigned)__cvta_generic_to_shared(behind) + (unsigned)T * kBoidSlots * kForceThreads * 4u;
    const unsigned mlst = mmem + (threadIdx.x >> 5) * (kListLen * 32u), mlst_v = mlst + kListLen * 16u;
    // The home range.  Slab launches whose ranges live on the device are sized for an upper bound:
    // a surplus block leaves here, before it has touched shared memory or a barrier.
    int home_begin = a.home_begin, home_count = a.home_count;
    int home_begin2 = a.home_begin2, home_count2 = a.home_count2;
    if (a.interior_words) {  // owned particles that are in neither boundary layer
        const int live = a.interior_words[0], lo = a.interior_words[1], hi = a.interior_words[2];
        home_begin = lo;
        home_count = live - lo - hi;
        if ((int)blockIdx.x * kHomes >= home_count) return;  // (block-uniform; also: no interior layer)
    }
    if (a.boundary_words) {  // the bottom and the top owned layer
        const int live = a.boundary_words[0], lo = a.boundary_words[1], hi = a.boundary_words[2];
        home_begin = 0;
        home_count = lo;
        home_begin2 = live - hi;
        home_count2 = hi;
        if ((int)blockIdx.x * kHomes >= lo + hi) return;
    }
    for (int q = threadIdx.x; q < T * T; q += blockDim.x) {
        const PairParam pp = a.table[q];
        tab[q] = pp;
        const int r = q / T, c = q - r * T;
        float4 ea = pp.a, eb = make_float4(pp.b.x, pp.b.y, 0.0f, 0.0f);
        ea.x = pp.a.x * pp.a.x;  // clusters_scale compares squared distances
        if (KIND == kKindMixed && ((a.kind_rows[r] >> c) & 1u)) {  // Boids-kind pair: see pair_mixed
            ea = make_float4(pp.a.x * pp.a.x, 0.0f, -pp.a.y, 0.0f);
            eb = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        }
        *reinterpret_cast<float4*>(tabAB + r * kRowBytes + c * 16) = ea;
        *reinterpret_cast<float4*>(tabAB + r * kRowBytes + kRowB + c * 16) = eb;
    }
    BoidQueue queue;
    queue.base = 0u;
    queue.next = 0u;
    queue.fx = queue.fy = queue.fz = 0.0f;
    if (BOIDS) {
        for (int q = 0; q < T; ++q) {
            *boid_vec(boid_acc, q, 0) = make_float4(0.f, 0.f, 0.f, 0.f);
            *boid_vec(boid_acc, q, 1) = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        // the queue sits behind the per-type sums; entry k of this lane at base + k * 512
        queue.base = (unsigned)__cvta_generic_to_shared(reinterpret_cast<float*>(behind) + T * kBoidSlots * kForceThreads +
                                                        threadIdx.x);
        queue.next = queue.base;
    }
    __syncthreads();

    // Block -> particles mapping, rotated so that the blocks of the TOP z layer run first, then the
    // bottom layer, then the rest: in GOF_WRAP_REFERENCE mode those two layers take the literal path
    // (~3x the work per particle) and would otherwise sit at the very end of the grid, stretching
    // the kernel's tail.  (Single periodic box only: a slab launches its boundary layers apart.)
    // Block -> particles mapping.  In GOF_WRAP_REFERENCE mode the blocks of the top and the bottom z
    // layer take the literal z-wrap paths and run ~2.5x as long as the others (far more than their
    // instruction count: a second hot loop on an SM costs instruction-cache misses for everybody).
    // At the end of the grid they would stretch the kernel's tail (1.12 ms instead of 0.97 at
    // N = 1M); spread evenly over the launch they slow every SM down (1.23 ms).  They are launched
    // first, top layer then bottom layer.  (Single periodic box only: a slab launches its boundary
    // layers apart.)
    int blk = (int)blockIdx.x;
#ifndef GOF_NO_ROTATE
    if (g.is3d && g.index_wrap_z && a.wrap_mode == kWrapReference && home_count2 == 0 && !a.interior_words) {
        const int grid = (int)gridDim.x;
        const int top0 = (a.pstart[(g.nlz - 1) * g.nbx * g.nby] - home_begin) / kHomes;  // first block of the top layer
        const int n_top = grid - top0;
        int n_bot = (a.pstart[g.nbx * g.nby] - home_begin + kHomes - 1) / kHomes;        // blocks touching the bottom layer
        if (n_bot > top0) n_bot = top0;
        const int nq = n_top + n_bot;  // literal blocks
        // position r in the order "top layer, bottom layer, the rest" for launch index p
        const int p = (int)blockIdx.x;
        int r = p;
        // The first wave (kWaveSMs x kWaveBlocks launch indices, handed out round-robin over the SMs)
        // gives the literal blocks to the first W SMs only, so that the two hot loops do not share an SM.
        constexpr int kWaveBlocks = SPLIT == 1 ? GOF_FORCE_MINBLOCKS : 2;
        const int kWaveSMs = a.wave_sms;
        const int W = (nq + kWaveBlocks - 1) / kWaveBlocks;
        if (W <= kWaveSMs && grid >= kWaveSMs * kWaveBlocks) {
            if (p < kWaveSMs * kWaveBlocks) {
                const int k = p / kWaveSMs, c = p - k * kWaveSMs;
                int before = 0;  // literal positions before p
                for (int kk = 0; kk < k; ++kk) {
                    const int left = nq - kk * W;
                    before += left < 0 ? 0 : (left < W ? left : W);
                }
                const int left = nq - k * W;
                const int wk = left < 0 ? 0 : (left < W ? left : W);
                if (c < wk) r = k * W + c;
                else r = nq + p - (before + wk);
            }  // else: every literal block has been placed, r = p
        }
        blk = r < n_top ? top0 + r : r - n_top;
    }
#endif
    const int local = blk * kHomes + hl;
    if (a.interior_words && local - hl >= home_count) return;  // whole block past the range
    if (a.boundary_words && local - hl >= home_count + home_count2) return;
    const bool valid = local < home_count + home_count2;
    int k = home_begin + (local < home_count ? local : home_count - 1);
    if (local >= home_count && home_count2 > 0)
        k = home_begin2 + (valid ? local - home_count : home_count2 - 1);

    const float4 pi = a.posw[k];
    const float4 vi = a.veli[k];
    Home h;
    h.x = pi.x; h.y = pi.y; h.z = pi.z;
    h.vx = vi.x; h.vy = vi.y; h.vz = vi.z;
    h.ax = h.ay = h.az = 0.0f;
    h.type = __float_as_int(pi.w);
    h.kinds = BOIDS ? a.kind_rows[h.type] : 0u;
    const PairParam* tabrow = tab + h.type * T;
    RowAddr tabrow_s;
    tabrow_s.a = (unsigned)__cvta_generic_to_shared(tabAB + h.type * kRowBytes);

    const int c = a.cell_sorted[k];
    const int cx = c % g.nbx;
    const int cyz = c / g.nbx;
    const int cy = cyz % g.nby;
    const int lz = cyz / g.nby;
    const int gz = lz + g.layer_base;  // global layer

    // Home cells on the bottom / top layer take the literal z-wrap path.  The choice is made
    // PER LANE (a warp that straddles the layer boundary runs both paths one after the other):
    // what a particle computes must not depend on its warp-mates, or the result would change
    // with the decomposition.  The votes inside the fast path use the active mask.
    //
    // Which home cells really need it.  With the home particle or the neighbour as the reference's
    // thread i, the literal rules (acceleration_updater.py:325-328) come to this for nbz >= 4: a pair
    // ACROSS the periodic z boundary interacts only if its upper particle has y > Lz - r_max (the
    // `pos_y[i]` of the quirk), and a pair on one side of it is lost if the upper-role particle has
    // such a y and the other one z < r_max.  Away from that band of y nothing else is left of the
    // rules than "no interaction across the z boundary".  So only the home cells whose own row of
    // cells, or one of the two next to it in y, reaches into the band take the literal paths; the
    // other cells of the bottom / top layer walk like interior cells and skip the three stencil rows
    // across the boundary (`skip_dz`).  The choice depends on the home cell alone.
    bool quirk = g.is3d && a.wrap_mode == kWrapReference && (gz == 0 || gz == g.nbz - 1);
    int skip_dz = 2;  // stencil dz to leave out (2: none)
    if (quirk && g.nbz >= 4) {
        const double band = g.Lzd - g.rmaxd - 1e-6 * g.Lzd;  // (a little below the threshold of the test)
        const int rm = cy > 0 ? cy - 1 : g.nby - 1, rp = cy < g.nby - 1 ? cy + 1 : 0;
        const bool in_band = (double)(cy + 1) * g.bsy > band || (double)(rm + 1) * g.bsy > band ||
                             (double)(rp + 1) * g.bsy > band;
        if (!in_band) {
            quirk = false;
            skip_dz = gz == 0 ? -1 : 1;
        }
    }

    const int dz_lo = g.is3d ? -1 : 0, dz_hi = g.is3d ? 1 : 0;
    // Clusters kind: the whole warp walks as one when its home particles share a row of cells (all
    // lanes are converged here, padding lanes included: they repeat the last particle and store
    // nothing).
    bool as_warp = false;
#ifndef GOF_NO_WARP_WALK
    if (KIND == kKindClusters || KIND == kKindMixed)  // (one row of cells: one layer, so `quirk` is the same for all lanes)
        as_warp = __reduce_min_sync(0xffffffffu, (unsigned)cyz) == __reduce_max_sync(0xffffffffu, (unsigned)cyz);
    if (KIND == kKindMixed && quirk) as_warp = false;  // the literal z-wrap layers of a mixed system walk per lane
#endif
    if (KIND == kKindMixed && as_warp) {
        // mixed kind: one walk for the warp, unified branch-free pairs (walk_culled_mixed)
        const Box box = warp_box(h);
        const int cx_min = (int)__reduce_min_sync(0xffffffffu, (unsigned)cx);
        const int cx_max = (int)__reduce_max_sync(0xffffffffu, (unsigned)cx);
        BoidRun run;
        run.c = run.rx = run.ry = run.rz = run.vx = run.vy = run.vz = 0.0f;
        run.t = -1;
        for (int dzi = dz_lo; dzi <= dz_hi; ++dzi) {
            int nl = lz + dzi, sz = 0;
            if (g.is3d) {
                const int ngz = gz + dzi;
                if (ngz < 0) { sz = 1; if (g.index_wrap_z) nl += g.nbz; }
                else if (ngz >= g.nbz) { sz = -1; if (g.index_wrap_z) nl -= g.nbz; }
            }
            if (dzi == skip_dz) continue;  // across the z boundary, reference wrap: no interaction
            for (int dyi = -1; dyi <= 1; ++dyi) {
                int ny = cy + dyi, sy = 0;
                if (ny < 0) { ny += g.nby; sy = 1; }
                else if (ny >= g.nby) { ny -= g.nby; sy = -1; }
                const int rb = (nl * g.nby + ny) * g.nbx;
                const int lo = cx_min > 0 ? cx_min - 1 : 0;
                const int hi = cx_max < g.nbx - 1 ? cx_max + 1 : g.nbx - 1;
                walk_image_mixed(h, run, tabrow_s, tabrow, boid_acc, a, g, box, 0, sy, sz, a.pstart[rb + lo],
                                 a.pstart[rb + hi] + a.count[rb + hi], mlst, mlst_v);
                const float ax0 = h.ax, ay0 = h.ay, az0 = h.az;
                h.ax = h.ay = h.az = 0.0f;
                if (cx_min == 0) {
                    const int s = a.pstart[rb + g.nbx - 1];
                    walk_image_mixed(h, run, tabrow_s, tabrow, boid_acc, a, g, box, 1, sy, sz, s,
                                     s + a.count[rb + g.nbx - 1], mlst, mlst_v);
                }
                if (cx_max == g.nbx - 1) {
                    const int s = a.pstart[rb];
                    walk_image_mixed(h, run, tabrow_s, tabrow, boid_acc, a, g, box, -1, sy, sz, s, s + a.count[rb], mlst,
                                     mlst_v);
                }
                h.ax += ax0; h.ay += ay0; h.az += az0;
            }
        }
        boidrun_store(run, boid_acc);  // the epilogue reads the slots
    } else if (as_warp && quirk) {
        // bottom / top layer in GOF_WRAP_REFERENCE mode: the same walk by the literal rules, out of line
        const Box box = warp_box(h);
        const int cx_min = (int)__reduce_min_sync(0xffffffffu, (unsigned)cx);
        const int cx_max = (int)__reduce_max_sync(0xffffffffu, (unsigned)cx);
        const float3 f = quirk_walk<SPLIT>(h.x, h.y, h.z, h.type, tabrow_s, tabrow, a, g, box.lox, box.loy, box.loz,
                                           box.hix, box.hiy, box.hiz, cx, cy, lz, gz, cx_min, cx_max, lst, lst_j,
                                           share, rows_per_share, part);
        h.ax = f.x; h.ay = f.y; h.az = f.z;
    } else if (as_warp) {
        const Box box = warp_box(h);
        const int cx_min = (int)__reduce_min_sync(0xffffffffu, (unsigned)cx);
        const int cx_max = (int)__reduce_max_sync(0xffffffffu, (unsigned)cx);
        int row = 0;
        for (int dzi = dz_lo; dzi <= dz_hi; ++dzi) {
            int nl = lz + dzi, sz = 0;
            if (g.is3d) {
                const int ngz = gz + dzi;
                if (ngz < 0) { sz = 1; if (g.index_wrap_z) nl += g.nbz; }
                else if (ngz >= g.nbz) { sz = -1; if (g.index_wrap_z) nl -= g.nbz; }
            }
            for (int dyi = -1; dyi <= 1; ++dyi, ++row) {
                if (SPLIT > 1 && row / rows_per_share != share) continue;
                if (dzi == skip_dz) {  // across the z boundary, reference wrap: no interaction
                    if (SPLIT > 1) {
                        store_partial(part, 2 * row, h);
                        store_partial(part, 2 * row + 1, h);
                    }
                    continue;
                }
                int ny = cy + dyi, sy = 0;
                if (ny < 0) { ny += g.nby; sy = 1; }
                else if (ny >= g.nby) { ny -= g.nby; sy = -1; }
                const int rb = (nl * g.nby + ny) * g.nbx;
                // the union of the lanes' un-shifted runs: cells cx_min - 1 .. cx_max + 1 of the row
                const int lo = cx_min > 0 ? cx_min - 1 : 0;
                const int hi = cx_max < g.nbx - 1 ? cx_max + 1 : g.nbx - 1;
                walk_image(h, tabrow_s, tabrow, a, g, box, 0, sy, sz, a.pstart[rb + lo],
                           a.pstart[rb + hi] + a.count[rb + hi], lst);
                if (SPLIT > 1) {
                    store_partial(part, 2 * row, h);
                    h.ax = h.ay = h.az = 0.0f;
                }
                // the cells across the periodic x boundary, for the lanes in the first / last cell of
                // the row (the shifted image of any other lane is more than a bin away from them; a
                // lane gets a non-zero sum from at most one of the two)
                const float ax0 = h.ax, ay0 = h.ay, az0 = h.az;
                h.ax = h.ay = h.az = 0.0f;
                if (cx_min == 0) {
                    const int s = a.pstart[rb + g.nbx - 1];
                    walk_image(h, tabrow_s, tabrow, a, g, box, 1, sy, sz, s, s + a.count[rb + g.nbx - 1], lst);
                }
                if (cx_max == g.nbx - 1) {
                    const int s = a.pstart[rb];
                    walk_image(h, tabrow_s, tabrow, a, g, box, -1, sy, sz, s, s + a.count[rb], lst);
                }
                if (SPLIT > 1) {
                    store_partial(part, 2 * row + 1, h);
                    h.ax = h.ay = h.az = 0.0f;
                } else {
                    h.ax += ax0; h.ay += ay0; h.az += az0;
                }
            }
        }
    } else if (!quirk) {
        int row = 0;
        for (int dzi = dz_lo; dzi <= dz_hi; ++dzi) {
            int nl = lz + dzi, sz = 0;
            if (g.is3d) {
                const int ngz = gz + dzi;
                if (ngz < 0) { sz = 1; if (g.index_wrap_z) nl += g.nbz; }
                else if (ngz >= g.nbz) { sz = -1; if (g.index_wrap_z) nl -= g.nbz; }
            }
            for (int dyi = -1; dyi <= 1; ++dyi, ++row) {
                if (SPLIT > 1 && row / rows_per_share != share) continue;
                if (dzi == skip_dz) {  // across the z boundary, reference wrap: no interaction
                    if (SPLIT > 1) {
                        store_partial(part, 2 * row, h);
                        store_partial(part, 2 * row + 1, h);
                    }
                    continue;
                }
                int ny = cy + dyi, sy = 0;
                if (ny < 0) { ny += g.nby; sy = 1; }
                else if (ny >= g.nby) { ny -= g.nby; sy = -1; }
                const int rb = (nl * g.nby + ny) * g.nbx;
                // the x-adjacent cells that need no shift are one contiguous run
                const int lo = cx > 0 ? cx - 1 : 0;
                const int hi = cx < g.nbx - 1 ? cx + 1 : g.nbx - 1;
                int s = a.pstart[rb + lo];
                int e = a.pstart[rb + hi] + a.count[rb + hi];
                if (!valid) e = s;
                run_image<KIND>(h, tabrow_s, tabrow, boid_acc, a, g, 0, sy, sz, s, e, queue);
                if (SPLIT > 1) {
                    store_partial(part, 2 * row, h);
                    h.ax = h.ay = h.az = 0.0f;
                }
                // the cell across the periodic x boundary (an empty run for interior home cells:
                // every lane makes the same calls, so the warp votes inside stay convergent)
                const bool xedge = cx == 0 || cx == g.nbx - 1;
                if (__any_sync(__activemask(), valid && xedge)) {  // (an empty run leaves h as it is: skip it)
                    const int wc = cx == 0 ? g.nbx - 1 : 0;
                    s = a.pstart[rb + wc];
                    e = (valid && xedge) ? s + a.count[rb + wc] : s;
                    run_image<KIND>(h, tabrow_s, tabrow, boid_acc, a, g, cx == 0 ? 1 : -1, sy, sz, s, e, queue);
                }
                if (SPLIT > 1) {
                    store_partial(part, 2 * row + 1, h);
                    h.ax = h.ay = h.az = 0.0f;
                }
            }
        }
        if (KIND == kKindBoids) {  // whatever is still queued; all forces of this path come from the queue
            drain(queue, h, tabrow, boid_acc, a, g);
            h.ax = queue.fx; h.ay = queue.fy; h.az = queue.fz;
        }
    } else {
        const float3 f = quirk_home<BOIDS, SPLIT>(h.x, h.y, h.z, h.type, h.kinds, tabrow_s, tabrow, boid_acc, a, g,
                                                  cx, cy, lz, gz, valid, share, rows_per_share, part);
        h.ax = f.x; h.ay = f.y; h.az = f.z;
    }
    if (SPLIT > 1) {
        // fold the run partials exactly as the one-thread kernel does: total = partial + total
        __syncthreads();
        if (share != 0) return;
        const int runs = g.is3d ? kSplitRuns : kSplitRuns / 3;
        float tx = 0.0f, ty = 0.0f, tz = 0.0f;
        for (int q = 0; q < runs; ++q) {
            tx = part[(q * 3 + 0) * kSplitHomes] + tx;
            ty = part[(q * 3 + 1) * kSplitHomes] + ty;
            tz = part[(q * 3 + 2) * kSplitHomes] + tz;
        }
        h.ax = tx; h.ay = ty; h.az = tz;
    }

    // ---- cohesion and alignment from the per-type means (acceleration_updater.py:422-458) ----
    const int id = __float_as_int(vi.w);
    if (BOIDS) {
        const float spd = sqrtf(h.vx * h.vx + h.vy * h.vy + h.vz * h.vz);
        for (int t = 0; t < T; ++t) {
            const float4 sc = *boid_vec(boid_acc, t, 0), sv = *boid_vec(boid_acc, t, 1);
            const float cnt = sc.x;
            if (MODE == kModeScatter && a.boid_counts && valid) {
                const size_t o = (size_t)t * a.n_total + id;
                a.boid_counts[o] = (int)cnt;
                a.boid_acc_x[o] = fmaf(cnt, h.x, sc.y);
                a.boid_acc_y[o] = fmaf(cnt, h.y, sc.z);
                a.boid_acc_z[o] = fmaf(cnt, h.z, sc.w);
                a.boid_vel_x[o] = sv.x;
                a.boid_vel_y[o] = sv.y;
                a.boid_vel_z[o] = sv.z;
            }
            if (((h.kinds >> t) & 1u) && cnt > 0.0f) {
                const PairParam q = tab[h.type * T + t];
                const float ic = 1.0f / cnt;
                const float mx = sv.x * ic, my = sv.y * ic, mz = sv.z * ic;
                const float proj = spd > 10.0f ? (mx * h.vx + my * h.vy + mz * h.vz) / spd : 0.0f;
                h.ax += q.a.w * (sc.y * ic) + q.a.z * (mx - proj * h.vx);
                h.ay += q.a.w * (sc.z * ic) + q.a.z * (my - proj * h.vy);
                h.az += q.a.w * (sc.w * ic) + q.a.z * (mz - proj * h.vz);
            }
        }
    }

    if (MODE == kModeScatter) {
        if (valid) {
            a.acc_x[id] = h.ax;
            a.acc_y[id] = h.ay;
            a.acc_z[id] = h.az;
        }
        return;
    }

    // ---- integrator.step2 + boundary_condition (main/integrator.py:58-114) ----
    float x = h.x, y = h.y, z = h.z, vx = h.vx, vy = h.vy, vz = h.vz;
    if (a.integrate) {
        const float dt = a.scalars->dt;
        x += (h.vx + 0.5f * h.ax * dt) * dt;
        y += (h.vy + 0.5f * h.ay * dt) * dt;
        z += (h.vz + 0.5f * h.az * dt) * dt;
        vx = fmaf(h.ax, 0.5f * dt, h.vx);
        vy = fmaf(h.ay, 0.5f * dt, h.vy);
        vz = fmaf(h.az, 0.5f * dt, h.vz);
        if (a.self_kind[h.type] == 0) { vx *= kDamping; vy *= kDamping; vz *= kDamping; }
        if (x < 0.0f) x += g.Lx; else if (x > g.Lx) x -= g.Lx;
        if (y < 0.0f) y += g.Ly; else if (y > g.Ly) y -= g.Ly;
        if (z < 0.0f) z += g.Lz; else if (z > g.Lz) z -= g.Lz;
    }
    float sq = 0.0f;
    if (valid) {
        a.posw_out[k] = make_float4(x, y, z, pi.w);
        a.veli_out[k] = make_float4(vx, vy, vz, vi.w);
        a.acc_out[k] = make_float4(h.ax, h.ay, h.az, 0.0f);
        sq = vx * vx + vy * vy + vz * vz;
    }
    // max |v|^2 for the next step's adaptive timestep (integrator.py:174-188)
    const unsigned m = __reduce_max_sync(0xffffffffu, __float_as_uint(sq));
    if (lane_id() == 0) atomicMax(&a.scalars->max_sq_speed_bits[a.parity_out], m);
}

// ---- diagnostics (bench.py): how much of the neighbour search is useful work --------------------
// For the snapshot the last force launch read, per home particle:
//   candidates  population of its 9 / 27 stencil cells (what a plain cell-list walk would test)
//   evaluated   force evaluations the Clusters kernel spends on it: the survivors of its warp's
//               box test, every run padded to a multiple of four (walk_culled); lanes on the other
//               paths (a warp that straddles two rows of cells, the literal z-wrap layers) and the
//               Boids / mixed kernels test every candidate
//   in_range    candidates closer than r_max (fp32, periodic image of the run; no float64 repair)
// Summed over the particles with one atomic per warp.  Not part of a step; never timed.
struct PairStats { unsigned long long candidates, evaluated, in_range; };

__global__ void __launch_bounds__(kForceThreads)
k_pair_stats(const __grid_constant__ Geometry g, const float4* __restrict__ posw,
             const int* __restrict__ cell_sorted, const int* __restrict__ start,
             const int* __restrict__ count, int n, int wrap_mode, int clusters, PairStats* __restrict__ out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = i < n;
    const int k = valid ? i : n - 1;
    const float4 pi = posw[k];
    Home h;
    h.x = pi.x; h.y = pi.y; h.z = pi.z;
    const int c = cell_sorted[k];
    const int cx = c % g.nbx, cyz = c / g.nbx, cy = cyz % g.nby, lz = cyz / g.nby, gz = lz + g.layer_base;
    bool quirk = g.is3d && wrap_mode == kWrapReference && (gz == 0 || gz == g.nbz - 1);
    int skip_dz = 2;
    if (quirk && g.nbz >= 4) {
        const double band = g.Lzd - g.rmaxd - 1e-6 * g.Lzd;
        const int rm = cy > 0 ? cy - 1 : g.nby - 1, rp = cy < g.nby - 1 ? cy + 1 : 0;
        const bool in_band = (double)(cy + 1) * g.bsy > band || (double)(rm + 1) * g.bsy > band ||
                             (double)(rp + 1) * g.bsy > band;
        if (!in_band) { quirk = false; skip_dz = gz == 0 ? -1 : 1; }
    }
    // (every lane takes part in the reductions: no short-circuit in front of them)
    const Box b = warp_box(h);
    const int cxl = (int)__reduce_min_sync(0xffffffffu, (unsigned)cx), cxh = (int)__reduce_max_sync(0xffffffffu, (unsigned)cx);
    const bool one_row = __reduce_min_sync(0xffffffffu, (unsigned)cyz) == __reduce_max_sync(0xffffffffu, (unsigned)cyz);
    const bool as_warp = clusters && one_row && __all_sync(0xffffffffu, !quirk);
    unsigned long long cand = 0, eval = 0, inr = 0;
    const int nrows = g.is3d ? 9 : 3;
    for (int r = 0; r < nrows; ++r) {
        const int dzi = g.is3d ? r / 3 - 1 : 0, dyi = g.is3d ? r % 3 - 1 : r - 1;
        int nl = lz + dzi, sz = 0;
        if (g.is3d) {
            const int ngz = gz + dzi;
            if (ngz < 0) { sz = 1; if (g.index_wrap_z) nl += g.nbz; }
            else if (ngz >= g.nbz) { sz = -1; if (g.index_wrap_z) nl -= g.nbz; }
        }
        if (g.is3d && dzi == skip_dz) continue;
        int ny = cy + dyi, sy = 0;
        if (ny < 0) { ny += g.nby; sy = 1; }
        else if (ny >= g.nby) { ny -= g.nby; sy = -1; }
        const int rb = (nl * g.nby + ny) * g.nbx;
        for (int part = 0; part < 3; ++part) {
            // own run of this lane (part 0: cells cx - 1 .. cx + 1; 1 / 2: the cell across the x edge)
            int s, e, sx = 0;
            int ws, we;  // the warp's run
            if (part == 0) {
                const int lo = cx > 0 ? cx - 1 : 0, hi = cx < g.nbx - 1 ? cx + 1 : g.nbx - 1;
                s = start[rb + lo]; e = start[rb + hi] + count[rb + hi];
                const int wlo = cxl > 0 ? cxl - 1 : 0, whi = cxh < g.nbx - 1 ? cxh + 1 : g.nbx - 1;
                ws = start[rb + wlo]; we = start[rb + whi] + count[rb + whi];
            } else if (part == 1) {
                sx = 1;
                s = start[rb + g.nbx - 1]; e = cx == 0 ? s + count[rb + g.nbx - 1] : s;
                ws = s; we = cxl == 0 ? s + count[rb + g.nbx - 1] : s;
            } else {
                sx = -1;
                s = start[rb]; e = cx == g.nbx - 1 ? s + count[rb] : s;
                ws = s; we = cxh == g.nbx - 1 ? s + count[rb] : s;
            }
            cand += (unsigned long long)(e - s);
            const Image im = make_image(h, g, sx, sy, sz);
            for (int j = s; j < e; ++j) {
                const float4 pj = ldg4(posw + j);
                const float dx = (im.x - pj.x) + im.lx, dy = (im.y - pj.y) + im.ly, dz = (im.z - pj.z) + im.lz;
                const float d2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
                inr += (d2 < g.r2_lo && d2 > 1e-10f) ? 1u : 0u;
            }
            if (as_warp) {
                const float tx = sx * g.Lx, ty = sy * g.Ly, tz = sz * g.Lz;
                const float lox = b.lox + tx - g.cull_slack, hix = b.hix + tx + g.cull_slack;
                const float loy = b.loy + ty - g.cull_slack, hiy = b.hiy + ty + g.cull_slack;
                const float loz = b.loz + tz - g.cull_slack, hiz = b.hiz + tz + g.cull_slack;
                unsigned kept = 0;
                for (int j = ws; j < we; ++j) {
                    const float4 p = ldg4(posw + j);
                    const float ex = fmaxf(fmaxf(lox - p.x, p.x - hix), 0.0f);
                    const float ey = fmaxf(fmaxf(loy - p.y, p.y - hiy), 0.0f);
                    const float ez = fmaxf(fmaxf(loz - p.z, p.z - hiz), 0.0f);
                    kept += fmaf(ez, ez, fmaf(ey, ey, ex * ex)) < g.cull_r2 ? 1u : 0u;
                }
                eval += (kept + 3u) & ~3u;
            } else {
                eval += (unsigned long long)(e - s);
            }
        }
    }
    if (!valid) cand = eval = inr = 0;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        cand += __shfl_xor_sync(0xffffffffu, cand, d);
        eval += __shfl_xor_sync(0xffffffffu, eval, d);
        inr += __shfl_xor_sync(0xffffffffu, inr, d);
    }
    if (lane_id() == 0) {
        atomicAdd(&out->candidates, cand);
        atomicAdd(&out->evaluated, eval);
        atomicAdd(&out->in_range, inr);
    }
}

}  // namespace gof
