// gof_kernels.cuh -- device code of the B200-native Game_of_Flies engine.
//
// Written for sm_100a only.  The work is HBM/ALU-bound integer and fp32 SIMT
// work (cell hashing, counting sort, pairwise stencil forces): no tensor-core
// contraction exists in this path, so the Blackwell rules that matter are
// coalesced 16-byte accesses on cell-sorted float4 arrays, L1/shared reuse of
// neighbour cells, warp-aggregated atomics and grid sizing.
//
// Reference behaviour restated here (never copied): main/acceleration_updater.py
// (bin_particles :38-76, cumsum :81-121, set_indices :125-135, accelerator
// :219-458) and main/integrator.py (step1 :33-55, step2 :58-88,
// boundary_condition :91-114, adaptive timestep :174-188).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace gof {

constexpr int kMaxTypes = 16;
constexpr float kMaxSpeed = 20.0f;  // integrator.py:14  max_spd
constexpr float kDamping = 0.95f;   // integrator.py:13  fac
constexpr int kDead = -1;           // type of a particle that left the slab
constexpr int kWrapReference = 0;   // GOF_WRAP_REFERENCE
constexpr int kWrapMinImage = 1;    // GOF_WRAP_MINIMAGE

// One entry per ordered species pair (p_i, p_j): what particle i feels from j.
// Clusters kind (parameter_matrix[-1] == 0), args = r_min, r_max, f_min, f_max:
//   d <  r_min : f = f_min - k0 * d                      (k0 = f_min / r_min)
//   d >= r_min : f = f_max - s1 * |d - mid|              (triangle; s1 = 2 f_max / (r_max - r_min),
//                                                         mid = (r_min + r_max) / 2; no clamp past r_max)
// Boids kind (== 1), args = r_max, separation, alignment, cohesion:
//   a = {r_sep = args[0] / 5, separation, alignment, cohesion}
struct __align__(16) PairParam {
    float4 a;  // clusters: {r_min, k0, f_min, mid}   boids: {r_sep, separation, alignment, cohesion}
    float4 b;  // clusters: {s1, f_max, 0, 0}         boids: unused
};

// Geometry and constants shared by the kernels (passed by value as a grid constant).
struct Geometry {
    int nbx, nby;        // bins along x, y
    int nlz;             // local layers along z held by this engine (1 in 2D)
    int nbz;             // global layers along z (0 in 2D)
    int layer_base;      // global layer index of local layer 0 (slab mode: lo - 1, may be -1)
    int index_wrap_z;    // 1: the local grid is the whole periodic box (single GPU)
    int num_cells;       // nbx * nby * nlz
    int is3d;
    double bsx, bsy, bsz;  // bin sizes, float64 like the reference
    double Lxd, Lyd, Lzd, rmaxd, r2d;
    float Lx, Ly, Lz, rmax;
    float Lx_lo, Ly_lo, Lz_lo;  // what (float)L lost: L = L + L_lo to ~2^-48 (exact periodic images)
    float r2_lo, r2_hi;  // guard band around r_max^2: inside it the cut-off is re-decided in float64
    // warp-level culling of candidates against the box of a warp's home particles: a candidate
    // farther than sqrt(cull_r2) from the box grown by cull_slack is out of range for all of them
    float cull_r2, cull_slack;
    // thresholds of the reference's z-wrap tests, pre-rounded so that fp32 comparisons of fp32
    // coordinates decide exactly like the reference's float64 ones:
    //   z < r_max  <=>  z < zq_r     (zq_r   = smallest float >= r_max)
    //   z > Lz - r_max  <=>  z > zq_top  (zq_top = largest float <= Lz - r_max)
    float zq_r, zq_top;
};

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }

// floor(x / bs) exactly as Python's float `//` yields it for x >= 0 (the reference
// bins with `pos // bin_size` in float64, acceleration_updater.py:62-66): take the
// rounded quotient, then repair the one case where rounding crossed an integer by
// testing q * bs <= x exactly with a fused multiply-add.
__device__ __forceinline__ int bin_coord(float x, double bs, int nb)
{
    const double xd = (double)x;
    double q = floor(xd / bs);
    if (fma(q, bs, -xd) > 0.0) q -= 1.0;
    int c = (q > 0.0) ? (int)q : 0;  // negative or NaN -> 0
    return c < nb ? c : nb - 1;      // x == limit (reference would index past the row) -> last bin
}

__device__ __forceinline__ float4 ldg4(const float4* p) { return __ldg(p); }

}  // namespace gof
