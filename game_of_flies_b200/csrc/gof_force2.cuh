// gof_force2.cuh -- the Clusters force kernel of the resident engine: two home particles
// per lane, Blackwell packed fp32 arithmetic (FFMA2 / FADD2 / FMUL2, sm_100+).
//
// Why this shape.  ncu on the one-home kernel (profiles/r01_force_scalar_*.txt) showed it
// limited by two per-SM rates at once: instruction issue (~25 slots per candidate pair)
// and the 128 B/clk return path from L1/shared to the register file (16 B of neighbour
// + 24 B of species parameters per lane and pair).  Here a lane owns TWO consecutive
// particles of the same cell, so that
//   * every broadcast neighbour load (one 16-byte {x,y,z,type}) feeds two pairs: 8 B per lane
//     and pair through the L1 -> register path, which is what limits the one-home kernel
//     (its test-only loop already takes 0.93 ms at 26 % issue utilisation);
//   * separations, squared distances, profile and accumulation are one packed instruction
//     for both pairs: ~15 issue slots per pair;
//   * the species parameters live in registers and are reloaded only when the neighbour
//     type changes (particles are ordered by id inside a cell and ids are numbered species
//     by species, so that happens ~once per 6 neighbours).
// The cell segments are padded to an even number of slots (gof_binning.cuh) so that the
// two particles of a lane always share a cell and therefore walk the same runs.
//
// Semantics are those of k_force<false, kModeIntegrate> (gof_force.cuh): same stencil,
// same guard band / float64 repair of the cut-off, same reference z-wrap path for the
// edge layers (taken per home particle through the scalar code), same fused integrator.
#pragma once
#include "gof_force.cuh"

namespace gof {

constexpr int kForce2Threads = 128;  // 256 particles per block

__device__ __forceinline__ float2 f2(float a, float b) { return make_float2(a, b); }
__device__ __forceinline__ float2 add2(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 mul2(float2 a, float2 b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }

// Packed species parameters of (home type 0, home type 1) x current neighbour type for
//   f(d) = s1 * (hw - |d - mid|) + cc * min(d - r_min, 0),   cc = -k0 - s1
// which equals the reference profile: the triangle for d >= r_min, f_min - k0 d below.
// They come out of shared memory already paired: the block builds a T^3 table indexed
// (t0, t1, tj) whose entries are {-r_min0, -r_min1, -mid0, -mid1}, {-s1_0, -s1_1, f_max0,
// f_max1}, {cc0, cc1}, so that one 16-byte load lands two operands of a packed instruction
// in an aligned register pair and nothing has to be shuffled.
struct PackedParams {
    float4 q0;  // -r_min (x, y), -mid (z, w)
    float4 q1;  // -s1 (x, y), f_max (z, w)
    float2 q2;  // cc
    int type;   // neighbour type these belong to
};

constexpr int kPackedMaxTypes = 9;  // T^3 * 40 B of shared memory: 29 KB at T = 9 (the reference UI's maximum)

struct Derived {
    const float4* q0;
    const float4* q1;
    const float2* q2;
    int T;
    int row;  // (t0 * T + t1) * T of this lane
};

__device__ __forceinline__ void load_params(PackedParams& p, const Derived& d, int tj)
{
    const int k = d.row + tj;
    p.q0 = d.q0[k];
    p.q1 = d.q1[k];
    p.q2 = d.q2[k];
    p.type = tj;
}

// The two home particles of a lane as seen by one run of neighbours.
struct Image2 {
    float2 nx, ny, nz;     // -(home + s * L), rounded
    float2 nlx, nly, nlz;  // -(what the rounded sum lost); only used by LOWFIX runs
    int shifts;
    bool up;
};

struct Home2 {
    float2 x, y, z;
    float2 ax, ay, az;
    int t0, t1;
};

__device__ __forceinline__ Image2 make_image2(const Home2& h, const Geometry& g, int sx, int sy, int sz)
{
    Image2 im;
    float hi, lo;
    shifted(h.x.x, sx, g.Lx, g.Lx_lo, hi, lo); im.nx.x = -hi; im.nlx.x = -lo;
    shifted(h.x.y, sx, g.Lx, g.Lx_lo, hi, lo); im.nx.y = -hi; im.nlx.y = -lo;
    shifted(h.y.x, sy, g.Ly, g.Ly_lo, hi, lo); im.ny.x = -hi; im.nly.x = -lo;
    shifted(h.y.y, sy, g.Ly, g.Ly_lo, hi, lo); im.ny.y = -hi; im.nly.y = -lo;
    shifted(h.z.x, sz, g.Lz, g.Lz_lo, hi, lo); im.nz.x = -hi; im.nlz.x = -lo;
    shifted(h.z.y, sz, g.Lz, g.Lz_lo, hi, lo); im.nz.y = -hi; im.nlz.y = -lo;
    im.shifts = pack_shifts(sx, sy, sz);
    im.up = (im.nlx.x != 0.0f) | (im.nlx.y != 0.0f) | (im.nly.x != 0.0f) | (im.nly.y != 0.0f) |
            (im.nlz.x != 0.0f) | (im.nlz.y != 0.0f);
    return im;
}

// Cold path: the hot loop added the pair's force iff `d2_hot < r2_lo`; take the
// reference's float64 decision and return the difference (see repair_pair_clusters).
__device__ __noinline__ float3 repair_pair2(const Geometry& g, float xi, float yi, float zi,
                                            float nimx, float nimy, float nimz, float nlx,
                                            float nly, float nlz, int shifts, float xj, float yj,
                                            float zj, float d2_hot, const PairParam* q)
{
    const bool added = d2_hot < g.r2_lo;
    const double e = exact_dist2(g, xi, yi, zi, xj, yj, zj, shifts);
    const bool wanted = (1e-10 < e) && (e < g.r2d);
    if (added == wanted) return make_float3(0.f, 0.f, 0.f);
    // home - neighbour, from the same operands as the hot loop (signs mirrored exactly)
    const float dx = -((xj + nimx) + nlx), dy = -((yj + nimy) + nly), dz = -((zj + nimz) + nlz);
    const float d2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
    const float4 qa = q->a;
    const float4 qb = q->b;
    const float inv = fast_rsqrt(d2 + kTinyD2);
    const float dist = d2 * inv;
    const float f_in = fmaf(-qa.y, dist, qa.z);
    const float f_tri = fmaf(-qb.x, fabsf(dist - qa.w), qb.y);
    float s = (dist < qa.x ? f_in : f_tri) * inv;
    if (added) s = -s;  // take it back
    return make_float3(-s * dx, -s * dy, -s * dz);
}

// One neighbour against the lane's two home particles.  Returns the two squared
// distances; `band` / `dmin` collect what the caller's repair vote needs.
template <bool LOWFIX>
__device__ __forceinline__ float2 neighbour2(Home2& h, PackedParams& pp, const Derived& der,
                                             const Geometry& g, const Image2& im, float2 nmid2,
                                             float hw2, const float4 P, bool& band, float& dmin)
{
    // One 16-byte neighbour {x, y, z, type} serves both home particles: the duplicated operands
    // of the packed instructions are formed in registers (a few MOVs on an idle pipe), not
    // loaded -- the L1 -> register return path (128 B/clk/SM, 512 B per broadcast LDG.128) is
    // what limits this loop, not the issue slots.
    // neighbour - home (the force below is accumulated with the matching sign)
    float2 dx = add2(f2(P.x, P.x), im.nx), dy = add2(f2(P.y, P.y), im.ny), dz = add2(f2(P.z, P.z), im.nz);
    if (LOWFIX) { dx = add2(dx, im.nlx); dy = add2(dy, im.nly); dz = add2(dz, im.nlz); }
    const float2 d2 = fma2(dz, dz, fma2(dy, dy, fma2(dx, dx, f2(kTinyD2, kTinyD2))));
    const int tj = __float_as_int(P.w);
    // neighbours are ordered by id inside a cell and ids are numbered species by species: the
    // type changes about once per six neighbours.  The vote makes this a real (mostly
    // not-taken) branch instead of a handful of predicated loads on every neighbour.
    if (__any_sync(__activemask(), tj != pp.type)) load_params(pp, der, tj);
    const float2 inv = f2(fast_rsqrt(d2.x), fast_rsqrt(d2.y));
    const float2 u = fma2(d2, inv, f2(pp.q0.x, pp.q0.y));  // d - r_min
    const float2 t = fma2(d2, inv, f2(pp.q0.z, pp.q0.w));  // d - mid
    float2 f = fma2(f2(pp.q1.x, pp.q1.y), f2(fabsf(t.x), fabsf(t.y)), f2(pp.q1.z, pp.q1.w));
    f = fma2(pp.q2, f2(fminf(u.x, 0.0f), fminf(u.y, 0.0f)), f);
    const float2 mask = f2(d2.x < g.r2_lo ? 1.0f : 0.0f, d2.y < g.r2_lo ? 1.0f : 0.0f);
    const float2 s = mul2(mul2(f, inv), mask);
    // force on home = -f * (home - neighbour) / dist = +s * (neighbour - home); the self
    // pair contributes s * 0 = 0 with s finite
    h.ax = fma2(s, dx, h.ax);
    h.ay = fma2(s, dy, h.ay);
    h.az = fma2(s, dz, h.az);
    // repair vote: inside the guard band (|d2 - centre| < half width) or at/below 1e-10
    const float2 c = add2(d2, nmid2);
    band |= (fabsf(c.x) < hw2) | (fabsf(c.y) < hw2);
    dmin = fminf(dmin, fminf(d2.x, d2.y));
    return d2;
}

__device__ __forceinline__ bool needs_repair2(const Geometry& g, float2 d2)
{
    return needs_repair(g, d2.x) || needs_repair(g, d2.y);
}

__device__ __forceinline__ void repair2(Home2& h, const PairParam* __restrict__ tab, int T,
                                        const Geometry& g, const Image2& im, const float4 P, float2 d2)
{
    const int tj = __float_as_int(P.w);
    if (needs_repair(g, d2.x)) {
        const float3 f = repair_pair2(g, h.x.x, h.y.x, h.z.x, im.nx.x, im.ny.x, im.nz.x, im.nlx.x,
                                      im.nly.x, im.nlz.x, im.shifts, P.x, P.y, P.z, d2.x,
                                      tab + h.t0 * T + tj);
        h.ax.x += f.x; h.ay.x += f.y; h.az.x += f.z;
    }
    if (needs_repair(g, d2.y)) {
        const float3 f = repair_pair2(g, h.x.y, h.y.y, h.z.y, im.nx.y, im.ny.y, im.nz.y, im.nlx.y,
                                      im.nly.y, im.nlz.y, im.shifts, P.x, P.y, P.z, d2.y,
                                      tab + h.t1 * T + tj);
        h.ax.y += f.x; h.ay.y += f.y; h.az.y += f.z;
    }
}

template <bool LOWFIX>
__device__ __forceinline__ void run_pairs2(Home2& h, PackedParams& pp, const Derived& der,
                                           const PairParam* __restrict__ tab, const ForceArgs& a,
                                           const Geometry& g, const Image2& im, int s, int e)
{
    // the run sums into its own accumulators (shorter partial sums: less fp32 rounding)
    const float2 ax0 = h.ax, ay0 = h.ay, az0 = h.az;
    h.ax = h.ay = h.az = f2(0.f, 0.f);
    const float4* __restrict__ src = a.posw;  // padded snapshot {x, y, z, type}
    // centre and half width of the guard band [r2_lo, r2_hi) (a little wider: a false alarm only
    // costs a trip through the repair code, which decides exactly)
    const float mid = 0.5f * (g.r2_lo + g.r2_hi), hw2 = 0.51f * (g.r2_hi - g.r2_lo) + 1e-30f;
    const float2 nmid2 = f2(-mid, -mid);
    int j = s;
    for (; j + 4 <= e; j += 4) {
        const float4 P0 = ldg4(src + j), P1 = ldg4(src + j + 1), P2 = ldg4(src + j + 2), P3 = ldg4(src + j + 3);
        bool band = false;
        float dmin = 1.0f;
        const float2 d0 = neighbour2<LOWFIX>(h, pp, der, g, im, nmid2, hw2, P0, band, dmin);
        const float2 d1 = neighbour2<LOWFIX>(h, pp, der, g, im, nmid2, hw2, P1, band, dmin);
        const float2 d2 = neighbour2<LOWFIX>(h, pp, der, g, im, nmid2, hw2, P2, band, dmin);
        const float2 d3 = neighbour2<LOWFIX>(h, pp, der, g, im, nmid2, hw2, P3, band, dmin);
        if (band | (dmin <= 1e-10f)) {
            if (needs_repair2(g, d0)) repair2(h, tab, der.T, g, im, P0, d0);
            if (needs_repair2(g, d1)) repair2(h, tab, der.T, g, im, P1, d1);
            if (needs_repair2(g, d2)) repair2(h, tab, der.T, g, im, P2, d2);
            if (needs_repair2(g, d3)) repair2(h, tab, der.T, g, im, P3, d3);
        }
    }
    for (; j < e; ++j) {
        const float4 P0 = ldg4(src + j);
        bool band = false;
        float dmin = 1.0f;
        const float2 d0 = neighbour2<LOWFIX>(h, pp, der, g, im, nmid2, hw2, P0, band, dmin);
        if (needs_repair2(g, d0)) repair2(h, tab, der.T, g, im, P0, d0);
    }
    h.ax = add2(h.ax, ax0); h.ay = add2(h.ay, ay0); h.az = add2(h.az, az0);
}

__device__ __forceinline__ void run_image2(Home2& h, PackedParams& pp, const Derived& der,
                                           const PairParam* __restrict__ tab, const ForceArgs& a,
                                           const Geometry& g, int sx, int sy, int sz, int s, int e)
{
    const Image2 im = make_image2(h, g, sx, sy, sz);
    if (__any_sync(__activemask(), im.up && e > s))
        run_pairs2<true>(h, pp, der, tab, a, g, im, s, e);
    else
        run_pairs2<false>(h, pp, der, tab, a, g, im, s, e);
}

#ifndef GOF_FORCE2_MINBLOCKS
#define GOF_FORCE2_MINBLOCKS 4
#endif

__global__ void __launch_bounds__(kForce2Threads, GOF_FORCE2_MINBLOCKS)
k_force2(const __grid_constant__ Geometry g, const __grid_constant__ ForceArgs a)
{
    extern __shared__ float4 smem_raw[];
    const int T = a.T;
    PairParam* tab = reinterpret_cast<PairParam*>(smem_raw);  // AoS table: cold paths
    float4* q0 = reinterpret_cast<float4*>(tab + T * T);
    float4* q1 = q0 + T * T * T;
    float2* q2 = reinterpret_cast<float2*>(q1 + T * T * T);
    // split tables of the one-home code (literal z-wrap path of the edge layers)
    const int TA = T | 1;
    float4* tabA = reinterpret_cast<float4*>(q2 + T * T * T + (T * T * T & 1));
    float2* tabB = reinterpret_cast<float2*>(tabA + T * TA);
    for (int q = threadIdx.x; q < T * T; q += blockDim.x) {
        const PairParam pp = a.table[q];
        tab[q] = pp;
        const int r = q / T, c = q - r * T;
        tabA[r * TA + c] = pp.a;
        tabB[r * TA + c] = make_float2(pp.b.x, pp.b.y);
    }
    for (int q = threadIdx.x; q < T * T * T; q += blockDim.x) {
        const int tj = q % T, t1 = (q / T) % T, t0 = q / (T * T);
        // clusters: a = {r_min, k0, f_min, mid}, b = {s1, f_max}
        const PairParam p0 = a.table[t0 * T + tj], p1 = a.table[t1 * T + tj];
        q0[q] = make_float4(-p0.a.x, -p1.a.x, -p0.a.w, -p1.a.w);
        q1[q] = make_float4(-p0.b.x, -p1.b.x, p0.b.y, p1.b.y);
        q2[q] = make_float2(-p0.a.y - p0.b.x, -p1.a.y - p1.b.x);
    }
    __syncthreads();
    Derived der;
    der.q0 = q0;
    der.q1 = q1;
    der.q2 = q2;
    der.T = T;

    // padded slot index space: this lane owns slots k0 and k0 + 1 of one cell
    const int n_pad = a.pstart[g.num_cells];
    const int slot = 2 * (blockIdx.x * blockDim.x + threadIdx.x);
    const bool in_range = slot < n_pad;
    const int k0 = in_range ? slot : 0;
    const int c = a.cell_sorted[k0];
    const int first = a.pstart[c], cnt = a.count[c];
    const bool valid0 = in_range && (k0 - first) < cnt;
    const bool valid1 = in_range && (k0 + 1 - first) < cnt;  // false: the cell's padding slot

    const float4 p0 = a.posw[k0], p1 = a.posw[k0 + 1];
    const float4 v0 = a.veli[k0], v1 = a.veli[k0 + 1];
    Home2 h;
    h.x = f2(p0.x, p1.x); h.y = f2(p0.y, p1.y); h.z = f2(p0.z, p1.z);
    h.ax = h.ay = h.az = f2(0.f, 0.f);
    h.t0 = __float_as_int(p0.w);
    h.t1 = __float_as_int(p1.w);  // padding slots carry type 0
    der.row = (h.t0 * T + h.t1) * T;
    PackedParams pp;
    pp.type = -1;
    pp.q0 = pp.q1 = make_float4(0.f, 0.f, 0.f, 0.f);
    pp.q2 = f2(0.f, 0.f);

    const int cx = c % g.nbx;
    const int cyz = c / g.nbx;
    const int cy = cyz % g.nby;
    const int lz = cyz / g.nby;
    const int gz = lz + g.layer_base;
    // per lane, like the one-home kernel: a particle's arithmetic never depends on warp-mates
    const bool quirk = g.is3d && a.wrap_mode == kWrapReference && (gz == 0 || gz == g.nbz - 1);

    const int dz_lo = g.is3d ? -1 : 0, dz_hi = g.is3d ? 1 : 0;
    if (!quirk) {
        for (int dzi = dz_lo; dzi <= dz_hi; ++dzi) {
            int nl = lz + dzi, sz = 0;
            if (g.is3d) {
                const int ngz = gz + dzi;
                if (ngz < 0) { sz = 1; if (g.index_wrap_z) nl += g.nbz; }
                else if (ngz >= g.nbz) { sz = -1; if (g.index_wrap_z) nl -= g.nbz; }
            }
            for (int dyi = -1; dyi <= 1; ++dyi) {
                int ny = cy + dyi, sy = 0;
                if (ny < 0) { ny += g.nby; sy = 1; }
                else if (ny >= g.nby) { ny -= g.nby; sy = -1; }
                const int rb = (nl * g.nby + ny) * g.nbx;
                const int lo = cx > 0 ? cx - 1 : 0;
                const int hi = cx < g.nbx - 1 ? cx + 1 : g.nbx - 1;
                // padded starts: the run also covers the padding slots of its inner cells,
                // which sit at 1e18 and fail the distance test like any far particle
                int s = a.pstart[rb + lo];
                int e = a.pstart[rb + hi] + a.count[rb + hi];
                if (!valid0) e = s;
                run_image2(h, pp, der, tab, a, g, 0, sy, sz, s, e);
                const bool xedge = cx == 0 || cx == g.nbx - 1;
                const int wc = cx == 0 ? g.nbx - 1 : 0;
                s = a.pstart[rb + wc];
                e = (valid0 && xedge) ? s + a.count[rb + wc] : s;
                run_image2(h, pp, der, tab, a, g, cx == 0 ? 1 : -1, sy, sz, s, e);
            }
        }
    } else {
        // reference z-wrap on the bottom / top layer: the literal per-pair path of the
        // one-home kernel, once per home particle (gof_force.cuh run_pairs_zquirk)
        for (int which = 0; which < 2; ++which) {
            const bool valid = which ? valid1 : valid0;
            const int type = which ? h.t1 : h.t0;
            RowAddr tabrow_s;
            tabrow_s.a = (unsigned)__cvta_generic_to_shared(tabA + type * TA);
            tabrow_s.b = (unsigned)__cvta_generic_to_shared(tabB + type * TA);
            const float3 f = quirk_home<false, 1>(which ? h.x.y : h.x.x, which ? h.y.y : h.y.x, which ? h.z.y : h.z.x,
                                                  type, 0u, tabrow_s, tab + type * T, nullptr, a, g, cx, cy, lz, gz,
                                                  valid, 0, 9, nullptr);
            if (which) { h.ax.y = f.x; h.ay.y = f.y; h.az.y = f.z; }
            else       { h.ax.x = f.x; h.ay.x = f.y; h.az.x = f.z; }
        }
    }

    // ---- integrator.step2 + boundary_condition (main/integrator.py:58-114), per home ----
    const int out0 = a.start[c] + (k0 - first);  // compact (unpadded) rest order
    float sqmax = 0.0f;
#pragma unroll
    for (int which = 0; which < 2; ++which) {
        const bool valid = which ? valid1 : valid0;
        const float4 p = which ? p1 : p0;
        const float4 v = which ? v1 : v0;
        const float ax = which ? h.ax.y : h.ax.x, ay = which ? h.ay.y : h.ay.x, az = which ? h.az.y : h.az.x;
        float x = p.x, y = p.y, z = p.z, vx = v.x, vy = v.y, vz = v.z;
        if (a.integrate) {
            const float dt = a.scalars->dt;
            x += (v.x + 0.5f * ax * dt) * dt;
            y += (v.y + 0.5f * ay * dt) * dt;
            z += (v.z + 0.5f * az * dt) * dt;
            vx = fmaf(ax, 0.5f * dt, v.x);
            vy = fmaf(ay, 0.5f * dt, v.y);
            vz = fmaf(az, 0.5f * dt, v.z);
            if (a.self_kind[__float_as_int(p.w)] == 0) { vx *= kDamping; vy *= kDamping; vz *= kDamping; }
            if (x < 0.0f) x += g.Lx; else if (x > g.Lx) x -= g.Lx;
            if (y < 0.0f) y += g.Ly; else if (y > g.Ly) y -= g.Ly;
            if (z < 0.0f) z += g.Lz; else if (z > g.Lz) z -= g.Lz;
        }
        if (valid) {
            const int o = out0 + which;
            a.posw_out[o] = make_float4(x, y, z, p.w);
            a.veli_out[o] = make_float4(vx, vy, vz, v.w);
            a.acc_out[o] = make_float4(ax, ay, az, 0.0f);
            sqmax = fmaxf(sqmax, vx * vx + vy * vy + vz * vz);
        }
    }
    const unsigned m = __reduce_max_sync(0xffffffffu, __float_as_uint(sqmax));
    if (lane_id() == 0) atomicMax(&a.scalars->max_sq_speed_bits[a.parity_out], m);
}

}  // namespace gof
