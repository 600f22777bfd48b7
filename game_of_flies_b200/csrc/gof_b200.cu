// gof_b200.cu -- C ABI (include/gof_b200.h) of the B200-native Game_of_Flies engine.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -shared
//        -Xcompiler -fPIC  (see game_of_flies_b200/build.py)
//
// The host code below only sequences kernel launches and carves caller-provided
// device memory; every number a user sees is produced by the kernels in
// gof_binning.cuh / gof_force.cuh and the small element-wise kernels here.
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "../../include/gof_b200.h"
#include "gof_force.cuh"
#include "gof_slab_p2p.cuh"

using namespace gof;

// --------------------------------------------------------------------------------------
// errors, launch accounting
// --------------------------------------------------------------------------------------
static thread_local std::string g_last_error;
static std::atomic<uint64_t> g_launches{0};

static int fail(int code, const std::string& msg)
{
    g_last_error = msg;
    return code;
}

#define GOF_CUDA(expr)                                                                        \
    do {                                                                                      \
        cudaError_t err__ = (expr);                                                           \
        if (err__ != cudaSuccess)                                                             \
            return fail((int)err__, std::string(#expr) + ": " + cudaGetErrorString(err__));   \
    } while (0)

#define GOF_LAUNCH(kernel, grid, block, smem, stream, ...)                                    \
    do {                                                                                      \
        kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__);                           \
        g_launches.fetch_add(1, std::memory_order_relaxed);                                   \
        cudaError_t err__ = cudaGetLastError();                                               \
        if (err__ != cudaSuccess)                                                             \
            return fail((int)err__, std::string(#kernel) + ": " + cudaGetErrorString(err__)); \
    } while (0)

static inline int cdiv(int a, int b) { return (a + b - 1) / b; }

// Every entry point runs on its engine's (or its buffers') device and leaves the caller's
// current device as it found it.
struct DeviceGuard {
    int prev = -1, want = -1;
    explicit DeviceGuard(int device) : want(device)
    {
        if (device < 0 || cudaGetDevice(&prev) != cudaSuccess) { prev = -1; return; }
        if (prev != device && cudaSetDevice(device) != cudaSuccess) prev = -1;
    }
    ~DeviceGuard()
    {
        if (prev >= 0 && prev != want) cudaSetDevice(prev);
    }
    DeviceGuard(const DeviceGuard&) = delete;
    DeviceGuard& operator=(const DeviceGuard&) = delete;
};

// device that owns a device pointer (-1: not a device pointer / unknown: stay on the current device)
static int device_of(const void* p)
{
    cudaPointerAttributes at{};
    if (!p || cudaPointerGetAttributes(&at, p) != cudaSuccess) {
        cudaGetLastError();
        return -1;
    }
    return (at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged) ? at.device : -1;
}

// SM count of the current device (queried once per device)
static int sm_count()
{
    static std::atomic<int> cache[64];
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    int v = cache[dev].load(std::memory_order_relaxed);
    if (v > 0) return v;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
    cache[dev].store(v, std::memory_order_relaxed);
    return v;
}
static inline size_t align_up(size_t v, size_t a = 256) { return (v + a - 1) / a * a; }


extern "C" int gof_abi_version(void) { return GOF_ABI_VERSION; }
extern "C" const char* gof_last_error(void) { return g_last_error.c_str(); }
extern "C" uint64_t gof_launch_count(void) { return g_launches.load(); }
extern "C" void gof_reset_launch_count(void) { g_launches.store(0); }
extern "C" int gof_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

// --------------------------------------------------------------------------------------
// host-side geometry (reference main/simulation.py:103-139)
// --------------------------------------------------------------------------------------
extern "C" int gof_grid_init(gof_grid* grid, double limit_x, double limit_y, double limit_z,
                             double r_max)
{
    if (!grid || !(r_max > 0) || !(limit_x > 0) || !(limit_y > 0) || limit_z < 0)
        return fail(GOF_E_INVALID, "gof_grid_init: bad limits or r_max");
    grid->limit_x = limit_x;
    grid->limit_y = limit_y;
    grid->limit_z = limit_z;
    grid->r_max = r_max;
    grid->num_bin_x = (int)std::floor(limit_x / r_max);
    grid->num_bin_y = (int)std::floor(limit_y / r_max);
    grid->num_bin_z = limit_z != 0 ? (int)std::floor(limit_z / r_max) : 0;
    if (grid->num_bin_x < 1 || grid->num_bin_y < 1 || (limit_z != 0 && grid->num_bin_z < 1))
        return fail(GOF_E_GRID, "gof_grid_init: r_max larger than the box");
    grid->bin_size_x = limit_x / grid->num_bin_x;
    grid->bin_size_y = limit_y / grid->num_bin_y;
    grid->bin_size_z = limit_z != 0 ? limit_z / grid->num_bin_z : 0.0;
    grid->num_cells = grid->num_bin_x * grid->num_bin_y * (grid->num_bin_z > 0 ? grid->num_bin_z : 1);
    if (grid->num_bin_x < 3 || grid->num_bin_y < 3 || (limit_z != 0 && grid->num_bin_z < 3))
        return fail(GOF_E_GRID,
                    "fewer than 3 bins along an axis: the reference assumes r_max < min(limits) / 3 "
                    "(acceleration_updater.py:302)");
    return 0;
}

extern "C" int gof_set_bin_neighbours(int nbx, int nby, int nbz, int32_t* out)
{
    if (!out || nbx < 1 || nby < 1) return fail(GOF_E_INVALID, "gof_set_bin_neighbours: bad arguments");
    auto wrap = [](int v, int n) { return v < 0 ? v + n : (v >= n ? v - n : v); };
    // offsets of the half stencil, in the order of the reference table
    static const int o2[5][2] = {{0, 0}, {1, 0}, {-1, 1}, {0, 1}, {1, 1}};
    static const int o3[9][2] = {{1, 0}, {-1, 0}, {0, 1}, {0, -1}, {1, 1}, {1, -1}, {-1, 1}, {-1, -1}, {0, 0}};
    const int width = nbz < 1 ? 5 : 14;
    const int layers = nbz < 1 ? 1 : nbz;
    for (int k = 0; k < layers; ++k)
        for (int j = 0; j < nby; ++j)
            for (int i = 0; i < nbx; ++i) {
                int32_t* row = out + (size_t)(i + nbx * (j + nby * k)) * width;
                for (int q = 0; q < 5; ++q)
                    row[q] = wrap(i + o2[q][0], nbx) + nbx * (wrap(j + o2[q][1], nby) + nby * k);
                if (nbz >= 1) {
                    const int kp = wrap(k + 1, nbz);
                    for (int q = 0; q < 9; ++q)
                        row[5 + q] = wrap(i + o3[q][0], nbx) + nbx * (wrap(j + o3[q][1], nby) + nby * kp);
                }
            }
    return 0;
}

static float next_down(float v) { return std::nextafterf(v, -INFINITY); }
static float next_up(float v) { return std::nextafterf(v, INFINITY); }

// local_layers == 0: the engine holds the whole periodic box.
static int make_geometry(const gof_grid* grid, int layer_base, int local_layers, Geometry* out)
{
    if (!grid) return fail(GOF_E_INVALID, "null grid");
    const bool is3d = grid->num_bin_z > 0;
    if (grid->num_bin_x < 3 || grid->num_bin_y < 3 || (is3d && grid->num_bin_z < 3))
        return fail(GOF_E_GRID, "fewer than 3 bins along an axis (reference assumes r_max < min(limits) / 3)");
    if (is3d && !(grid->bin_size_z > 1.0))
        return fail(GOF_E_GRID, "3D box with bin_size_z <= 1: the reference bins such a box as 2D "
                                "(acceleration_updater.py:54); refuse instead of diverging");
    Geometry g{};
    g.nbx = grid->num_bin_x;
    g.nby = grid->num_bin_y;
    g.nbz = is3d ? grid->num_bin_z : 0;
    g.is3d = is3d;
    if (local_layers > 0) {
        g.nlz = local_layers;
        g.layer_base = layer_base;
        g.index_wrap_z = 0;
    } else {
        g.nlz = is3d ? grid->num_bin_z : 1;
        g.layer_base = 0;
        g.index_wrap_z = 1;
    }
    g.num_cells = g.nbx * g.nby * g.nlz;
    g.bsx = grid->bin_size_x;
    g.bsy = grid->bin_size_y;
    g.bsz = grid->bin_size_z;
    g.Lxd = grid->limit_x;
    g.Lyd = grid->limit_y;
    g.Lzd = grid->limit_z;
    g.rmaxd = grid->r_max;
    g.r2d = grid->r_max * grid->r_max;
    g.Lx = (float)g.Lxd;
    g.Ly = (float)g.Lyd;
    g.Lz = (float)g.Lzd;
    g.rmax = (float)g.rmaxd;
    g.Lx_lo = (float)(g.Lxd - (double)g.Lx);
    g.Ly_lo = (float)(g.Lyd - (double)g.Ly);
    g.Lz_lo = (float)(g.Lzd - (double)g.Lz);
    {
        float r = (float)g.rmaxd;
        if ((double)r < g.rmaxd) r = next_up(r);
        g.zq_r = r;
        const double top = g.Lzd - g.rmaxd;
        float t = (float)top;
        if ((double)t > top) t = next_down(t);
        g.zq_top = t;
    }
    // Guard band: the fp32 squared distance of a pair whose true distance is r_max can be
    // off by the rounding of the shifted coordinate (<= ulp(L + r_max)), the rounding of
    // the difference and of the three squares.  Inside the band the kernel decides in fp64.
    const double lmax = std::fmax(g.Lxd, std::fmax(g.Lyd, g.Lzd)) + g.rmaxd;
    const double ulp = (double)next_up((float)lmax) - (double)(float)lmax;
    const double band = 2.0 * std::sqrt(3.0) * g.rmaxd * (3.0 * ulp) + 32.0 * 1.1920929e-7 * g.r2d;
    g.r2_lo = next_down((float)(g.r2d - band));
    g.r2_hi = next_up((float)(g.r2d + band));
    g.cull_r2 = g.r2_hi * 1.0001f;
    {
        float m = g.Lx > g.Ly ? g.Lx : g.Ly;
        if (g.Lz > m) m = g.Lz;
        g.cull_slack = 2e-6f * m;  // a few ulps of a coordinate: images are rounded sums
    }
    *out = g;
    return 0;
}

// --------------------------------------------------------------------------------------
// species table
// --------------------------------------------------------------------------------------
struct HostSpecies {
    int T = 0;
    bool has_boids = false;     // some pair is of the Boids kind
    bool has_clusters = false;  // some pair is not
    int kind() const { return !has_boids ? kKindClusters : (has_clusters ? kKindMixed : kKindBoids); }
    std::vector<PairParam> table;
    std::vector<unsigned> kind_rows;
    std::vector<double> sep_exact;
    std::vector<int> self_kind;
    size_t bytes() const
    {
        return align_up(table.size() * sizeof(PairParam)) + align_up(kind_rows.size() * sizeof(unsigned)) +
               align_up(sep_exact.size() * sizeof(double)) + align_up(self_kind.size() * sizeof(int));
    }
};

struct DeviceSpecies {
    PairParam* table;
    unsigned* kind_rows;
    double* sep_exact;
    int* self_kind;
};

static size_t species_bytes(int T)
{
    return align_up((size_t)T * T * sizeof(PairParam)) + align_up(T * sizeof(unsigned)) +
           align_up((size_t)T * T * sizeof(double)) + align_up(T * sizeof(int));
}

// parameter_matrix: float64 [5][T][T] as built by Simulation.__init__ (simulation.py:54-98).
static int build_species(const double* pm, int T, HostSpecies* out)
{
    if (!pm || T < 1 || T > GOF_MAX_TYPES) return fail(GOF_E_INVALID, "num_types must be in [1, 16]");
    out->T = T;
    out->has_boids = out->has_clusters = false;
    out->table.assign((size_t)T * T, PairParam{});
    out->kind_rows.assign(T, 0u);
    out->sep_exact.assign((size_t)T * T, 0.0);
    out->self_kind.assign(T, 0);
    const int TT = T * T;
    for (int i = 0; i < T; ++i)
        for (int j = 0; j < T; ++j) {
            const int q = i * T + j;
            const double a0 = pm[0 * TT + q], a1 = pm[1 * TT + q], a2 = pm[2 * TT + q], a3 = pm[3 * TT + q];
            const double kind = pm[4 * TT + q];
            PairParam p{};
            if (kind != 1.0) out->has_clusters = true;
            if (kind == 1.0) {
                out->has_boids = true;
                out->kind_rows[i] |= 1u << j;
                p.a = make_float4((float)(a0 / 5), (float)a1, (float)a2, (float)a3);
                p.b = make_float4(0.f, 0.f, 0.f, 0.f);
                out->sep_exact[q] = a0 / 5;
            } else if (kind == 0.0) {
                // clusters(): r_min, r_max, f_min, f_max
                const double s1 = 2 * a3 / (a1 - a0);
                p.a = make_float4((float)a0, (float)(a2 / a0), (float)a2, (float)((a0 + a1) / 2));
                // f_tri(d) = s1 * min(d - r_min, r_max - d) = s1 * ((r_max - r_min)/2 - |d - mid|)
                p.b = make_float4((float)s1, (float)(s1 * (a1 - a0) / 2), 0.f, 0.f);
            } else {
                // the reference ignores any other value (neither branch of lines 339 / 350 fires)
                p.a = make_float4(-1.f, 0.f, 0.f, 0.f);
                p.b = make_float4(0.f, 0.f, 0.f, 0.f);
            }
            out->table[q] = p;
            if (i == j) out->self_kind[i] = (kind == 0.0) ? 0 : 1;
        }
    return 0;
}

static DeviceSpecies carve_species(char* base, int T)
{
    DeviceSpecies d;
    d.table = reinterpret_cast<PairParam*>(base);
    base += align_up((size_t)T * T * sizeof(PairParam));
    d.kind_rows = reinterpret_cast<unsigned*>(base);
    base += align_up(T * sizeof(unsigned));
    d.sep_exact = reinterpret_cast<double*>(base);
    base += align_up((size_t)T * T * sizeof(double));
    d.self_kind = reinterpret_cast<int*>(base);
    return d;
}

static int upload_species(const HostSpecies& h, const DeviceSpecies& d, cudaStream_t st)
{
    // pageable sources: the runtime stages them before returning, so the vectors may change afterwards
    GOF_CUDA(cudaMemcpyAsync(d.table, h.table.data(), h.table.size() * sizeof(PairParam), cudaMemcpyHostToDevice, st));
    GOF_CUDA(cudaMemcpyAsync(d.kind_rows, h.kind_rows.data(), h.kind_rows.size() * sizeof(unsigned), cudaMemcpyHostToDevice, st));
    GOF_CUDA(cudaMemcpyAsync(d.sep_exact, h.sep_exact.data(), h.sep_exact.size() * sizeof(double), cudaMemcpyHostToDevice, st));
    GOF_CUDA(cudaMemcpyAsync(d.self_kind, h.self_kind.data(), h.self_kind.size() * sizeof(int), cudaMemcpyHostToDevice, st));
    return 0;
}

static size_t force_smem_bytes(int T, int kind, int split)
{
    return (size_t)T * T * sizeof(PairParam) + (size_t)T * kRowBytes +
           (kind != kKindClusters ? (size_t)T * kBoidSlots * kForceThreads * sizeof(float) : 0) +
           (kind == kKindBoids ? (size_t)kQueueDepth * kForceThreads * sizeof(unsigned) : 0) +
           (kind == kKindMixed ? (size_t)(kForceThreads / 32) * kListLen * 2 * sizeof(float4) : 0) +
           (split > 1 ? (size_t)kSplitRuns * 3 * kSplitHomes * sizeof(float) : 0) +
           (kind == kKindClusters ? (size_t)(split > 1 ? split : kForceThreads / 32) * kWalkCap * (sizeof(float4) + sizeof(int)) : 0);
}

// cudaFuncSetAttribute is per device: remember per kernel instantiation which devices have it
// (a process may drive engines on several GPUs, from several threads).
static bool attr_needed(std::atomic<unsigned long long>& done)
{
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return true;
    return !((done.load(std::memory_order_relaxed) >> dev) & 1ull);
}
static void attr_done(std::atomic<unsigned long long>& done)
{
    int dev = 0;
    if (cudaGetDevice(&dev) == cudaSuccess && dev >= 0 && dev < 64) done.fetch_or(1ull << dev, std::memory_order_relaxed);
}

template <int KIND, int MODE, int SPLIT>
static int launch_force_t(const Geometry& g, const ForceArgs& a, cudaStream_t st)
{
    const size_t smem = force_smem_bytes(a.T, KIND, SPLIT);
    static std::atomic<unsigned long long> attr_set{0};  // per instantiation, one bit per device
    if (attr_needed(attr_set)) {
        GOF_CUDA(cudaFuncSetAttribute(k_force<KIND, MODE, SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      100 * 1024));
        attr_done(attr_set);
    }
    const int homes = SPLIT == 1 ? kForceThreads : kSplitHomes;
    GOF_LAUNCH((k_force<KIND, MODE, SPLIT>), cdiv(a.home_count + a.home_count2, homes), homes * SPLIT, smem, st, g, a);
    return 0;
}

// How many warps share one home particle's stencil (k_force's SPLIT).  Small Clusters-kind
// systems are latency-bound with one thread per particle: split the 9 (3 in 2D) stencil rows over
// warps while the launch still fits the GPU a few times over.  Results do not depend on it.
static int g_force_split = -1;  // -1: automatic

static int pick_split(const Geometry& g, int homes, bool boids)
{
    if (boids) return 1;
    if (g_force_split > 0) return (!g.is3d && g_force_split == 9) ? 3 : g_force_split;
    // thresholds measured on B200 (step time against N for every split, uniform Clusters systems
    // at the reference density): 3D -- 9 warps up to ~150k particles, 3 warps up to ~350k;
    // 2D -- 3 warps up to ~50k
    const long warps = cdiv(homes, 32);
    const long kResidentWarps = 20L * sm_count();  // (thresholds measured on 148 SMs)
    if (g.is3d) {
        if (warps * 9 <= 14L * kResidentWarps) return 9;
        if (warps * 3 <= 11L * kResidentWarps) return 3;
        return 1;
    }
    return 5 * warps * 3 <= 8L * kResidentWarps ? 3 : 1;
}

extern "C" int gof_set_force_split(int split)
{
    if (split != -1 && split != 1 && split != 3 && split != 9)
        return fail(GOF_E_INVALID, "gof_set_force_split: expected -1 (automatic), 1, 3 or 9");
    g_force_split = split;
    return 0;
}

template <int MODE>
static int launch_force_m(const Geometry& g, const ForceArgs& a, int kind, cudaStream_t st)
{
    if (kind == kKindBoids) {
        // queue entries hold a slot index in kQueueSlotBits bits (slots: owned + halo particles)
        if (a.slot_limit > (1 << kQueueSlotBits))
            return fail(GOF_E_INVALID, "at most 2^26 particle slots per engine while a species pair is Boids-kind");
        return launch_force_t<kKindBoids, MODE, 1>(g, a, st);
    }
    if (kind == kKindMixed) return launch_force_t<kKindMixed, MODE, 1>(g, a, st);
    // (a slab launches its interior and its two boundary layers apart, side by side: what decides
    // whether the GPU is full is the slab's population, not the 5 % of it in the boundary launch --
    // which used to get the 9-warp kernel of small systems, three times the SM time per particle)
    switch (pick_split(g, a.split_homes > 0 ? a.split_homes : a.home_count + a.home_count2, false)) {
        case 9: return launch_force_t<kKindClusters, MODE, 9>(g, a, st);
        case 3: return launch_force_t<kKindClusters, MODE, 3>(g, a, st);
        default: return launch_force_t<kKindClusters, MODE, 1>(g, a, st);
    }
}

// a.home_count is the number of threads to launch (an upper bound when a.interior_words is set)
static int launch_force(const Geometry& g, const ForceArgs& a0, int kind, int mode, cudaStream_t st)
{
    if (a0.home_count + a0.home_count2 <= 0) return 0;
    ForceArgs a = a0;
    a.wave_sms = sm_count();
    if (mode == kModeIntegrate) return launch_force_m<kModeIntegrate>(g, a, kind, st);
    return launch_force_m<kModeScatter>(g, a, kind, st);
}

// exclusive scan of m ints; tile_sums needs cdiv(m, kScanTile) ints
static int launch_scan(const int* in, int m, int* tile_sums, int* out, cudaStream_t st)
{
    const int tiles = cdiv(m, kScanTile);
    if (tiles <= 1) {
        GOF_LAUNCH(k_scan_apply, 1, kScanThreads, 0, st, in, m, (const int*)nullptr, out);
        return 0;
    }
    GOF_LAUNCH(k_scan_tile_sums, tiles, kScanThreads, 0, st, in, m, tile_sums);
    GOF_LAUNCH(k_scan_tile_prefix, 1, kScanThreads, 0, st, tile_sums, tiles);
    GOF_LAUNCH(k_scan_apply, tiles, kScanThreads, 0, st, in, m, (const int*)tile_sums, out);
    return 0;
}

// --------------------------------------------------------------------------------------
// small element-wise kernels
// --------------------------------------------------------------------------------------
namespace gof {

constexpr int kEwThreads = 256;

// integrator.step1 on SoA arrays (main/integrator.py:33-55)
__global__ void __launch_bounds__(kEwThreads)
k_step1(float* __restrict__ vx, float* __restrict__ vy, float* __restrict__ vz,
        const float* __restrict__ ax, const float* __restrict__ ay, const float* __restrict__ az,
        int n, float dt)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float h = 0.5f * dt;
    float x = fmaf(ax[i], h, vx[i]), y = fmaf(ay[i], h, vy[i]), z = fmaf(az[i], h, vz[i]);
    const float sq = x * x + y * y + z * z;
    if (sq > kMaxSpeed * kMaxSpeed) {
        const float k = kMaxSpeed / sqrtf(sq);
        x *= k; y *= k; z *= k;
    }
    vx[i] = x; vy[i] = y; vz[i] = z;
}

// integrator.step2 on SoA arrays (main/integrator.py:58-88)
__global__ void __launch_bounds__(kEwThreads)
k_step2(float* __restrict__ px, float* __restrict__ py, float* __restrict__ pz,
        float* __restrict__ vx, float* __restrict__ vy, float* __restrict__ vz,
        const float* __restrict__ ax, const float* __restrict__ ay, const float* __restrict__ az,
        const int* __restrict__ tia, const int* __restrict__ self_kind, int n, float dt)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float a0 = ax[i], a1 = ay[i], a2 = az[i];
    float v0 = vx[i], v1 = vy[i], v2 = vz[i];
    px[i] += (v0 + 0.5f * a0 * dt) * dt;
    py[i] += (v1 + 0.5f * a1 * dt) * dt;
    pz[i] += (v2 + 0.5f * a2 * dt) * dt;
    v0 = fmaf(a0, 0.5f * dt, v0);
    v1 = fmaf(a1, 0.5f * dt, v1);
    v2 = fmaf(a2, 0.5f * dt, v2);
    if (self_kind[tia[i]] == 0) { v0 *= kDamping; v1 *= kDamping; v2 *= kDamping; }
    vx[i] = v0; vy[i] = v1; vz[i] = v2;
}

// integrator.boundary_condition (main/integrator.py:91-114)
__global__ void __launch_bounds__(kEwThreads)
k_boundary(float* __restrict__ px, float* __restrict__ py, float* __restrict__ pz, float Lx,
           float Ly, float Lz, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float x = px[i], y = py[i], z = pz[i];
    if (x < 0.0f) x += Lx; else if (x > Lx) x -= Lx;
    if (y < 0.0f) y += Ly; else if (y > Ly) y -= Ly;
    if (z < 0.0f) z += Lz; else if (z > Lz) z -= Lz;
    px[i] = x; py[i] = y; pz[i] = z;
}

// set_sq_speed + max_reduction (main/integrator.py:20-30, 117-129): grid-stride, one atomic per warp
__global__ void __launch_bounds__(kEwThreads)
k_max_sq_speed(const float* __restrict__ vx, const float* __restrict__ vy,
               const float* __restrict__ vz, int n, unsigned* __restrict__ out_bits)
{
    float m = 0.0f;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float s = vx[i] * vx[i] + vy[i] * vy[i] + vz[i] * vz[i];
        m = fmaxf(m, s);
    }
    const unsigned w = __reduce_max_sync(0xffffffffu, __float_as_uint(m));
    if (lane_id() == 0) atomicMax(out_bits, w);
}

// SoA (particle order) -> engine float4 state; also the maximum squared speed of the upload
__global__ void __launch_bounds__(kEwThreads)
k_pack(const float* __restrict__ px, const float* __restrict__ py, const float* __restrict__ pz,
       const float* __restrict__ vx, const float* __restrict__ vy, const float* __restrict__ vz,
       const float* __restrict__ ax, const float* __restrict__ ay, const float* __restrict__ az,
       const int* __restrict__ types, int n, float4* __restrict__ posw, float4* __restrict__ veli,
       float4* __restrict__ acc, unsigned* __restrict__ max_bits)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    float sq = 0.0f;
    if (i < n) {
        const float4 v = make_float4(vx ? vx[i] : 0.f, vy ? vy[i] : 0.f, vz ? vz[i] : 0.f, __int_as_float(i));
        posw[i] = make_float4(px[i], py[i], pz[i], __int_as_float(types[i]));
        veli[i] = v;
        acc[i] = make_float4(ax ? ax[i] : 0.f, ay ? ay[i] : 0.f, az ? az[i] : 0.f, 0.f);
        sq = v.x * v.x + v.y * v.y + v.z * v.z;
    }
    const unsigned w = __reduce_max_sync(0xffffffffu, __float_as_uint(sq));
    if (lane_id() == 0) atomicMax(max_bits, w);
}

// engine float4 state (cell order) -> SoA in particle order; null outputs are skipped
__global__ void __launch_bounds__(kEwThreads)
k_unpack(const float4* __restrict__ posw, const float4* __restrict__ veli,
         const float4* __restrict__ acc, int n, float* __restrict__ px, float* __restrict__ py,
         float* __restrict__ pz, float* __restrict__ vx, float* __restrict__ vy,
         float* __restrict__ vz, float* __restrict__ ax, float* __restrict__ ay,
         float* __restrict__ az)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const float4 v = veli[k];
    const int id = __float_as_int(v.w);
    if (px) { const float4 p = posw[k]; px[id] = p.x; py[id] = p.y; pz[id] = p.z; }
    if (vx) { vx[id] = v.x; vy[id] = v.y; vz[id] = v.z; }
    if (ax) { const float4 a = acc[k]; ax[id] = a.x; ay[id] = a.y; az[id] = a.z; }
}

// Simulation.update's output array (simulation.py:300-304): [N][3] in particle order
__global__ void __launch_bounds__(kEwThreads)
k_snapshot(const float4* __restrict__ posw, const float4* __restrict__ veli, int n,
           float* __restrict__ xyz)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const float4 p = posw[k];
    const int id = __float_as_int(veli[k].w);
    xyz[3 * (size_t)id + 0] = p.x;
    xyz[3 * (size_t)id + 1] = p.y;
    xyz[3 * (size_t)id + 2] = p.z;
}

// bins of the last step in the reference's layout
__global__ void __launch_bounds__(kEwThreads)
k_export_bins(const float4* __restrict__ veli_sorted, const int* __restrict__ cell_sorted,
              const int* __restrict__ start, const int* __restrict__ pstart,
              const int* __restrict__ count, int num_cells, int* __restrict__ particle_bins,
              int* __restrict__ particle_indices)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= pstart[num_cells]) return;
    const int c = cell_sorted[k];
    const int r = k - pstart[c];
    if (r >= count[c]) return;
    const int id = __float_as_int(veli_sorted[k].w);
    if (particle_bins) particle_bins[id] = c;
    if (particle_indices) {
        // the reference's order inside a bin is by particle id; the engine's is by x
        int by_id = 0;
        for (int t = pstart[c], e = pstart[c] + count[c]; t < e; ++t)
            by_id += __float_as_int(veli_sorted[t].w) < id;
        particle_indices[start[c] + by_id] = id;
    }
}

// reference-layout inputs -> cell-sorted float4 scratch for gof_accelerator
__global__ void __launch_bounds__(kEwThreads)
k_gather_sorted(const float* __restrict__ px, const float* __restrict__ py,
                const float* __restrict__ pz, const float* __restrict__ vx,
                const float* __restrict__ vy, const float* __restrict__ vz,
                const int* __restrict__ types, const int* __restrict__ particle_bins,
                const int* __restrict__ particle_indices, int n, float4* __restrict__ posw,
                float4* __restrict__ veli, int* __restrict__ cell_sorted)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const int id = particle_indices[k];
    posw[k] = make_float4(px[id], py[id], pz[id], __int_as_float(types[id]));
    veli[k] = make_float4(vx[id], vy[id], vz[id], __int_as_float(id));
    cell_sorted[k] = particle_bins[id];
}

// The timestep of the step being enqueued (integrator.py:174-188) and the reset of the
// accumulator the force kernel's epilogue will fill for the next one.
__global__ void k_prepare_dt(StepScalars* s, int parity, float fixed_dt)
{
    float dt = fixed_dt;
    if (fixed_dt < 0.0f) {
        const double m = sqrt((double)__uint_as_float(s->max_sq_speed_bits[parity]));
        dt = (m > 1e-6) ? (float)(0.4 / m) : 0.1f;
    }
    s->dt = dt;
    s->max_sq_speed_bits[parity ^ 1] = 0u;
}

__global__ void k_bits_to_float(const unsigned* bits, float* out) { *out = __uint_as_float(*bits); }
__global__ void k_set_word(int* p, int v) { *p = v; }

}  // namespace gof

// --------------------------------------------------------------------------------------
// reference-layout entry points
// --------------------------------------------------------------------------------------
extern "C" size_t gof_workspace_bytes(int n, int num_bins, int num_types)
{
    if (n < 0) n = 0;
    size_t b = 0;
    b += 2 * align_up((size_t)n * sizeof(float4));        // posw, veli (accelerator)
    b += 3 * align_up((size_t)n * sizeof(int));           // slot ids / cell_sorted / spare
    b += align_up((size_t)(cdiv(num_bins > 0 ? num_bins : 1, kScanTile) + 1) * sizeof(int));
    b += species_bytes(num_types > 0 ? num_types : 1);
    b += 256;
    return b;
}

extern "C" int gof_setup_bins(const float* d_pos_x, const float* d_pos_y, const float* d_pos_z,
                              int n, const gof_grid* grid, int num_bins, int32_t* d_particle_bins,
                              int32_t* d_counts, int32_t* d_bin_offsets, int32_t* d_starts,
                              int32_t* d_particle_indices, void* d_workspace, size_t workspace_bytes,
                              void* stream)
{
    DeviceGuard guard(device_of(d_pos_x));  // the arrays' device, whatever the caller's current one
    cudaStream_t st = (cudaStream_t)stream;
    Geometry g;
    if (int rc = make_geometry(grid, 0, 0, &g)) return rc;
    if (n < 0 || num_bins < g.num_cells) return fail(GOF_E_INVALID, "gof_setup_bins: num_bins < num_cells");
    if (!d_pos_x || !d_pos_y || !d_pos_z || !d_particle_bins || !d_counts || !d_bin_offsets || !d_starts ||
        !d_particle_indices || !d_workspace)
        return fail(GOF_E_INVALID, "gof_setup_bins: null pointer");
    if (workspace_bytes < gof_workspace_bytes(n, num_bins, 1)) return fail(GOF_E_NOMEM, "gof_setup_bins: workspace too small");
    char* w = (char*)d_workspace;
    int* slot_id = (int*)w;
    w += align_up((size_t)n * sizeof(int));
    int* tile_sums = (int*)w;

    GOF_CUDA(cudaMemsetAsync(d_counts, 0, (size_t)num_bins * sizeof(int), st));
    if (n > 0) {
        MigrateOut none{};
        GOF_LAUNCH(k_hash_count<true>, cdiv(n, kBinThreads), kBinThreads, 0, st, (const float4*)nullptr,
                   (const float4*)nullptr, (const float4*)nullptr, d_pos_x, d_pos_y, d_pos_z, n, (const int*)nullptr, g,
                   d_particle_bins, d_bin_offsets, d_counts, none);
    }
    if (int rc = launch_scan(d_counts, num_bins, tile_sums, d_starts, st)) return rc;
    if (n > 0) {
        GOF_LAUNCH(k_scatter_ids<true>, cdiv(n, kBinThreads), kBinThreads, 0, st, (const float4*)nullptr, n,
                   (const int*)nullptr, d_particle_bins, d_bin_offsets, d_starts, slot_id, (const float4*)nullptr,
                   (unsigned*)nullptr);
        GOF_LAUNCH(k_rank_indices, cdiv(n, kBinThreads), kBinThreads, 0, st, n, d_particle_bins, d_starts,
                   d_counts, slot_id, d_bin_offsets, d_particle_indices);
    }
    return 0;
}

extern "C" int gof_accelerator(const float* d_pos_x, const float* d_pos_y, const float* d_pos_z,
                               const float* d_vel_x, const float* d_vel_y, const float* d_vel_z, int n,
                               const gof_grid* grid, const double* pm_host, int num_types,
                               const int32_t* d_types, float* d_acc_x, float* d_acc_y, float* d_acc_z,
                               float* d_boid_acc_x, float* d_boid_acc_y, float* d_boid_acc_z,
                               float* d_boid_vel_x, float* d_boid_vel_y, float* d_boid_vel_z,
                               int32_t* d_boid_counts, const int32_t* d_particle_bins,
                               const int32_t* d_particle_indices, const int32_t* d_bin_starts,
                               const int32_t* d_bin_counts, int wrap_mode, void* d_workspace,
                               size_t workspace_bytes, void* stream)
{
    DeviceGuard guard(device_of(d_pos_x));  // the arrays' device, whatever the caller's current one
    cudaStream_t st = (cudaStream_t)stream;
    Geometry g;
    if (int rc = make_geometry(grid, 0, 0, &g)) return rc;
    HostSpecies hs;
    if (int rc = build_species(pm_host, num_types, &hs)) return rc;
    if (!d_pos_x || !d_pos_y || !d_pos_z || !d_vel_x || !d_vel_y || !d_vel_z || !d_types || !d_acc_x ||
        !d_acc_y || !d_acc_z || !d_particle_bins || !d_particle_indices || !d_bin_starts || !d_bin_counts ||
        !d_workspace)
        return fail(GOF_E_INVALID, "gof_accelerator: null pointer");
    const bool boid_out = d_boid_counts != nullptr;
    if (boid_out && (!d_boid_acc_x || !d_boid_acc_y || !d_boid_acc_z || !d_boid_vel_x || !d_boid_vel_y || !d_boid_vel_z))
        return fail(GOF_E_INVALID, "gof_accelerator: boid outputs must be given together");
    if (wrap_mode != GOF_WRAP_REFERENCE && wrap_mode != GOF_WRAP_MINIMAGE)
        return fail(GOF_E_INVALID, "gof_accelerator: bad wrap_mode");
    if (workspace_bytes < gof_workspace_bytes(n, g.num_cells, num_types))
        return fail(GOF_E_NOMEM, "gof_accelerator: workspace too small");
    if (n <= 0) return 0;

    char* w = (char*)d_workspace;
    float4* posw = (float4*)w;
    w += align_up((size_t)n * sizeof(float4));
    float4* veli = (float4*)w;
    w += align_up((size_t)n * sizeof(float4));
    int* cell_sorted = (int*)w;
    w += 3 * align_up((size_t)n * sizeof(int));
    w += align_up((size_t)(cdiv(g.num_cells, kScanTile) + 1) * sizeof(int));
    DeviceSpecies ds = carve_species(w, num_types);
    if (int rc = upload_species(hs, ds, st)) return rc;

    if (boid_out) {
        const size_t nt = (size_t)n * num_types;
        GOF_CUDA(cudaMemsetAsync(d_boid_counts, 0, nt * sizeof(int), st));
        float* fs[6] = {d_boid_acc_x, d_boid_acc_y, d_boid_acc_z, d_boid_vel_x, d_boid_vel_y, d_boid_vel_z};
        for (float* f : fs) GOF_CUDA(cudaMemsetAsync(f, 0, nt * sizeof(float), st));
    }
    GOF_LAUNCH(k_gather_sorted, cdiv(n, kEwThreads), kEwThreads, 0, st, d_pos_x, d_pos_y, d_pos_z, d_vel_x,
               d_vel_y, d_vel_z, d_types, d_particle_bins, d_particle_indices, n, posw, veli, cell_sorted);
    ForceArgs a{};
    a.posw = posw;
    a.veli = veli;
    a.cell_sorted = cell_sorted;
    a.start = d_bin_starts;
    a.pstart = d_bin_starts;
    a.count = d_bin_counts;
    a.home_begin = 0;
    a.home_count = n;
    a.table = ds.table;
    a.kind_rows = ds.kind_rows;
    a.sep_exact = ds.sep_exact;
    a.self_kind = ds.self_kind;
    a.T = num_types;
    a.wrap_mode = wrap_mode;
    a.acc_x = d_acc_x;
    a.acc_y = d_acc_y;
    a.acc_z = d_acc_z;
    a.boid_counts = d_boid_counts;
    a.boid_acc_x = d_boid_acc_x;
    a.boid_acc_y = d_boid_acc_y;
    a.boid_acc_z = d_boid_acc_z;
    a.boid_vel_x = d_boid_vel_x;
    a.boid_vel_y = d_boid_vel_y;
    a.boid_vel_z = d_boid_vel_z;
    a.n_total = n;
    a.slot_limit = n;
    return launch_force(g, a, hs.kind(), kModeScatter, st);
}

extern "C" int gof_step1(float* vx, float* vy, float* vz, const float* ax, const float* ay, const float* az,
                         int n, float timestep, void* stream)
{
    DeviceGuard guard(device_of(vx));  // the arrays' device, whatever the caller's current one
    if (n <= 0) return 0;
    if (!vx || !vy || !vz || !ax || !ay || !az) return fail(GOF_E_INVALID, "gof_step1: null pointer");
    GOF_LAUNCH(k_step1, cdiv(n, kEwThreads), kEwThreads, 0, (cudaStream_t)stream, vx, vy, vz, ax, ay, az, n, timestep);
    return 0;
}

extern "C" int gof_step2(float* px, float* py, float* pz, float* vx, float* vy, float* vz, const float* ax,
                         const float* ay, const float* az, const int32_t* tia, const int32_t* self_kind, int n,
                         float timestep, void* stream)
{
    DeviceGuard guard(device_of(px));  // the arrays' device, whatever the caller's current one
    if (n <= 0) return 0;
    if (!px || !py || !pz || !vx || !vy || !vz || !ax || !ay || !az || !tia || !self_kind)
        return fail(GOF_E_INVALID, "gof_step2: null pointer");
    GOF_LAUNCH(k_step2, cdiv(n, kEwThreads), kEwThreads, 0, (cudaStream_t)stream, px, py, pz, vx, vy, vz, ax, ay,
               az, tia, self_kind, n, timestep);
    return 0;
}

extern "C" int gof_boundary_condition(float* px, float* py, float* pz, float lx, float ly, float lz, int n,
                                      void* stream)
{
    DeviceGuard guard(device_of(px));  // the arrays' device, whatever the caller's current one
    if (n <= 0) return 0;
    if (!px || !py || !pz) return fail(GOF_E_INVALID, "gof_boundary_condition: null pointer");
    GOF_LAUNCH(k_boundary, cdiv(n, kEwThreads), kEwThreads, 0, (cudaStream_t)stream, px, py, pz, lx, ly, lz, n);
    return 0;
}

extern "C" int gof_max_sq_speed(const float* vx, const float* vy, const float* vz, int n, float* d_out,
                                void* stream)
{
    DeviceGuard guard(device_of(vx));  // the arrays' device, whatever the caller's current one
    cudaStream_t st = (cudaStream_t)stream;
    if (!d_out) return fail(GOF_E_INVALID, "gof_max_sq_speed: null output");
    GOF_CUDA(cudaMemsetAsync(d_out, 0, sizeof(float), st));
    if (n <= 0) return 0;
    if (!vx || !vy || !vz) return fail(GOF_E_INVALID, "gof_max_sq_speed: null pointer");
    int blocks = cdiv(n, kEwThreads);
    if (blocks > 148 * 8) blocks = 148 * 8;
    // non-negative floats order like their bit patterns, so the float is reduced in place as a uint
    GOF_LAUNCH(k_max_sq_speed, blocks, kEwThreads, 0, st, vx, vy, vz, n, reinterpret_cast<unsigned*>(d_out));
    return 0;
}

// --------------------------------------------------------------------------------------
// the resident engine
// --------------------------------------------------------------------------------------
struct gof_engine {
    int device = 0;
    int n = 0;          // live particles
    int capacity = 0;   // slots in the per-particle arrays
    int T = 0;
    gof_grid grid{};
    Geometry geom{};
    HostSpecies species;
    bool species_dirty = true;
    bool bound = false, uploaded = false, binned = false;
    int parity = 0;  // which max_sq_speed slot describes the resident velocities

    // carved from the arena
    size_t arena_bytes = 0;
    float4 *posw[2] = {nullptr, nullptr}, *veli[2] = {nullptr, nullptr}, *acc = nullptr;
    int *cell_of = nullptr, *off = nullptr, *slot_id = nullptr, *cell_sorted = nullptr;
    unsigned* slot_key = nullptr;  // sort key inside a cell (bits of x), parallel to slot_id
    int *count = nullptr, *start = nullptr, *tile_sums = nullptr;
    int slot_capacity = 0;                      // capacity (+ the two halo layers in slab mode)
    StepScalars* scalars = nullptr;
    DeviceSpecies dspecies{};
    float* stage[9] = {};  // SoA staging: pos xyz, vel xyz, acc xyz
    int* stage_types = nullptr;
    int* stage_bins = nullptr;     // particle_bins / particle_indices export
    int* stage_indices = nullptr;
    float* stage_xyz = nullptr;    // snapshot [N][3]
    // gof_engine_step_host: a second SoA staging set for the way back, so that the next call's
    // uploads can run while this call's downloads are still leaving
    float* stage_down[9] = {};
    cudaStream_t host_side = nullptr;          // the downloads' stream
    cudaEvent_t ev_ready = nullptr;            // un-permuted state staged
    cudaEvent_t ev_down[9] = {};               // download k of the last call complete
    const void* down_ptr[9] = {};              // ... and where it goes
    bool down_pending = false;

    // slab mode (multi-GPU): this engine owns the global z-layers [layer_lo, layer_hi)
    bool slab = false;
    int layer_lo = 0, layer_hi = 0;
    int mig_capacity = 0;   // migration records per direction
    int halo_capacity = 0;  // halo particles per side
    float4 *mig_send[2] = {nullptr, nullptr};  // [0] down, [1] up
    float4 *mig_recv[2] = {nullptr, nullptr};  // [0] from the slab below, [1] from above
    int* counters = nullptr;                   // [0..2] leavers down / up / error, [4..7] sort results
    int n_slots = 0;        // resident slots before the sort (live + dead + arrivals)
    int lo_pop = 0, hi_pop = 0;  // populations of the two owned boundary layers (gof_slab_commit)

    // peer-memory transport (gof_slab_connect / gof_slab_step_p2p): nothing about a step's sizes
    // reaches the host; the neighbours' arenas are mapped into this process
    char* arena_base = nullptr;
    struct P2P {
        bool on = false;
        int rank = 0, world = 0;
        PeerSide below{}, above{};
        unsigned step = 0;          // tag of the last enqueued step
        bool need_publish = true;   // the maximum speed of the resident state has not been published yet
    } p2p;
    unsigned long long* p2p_flags = nullptr;    // [0] records from below, [1] from above, [2] halo from below, [3] from above
    unsigned long long* speed_in = nullptr;     // [2][kMaxWorld]: every slab's published maximum, per step parity
    unsigned long long** speed_ptrs = nullptr;  // [2][kMaxWorld]: where this slab publishes to (device array)
    int* push_done = nullptr;                   // [2] block counters of k_push_halo
    int* halo_n = nullptr;                      // [2] sizes of the received halo layers
    int* halo_count_in = nullptr;               // [2][cells per layer]: the neighbours' per-cell counts of their layers

    // One core_step captured as a CUDA graph per parity (8 kernels + a memset become one launch;
    // small systems are launch-bound).  Re-captured when anything baked into the nodes changes.
    struct StepGraph {
        cudaGraphExec_t exec = nullptr;
        float timestep = 0.f;
        int wrap_mode = 0, split = 0, generation = -1, launches = 0;
    } graphs[2];
    bool use_graph = true;
    int generation = 0;             // bumped by whatever invalidates captured graphs
    cudaStream_t capture_stream = nullptr;

    // optional per-launch timing of the force kernel (bench.py roofline)
    bool timing = false;
    std::vector<cudaEvent_t> ev;  // pairs: start, stop
    int ev_used = 0;
    // ... and of the phases of a peer-memory slab step: five events per step (start, sorted,
    // interior forces done, boundary forces done [side stream], end)
    std::vector<cudaEvent_t> pev;
    int pev_steps = 0;
};

static constexpr int kMaxTimedLaunches = 4096;

static size_t engine_layout(gof_engine* e, char* base)
{
    // base == nullptr: only compute the size
    size_t o = 0;
    auto take = [&](size_t bytes) -> char* {
        char* p = base ? base + o : nullptr;
        o += align_up(bytes);
        return p;
    };
    const size_t cap = (size_t)e->capacity;
    const size_t cells = (size_t)e->geom.num_cells + 2;
    // slab mode appends the two halo layers behind the owned particles
    const size_t slots = cap + 2 * (size_t)e->halo_capacity;
    e->slot_capacity = (int)slots;
    // [0] = rest state (compact, last step's cell order), [1] = this step's snapshot
    e->posw[0] = (float4*)take(cap * sizeof(float4));
    e->veli[0] = (float4*)take(cap * sizeof(float4));
    e->posw[1] = (float4*)take(slots * sizeof(float4));
    e->veli[1] = (float4*)take(slots * sizeof(float4));
    e->acc = (float4*)take(cap * sizeof(float4));
    e->cell_of = (int*)take(cap * sizeof(int));
    e->off = (int*)take(cap * sizeof(int));
    e->slot_id = (int*)take(slots * sizeof(int));
    e->slot_key = (unsigned*)take(slots * sizeof(unsigned));
    e->cell_sorted = (int*)take(slots * sizeof(int));
    e->count = (int*)take(cells * sizeof(int));
    e->start = (int*)take(cells * sizeof(int));
    e->tile_sums = (int*)take((size_t)(cdiv((int)cells, kScanTile) + 1) * sizeof(int));
    e->scalars = (StepScalars*)take(sizeof(StepScalars));
    char* sp = take(species_bytes(e->T));
    if (base) e->dspecies = carve_species(sp, e->T);
    for (int k = 0; k < 9; ++k) e->stage[k] = (float*)take(cap * sizeof(float));
    e->stage_types = (int*)take(cap * sizeof(int));
    e->stage_bins = (int*)take(cap * sizeof(int));
    e->stage_indices = (int*)take(cap * sizeof(int));
    e->stage_xyz = (float*)take(cap * 3 * sizeof(float));
    if (!e->slab)
        for (int k = 0; k < 9; ++k) e->stage_down[k] = (float*)take(cap * sizeof(float));
    if (e->slab) {
        for (int d = 0; d < 2; ++d) {
            e->mig_send[d] = (float4*)take(((size_t)e->mig_capacity * 3 + 1) * sizeof(float4));
            e->mig_recv[d] = (float4*)take(((size_t)e->mig_capacity * 3 + 1) * sizeof(float4));
        }
        e->counters = (int*)take(8 * sizeof(int));
        // one block that peers write into / that holds peer pointers (offsets are part of gof_slab_window)
        char* blk = take(1024);
        e->p2p_flags = (unsigned long long*)blk;
        e->speed_in = (unsigned long long*)(blk + 64);
        e->speed_ptrs = (unsigned long long**)(blk + 64 + 2 * kMaxWorld * 8);
        e->push_done = (int*)(blk + 64 + 4 * kMaxWorld * 8);
        e->halo_n = e->push_done + 2;
        e->halo_count_in = (int*)take(2 * (size_t)e->geom.nbx * e->geom.nby * sizeof(int));
    }
    return o;
}

extern "C" int gof_engine_create(gof_engine** out, int device, int num_particles, int num_types,
                                 const gof_grid* grid, const double* pm_host)
{
    if (!out || num_particles < 1 || !grid) return fail(GOF_E_INVALID, "gof_engine_create: bad arguments");
    gof_engine* e = new (std::nothrow) gof_engine();
    if (!e) return fail(GOF_E_NOMEM, "gof_engine_create: host allocation failed");
    e->device = device;
    e->n = e->capacity = num_particles;
    e->T = num_types;
    e->grid = *grid;
    int rc = make_geometry(grid, 0, 0, &e->geom);
    if (!rc) rc = build_species(pm_host, num_types, &e->species);
    if (rc) {
        delete e;
        return rc;
    }
    e->arena_bytes = engine_layout(e, nullptr);
    *out = e;
    return 0;
}

extern "C" void gof_engine_destroy(gof_engine* e)
{
    if (!e) return;
    DeviceGuard guard(e->device);
    for (cudaEvent_t v : e->ev) cudaEventDestroy(v);
    for (cudaEvent_t v : e->pev) cudaEventDestroy(v);
    for (auto& g : e->graphs)
        if (g.exec) cudaGraphExecDestroy(g.exec);
    if (e->capture_stream) cudaStreamDestroy(e->capture_stream);
    if (e->host_side) {
        cudaStreamSynchronize(e->host_side);
        cudaStreamDestroy(e->host_side);
    }
    if (e->ev_ready) cudaEventDestroy(e->ev_ready);
    for (cudaEvent_t v : e->ev_down)
        if (v) cudaEventDestroy(v);
    delete e;
}

extern "C" int gof_engine_set_graph(gof_engine* e, int enable)
{
    if (!e) return fail(GOF_E_INVALID, "null engine");
    e->use_graph = enable != 0;
    return 0;
}

extern "C" int gof_engine_timing(gof_engine* e, int enable)
{
    if (!e) return fail(GOF_E_INVALID, "null engine");
    DeviceGuard guard(e->device);
    e->timing = enable != 0;
    e->ev_used = 0;
    e->pev_steps = 0;
    return 0;
}

extern "C" int gof_engine_force_time(gof_engine* e, float* total_ms, int* launches)
{
    if (!e || !total_ms || !launches) return fail(GOF_E_INVALID, "gof_engine_force_time: null pointer");
    DeviceGuard guard(e->device);
    float sum = 0.f;
    for (int k = 0; k < e->ev_used; ++k) {
        float ms = 0.f;
        GOF_CUDA(cudaEventSynchronize(e->ev[2 * k + 1]));
        GOF_CUDA(cudaEventElapsedTime(&ms, e->ev[2 * k], e->ev[2 * k + 1]));
        sum += ms;
    }
    *total_ms = sum;
    *launches = e->ev_used;
    return 0;
}

extern "C" size_t gof_engine_arena_bytes(const gof_engine* e) { return e ? e->arena_bytes : 0; }

extern "C" int gof_engine_bind_arena(gof_engine* e, void* d_arena, size_t bytes)
{
    if (!e || !d_arena) return fail(GOF_E_INVALID, "gof_engine_bind_arena: null pointer");
    if (bytes < e->arena_bytes) return fail(GOF_E_NOMEM, "gof_engine_bind_arena: arena too small");
    if ((uintptr_t)d_arena % 256) return fail(GOF_E_INVALID, "gof_engine_bind_arena: arena must be 256-byte aligned");
    DeviceGuard guard(e->device);
    engine_layout(e, (char*)d_arena);
    e->arena_base = (char*)d_arena;
    e->p2p = gof_engine::P2P{};
    e->bound = true;
    e->uploaded = e->binned = false;
    e->species_dirty = true;
    ++e->generation;
    return 0;
}

extern "C" int gof_engine_set_parameters(gof_engine* e, const double* pm_host, void* stream)
{
    if (!e) return fail(GOF_E_INVALID, "null engine");
    if (int rc = build_species(pm_host, e->T, &e->species)) return rc;
    e->species_dirty = true;
    ++e->generation;  // the kinds decide which kernel instantiation a step launches
    if (e->bound) {
        DeviceGuard guard(e->device);
        if (int rc = upload_species(e->species, e->dspecies, (cudaStream_t)stream)) return rc;
        e->species_dirty = false;
    }
    return 0;
}

static int engine_ready(gof_engine* e, bool need_state)
{
    if (!e) return fail(GOF_E_INVALID, "null engine");
    if (!e->bound) return fail(GOF_E_STATE, "engine has no arena: call gof_engine_bind_arena first");
    if (need_state && !e->uploaded) return fail(GOF_E_STATE, "engine has no state: call gof_engine_upload first");
    return 0;
}

extern "C" int gof_engine_upload(gof_engine* e, const float* px, const float* py, const float* pz,
                                 const float* vx, const float* vy, const float* vz, const float* ax,
                                 const float* ay, const float* az, const int32_t* types, void* stream)
{
    DeviceGuard guard(e ? e->device : -1);
    if (int rc = engine_ready(e, false)) return rc;
    if (e->slab) return fail(GOF_E_STATE, "gof_engine_upload: a slab engine takes its particles through gof_slab_upload");
    if (!px || !py || !pz || !types) return fail(GOF_E_INVALID, "gof_engine_upload: positions and types are required");
    cudaStream_t st = (cudaStream_t)stream;
    const int n = e->capacity;
    const float* src[9] = {px, py, pz, vx, vy, vz, ax, ay, az};
    for (int k = 0; k < 9; ++k)
        if (src[k]) GOF_CUDA(cudaMemcpyAsync(e->stage[k], src[k], (size_t)n * sizeof(float), cudaMemcpyDefault, st));
    GOF_CUDA(cudaMemcpyAsync(e->stage_types, types, (size_t)n * sizeof(int), cudaMemcpyDefault, st));
    if (e->species_dirty) {
        if (int rc = upload_species(e->species, e->dspecies, st)) return rc;
        e->species_dirty = false;
    }
    GOF_CUDA(cudaMemsetAsync(e->scalars, 0, sizeof(StepScalars), st));
    e->parity = 0;
    e->n = n;
    GOF_LAUNCH(k_pack, cdiv(n, kEwThreads), kEwThreads, 0, st, e->stage[0], e->stage[1], e->stage[2],
               vx ? e->stage[3] : nullptr, vy ? e->stage[4] : nullptr, vz ? e->stage[5] : nullptr,
               ax ? e->stage[6] : nullptr, ay ? e->stage[7] : nullptr, az ? e->stage[8] : nullptr,
               e->stage_types, n, e->posw[0], e->veli[0], e->acc, &e->scalars->max_sq_speed_bits[0]);
    e->uploaded = true;
    e->binned = false;
    return 0;
}

// The order inside a cell: by x (see k_rank_gather_kick), or by id when built with GOF_NO_XSORT.
static unsigned* engine_keys(gof_engine* e)
{
#ifdef GOF_NO_XSORT
    (void)e;
    return nullptr;
#else
    return e->slot_key;
#endif
}

// setup_bins on the resident state: posw[0]/veli[0]/acc (rest order) -> posw[1]/veli[1]
// (cell order, velocities kicked by step1 when `kick`), cell tables filled.
static int engine_bin(gof_engine* e, bool kick, cudaStream_t st)
{
    const Geometry& g = e->geom;
    const int n = e->n;
    GOF_CUDA(cudaMemsetAsync(e->count, 0, (size_t)(g.num_cells + 2) * sizeof(int), st));
    MigrateOut none{};
    GOF_LAUNCH(k_hash_count<false>, cdiv(n, kBinThreads), kBinThreads, 0, st, e->posw[0], e->veli[0], e->acc,
               (const float*)nullptr, (const float*)nullptr, (const float*)nullptr, n, (const int*)nullptr, g,
               e->cell_of, e->off, e->count, none);
    if (int rc = launch_scan(e->count, g.num_cells + 1, e->tile_sums, e->start, st)) return rc;
    const int* layout = e->start;
    GOF_LAUNCH(k_scatter_ids<false>, cdiv(n, kBinThreads), kBinThreads, 0, st, e->veli[0], n, (const int*)nullptr,
               e->cell_of, e->off, layout, e->slot_id, e->posw[0], engine_keys(e));
    GOF_LAUNCH(k_rank_gather_kick, cdiv(n, kBinThreads), kBinThreads, 0, st, n, (const int*)nullptr, e->cell_of,
               layout, e->count,
               e->slot_id, engine_keys(e), g.num_cells, e->posw[0], e->veli[0], e->acc, e->posw[1], e->veli[1], e->cell_sorted,
               e->scalars, kick ? 1 : 0);
    e->binned = true;
    return 0;
}

// force (+ integrate) on the snapshot engine_bin just built
static int engine_force(gof_engine* e, const ForceArgs& a, cudaStream_t st)
{
    return launch_force(e->geom, a, e->species.kind(), kModeIntegrate, st);
}

static void engine_force_args(gof_engine* e, int wrap_mode, ForceArgs* a)
{
    *a = ForceArgs{};
    a->posw = e->posw[1];
    a->veli = e->veli[1];
    a->cell_sorted = e->cell_sorted;
    a->start = e->start;
    a->pstart = e->start;
    a->count = e->count;
    a->home_begin = 0;
    a->home_count = e->n;
    a->table = e->dspecies.table;
    a->kind_rows = e->dspecies.kind_rows;
    a->sep_exact = e->dspecies.sep_exact;
    a->self_kind = e->dspecies.self_kind;
    a->T = e->T;
    a->wrap_mode = wrap_mode;
    a->posw_out = e->posw[0];
    a->veli_out = e->veli[0];
    a->acc_out = e->acc;
    a->scalars = e->scalars;
    a->n_total = e->n;
    a->slot_limit = e->slot_capacity;
    a->split_homes = e->slab ? e->capacity : 0;  // (a slab's launches cover parts of it: see launch_force_m)
}

// the launches of one core_step at the engine's current parity (which the caller flips)
static int engine_enqueue_step(gof_engine* e, float timestep, int wrap_mode, cudaStream_t st, bool timed)
{
    GOF_LAUNCH(k_prepare_dt, 1, 1, 0, st, e->scalars, e->parity, timestep < 0 ? -1.0f : timestep);
    if (int rc = engine_bin(e, true, st)) return rc;
    ForceArgs a;
    engine_force_args(e, wrap_mode, &a);
    a.integrate = 1;
    a.parity_out = e->parity ^ 1;
    if (timed) {
        while ((int)e->ev.size() < 2 * (e->ev_used + 1)) {
            cudaEvent_t v;
            GOF_CUDA(cudaEventCreate(&v));
            e->ev.push_back(v);
        }
        GOF_CUDA(cudaEventRecord(e->ev[2 * e->ev_used], st));
    }
    if (int rc = engine_force(e, a, st)) return rc;
    if (timed) {
        GOF_CUDA(cudaEventRecord(e->ev[2 * e->ev_used + 1], st));
        ++e->ev_used;
    }
    return 0;
}

// (Re)capture the step at the current parity.  Captured on the engine's own stream, so that the
// caller's stream may be the legacy default stream (which cannot be captured).
static int engine_capture_step(gof_engine* e, float timestep, int wrap_mode)
{
    gof_engine::StepGraph& sg = e->graphs[e->parity];
    if (sg.exec) {
        cudaGraphExecDestroy(sg.exec);
        sg.exec = nullptr;
    }
    if (!e->capture_stream) GOF_CUDA(cudaStreamCreateWithFlags(&e->capture_stream, cudaStreamNonBlocking));
    const uint64_t before = g_launches.load();
    GOF_CUDA(cudaStreamBeginCapture(e->capture_stream, cudaStreamCaptureModeRelaxed));
    const int rc = engine_enqueue_step(e, timestep, wrap_mode, e->capture_stream, false);
    cudaGraph_t graph = nullptr;
    const cudaError_t end = cudaStreamEndCapture(e->capture_stream, &graph);
    const int launches = (int)(g_launches.load() - before);
    g_launches.store(before);  // captured, not launched
    if (rc || end != cudaSuccess || !graph) {
        if (graph) cudaGraphDestroy(graph);
        cudaGetLastError();
        return rc ? rc : fail((int)end, std::string("step capture: ") + cudaGetErrorString(end));
    }
    const cudaError_t inst = cudaGraphInstantiate(&sg.exec, graph, 0);
    cudaGraphDestroy(graph);
    if (inst != cudaSuccess) {
        sg.exec = nullptr;
        return fail((int)inst, std::string("cudaGraphInstantiate: ") + cudaGetErrorString(inst));
    }
    sg.timestep = timestep;
    sg.wrap_mode = wrap_mode;
    sg.split = g_force_split;
    sg.generation = e->generation;
    sg.launches = launches;
    return 0;
}

extern "C" int gof_engine_step(gof_engine* e, int n_steps, float timestep, int wrap_mode, void* stream)
{
    DeviceGuard guard(e ? e->device : -1);
    if (int rc = engine_ready(e, true)) return rc;
    if (wrap_mode != GOF_WRAP_REFERENCE && wrap_mode != GOF_WRAP_MINIMAGE)
        return fail(GOF_E_INVALID, "gof_engine_step: bad wrap_mode");
    if (timestep == 0.0f || timestep != timestep) return fail(GOF_E_INVALID, "gof_engine_step: bad timestep");
    if (e->slab) return fail(GOF_E_STATE, "gof_engine_step: a slab engine steps through the gof_slab_* phases");
    if (timestep < 0) timestep = -1.0f;  // every negative value means "adaptive": one graph key
    cudaStream_t st = (cudaStream_t)stream;
    for (int s = 0; s < n_steps; ++s) {
        const bool timed = e->timing && e->ev_used < kMaxTimedLaunches;
        if (e->use_graph && !e->timing && !e->slab) {
            gof_engine::StepGraph& sg = e->graphs[e->parity];
            if (!sg.exec || sg.timestep != timestep || sg.wrap_mode != wrap_mode || sg.split != g_force_split ||
                sg.generation != e->generation)
                if (int rc = engine_capture_step(e, timestep, wrap_mode)) return rc;
            GOF_CUDA(cudaGraphLaunch(sg.exec, st));
            g_launches.fetch_add((uint64_t)sg.launches, std::memory_order_relaxed);
            e->binned = true;
        } else {
            if (int rc = engine_enqueue_step(e, timestep, wrap_mode, st, timed)) return rc;
        }
        e->parity ^= 1;
    }
    return 0;
}

extern "C" int gof_engine_accelerate(gof_engine* e, int wrap_mode, void* stream)
{
    DeviceGuard guard(e ? e->device : -1);
    if (int rc = engine_ready(e, true)) return rc;
    if (wrap_mode != GOF_WRAP_REFERENCE && wrap_mode != GOF_WRAP_MINIMAGE)
        return fail(GOF_E_INVALID, "gof_engine_accelerate: bad wrap_mode");
    if (e->slab) return fail(GOF_E_STATE, "gof_engine_accelerate: not available on a slab engine");
    cudaStream_t st = (cudaStream_t)stream;
    if (int rc = engine_bin(e, false, st)) return rc;
    ForceArgs a;
    engine_force_args(e, wrap_mode, &a);
    a.integrate = 0;
    a.parity_out = e->parity;  // velocities unchanged: the maximum is re-accumulated into the same slot
    return engine_force(e, a, st);
}

extern "C" int gof_engine_download(gof_engine* e, float* px, float* py, float* pz, float* vx, float* vy,
                                   float* vz, float* ax, float* ay, float* az, void* stream)
{
    DeviceGuard guard(e ? e->device : -1);
    if (int rc = engine_ready(e, true)) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const int n = e->n;
    const bool wp = px || py || pz, wv = vx || vy || vz, wa = ax || ay || az;
    GOF_LAUNCH(k_unpack, cdiv(n, kEwThreads), kEwThreads, 0, st, e->posw[0], e->veli[0], e->acc, n,
               wp ? e->stage[0] : nullptr, e->stage[1], e->stage[2], wv ? e->stage[3] : nullptr, e->stage[4],
               e->stage[5], wa ? e->stage[6] : nullptr, e->stage[7], e->stage[8]);
    float* dst[9] = {px, py, pz, vx, vy, vz, ax, ay, az};
    for (int k = 0; k < 9; ++k)
        if (dst[k]) GOF_CUDA(cudaMemcpyAsync(dst[k], e->stage[k], (size_t)n * sizeof(float), cudaMemcpyDefault, st));
    return 0;
}

/* One step of a host-driven loop with the transfers in both directions at once (see the header). */
extern "C" int gof_engine_step_host(gof_engine* e, const float* const* in9, const int32_t* types,
                                    float* const* out9, float timestep, int wrap_mode, void* stream)
{
    DeviceGuard guard(e ? e->device : -1);
    if (int rc = engine_ready(e, false)) return rc;
    if (e->slab) return fail(GOF_E_STATE, "gof_engine_step_host: not available on a slab engine");
    if (!in9 || !out9 || !types) return fail(GOF_E_INVALID, "gof_engine_step_host: null argument");
    for (int k = 0; k < 9; ++k)
        if (!in9[k] || !out9[k]) return fail(GOF_E_INVALID, "gof_engine_step_host: all nine arrays are required both ways");
    cudaStream_t st = (cudaStream_t)stream;
    if (!e->host_side) {
        GOF_CUDA(cudaStreamCreateWithFlags(&e->host_side, cudaStreamNonBlocking));
        GOF_CUDA(cudaEventCreateWithFlags(&e->ev_ready, cudaEventDisableTiming));
        for (int k = 0; k < 9; ++k) GOF_CUDA(cudaEventCreateWithFlags(&e->ev_down[k], cudaEventDisableTiming));
    }
    const int n = e->capacity;
    // uploads: an array that the previous call is still downloading into waits for exactly that copy
    for (int k = 0; k < 9; ++k) {
        if (e->down_pending)
            for (int m = 0; m < 9; ++m)
                if (e->down_ptr[m] == (const void*)in9[k]) GOF_CUDA(cudaStreamWaitEvent(st, e->ev_down[m], 0));
        GOF_CUDA(cudaMemcpyAsync(e->stage[k], in9[k], (size_t)n * sizeof(float), cudaMemcpyDefault, st));
    }
    GOF_CUDA(cudaMemcpyAsync(e->stage_types, types, (size_t)n * sizeof(int), cudaMemcpyDefault, st));
    if (e->species_dirty) {
        if (int rc = upload_species(e->species, e->dspecies, st)) return rc;
        e->species_dirty = false;
    }
    GOF_CUDA(cudaMemsetAsync(e->scalars, 0, sizeof(StepScalars), st));
    e->parity = 0;
    e->n = n;
    GOF_LAUNCH(k_pack, cdiv(n, kEwThreads), kEwThreads, 0, st, e->stage[0], e->stage[1], e->stage[2], e->stage[3],
               e->stage[4], e->stage[5], e->stage[6], e->stage[7], e->stage[8], e->stage_types, n, e->posw[0],
               e->veli[0], e->acc, &e->scalars->max_sq_speed_bits[0]);
    e->uploaded = true;
    e->binned = false;
    if (int rc = gof_engine_step(e, 1, timestep, wrap_mode, stream)) return rc;
    // the way back: un-permute into the second staging set once the previous call's copies have left it
    if (e->down_pending) GOF_CUDA(cudaStreamWaitEvent(st, e->ev_down[8], 0));
    GOF_LAUNCH(k_unpack, cdiv(n, kEwThreads), kEwThreads, 0, st, e->posw[0], e->veli[0], e->acc, n, e->stage_down[0],
               e->stage_down[1], e->stage_down[2], e->stage_down[3], e->stage_down[4], e->stage_down[5],
               e->stage_down[6], e->stage_down[7], e->stage_down[8]);
    GOF_CUDA(cudaEventRecord(e->ev_ready, st));
    GOF_CUDA(cudaStreamWaitEvent(e->host_side, e->ev_ready, 0));
    for (int k = 0; k < 9; ++k) {
        GOF_CUDA(cudaMemcpyAsync(out9[k], e->stage_down[k], (size_t)n * sizeof(float), cudaMemcpyDefault, e->host_side));
        GOF_CUDA(cudaEventRecord(e->ev_down[k], e->host_side));
        e->down_ptr[k] = out9[k];
    }
    e->down_pending = true;
    return 0;
}

extern "C" int gof_engine_host_wait(gof_engine* e)
{
    if (!e) return fail(GOF_E_INVALID, "null engine");
    DeviceGuard guard(e->device);
    if (e->down_pending) GOF_CUDA(cudaEventSynchronize(e->ev_down[8]));
    return 0;
}

extern "C" int gof_engine_snapshot(gof_engine* e, float* xyz, void* stream)
{
    DeviceGuard guard(e ? e->device : -1);
    if (int rc = engine_ready(e, true)) return rc;
    if (!xyz) return fail(GOF_E_INVALID, "gof_engine_snapshot: null output");
    cudaStream_t st = (cudaStream_t)stream;
    const int n = e->n;
    GOF_LAUNCH(k_snapshot, cdiv(n, kEwThreads), kEwThreads, 0, st, e->posw[0], e->veli[0], n, e->stage_xyz);
    GOF_CUDA(cudaMemcpyAsync(xyz, e->stage_xyz, (size_t)n * 3 * sizeof(float), cudaMemcpyDefault, st));
    return 0;
}

/* The two halves of gof_engine_snapshot, for double buffering: `stage` runs the un-permute
 * kernel into the engine's device staging buffer on `stream` (it must precede the next step,
 * which rewrites the state); `copy` moves the staged frame to xyz_any on `stream`, typically a
 * side stream that waited for the stage, so that the next step overlaps the transfer.  The
 * next `stage` must wait for the previous `copy`. */
extern "C" int gof_engine_snapshot_stage(gof_engine* e, void* stream)
{
    DeviceGuard guard(e ? e->device : -1);
    if (int rc = engine_ready(e, true)) return rc;
    const int n = e->n;
    GOF_LAUNCH(k_snapshot, cdiv(n, kEwThreads), kEwThreads, 0, (cudaStream_t)stream, e->posw[0], e->veli[0], n,
               e->stage_xyz);
    return 0;
}

extern "C" int gof_engine_snapshot_copy(gof_engine* e, float* xyz, void* stream)
{
    DeviceGuard guard(e ? e->device : -1);
    if (int rc = engine_ready(e, true)) return rc;
    if (!xyz) return fail(GOF_E_INVALID, "gof_engine_snapshot_copy: null output");
    GOF_CUDA(cudaMemcpyAsync(xyz, e->stage_xyz, (size_t)e->n * 3 * sizeof(float), cudaMemcpyDefault,
                             (cudaStream_t)stream));
    return 0;
}

extern "C" int gof_engine_read_bins(gof_engine* e, int32_t* particle_bins, int32_t* bin_counts,
                                    int32_t* bin_starts, int32_t* particle_indices, void* stream)
{
    DeviceGuard guard(e ? e->device : -1);
    if (int rc = engine_ready(e, true)) return rc;
    if (!e->binned) return fail(GOF_E_STATE, "gof_engine_read_bins: no step or accelerate has run since the upload");
    cudaStream_t st = (cudaStream_t)stream;
    const int n = e->n;
    // posw[1]/veli[1] still hold the cell-ordered snapshot the last force kernel read
    GOF_LAUNCH(k_export_bins, cdiv(e->slot_capacity, kEwThreads), kEwThreads, 0, st, e->veli[1], e->cell_sorted,
               e->start, e->start, e->count, e->geom.num_cells, e->stage_bins,
               e->stage_indices);
    if (particle_bins) GOF_CUDA(cudaMemcpyAsync(particle_bins, e->stage_bins, (size_t)n * sizeof(int), cudaMemcpyDefault, st));
    if (particle_indices) GOF_CUDA(cudaMemcpyAsync(particle_indices, e->stage_indices, (size_t)n * sizeof(int), cudaMemcpyDefault, st));
    const size_t cb = (size_t)e->geom.num_cells * sizeof(int);
    if (bin_counts) GOF_CUDA(cudaMemcpyAsync(bin_counts, e->count, cb, cudaMemcpyDefault, st));
    if (bin_starts) GOF_CUDA(cudaMemcpyAsync(bin_starts, e->start, cb, cudaMemcpyDefault, st));
    return 0;
}

/* Diagnostics for bench.py: stencil candidates, force evaluations and pairs in range, summed
 * over the particles of the snapshot the last step / accelerate read (see k_pair_stats).
 * Synchronises `stream`. */
extern "C" int gof_engine_pair_stats(gof_engine* e, int wrap_mode, unsigned long long* out3, void* stream)
{
    DeviceGuard guard(e ? e->device : -1);
    if (int rc = engine_ready(e, true)) return rc;
    if (!out3) return fail(GOF_E_INVALID, "gof_engine_pair_stats: null output");
    if (!e->binned) return fail(GOF_E_STATE, "gof_engine_pair_stats: no step or accelerate has run since the upload");
    if (e->n <= 0) { out3[0] = out3[1] = out3[2] = 0; return 0; }
    cudaStream_t st = (cudaStream_t)stream;
    // the slot/rank scratch is free between steps: borrow three 64-bit words of it
    PairStats* d = reinterpret_cast<PairStats*>(e->slot_id);
    GOF_CUDA(cudaMemsetAsync(d, 0, sizeof(PairStats), st));
    GOF_LAUNCH(k_pair_stats, cdiv(e->n, kForceThreads), kForceThreads, 0, st, e->geom, (const float4*)e->posw[1],
               (const int*)e->cell_sorted, (const int*)e->start, (const int*)e->count, e->n, wrap_mode,
               e->species.kind() != kKindBoids ? 1 : 0, d);  // (the Clusters and the mixed kind cull per warp)
    PairStats hst{};
    GOF_CUDA(cudaMemcpyAsync(&hst, d, sizeof(PairStats), cudaMemcpyDeviceToHost, st));
    GOF_CUDA(cudaStreamSynchronize(st));
    out3[0] = hst.candidates; out3[1] = hst.evaluated; out3[2] = hst.in_range;
    return 0;
}

extern "C" int gof_engine_last_timestep(gof_engine* e, float* out, void* stream)
{
    DeviceGuard guard(e ? e->device : -1);
    if (int rc = engine_ready(e, true)) return rc;
    if (!out) return fail(GOF_E_INVALID, "gof_engine_last_timestep: null output");
    cudaStream_t st = (cudaStream_t)stream;
    GOF_CUDA(cudaMemcpyAsync(out, &e->scalars->dt, sizeof(float), cudaMemcpyDeviceToHost, st));
    GOF_CUDA(cudaStreamSynchronize(st));
    return 0;
}

// --------------------------------------------------------------------------------------
// slab decomposition (multi-GPU): one engine per z-slab, host exchanges via NCCL send/recv
// --------------------------------------------------------------------------------------
extern "C" int gof_slab_create(gof_engine** out, int device, int capacity, int num_types,
                               const gof_grid* grid, const double* pm_host, int layer_lo, int layer_hi,
                               int migrate_capacity, int halo_capacity)
{
    if (!out || capacity < 1 || !grid) return fail(GOF_E_INVALID, "gof_slab_create: bad arguments");
    if (grid->num_bin_z < 4) return fail(GOF_E_GRID, "gof_slab_create: slab decomposition needs a 3D box with >= 4 layers");
    const int owned = layer_hi - layer_lo;
    if (layer_lo < 0 || layer_hi > grid->num_bin_z || owned < 1 || owned > grid->num_bin_z - 2)
        return fail(GOF_E_INVALID, "gof_slab_create: a slab owns between 1 and num_bin_z - 2 layers");
    if (migrate_capacity < 1 || halo_capacity < 1) return fail(GOF_E_INVALID, "gof_slab_create: bad capacities");
    gof_engine* e = new (std::nothrow) gof_engine();
    if (!e) return fail(GOF_E_NOMEM, "gof_slab_create: host allocation failed");
    e->device = device;
    e->capacity = capacity;
    e->n = 0;
    e->T = num_types;
    e->grid = *grid;
    e->slab = true;
    e->layer_lo = layer_lo;
    e->layer_hi = layer_hi;
    e->mig_capacity = migrate_capacity;
    e->halo_capacity = halo_capacity;
    int rc = make_geometry(grid, layer_lo - 1, owned + 2, &e->geom);
    if (!rc) rc = build_species(pm_host, num_types, &e->species);
    if (rc) {
        delete e;
        return rc;
    }
    e->arena_bytes = engine_layout(e, nullptr);
    *out = e;
    return 0;
}

namespace gof {
// slab upload: SoA + explicit global ids -> float4 state, and the maximum squared speed
__global__ void __launch_bounds__(kEwThreads)
k_pack_ids(const float* __restrict__ px, const float* __restrict__ py, const float* __restrict__ pz,
           const float* __restrict__ vx, const float* __restrict__ vy, const float* __restrict__ vz,
           const float* __restrict__ ax, const float* __restrict__ ay, const float* __restrict__ az,
           const int* __restrict__ types, const int* __restrict__ ids, int n, float4* __restrict__ posw,
           float4* __restrict__ veli, float4* __restrict__ acc, unsigned* __restrict__ max_bits)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    float sq = 0.0f;
    if (i < n) {
        const float4 v = make_float4(vx ? vx[i] : 0.f, vy ? vy[i] : 0.f, vz ? vz[i] : 0.f, __int_as_float(ids[i]));
        posw[i] = make_float4(px[i], py[i], pz[i], __int_as_float(types[i]));
        veli[i] = v;
        acc[i] = make_float4(ax ? ax[i] : 0.f, ay ? ay[i] : 0.f, az ? az[i] : 0.f, 0.f);
        sq = v.x * v.x + v.y * v.y + v.z * v.z;
    }
    const unsigned w = __reduce_max_sync(0xffffffffu, __float_as_uint(sq));
    if (lane_id() == 0) atomicMax(max_bits, w);
}

// rest order -> SoA staging without un-permuting (slab download: the host scatters by id)
__global__ void __launch_bounds__(kEwThreads)
k_unpack_rest(const float4* __restrict__ posw, const float4* __restrict__ veli, const float4* __restrict__ acc,
              int n, float* __restrict__ px, float* __restrict__ py, float* __restrict__ pz,
              float* __restrict__ vx, float* __restrict__ vy, float* __restrict__ vz,
              float* __restrict__ ax, float* __restrict__ ay, float* __restrict__ az, int* __restrict__ ids,
              int* __restrict__ types)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const float4 p = posw[k], v = veli[k], a = acc[k];
    px[k] = p.x; py[k] = p.y; pz[k] = p.z;
    vx[k] = v.x; vy[k] = v.y; vz[k] = v.z;
    ax[k] = a.x; ay[k] = a.y; az[k] = a.z;
    ids[k] = __float_as_int(v.w);
    types[k] = __float_as_int(p.w);
}
}  // namespace gof

static int slab_ready(gof_engine* e, bool need_state)
{
    if (int rc = engine_ready(e, need_state)) return rc;
    if (!e->slab) return fail(GOF_E_STATE, "not a slab engine (use gof_slab_create)");
    return 0;
}

extern "C" int gof_slab_upload(gof_engine* e, int n, const float* px, const float* py, const float* pz,
                               const float* vx, const float* vy, const float* vz, const float* ax,
                               const float* ay, const float* az, const int32_t* types, const int32_t* ids,
                               void* stream)
{
    DeviceGuard guard(e ? e->device : -1);
    if (int rc = slab_ready(e, false)) return rc;
    if (n < 0 || n > e->capacity) return fail(GOF_E_NOMEM, "gof_slab_upload: more particles than capacity");
    if (n > 0 && (!px || !py || !pz || !types || !ids)) return fail(GOF_E_INVALID, "gof_slab_upload: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    const float* src[9] = {px, py, pz, vx, vy, vz, ax, ay, az};
    for (int k = 0; k < 9; ++k)
        if (src[k] && n) GOF_CUDA(cudaMemcpyAsync(e->stage[k], src[k], (size_t)n * sizeof(float), cudaMemcpyDefault, st));
    if (n) {
        GOF_CUDA(cudaMemcpyAsync(e->stage_types, types, (size_t)n * sizeof(int), cudaMemcpyDefault, st));
        GOF_CUDA(cudaMemcpyAsync(e->stage_bins, ids, (size_t)n * sizeof(int), cudaMemcpyDefault, st));
    }
    if (e->species_dirty) {
        if (int rc = upload_species(e->species, e->dspecies, st)) return rc;
        e->species_dirty = false;
    }
    GOF_CUDA(cudaMemsetAsync(e->scalars, 0, sizeof(StepScalars), st));
    e->parity = 0;
    e->n = n;
    if (n)
        GOF_LAUNCH(k_pack_ids, cdiv(n, kEwThreads), kEwThreads, 0, st, e->stage[0], e->stage[1], e->stage[2],
                   vx ? e->stage[3] : nullptr, vy ? e->stage[4] : nullptr, vz ? e->stage[5] : nullptr,
                   ax ? e->stage[6] : nullptr, ay ? e->stage[7] : nullptr, az ? e->stage[8] : nullptr,
                   e->stage_types, e->stage_bins, n, e->posw[0], e->veli[0], e->acc,
                   &e->scalars->max_sq_speed_bits[0]);
    // the live count and the sticky error flag as the device-driven step reads them
    GOF_CUDA(cudaMemsetAsync(e->counters, 0, 8 * sizeof(int), st));
    GOF_CUDA(cudaMemsetAsync(e->mig_send[0], 0, sizeof(float4), st));
    GOF_CUDA(cudaMemsetAsync(e->mig_send[1], 0, sizeof(float4), st));
    GOF_LAUNCH(k_set_word, 1, 1, 0, st, e->counters + 4, n);
    e->p2p.need_publish = true;
    e->uploaded = true;
    e->binned = false;
    return 0;
}

/* Device pointers the host wraps as tensors for send/recv (all inside the bound arena). */
extern "C" int gof_slab_pointers(gof_engine* e, void** ptrs, int n_ptrs)
{
    DeviceGuard guard(e ? e->device : -1);
    if (int rc = slab_ready(e, false)) return rc;
    if (!ptrs || n_ptrs < 11) return fail(GOF_E_INVALID, "gof_slab_pointers: need room for 11 pointers");
    ptrs[0] = e->mig_send[0];
    ptrs[1] = e->mig_send[1];
    ptrs[2] = e->mig_recv[0];
    ptrs[3] = e->mig_recv[1];
    ptrs[4] = e->posw[1];
    ptrs[5] = e->veli[1];
    ptrs[6] = e->count;
    ptrs[7] = &e->scalars->max_sq_speed_bits[0];
    ptrs[8] = &e->scalars->max_sq_speed_bits[1];
    ptrs[9] = e->start;
    ptrs[10] = e->counters;
    return 0;
}

namespace gof {
// {live particles, population of the bottom owned layer, of the top owned layer, error flag}
__global__ void k_slab_counts(const int* __restrict__ start, const int* __restrict__ error,
                              int* __restrict__ out, int num_cells, int per_layer, int top)
{
    out[0] = start[num_cells];
    out[1] = start[2 * per_layer] - start[per_layer];
    out[2] = start[(top + 1) * per_layer] - start[top * per_layer];
    out[3] = *error;
}
}  // namespace gof

/* Phase 1: bin the resident particles; the ones that left the slab are packed into the two
 * send buffers (header word = count).  Nothing comes back to the host: the buffers travel
 * whole (fixed size) and phase 2 reads the counts on the device. */
extern "C" int gof_slab_phase_hash_async(gof_engine* e, void* stream)
{
    DeviceGuard guard(e ? e->device : -1);
    if (int rc = slab_ready(e, true)) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const Geometry& g = e->geom;
    GOF_CUDA(cudaMemsetAsync(e->count, 0, (size_t)(g.num_cells + 2) * sizeof(int), st));
    GOF_CUDA(cudaMemsetAsync(e->counters, 0, 8 * sizeof(int), st));
    GOF_CUDA(cudaMemsetAsync(e->mig_send[0], 0, sizeof(float4), st));
    GOF_CUDA(cudaMemsetAsync(e->mig_send[1], 0, sizeof(float4), st));
    if (e->n > 0) {
        MigrateOut mig{e->mig_send[0], e->mig_send[1], e->counters + 2, e->mig_capacity};
        GOF_LAUNCH(k_hash_count<false>, cdiv(e->n, kBinThreads), kBinThreads, 0, st, e->posw[0], e->veli[0], e->acc,
                   (const float*)nullptr, (const float*)nullptr, (const float*)nullptr, e->n, (const int*)nullptr, g,
                   e->cell_of, e->off, e->count, mig);
    }
    e->n_slots = e->n;
    return 0;
}

/* Phase 2: append the arrivals (the two received buffers), bin them, sort everything by cell
 * and kick the velocities.  The maximum squared speed must have been all-reduced over the
 * slabs before this call when the timestep is adaptive.  Afterwards counters[4..7] hold {live
 * particles, population of the bottom owned layer, of the top one, error flag}; the host
 * all-gathers them and reports them back with gof_slab_commit. */
extern "C" int gof_slab_phase_sort_async(gof_engine* e, float timestep, void* stream)
{
    DeviceGuard guard(e ? e->device : -1);
    if (int rc = slab_ready(e, true)) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const Geometry& g = e->geom;
    const int max_arrivals = 2 * e->mig_capacity;
    int bound = e->n_slots + max_arrivals;  // slots that may be in use after the append
    if (bound > e->capacity) bound = e->capacity;
    // (no room left for arrivals: a device flag, so that it travels with the words every rank
    // gathers and all of them raise, instead of one rank failing here while its peers wait)
    if (e->n_slots + 64 > e->capacity) GOF_LAUNCH(k_set_word, 1, 1, 0, st, e->counters + 2, 5);
    const int room = bound - e->n_slots > 0 ? bound - e->n_slots : 0;
    int* words = e->counters;  // [0] arrivals, [1] slots in use ([2] is the error flag)
    GOF_LAUNCH(k_append_arrivals, cdiv(max_arrivals, kBinThreads), kBinThreads, 0, st, e->mig_recv[0], e->mig_recv[1],
               e->mig_capacity < room / 2 ? e->mig_capacity : room / 2, e->n_slots, e->posw[0], e->veli[0], e->acc,
               words);
    MigrateOut flag_only{nullptr, nullptr, e->counters + 2, 0};
    if (room > 0)
        GOF_LAUNCH(k_hash_count<false>, cdiv(room, kBinThreads), kBinThreads, 0, st, e->posw[0] + e->n_slots,
                   e->veli[0] + e->n_slots, e->acc + e->n_slots, (const float*)nullptr, (const float*)nullptr,
                   (const float*)nullptr, room, (const int*)words, g, e->cell_of + e->n_slots, e->off + e->n_slots,
                   e->count, flag_only);
    GOF_LAUNCH(k_prepare_dt, 1, 1, 0, st, e->scalars, e->parity, timestep < 0 ? -1.0f : timestep);
    if (int rc = launch_scan(e->count, g.num_cells + 1, e->tile_sums, e->start, st)) return rc;
    GOF_LAUNCH(k_scatter_ids<false>, cdiv(bound, kBinThreads), kBinThreads, 0, st, e->veli[0], bound,
               (const int*)(words + 1), e->cell_of, e->off, e->start, e->slot_id, e->posw[0], engine_keys(e));
    GOF_LAUNCH(k_rank_gather_kick, cdiv(bound, kBinThreads), kBinThreads, 0, st, bound, (const int*)(words + 1),
               e->cell_of, e->start, e->count, e->slot_id, engine_keys(e), g.num_cells, e->posw[0], e->veli[0], e->acc,
               e->posw[1],
               e->veli[1], e->cell_sorted, e->scalars, 1);
    GOF_LAUNCH(k_slab_counts, 1, 1, 0, st, e->start, e->counters + 2, e->counters + 4, g.num_cells, g.nbx * g.nby,
               g.nlz - 2);
    e->n_slots = bound;  // upper bound until gof_slab_commit
    e->binned = true;
    return 0;
}

/* The host's copy of counters[4..7] after phase 2 (it needs them for the halo messages anyway). */
extern "C" int gof_slab_commit(gof_engine* e, int n_live, int lo_pop, int hi_pop, int error)
{
    DeviceGuard guard(e ? e->device : -1);
    if (int rc = slab_ready(e, true)) return rc;
    if (error == 1) return fail(GOF_E_STATE, "slab: a particle moved more than one layer in a step, or arrived in the wrong slab");
    if (error == 2) return fail(GOF_E_NOMEM, "slab: migration buffer overflow (raise migrate_capacity)");
    if (error == 5) return fail(GOF_E_NOMEM, "slab: particle capacity exceeded");
    if (n_live < 0 || n_live > e->capacity) return fail(GOF_E_NOMEM, "slab: particle capacity exceeded");
    if (lo_pop > e->halo_capacity || hi_pop > e->halo_capacity)
        return fail(GOF_E_NOMEM, "slab: boundary layer larger than halo_capacity");
    e->n = n_live;
    e->lo_pop = lo_pop;
    e->hi_pop = hi_pop;
    return 0;
}

static int slab_timed_force(gof_engine* e, const ForceArgs& a, cudaStream_t st)
{
    const bool timed = e->timing && e->ev_used < kMaxTimedLaunches;
    if (timed) {
        while ((int)e->ev.size() < 2 * (e->ev_used + 1)) {
            cudaEvent_t v;
            GOF_CUDA(cudaEventCreate(&v));
            e->ev.push_back(v);
        }
        GOF_CUDA(cudaEventRecord(e->ev[2 * e->ev_used], st));
    }
    if (int rc = launch_force(e->geom, a, e->species.kind(), kModeIntegrate, st)) return rc;
    if (timed) {
        GOF_CUDA(cudaEventRecord(e->ev[2 * e->ev_used + 1], st));
        ++e->ev_used;
    }
    return 0;
}

/* Phase 3a: forces + integration of the owned particles that are in neither boundary layer.
 * They need no halo, so this can be enqueued right after gof_slab_phase_sort_async while the
 * host is still gathering sizes and moving the halo layers on another stream: the range is
 * read from the device words of phase 2. */
extern "C" int gof_slab_phase_force_interior(gof_engine* e, int wrap_mode, void* stream)
{
    DeviceGuard guard(e ? e->device : -1);
    if (int rc = slab_ready(e, true)) return rc;
    if (wrap_mode != GOF_WRAP_REFERENCE && wrap_mode != GOF_WRAP_MINIMAGE)
        return fail(GOF_E_INVALID, "gof_slab_phase_force_interior: bad wrap_mode");
    if (e->layer_hi - e->layer_lo < 3) return 0;  // no interior layer
    ForceArgs a;
    engine_force_args(e, wrap_mode, &a);
    a.pstart = e->start;
    a.integrate = 1;
    a.parity_out = e->parity ^ 1;
    a.interior_words = e->counters + 4;
    a.home_begin = 0;
    a.home_count = e->n_slots;  // upper bound of the owned particles (set by phase 2)
    return slab_timed_force(e, a, (cudaStream_t)stream);
}

/* Phase 3b: the halo particles (n_lo from the slab below, then n_hi from the slab above) have
 * been received behind the owned ones and their per-cell counts into the count rows of the two
 * halo layers: halo cell table, then the two boundary layers.  `skip_interior` = 0 also runs
 * the interior here (the one-call form gof_slab_phase_force). */
static int slab_force_boundary(gof_engine* e, int n_halo_lo, int n_halo_hi, int wrap_mode, bool with_interior,
                               cudaStream_t st)
{
    if (n_halo_lo < 0 || n_halo_hi < 0 || n_halo_lo > e->halo_capacity || n_halo_hi > e->halo_capacity)
        return fail(GOF_E_NOMEM, "slab: halo larger than halo_capacity");
    if (wrap_mode != GOF_WRAP_REFERENCE && wrap_mode != GOF_WRAP_MINIMAGE)
        return fail(GOF_E_INVALID, "gof_slab_phase_force: bad wrap_mode");
    const Geometry& g = e->geom;
    const int per_layer = g.nbx * g.nby;
    GOF_LAUNCH(k_halo_starts, 2, kScanThreads, 0, st, e->count, e->start, per_layer, (g.nlz - 1) * per_layer, e->n,
               e->n + n_halo_lo);
    ForceArgs a;
    engine_force_args(e, wrap_mode, &a);
    a.pstart = e->start;
    a.integrate = 1;
    a.parity_out = e->parity ^ 1;
    const int owned = e->layer_hi - e->layer_lo;
    if (with_interior || owned < 3) {
        // everything in one launch (also the case when there is no interior layer)
        if (int rc = slab_timed_force(e, a, st)) return rc;
    } else {
        // the bottom and the top owned layer in one launch
        a.home_begin = 0;
        a.home_count = e->lo_pop;
        a.home_begin2 = e->n - e->hi_pop;
        a.home_count2 = e->hi_pop;
        if (int rc = slab_timed_force(e, a, st)) return rc;
    }
    e->parity ^= 1;
    return 0;
}

extern "C" int gof_slab_phase_force_boundary(gof_engine* e, int n_halo_lo, int n_halo_hi, int wrap_mode,
                                             void* stream)
{
    DeviceGuard guard(e ? e->device : -1);
    if (int rc = slab_ready(e, true)) return rc;
    return slab_force_boundary(e, n_halo_lo, n_halo_hi, wrap_mode, false, (cudaStream_t)stream);
}

/* Phase 3 in one call (no overlap): halo table + all owned particles. */
extern "C" int gof_slab_phase_force(gof_engine* e, int n_halo_lo, int n_halo_hi, int wrap_mode, void* stream)
{
    DeviceGuard guard(e ? e->device : -1);
    if (int rc = slab_ready(e, true)) return rc;
    return slab_force_boundary(e, n_halo_lo, n_halo_hi, wrap_mode, true, (cudaStream_t)stream);
}

// --------------------------------------------------------------------------------------
// slab decomposition over peer memory: the neighbours' arenas are mapped into this process
// (CUDA IPC between processes, plain pointers inside one), every exchange is a kernel that
// stores into the neighbour and raises a tagged flag there, every wait is a one-block kernel
// on the consumer's stream.  No size returns to the host during a run.
// --------------------------------------------------------------------------------------
static const char* p2p_error_text(int code)
{
    switch (code) {
        case 1: return "slab: a particle moved more than one layer in a step, or arrived in the wrong slab";
        case 2: return "slab: migration buffer overflow (raise migrate_capacity)";
        case kErrPeerTimeout: return "slab: a neighbour's data did not arrive within 4 s (peer stalled or failed)";
        case kErrHaloOverflow: return "slab: a boundary layer is larger than the neighbour's halo_capacity";
        case 5: return "slab: particle capacity exceeded by arrivals";
        default: return "slab: device-side error";
    }
}

/* Synchronise `st`, fetch {live, bottom, top, -} and the sticky error flag, raise if it is set. */
static int slab_p2p_sync_count(gof_engine* e, cudaStream_t st)
{
    int w[8];
    GOF_CUDA(cudaMemcpyAsync(w, e->counters, sizeof(w), cudaMemcpyDeviceToHost, st));
    GOF_CUDA(cudaStreamSynchronize(st));
    if (w[2] != 0) return fail(w[2] == 2 || w[2] == 5 || w[2] == kErrHaloOverflow ? GOF_E_NOMEM : GOF_E_STATE, p2p_error_text(w[2]));
    if (w[4] < 0 || w[4] > e->capacity) return fail(GOF_E_NOMEM, "slab: particle capacity exceeded");
    e->n = w[4];
    e->lo_pop = w[5];
    e->hi_pop = w[6];
    return 0;
}

extern "C" int gof_slab_check(gof_engine* e, int* n_live, void* stream)
{
    DeviceGuard guard(e ? e->device : -1);
    if (int rc = slab_ready(e, true)) return rc;
    if (int rc = slab_p2p_sync_count(e, (cudaStream_t)stream)) return rc;
    if (n_live) *n_live = e->n;
    return 0;
}

extern "C" int gof_slab_window(gof_engine* e, int64_t* info, int n_info)
{
    DeviceGuard guard(e ? e->device : -1);
    if (int rc = slab_ready(e, false)) return rc;
    if (!info || n_info < 12) return fail(GOF_E_INVALID, "gof_slab_window: need room for 12 words");
    auto off = [&](const void* p) { return (int64_t)((const char*)p - e->arena_base); };
    info[0] = off(e->mig_recv[0]);
    info[1] = off(e->mig_recv[1]);
    info[2] = off(e->posw[1]);
    info[3] = off(e->veli[1]);
    info[4] = off(e->count);
    info[5] = off(e->p2p_flags);
    info[6] = e->capacity;
    info[7] = e->halo_capacity;
    info[8] = e->geom.nlz;
    info[9] = e->mig_capacity;
    info[10] = (int64_t)e->geom.nbx * e->geom.nby;
    info[11] = off(e->halo_count_in);
    return 0;
}

extern "C" int gof_slab_connect(gof_engine* e, int rank, int world, void* const* bases, const int64_t* infos)
{
    DeviceGuard guard(e ? e->device : -1);
    if (int rc = slab_ready(e, false)) return rc;
    if (!bases || !infos || world < 2 || world > kMaxWorld || rank < 0 || rank >= world)
        return fail(GOF_E_INVALID, "gof_slab_connect: bad arguments (2 <= world <= 16)");
    const int64_t* mine = infos + 12 * rank;
    if ((char*)bases[rank] != e->arena_base || mine[6] != e->capacity || mine[8] != e->geom.nlz)
        return fail(GOF_E_INVALID, "gof_slab_connect: entry `rank` does not describe this slab");
    if (e->layer_hi - e->layer_lo < 2)
        return fail(GOF_E_INVALID, "gof_slab_connect: a slab needs at least two owned layers");
    auto side = [&](int peer, bool peer_is_below, PeerSide* out) -> int {
        const int64_t* pi = infos + 12 * peer;
        char* b = (char*)bases[peer];
        if (!b) return fail(GOF_E_INVALID, "gof_slab_connect: null peer base");
        if (pi[9] != e->mig_capacity || pi[10] != (int64_t)e->geom.nbx * e->geom.nby)
            return fail(GOF_E_INVALID, "gof_slab_connect: migrate_capacity and the layer shape must match on all slabs");
        const int64_t cap = pi[6], hcap = pi[7], per_layer = pi[10];
        unsigned long long* flags = (unsigned long long*)(b + pi[5]);
        if (peer_is_below) {
            // this slab is ABOVE the peer: records arrive there "from above", the bottom layer
            // becomes the peer's UPPER halo (second halo region, count row of its last layer)
            out->mig_recv = (float4*)(b + pi[1]);
            out->halo_posw = (float4*)(b + pi[2]) + cap + hcap;
            out->halo_veli = (float4*)(b + pi[3]) + cap + hcap;
            out->halo_count = (int*)(b + pi[11]) + per_layer;
            out->mig_flag = flags + 1;
            out->halo_flag = flags + 3;
        } else {
            out->mig_recv = (float4*)(b + pi[0]);
            out->halo_posw = (float4*)(b + pi[2]) + cap;
            out->halo_veli = (float4*)(b + pi[3]) + cap;
            out->halo_count = (int*)(b + pi[11]);
            out->mig_flag = flags + 0;
            out->halo_flag = flags + 2;
        }
        out->halo_capacity = (int)hcap;
        return 0;
    };
    if (int rc = side((rank + world - 1) % world, true, &e->p2p.below)) return rc;
    if (int rc = side((rank + 1) % world, false, &e->p2p.above)) return rc;
    // where this slab publishes its maximum speed: slot `rank` of every slab's array, per parity
    unsigned long long* ptrs[2 * kMaxWorld] = {};
    for (int par = 0; par < 2; ++par)
        for (int t = 0; t < world; ++t)
            ptrs[par * kMaxWorld + t] = (unsigned long long*)((char*)bases[t] + infos[12 * t + 5] + 64) + par * kMaxWorld + rank;
    GOF_CUDA(cudaMemcpy(e->speed_ptrs, ptrs, sizeof(ptrs), cudaMemcpyHostToDevice));
    GOF_CUDA(cudaMemset(e->p2p_flags, 0, 64 + 2 * kMaxWorld * 8));
    GOF_CUDA(cudaMemset(e->push_done, 0, 4 * sizeof(int)));
    GOF_CUDA(cudaDeviceSynchronize());
    e->p2p.on = true;
    e->p2p.rank = rank;
    e->p2p.world = world;
    e->p2p.step = 0;
    e->p2p.need_publish = true;
    return 0;
}

/* n_steps x core_step over the local slabs of a job whose slabs are connected with
 * gof_slab_connect.  Phases are enqueued slab by slab so that, when several slabs share a GPU
 * (one process, tests), every store into a neighbour is enqueued before the wait for it.
 * `side_stream` carries the halo traffic and the boundary layers under the interior forces. */
extern "C" int gof_slab_step_p2p(gof_engine* const* slabs, int n_slabs, int n_steps, float timestep, int wrap_mode,
                                 void* main_stream, void* side_stream)
{
    if (!slabs || n_slabs < 1 || !slabs[0]) return fail(GOF_E_INVALID, "gof_slab_step_p2p: no slabs");
    DeviceGuard guard(slabs[0]->device);  // (the slabs of one call share a device and its two streams)
    for (int s = 1; s < n_slabs; ++s)
        if (!slabs[s] || slabs[s]->device != slabs[0]->device)
            return fail(GOF_E_INVALID, "gof_slab_step_p2p: the slabs of one call must live on one device");
    if (wrap_mode != GOF_WRAP_REFERENCE && wrap_mode != GOF_WRAP_MINIMAGE)
        return fail(GOF_E_INVALID, "gof_slab_step_p2p: bad wrap_mode");
    if (timestep == 0.0f || timestep != timestep) return fail(GOF_E_INVALID, "gof_slab_step_p2p: bad timestep");
    for (int s = 0; s < n_slabs; ++s) {
        if (int rc = slab_ready(slabs[s], true)) return rc;
        if (!slabs[s]->p2p.on) return fail(GOF_E_STATE, "gof_slab_step_p2p: call gof_slab_connect first");
    }
    cudaStream_t ms = (cudaStream_t)main_stream, ss = (cudaStream_t)side_stream;
    if (!ss || ss == ms) return fail(GOF_E_INVALID, "gof_slab_step_p2p: needs a side stream");
    const bool adaptive = timestep < 0;
    static thread_local cudaEvent_t ev_sorted = nullptr, ev_done = nullptr;
    if (!ev_sorted) {
        GOF_CUDA(cudaEventCreateWithFlags(&ev_sorted, cudaEventDisableTiming));
        GOF_CUDA(cudaEventCreateWithFlags(&ev_done, cudaEventDisableTiming));
    }
    gof_engine* te = slabs[0];  // phase timing (bench.py) follows the first local slab
    auto mark = [&](int which, cudaStream_t st) -> cudaError_t {
        if (!te->timing || te->pev_steps >= kMaxTimedLaunches) return cudaSuccess;
        const size_t idx = (size_t)te->pev_steps * 5 + which;
        while (te->pev.size() <= idx) {
            cudaEvent_t v;
            if (cudaError_t err = cudaEventCreate(&v)) return err;
            te->pev.push_back(v);
        }
        return cudaEventRecord(te->pev[idx], st);
    };
    for (int step = 0; step < n_steps; ++step) {
        GOF_CUDA(mark(0, ms));
        // ---- 1. bin, pack the leavers, push them to the neighbours ------------------------------
        for (int s = 0; s < n_slabs; ++s) {
            gof_engine* e = slabs[s];
            const Geometry& g = e->geom;
            const unsigned tag = ++e->p2p.step;
            if (adaptive && e->p2p.need_publish)
                GOF_LAUNCH(k_publish_speed, 1, 32, 0, ms, (const StepScalars*)e->scalars, e->parity,
                           (unsigned long long* const*)(e->speed_ptrs + e->parity * kMaxWorld), e->p2p.world, tag);
            e->p2p.need_publish = false;
            // (the record counters of the send buffers were cleared by the previous step's arrivals
            // kernel, or by the upload)
            GOF_CUDA(cudaMemsetAsync(e->count, 0, (size_t)(g.num_cells + 2) * sizeof(int), ms));
            MigrateOut mig{e->mig_send[0], e->mig_send[1], e->counters + 2, e->mig_capacity};
            // (the live count of the resident state is a device word: launch over the capacity)
            GOF_LAUNCH(k_hash_count<false>, cdiv(e->capacity, kBinThreads), kBinThreads, 0, ms, e->posw[0], e->veli[0],
                       e->acc, (const float*)nullptr, (const float*)nullptr, (const float*)nullptr, e->capacity,
                       (const int*)(e->counters + 4), g, e->cell_of, e->off, e->count, mig);
            GOF_LAUNCH(k_push_migration, 2, kPushThreads, 0, ms, (const float4*)e->mig_send[0], (const float4*)e->mig_send[1],
                       e->p2p.below, e->p2p.above, e->mig_capacity, tag);
        }
        // ---- 2. arrivals in, timestep, counting sort + kick ------------------------------------
        for (int s = 0; s < n_slabs; ++s) {
            gof_engine* e = slabs[s];
            const Geometry& g = e->geom;
            const unsigned tag = e->p2p.step;
            int* words = e->counters;  // [0] arrivals, [1] slots in use, [2] sticky error, [4..7] sort results
            // wait for the two record buffers, settle the timestep, append + bin the arrivals: one launch
            GOF_LAUNCH(k_arrivals_p2p, cdiv(2 * e->mig_capacity, kBinThreads), kBinThreads, 0, ms,
                       (const unsigned long long*)(e->p2p_flags + 0), (const unsigned long long*)(e->p2p_flags + 1), tag,
                       (const float4*)e->mig_recv[0], (const float4*)e->mig_recv[1], e->mig_capacity, e->posw[0],
                       e->veli[0], e->acc, words, e->capacity, g, e->cell_of, e->off, e->count, e->scalars, e->parity,
                       adaptive ? -1.0f : timestep, (const unsigned long long*)(e->speed_in + e->parity * kMaxWorld),
                       e->p2p.world, e->mig_send[0], e->mig_send[1]);
            if (int rc = launch_scan(e->count, g.num_cells + 1, e->tile_sums, e->start, ms)) return rc;
            GOF_LAUNCH(k_scatter_ids<false>, cdiv(e->capacity, kBinThreads), kBinThreads, 0, ms, e->veli[0], e->capacity,
                       (const int*)(words + 1), e->cell_of, e->off, e->start, e->slot_id, e->posw[0], engine_keys(e),
                       words + 4, (const int*)(words + 2), g.num_cells, g.nbx * g.nby, g.nlz - 2);
            GOF_LAUNCH(k_rank_gather_kick, cdiv(e->capacity, kBinThreads), kBinThreads, 0, ms, e->capacity,
                       (const int*)(words + 1), e->cell_of, e->start, e->count, e->slot_id, engine_keys(e), g.num_cells,
                       e->posw[0], e->veli[0], e->acc, e->posw[1], e->veli[1], e->cell_sorted, e->scalars, 1);
            e->binned = true;
        }
        GOF_CUDA(mark(1, ms));
        GOF_CUDA(cudaEventRecord(ev_sorted, ms));
        GOF_CUDA(cudaStreamWaitEvent(ss, ev_sorted, 0));
        // ---- 3a. interior forces (main stream) ------------------------------------------------
        for (int s = 0; s < n_slabs; ++s) {
            gof_engine* e = slabs[s];
            if (e->layer_hi - e->layer_lo < 3) continue;  // no interior layer
            ForceArgs a;
            engine_force_args(e, wrap_mode, &a);
            a.integrate = 1;
            a.parity_out = e->parity ^ 1;
            a.interior_words = e->counters + 4;
            a.home_begin = 0;
            a.home_count = e->capacity;  // upper bound of the owned particles
            a.n_total = e->capacity;
            if (int rc = slab_timed_force(e, a, ms)) return rc;
        }
        GOF_CUDA(mark(2, ms));
        // ---- 3b. halo layers to the neighbours, then the boundary layers (side stream) ---------
        for (int s = 0; s < n_slabs; ++s) {
            gof_engine* e = slabs[s];
            const Geometry& g = e->geom;
            GOF_LAUNCH(k_push_halo, 2 * kHaloPushBlocks, kPushThreads, 0, ss, (const float4*)e->posw[1],
                       (const float4*)e->veli[1], (const int*)e->count, (const int*)(e->counters + 4), g.nbx * g.nby,
                       g.nlz - 2, e->species.has_boids ? 1 : 0, e->p2p.below, e->p2p.above, e->p2p.step, e->push_done);
        }
        for (int s = 0; s < n_slabs; ++s) {
            gof_engine* e = slabs[s];
            const Geometry& g = e->geom;
            const int per_layer = g.nbx * g.nby;
            GOF_LAUNCH(k_wait_tags, 1, 32, 0, ss, (const unsigned long long*)(e->p2p_flags + 2),
                       (const unsigned long long*)(e->p2p_flags + 3), e->p2p.step, e->halo_n, e->counters + 2);
            GOF_LAUNCH(k_halo_starts_p2p, 2, kScanThreads, 0, ss, (const int*)e->halo_count_in, e->count, e->start,
                       per_layer, (g.nlz - 1) * per_layer, e->capacity, e->capacity + e->halo_capacity);
            // (On the side stream, next to the interior kernel: measured against running them on the
            // main stream behind it, 0.860 against 0.893 ms of forces per step at 1M particles per slab.)
            // the bottom and the top owned layer (with two owned layers: everything), ranges from
            // the device words; the grid covers an upper bound of their population
            ForceArgs a;
            engine_force_args(e, wrap_mode, &a);
            a.integrate = 1;
            a.parity_out = e->parity ^ 1;
            a.n_total = e->capacity;
            a.boundary_words = e->counters + 4;
            const int cap2 = 2 * e->halo_capacity;
            a.home_count = cap2 < e->capacity ? cap2 : e->capacity;
            if (int rc = slab_timed_force(e, a, ss)) return rc;
        }
        GOF_CUDA(mark(3, ss));
        GOF_CUDA(cudaEventRecord(ev_done, ss));
        GOF_CUDA(cudaStreamWaitEvent(ms, ev_done, 0));
        // ---- 4. the new velocities' maximum goes to every slab (next step's timestep) -----------
        for (int s = 0; s < n_slabs; ++s) {
            gof_engine* e = slabs[s];
            e->parity ^= 1;
            if (adaptive)
                GOF_LAUNCH(k_publish_speed, 1, 32, 0, ms, (const StepScalars*)e->scalars, e->parity,
                           (unsigned long long* const*)(e->speed_ptrs + e->parity * kMaxWorld), e->p2p.world,
                           e->p2p.step + 1);
            else
                e->p2p.need_publish = true;  // (a later adaptive step must publish first)
        }
        GOF_CUDA(mark(4, ms));
        if (te->timing && te->pev_steps < kMaxTimedLaunches) ++te->pev_steps;
    }
    return 0;
}

/* Phase times of the peer-memory steps since gof_engine_timing(e, 1), summed over `*steps` steps, in
 * milliseconds: out4 = {start -> sorted (hash, migration, sort), sorted -> interior forces done,
 * sorted -> boundary forces done (halo push, wait, boundary layers; side stream), start -> end}.
 * Synchronises on the recorded events. */
extern "C" int gof_slab_phase_times(gof_engine* e, float* out4, int* steps)
{
    DeviceGuard guard(e ? e->device : -1);
    if (int rc = slab_ready(e, false)) return rc;
    if (!out4 || !steps) return fail(GOF_E_INVALID, "gof_slab_phase_times: null pointer");
    out4[0] = out4[1] = out4[2] = out4[3] = 0.f;
    for (int k = 0; k < e->pev_steps; ++k) {
        cudaEvent_t* v = &e->pev[(size_t)k * 5];
        GOF_CUDA(cudaEventSynchronize(v[4]));
        GOF_CUDA(cudaEventSynchronize(v[3]));
        float t;
        GOF_CUDA(cudaEventElapsedTime(&t, v[0], v[1])); out4[0] += t;
        GOF_CUDA(cudaEventElapsedTime(&t, v[1], v[2])); out4[1] += t;
        GOF_CUDA(cudaEventElapsedTime(&t, v[1], v[3])); out4[2] += t;
        GOF_CUDA(cudaEventElapsedTime(&t, v[0], v[4])); out4[3] += t;
    }
    *steps = e->pev_steps;
    return 0;
}

/* CUDA IPC plumbing for gof_slab_connect between processes: export the allocation that holds
 * `d_ptr` (64-byte handle + offset of d_ptr inside it), open a peer's handle, close it. */
extern "C" int gof_ipc_export(const void* d_ptr, void* handle64, int64_t* offset)
{
    if (!d_ptr || !handle64 || !offset) return fail(GOF_E_INVALID, "gof_ipc_export: null pointer");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
    // base of the allocation: driver entry point through the runtime (no link-time libcuda)
    typedef int (*range_fn)(unsigned long long*, size_t*, unsigned long long);
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qr;
    GOF_CUDA(cudaGetDriverEntryPoint("cuMemGetAddressRange", &fn, cudaEnableDefault, &qr));
    if (!fn || qr != cudaDriverEntryPointSuccess) return fail(GOF_E_STATE, "gof_ipc_export: cuMemGetAddressRange unavailable");
    unsigned long long base = 0;
    size_t size = 0;
    if (int cr = ((range_fn)fn)(&base, &size, (unsigned long long)(uintptr_t)d_ptr))
        return fail(GOF_E_STATE, "gof_ipc_export: cuMemGetAddressRange failed (" + std::to_string(cr) + ")");
    cudaIpcMemHandle_t h;
    GOF_CUDA(cudaIpcGetMemHandle(&h, (void*)(uintptr_t)base));
    std::memcpy(handle64, &h, 64);
    *offset = (int64_t)((unsigned long long)(uintptr_t)d_ptr - base);
    return 0;
}

extern "C" int gof_ipc_open(const void* handle64, void** d_base)
{
    if (!handle64 || !d_base) return fail(GOF_E_INVALID, "gof_ipc_open: null pointer");
    cudaIpcMemHandle_t h;
    std::memcpy(&h, handle64, 64);
    GOF_CUDA(cudaIpcOpenMemHandle(d_base, h, cudaIpcMemLazyEnablePeerAccess));
    return 0;
}

extern "C" int gof_ipc_close(void* d_base)
{
    if (!d_base) return 0;
    GOF_CUDA(cudaIpcCloseMemHandle(d_base));
    return 0;
}

/* Current parity (which of the two max-speed slots describes the resident velocities). */
extern "C" int gof_slab_parity(const gof_engine* e) { return e ? e->parity : 0; }
extern "C" int gof_slab_count(const gof_engine* e) { return e ? e->n : 0; }

/* State in rest order with the particle ids (the host scatters by id).  Arrays hold
 * gof_slab_count() entries; pointers may be host or device. */
extern "C" int gof_slab_download(gof_engine* e, float* px, float* py, float* pz, float* vx, float* vy, float* vz,
                                 float* ax, float* ay, float* az, int32_t* ids, int32_t* types, void* stream)
{
    DeviceGuard guard(e ? e->device : -1);
    if (int rc = slab_ready(e, true)) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    if (e->p2p.on)
        if (int rc = slab_p2p_sync_count(e, st)) return rc;
    const int n = e->n;
    if (n == 0) return 0;
    GOF_LAUNCH(k_unpack_rest, cdiv(n, kEwThreads), kEwThreads, 0, st, e->posw[0], e->veli[0], e->acc, n, e->stage[0],
               e->stage[1], e->stage[2], e->stage[3], e->stage[4], e->stage[5], e->stage[6], e->stage[7],
               e->stage[8], e->stage_bins, e->stage_types);
    float* dst[9] = {px, py, pz, vx, vy, vz, ax, ay, az};
    for (int k = 0; k < 9; ++k)
        if (dst[k]) GOF_CUDA(cudaMemcpyAsync(dst[k], e->stage[k], (size_t)n * sizeof(float), cudaMemcpyDefault, st));
    if (ids) GOF_CUDA(cudaMemcpyAsync(ids, e->stage_bins, (size_t)n * sizeof(int), cudaMemcpyDefault, st));
    if (types) GOF_CUDA(cudaMemcpyAsync(types, e->stage_types, (size_t)n * sizeof(int), cudaMemcpyDefault, st));
    return 0;
}
