// common.cuh -- shared host/device helpers for the nimrud_b200 CUDA library (sm_100a).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>
#include <string>
#include <utility>

#include "../../include/nimrud_b200.h"

namespace nbr {

// ------------------------------------------------------------------------------------------------
// errors, launch counter
// ------------------------------------------------------------------------------------------------
void set_error(const std::string &msg);
int fail(int code, const std::string &msg);
extern std::atomic<int64_t> g_launches;

#define NBR_CUDA(expr)                                                                          \
    do {                                                                                        \
        cudaError_t _e = (expr);                                                                \
        if (_e != cudaSuccess)                                                                  \
            return ::nbr::fail(NBR_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)); \
    } while (0)

#define NBR_TRY(expr)                  \
    do {                               \
        int _rc = (expr);              \
        if (_rc != NBR_OK) return _rc; \
    } while (0)

// count + check a kernel launch
#define NBR_LAUNCHED()                                                                        \
    do {                                                                                      \
        ::nbr::g_launches.fetch_add(1, std::memory_order_relaxed);                            \
        cudaError_t _e = cudaGetLastError();                                                  \
        if (_e != cudaSuccess)                                                                \
            return ::nbr::fail(NBR_ERR_CUDA, std::string("kernel launch: ") + cudaGetErrorString(_e)); \
    } while (0)

// stream-ordered scratch buffer (RAII; freed on the same stream)
struct Scratch {
    void *ptr = nullptr;
    cudaStream_t stream = nullptr;
    Scratch() {}
    Scratch(const Scratch &) = delete;
    Scratch &operator=(const Scratch &) = delete;
    ~Scratch() { release(); }
    int alloc(size_t bytes, cudaStream_t s);
    void release();
    template <typename T>
    T *as() const { return reinterpret_cast<T *>(ptr); }
    void swap(Scratch &o) { std::swap(ptr, o.ptr); std::swap(stream, o.stream); }
};

int device_sm_count();
// stream-ordered allocation from the library's PRIVATE memory pool of the current device (the default pool and its
// attributes belong to the application).  freed with cudaFreeAsync.  the pool keeps up to NBR_POOL_KEEP_MB
// (default: 5 % of the device's memory, at least 2 GB) of freed scratch for the next call; nbr_trim_memory() drops it.
cudaError_t pool_alloc(void **ptr, size_t bytes, cudaStream_t stream);
// true the first time it is called with this flag word on the current device (kernel attributes such as the
// dynamic shared memory limit are per device: a process that drives several GPUs has to set them on each)
bool first_use_on_device(std::atomic<uint64_t> &seen);

// optional per-phase device timing (CUDA events on the launching stream); see nbr_timing_*
enum Phase { PHASE_BBOX = 0, PHASE_INDEX = 1, PHASE_ORDER = 2, PHASE_FEATURES = 3, PHASE_BOXES = 4, PHASE_PUSH = 5, PHASE_HALO_WAIT = 6, PHASE_COUNT = 8 };
struct PhaseTimer {
    int phase;
    cudaStream_t stream;
    cudaEvent_t start = nullptr, stop = nullptr;
    PhaseTimer(int phase, cudaStream_t stream);
    ~PhaseTimer();
};

static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// ------------------------------------------------------------------------------------------------
// device-side grid description (copied by value into kernels)
// ------------------------------------------------------------------------------------------------
struct GridDev {
    double minc[3];
    double edge;
    double inv_edge;
    int32_t widths[3];
    int32_t shifts[3];
    int32_t ncell[3];   // cells covered by this lattice's directory (local cell coordinates 0..ncell-1)
    int32_t cell_lo[3]; // global cell index of local cell 0: local = floor((p - minc)/e) - cell_lo.  non-zero
                        // when the lattice is anchored on a global box (multi-GPU) but only covers a tile
    int32_t ndim;
};

// brick geometry: 32 (x) * 8 (y) * 4 (z) cells = 32 words of 32 bits = 128 bytes
constexpr int BRICK_X = 32, BRICK_Y = 8, BRICK_Z = 4, BRICK_WORDS = 32;
constexpr int BRICK_YS = 3, BRICK_ZS = 2, BRICK_XS = 5;

struct LatticeDev {
    GridDev g;
    int32_t nbx, nby, nbz;        // directory dimensions (bricks)
    const uint32_t *dir;          // [nbz][nby][nbx] -> slot; 0 = empty (slot 0 is an all-zero brick)
    const uint32_t *pool;         // [slots][32] occupancy words; word = (z&3)*8 + (y&7), bit = x&31
    const uint32_t *rowbase;      // [slots][32] index (np.unique order) of the first voxel of the row; or NULL
    const uint64_t *ukeys;        // sorted unique packed addresses (np.unique order); or NULL
};

// what the query order (order.cu) leaves behind for the lattice build when its cells ARE the bricks of the finest
// lattice (cell coordinates from the lattice's own exact cell arithmetic): the first ordered point of every cell.
// cells are numbered in blocks of 8 x 8 x 8, Z-curve inside a block (order.cu); the bricks a cell's points can touch
// on every lattice follow from the cell alone, so the build marks bricks per occupied CELL instead of per point.
struct CellOrderInfo {
    Scratch offsets;            // uint32 [n_cells]: exclusive scan of the cell populations
    int64_t n_cells = 0;
    int64_t n_points = 0;
    int32_t bdims[3] = {0, 0, 0};   // blocks of 8 x 8 x 8 cells per axis
    int32_t dims[3] = {0, 0, 0};    // cells per axis (== bricks of the finest lattice)
    double finest = 0.0;            // edge of the lattice whose bricks the cells are
    bool valid = false;
};

#ifdef __CUDACC__
// ------------------------------------------------------------------------------------------------
// exact float64 arithmetic of the reference (no fma contraction anywhere on these paths)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ double load_coord(const void *xyz, int dtype, int64_t i, int ndim, int a)
{
    if (a >= ndim) return 0.0;
    return dtype == NBR_F32 ? (double)reinterpret_cast<const float *>(xyz)[i * ndim + a]
                            : reinterpret_cast<const double *>(xyz)[i * ndim + a];
}

// floor((p - min_corner) / e)  -- utils/geometry.py:107
__device__ __forceinline__ double cell_coord_f(double p, double minc, double edge)
{
    return floor(__ddiv_rn(__dsub_rn(p, minc), edge));
}

// the same value without the division in the common case: u = (p - min_corner) * (1/e) is within a few
// ulp of the quotient, so floor(u) can only differ from floor(quotient) when u sits next to an integer;
// only then is the division done.  bit-identical to cell_coord_f.
__device__ __forceinline__ double cell_coord_fast(double p, double minc, double edge, double inv_edge)
{
    const double d = __dsub_rn(p, minc);
    const double u = d * inv_edge;
    const double r = rint(u);
    if (fabs(u - r) <= 4.0e-15 * fabs(u) + 1e-300) return floor(__ddiv_rn(d, edge));
    return floor(u);
}

// (k*e + min_corner) + e*0.5 -- utils/geometry.py:137
__device__ __forceinline__ double cell_centre(int64_t k, double minc, double edge)
{
    return __dadd_rn(__dadd_rn(__dmul_rn((double)k, edge), minc), __dmul_rn(edge, 0.5));
}

// centre of LOCAL cell k on axis a
struct GridDev;
__device__ __forceinline__ double grid_centre(const GridDev &g, long long k, int a);

// one squared coordinate difference; the membership sum is ((dx2 + dy2) + dz2) <= r2
__device__ __forceinline__ double sqdiff(double q, double c)
{
    double d = __dsub_rn(q, c);
    return __dmul_rn(d, d);
}

__device__ __forceinline__ double grid_centre(const GridDev &g, long long k, int a)
{
    return cell_centre(k + (long long)g.cell_lo[a], g.minc[a], g.edge);
}

__device__ __forceinline__ uint32_t lanemask_lt()
{
    uint32_t m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}

__device__ __forceinline__ int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

// rare path of point_cell: u = (p - minc) * (1/e) sits next to an integer, only the division tells the cell
static __device__ __noinline__ int cell_by_division(double d, double edge)
{
    return (int)fmin(fmax(floor(__ddiv_rn(d, edge)), -2.0e9), 2.0e9);
}

// LOCAL cell of point i on every axis: floor((p - min_corner) / e) - cell_lo  (utils/geometry.py:107),
// bit-identical to cell_coord_f but with one multiply, one float64 -> int conversion and a short test in
// the common case (the quotient is only formed when the product is within a few ulp of an integer)
template <typename T>
__device__ __forceinline__ void point_cell(const T *__restrict__ xyz, int64_t i, const GridDev &g, int c[3])
{
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const double d = __dsub_rn((double)xyz[i * 3 + a], g.minc[a]);
        const double u = d * g.inv_edge;
        int k = __double2int_rd(u);
        const double fr = u - (double)k;                  // in [0, 1] when u is in int range
        const double tol = 4.0e-15 * fabs(u) + 1e-300;
        if (!(fr > tol && fr < 1.0 - tol)) k = cell_by_division(d, g.edge);
        // search points lie inside the covered range by construction; the clamp defends against a
        // caller-supplied box that does not contain them.
        c[a] = clampi(k - g.cell_lo[a], 0, g.ncell[a] - 1);
    }
}

#endif

}  // namespace nbr
