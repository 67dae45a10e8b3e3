// index.cu -- the spatial index: voxel grid parameters, packed addresses, unique voxels, centres, and
// the occupancy bit-brick lattice that the query kernels read.
//
// reference: nimrud/utils/geometry.py:16-154 (VoxelFilter) and the search_voxels/cKDTree build of
// nimrud/minimal/multiscale.py:75-87.
#include <math.h>

#include <stdlib.h>

#include <memory>
#include <string>
#include <vector>

#include "common.cuh"
#include "lattice.cuh"
#include "mailbox.cuh"
#include "scan.cuh"

namespace nbr {

int sort_keys(uint64_t *keys, uint64_t *tmp, int64_t n, int begin_bit, int end_bit, cudaStream_t stream);

// ------------------------------------------------------------------------------------------------
// bounding box
// ------------------------------------------------------------------------------------------------
constexpr int BBOX_THREADS = 256;

__device__ __forceinline__ double warp_min(double v)
{
#pragma unroll
    for (int o = 16; o; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ double warp_max(double v)
{
#pragma unroll
    for (int o = 16; o; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// partial[block][6].  one point per thread and trip; float32 clouds are reduced in float32 (exact).
template <typename T, int NDIM>
__global__ void __launch_bounds__(BBOX_THREADS)
bbox_partial_kernel(const T *__restrict__ xyz, int64_t n, double *__restrict__ partial)
{
    T lo[3], hi[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) { lo[k] = (T)INFINITY; hi[k] = (T)-INFINITY; }
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
#pragma unroll
        for (int k = 0; k < NDIM; ++k) {
            const T v = xyz[i * NDIM + k];
            lo[k] = v < lo[k] ? v : lo[k];
            hi[k] = v > hi[k] ? v : hi[k];
        }
    }
    __shared__ double s[BBOX_THREADS / 32][6];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        double a = warp_min((double)lo[k]), b = warp_max((double)hi[k]);
        if (lane == 0) { s[warp][k] = a; s[warp][3 + k] = b; }
    }
    __syncthreads();
    if (threadIdx.x < 6) {
        double r = s[0][threadIdx.x];
        for (int w = 1; w < BBOX_THREADS / 32; ++w)
            r = threadIdx.x < 3 ? fmin(r, s[w][threadIdx.x]) : fmax(r, s[w][threadIdx.x]);
        partial[blockIdx.x * 6 + threadIdx.x] = r;
    }
}

// 6 warps, one per output value
__global__ void bbox_final_kernel(const double *__restrict__ partial, int blocks, int ndim, double *__restrict__ out)
{
    const int k = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double r = k < 3 ? INFINITY : -INFINITY;
    for (int b = lane; b < blocks; b += 32) r = k < 3 ? fmin(r, partial[b * 6 + k]) : fmax(r, partial[b * 6 + k]);
    r = k < 3 ? warp_min(r) : warp_max(r);
    if ((k % 3) >= ndim) r = 0.0;
    if (lane == 0) out[k] = r;
}

int bbox(const void *xyz, int dtype, int64_t n, int ndim, double *lohi_dev, cudaStream_t stream)
{
    if (n < 1) return fail(NBR_ERR_TOO_FEW_POINTS, "need at least 1 point for a bounding box");
    int blocks = (int)std::min<int64_t>(ceil_div(n, BBOX_THREADS * 4), device_sm_count() * 8);
    if (blocks < 1) blocks = 1;
    Scratch partial;
    NBR_TRY(partial.alloc(sizeof(double) * 6 * blocks, stream));
    double *pp = partial.as<double>();
    if (dtype == NBR_F32 && ndim == 3)
        bbox_partial_kernel<float, 3><<<blocks, BBOX_THREADS, 0, stream>>>((const float *)xyz, n, pp);
    else if (dtype == NBR_F32)
        bbox_partial_kernel<float, 2><<<blocks, BBOX_THREADS, 0, stream>>>((const float *)xyz, n, pp);
    else if (ndim == 3)
        bbox_partial_kernel<double, 3><<<blocks, BBOX_THREADS, 0, stream>>>((const double *)xyz, n, pp);
    else
        bbox_partial_kernel<double, 2><<<blocks, BBOX_THREADS, 0, stream>>>((const double *)xyz, n, pp);
    NBR_LAUNCHED();
    bbox_final_kernel<<<1, 192, 0, stream>>>(partial.as<double>(), blocks, ndim, lohi_dev);
    NBR_LAUNCHED();
    return NBR_OK;
}

// ------------------------------------------------------------------------------------------------
// grid parameters (host)   utils/geometry.py:37-62
// ------------------------------------------------------------------------------------------------
int grid_from_bbox(const double lo[3], const double hi[3], double edge, int ndim, nbr_grid *out)
{
    if (!(edge > 0) || (ndim != 2 && ndim != 3)) return fail(NBR_ERR_INVALID, "grid: edge must be > 0, ndim 2 or 3");
    memset(out, 0, sizeof(*out));
    out->edge = edge;
    out->ndim = ndim;
    double total = 0;
    int shift = 0;
    for (int a = 0; a < 3; ++a) {
        if (a >= ndim) { out->widths[a] = 0; out->shifts[a] = shift; continue; }
        out->min_corner[a] = lo[a] - edge / 2;
        out->max_corner[a] = hi[a] + edge / 2;
        double span = out->max_corner[a] - out->min_corner[a];
        double w = ceil(log2(span / edge));
        if (!(w >= 0)) w = 0;     // degenerate (span <= e): the reference breaks here (int("0b")), we use 0 bits
        total += w;
        out->widths[a] = (int32_t)w;
        out->shifts[a] = shift;
        shift += (int32_t)w;
    }
    if (total > 64) return fail(NBR_ERR_ADDRESS_BITS, "edge length is too small to address this space");
    return NBR_OK;
}

// local_lohi (optional): bounding box of the points this lattice will hold; the directory then only
// covers that part of the (possibly much larger, global) grid.
int grid_to_dev(const nbr_grid *g, GridDev *d, const double *local_lohi)
{
    for (int a = 0; a < 3; ++a) {
        d->minc[a] = g->min_corner[a];
        d->widths[a] = g->widths[a];
        d->shifts[a] = g->shifts[a];
        double first = 0.0;
        double last = a < g->ndim ? floor((g->max_corner[a] - g->min_corner[a]) / g->edge) : 0.0;
        if (local_lohi && a < g->ndim) {
            // one cell of slack on both sides: host and device floor() agree to the last bit only in theory
            first = fmax(first, floor((local_lohi[a] - g->min_corner[a]) / g->edge) - 1.0);
            last = fmin(last, floor((local_lohi[3 + a] - g->min_corner[a]) / g->edge) + 1.0);
            if (last < first) last = first;
        }
        if (last - first + 1.0 > 2147483000.0 || first > 2147483000.0)
            return fail(NBR_ERR_UNSUPPORTED, "more than 2^31 cells along one axis");
        d->cell_lo[a] = (int32_t)first;
        d->ncell[a] = (int32_t)(last - first + 1.0);
    }
    d->edge = g->edge;
    d->inv_edge = 1.0 / g->edge;
    d->ndim = g->ndim;
    return NBR_OK;
}

// ------------------------------------------------------------------------------------------------
// addresses / centres      utils/geometry.py:103-138
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
address_kernel(const void *__restrict__ xyz, int dtype, int64_t n, GridDev g, double maxc0, double maxc1, double maxc2,
               int64_t *__restrict__ addr, int32_t *__restrict__ oob)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double maxc[3] = {maxc0, maxc1, maxc2};
    int64_t a64 = 0;
    bool bad = false;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        if (a < g.ndim) {
            double p = load_coord(xyz, dtype, i, g.ndim, a);
            bad |= (p < g.minc[a]) | (p > maxc[a]);
            int64_t k = (int64_t)cell_coord_fast(p, g.minc[a], g.edge, g.inv_edge);
            a64 += k << g.shifts[a];
        }
    }
    addr[i] = a64;
    if (bad && oob) *oob = 1;
}

__global__ void __launch_bounds__(256)
centre_kernel(const int64_t *__restrict__ addr, int64_t n, GridDev g, double *__restrict__ out)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int64_t k = addr[i];
#pragma unroll
    for (int a = 0; a < 3; ++a)
        if (a < g.ndim) {
            int64_t mask = g.widths[a] >= 64 ? -1ll : ((1ll << g.widths[a]) - 1);
            int64_t c = (k >> g.shifts[a]) & mask;
            out[i * g.ndim + a] = cell_centre(c, g.minc[a], g.edge);
        }
}

// ------------------------------------------------------------------------------------------------
// unique of a sorted array
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
head_flag_kernel(const uint64_t *__restrict__ keys, int64_t n, uint32_t *__restrict__ flags)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    flags[i] = (i == 0 || keys[i] != keys[i - 1]) ? 1u : 0u;
}

__global__ void __launch_bounds__(256)
compact_kernel(const uint64_t *__restrict__ keys, const uint32_t *__restrict__ slots, int64_t n,
               uint64_t *__restrict__ out)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t s = slots[i];
    if (s) out[s - 1] = keys[i];
}

__global__ void widen_count_kernel(const uint32_t *c, int64_t *out) { *out = (int64_t)*c; }

int unique_sorted(const uint64_t *sorted, int64_t n, uint64_t *out, int64_t *n_out_dev, cudaStream_t stream)
{
    if (n <= 0) {
        NBR_CUDA(cudaMemsetAsync(n_out_dev, 0, sizeof(int64_t), stream));
        return NBR_OK;
    }
    Scratch flags, count;
    NBR_TRY(flags.alloc(sizeof(uint32_t) * n, stream));
    NBR_TRY(count.alloc(sizeof(uint32_t), stream));
    const unsigned blocks = (unsigned)ceil_div(n, 256);
    head_flag_kernel<<<blocks, 256, 0, stream>>>(sorted, n, flags.as<uint32_t>());
    NBR_LAUNCHED();
    NBR_TRY(flags_to_slots(flags.as<uint32_t>(), n, count.as<uint32_t>(), stream));
    compact_kernel<<<blocks, 256, 0, stream>>>(sorted, flags.as<uint32_t>(), n, out);
    NBR_LAUNCHED();
    widen_count_kernel<<<1, 1, 0, stream>>>(count.as<uint32_t>(), n_out_dev);
    NBR_LAUNCHED();
    return NBR_OK;
}

// ------------------------------------------------------------------------------------------------
// lattice (bit bricks)
// ------------------------------------------------------------------------------------------------
// mark / fill handle PTS points per thread with the loads of all of them in flight together: the kernels are
// bound by the latency of the dependent directory / pool accesses, not by bandwidth
#ifndef NBR_PTS
#define NBR_PTS 4
#endif
constexpr int PTS = NBR_PTS;
#ifndef NBR_FILL_AGGREGATE
#define NBR_FILL_AGGREGATE 1
#endif

template <typename T>
__global__ void __launch_bounds__(256)
brick_mark_kernel(const T *__restrict__ xyz, int64_t n, GridDev g, int nbx, int nby, uint32_t *__restrict__ dir)
{
    const int64_t base = ((int64_t)blockIdx.x * PTS) * blockDim.x + threadIdx.x;
    int64_t b[PTS];
#pragma unroll
    for (int k = 0; k < PTS; ++k) {
        const int64_t i = base + (int64_t)k * blockDim.x;
        b[k] = -1;
        if (i < n) {
            int c[3];
            point_cell<T>(xyz, i, g, c);
            b[k] = ((int64_t)(c[2] >> BRICK_ZS) * nby + (c[1] >> BRICK_YS)) * nbx + (c[0] >> BRICK_XS);
        }
    }
    uint32_t seen[PTS];
#pragma unroll
    for (int k = 0; k < PTS; ++k) seen[k] = b[k] >= 0 ? dir[b[k]] : 1u;
#pragma unroll
    for (int k = 0; k < PTS; ++k)
        if (seen[k] == 0) dir[b[k]] = 1;   // benign race: every writer stores 1
}

__global__ void __launch_bounds__(256)
pool_zero_kernel(uint32_t *__restrict__ pool, const uint32_t *__restrict__ n_bricks)
{
    const int64_t words = ((int64_t)*n_bricks + 1) * BRICK_WORDS;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < words; i += (int64_t)gridDim.x * blockDim.x)
        pool[i] = 0;
}

template <typename T>
__global__ void __launch_bounds__(256)
brick_fill_kernel(const T *__restrict__ xyz, int64_t n, GridDev g, int nbx, int nby, const uint32_t *__restrict__ dir,
                  uint32_t *__restrict__ pool)
{
    const int64_t base = ((int64_t)blockIdx.x * PTS) * blockDim.x + threadIdx.x;
    int64_t b[PTS];
    int word[PTS];
    uint32_t bit[PTS];
#pragma unroll
    for (int k = 0; k < PTS; ++k) {
        const int64_t i = base + (int64_t)k * blockDim.x;
        b[k] = -1;
        word[k] = 0;
        bit[k] = 0;
        if (i < n) {
            int c[3];
            point_cell<T>(xyz, i, g, c);
            b[k] = ((int64_t)(c[2] >> BRICK_ZS) * nby + (c[1] >> BRICK_YS)) * nbx + (c[0] >> BRICK_XS);
            word[k] = ((c[2] & (BRICK_Z - 1)) << BRICK_YS) | (c[1] & (BRICK_Y - 1));
            bit[k] = 1u << (c[0] & 31);
        }
    }
    uint32_t *w[PTS];
#pragma unroll
    for (int k = 0; k < PTS; ++k) w[k] = pool + (int64_t)(b[k] >= 0 ? dir[b[k]] : 0u) * BRICK_WORDS + word[k];
    uint32_t have[PTS];
#pragma unroll
    for (int k = 0; k < PTS; ++k) have[k] = bit[k] ? *w[k] : ~0u;
#pragma unroll
    for (int k = 0; k < PTS; ++k)
        if ((have[k] & bit[k]) == 0 && bit[k]) atomicOr(w[k], bit[k]);
}

// number of occupied voxels = popcount of the pool
__global__ void __launch_bounds__(256)
pool_count_kernel(const uint32_t *__restrict__ pool, const uint32_t *__restrict__ n_bricks,
                  unsigned long long *__restrict__ total)
{
    const int64_t words = ((int64_t)*n_bricks + 1) * BRICK_WORDS;
    unsigned long long acc = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < words; i += (int64_t)gridDim.x * blockDim.x)
        acc += __popc(pool[i]);
#pragma unroll
    for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0 && acc) atomicAdd(total, acc);
}

// for every sorted unique address that starts a (brick,row) word: rowbase[slot][word] = its rank
__global__ void __launch_bounds__(256)
rowbase_kernel(const uint64_t *__restrict__ ukeys, const int64_t *__restrict__ n_unique, GridDev g, int nbx,
               int nby, const uint32_t *__restrict__ dir, uint32_t *__restrict__ rowbase)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= *n_unique) return;
    int c[3], p[3];
    const uint64_t k = ukeys[i];
    const uint64_t kp = i ? ukeys[i - 1] : 0;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        uint64_t mask = g.widths[a] >= 64 ? ~0ull : ((1ull << g.widths[a]) - 1);
        c[a] = (int)((k >> g.shifts[a]) & mask) - g.cell_lo[a];
        p[a] = (int)((kp >> g.shifts[a]) & mask) - g.cell_lo[a];
    }
    const bool head = i == 0 || c[2] != p[2] || c[1] != p[1] || (c[0] >> BRICK_XS) != (p[0] >> BRICK_XS);
    if (!head) return;
    const int64_t b = ((int64_t)(c[2] >> BRICK_ZS) * nby + (c[1] >> BRICK_YS)) * nbx + (c[0] >> BRICK_XS);
    const int word = ((c[2] & (BRICK_Z - 1)) << BRICK_YS) | (c[1] & (BRICK_Y - 1));
    rowbase[(int64_t)dir[b] * BRICK_WORDS + word] = (uint32_t)i;
}

Lattice::SharedBuffers::~SharedBuffers()
{
    if (dir) cudaFreeAsync(dir, stream);
    if (pool) cudaFreeAsync(pool, stream);
    if (counters) cudaFreeAsync(counters, stream);
}

Lattice::~Lattice()
{
    if (shared) return;     // the batch owns the buffers
    // buffers were allocated stream-ordered; free them the same way
    if (dir) cudaFreeAsync(dir, stream);
    if (pool) cudaFreeAsync(pool, stream);
    if (rowbase) cudaFreeAsync(rowbase, stream);
    if (ukeys) cudaFreeAsync(ukeys, stream);
    if (counters) cudaFreeAsync(counters, stream);
}

LatticeDev Lattice::dev() const
{
    LatticeDev d;
    d.g = gdev;
    d.nbx = nbx; d.nby = nby; d.nbz = nbz;
    d.dir = dir; d.pool = pool; d.rowbase = rowbase; d.ukeys = ukeys;
    return d;
}

int lattice_create(Lattice **out, const void *xyz, int dtype, int64_t n, const nbr_grid *grid, int flags,
                   cudaStream_t stream, const double *local_lohi)
{
    if (!out || !xyz || !grid) return fail(NBR_ERR_INVALID, "lattice_create: null argument");
    if (dtype != NBR_F32 && dtype != NBR_F64) return fail(NBR_ERR_INVALID, "lattice_create: bad dtype");
    if (n < 1) return fail(NBR_ERR_TOO_FEW_POINTS, "lattice_create: empty search cloud");
    if (n >= (int64_t)1 << 31) return fail(NBR_ERR_UNSUPPORTED, "lattice_create: more than 2^31 search points");
    if (grid->ndim != 3) return fail(NBR_ERR_INVALID, "lattice_create: 3-D grids only");
    Lattice *L = new Lattice();
    L->stream = stream;
    L->grid = *grid;
    L->n_search = n;
    int rc = grid_to_dev(grid, &L->gdev, local_lohi);
    if (rc) { delete L; return rc; }
    L->nbx = (L->gdev.ncell[0] + BRICK_X - 1) / BRICK_X;
    L->nby = (L->gdev.ncell[1] + BRICK_Y - 1) / BRICK_Y;
    L->nbz = (L->gdev.ncell[2] + BRICK_Z - 1) / BRICK_Z;
    const double dir_entries = (double)L->nbx * L->nby * L->nbz;
    if (dir_entries > 3.0e9) {
        delete L;
        return fail(NBR_ERR_UNSUPPORTED, "brick directory would exceed 12 GB; extent / edge too large");
    }
    L->n_dir = (int64_t)L->nbx * L->nby * L->nbz;
    L->pool_slots = std::min<int64_t>(n, L->n_dir) + 1;

#define L_CUDA(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) { delete L; \
        return fail(NBR_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)); } } while (0)
#define L_TRY(expr) do { int _rc = (expr); if (_rc) { delete L; return _rc; } } while (0)
#define L_LAUNCHED() do { g_launches.fetch_add(1, std::memory_order_relaxed); cudaError_t _e = cudaGetLastError(); \
        if (_e != cudaSuccess) { delete L; return fail(NBR_ERR_CUDA, std::string("kernel launch: ") + cudaGetErrorString(_e)); } } while (0)

    L_CUDA(pool_alloc((void **)&L->dir, sizeof(uint32_t) * L->n_dir, stream));
    L_CUDA(pool_alloc((void **)&L->pool, sizeof(uint32_t) * BRICK_WORDS * L->pool_slots, stream));
    L_CUDA(pool_alloc((void **)&L->counters, 64, stream));
    L_CUDA(cudaMemsetAsync(L->dir, 0, sizeof(uint32_t) * L->n_dir, stream));
    L_CUDA(cudaMemsetAsync(L->counters, 0, 64, stream));
    uint32_t *n_bricks_dev = reinterpret_cast<uint32_t *>(L->counters);
    unsigned long long *n_vox_dev = reinterpret_cast<unsigned long long *>(L->counters + 8);
    int64_t *n_unique_dev = reinterpret_cast<int64_t *>(L->counters + 16);

    const unsigned blocks = (unsigned)ceil_div(n, 256);
    const unsigned pt_blocks = (unsigned)ceil_div(n, 256 * PTS);
    const int sweep_blocks = device_sm_count() * 8;
    if (dtype == NBR_F32)
        brick_mark_kernel<float><<<pt_blocks, 256, 0, stream>>>((const float *)xyz, n, L->gdev, L->nbx, L->nby, L->dir);
    else
        brick_mark_kernel<double><<<pt_blocks, 256, 0, stream>>>((const double *)xyz, n, L->gdev, L->nbx, L->nby, L->dir);
    L_LAUNCHED();
    L_TRY(flags_to_slots(L->dir, L->n_dir, n_bricks_dev, stream));
    pool_zero_kernel<<<sweep_blocks, 256, 0, stream>>>(L->pool, n_bricks_dev);
    L_LAUNCHED();
    if (dtype == NBR_F32)
        brick_fill_kernel<float><<<pt_blocks, 256, 0, stream>>>((const float *)xyz, n, L->gdev, L->nbx, L->nby, L->dir, L->pool);
    else
        brick_fill_kernel<double><<<pt_blocks, 256, 0, stream>>>((const double *)xyz, n, L->gdev, L->nbx, L->nby, L->dir, L->pool);
    L_LAUNCHED();
    pool_count_kernel<<<sweep_blocks, 256, 0, stream>>>(L->pool, n_bricks_dev, n_vox_dev);
    L_LAUNCHED();

    if (flags & NBR_LATTICE_INDEXED) {
        // np.unique order: sort the packed addresses, dedup, remember the rank of each row head
        Scratch addr, tmp;
        L_TRY(addr.alloc(sizeof(uint64_t) * n, stream));
        L_TRY(tmp.alloc(sizeof(uint64_t) * n, stream));
        L_CUDA(pool_alloc((void **)&L->ukeys, sizeof(uint64_t) * n, stream));
        L_CUDA(pool_alloc((void **)&L->rowbase, sizeof(uint32_t) * BRICK_WORDS * L->pool_slots, stream));
        address_kernel<<<blocks, 256, 0, stream>>>(xyz, dtype, n, L->gdev, grid->max_corner[0], grid->max_corner[1],
                                                   grid->max_corner[2], addr.as<int64_t>(), nullptr);
        L_LAUNCHED();
        const int bits = grid->widths[0] + grid->widths[1] + grid->widths[2];
        L_TRY(sort_keys(addr.as<uint64_t>(), tmp.as<uint64_t>(), n, 0, bits, stream));
        L_TRY(unique_sorted(addr.as<uint64_t>(), n, L->ukeys, n_unique_dev, stream));
        rowbase_kernel<<<blocks, 256, 0, stream>>>(L->ukeys, n_unique_dev, L->gdev, L->nbx, L->nby, L->dir, L->rowbase);
        L_LAUNCHED();
        L->indexed = true;
    }
    *out = L;
    return NBR_OK;
#undef L_CUDA
#undef L_TRY
#undef L_LAUNCHED
}

// ------------------------------------------------------------------------------------------------
// batch build: every lattice of a call in ONE pass over the points per phase.  the directories are
// concatenated, one scan numbers the occupied bricks of all lattices (slot ids are global, one shared pool,
// slot 0 = the all-zero brick), the fill pass counts each lattice's voxels as it sets the bits (the thread
// whose atomicOr flips a bit owns that voxel).  7 launches per call instead of ~10 per lattice, and the
// points are read twice instead of 2 x n_lat times.
// ------------------------------------------------------------------------------------------------
struct BatchDev {
    GridDev g[LATTICE_BATCH];
    int64_t dir_off[LATTICE_BATCH];
    int nbx[LATTICE_BATCH], nby[LATTICE_BATCH];
    int n;
};

// DEV_N: the number of points is only known on the device (halo mailbox): n_bound sizes a capped grid that strides
// over the chunks.  otherwise one chunk per block (the common path keeps its straight-line code)
template <typename T, bool DEV_N>
__global__ void __launch_bounds__(256)
batch_mark_kernel(const T *__restrict__ xyz, int64_t n_bound, const unsigned long long *__restrict__ n_dev,
                  const __grid_constant__ BatchDev B, uint32_t *__restrict__ dir)
{
    const int64_t n = DEV_N ? min(n_bound, (int64_t)*n_dev) : n_bound;
    for (int64_t chunk = blockIdx.x; DEV_N ? chunk * PTS * blockDim.x < n : chunk == blockIdx.x; chunk += gridDim.x) {
    const int64_t base = (chunk * PTS) * blockDim.x + threadIdx.x;
    if (DEV_N && base - (threadIdx.x & 31) >= n) return;    // the whole warp is past the end
    for (int l = 0; l < B.n; ++l) {
        int64_t b[PTS];
#pragma unroll
        for (int k = 0; k < PTS; ++k) {
            const int64_t i = base + (int64_t)k * blockDim.x;
            b[k] = -1;
            if (i < n) {
                int c[3];
                point_cell<T>(xyz, i, B.g[l], c);
                b[k] = B.dir_off[l] + ((int64_t)(c[2] >> BRICK_ZS) * B.nby[l] + (c[1] >> BRICK_YS)) * B.nbx[l] + (c[0] >> BRICK_XS);
            }
        }
        uint32_t seen[PTS];
#pragma unroll
        for (int k = 0; k < PTS; ++k) seen[k] = b[k] >= 0 ? dir[b[k]] : 1u;
#pragma unroll
        for (int k = 0; k < PTS; ++k)
            if (seen[k] == 0) dir[b[k]] = 1;   // benign race: every writer stores 1
    }
    }
}

template <typename T, bool DEV_N>
__global__ void __launch_bounds__(256)
batch_fill_kernel(const T *__restrict__ xyz, int64_t n_bound, const unsigned long long *__restrict__ n_dev,
                  const __grid_constant__ BatchDev B, const uint32_t *__restrict__ dir,
                  uint32_t *__restrict__ pool, unsigned char *__restrict__ counters)
{
    const int64_t n = DEV_N ? min(n_bound, (int64_t)*n_dev) : n_bound;
    for (int64_t chunk = blockIdx.x; DEV_N ? chunk * PTS * blockDim.x < n : chunk == blockIdx.x; chunk += gridDim.x) {
    const int64_t base = (chunk * PTS) * blockDim.x + threadIdx.x;
    if (DEV_N && base - (threadIdx.x & 31) >= n) return;    // whole warp past the end: nothing to vote on
    uint32_t pend_old[PTS], pend_bit[PTS];
#pragma unroll
    for (int k = 0; k < PTS; ++k) { pend_old[k] = ~0u; pend_bit[k] = 0; }
    for (int l = 0; l < B.n; ++l) {
        int64_t b[PTS];
        int word[PTS];
        uint32_t bit[PTS];
#pragma unroll
        for (int k = 0; k < PTS; ++k) {
            const int64_t i = base + (int64_t)k * blockDim.x;
            b[k] = -1;
            word[k] = 0;
            bit[k] = 0;
            if (i < n) {
                int c[3];
                point_cell<T>(xyz, i, B.g[l], c);
                b[k] = B.dir_off[l] + ((int64_t)(c[2] >> BRICK_ZS) * B.nby[l] + (c[1] >> BRICK_YS)) * B.nbx[l] + (c[0] >> BRICK_XS);
                word[k] = ((c[2] & (BRICK_Z - 1)) << BRICK_YS) | (c[1] & (BRICK_Y - 1));
                bit[k] = 1u << (c[0] & 31);
            }
        }
        uint32_t *w[PTS];
#pragma unroll
        for (int k = 0; k < PTS; ++k) w[k] = pool + (int64_t)(b[k] >= 0 ? dir[b[k]] : 0u) * BRICK_WORDS + word[k];
        uint32_t have[PTS];
#pragma unroll
        for (int k = 0; k < PTS; ++k) have[k] = bit[k] ? *w[k] : ~0u;
        // the atomics of the previous lattice have had this lattice's loads to complete: count its new voxels now
        if (l > 0) {
            int fresh = 0;
#pragma unroll
            for (int k = 0; k < PTS; ++k) fresh += __popc(pend_bit[k] & ~pend_old[k]);
            fresh = __reduce_add_sync(0xffffffffu, fresh);
            if ((threadIdx.x & 31) == 0 && fresh)
                atomicAdd(reinterpret_cast<unsigned long long *>(counters + 64 * (l - 1) + 8), (unsigned long long)fresh);
        }
#pragma unroll
        for (int k = 0; k < PTS; ++k) {
#if NBR_FILL_AGGREGATE
            // the points are in cell order: on the coarser lattices the 32 points of a warp usually share one
            // occupancy word.  then one lane sets all their bits with a single atomic (same-address atomics
            // serialise in L2) and owns the voxels that were new
            int same = 0;
            __match_all_sync(0xffffffffu, (unsigned long long)w[k], &same);
            if (same) {
                const uint32_t bits = __reduce_or_sync(0xffffffffu, bit[k]);
                const uint32_t seen = __reduce_and_sync(0xffffffffu, have[k]);      // every lane read the same word
                const bool lead = (threadIdx.x & 31) == 0;
                const bool need = lead && (bits & ~seen) != 0;
                pend_old[k] = need ? atomicOr(w[k], bits) : ~0u;
                pend_bit[k] = lead ? bits : 0u;
            } else
#endif
            {
                const bool need = (have[k] & bit[k]) == 0 && bit[k];
                pend_old[k] = need ? atomicOr(w[k], bit[k]) : ~0u;
                pend_bit[k] = bit[k];
            }
        }
    }
    {
        int fresh = 0;
#pragma unroll
        for (int k = 0; k < PTS; ++k) fresh += __popc(pend_bit[k] & ~pend_old[k]);
        fresh = __reduce_add_sync(0xffffffffu, fresh);
        if ((threadIdx.x & 31) == 0 && fresh && B.n > 0)
            atomicAdd(reinterpret_cast<unsigned long long *>(counters + 64 * (B.n - 1) + 8), (unsigned long long)fresh);
    }
    }
}

// bricks of every lattice that the points of an occupied ORDER CELL can touch (order.cu: the cells are the bricks of
// the finest lattice of the batch, so for that lattice the cell IS the brick; for a coarser lattice the cell's box
// [P_lo, P_hi) maps to a range of its cells, at most 2 x 2 x 2 bricks).  one thread per cell of the dense cell array
// instead of one directory read + store per point and lattice.  a marked brick may stay empty (the range is a
// superset by a 1e-6-cell slack against the rounding of the two cell computations): it costs a zero brick.
struct CellsDev {
    int bdims[3];
    int dims[3];
    int finest;                          // index of the lattice whose bricks the cells are
    // cell coordinate c (in bricks of the finest lattice) -> position of the cell's low face on axis a, in cells of
    // lattice l (global numbering):  alpha[l][a] + beta[l][a] * c ;  the high face is one beta further
    double alpha[LATTICE_BATCH][3], beta[LATTICE_BATCH][3];
};

// one WARP per 512 cells of the order (8 x 8 x 8, Z-curve inside; no block-wide barrier: the latency chains of the
// 64 warps of an SM overlap): the occupied cells of the group are compacted in the warp's shared memory, then a lane
// per occupied cell (a few dozen per group) takes the lattices in turn
constexpr int CM_WARPS = 8;
__global__ void __launch_bounds__(CM_WARPS * 32)
cells_mark_kernel(const uint32_t *__restrict__ offsets, int64_t n_cells, int64_t n_points, int64_t n_groups,
                  const __grid_constant__ BatchDev B, const __grid_constant__ CellsDev C, uint32_t *__restrict__ dir)
{
    __shared__ unsigned short s_list[CM_WARPS][512];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t group = (int64_t)blockIdx.x * CM_WARPS + warp;
    if (group >= n_groups) return;
    const int64_t c0 = group * 512;
    // lane owns cells c0 + 16 lane .. + 15: 17 consecutive offsets
    uint32_t off[17];
#pragma unroll
    for (int k = 0; k < 17; ++k) {
        const int64_t c = c0 + 16 * lane + k;
        off[k] = c < n_cells ? offsets[c] : (uint32_t)n_points;
    }
    uint32_t mine = 0;
#pragma unroll
    for (int k = 0; k < 16; ++k) mine |= (off[k + 1] > off[k] ? 1u : 0u) << k;
    if (!__any_sync(0xffffffffu, mine != 0)) return;     // no point in these 512 cells
    // exclusive prefix of the lanes' counts
    const int cnt = __popc(mine);
    int pre = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, pre, o);
        if (lane >= o) pre += t;
    }
    const int n_occ = __shfl_sync(0xffffffffu, pre, 31);
    int pos = pre - cnt;
    for (uint32_t m = mine; m; m &= m - 1) {
        const uint32_t i = 16u * (uint32_t)lane + (uint32_t)(__ffs(m) - 1);      // Z-curve position inside the group
        uint32_t lx = 0, ly = 0, lz = 0;
#pragma unroll
        for (int b = 0; b < 3; ++b) {
            lx |= ((i >> (3 * b)) & 1u) << b;
            ly |= ((i >> (3 * b + 1)) & 1u) << b;
            lz |= ((i >> (3 * b + 2)) & 1u) << b;
        }
        s_list[warp][pos++] = (unsigned short)(lx | (ly << 3) | (lz << 6));
    }
    __syncwarp();
    const uint32_t block = (uint32_t)group;
    const int sb[3] = {(int)(block % (uint32_t)C.bdims[0]) * 8, (int)((block / (uint32_t)C.bdims[0]) % (uint32_t)C.bdims[1]) * 8,
                       (int)(block / ((uint32_t)C.bdims[0] * (uint32_t)C.bdims[1])) * 8};
    for (int ci = lane; ci < n_occ; ci += 32) {
        const uint32_t pk = s_list[warp][ci];
        const int cc[3] = {sb[0] + (int)(pk & 7u), sb[1] + (int)((pk >> 3) & 7u), sb[2] + (int)(pk >> 6)};
        for (int l = 0; l < B.n; ++l) {
            int b0[3], nb[3];
            if (l == C.finest) {
#pragma unroll
                for (int a = 0; a < 3; ++a) { b0[a] = cc[a]; nb[a] = 0; }
            } else {
                const GridDev &g = B.g[l];
                const int shift[3] = {BRICK_XS, BRICK_YS, BRICK_ZS};
#pragma unroll
                for (int a = 0; a < 3; ++a) {
                    // saturating floor conversions; the slack covers the rounding of this map and of the two cell computations
                    const double u0 = fma(C.beta[l][a], (double)cc[a], C.alpha[l][a]);
                    const int k0 = max(min(__double2int_rd(u0 - 1.0e-6) - g.cell_lo[a], g.ncell[a] - 1), 0);
                    const int k1 = max(min(__double2int_rd(u0 + C.beta[l][a] + 1.0e-6) - g.cell_lo[a], g.ncell[a] - 1), 0);
                    b0[a] = k0 >> shift[a];
                    nb[a] = (k1 >> shift[a]) - b0[a];            // 0 or 1: a cell is never longer than a brick of a coarser lattice
                }
            }
            uint32_t *e = dir + B.dir_off[l] + ((int64_t)b0[2] * B.nby[l] + b0[1]) * B.nbx[l] + b0[0];
            const int64_t sy = B.nbx[l], sz = (int64_t)B.nby[l] * B.nbx[l];
            // unconditional stores (a read first would put its latency on every cell); every writer stores 1
            for (int z = 0; z <= nb[2]; ++z)
                for (int y = 0; y <= nb[1]; ++y)
                    for (int x = 0; x <= nb[0]; ++x) e[z * sz + y * sy + x] = 1u;
        }
    }
}

int halo_wait(Mailbox *M, cudaStream_t stream);

int lattices_create_batch(Lattice **out, int n_lat, const nbr_grid *grids, const void *xyz, int dtype, int64_t n,
                          cudaStream_t stream, const double *local_lohi, const void *xyz2, int64_t n2, Mailbox *mailbox,
                          const CellOrderInfo *order)
{
    // mailbox: the second part of the search cloud is what the peers pushed into this rank's halo mailbox; its
    // size stays on the device, the grids are sized for the mailbox's capacity
    const unsigned long long *n2_dev = nullptr;
    if (mailbox) {
        if (mailbox->dtype != dtype) return fail(NBR_ERR_INVALID, "lattices_create_batch: the mailbox holds another dtype");
        xyz2 = mailbox->rows();
        n2 = mailbox->capacity;
        n2_dev = mailbox->count_dev();
    }
    if (!out || !xyz || !grids || n_lat < 1 || n_lat > LATTICE_BATCH) return fail(NBR_ERR_INVALID, "lattices_create_batch: bad argument");
    if (dtype != NBR_F32 && dtype != NBR_F64) return fail(NBR_ERR_INVALID, "lattices_create_batch: bad dtype");
    if (n < 1) return fail(NBR_ERR_TOO_FEW_POINTS, "lattices_create_batch: empty search cloud");
    if (n + n2 >= (int64_t)1 << 31) return fail(NBR_ERR_UNSUPPORTED, "lattices_create_batch: more than 2^31 search points");
    if (n2 < 0 || (n2 > 0 && !xyz2)) return fail(NBR_ERR_INVALID, "lattices_create_batch: bad second point set");
    std::vector<std::unique_ptr<Lattice>> lat(n_lat);
    BatchDev B;
    memset(&B, 0, sizeof(B));
    B.n = n_lat;
    int64_t dir_total = 0, slot_bound = 1;
    for (int l = 0; l < n_lat; ++l) {
        if (grids[l].ndim != 3) return fail(NBR_ERR_INVALID, "lattices_create_batch: 3-D grids only");
        Lattice *L = new Lattice();
        lat[l].reset(L);
        L->stream = stream;
        L->grid = grids[l];
        L->n_search = n + n2;
        NBR_TRY(grid_to_dev(&grids[l], &L->gdev, local_lohi));
        L->nbx = (L->gdev.ncell[0] + BRICK_X - 1) / BRICK_X;
        L->nby = (L->gdev.ncell[1] + BRICK_Y - 1) / BRICK_Y;
        L->nbz = (L->gdev.ncell[2] + BRICK_Z - 1) / BRICK_Z;
        const double dir_entries = (double)L->nbx * L->nby * L->nbz;
        if (dir_entries > 3.0e9) return fail(NBR_ERR_UNSUPPORTED, "brick directory would exceed 12 GB; extent / edge too large");
        L->n_dir = (int64_t)L->nbx * L->nby * L->nbz;
        L->pool_slots = std::min<int64_t>(n + n2, L->n_dir);
        B.g[l] = L->gdev;
        B.dir_off[l] = dir_total;
        B.nbx[l] = L->nbx;
        B.nby[l] = L->nby;
        dir_total += L->n_dir;
        slot_bound += L->pool_slots;
    }
    if (slot_bound >= (int64_t)1 << 32 || dir_total > (int64_t)6e9)
        return fail(NBR_ERR_UNSUPPORTED, "lattices_create_batch: the batch needs more than 2^32 bricks");
    auto shared = std::make_shared<Lattice::SharedBuffers>();
    shared->stream = stream;
    NBR_CUDA(pool_alloc((void **)&shared->dir, sizeof(uint32_t) * dir_total, stream));
    NBR_CUDA(pool_alloc((void **)&shared->pool, sizeof(uint32_t) * BRICK_WORDS * slot_bound, stream));
    NBR_CUDA(pool_alloc((void **)&shared->counters, 64 * (n_lat + 1), stream));
    NBR_CUDA(cudaMemsetAsync(shared->dir, 0, sizeof(uint32_t) * dir_total, stream));
    NBR_CUDA(cudaMemsetAsync(shared->counters, 0, 64 * (n_lat + 1), stream));
    uint32_t *dir = reinterpret_cast<uint32_t *>(shared->dir);
    uint32_t *pool = reinterpret_cast<uint32_t *>(shared->pool);
    unsigned char *counters = reinterpret_cast<unsigned char *>(shared->counters);
    uint32_t *n_bricks_total = reinterpret_cast<uint32_t *>(counters + 64 * n_lat);

    const void *parts[2] = {xyz, xyz2};
    const int64_t part_n[2] = {n, n2};
    const unsigned long long *part_dev[2] = {nullptr, n2_dev};
    // ordered cloud whose order cells are the bricks of the finest lattice of this batch: mark per occupied cell
    int by_cells = -1;
    static const bool no_cells = getenv("NBR_INDEX") && std::string(getenv("NBR_INDEX")) == "points";
    if (!no_cells && order && order->valid && order->n_points == n && order->n_cells > 0)
        for (int l = 0; l < n_lat; ++l)
            if (grids[l].edge == order->finest && lat[l]->nbx == order->dims[0] && lat[l]->nby == order->dims[1] &&
                lat[l]->nbz == order->dims[2])
                by_cells = l;
    if (by_cells >= 0)
        for (int l = 0; l < n_lat; ++l)
            if (grids[l].edge < order->finest) by_cells = -1;            // a finer lattice in the batch: its bricks are smaller than the cells
    for (int p = 0; p < 2; ++p) {
        // the tile's own bricks are marked while the peers' pushes are still in flight
        if (p == 1 && mailbox) NBR_TRY(halo_wait(mailbox, stream));
        if (part_n[p] <= 0) continue;
        if (p == 0 && by_cells >= 0) {
            CellsDev C;
            memset(&C, 0, sizeof(C));
            for (int a = 0; a < 3; ++a) { C.bdims[a] = order->bdims[a]; C.dims[a] = order->dims[a]; }
            C.finest = by_cells;
            const GridDev &f = B.g[by_cells];
            const int span[3] = {BRICK_X, BRICK_Y, BRICK_Z};
            for (int l = 0; l < n_lat; ++l)
                for (int a = 0; a < 3; ++a) {
                    C.alpha[l][a] = (f.minc[a] + (double)f.cell_lo[a] * f.edge - B.g[l].minc[a]) * B.g[l].inv_edge;
                    C.beta[l][a] = span[a] * f.edge * B.g[l].inv_edge;
                }
            const int64_t n_groups = ceil_div(order->n_cells, 512);
            cells_mark_kernel<<<(unsigned)ceil_div(n_groups, CM_WARPS), CM_WARPS * 32, 0, stream>>>(order->offsets.as<uint32_t>(), order->n_cells, n,
                                                                                                n_groups, B, C, dir);
            NBR_LAUNCHED();
            continue;
        }
        // a part whose size is only known on the device gets a capped grid that strides over its chunks
        const unsigned pt_blocks = (unsigned)std::min<int64_t>(ceil_div(part_n[p], 256 * PTS), part_dev[p] ? device_sm_count() * 8 : INT64_MAX);
        if (part_dev[p]) {
            if (dtype == NBR_F32) batch_mark_kernel<float, true><<<pt_blocks, 256, 0, stream>>>((const float *)parts[p], part_n[p], part_dev[p], B, dir);
            else                  batch_mark_kernel<double, true><<<pt_blocks, 256, 0, stream>>>((const double *)parts[p], part_n[p], part_dev[p], B, dir);
        } else {
            if (dtype == NBR_F32) batch_mark_kernel<float, false><<<pt_blocks, 256, 0, stream>>>((const float *)parts[p], part_n[p], nullptr, B, dir);
            else                  batch_mark_kernel<double, false><<<pt_blocks, 256, 0, stream>>>((const double *)parts[p], part_n[p], nullptr, B, dir);
        }
        NBR_LAUNCHED();
    }
    NBR_TRY(flags_to_slots(dir, dir_total, n_bricks_total, stream));
    pool_zero_kernel<<<device_sm_count() * 8, 256, 0, stream>>>(pool, n_bricks_total);
    NBR_LAUNCHED();
    for (int p = 0; p < 2; ++p) {
        if (part_n[p] <= 0) continue;
        const unsigned pt_blocks = (unsigned)std::min<int64_t>(ceil_div(part_n[p], 256 * PTS), part_dev[p] ? device_sm_count() * 8 : INT64_MAX);
        if (part_dev[p]) {
            if (dtype == NBR_F32) batch_fill_kernel<float, true><<<pt_blocks, 256, 0, stream>>>((const float *)parts[p], part_n[p], part_dev[p], B, dir, pool, counters);
            else                  batch_fill_kernel<double, true><<<pt_blocks, 256, 0, stream>>>((const double *)parts[p], part_n[p], part_dev[p], B, dir, pool, counters);
        } else {
            if (dtype == NBR_F32) batch_fill_kernel<float, false><<<pt_blocks, 256, 0, stream>>>((const float *)parts[p], part_n[p], nullptr, B, dir, pool, counters);
            else                  batch_fill_kernel<double, false><<<pt_blocks, 256, 0, stream>>>((const double *)parts[p], part_n[p], nullptr, B, dir, pool, counters);
        }
        NBR_LAUNCHED();
    }
    for (int l = 0; l < n_lat; ++l) {
        Lattice *L = lat[l].get();
        L->shared = shared;
        L->dir = dir + B.dir_off[l];
        L->pool = pool;                         // slot ids are global
        L->counters = counters + 64 * l;        // n_voxels; the brick count is only known for the whole batch
        out[l] = lat[l].release();
    }
    return NBR_OK;
}

// centres of the nv unique voxels of an INDEXED lattice, np.unique order (utils/geometry.py:120-138)
int lattice_centres(const Lattice *L, int64_t nv, double *centres, cudaStream_t stream)
{
    if (!L->indexed) return fail(NBR_ERR_INVALID, "lattice_centres: lattice was built without NBR_LATTICE_INDEXED");
    if (nv <= 0) return NBR_OK;
    centre_kernel<<<(unsigned)ceil_div(nv, 256), 256, 0, stream>>>(reinterpret_cast<const int64_t *>(L->ukeys), nv, L->gdev, centres);
    NBR_LAUNCHED();
    return NBR_OK;
}

int lattice_counts(const Lattice *L, int64_t *n_voxels, int64_t *n_bricks)
{
    unsigned char host[64];
    NBR_CUDA(cudaMemcpyAsync(host, L->counters, 64, cudaMemcpyDeviceToHost, L->stream));
    NBR_CUDA(cudaStreamSynchronize(L->stream));
    if (n_bricks) *n_bricks = L->shared ? -1 : (int64_t) * reinterpret_cast<uint32_t *>(host);   // batch: not tracked per lattice
    if (n_voxels) *n_voxels = (int64_t) * reinterpret_cast<unsigned long long *>(host + 8);
    return NBR_OK;
}

}  // namespace nbr

// ------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------
using namespace nbr;

extern "C" int nbr_bbox(const void *xyz, int dtype, int64_t n, int ndim, double *lohi_dev, void *stream)
{
    if (!xyz || !lohi_dev || (ndim != 2 && ndim != 3)) return fail(NBR_ERR_INVALID, "nbr_bbox: bad argument");
    return bbox(xyz, dtype, n, ndim, lohi_dev, (cudaStream_t)stream);
}

extern "C" int nbr_grid_from_bbox(const double lo[3], const double hi[3], double edge, int ndim, nbr_grid *out)
{
    if (!lo || !hi || !out) return fail(NBR_ERR_INVALID, "nbr_grid_from_bbox: null argument");
    return grid_from_bbox(lo, hi, edge, ndim, out);
}

extern "C" int nbr_voxel_addresses(const void *xyz, int dtype, int64_t n, const nbr_grid *grid, int64_t *addresses,
                                   int32_t *oob_dev, void *stream)
{
    if (!xyz || !grid || !addresses) return fail(NBR_ERR_INVALID, "nbr_voxel_addresses: null argument");
    if (n <= 0) return NBR_OK;
    GridDev g;
    NBR_TRY(grid_to_dev(grid, &g, nullptr));
    if (oob_dev) NBR_CUDA(cudaMemsetAsync(oob_dev, 0, sizeof(int32_t), (cudaStream_t)stream));
    address_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(
        xyz, dtype, n, g, grid->max_corner[0], grid->max_corner[1], grid->max_corner[2], addresses, oob_dev);
    NBR_LAUNCHED();
    return NBR_OK;
}

extern "C" int nbr_unique_u64(const uint64_t *sorted, int64_t n, uint64_t *out, int64_t *n_out_dev, void *stream)
{
    if (!sorted || !out || !n_out_dev) return fail(NBR_ERR_INVALID, "nbr_unique_u64: null argument");
    return unique_sorted(sorted, n, out, n_out_dev, (cudaStream_t)stream);
}

extern "C" int nbr_voxel_centres(const int64_t *addresses, int64_t n, const nbr_grid *grid, double *xyz_out, void *stream)
{
    if (!addresses || !grid || !xyz_out) return fail(NBR_ERR_INVALID, "nbr_voxel_centres: null argument");
    if (n <= 0) return NBR_OK;
    GridDev g;
    NBR_TRY(grid_to_dev(grid, &g, nullptr));
    centre_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(addresses, n, g, xyz_out);
    NBR_LAUNCHED();
    return NBR_OK;
}

extern "C" int nbr_lattice_create(nbr_lattice **out, const void *search_xyz, int dtype, int64_t n_search,
                                  const nbr_grid *grid, int flags, void *stream)
{
    return lattice_create(reinterpret_cast<Lattice **>(out), search_xyz, dtype, n_search, grid, flags,
                          (cudaStream_t)stream);
}

extern "C" void nbr_lattice_destroy(nbr_lattice *lattice) { delete reinterpret_cast<Lattice *>(lattice); }

extern "C" int nbr_lattice_info(const nbr_lattice *lattice, int64_t *n_voxels, int64_t *n_bricks)
{
    if (!lattice) return fail(NBR_ERR_INVALID, "nbr_lattice_info: null lattice");
    return lattice_counts(reinterpret_cast<const Lattice *>(lattice), n_voxels, n_bricks);
}

extern "C" int nbr_lattice_export(const nbr_lattice *lattice, int64_t *addresses, double *centres, void *stream)
{
    const Lattice *L = reinterpret_cast<const Lattice *>(lattice);
    if (!L) return fail(NBR_ERR_INVALID, "nbr_lattice_export: null lattice");
    if (!L->indexed) return fail(NBR_ERR_INVALID, "nbr_lattice_export: lattice was built without NBR_LATTICE_INDEXED");
    int64_t nv = 0;
    NBR_TRY(lattice_counts(L, &nv, nullptr));
    cudaStream_t s = (cudaStream_t)stream;
    if (addresses) NBR_CUDA(cudaMemcpyAsync(addresses, L->ukeys, sizeof(int64_t) * nv, cudaMemcpyDeviceToDevice, s));
    if (centres && nv) {
        centre_kernel<<<(unsigned)ceil_div(nv, 256), 256, 0, s>>>(reinterpret_cast<const int64_t *>(L->ukeys), nv,
                                                                L->gdev, centres);
        NBR_LAUNCHED();
    }
    return NBR_OK;
}
