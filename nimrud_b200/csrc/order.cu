// order.cu -- spatially coherent processing order of the query cloud: Morton (Z-curve) keys at a
// granularity of a few finest voxels, sorted with the hand-written radix sort.  consecutive queries
// of the order are close in space at every scale, which is what lets one warp of the feature kernel
// share a staged occupancy window.  results never depend on the order (each query is independent).
#include "common.cuh"
#include "scan.cuh"

namespace nbr {

int sort_pairs(uint64_t *keys, uint64_t *keys_tmp, uint32_t *vals, uint32_t *vals_tmp, int64_t n, int begin_bit,
               int end_bit, cudaStream_t stream);

struct MortonParam {
    double lo[3];
    double inv_cell;
    int bits[3];
    int levels;
    int common;   // levels present on all three axes
    // position of bit l of axis a in the key (axes with fewer bits drop out of the upper levels)
    unsigned char pos[3][21];
};

// spread the low 21 bits of x so that bit i lands at 3*i
__device__ __forceinline__ uint64_t spread3(uint32_t v)
{
    uint64_t x = v & 0x1fffffu;
    x = (x | x << 32) & 0x1f00000000ffffull;
    x = (x | x << 16) & 0x1f0000ff0000ffull;
    x = (x | x << 8) & 0x100f00f00f00f00full;
    x = (x | x << 4) & 0x10c30c30c30c30c3ull;
    x = (x | x << 2) & 0x1249249249249249ull;
    return x;
}

__global__ void __launch_bounds__(256)
morton_kernel(const void *__restrict__ xyz, int dtype, int64_t n, MortonParam P, uint64_t *__restrict__ keys,
              uint32_t *__restrict__ vals)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    // the order only has to be spatially coherent, so float32 quantisation is fine here
    uint32_t c[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        float v = (float)(load_coord(xyz, dtype, i, 3, a) - P.lo[a]) * (float)P.inv_cell;
        const float top = (float)((1u << P.bits[a]) - 1u);
        v = fminf(fmaxf(v, 0.0f), top);
        c[a] = (uint32_t)v;
    }
    // levels shared by all three axes: classic 3-way bit interleave; the rest (axes with more bits) generically
    const uint32_t low = (1u << P.common) - 1u;
    uint64_t key = spread3(c[0] & low) | (spread3(c[1] & low) << 1) | (spread3(c[2] & low) << 2);
#pragma unroll
    for (int a = 0; a < 3; ++a)
        for (int l = P.common; l < P.bits[a]; ++l) key |= (uint64_t)((c[a] >> l) & 1u) << P.pos[a][l];
    keys[i] = key;
    vals[i] = (uint32_t)i;
}

// sorted_out[i] = xyz[perm[i]] (same dtype as the input)
__global__ void __launch_bounds__(256)
gather_kernel(const void *__restrict__ xyz, int dtype, int64_t n, const uint32_t *__restrict__ perm,
              void *__restrict__ sorted_out)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int64_t j = perm[i];
    if (dtype == NBR_F32) {
        const float *src = reinterpret_cast<const float *>(xyz) + j * 3;
        float *dst = reinterpret_cast<float *>(sorted_out) + i * 3;
        dst[0] = src[0]; dst[1] = src[1]; dst[2] = src[2];
    } else {
        const double *src = reinterpret_cast<const double *>(xyz) + j * 3;
        double *dst = reinterpret_cast<double *>(sorted_out) + i * 3;
        dst[0] = src[0]; dst[1] = src[1]; dst[2] = src[2];
    }
}

// perm_out[n]: query indices in Morton order; sorted_xyz_out (optional): the cloud in that order.
// lohi = bounding box of the query cloud (host).
int morton_order(const void *xyz, int dtype, int64_t n, const double lohi[6], double cell, uint32_t *perm_out,
                 void *sorted_xyz_out, cudaStream_t stream)
{
    if (n <= 0) return NBR_OK;
    if (n >= (int64_t)1 << 32) return fail(NBR_ERR_UNSUPPORTED, "morton_order: more than 2^32 queries");
    MortonParam P;
    P.inv_cell = 1.0 / cell;
    P.levels = 0;
    int total = 0;
    for (int a = 0; a < 3; ++a) {
        P.lo[a] = lohi[a];
        double cells = (lohi[3 + a] - lohi[a]) / cell + 1.0;
        int b = 1;
        while (b < 21 && (double)(1u << b) < cells) ++b;
        P.bits[a] = b;
        P.levels = std::max(P.levels, b);
        total += b;
    }
    P.common = std::min(P.bits[0], std::min(P.bits[1], P.bits[2]));
    {
        int pos = 0;
        for (int l = 0; l < P.levels; ++l)
            for (int a = 0; a < 3; ++a)
                if (l < P.bits[a]) P.pos[a][l] = (unsigned char)pos++;
    }
    Scratch keys, keys_tmp, vals_tmp;
    NBR_TRY(keys.alloc(sizeof(uint64_t) * n, stream));
    NBR_TRY(keys_tmp.alloc(sizeof(uint64_t) * n, stream));
    NBR_TRY(vals_tmp.alloc(sizeof(uint32_t) * n, stream));
    morton_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, stream>>>(xyz, dtype, n, P, keys.as<uint64_t>(), perm_out);
    NBR_LAUNCHED();
    NBR_TRY(sort_pairs(keys.as<uint64_t>(), keys_tmp.as<uint64_t>(), perm_out, vals_tmp.as<uint32_t>(), n, 0, total, stream));
    if (sorted_xyz_out) {
        gather_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, stream>>>(xyz, dtype, n, perm_out, sorted_xyz_out);
        NBR_LAUNCHED();
    }
    return NBR_OK;
}

// ------------------------------------------------------------------------------------------------
// cell order: counting sort of the cloud on a dense grid of brick-shaped cells (32 x 8 x 4 finest voxels,
// aligned with the bricks of the finest lattice).  every point of a cell needs the same <= 3 x 3 x 3 bricks
// of that lattice (fewer on coarser ones), so a warp of 32 consecutive points always finds its window in
// the staging buffer of the feature kernels.  two passes over the points (count + place) and one scan of
// the cell counters instead of a 4-pass radix sort.  the order inside a cell is whatever the atomics
// produce; results do not depend on it (every query is independent).
// ------------------------------------------------------------------------------------------------
struct CellGrid {
    double origin[3];
    double inv_cell[3];
    int first[3];
    int dims[3];
    int bdims[3];                 // blocks of 8 x 8 x 8 cells
    int exact;                    // cells = bricks of `g` (the finest lattice), from its own exact cell arithmetic
    GridDev g;
};

// cells are numbered block by block (blocks of 8 x 8 x 8 cells, row-major) and along a Z-curve inside a block:
// cells that follow each other in the order are neighbors in space even where most cells are empty, and the
// dense counter array is at most a few percent larger than the grid (a Z-curve over the whole grid pads every
// axis to a power of two: 3.2x on the config-2 scene)
__device__ __forceinline__ uint32_t cell_of(const void *xyz, int dtype, int64_t i, const CellGrid &G)
{
    uint32_t c[3];
    if (G.exact) {
        // the brick of the finest lattice that holds the point, bit-identical to what the lattice build computes
        int k[3];
        if (dtype == NBR_F32) point_cell<float>(reinterpret_cast<const float *>(xyz), i, G.g, k);
        else point_cell<double>(reinterpret_cast<const double *>(xyz), i, G.g, k);
        c[0] = (uint32_t)(k[0] >> BRICK_XS); c[1] = (uint32_t)(k[1] >> BRICK_YS); c[2] = (uint32_t)(k[2] >> BRICK_ZS);
    } else {
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const double u = (load_coord(xyz, dtype, i, 3, a) - G.origin[a]) * G.inv_cell[a];
            c[a] = (uint32_t)clampi((int)floor(u) - G.first[a], 0, G.dims[a] - 1);
        }
    }
    const uint32_t block = ((c[2] >> 3) * (uint32_t)G.bdims[1] + (c[1] >> 3)) * (uint32_t)G.bdims[0] + (c[0] >> 3);
    const uint32_t z9 = (uint32_t)(spread3(c[0] & 7u) | (spread3(c[1] & 7u) << 1) | (spread3(c[2] & 7u) << 2));
    return block * 512u + z9;
}

__global__ void __launch_bounds__(256)
cell_count_kernel(const void *__restrict__ xyz, int dtype, int64_t n, CellGrid G, uint32_t *__restrict__ counts,
                  uint32_t *__restrict__ cell, uint32_t *__restrict__ rank)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t c = cell_of(xyz, dtype, i, G);
    cell[i] = c;
    rank[i] = atomicAdd(&counts[c], 1u);
}

__global__ void __launch_bounds__(256)
cell_place_kernel(int64_t n, const uint32_t *__restrict__ offsets, const uint32_t *__restrict__ cell,
                  const uint32_t *__restrict__ rank, uint32_t *__restrict__ perm)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    perm[offsets[cell[i]] + rank[i]] = (uint32_t)i;
}

// lohi: bounding box of the cloud (host).  origin / cell: corner and edge lengths of the cell grid to align
// with (cells are doubled until the dense counter array fits 2^26 entries).
int cell_order(const void *xyz, int dtype, int64_t n, const double lohi[6], const double origin[3],
               const double cell_in[3], uint32_t *perm_out, void *sorted_xyz_out, cudaStream_t stream,
               const GridDev *finest, CellOrderInfo *info)
{
    if (info) info->valid = false;
    if (n <= 0) return NBR_OK;
    if (n >= (int64_t)1 << 32) return fail(NBR_ERR_UNSUPPORTED, "cell_order: more than 2^32 points");
    CellGrid G;
    memset(&G, 0, sizeof(G));
    double cell[3] = {cell_in[0], cell_in[1], cell_in[2]};
    double total = -1;
    if (finest) {
        // cells = the bricks of the finest lattice over its whole directory (every point of the cloud lies inside)
        double cells = 512.0;
        for (int a = 0; a < 3; ++a) {
            G.origin[a] = origin[a];
            G.inv_cell[a] = 1.0 / cell[a];
            G.first[a] = 0;
        }
        G.dims[0] = (finest->ncell[0] + BRICK_X - 1) / BRICK_X;
        G.dims[1] = (finest->ncell[1] + BRICK_Y - 1) / BRICK_Y;
        G.dims[2] = (finest->ncell[2] + BRICK_Z - 1) / BRICK_Z;
        for (int a = 0; a < 3; ++a) { G.bdims[a] = (G.dims[a] + 7) / 8; cells *= (double)G.bdims[a]; }
        if (cells <= 67108864.0) { total = cells; G.exact = 1; G.g = *finest; }
    }
    for (int attempt = 0; total < 0 && attempt < 64; ++attempt) {
        double cells = 512.0;
        bool ok = true;
        for (int a = 0; a < 3; ++a) {
            G.origin[a] = origin[a];
            G.inv_cell[a] = 1.0 / cell[a];
            const double f = floor((lohi[a] - origin[a]) / cell[a]), l = floor((lohi[3 + a] - origin[a]) / cell[a]);
            if (!(fabs(f) < 2.0e9 && fabs(l) < 2.0e9 && l - f + 1.0 <= 1048576.0)) { ok = false; break; }
            G.first[a] = (int)f;
            G.dims[a] = (int)(l - f + 1.0);
            G.bdims[a] = (G.dims[a] + 7) / 8;
            cells *= (double)G.bdims[a];
        }
        if (ok && cells <= 67108864.0) { total = cells; break; }
        for (int a = 0; a < 3; ++a) cell[a] *= 2.0;
    }
    if (total < 0) return fail(NBR_ERR_UNSUPPORTED, "cell_order: cannot cover the cloud with a cell grid");
    const int64_t nc = (int64_t)total;
    Scratch counts, cellid, rank;
    NBR_TRY(counts.alloc(sizeof(uint32_t) * nc, stream));
    NBR_TRY(cellid.alloc(sizeof(uint32_t) * n, stream));
    NBR_TRY(rank.alloc(sizeof(uint32_t) * n, stream));
    NBR_CUDA(cudaMemsetAsync(counts.ptr, 0, sizeof(uint32_t) * nc, stream));
    const unsigned blocks = (unsigned)ceil_div(n, 256);
    cell_count_kernel<<<blocks, 256, 0, stream>>>(xyz, dtype, n, G, counts.as<uint32_t>(), cellid.as<uint32_t>(),
                                                  rank.as<uint32_t>());
    NBR_LAUNCHED();
    NBR_TRY((exclusive_scan<uint32_t, uint32_t>(counts.as<uint32_t>(), counts.as<uint32_t>(), nc, stream)));
    cell_place_kernel<<<blocks, 256, 0, stream>>>(n, counts.as<uint32_t>(), cellid.as<uint32_t>(), rank.as<uint32_t>(),
                                                  perm_out);
    NBR_LAUNCHED();
    if (sorted_xyz_out) {
        gather_kernel<<<blocks, 256, 0, stream>>>(xyz, dtype, n, perm_out, sorted_xyz_out);
        NBR_LAUNCHED();
    }
    if (info && G.exact && sorted_xyz_out) {
        info->offsets.swap(counts);
        info->n_cells = nc;
        info->n_points = n;
        for (int a = 0; a < 3; ++a) { info->bdims[a] = G.bdims[a]; info->dims[a] = G.dims[a]; }
        info->finest = finest->edge;
        info->valid = true;
    }
    return NBR_OK;
}

// ------------------------------------------------------------------------------------------------
// queries = the first nq points of an ordered cloud of n points (tile + halo: the tile's points are the
// queries, multi-GPU path).  keeps the order, drops the others: perm_q / sorted_q hold the nq entries whose
// original index is < nq.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
prefix_flag_kernel(const uint32_t *__restrict__ perm, int64_t n, uint32_t nq, uint32_t *__restrict__ flags)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) flags[i] = perm[i] < nq ? 1u : 0u;
}

__global__ void __launch_bounds__(256)
prefix_compact_kernel(const uint32_t *__restrict__ perm, const void *__restrict__ sorted, int dtype, int64_t n,
                      const uint32_t *__restrict__ slots, uint32_t *__restrict__ perm_q, void *__restrict__ sorted_q)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t s = slots[i];
    if (!s) return;
    const int64_t pos = (int64_t)s - 1;
    perm_q[pos] = perm[i];
    if (dtype == NBR_F32) {
        const float *src = reinterpret_cast<const float *>(sorted) + i * 3;
        float *dst = reinterpret_cast<float *>(sorted_q) + pos * 3;
        dst[0] = src[0]; dst[1] = src[1]; dst[2] = src[2];
    } else {
        const double *src = reinterpret_cast<const double *>(sorted) + i * 3;
        double *dst = reinterpret_cast<double *>(sorted_q) + pos * 3;
        dst[0] = src[0]; dst[1] = src[1]; dst[2] = src[2];
    }
}

int compact_prefix_queries(const uint32_t *perm, const void *sorted, int dtype, int64_t n, int64_t nq, uint32_t *perm_q,
                           void *sorted_q, cudaStream_t stream)
{
    if (n <= 0 || nq <= 0) return NBR_OK;
    Scratch flags, count;
    NBR_TRY(flags.alloc(sizeof(uint32_t) * n, stream));
    NBR_TRY(count.alloc(sizeof(uint32_t), stream));
    const unsigned blocks = (unsigned)ceil_div(n, 256);
    prefix_flag_kernel<<<blocks, 256, 0, stream>>>(perm, n, (uint32_t)nq, flags.as<uint32_t>());
    NBR_LAUNCHED();
    NBR_TRY(flags_to_slots(flags.as<uint32_t>(), n, count.as<uint32_t>(), stream));
    prefix_compact_kernel<<<blocks, 256, 0, stream>>>(perm, sorted, dtype, n, flags.as<uint32_t>(), perm_q, sorted_q);
    NBR_LAUNCHED();
    return NBR_OK;
}

}  // namespace nbr
