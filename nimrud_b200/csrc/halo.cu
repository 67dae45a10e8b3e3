// halo.cu -- multi-GPU halo selection: which points of this rank's tile does each other tile need?
// (the data-parallel half of nimrud_b200/distributed.py; precedent in the reference: nested_regions,
// nimrud/utils/geometry.py:203-253 -- inclusive box +- buffer radius.)
//
// two passes over the tile for up to 8 destination boxes at a time: count (warp-aggregated atomics), then
// fill the per-destination segments of the send buffer.  the inclusive float64 test is the same code in
// both passes; the order inside a segment is arbitrary (the voxel lattices do not depend on it).
#include "common.cuh"

namespace nbr {

constexpr int HALO_MAX_DST = 8;

struct HaloBoxes {
    double lo[HALO_MAX_DST][3], hi[HALO_MAX_DST][3];
    long long offset[HALO_MAX_DST];      // first row of each destination's segment (fill pass)
    int n;
};

__device__ __forceinline__ uint32_t halo_mask(const void *xyz, int dtype, int64_t i, const HaloBoxes &B)
{
    const double x = load_coord(xyz, dtype, i, 3, 0), y = load_coord(xyz, dtype, i, 3, 1), z = load_coord(xyz, dtype, i, 3, 2);
    uint32_t m = 0;
    for (int d = 0; d < B.n; ++d)
        if (x >= B.lo[d][0] && x <= B.hi[d][0] && y >= B.lo[d][1] && y <= B.hi[d][1] && z >= B.lo[d][2] && z <= B.hi[d][2])
            m |= 1u << d;
    return m;
}

__global__ void __launch_bounds__(256)
halo_count_kernel(const void *__restrict__ xyz, int dtype, int64_t n, const __grid_constant__ HaloBoxes B,
                  unsigned long long *__restrict__ counts)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t m = i < n ? halo_mask(xyz, dtype, i, B) : 0u;
    for (int d = 0; d < B.n; ++d) {
        const uint32_t votes = __ballot_sync(0xffffffffu, (m >> d) & 1u);
        if (votes && (threadIdx.x & 31) == 0) atomicAdd(counts + d, (unsigned long long)__popc(votes));
    }
}

template <typename T>
__global__ void __launch_bounds__(256)
halo_fill_kernel(const T *__restrict__ xyz, int dtype, int64_t n, const __grid_constant__ HaloBoxes B,
                 unsigned long long *__restrict__ cursors, T *__restrict__ out)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t m = i < n ? halo_mask(xyz, dtype, i, B) : 0u;
    const uint32_t lt = lanemask_lt();
    for (int d = 0; d < B.n; ++d) {
        const bool mine = (m >> d) & 1u;
        const uint32_t votes = __ballot_sync(0xffffffffu, mine);
        if (!votes) continue;
        unsigned long long base = 0;
        const int leader = __ffs(votes) - 1;
        if ((int)(threadIdx.x & 31) == leader) base = atomicAdd(cursors + d, (unsigned long long)__popc(votes));
        base = __shfl_sync(0xffffffffu, base, leader);
        if (mine) {
            const long long row = B.offset[d] + (long long)base + __popc(votes & lt);
            out[row * 3 + 0] = xyz[i * 3 + 0];
            out[row * 3 + 1] = xyz[i * 3 + 1];
            out[row * 3 + 2] = xyz[i * 3 + 2];
        }
    }
}

static int fill_boxes(const double *boxes_host, int ndst, const int64_t *offsets_host, HaloBoxes *B)
{
    if (!boxes_host || ndst < 0 || ndst > HALO_MAX_DST) return fail(NBR_ERR_INVALID, "halo: between 0 and 8 destination boxes per call");
    memset(B, 0, sizeof(*B));
    B->n = ndst;
    for (int d = 0; d < ndst; ++d)
        for (int a = 0; a < 3; ++a) {
            B->lo[d][a] = boxes_host[d * 6 + a];
            B->hi[d][a] = boxes_host[d * 6 + 3 + a];
        }
    if (offsets_host)
        for (int d = 0; d < ndst; ++d) B->offset[d] = offsets_host[d];
    return NBR_OK;
}

}  // namespace nbr

using namespace nbr;

// counts_dev[ndst] (uint64, device) += number of points inside each inclusive box [lo, hi]
extern "C" int nbr_halo_count(const void *xyz, int dtype, int64_t n, const double *boxes_host, int32_t ndst,
                              uint64_t *counts_dev, void *stream)
{
    if (!xyz || !counts_dev) return fail(NBR_ERR_INVALID, "nbr_halo_count: null argument");
    if (dtype != NBR_F32 && dtype != NBR_F64) return fail(NBR_ERR_INVALID, "nbr_halo_count: bad dtype");
    HaloBoxes B;
    NBR_TRY(fill_boxes(boxes_host, ndst, nullptr, &B));
    if (n <= 0 || ndst == 0) return NBR_OK;
    halo_count_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(
        xyz, dtype, n, B, reinterpret_cast<unsigned long long *>(counts_dev));
    NBR_LAUNCHED();
    return NBR_OK;
}

// out (rows of 3, same dtype as xyz): the points inside box d go to rows [offsets_host[d], offsets_host[d] + count_d);
// cursors_dev[ndst] (uint64, device) must be zero on entry
extern "C" int nbr_halo_fill(const void *xyz, int dtype, int64_t n, const double *boxes_host, int32_t ndst,
                             const int64_t *offsets_host, uint64_t *cursors_dev, void *out, void *stream)
{
    if (!xyz || !cursors_dev || !out || !offsets_host) return fail(NBR_ERR_INVALID, "nbr_halo_fill: null argument");
    if (dtype != NBR_F32 && dtype != NBR_F64) return fail(NBR_ERR_INVALID, "nbr_halo_fill: bad dtype");
    HaloBoxes B;
    NBR_TRY(fill_boxes(boxes_host, ndst, offsets_host, &B));
    if (n <= 0 || ndst == 0) return NBR_OK;
    const unsigned blocks = (unsigned)ceil_div(n, 256);
    unsigned long long *cur = reinterpret_cast<unsigned long long *>(cursors_dev);
    if (dtype == NBR_F32)
        halo_fill_kernel<float><<<blocks, 256, 0, (cudaStream_t)stream>>>((const float *)xyz, dtype, n, B, cur, (float *)out);
    else
        halo_fill_kernel<double><<<blocks, 256, 0, (cudaStream_t)stream>>>((const double *)xyz, dtype, n, B, cur, (double *)out);
    NBR_LAUNCHED();
    return NBR_OK;
}
