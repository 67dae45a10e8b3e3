// vector_field.cu -- the vector-field multiscale operator (legacy precedent: V_MSO,
// nimrud/prototypes/mso.py:12-175 and vec_field_interp :178-257): every search point carries a feature
// vector; the vectors are averaged per voxel of the lattice ("interpolated to the voxel grid"), and for every
// query and radius the mean of the voxel vectors over the voxels within the radius is returned.
//
// membership is the same exact, inclusive float64 test as the eigenfeature path (the minimal/ semantics, not
// the prototype's strict float32 one); no reference code runs this on the minimal/ path: parity unpinned.
// no neighbor list is materialised: the visitor of radius_exact.cu accumulates the voxel vectors directly.
#include "common.cuh"
#include "finalize.cuh"
#include "lattice.cuh"
#include "members.cuh"

namespace nbr {

constexpr int VF_CHUNK = 8;      // vector components accumulated per pass over a query's neighbors

// rank (np.unique order) of the voxel holding point i, or -1
__device__ __forceinline__ int64_t voxel_rank_of_point(const LatticeDev &L, const void *xyz, int dtype, int64_t i)
{
    const GridDev &g = L.g;
    int c[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const double p = load_coord(xyz, dtype, i, 3, a);
        const double k = cell_coord_fast(p, g.minc[a], g.edge, g.inv_edge) - (double)g.cell_lo[a];
        if (!(k >= 0.0 && k <= (double)(g.ncell[a] - 1))) return -1;
        c[a] = (int)k;
    }
    const uint32_t slot = L.dir[((int64_t)(c[2] >> BRICK_ZS) * L.nby + (c[1] >> BRICK_YS)) * L.nbx + (c[0] >> BRICK_XS)];
    if (!slot) return -1;
    const int64_t w = (int64_t)slot * BRICK_WORDS + (((c[2] & (BRICK_Z - 1)) << BRICK_YS) | (c[1] & (BRICK_Y - 1)));
    const uint32_t bits = L.pool[w];
    const int b = c[0] & 31;
    if (!((bits >> b) & 1u)) return -1;
    return (int64_t)L.rowbase[w] + __popc(bits & ((1u << b) - 1u));
}

__global__ void __launch_bounds__(256)
voxel_vector_sum_kernel(LatticeDev L, const void *__restrict__ xyz, int dtype, int64_t n, const float *__restrict__ vec,
                        int F, double *__restrict__ sums, unsigned int *__restrict__ counts)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int64_t r = voxel_rank_of_point(L, xyz, dtype, i);
    if (r < 0) return;
    atomicAdd(counts + r, 1u);
    for (int f = 0; f < F; ++f) atomicAdd(sums + r * F + f, (double)vec[i * F + f]);
}

__global__ void __launch_bounds__(256)
voxel_vector_mean_kernel(const double *__restrict__ sums, const unsigned int *__restrict__ counts, int64_t nv, int F,
                         float *__restrict__ out)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nv * F) return;
    const unsigned int c = counts[i / F];
    out[i] = c ? (float)(sums[i] / (double)c) : 0.0f;
}

template <typename OutT>
__global__ void __launch_bounds__(128)
radius_vector_mean_kernel(LatticeDev L, const void *__restrict__ query, int dtype, int64_t nq, double radius,
                          const float *__restrict__ voxvec, int F, OutT *__restrict__ out, int64_t row_stride,
                          int col_offset)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nq) return;
    double q[3], fr;
    int c[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        q[a] = load_coord(query, dtype, i, 3, a);
        query_anchor(q[a], L.g, a, c[a], fr);
    }
    OutT *dst = out + i * row_stride + col_offset;
    for (int f0 = 0; f0 < F; f0 += VF_CHUNK) {
        double acc[VF_CHUNK];
#pragma unroll
        for (int t = 0; t < VF_CHUNK; ++t) acc[t] = 0.0;
        long long n = 0;
        for_each_member(L, q, c, radius, [&](int, int, int, uint32_t slot, int word, int b) {
            const int64_t w = (int64_t)slot * BRICK_WORDS + word;
            const int64_t r = (int64_t)L.rowbase[w] + __popc(L.pool[w] & ((1u << b) - 1u));
            const float *v = voxvec + r * F + f0;
#pragma unroll
            for (int t = 0; t < VF_CHUNK; ++t)
                if (f0 + t < F) acc[t] += (double)v[t];
            ++n;
        });
#pragma unroll
        for (int t = 0; t < VF_CHUNK; ++t)
            if (f0 + t < F) dst[f0 + t] = n ? (OutT)(acc[t] / (double)n) : (OutT)0;      // undefined = 0
    }
}

}  // namespace nbr

using namespace nbr;

extern "C" int nbr_voxel_vector_means(const nbr_lattice *lattice, const void *search_xyz, int dtype, int64_t n_search,
                                      const float *vectors, int32_t n_components, float *voxvec_out, void *stream)
{
    const Lattice *L = reinterpret_cast<const Lattice *>(lattice);
    if (!L || !search_xyz || !vectors || !voxvec_out) return fail(NBR_ERR_INVALID, "nbr_voxel_vector_means: null argument");
    if (!L->indexed) return fail(NBR_ERR_INVALID, "nbr_voxel_vector_means: lattice was built without NBR_LATTICE_INDEXED");
    if (dtype != NBR_F32 && dtype != NBR_F64) return fail(NBR_ERR_INVALID, "nbr_voxel_vector_means: bad dtype");
    if (n_components < 1) return fail(NBR_ERR_INVALID, "nbr_voxel_vector_means: need at least one component");
    int64_t nv = 0;
    NBR_TRY(lattice_counts(L, &nv, nullptr));
    if (nv <= 0 || n_search <= 0) return NBR_OK;
    cudaStream_t s = (cudaStream_t)stream;
    Scratch sums, counts;
    NBR_TRY(sums.alloc(sizeof(double) * nv * n_components, s));
    NBR_TRY(counts.alloc(sizeof(unsigned int) * nv, s));
    NBR_CUDA(cudaMemsetAsync(sums.ptr, 0, sizeof(double) * nv * n_components, s));
    NBR_CUDA(cudaMemsetAsync(counts.ptr, 0, sizeof(unsigned int) * nv, s));
    voxel_vector_sum_kernel<<<(unsigned)ceil_div(n_search, 256), 256, 0, s>>>(L->dev(), search_xyz, dtype, n_search, vectors,
                                                                              n_components, sums.as<double>(), counts.as<unsigned int>());
    NBR_LAUNCHED();
    voxel_vector_mean_kernel<<<(unsigned)ceil_div(nv * n_components, 256), 256, 0, s>>>(sums.as<double>(), counts.as<unsigned int>(),
                                                                                       nv, n_components, voxvec_out);
    NBR_LAUNCHED();
    return NBR_OK;
}

extern "C" int nbr_radius_vector_means(const nbr_lattice *lattice, const void *query_xyz, int dtype, int64_t n_query,
                                       double radius, const float *voxvec, int32_t n_components, void *out, int out_dtype,
                                       int64_t out_row_stride, int32_t col_offset, void *stream)
{
    const Lattice *L = reinterpret_cast<const Lattice *>(lattice);
    if (!L || !query_xyz || !voxvec || !out) return fail(NBR_ERR_INVALID, "nbr_radius_vector_means: null argument");
    if (!L->indexed) return fail(NBR_ERR_INVALID, "nbr_radius_vector_means: lattice was built without NBR_LATTICE_INDEXED");
    if (dtype != NBR_F32 && dtype != NBR_F64) return fail(NBR_ERR_INVALID, "nbr_radius_vector_means: bad dtype");
    if (out_dtype != NBR_F32 && out_dtype != NBR_F64) return fail(NBR_ERR_INVALID, "nbr_radius_vector_means: bad out_dtype");
    if (n_components < 1 || !(radius >= 0)) return fail(NBR_ERR_INVALID, "nbr_radius_vector_means: bad argument");
    if (n_query <= 0) return NBR_OK;
    cudaStream_t s = (cudaStream_t)stream;
    const unsigned blocks = (unsigned)ceil_div(n_query, 128);
    if (out_dtype == NBR_F32)
        radius_vector_mean_kernel<float><<<blocks, 128, 0, s>>>(L->dev(), query_xyz, dtype, n_query, radius, voxvec,
                                                                n_components, (float *)out, out_row_stride, col_offset);
    else
        radius_vector_mean_kernel<double><<<blocks, 128, 0, s>>>(L->dev(), query_xyz, dtype, n_query, radius, voxvec,
                                                                 n_components, (double *)out, out_row_stride, col_offset);
    NBR_LAUNCHED();
    return NBR_OK;
}
