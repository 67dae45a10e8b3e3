// radius_exact.cu -- per-candidate radius queries on the bit-brick lattice, every membership test in
// the reference's exact float64 arithmetic.  Two users:
//   * radius_features_exact_kernel: the general feature kernel (any r/e); also the in-library
//     cross-check of the row-interval kernel (radius_rows.cu).
//   * radius_sets_*: neighbor index sets in CSR form, the parity path against
//     cKDTree.query_ball_tree (nimrud/minimal/multiscale.py:103).
#include "common.cuh"
#include "finalize.cuh"
#include "lattice.cuh"
#include "scan.cuh"
#include "members.cuh"

namespace nbr {

struct RadiiParam {
    double r[16];
    int n;
};

template <typename OutT>
__global__ void __launch_bounds__(128)
radius_features_exact_kernel(LatticeDev L, const void *__restrict__ query, int dtype,
                             const uint32_t *__restrict__ perm, int64_t nq, RadiiParam rp, OutT *__restrict__ out,
                             int64_t row_stride, int col_offset, int descriptor_mask)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nq) return;
    const int ri = blockIdx.y;
    double q[3], f[3];
    int c[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        q[a] = load_coord(query, dtype, i, 3, a);
        query_anchor(q[a], L.g, a, c[a], f[a]);
    }
    Moments m;
    m.n = 0;
#pragma unroll
    for (int k = 0; k < 3; ++k) m.s1[k] = 0;
#pragma unroll
    for (int k = 0; k < 6; ++k) m.s2[k] = 0;
    for_each_member(L, q, c, rp.r[ri], [&](int jx, int jy, int jz, uint32_t, int, int) {
        m.n += 1;
        m.s1[0] += jx; m.s1[1] += jy; m.s1[2] += jz;
        m.s2[0] += (long long)jx * jx; m.s2[1] += (long long)jx * jy; m.s2[2] += (long long)jx * jz;
        m.s2[3] += (long long)jy * jy; m.s2[4] += (long long)jy * jz; m.s2[5] += (long long)jz * jz;
    });
    const int ncol = (descriptor_mask & NBR_DESC_EXTENDED) ? NBR_COLS_EXTENDED : NBR_COLS_REFERENCE;
    const int64_t row = perm ? (int64_t)perm[i] : i;       // queries may come in a sorted order
    emit_features<OutT>(m, f, L.g.edge, out + row * row_stride + col_offset + ri * ncol, descriptor_mask);
}

int radius_features_exact(const Lattice *lat, const void *query, int dtype, const uint32_t *perm, int64_t nq,
                          const double *radii, int nr, void *out, int out_dtype, int64_t row_stride, int col_offset,
                          int descriptor_mask, cudaStream_t stream)
{
    if (nq <= 0 || nr <= 0) return NBR_OK;
    for (int base = 0; base < nr; base += 16) {
        RadiiParam rp;
        rp.n = std::min(16, nr - base);
        for (int k = 0; k < rp.n; ++k) rp.r[k] = radii[base + k];
        const int ncol = (descriptor_mask & NBR_DESC_EXTENDED) ? NBR_COLS_EXTENDED : NBR_COLS_REFERENCE;
        dim3 grid((unsigned)ceil_div(nq, 128), rp.n);
        if (out_dtype == NBR_F32)
            radius_features_exact_kernel<float><<<grid, 128, 0, stream>>>(lat->dev(), query, dtype, perm, nq, rp, (float *)out,
                                                                           row_stride, col_offset + base * ncol,
                                                                           descriptor_mask);
        else
            radius_features_exact_kernel<double><<<grid, 128, 0, stream>>>(lat->dev(), query, dtype, perm, nq, rp, (double *)out,
                                                                            row_stride, col_offset + base * ncol,
                                                                            descriptor_mask);
        NBR_LAUNCHED();
    }
    return NBR_OK;
}

// ------------------------------------------------------------------------------------------------
// neighbor index sets
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
radius_count_kernel(LatticeDev L, const void *__restrict__ query, int dtype, int64_t nq, double radius,
                    int64_t *__restrict__ counts)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nq) return;
    double q[3], f;
    int c[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        q[a] = load_coord(query, dtype, i, 3, a);
        query_anchor(q[a], L.g, a, c[a], f);
    }
    long long n = 0;
    for_each_member(L, q, c, radius, [&](int, int, int, uint32_t, int, int) { ++n; });
    counts[i] = n;
}

__global__ void __launch_bounds__(128)
radius_fill_kernel(LatticeDev L, const void *__restrict__ query, int dtype, int64_t nq, double radius,
                   const int64_t *__restrict__ offsets, int32_t *__restrict__ indices)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nq) return;
    double q[3], f;
    int c[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        q[a] = load_coord(query, dtype, i, 3, a);
        query_anchor(q[a], L.g, a, c[a], f);
    }
    int32_t *dst = indices + offsets[i];
    for_each_member(L, q, c, radius, [&](int, int, int, uint32_t slot, int word, int b) {
        const int64_t w = (int64_t)slot * BRICK_WORDS + word;
        const uint32_t below = L.pool[w] & ((1u << b) - 1u);
        *dst++ = (int32_t)(L.rowbase[w] + __popc(below));
    });
}

int radius_sets(const Lattice *lat, const void *query, int dtype, int64_t nq, double radius, int64_t *offsets,
                int32_t *indices, cudaStream_t stream)
{
    if (!lat->indexed) return fail(NBR_ERR_INVALID, "radius_sets: lattice was built without NBR_LATTICE_INDEXED");
    if (nq <= 0) {
        if (!indices) NBR_CUDA(cudaMemsetAsync(offsets, 0, sizeof(int64_t), stream));
        return NBR_OK;
    }
    const unsigned blocks = (unsigned)ceil_div(nq, 128);
    if (!indices) {
        // counts go to offsets[1..nq]; then an exclusive scan in place over offsets[0..nq]
        NBR_CUDA(cudaMemsetAsync(offsets, 0, sizeof(int64_t), stream));
        Scratch counts;
        NBR_TRY(counts.alloc(sizeof(int64_t) * (nq + 1), stream));
        radius_count_kernel<<<blocks, 128, 0, stream>>>(lat->dev(), query, dtype, nq, radius, counts.as<int64_t>());
        NBR_LAUNCHED();
        NBR_CUDA(cudaMemsetAsync(counts.as<int64_t>() + nq, 0, sizeof(int64_t), stream));
        NBR_TRY((exclusive_scan<int64_t, int64_t>(counts.as<int64_t>(), offsets, nq + 1, stream)));
    } else {
        radius_fill_kernel<<<blocks, 128, 0, stream>>>(lat->dev(), query, dtype, nq, radius, offsets, indices);
        NBR_LAUNCHED();
    }
    return NBR_OK;
}

}  // namespace nbr
