// knn.cu -- k nearest voxels per query on the bit-brick lattice; total order (d^2 as float64, index).
// No reference counterpart (extension, SURVEY 8c / BASELINE config 3).
//
// one warp per query.  the warp sweeps a cubic window of cells around the query's anchor cell: each
// 32-cell occupancy word maps bit b -> lane b, every lane evaluates its own cell with the exact
// float64 distance, candidates that beat the current k-th are merged into a sorted list kept in
// shared memory.  the window doubles (and the sweep restarts) until the k-th distance is provably
// smaller than the distance to any cell outside the window.
#include "common.cuh"
#include "finalize.cuh"
#include "lattice.cuh"

namespace nbr {

constexpr int KNN_WARPS = 4;
constexpr int KNN_MAX_K = 128;
constexpr int KNN_PER_LANE = KNN_MAX_K / 32;

struct KnnEntry {
    double d2;
    int32_t idx;
    int32_t kx, ky, kz;
};

struct KsParam {
    int32_t k[16];
    int32_t n;
};

__device__ __forceinline__ bool entry_less(double d2a, int ia, double d2b, int ib)
{
    return d2a < d2b || (d2a == d2b && ia < ib);
}

// insert e into list[0..have) (sorted ascending), keeping at most k entries.  whole warp calls.
__device__ __forceinline__ int warp_insert(KnnEntry *list, int have, int k, const KnnEntry &e, int lane)
{
    // position = number of entries that sort before e
    KnnEntry mine[KNN_PER_LANE];
    int pos = 0;
#pragma unroll
    for (int t = 0; t < KNN_PER_LANE; ++t) {
        const int p = lane + 32 * t;
        bool before = false;
        if (p < have) {
            mine[t] = list[p];
            before = entry_less(mine[t].d2, mine[t].idx, e.d2, e.idx);
        }
        pos += __popc(__ballot_sync(0xffffffffu, before));
    }
    __syncwarp();
#pragma unroll
    for (int t = 0; t < KNN_PER_LANE; ++t) {
        const int p = lane + 32 * t;
        if (p < have && p >= pos && p + 1 < k) list[p + 1] = mine[t];
    }
    if (lane == 0 && pos < k) list[pos] = e;
    __syncwarp();
    return have < k ? have + 1 : k;
}

template <typename OutT>
__global__ void __launch_bounds__(KNN_WARPS * 32)
knn_kernel(LatticeDev L, const void *__restrict__ query, int dtype, int64_t nq, int k, int32_t *__restrict__ idx_out,
           double *__restrict__ d2_out, KsParam ks, OutT *__restrict__ feats, int64_t row_stride, int col_offset,
           int descriptor_mask)
{
    extern __shared__ unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    KnnEntry *list = reinterpret_cast<KnnEntry *>(smem_raw) + (size_t)warp * k;
    const GridDev &g = L.g;
    const int ncol = (descriptor_mask & NBR_DESC_EXTENDED) ? NBR_COLS_EXTENDED : NBR_COLS_REFERENCE;

    for (int64_t qi = (int64_t)blockIdx.x * KNN_WARPS + warp; qi < nq; qi += (int64_t)gridDim.x * KNN_WARPS) {
        double q[3], f[3];
        int c[3];
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            q[a] = load_coord(query, dtype, qi, 3, a);
            query_anchor(q[a], g, a, c[a], f[a]);
        }
        int have = 0;
        long long rho = 2;
        for (;;) {
            have = 0;
            int lo[3], hi[3];
            bool covers_all = true, empty = false;
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                long long l = (long long)c[a] - rho, h = (long long)c[a] + rho;
                covers_all &= (l <= 0) & (h >= g.ncell[a] - 1);
                lo[a] = (int)(l < 0 ? 0 : l);
                hi[a] = (int)(h > g.ncell[a] - 1 ? g.ncell[a] - 1 : h);
                empty |= lo[a] > hi[a];
            }
            double worst_d2 = INFINITY;
            int worst_idx = 0x7fffffff;
            if (!empty) {
                for (int kz = lo[2]; kz <= hi[2]; ++kz) {
                    const double dz2 = sqdiff(q[2], grid_centre(g, kz, 2));
                    for (int ky = lo[1]; ky <= hi[1]; ++ky) {
                        const double dy2 = sqdiff(q[1], grid_centre(g, ky, 1));
                        const int word = ((kz & (BRICK_Z - 1)) << BRICK_YS) | (ky & (BRICK_Y - 1));
                        const int64_t rowb = ((int64_t)(kz >> BRICK_ZS) * L.nby + (ky >> BRICK_YS)) * L.nbx;
                        for (int bx = lo[0] >> BRICK_XS; bx <= hi[0] >> BRICK_XS; ++bx) {
                            const uint32_t slot = L.dir[rowb + bx];
                            if (!slot) continue;
                            const int64_t wi = (int64_t)slot * BRICK_WORDS + word;
                            const uint32_t full = L.pool[wi];
                            uint32_t w = full;
                            const int x0 = bx << BRICK_XS;
                            if (lo[0] > x0) w &= ~0u << (lo[0] - x0);
                            if (hi[0] < x0 + 31) w &= ~0u >> (x0 + 31 - hi[0]);
                            if (!w) continue;
                            // bit b -> lane b
                            KnnEntry e;
                            bool cand = (w >> lane) & 1u;
                            e.kx = x0 + lane; e.ky = ky; e.kz = kz;
                            e.d2 = INFINITY; e.idx = 0;
                            if (cand) {
                                double s = sqdiff(q[0], grid_centre(g, e.kx, 0));
                                s = __dadd_rn(s, dy2);
                                s = __dadd_rn(s, dz2);
                                e.d2 = s;
                                e.idx = (int32_t)(L.rowbase[wi] + __popc(full & ((1u << lane) - 1u)));
                                cand = have < k || entry_less(s, e.idx, worst_d2, worst_idx);
                            }
                            uint32_t todo = __ballot_sync(0xffffffffu, cand);
                            while (todo) {
                                const int src = __ffs(todo) - 1;
                                todo &= todo - 1;
                                KnnEntry b;
                                b.d2 = __shfl_sync(0xffffffffu, e.d2, src);
                                b.idx = __shfl_sync(0xffffffffu, e.idx, src);
                                b.kx = x0 + src; b.ky = ky; b.kz = kz;
                                if (have == k && !entry_less(b.d2, b.idx, worst_d2, worst_idx)) continue;
                                have = warp_insert(list, have, k, b, lane);
                                if (have == k) {
                                    worst_d2 = list[k - 1].d2;
                                    worst_idx = list[k - 1].idx;
                                }
                            }
                        }
                    }
                }
            }
            if (covers_all) break;
            if (have == k) {
                const double bound = ((double)rho + 0.5) * g.edge * (1.0 - 1e-9);
                if (worst_d2 < bound * bound) break;
            }
            rho *= 2;
            if (rho > (1ll << 31)) break;
        }
        __syncwarp();
        // outputs
        for (int p = lane; p < k; p += 32) {
            if (idx_out) idx_out[qi * k + p] = p < have ? list[p].idx : -1;
            if (d2_out) d2_out[qi * k + p] = p < have ? list[p].d2 : INFINITY;
        }
        if (feats) {
            for (int s = 0; s < ks.n; ++s) {
                const int kk = ks.k[s] < have ? ks.k[s] : have;
                long long acc[10];
#pragma unroll
                for (int t = 0; t < 10; ++t) acc[t] = 0;
                for (int p = lane; p < kk; p += 32) {
                    const long long jx = list[p].kx - c[0], jy = list[p].ky - c[1], jz = list[p].kz - c[2];
                    acc[0] += 1; acc[1] += jx; acc[2] += jy; acc[3] += jz;
                    acc[4] += jx * jx; acc[5] += jx * jy; acc[6] += jx * jz;
                    acc[7] += jy * jy; acc[8] += jy * jz; acc[9] += jz * jz;
                }
#pragma unroll
                for (int t = 0; t < 10; ++t)
#pragma unroll
                    for (int o = 16; o; o >>= 1) acc[t] += __shfl_xor_sync(0xffffffffu, acc[t], o);
                if (lane == 0) {
                    Moments m;
                    m.n = acc[0];
                    m.s1[0] = acc[1]; m.s1[1] = acc[2]; m.s1[2] = acc[3];
#pragma unroll
                    for (int t = 0; t < 6; ++t) m.s2[t] = acc[4 + t];
                    emit_features<OutT>(m, f, g.edge, feats + qi * row_stride + col_offset + s * ncol, descriptor_mask);
                }
            }
        }
        __syncwarp();
    }
}

int knn(const Lattice *lat, const void *query, int dtype, int64_t nq, int k, int32_t *idx_out, double *d2_out,
        const int32_t *ks, int n_k, void *feats, int out_dtype, int64_t row_stride, int col_offset, int descriptor_mask,
        cudaStream_t stream)
{
    if (!lat->indexed) return fail(NBR_ERR_INVALID, "knn: lattice was built without NBR_LATTICE_INDEXED");
    if (lat->grid.ndim != 3) return fail(NBR_ERR_INVALID, "knn: 3-D lattices only");
    if (k < 1 || k > KNN_MAX_K) return fail(NBR_ERR_UNSUPPORTED, "knn: k must be in [1, 128]");
    if (n_k < 0 || n_k > 16) return fail(NBR_ERR_UNSUPPORTED, "knn: at most 16 values of k");
    if (nq <= 0) return NBR_OK;
    KsParam kp;
    kp.n = feats ? n_k : 0;
    for (int i = 0; i < kp.n; ++i) {
        if (ks[i] < 1 || ks[i] > k) return fail(NBR_ERR_INVALID, "knn: every ks[i] must be in [1, k]");
        kp.k[i] = ks[i];
    }
    const size_t smem = sizeof(KnnEntry) * (size_t)k * KNN_WARPS;
    const int blocks = (int)std::min<int64_t>(ceil_div(nq, KNN_WARPS), (int64_t)device_sm_count() * 16);
    if (out_dtype == NBR_F32)
        knn_kernel<float><<<blocks, KNN_WARPS * 32, smem, stream>>>(lat->dev(), query, dtype, nq, k, idx_out, d2_out, kp,
                                                                     (float *)feats, row_stride, col_offset, descriptor_mask);
    else
        knn_kernel<double><<<blocks, KNN_WARPS * 32, smem, stream>>>(lat->dev(), query, dtype, nq, k, idx_out, d2_out, kp,
                                                                      (double *)feats, row_stride, col_offset, descriptor_mask);
    NBR_LAUNCHED();
    return NBR_OK;
}

}  // namespace nbr

extern "C" int nbr_knn(const nbr_lattice *lattice, const void *query_xyz, int dtype, int64_t n_query, int32_t k,
                       int32_t *idx_out, double *d2_out, const int32_t *ks_host, int32_t n_k, void *feats_out,
                       int out_dtype, int64_t out_row_stride, int32_t col_offset, int32_t descriptor_mask, void *stream)
{
    if (!lattice || !query_xyz) return nbr::fail(NBR_ERR_INVALID, "nbr_knn: null argument");
    if (dtype != NBR_F32 && dtype != NBR_F64) return nbr::fail(NBR_ERR_INVALID, "nbr_knn: bad dtype");
    if (feats_out && (!ks_host || n_k < 1)) return nbr::fail(NBR_ERR_INVALID, "nbr_knn: feats_out needs ks");
    return nbr::knn(reinterpret_cast<const nbr::Lattice *>(lattice), query_xyz, dtype, n_query, k, idx_out, d2_out,
                    ks_host, n_k, feats_out, out_dtype, out_row_stride, col_offset, descriptor_mask,
                    (cudaStream_t)stream);
}
