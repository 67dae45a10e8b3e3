// knn.cu -- k nearest voxels per query on the bit-brick lattice; total order (d^2 as float64, index).
// No reference counterpart (extension, SURVEY 8c / BASELINE config 3).
//
// one warp per query, three steps:
//   1. candidate ball.  the warp sweeps the bricks under a cubic window of half-width W cells around the
//      query's anchor cell: lane b resolves brick b through the directory, every non-empty brick is then
//      read as one coalesced 128-byte line (lane = row of the brick).  cells outside the window are at
//      least (W + 0.5) e away, so once >= k cells lie inside a ball of that radius the k nearest are
//      among them.  the ball test runs in float32 with a relative guard band; W grows until the ball is
//      populated, and the ball shrinks (bisection on r^2) if it holds more candidates than the buffer.
//   2. every candidate gets the reference's float64 squared distance ((dx^2 + dy^2) + dz^2, no fma) and
//      its index in np.unique order (row base + popcount); records go to shared memory.
//   3. the warp sorts the records with a bitonic network on (d^2, index) -- a strict total order, ties
//      included -- writes the first k, and accumulates the integer moments of every k in ks for the
//      feature columns (finalize.cuh), one lane per k.
#include "common.cuh"
#include "finalize.cuh"
#include "lattice.cuh"

#include <stdlib.h>

#include <string>

namespace nbr {

int knn_points(const void *search, int s_dtype, int64_t ns, const void *query, int q_dtype, int64_t nq, int k, double cell_edge,
               int32_t *idx_out, double *d2_out, const int32_t *ks, int n_k, void *feats, int out_dtype, int64_t row_stride,
               int col_offset, int descriptor_mask, cudaStream_t stream);
int lattice_centres(const Lattice *L, int64_t nv, double *centres, cudaStream_t stream);

constexpr int KNN_WARPS = 4;
constexpr int KNN_MAX_K = 128;
constexpr int KNN_CAP = 512;            // candidate records per warp (8 KB)

struct __align__(16) KnnRec {
    double d2;
    int32_t idx;
    int32_t pad;
};

struct KsParam {
    int32_t k[16];
    int32_t n;
};

__device__ __forceinline__ bool rec_less(const KnnRec &a, const KnnRec &b)
{
    return a.d2 < b.d2 || (a.d2 == b.d2 && a.idx < b.idx);
}

template <typename OutT>
__global__ void __launch_bounds__(KNN_WARPS * 32)
knn_kernel(LatticeDev L, const void *__restrict__ query, int dtype, int64_t nq, int k, int32_t *__restrict__ idx_out,
           double *__restrict__ d2_out, KsParam ks, OutT *__restrict__ feats, int64_t row_stride, int col_offset,
           int descriptor_mask)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    KnnRec *rec = reinterpret_cast<KnnRec *>(smem_raw) + (size_t)warp * KNN_CAP;
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
        // query position relative to the anchor cell's centre, cell units (float32 guide only)
        const float px = (float)f[0] - 0.5f, py = (float)f[1] - 0.5f, pz = (float)f[2] - 0.5f;

        // ---- 1. candidate ball
        // first window: the ball of radius W + 0.5 cells cuts about k cells out of a populated plane
        long long W = (long long)ceilf(sqrtf((float)k * 0.318309886f) - 0.2f);
        if (W < 1) W = 1;
        float lo2 = 0.0f, hi2 = INFINITY, r2 = 0.0f;
        bool ball_of_w = true;                                    // r2 is the full ball of the current window
        int n_cand = 0;
        for (int attempt = 0; attempt < 200; ++attempt) {
            int wlo[3], whi[3];
            bool covers_all = true, empty = false;
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                const long long l = (long long)c[a] - W, h = (long long)c[a] + W;
                covers_all &= (l <= 0) & (h >= g.ncell[a] - 1);
                wlo[a] = (int)(l < 0 ? 0 : l);
                whi[a] = (int)(h > g.ncell[a] - 1 ? g.ncell[a] - 1 : h);
                empty |= wlo[a] > whi[a];
            }
            if (ball_of_w) {
                const float wf = (float)W + 0.5f;
                r2 = covers_all ? INFINITY : wf * wf * (1.0f - 1.0e-4f);
            }
            const float r2_sure = r2 * (1.0f - 4.0e-5f);
            int n_sure = 0;
            n_cand = 0;
            if (!empty) {
                const int bx0 = wlo[0] >> BRICK_XS, by0 = wlo[1] >> BRICK_YS, bz0 = wlo[2] >> BRICK_ZS;
                const int nbx = (whi[0] >> BRICK_XS) - bx0 + 1, nby = (whi[1] >> BRICK_YS) - by0 + 1;
                const int nbz = (whi[2] >> BRICK_ZS) - bz0 + 1;
                const long long nbricks = (long long)nbx * nby * nbz;
                for (long long base = 0; base < nbricks; base += 32) {
                    const long long b = base + lane;
                    uint32_t slot = 0;
                    int gx = 0, gy = 0, gz = 0;
                    if (b < nbricks) {
                        gx = bx0 + (int)(b % nbx);
                        gy = by0 + (int)((b / nbx) % nby);
                        gz = bz0 + (int)(b / ((long long)nbx * nby));
                        slot = L.dir[((int64_t)gz * L.nby + gy) * L.nbx + gx];
                    }
                    uint32_t todo = __ballot_sync(0xffffffffu, slot != 0);
                    // the line of the next non-empty brick is requested before the current one is processed
                    uint32_t next_full = 0;
                    if (todo) next_full = L.pool[(int64_t)__shfl_sync(0xffffffffu, slot, __ffs(todo) - 1) * BRICK_WORDS + lane];
                    while (todo) {
                        const int src = __ffs(todo) - 1;
                        todo &= todo - 1;
                        const uint32_t s = __shfl_sync(0xffffffffu, slot, src);
                        const int x0 = __shfl_sync(0xffffffffu, gx, src) << BRICK_XS;
                        const int ky = (__shfl_sync(0xffffffffu, gy, src) << BRICK_YS) | (lane & (BRICK_Y - 1));
                        const int kz = (__shfl_sync(0xffffffffu, gz, src) << BRICK_ZS) | (lane >> BRICK_YS);
                        const int64_t wi = (int64_t)s * BRICK_WORDS + lane;
                        const uint32_t full = next_full;                // one 128-byte line per brick
                        if (todo) next_full = L.pool[(int64_t)__shfl_sync(0xffffffffu, slot, __ffs(todo) - 1) * BRICK_WORDS + lane];
                        uint32_t w = full;
                        if (wlo[0] > x0) w &= ~0u << (wlo[0] - x0);
                        if (whi[0] < x0 + 31) w &= ~0u >> (x0 + 31 - whi[0]);
                        if (ky < wlo[1] || ky > whi[1] || kz < wlo[2] || kz > whi[2]) w = 0;
                        // float32 ball test of the row's cells
                        const float dy = (float)(ky - c[1]) - py, dz = (float)(kz - c[2]) - pz;
                        const float row2 = dy * dy + dz * dz;
                        uint32_t cand = 0;
                        int sure = 0;
                        for (uint32_t u = w; u; u &= u - 1) {
                            const int bit = __ffs(u) - 1;
                            const float dx = (float)(x0 + bit - c[0]) - px;
                            const float d2 = fmaf(dx, dx, row2);
                            if (d2 <= r2) cand |= 1u << bit;
                            sure += d2 <= r2_sure;
                        }
                        n_sure += sure;
                        // records: position by a warp prefix sum of the per-row counts
                        const int mine = __popc(cand);
                        const int total = __reduce_add_sync(0xffffffffu, mine);
                        if (total == 0) continue;
                        int pre = mine;
#pragma unroll
                        for (int o = 1; o < 32; o <<= 1) {
                            const int t = __shfl_up_sync(0xffffffffu, pre, o);
                            if (lane >= o) pre += t;
                        }
                        int pos = n_cand + pre - mine;
                        n_cand += total;
                        if (cand && n_cand <= KNN_CAP) {
                            const double dy2 = sqdiff(q[1], grid_centre(g, ky, 1));
                            const double dz2 = sqdiff(q[2], grid_centre(g, kz, 2));
                            const uint32_t rb = L.rowbase[wi];
                            for (uint32_t u = cand; u; u &= u - 1) {
                                const int bit = __ffs(u) - 1;
                                double s2 = sqdiff(q[0], grid_centre(g, x0 + bit, 0));
                                s2 = __dadd_rn(s2, dy2);
                                s2 = __dadd_rn(s2, dz2);
                                KnnRec r;
                                r.d2 = s2;
                                r.idx = (int32_t)(rb + __popc(full & ((1u << bit) - 1u)));
                                // cell offset from the anchor cell, 10 bits per axis (bit 31 set: too far, use the key)
                                const int jx = x0 + bit - c[0], jy = ky - c[1], jz = kz - c[2];
                                const bool fits = (unsigned)(jx + 512) < 1024u && (unsigned)(jy + 512) < 1024u && (unsigned)(jz + 512) < 1024u;
                                r.pad = fits ? (jx + 512) | ((jy + 512) << 10) | ((jz + 512) << 20) : (int)0x80000000;
                                rec[pos] = r;
                                ++pos;
                            }
                        }
                    }
                }
            }
            n_sure = __reduce_add_sync(0xffffffffu, n_sure);
            // ---- decide: enough sure cells, and the candidates fit?
            // invariants: fewer than k sure cells at lo2; more than KNN_CAP candidates at hi2
            if (n_sure >= k && n_cand <= KNN_CAP) break;
            if (n_sure >= k) {
                hi2 = r2;                                   // too many: shrink the ball inside the same window
                ball_of_w = false;
            } else if (!ball_of_w) {
                lo2 = r2;                                   // the shrunken ball lost the k-th neighbor
            } else if (covers_all) {
                break;                                      // the whole lattice holds fewer than k voxels
            } else {
                lo2 = r2;                                   // the ball of this window is not populated: widen it
                W = W < 4 ? W + 1 : (W * 3) / 2;
                if (W > (1ll << 31)) break;
                continue;
            }
            const float mid = hi2 < INFINITY ? 0.5f * (lo2 + hi2) : fmaxf(4.0f * lo2, 1.0f);
            if (!(mid > lo2 && mid < hi2)) break;           // unreachable: > KNN_CAP - k cells within one ulp of distance
            r2 = mid;
        }
        if (n_cand > KNN_CAP) n_cand = KNN_CAP;             // unreachable guard
        __syncwarp();

        // ---- 3. bitonic sort of the records (padded with +inf to a power of two)
        int npow = 32;
        while (npow < n_cand) npow <<= 1;
        for (int p = n_cand + lane; p < npow; p += 32) {
            KnnRec r;
            r.d2 = INFINITY; r.idx = 0x7fffffff; r.pad = 0;
            rec[p] = r;
        }
        __syncwarp();
        for (int size = 2; size <= npow; size <<= 1) {
            for (int stride = size >> 1; stride > 0; stride >>= 1) {
                for (int t = lane; t < (npow >> 1); t += 32) {
                    const int i = 2 * t - (t & (stride - 1));         // lower index of the pair
                    const int j = i + stride;
                    const bool up = (i & size) == 0;
                    const KnnRec a = rec[i], b = rec[j];
                    if (rec_less(b, a) == up) { rec[i] = b; rec[j] = a; }
                }
                __syncwarp();
            }
        }
        const int have = n_cand < k ? n_cand : k;

        // ---- outputs
        for (int p = lane; p < k; p += 32) {
            if (idx_out) idx_out[qi * k + p] = p < have ? rec[p].idx : -1;
            if (d2_out) d2_out[qi * k + p] = p < have ? rec[p].d2 : INFINITY;
        }
        if (feats) {
            // integer moments of the first ks[s] records; lane s finalises scale s
            Moments mine;
            mine.n = 0;
#pragma unroll
            for (int t = 0; t < 3; ++t) mine.s1[t] = 0;
#pragma unroll
            for (int t = 0; t < 6; ++t) mine.s2[t] = 0;
            // ks ascending: every record is read once, the sums of [ks[s-1], ks[s]) are added to running totals
            long long run[10];
#pragma unroll
            for (int t = 0; t < 10; ++t) run[t] = 0;
            int first = 0;
            for (int s = 0; s < ks.n; ++s) {
                const int kk = ks.k[s] < have ? ks.k[s] : have;
                int acc[10];
#pragma unroll
                for (int t = 0; t < 10; ++t) acc[t] = 0;
                bool wide = false;
                for (int p = first + lane; p < kk; p += 32) {
                    const int pk = rec[p].pad;
                    wide |= pk < 0;
                    const int jx = (pk & 1023) - 512, jy = ((pk >> 10) & 1023) - 512, jz = ((pk >> 20) & 1023) - 512;
                    acc[0] += 1; acc[1] += jx; acc[2] += jy; acc[3] += jz;
                    acc[4] += jx * jx; acc[5] += jx * jy; acc[6] += jx * jz;
                    acc[7] += jy * jy; acc[8] += jy * jz; acc[9] += jz * jz;
                }
                if (__any_sync(0xffffffffu, wide)) {
                    // a window wider than 511 cells: offsets from the packed addresses (utils/geometry.py:120-131), 64-bit sums
                    long long big[10];
#pragma unroll
                    for (int t = 0; t < 10; ++t) big[t] = 0;
                    for (int p = first + lane; p < kk; p += 32) {
                        const uint64_t key = L.ukeys[rec[p].idx];
                        long long j[3];
#pragma unroll
                        for (int a = 0; a < 3; ++a) {
                            const uint64_t mask = g.widths[a] >= 64 ? ~0ull : ((1ull << g.widths[a]) - 1);
                            j[a] = (long long)((key >> g.shifts[a]) & mask) - g.cell_lo[a] - c[a];
                        }
                        big[0] += 1; big[1] += j[0]; big[2] += j[1]; big[3] += j[2];
                        big[4] += j[0] * j[0]; big[5] += j[0] * j[1]; big[6] += j[0] * j[2];
                        big[7] += j[1] * j[1]; big[8] += j[1] * j[2]; big[9] += j[2] * j[2];
                    }
#pragma unroll
                    for (int t = 0; t < 10; ++t) {
#pragma unroll
                        for (int o = 16; o; o >>= 1) big[t] += __shfl_xor_sync(0xffffffffu, big[t], o);
                        run[t] += big[t];
                    }
                } else {
#pragma unroll
                    for (int t = 0; t < 10; ++t) run[t] += (long long)__reduce_add_sync(0xffffffffu, acc[t]);
                }
                first = kk > first ? kk : first;
                if (lane == s) {
                    mine.n = run[0];
                    mine.s1[0] = run[1]; mine.s1[1] = run[2]; mine.s1[2] = run[3];
#pragma unroll
                    for (int t = 0; t < 6; ++t) mine.s2[t] = run[4 + t];
                }
            }
            if (lane < ks.n)
                emit_features<OutT>(mine, f, g.edge, feats + qi * row_stride + col_offset + lane * ncol, descriptor_mask);
        }
        __syncwarp();
    }
}

static size_t knn_smem() { return sizeof(KnnRec) * (size_t)KNN_CAP * KNN_WARPS; }

int knn(const Lattice *lat, const void *query, int dtype, int64_t nq, int k, int32_t *idx_out, double *d2_out,
        const int32_t *ks, int n_k, void *feats, int out_dtype, int64_t row_stride, int col_offset, int descriptor_mask,
        cudaStream_t stream)
{
    if (!lat->indexed) return fail(NBR_ERR_INVALID, "knn: lattice was built without NBR_LATTICE_INDEXED");
    if (lat->grid.ndim != 3) return fail(NBR_ERR_INVALID, "knn: 3-D lattices only");
    if (k < 1 || k > KNN_MAX_K) return fail(NBR_ERR_UNSUPPORTED, "knn: k must be in [1, 128]");
    if (n_k < 0 || n_k > 16) return fail(NBR_ERR_UNSUPPORTED, "knn: at most 16 values of k");
    if (nq <= 0) return NBR_OK;
    // default route: the voxel centres (np.unique order, so a centre's position IS its index) go through the point
    // kernel of knn_points.cu -- same float64 distances on the same centres, same (d^2, index) order, measured 46 M
    // against 35 M queries/s at k = 50 on 9.4M voxels.  NBR_KNN=bricks keeps the brick sweep below.
    static const bool use_bricks = getenv("NBR_KNN") && std::string(getenv("NBR_KNN")) == "bricks";
    if (!use_bricks) {
        int64_t nv = 0;
        NBR_TRY(lattice_counts(lat, &nv, nullptr));
        if (nv >= 1) {
            Scratch centres;
            NBR_TRY(centres.alloc(sizeof(double) * 3 * (size_t)nv, stream));
            NBR_TRY(lattice_centres(lat, nv, centres.as<double>(), stream));
            return knn_points(centres.ptr, NBR_F64, nv, query, dtype, nq, k, 0.0, idx_out, d2_out, ks, n_k, feats, out_dtype, row_stride,
                              col_offset, descriptor_mask, stream);
        }
    }
    KsParam kp;
    kp.n = feats ? n_k : 0;
    for (int i = 0; i < kp.n; ++i) {
        if (ks[i] < 1 || ks[i] > k) return fail(NBR_ERR_INVALID, "knn: every ks[i] must be in [1, k]");
        if (i > 0 && ks[i] <= ks[i - 1]) return fail(NBR_ERR_INVALID, "knn: ks must be ascending");
        kp.k[i] = ks[i];
    }
    const size_t smem = knn_smem();
    static std::atomic<uint64_t> configured{0};
    if (first_use_on_device(configured)) {
        NBR_CUDA(cudaFuncSetAttribute(knn_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        NBR_CUDA(cudaFuncSetAttribute(knn_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    const int blocks = (int)std::min<int64_t>(ceil_div(nq, KNN_WARPS), (int64_t)device_sm_count() * 8);
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
