// knn_points.cu -- k nearest RAW POINTS per query (no voxel filter on the search set); total order (d^2 as float64,
// index in the search cloud).  no reference counterpart in nimrud/minimal (extension, BASELINE config 3 "raw-point
// variant"); the precedent is the legacy pipeline's sspedge = 0, which searches the unfiltered cloud
// (nimrud/prototypes/mso.py:277,303-308).
//
// index: counting sort of the search points on a dense grid of cubic cells (a few points per occupied cell; the edge
// is doubled until the grid fits 2^26 cells), float64 copies of the points in cell order + their original indices.
// query: one warp per query.  the warp sweeps the cells of a cubic window of half-width W cells around the query's
// cell; every point outside the window is farther than the distance rb to the window's nearest open face, so once the
// ball of radius rb holds >= k points the k nearest are among them.  squared distances are the float64 expression
// ((dx*dx + dy*dy) + dz*dz) without fma (the same expression a float64 brute force evaluates).  W grows until the ball is
// populated; the ball shrinks by bisection if it holds more candidates than the record buffer.  the records are sorted
// by a bitonic network on (d^2, index) -- ties included -- and the first k written; for every k in ks the float64
// moments of the k nearest points (relative to the query) go through finalize.cuh like every other neighborhood.
#include "common.cuh"
#include "finalize.cuh"
#include "scan.cuh"

namespace nbr {

int bbox(const void *xyz, int dtype, int64_t n, int ndim, double *lohi_dev, cudaStream_t stream);

constexpr int KP_WARPS = 4;
constexpr int KP_MAX_K = 128;
constexpr int KP_CAP = 512;             // candidate records per warp (8 KB)

struct __align__(16) KpRec {
    double d2;
    int32_t idx;                        // index in the caller's search cloud
    int32_t pos;                        // position in the cell-ordered copy
};

struct KpKs {
    int32_t k[16];
    int32_t n;
};

struct PGrid {
    double origin[3];
    double g, inv_g;
    int dims[3];
};

__device__ __forceinline__ bool kp_less(const KpRec &a, const KpRec &b)
{
    return a.d2 < b.d2 || (a.d2 == b.d2 && a.idx < b.idx);
}

__device__ __forceinline__ int kp_cell(double p, double origin, double inv_g, int dim)
{
    return clampi((int)fmin(fmax(floor((p - origin) * inv_g), -1.0e9), 1.0e9), 0, dim - 1);
}

__global__ void __launch_bounds__(256)
kp_count_kernel(const void *__restrict__ xyz, int dtype, int64_t n, PGrid G, uint32_t *__restrict__ counts,
                uint32_t *__restrict__ cell, uint32_t *__restrict__ rank)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int c[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) c[a] = kp_cell(load_coord(xyz, dtype, i, 3, a), G.origin[a], G.inv_g, G.dims[a]);
    const uint32_t id = ((uint32_t)c[2] * (uint32_t)G.dims[1] + (uint32_t)c[1]) * (uint32_t)G.dims[0] + (uint32_t)c[0];
    cell[i] = id;
    rank[i] = atomicAdd(&counts[id], 1u);
}

__global__ void __launch_bounds__(256)
kp_place_kernel(const void *__restrict__ xyz, int dtype, int64_t n, const uint32_t *__restrict__ offsets,
                const uint32_t *__restrict__ cell, const uint32_t *__restrict__ rank, double *__restrict__ sorted,
                int32_t *__restrict__ orig)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int64_t pos = (int64_t)offsets[cell[i]] + rank[i];
#pragma unroll
    for (int a = 0; a < 3; ++a) sorted[pos * 3 + a] = load_coord(xyz, dtype, i, 3, a);
    orig[pos] = (int32_t)i;
}

template <typename OutT>
__global__ void __launch_bounds__(KP_WARPS * 32, 6)
knn_points_kernel(PGrid G, const double *__restrict__ pts, const int32_t *__restrict__ orig, const uint32_t *__restrict__ offsets,
                  int64_t n_search, const void *__restrict__ query, int q_dtype, int64_t nq, int k, int32_t *__restrict__ idx_out,
                  double *__restrict__ d2_out, KpKs ks, OutT *__restrict__ feats, int64_t row_stride, int col_offset,
                  int descriptor_mask, int cap)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ int s_count[KP_WARPS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    KpRec *rec = reinterpret_cast<KpRec *>(smem_raw) + (size_t)warp * cap;       // cap records per warp (a power of two >= 4 k)
    const int ncol = (descriptor_mask & NBR_DESC_EXTENDED) ? NBR_COLS_EXTENDED : NBR_COLS_REFERENCE;
    const int64_t n_cells = (int64_t)G.dims[0] * G.dims[1] * G.dims[2];

    for (int64_t qi = (int64_t)blockIdx.x * KP_WARPS + warp; qi < nq; qi += (int64_t)gridDim.x * KP_WARPS) {
        double q[3];
        long long ca[3];
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            q[a] = load_coord(query, q_dtype, qi, 3, a);
            ca[a] = (long long)fmin(fmax(floor((q[a] - G.origin[a]) * G.inv_g), -4.0e9), 4.0e9);   // not clamped to the grid
        }
        // first window: the cells hold ~4 points each where the cloud is a surface and the ball of a window of half-width
        // W has a radius of about W + 0.5 cells, i.e. about 4 pi (W + 0.5)^2 points: start where that reaches k.  (a window
        // one step too wide costs more than an extra attempt: 343 cells against 27 + 125)
        long long W = (long long)ceilf(sqrtf((float)k * 0.0796f) - 0.5f);
        if (W < 1) W = 1;
        double lo2 = 0.0, hi2 = INFINITY, r2 = 0.0;
        bool ball_of_w = true;
        int n_cand = 0;
        for (int attempt = 0; attempt < 400; ++attempt) {
            int wlo[3], whi[3];
            bool covers_all = true, empty = false;
            double rb = INFINITY;                       // distance to the nearest face of the window that has grid behind it
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                const long long l = ca[a] - W, h = ca[a] + W;
                if (l > 0) {
                    covers_all = false;
                    rb = fmin(rb, q[a] - (G.origin[a] + (double)l * G.g));
                }
                if (h < G.dims[a] - 1) {
                    covers_all = false;
                    rb = fmin(rb, (G.origin[a] + (double)(h + 1) * G.g) - q[a]);
                }
                wlo[a] = (int)(l < 0 ? 0 : (l > G.dims[a] - 1 ? G.dims[a] : l));
                whi[a] = (int)(h > G.dims[a] - 1 ? G.dims[a] - 1 : (h < 0 ? -1 : h));
                empty |= wlo[a] > whi[a];
            }
            if (ball_of_w) {
                // the faces are origin + l * g in float64 and a point's cell comes from floor((p - origin) / g): shrink the
                // bound by far more than their rounding (a smaller ball is only more conservative)
                rb = fmax(rb - (1.0e-6 * G.g + 1.0e-13 * (fabs(q[0]) + fabs(q[1]) + fabs(q[2]) + fabs(G.origin[0]) + fabs(G.origin[1]) + fabs(G.origin[2]))), 0.0);
                r2 = covers_all ? INFINITY : rb * rb;
            }
            if (lane == 0) s_count[warp] = 0;
            __syncwarp();
            if (!empty) {
                // the cells are numbered x-fastest, so the points of one x-row of the window (whi[0] - wlo[0] + 1 cells) are ONE
                // contiguous range of the cell-ordered copy: a lane per row, two offsets per row
                const int ncy = whi[1] - wlo[1] + 1, ncz = whi[2] - wlo[2] + 1;
                const long long nrows = (long long)ncy * ncz;
                for (long long t = lane; t < nrows; t += 32) {
                    const int iz = wlo[2] + (int)(t / ncy), iy = wlo[1] + (int)(t - (long long)(iz - wlo[2]) * ncy);
                    const int64_t id0 = ((int64_t)iz * G.dims[1] + iy) * G.dims[0] + wlo[0], id1 = id0 + (whi[0] - wlo[0]) + 1;
                    const uint32_t first = offsets[id0], last = id1 < n_cells ? offsets[id1] : (uint32_t)n_search;
                    for (uint32_t p = first; p < last; ++p) {
                        const double dx = __dsub_rn(q[0], pts[(size_t)p * 3]), dy = __dsub_rn(q[1], pts[(size_t)p * 3 + 1]),
                                     dz = __dsub_rn(q[2], pts[(size_t)p * 3 + 2]);
                        const double d2 = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
                        if (d2 <= r2) {
                            const int slot = atomicAdd(&s_count[warp], 1);
                            if (slot < cap) {
                                KpRec r;
                                r.d2 = d2; r.idx = orig[p]; r.pos = (int32_t)p;
                                rec[slot] = r;
                            }
                        }
                    }
                }
            }
            __syncwarp();
            n_cand = s_count[warp];
            __syncwarp();
            // invariants: fewer than k points inside lo2; more than cap inside hi2
            if (n_cand >= k && n_cand <= cap) break;
            if (n_cand >= k) {
                hi2 = r2;                                   // too many: shrink the ball inside the same window
                ball_of_w = false;
            } else if (!ball_of_w) {
                lo2 = r2;                                   // the shrunken ball lost the k-th neighbor
            } else if (covers_all) {
                break;                                      // the whole cloud holds fewer than k points
            } else {
                lo2 = r2;                                   // the ball of this window is not populated: widen it
                W = W < 4 ? W + 1 : (W * 3) / 2;
                if (W > (1ll << 31)) break;
                continue;
            }
            const double mid = hi2 < INFINITY ? 0.5 * (lo2 + hi2) : fmax(4.0 * lo2, 1.0e-300);
            if (!(mid > lo2 && mid < hi2)) break;           // more than KP_CAP - k points at one distance: keep the first KP_CAP found
            r2 = mid;
        }
        if (n_cand > cap) n_cand = cap;
        __syncwarp();

        // ---- thin the candidates before sorting: the ball usually holds 2-3 k points and the sort costs n log^2 n.  a
        // 32-bucket histogram of d^2 (lane = bucket) finds the bucket of the k-th smallest; every record up to and
        // including that bucket stays (all ties of the k-th distance share its bucket), the rest cannot be among the k
        if (n_cand > 64 && n_cand > k + 16) {
            double dmax = 0.0;
            for (int p = lane; p < n_cand; p += 32) dmax = fmax(dmax, rec[p].d2);
#pragma unroll
            for (int o = 16; o; o >>= 1) dmax = fmax(dmax, __shfl_xor_sync(0xffffffffu, dmax, o));
            const double scale = dmax > 0.0 ? 32.0 / dmax : 0.0;
            int mine = 0;                                            // records in bucket `lane`
            for (int p0 = 0; p0 < n_cand; p0 += 32) {
                const int p = p0 + lane;
                const int bkt = p < n_cand ? min((int)(rec[p].d2 * scale), 31) : -1;
#pragma unroll
                for (int b = 0; b < 32; ++b) {
                    const int c = __popc(__ballot_sync(0xffffffffu, bkt == b));
                    if (lane == b) mine += c;
                }
            }
            int cum = mine;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, cum, o);
                if (lane >= o) cum += t;
            }
            const int last_bucket = __ffs(__ballot_sync(0xffffffffu, cum >= k)) - 1;      // >= 0: cum reaches n_cand >= k
            int kept = 0;
            for (int p0 = 0; p0 < n_cand; p0 += 32) {                 // in-place compaction: writes stay behind the reads
                const int p = p0 + lane;
                KpRec r;
                r.d2 = INFINITY; r.idx = 0; r.pos = 0;
                if (p < n_cand) r = rec[p];
                const bool keep = p < n_cand && min((int)(r.d2 * scale), 31) <= last_bucket;
                const uint32_t m = __ballot_sync(0xffffffffu, keep);
                __syncwarp();
                if (keep) rec[kept + __popc(m & lanemask_lt())] = r;
                kept += __popc(m);
                __syncwarp();
            }
            n_cand = kept;
        }

        // ---- bitonic sort of the records (padded with +inf to a power of two)
        int npow = 32;
        while (npow < n_cand) npow <<= 1;
        for (int p = n_cand + lane; p < npow; p += 32) {
            KpRec r;
            r.d2 = INFINITY; r.idx = 0x7fffffff; r.pos = 0;
            rec[p] = r;
        }
        __syncwarp();
        for (int size = 2; size <= npow; size <<= 1) {
            for (int stride = size >> 1; stride > 0; stride >>= 1) {
                for (int t = lane; t < (npow >> 1); t += 32) {
                    const int i = 2 * t - (t & (stride - 1));
                    const int j = i + stride;
                    const bool up = (i & size) == 0;
                    const KpRec a = rec[i], b = rec[j];
                    if (kp_less(b, a) == up) { rec[i] = b; rec[j] = a; }
                }
                __syncwarp();
            }
        }
        const int have = n_cand < k ? n_cand : k;
        for (int p = lane; p < k; p += 32) {
            if (idx_out) idx_out[qi * k + p] = p < have ? rec[p].idx : -1;
            if (d2_out) d2_out[qi * k + p] = p < have ? rec[p].d2 : INFINITY;
        }
        if (feats) {
            // float64 moments of the k nearest points relative to the query; lane s finalises scale s
            double run[10], mine[10];
#pragma unroll
            for (int t = 0; t < 10; ++t) { run[t] = 0.0; mine[t] = 0.0; }
            int first = 0;
            for (int s = 0; s < ks.n; ++s) {
                const int kk = ks.k[s] < have ? ks.k[s] : have;
                double acc[10];
#pragma unroll
                for (int t = 0; t < 10; ++t) acc[t] = 0.0;
                for (int p = first + lane; p < kk; p += 32) {
                    const size_t at = (size_t)rec[p].pos * 3;
                    const double dx = pts[at] - q[0], dy = pts[at + 1] - q[1], dz = pts[at + 2] - q[2];
                    acc[0] += 1.0; acc[1] += dx; acc[2] += dy; acc[3] += dz;
                    acc[4] += dx * dx; acc[5] += dx * dy; acc[6] += dx * dz;
                    acc[7] += dy * dy; acc[8] += dy * dz; acc[9] += dz * dz;
                }
#pragma unroll
                for (int t = 0; t < 10; ++t) {
#pragma unroll
                    for (int o = 16; o; o >>= 1) acc[t] += __shfl_xor_sync(0xffffffffu, acc[t], o);
                    run[t] += acc[t];
                }
                first = kk > first ? kk : first;
                if (lane == s) {
#pragma unroll
                    for (int t = 0; t < 10; ++t) mine[t] = run[t];
                }
            }
            if (lane < ks.n) {
                const double n = mine[0];
                double centroid = 0.0;
                double a[6] = {0, 0, 0, 0, 0, 0};
                if (n > 0.0) {
                    const double inv = 1.0 / n;
                    centroid = sqrt(mine[1] * mine[1] + mine[2] * mine[2] + mine[3] * mine[3]) * inv;
                    a[0] = n * mine[4] - mine[1] * mine[1]; a[1] = n * mine[5] - mine[1] * mine[2]; a[2] = n * mine[6] - mine[1] * mine[3];
                    a[3] = n * mine[7] - mine[2] * mine[2]; a[4] = n * mine[8] - mine[2] * mine[3]; a[5] = n * mine[9] - mine[3] * mine[3];
                }
                emit_core<OutT>((long long)n, centroid, a, 1.0, feats + qi * row_stride + col_offset + lane * ncol, descriptor_mask);
            }
        }
        __syncwarp();
    }
}

int knn_points(const void *search, int s_dtype, int64_t ns, const void *query, int q_dtype, int64_t nq, int k, double cell_edge,
               int32_t *idx_out, double *d2_out, const int32_t *ks, int n_k, void *feats, int out_dtype, int64_t row_stride,
               int col_offset, int descriptor_mask, cudaStream_t stream)
{
    if (k < 1 || k > KP_MAX_K) return fail(NBR_ERR_UNSUPPORTED, "knn_points: k must be in [1, 128]");
    if (n_k < 0 || n_k > 16) return fail(NBR_ERR_UNSUPPORTED, "knn_points: at most 16 values of k");
    if (ns < 1) return fail(NBR_ERR_TOO_FEW_POINTS, "knn_points: empty search cloud");
    if (ns >= (int64_t)1 << 31) return fail(NBR_ERR_UNSUPPORTED, "knn_points: more than 2^31 search points");
    if (nq <= 0) return NBR_OK;
    KpKs kp;
    kp.n = feats ? n_k : 0;
    for (int i = 0; i < kp.n; ++i) {
        if (ks[i] < 1 || ks[i] > k) return fail(NBR_ERR_INVALID, "knn_points: every ks[i] must be in [1, k]");
        if (i > 0 && ks[i] <= ks[i - 1]) return fail(NBR_ERR_INVALID, "knn_points: ks must be ascending");
        kp.k[i] = ks[i];
    }
    // ---- grid over the search cloud's bounding box
    Scratch box;
    double lohi[6];
    NBR_TRY(box.alloc(sizeof(double) * 6, stream));
    NBR_TRY(bbox(search, s_dtype, ns, 3, box.as<double>(), stream));
    NBR_CUDA(cudaMemcpyAsync(lohi, box.ptr, sizeof(lohi), cudaMemcpyDeviceToHost, stream));
    NBR_CUDA(cudaStreamSynchronize(stream));
    PGrid G;
    double ext[3], longest = 0.0;
    for (int a = 0; a < 3; ++a) {
        G.origin[a] = lohi[a];
        ext[a] = std::max(lohi[3 + a] - lohi[a], 0.0);
        longest = std::max(longest, ext[a]);
    }
    if (!(cell_edge > 0)) {
        // surfaces, not volumes: ~4 points per cell if the points covered the box's largest face evenly
        double area = std::max(std::max(ext[0] * ext[1], ext[0] * ext[2]), ext[1] * ext[2]);
        cell_edge = area > 0 ? sqrt(4.0 * area / (double)ns) : (longest > 0 ? longest / 64.0 : 1.0);
    }
    if (!(cell_edge > 0) || !std::isfinite(cell_edge)) cell_edge = 1.0;
    for (int attempt = 0; attempt < 200; ++attempt) {
        double cells = 1.0;
        bool ok = true;
        for (int a = 0; a < 3; ++a) {
            const double d = floor(ext[a] / cell_edge) + 1.0;
            if (!(d <= 2097152.0)) { ok = false; break; }
            G.dims[a] = (int)d;
            cells *= d;
        }
        if (ok && cells <= 67108864.0) break;
        cell_edge *= 2.0;
    }
    G.g = cell_edge;
    G.inv_g = 1.0 / cell_edge;
    const int64_t nc = (int64_t)G.dims[0] * G.dims[1] * G.dims[2];
    // ---- counting sort of the search points by cell
    Scratch counts, cell, rank, sorted, orig;
    NBR_TRY(counts.alloc(sizeof(uint32_t) * nc, stream));
    NBR_TRY(cell.alloc(sizeof(uint32_t) * ns, stream));
    NBR_TRY(rank.alloc(sizeof(uint32_t) * ns, stream));
    NBR_TRY(sorted.alloc(sizeof(double) * 3 * ns, stream));
    NBR_TRY(orig.alloc(sizeof(int32_t) * ns, stream));
    NBR_CUDA(cudaMemsetAsync(counts.ptr, 0, sizeof(uint32_t) * nc, stream));
    const unsigned pblocks = (unsigned)ceil_div(ns, 256);
    kp_count_kernel<<<pblocks, 256, 0, stream>>>(search, s_dtype, ns, G, counts.as<uint32_t>(), cell.as<uint32_t>(), rank.as<uint32_t>());
    NBR_LAUNCHED();
    NBR_TRY((exclusive_scan<uint32_t, uint32_t>(counts.as<uint32_t>(), counts.as<uint32_t>(), nc, stream)));
    kp_place_kernel<<<pblocks, 256, 0, stream>>>(search, s_dtype, ns, counts.as<uint32_t>(), cell.as<uint32_t>(), rank.as<uint32_t>(),
                                                 sorted.as<double>(), orig.as<int32_t>());
    NBR_LAUNCHED();
    // ---- queries
    int cap = 128;
    while (cap < 4 * k && cap < KP_CAP) cap <<= 1;           // 128 records for k <= 32, 256 for k <= 64, 512 above
    const size_t smem = sizeof(KpRec) * (size_t)cap * KP_WARPS;
    const size_t smem_max = sizeof(KpRec) * (size_t)KP_CAP * KP_WARPS;
    static std::atomic<uint64_t> configured{0};
    if (first_use_on_device(configured)) {
        NBR_CUDA(cudaFuncSetAttribute(knn_points_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max));
        NBR_CUDA(cudaFuncSetAttribute(knn_points_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max));
    }
    const int blocks = (int)std::min<int64_t>(ceil_div(nq, KP_WARPS), (int64_t)device_sm_count() * 12);
    if (out_dtype == NBR_F32)
        knn_points_kernel<float><<<blocks, KP_WARPS * 32, smem, stream>>>(G, sorted.as<double>(), orig.as<int32_t>(), counts.as<uint32_t>(), ns,
                                                                          query, q_dtype, nq, k, idx_out, d2_out, kp, (float *)feats,
                                                                          row_stride, col_offset, descriptor_mask, cap);
    else
        knn_points_kernel<double><<<blocks, KP_WARPS * 32, smem, stream>>>(G, sorted.as<double>(), orig.as<int32_t>(), counts.as<uint32_t>(), ns,
                                                                           query, q_dtype, nq, k, idx_out, d2_out, kp, (double *)feats,
                                                                           row_stride, col_offset, descriptor_mask, cap);
    NBR_LAUNCHED();
    return NBR_OK;
}

}  // namespace nbr

extern "C" int nbr_knn_points(const void *search_xyz, int s_dtype, int64_t n_search, const void *query_xyz, int q_dtype,
                              int64_t n_query, int32_t k, double cell_edge, int32_t *idx_out, double *d2_out,
                              const int32_t *ks_host, int32_t n_k, void *feats_out, int out_dtype, int64_t out_row_stride,
                              int32_t col_offset, int32_t descriptor_mask, void *stream)
{
    if (!search_xyz || (n_query > 0 && !query_xyz)) return nbr::fail(NBR_ERR_INVALID, "nbr_knn_points: null argument");
    if ((s_dtype != NBR_F32 && s_dtype != NBR_F64) || (q_dtype != NBR_F32 && q_dtype != NBR_F64))
        return nbr::fail(NBR_ERR_INVALID, "nbr_knn_points: bad dtype");
    if (feats_out && (!ks_host || n_k < 1)) return nbr::fail(NBR_ERR_INVALID, "nbr_knn_points: feats_out needs ks");
    if (feats_out && out_dtype != NBR_F32 && out_dtype != NBR_F64) return nbr::fail(NBR_ERR_INVALID, "nbr_knn_points: bad out_dtype");
    return nbr::knn_points(search_xyz, s_dtype, n_search, query_xyz, q_dtype, n_query, k, cell_edge, idx_out, d2_out, ks_host, n_k,
                           feats_out, out_dtype, out_row_stride, col_offset, descriptor_mask, (cudaStream_t)stream);
}
