// api.cu -- C ABI glue: errors, scratch memory, the radius-feature dispatcher and the whole-path
// drivers that stand in for process_single_core (nimrud/minimal/multiscale.py:27-67).
#include <algorithm>
#include <mutex>
#include <string>
#include <vector>

#include "common.cuh"
#include "lattice.cuh"
#include "mailbox.cuh"
#include "plan.cuh"
#include "radius_rows.cuh"

namespace nbr {

static thread_local std::string g_error;
std::atomic<int64_t> g_launches{0};

void set_error(const std::string &msg) { g_error = msg; }

int fail(int code, const std::string &msg)
{
    g_error = msg;
    return code;
}

int Scratch::alloc(size_t bytes, cudaStream_t s)
{
    release();
    stream = s;
    if (bytes == 0) bytes = 16;
    cudaError_t e = pool_alloc(&ptr, bytes, s);
    if (e != cudaSuccess) {
        ptr = nullptr;
        return fail(NBR_ERR_CUDA, std::string("cudaMallocAsync(") + std::to_string(bytes) + "): " + cudaGetErrorString(e));
    }
    return NBR_OK;
}

void Scratch::release()
{
    if (ptr) cudaFreeAsync(ptr, stream);
    ptr = nullptr;
}

int device_sm_count()
{
    static std::atomic<int> sms_of[64];
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    int sms = sms_of[dev].load(std::memory_order_relaxed);
    if (!sms) {
        if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
        sms_of[dev].store(sms, std::memory_order_relaxed);
    }
    return sms;
}

static std::mutex g_pool_mutex;
static cudaMemPool_t g_pools[64] = {nullptr};

static cudaMemPool_t private_pool(int dev)
{
    std::lock_guard<std::mutex> lock(g_pool_mutex);
    if (g_pools[dev]) return g_pools[dev];
    cudaMemPoolProps props;
    memset(&props, 0, sizeof(props));
    props.allocType = cudaMemAllocationTypePinned;
    props.handleTypes = cudaMemHandleTypeNone;
    props.location.type = cudaMemLocationTypeDevice;
    props.location.id = dev;
    cudaMemPool_t pool = nullptr;
    if (cudaMemPoolCreate(&pool, &props) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    // keep a bounded amount of freed scratch: the path allocates the same sizes on every call
    uint64_t keep = 0;
    const char *env = getenv("NBR_POOL_KEEP_MB");
    if (env) keep = (uint64_t)atof(env) * (1ull << 20);
    else {
        size_t free_b = 0, total_b = 0;
        if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) total_b = 0;
        keep = std::max<uint64_t>(2ull << 30, total_b / 20);
    }
    cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    g_pools[dev] = pool;
    return pool;
}

cudaError_t pool_alloc(void **ptr, size_t bytes, cudaStream_t stream)
{
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    cudaMemPool_t pool = dev >= 0 && dev < 64 ? private_pool(dev) : nullptr;
    if (!pool) return cudaMallocAsync(ptr, bytes, stream);          // no private pool on this driver: the default one, untouched
    return cudaMallocFromPoolAsync(ptr, bytes, pool, stream);
}

bool first_use_on_device(std::atomic<uint64_t> &seen)
{
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return true;      // unknown: configure again
    const uint64_t bit = 1ull << dev;
    return (seen.fetch_or(bit) & bit) == 0;
}

// ---------------------------------------------------------------------------------------------
// phase timing
// ---------------------------------------------------------------------------------------------
static std::atomic<int> g_timing{0};
static std::mutex g_timing_mutex;
struct TimedSpan { int phase; cudaEvent_t start, stop; };
static std::vector<TimedSpan> g_spans;

PhaseTimer::PhaseTimer(int phase_, cudaStream_t stream_) : phase(phase_), stream(stream_)
{
    if (!g_timing.load(std::memory_order_relaxed)) return;
    if (cudaEventCreate(&start) != cudaSuccess || cudaEventCreate(&stop) != cudaSuccess) { start = stop = nullptr; return; }
    cudaEventRecord(start, stream);
}

PhaseTimer::~PhaseTimer()
{
    if (!start) return;
    cudaEventRecord(stop, stream);
    std::lock_guard<std::mutex> lock(g_timing_mutex);
    g_spans.push_back({phase, start, stop});
}

int radius_features_exact(const Lattice *lat, const void *query, int dtype, const uint32_t *perm, int64_t nq,
                          const double *radii, int nr, void *out, int out_dtype, int64_t row_stride, int col_offset,
                          int descriptor_mask, cudaStream_t stream);
int morton_order(const void *xyz, int dtype, int64_t n, const double lohi[6], double cell, uint32_t *perm_out,
                 void *sorted_xyz_out, cudaStream_t stream);
int cell_order(const void *xyz, int dtype, int64_t n, const double lohi[6], const double origin[3],
               const double cell_in[3], uint32_t *perm_out, void *sorted_xyz_out, cudaStream_t stream,
               const GridDev *finest = nullptr, CellOrderInfo *info = nullptr);
int compact_prefix_queries(const uint32_t *perm, const void *sorted, int dtype, int64_t n, int64_t nq, uint32_t *perm_q,
                           void *sorted_q, cudaStream_t stream);
int radius_sets(const Lattice *lat, const void *query, int dtype, int64_t nq, double radius, int64_t *offsets,
                int32_t *indices, cudaStream_t stream);
int halo_wait(Mailbox *M, cudaStream_t stream);

static int check_cloud_dtype(int dtype, const char *who)
{
    if (dtype != NBR_F32 && dtype != NBR_F64) return fail(NBR_ERR_INVALID, std::string(who) + ": dtype must be NBR_F32 or NBR_F64");
    return NBR_OK;
}

int radius_features(const Lattice *lat, const void *query, int dtype, const uint32_t *perm, int64_t nq,
                    const double *radii, int nr, void *out, int out_dtype, int64_t row_stride, int col_offset,
                    int descriptor_mask, int algorithm, cudaStream_t stream)
{
    if (lat->grid.ndim != 3) return fail(NBR_ERR_INVALID, "radius_features: the feature path is 3-D only");
    for (int k = 0; k < nr; ++k)
        if (!(radii[k] >= 0)) return fail(NBR_ERR_INVALID, "radius_features: radii must be >= 0");
    const int ncol = (descriptor_mask & NBR_DESC_EXTENDED) ? NBR_COLS_EXTENDED : NBR_COLS_REFERENCE;
    NBR_TRY(ball_tables_trim());
    NBR_TRY(ball_tables5_trim());
    if (algorithm != 1 && rows_supported(lat->grid.edge, radii, nr)) {
        // 7x7x7 windows go through the lean kernel, wider ones through the interval kernel
        R3Launch r3, r5;
        memset(&r3, 0, sizeof(r3));
        memset(&r5, 0, sizeof(r5));
        std::vector<double> rr;
        std::vector<int> cc;
        for (int k = 0; k < nr; ++k) {
            R3Entry E;
            int rc = NBR_OK;
            if (rows3_entry(lat, radii[k], col_offset + k * ncol, r3.n ? &r3.e[r3.n - 1] : nullptr, &E, &r3.tq, stream, &rc)) {
                r3.e[r3.n++] = E;
                if (r3.n == R3_MAX_ENTRIES) {
                    NBR_TRY(rows3_launch(&r3, query, dtype, perm, nq, out, out_dtype, row_stride, descriptor_mask, stream));
                    r3.n = 0;
                }
                continue;
            }
            NBR_TRY(rc);
            if (rows5_entry(lat, radii[k], col_offset + k * ncol, r5.n ? &r5.e[r5.n - 1] : nullptr, &E, &r5.tq, stream, &rc)) {
                r5.e[r5.n++] = E;
                if (r5.n == R3_MAX_ENTRIES) {
                    NBR_TRY(rows5_launch(&r5, query, dtype, perm, nq, out, out_dtype, row_stride, descriptor_mask, stream));
                    r5.n = 0;
                }
                continue;
            }
            NBR_TRY(rc);
            rr.push_back(radii[k]);
            cc.push_back(col_offset + k * ncol);
        }
        NBR_TRY(rows3_launch(&r3, query, dtype, perm, nq, out, out_dtype, row_stride, descriptor_mask, stream));
        NBR_TRY(rows5_launch(&r5, query, dtype, perm, nq, out, out_dtype, row_stride, descriptor_mask, stream));
        for (size_t base = 0; base < rr.size(); base += RW_MAX_RADII) {
            RowsLaunch launch;
            memset(&launch, 0, sizeof(launch));
            const int n = (int)std::min<size_t>(RW_MAX_RADII, rr.size() - base);
            launch.n_lat = 1;
            launch.lat[0] = lat->dev();
            NBR_TRY(rows_param(lat, rr.data() + base, cc.data() + base, n, &launch.rows[0], stream));
            NBR_TRY(radius_rows_launch(&launch, query, dtype, perm, nq, out, out_dtype, row_stride, descriptor_mask, stream));
        }
        return NBR_OK;
    }
    if (algorithm == 2) return fail(NBR_ERR_UNSUPPORTED, "radius_features: row-interval kernel does not cover this r/e");
    return radius_features_exact(lat, query, dtype, perm, nq, radii, nr, out, out_dtype, row_stride, col_offset,
                                 descriptor_mask, stream);
}

// ------------------------------------------------------------------------------------------------
// whole-path driver.  a Plan holds what depends on the SEARCH cloud only (one lattice per distinct
// edge); plan_run handles one batch of queries: Morton order, one fused launch over all lattices
// (all radii of a lattice share one pass over each query's window).  scales the row kernel cannot take
// (r/e > 9.5) go through the exact kernel.
// ------------------------------------------------------------------------------------------------
// bounding box of a device cloud, copied to the host (synchronises the stream)
static int host_bbox(const void *xyz, int dtype, int64_t n, double *lohi_host, cudaStream_t stream)
{
    Scratch box;
    NBR_TRY(box.alloc(sizeof(double) * 6, stream));
    {
        PhaseTimer t(PHASE_BBOX, stream);
        NBR_TRY(bbox(xyz, dtype, n, 3, box.as<double>(), stream));
    }
    NBR_CUDA(cudaMemcpyAsync(lohi_host, box.ptr, sizeof(double) * 6, cudaMemcpyDeviceToHost, stream));
    NBR_CUDA(cudaStreamSynchronize(stream));
    return NBR_OK;
}

static int brick_origin(const double *lohi, const double *local_box, double finest, double origin[3], GridDev *gdev = nullptr);
int order_queries(const void *query, int q_dtype, int64_t nq, const double *qbox_known, double finest,
                  const double *origin, Scratch &perm, Scratch &sorted, cudaStream_t stream,
                  const GridDev *finest_grid = nullptr, CellOrderInfo *info = nullptr);

int plan_create(Plan **out, const void *search, int s_dtype, int64_t ns, const double *edges, const double *radii,
                int n_scales, int descriptor_mask, const double *global_lohi, const double *known_local_box,
                cudaStream_t stream, const void *search2, int64_t ns2, Mailbox *mailbox, const CellOrderInfo *order)
{
    NBR_TRY(check_cloud_dtype(s_dtype, "multiscale_features"));
    if (n_scales < 0) return fail(NBR_ERR_INVALID, "multiscale_features: negative number of scales");
    if (ns + ns2 < 2 && !mailbox) return fail(NBR_ERR_TOO_FEW_POINTS, "need at least 2 points to define a voxel grid");
    if (!search || (n_scales > 0 && (!edges || !radii))) return fail(NBR_ERR_INVALID, "multiscale_features: null argument");
    for (int s = 0; s < n_scales; ++s) {
        if (!(edges[s] > 0)) return fail(NBR_ERR_INVALID, "multiscale_features: edge lengths must be > 0");
        if (!(radii[s] >= 0)) return fail(NBR_ERR_INVALID, "multiscale_features: radii must be >= 0");
    }
    Plan *P = new Plan();
    P->n_scales = n_scales;
    P->descriptor_mask = descriptor_mask;
    P->ncol = (descriptor_mask & NBR_DESC_EXTENDED) ? NBR_COLS_EXTENDED : NBR_COLS_REFERENCE;
    P->edges.assign(edges, edges + n_scales);
    P->radii.assign(radii, radii + n_scales);
    // bounding box of the search cloud given to this call.  with global_lohi (multi-GPU: the all-reduced
    // box) the lattices are ANCHORED on the global box but their directories only cover this local box.
    int rc = NBR_OK;
    if (known_local_box) {
        std::copy(known_local_box, known_local_box + 6, P->local_box);
    } else {
        rc = host_bbox(search, s_dtype, ns, P->local_box, stream);
    }
    if (rc) { delete P; return rc; }
    const double *lohi = global_lohi ? global_lohi : P->local_box;
    for (int s = 0; s < n_scales; ++s) {
        size_t gi = 0;
        while (gi < P->groups.size() && P->groups[gi].edge != edges[s]) ++gi;
        if (gi == P->groups.size()) P->groups.push_back({edges[s], nullptr, {}});
        P->groups[gi].scales.push_back(s);
        P->finest = s == 0 ? edges[s] : std::min(P->finest, edges[s]);
    }
    {
        PhaseTimer t(PHASE_INDEX, stream);
        // all lattices of the call in one pass over the points (batches of LATTICE_BATCH)
        for (size_t base = 0; !rc && base < P->groups.size(); base += LATTICE_BATCH) {
            const int nb = (int)std::min<size_t>(LATTICE_BATCH, P->groups.size() - base);
            nbr_grid grids[LATTICE_BATCH];
            Lattice *made[LATTICE_BATCH] = {nullptr};
            for (int k = 0; !rc && k < nb; ++k) rc = grid_from_bbox(lohi, lohi + 3, P->groups[base + k].edge, 3, &grids[k]);
            if (!rc) rc = lattices_create_batch(made, nb, grids, search, s_dtype, ns, stream, global_lohi ? P->local_box : nullptr, search2, ns2, mailbox, order);
            if (!rc) for (int k = 0; k < nb; ++k) P->groups[base + k].lat = made[k];
        }
    }
    if (!rc && n_scales > 0) rc = brick_origin(lohi, global_lohi ? P->local_box : nullptr, P->finest, P->order_origin);
    if (rc) { delete P; return rc; }
    *out = P;
    return NBR_OK;
}

// features of one batch of queries that is already in a coherent order: row i of `sorted` is query
// perm[i] (perm == NULL: identity) and its features go to row perm[i] of `out`
// dests (feature all-gather, nbr_tile_step_gather): `out` is this rank's share of the result; the rows also go to every
// peer's staging buffer -- from inside the fused kernel when one launch owns whole rows, otherwise as a push of the
// finished share
int plan_run_sorted(const Plan *P, const void *sorted, int q_dtype, const uint32_t *perm, int64_t nq, void *out,
                    int out_dtype, cudaStream_t stream, const RowDests *dests = nullptr)
{
    if (out_dtype != NBR_F32 && out_dtype != NBR_F64) return fail(NBR_ERR_INVALID, "multiscale_features: bad out_dtype");
    if (nq <= 0 || P->n_scales == 0) return NBR_OK;
    const int ncol = P->ncol;
    const int64_t row_stride = (int64_t)ncol * P->n_scales;
    NBR_TRY(ball_tables_trim());
    NBR_TRY(ball_tables5_trim());
    PhaseTimer tm(PHASE_FEATURES, stream);
    RowsLaunch launch;
    memset(&launch, 0, sizeof(launch));
    auto flush = [&]() -> int {
        if (launch.n_lat == 0) return NBR_OK;
        int r = radius_rows_launch(&launch, sorted, q_dtype, perm, nq, out, out_dtype, row_stride, P->descriptor_mask, stream);
        memset(&launch, 0, sizeof(launch));
        return r;
    };
    R3Launch r3, r5;
    memset(&r3, 0, sizeof(r3));
    memset(&r5, 0, sizeof(r5));
    bool only_rows3 = true;                   // every scale so far is an entry of the pending rows3 launch
    for (auto &g : P->groups) {
        std::vector<double> rr;
        std::vector<int> cc;
        for (int s : g.scales) {
            R3Entry E;
            int rc = NBR_OK;
            if (rows3_entry(g.lat, P->radii[s], s * ncol, r3.n ? &r3.e[r3.n - 1] : nullptr, &E, &r3.tq, stream, &rc)) {
                r3.e[r3.n++] = E;
                if (r3.n == R3_MAX_ENTRIES && P->n_scales > R3_MAX_ENTRIES) {
                    NBR_TRY(rows3_launch(&r3, sorted, q_dtype, perm, nq, out, out_dtype, row_stride, P->descriptor_mask, stream));
                    r3.n = 0;
                    only_rows3 = false;
                }
                continue;
            }
            NBR_TRY(rc);
            only_rows3 = false;
            if (rows5_entry(g.lat, P->radii[s], s * ncol, r5.n ? &r5.e[r5.n - 1] : nullptr, &E, &r5.tq, stream, &rc)) {
                r5.e[r5.n++] = E;
                if (r5.n == R3_MAX_ENTRIES) {
                    NBR_TRY(rows5_launch(&r5, sorted, q_dtype, perm, nq, out, out_dtype, row_stride, P->descriptor_mask, stream));
                    r5.n = 0;
                }
                continue;
            }
            NBR_TRY(rc);
            if (rows_supported(g.edge, &P->radii[s], 1)) { rr.push_back(P->radii[s]); cc.push_back(s * ncol); }
            else
                NBR_TRY(radius_features_exact(g.lat, sorted, q_dtype, perm, nq, &P->radii[s], 1, out, out_dtype, row_stride,
                                              s * ncol, P->descriptor_mask, stream));
        }
        for (size_t base = 0; base < rr.size(); base += RW_MAX_RADII) {
            const int n = (int)std::min<size_t>(RW_MAX_RADII, rr.size() - base);
            launch.lat[launch.n_lat] = g.lat->dev();
            NBR_TRY(rows_param(g.lat, rr.data() + base, cc.data() + base, n, &launch.rows[launch.n_lat], stream));
            ++launch.n_lat;
            if (launch.n_lat == RW_MAX_LATTICES) NBR_TRY(flush());
        }
    }
    const char *gather_env = getenv("NBR_GATHER");
    const bool peer_copies = gather_env && std::string(gather_env) == "copy";
    const bool fused = dests && dests->n > 1 && only_rows3 && !peer_copies && ((uintptr_t)out & 15) == 0 &&
                       rows3_owns_rows(r3.n, row_stride, out_dtype, P->descriptor_mask);
    NBR_TRY(rows3_launch(&r3, sorted, q_dtype, perm, nq, out, out_dtype, row_stride, P->descriptor_mask, stream, fused ? dests : nullptr));
    NBR_TRY(rows5_launch(&r5, sorted, q_dtype, perm, nq, out, out_dtype, row_stride, P->descriptor_mask, stream));
    NBR_TRY(flush());
    if (dests && !fused) NBR_TRY(gather_push_rows(dests, out, nq, (size_t)row_stride * (out_dtype == NBR_F32 ? 4 : 8), stream));
    return NBR_OK;
}

// corner of brick (0,0,0) of the lattice with edge `finest` that plan_create would build for this box
static int brick_origin(const double *lohi, const double *local_box, double finest, double origin[3], GridDev *gdev)
{
    nbr_grid grid;
    GridDev d;
    NBR_TRY(grid_from_bbox(lohi, lohi + 3, finest, 3, &grid));
    NBR_TRY(grid_to_dev(&grid, &d, local_box));
    for (int a = 0; a < 3; ++a) origin[a] = d.minc[a] + (double)d.cell_lo[a] * finest;
    if (gdev) *gdev = d;
    return NBR_OK;
}

// spatially coherent order of a cloud + the cloud in that order.  default: counting sort on cells shaped and
// aligned like the bricks of the finest lattice (origin = its minimum corner); NBR_ORDER=morton: Z-curve of
// cells of 4 finest voxels, radix sorted
int order_queries(const void *query, int q_dtype, int64_t nq, const double *qbox_known, double finest,
                  const double *origin, Scratch &perm, Scratch &sorted, cudaStream_t stream, const GridDev *finest_grid,
                  CellOrderInfo *info)
{
    NBR_TRY(perm.alloc(sizeof(uint32_t) * nq, stream));
    NBR_TRY(sorted.alloc((size_t)nq * 3 * (q_dtype == NBR_F32 ? 4 : 8), stream));
    double qbox[6];
    if (qbox_known) std::copy(qbox_known, qbox_known + 6, qbox);
    else NBR_TRY(host_bbox(query, q_dtype, nq, qbox, stream));
    PhaseTimer t(PHASE_ORDER, stream);
    static const bool use_morton = getenv("NBR_ORDER") && std::string(getenv("NBR_ORDER")) == "morton";
    if (use_morton) return morton_order(query, q_dtype, nq, qbox, 4.0 * finest, perm.as<uint32_t>(), sorted.ptr, stream);
    const double cell[3] = {BRICK_X * finest, BRICK_Y * finest, BRICK_Z * finest};
    const double org[3] = {origin ? origin[0] : qbox[0], origin ? origin[1] : qbox[1], origin ? origin[2] : qbox[2]};
    return cell_order(query, q_dtype, nq, qbox, org, cell, perm.as<uint32_t>(), sorted.ptr, stream, finest_grid, info);
}

// features of one batch of queries in arbitrary order -> rows [0, nq) of `out`
int plan_run(const Plan *P, const void *query, int q_dtype, int64_t nq, const double *qbox_known, void *out,
             int out_dtype, cudaStream_t stream)
{
    NBR_TRY(check_cloud_dtype(q_dtype, "multiscale_features"));
    if (nq <= 0 || P->n_scales == 0) return NBR_OK;
    if (!query || !out) return fail(NBR_ERR_INVALID, "multiscale_features: null argument");
    Scratch perm, sorted;
    NBR_TRY(order_queries(query, q_dtype, nq, qbox_known, P->finest, P->order_origin, perm, sorted, stream));
    return plan_run_sorted(P, sorted.ptr, q_dtype, perm.as<uint32_t>(), nq, out, out_dtype, stream);
}

int plan_voxel_counts(const Plan *P, int64_t *n_voxels_host)
{
    for (auto &g : P->groups) {
        int64_t nv = 0;
        NBR_TRY(lattice_counts(g.lat, &nv, nullptr));
        for (int s : g.scales) n_voxels_host[s] = nv;
    }
    return NBR_OK;
}

int multiscale_features(const void *query, int q_dtype, int64_t nq, const void *search, int s_dtype, int64_t ns,
                        const double *edges, const double *radii, int n_scales, void *out, int out_dtype,
                        int descriptor_mask, const double *global_lohi, int64_t *n_voxels_host, cudaStream_t stream)
{
    if (nq < 0) return fail(NBR_ERR_INVALID, "multiscale_features: negative size");
    NBR_TRY(check_cloud_dtype(q_dtype, "multiscale_features"));
    NBR_TRY(check_cloud_dtype(s_dtype, "multiscale_features"));
    if (ns < 2) return fail(NBR_ERR_TOO_FEW_POINTS, "need at least 2 points to define a voxel grid");
    if (!search) return fail(NBR_ERR_INVALID, "multiscale_features: null argument");
    // query cloud == search cloud, or its first nq points (tile + halo, multi-GPU): order the search cloud once
    // and build the lattices from the ordered copy too (coalesced directory / pool updates); the lattices do
    // not depend on the order of the points
    const bool same = query == search && nq <= ns && nq > 0 && q_dtype == s_dtype && n_scales > 0 && out;
    Plan *P = nullptr;
    int rc;
    if (same) {
        double box[6], finest = 0.0;
        for (int s = 0; s < n_scales; ++s) {
            if (!edges || !(edges[s] > 0)) return fail(NBR_ERR_INVALID, "multiscale_features: edge lengths must be > 0");
            finest = s == 0 ? edges[s] : std::min(finest, edges[s]);
        }
        NBR_TRY(host_bbox(search, s_dtype, ns, box, stream));
        Scratch perm, sorted, perm_q, sorted_q;
        double origin[3];
        GridDev fgrid;
        CellOrderInfo order;
        NBR_TRY(brick_origin(global_lohi ? global_lohi : box, global_lohi ? box : nullptr, finest, origin, &fgrid));
        NBR_TRY(order_queries(search, s_dtype, ns, box, finest, origin, perm, sorted, stream, &fgrid, &order));
        NBR_TRY(plan_create(&P, sorted.ptr, s_dtype, ns, edges, radii, n_scales, descriptor_mask, global_lohi, box, stream, nullptr, 0,
                            nullptr, &order));
        if (nq < ns) {
            rc = perm_q.alloc(sizeof(uint32_t) * nq, stream);
            if (!rc) rc = sorted_q.alloc((size_t)nq * 3 * (q_dtype == NBR_F32 ? 4 : 8), stream);
            if (!rc) {
                PhaseTimer t(PHASE_ORDER, stream);
                rc = compact_prefix_queries(perm.as<uint32_t>(), sorted.ptr, q_dtype, ns, nq, perm_q.as<uint32_t>(), sorted_q.ptr, stream);
            }
            if (!rc) rc = plan_run_sorted(P, sorted_q.ptr, q_dtype, perm_q.as<uint32_t>(), nq, out, out_dtype, stream);
        } else {
            rc = plan_run_sorted(P, sorted.ptr, q_dtype, perm.as<uint32_t>(), nq, out, out_dtype, stream);
        }
        if (rc == NBR_OK && n_voxels_host) rc = plan_voxel_counts(P, n_voxels_host);
        delete P;
        return rc;
    }
    NBR_TRY(plan_create(&P, search, s_dtype, ns, edges, radii, n_scales, descriptor_mask, global_lohi, nullptr, stream));
    rc = plan_run(P, query, q_dtype, nq, nullptr, out, out_dtype, stream);
    if (rc == NBR_OK && n_voxels_host) rc = plan_voxel_counts(P, n_voxels_host);
    delete P;
    return rc;
}

}  // namespace nbr

using namespace nbr;

extern "C" const char *nbr_last_error(void) { return g_error.c_str(); }
extern "C" int nbr_version(void) { return 200; }

// releases the scratch the library's private pool keeps cached on the current device (synchronises the device)
extern "C" int nbr_trim_memory(void)
{
    int dev = 0;
    NBR_CUDA(cudaGetDevice(&dev));
    NBR_CUDA(cudaDeviceSynchronize());
    cudaMemPool_t pool = nullptr;
    {
        std::lock_guard<std::mutex> lock(g_pool_mutex);
        pool = dev >= 0 && dev < 64 ? g_pools[dev] : nullptr;
    }
    if (pool) NBR_CUDA(cudaMemPoolTrimTo(pool, 0));
    return NBR_OK;
}
extern "C" int64_t nbr_kernel_launches(void) { return g_launches.load(); }

extern "C" void nbr_timing_enable(int on) { g_timing.store(on ? 1 : 0); }

extern "C" int nbr_timing_read(double *ms_out)
{
    if (!ms_out) return fail(NBR_ERR_INVALID, "nbr_timing_read: null argument");
    for (int i = 0; i < PHASE_COUNT; ++i) ms_out[i] = 0.0;
    std::lock_guard<std::mutex> lock(g_timing_mutex);
    for (auto &sp : g_spans) {
        float ms = 0.f;
        if (cudaEventSynchronize(sp.stop) == cudaSuccess && cudaEventElapsedTime(&ms, sp.start, sp.stop) == cudaSuccess)
            ms_out[sp.phase] += ms;
        cudaEventDestroy(sp.start);
        cudaEventDestroy(sp.stop);
    }
    g_spans.clear();
    return NBR_OK;
}

extern "C" int nbr_radius_features(const nbr_lattice *lattice, const void *query_xyz, int dtype, int64_t n_query,
                                   const double *radii_host, int32_t n_radii, void *out, int out_dtype,
                                   int64_t out_row_stride, int32_t col_offset, int32_t descriptor_mask,
                                   int32_t algorithm, void *stream)
{
    if (!lattice || !query_xyz || !radii_host || !out) return fail(NBR_ERR_INVALID, "nbr_radius_features: null argument");
    NBR_TRY(check_cloud_dtype(dtype, "nbr_radius_features"));
    if (out_dtype != NBR_F32 && out_dtype != NBR_F64) return fail(NBR_ERR_INVALID, "nbr_radius_features: bad out_dtype");
    return radius_features(reinterpret_cast<const Lattice *>(lattice), query_xyz, dtype, nullptr, n_query, radii_host,
                           n_radii, out, out_dtype, out_row_stride, col_offset, descriptor_mask, algorithm,
                           (cudaStream_t)stream);
}

extern "C" int nbr_radius_sets(const nbr_lattice *lattice, const void *query_xyz, int dtype, int64_t n_query,
                               double radius, int64_t *offsets, int32_t *indices, void *stream)
{
    if (!lattice || !query_xyz || !offsets) return fail(NBR_ERR_INVALID, "nbr_radius_sets: null argument");
    NBR_TRY(check_cloud_dtype(dtype, "nbr_radius_sets"));
    return radius_sets(reinterpret_cast<const Lattice *>(lattice), query_xyz, dtype, n_query, radius, offsets, indices,
                       (cudaStream_t)stream);
}

extern "C" int nbr_multiscale_features(const void *query_xyz, int q_dtype, int64_t n_query, const void *search_xyz,
                                       int s_dtype, int64_t n_search, const double *edges_host,
                                       const double *radii_host, int32_t n_scales, void *out, int out_dtype,
                                       int32_t descriptor_mask, const double *global_lohi_host,
                                       int64_t *n_voxels_host, void *stream)
{
    return multiscale_features(query_xyz, q_dtype, n_query, search_xyz, s_dtype, n_search, edges_host, radii_host,
                               n_scales, out, out_dtype, descriptor_mask, global_lohi_host, n_voxels_host,
                               (cudaStream_t)stream);
}

// ---- multi-GPU tile path: the tile's points are ordered while the halo exchange is still in flight, then the
// lattices are built from the ordered tile plus the received halo points and the tile's points are the queries
extern "C" int nbr_brick_origin(const double *global_lohi_host, const double *local_lohi_host, double finest_edge,
                                double *origin_out)
{
    if (!global_lohi_host || !origin_out) return fail(NBR_ERR_INVALID, "nbr_brick_origin: null argument");
    return brick_origin(global_lohi_host, local_lohi_host, finest_edge, origin_out);
}

extern "C" int nbr_order_cloud(const void *xyz, int dtype, int64_t n, const double *lohi_host, const double *origin_host,
                               double finest_edge, uint32_t *perm_out, void *sorted_out, void *stream)
{
    if (!xyz || !lohi_host || !origin_host || !perm_out || !sorted_out) return fail(NBR_ERR_INVALID, "nbr_order_cloud: null argument");
    NBR_TRY(check_cloud_dtype(dtype, "nbr_order_cloud"));
    if (!(finest_edge > 0)) return fail(NBR_ERR_INVALID, "nbr_order_cloud: edge must be > 0");
    if (n <= 0) return NBR_OK;
    PhaseTimer t(PHASE_ORDER, (cudaStream_t)stream);
    const double cell[3] = {BRICK_X * finest_edge, BRICK_Y * finest_edge, BRICK_Z * finest_edge};
    return cell_order(xyz, dtype, n, lohi_host, origin_host, cell, perm_out, sorted_out, (cudaStream_t)stream);
}

extern "C" int nbr_multiscale_features_tile(const void *sorted_xyz, const uint32_t *perm, int dtype, int64_t n,
                                            const void *halo_xyz, int64_t n_halo, const double *local_lohi_host,
                                            const double *global_lohi_host, const double *edges_host,
                                            const double *radii_host, int32_t n_scales, void *out, int out_dtype,
                                            int32_t descriptor_mask, int64_t *n_voxels_host, void *stream)
{
    if (!sorted_xyz || !perm || !local_lohi_host || !global_lohi_host || !out || (n_halo > 0 && !halo_xyz))
        return fail(NBR_ERR_INVALID, "nbr_multiscale_features_tile: null argument");
    NBR_TRY(check_cloud_dtype(dtype, "nbr_multiscale_features_tile"));
    if (n <= 0 || n_scales <= 0) return NBR_OK;
    Plan *P = nullptr;
    NBR_TRY(plan_create(&P, sorted_xyz, dtype, n, edges_host, radii_host, n_scales, descriptor_mask, global_lohi_host,
                        local_lohi_host, (cudaStream_t)stream, halo_xyz, n_halo));
    int rc = plan_run_sorted(P, sorted_xyz, dtype, perm, n, out, out_dtype, (cudaStream_t)stream);
    if (rc == NBR_OK && n_voxels_host) rc = plan_voxel_counts(P, n_voxels_host);
    delete P;
    return rc;
}

// same, the halo comes from this rank's mailbox (mailbox.cu): the wait for the peers' pushes is stream-ordered
// inside the lattice build and the number of halo points never visits the host
extern "C" int nbr_multiscale_features_tile_mb(const void *sorted_xyz, const uint32_t *perm, int dtype, int64_t n,
                                               nbr_mailbox *mailbox, const double *local_lohi_host,
                                               const double *global_lohi_host, const double *edges_host,
                                               const double *radii_host, int32_t n_scales, void *out, int out_dtype,
                                               int32_t descriptor_mask, int64_t *n_voxels_host, void *stream)
{
    if (!mailbox || !local_lohi_host || !global_lohi_host || (n > 0 && (!sorted_xyz || !perm || !out)))
        return fail(NBR_ERR_INVALID, "nbr_multiscale_features_tile_mb: null argument");
    NBR_TRY(check_cloud_dtype(dtype, "nbr_multiscale_features_tile_mb"));
    Mailbox *M = reinterpret_cast<Mailbox *>(mailbox);
    if (n <= 0 || n_scales <= 0) return halo_wait(M, (cudaStream_t)stream);      // an empty tile still drains its mailbox
    Plan *P = nullptr;
    NBR_TRY(plan_create(&P, sorted_xyz, dtype, n, edges_host, radii_host, n_scales, descriptor_mask, global_lohi_host,
                        local_lohi_host, (cudaStream_t)stream, nullptr, 0, M));
    int rc = plan_run_sorted(P, sorted_xyz, dtype, perm, n, out, out_dtype, (cudaStream_t)stream);
    if (rc == NBR_OK && n_voxels_host) rc = plan_voxel_counts(P, n_voxels_host);
    delete P;
    return rc;
}

// destinations of a rank's rows in the staging buffers of the peers (rows [row_offset, ...) of `total`).
// NBR_ERR_CAPACITY if one of the buffers cannot hold `total` rows and their row numbers
static int gather_dests(const Mailbox *M, int64_t row_offset, int64_t total, size_t row_bytes, RowDests *D)
{
    memset(D, 0, sizeof(*D));
    D->n = M->world;
    D->self = M->rank;
    D->row_offset = row_offset;
    if (row_bytes % 16 != 0) return fail(NBR_ERR_UNSUPPORTED, "feature gather: rows must be multiples of 16 bytes");
    const size_t need = gather_bytes_needed(total, row_bytes), perm_at = gather_perm_offset(total, row_bytes);
    for (int r = 0; r < M->world; ++r) {
        if (M->world > 1 && total > 0 && (!M->gather_peer[r] || need > M->gather_peer_bytes[r]))
            return fail(NBR_ERR_CAPACITY, "feature gather: the staging buffers are too small for the rows of all ranks");
        D->base[r] = M->gather_peer[r];
        D->perm[r] = M->gather_peer[r] ? reinterpret_cast<uint32_t *>(M->gather_peer[r] + perm_at) : nullptr;
    }
    return NBR_OK;
}

// nbr_multiscale_features_tile_mb whose rows also go to the staging buffer of every other rank (step-wise form of
// nbr_tile_step_gather: the caller finishes with nbr_gather_finish -- or, several tiles on one device in one process,
// relies on stream order -- and nbr_gather_unpermute).  out_all: the result for the rows of ALL ranks
extern "C" int nbr_multiscale_features_tile_mb_gather(const void *sorted_xyz, const uint32_t *perm, int dtype, int64_t n,
                                                      nbr_mailbox *mailbox, const double *local_lohi_host,
                                                      const double *global_lohi_host, const double *edges_host,
                                                      const double *radii_host, int32_t n_scales, void *out_all, int out_dtype,
                                                      int32_t descriptor_mask, int64_t row_offset, int64_t total_rows,
                                                      int64_t *n_voxels_host, void *stream)
{
    if (!mailbox || !local_lohi_host || !global_lohi_host || (n > 0 && (!sorted_xyz || !perm || !out_all)) || row_offset < 0 ||
        total_rows < row_offset + n)
        return fail(NBR_ERR_INVALID, "nbr_multiscale_features_tile_mb_gather: bad argument");
    NBR_TRY(check_cloud_dtype(dtype, "nbr_multiscale_features_tile_mb_gather"));
    if (out_dtype != NBR_F32 && out_dtype != NBR_F64) return fail(NBR_ERR_INVALID, "nbr_multiscale_features_tile_mb_gather: bad out_dtype");
    Mailbox *M = reinterpret_cast<Mailbox *>(mailbox);
    if (n <= 0 || n_scales <= 0) return halo_wait(M, (cudaStream_t)stream);
    const int ncol = (descriptor_mask & NBR_DESC_EXTENDED) ? NBR_COLS_EXTENDED : NBR_COLS_REFERENCE;
    const size_t row_bytes = (size_t)ncol * n_scales * (out_dtype == NBR_F32 ? 4 : 8);
    RowDests D;
    NBR_TRY(gather_dests(M, row_offset, total_rows, row_bytes, &D));
    Plan *P = nullptr;
    NBR_TRY(plan_create(&P, sorted_xyz, dtype, n, edges_host, radii_host, n_scales, descriptor_mask, global_lohi_host,
                        local_lohi_host, (cudaStream_t)stream, nullptr, 0, M));
    int rc = plan_run_sorted(P, sorted_xyz, dtype, perm, n, (unsigned char *)out_all + (size_t)row_offset * row_bytes, out_dtype,
                             (cudaStream_t)stream, &D);
    if (rc == NBR_OK && n_voxels_host) rc = plan_voxel_counts(P, n_voxels_host);
    delete P;
    return rc;
}

extern "C" int nbr_gather_finish(nbr_mailbox *mailbox, void *stream)
{
    if (!mailbox) return fail(NBR_ERR_INVALID, "nbr_gather_finish: null argument");
    return gather_finish(reinterpret_cast<Mailbox *>(mailbox), (cudaStream_t)stream);
}

extern "C" int nbr_gather_unpermute(nbr_mailbox *mailbox, const int64_t *row_offsets_host, int64_t row_bytes, void *out_all, void *stream)
{
    if (!mailbox || !row_offsets_host || row_bytes < 0 || !out_all) return fail(NBR_ERR_INVALID, "nbr_gather_unpermute: bad argument");
    return gather_unpermute(reinterpret_cast<Mailbox *>(mailbox), row_offsets_host, (size_t)row_bytes, out_all, (cudaStream_t)stream);
}

// one whole step of a rank in ONE call: box table (the step's only host synchronisation) -> halo push -> query
// order of the tile -> lattices from tile + mailbox -> features.  nothing but C++ runs between the synchronisation
// and the next launches, so the GPU idles for microseconds there, not for an interpreter's worth of time.
// h = max_s(r_s + e_s / 2) (+ slack): every voxel centre within r_s of a query of the tile holds a point within h
// per axis of the tile's box.  boxes_host_out (optional): [world][8] = lo, hi, n_points, 0 of every tile.
//
// tile_step_plan: everything up to the lattices.  *P_out stays NULL for an empty tile (its mailbox is drained).
namespace nbr {
int tile_step_plan(Mailbox *M, const void *xyz, int dtype, int64_t n, const double *edges_host, const double *radii_host,
                   int32_t n_scales, int32_t descriptor_mask, double *boxes_host_out, Scratch &perm, Scratch &sorted, Plan **P_out,
                   cudaStream_t s)
{
    *P_out = nullptr;
    for (int k = 0; k < n_scales; ++k)                     // before the first collective step: a bad argument must not strand the peers
        if (!edges_host || !radii_host || !(edges_host[k] > 0) || !(radii_host[k] >= 0))
            return fail(NBR_ERR_INVALID, "nbr_tile_step: edge lengths must be > 0, radii >= 0");
    nbr_mailbox *mailbox = reinterpret_cast<nbr_mailbox *>(M);
    void *stream = (void *)s;
    double boxes[MB_MAX_WORLD][8];
    NBR_TRY(nbr_tile_box_publish(mailbox, xyz, dtype, n, stream));
    NBR_TRY(nbr_tile_boxes_wait(mailbox, &boxes[0][0], stream));
    if (boxes_host_out) memcpy(boxes_host_out, boxes, sizeof(double) * 8 * M->world);
    double h = 0.0, finest = 0.0;
    for (int k = 0; k < n_scales; ++k) {
        h = std::max(h, radii_host[k] + edges_host[k] / 2);
        finest = k == 0 ? edges_host[k] : std::min(finest, edges_host[k]);
    }
    h *= 1.0 + 1e-6;
    NBR_TRY(nbr_halo_push(mailbox, xyz, dtype, n, &boxes[0][0], h, stream));
    if (n == 0 || n_scales == 0) return halo_wait(M, s);                    // an empty tile still drains its mailbox
    double glob[6] = {INFINITY, INFINITY, INFINITY, -INFINITY, -INFINITY, -INFINITY};
    for (int r = 0; r < M->world; ++r)
        if (boxes[r][6] > 0)
            for (int a = 0; a < 3; ++a) {
                glob[a] = std::min(glob[a], boxes[r][a]);
                glob[3 + a] = std::max(glob[3 + a], boxes[r][3 + a]);
            }
    double local[6], origin[3];
    const double *mine = boxes[M->rank];
    for (int a = 0; a < 3; ++a) {
        local[a] = std::max(mine[a] - h, glob[a]);
        local[3 + a] = std::min(mine[3 + a] + h, glob[3 + a]);
    }
    GridDev fgrid;
    CellOrderInfo order;
    NBR_TRY(brick_origin(glob, local, finest, origin, &fgrid));
    NBR_TRY(order_queries(xyz, dtype, n, mine, finest, origin, perm, sorted, s, &fgrid, &order));
    return plan_create(P_out, sorted.ptr, dtype, n, edges_host, radii_host, n_scales, descriptor_mask, glob, local, s, nullptr, 0, M,
                       &order);
}
}  // namespace nbr

// nbr_tile_step with the feature all-gather inside: out_all receives the rows of ALL ranks (rank order, every share in
// its tile's own order).  the fused kernel stores every finished row into the staging buffers of the peers as it goes,
// a stream-ordered signal + wait follows, then the staged rows of the peers are put in place.
// row_offsets_host[world + 1]: first row of every rank's share
extern "C" int nbr_tile_step_gather(nbr_mailbox *mailbox, const void *xyz, int dtype, int64_t n, const double *edges_host,
                                    const double *radii_host, int32_t n_scales, void *out_all, int64_t out_rows_capacity,
                                    int out_dtype, int32_t descriptor_mask, double *boxes_host_out, int64_t *n_voxels_host,
                                    int64_t *row_offsets_host, void *stream)
{
    Mailbox *M = reinterpret_cast<Mailbox *>(mailbox);
    if (!M || n < 0 || n_scales < 0 || (n_scales > 0 && (!edges_host || !radii_host)) || (n > 0 && !xyz) || out_rows_capacity < 0)
        return fail(NBR_ERR_INVALID, "nbr_tile_step_gather: bad argument");
    NBR_TRY(check_cloud_dtype(dtype, "nbr_tile_step_gather"));
    if (out_dtype != NBR_F32 && out_dtype != NBR_F64) return fail(NBR_ERR_INVALID, "nbr_tile_step_gather: bad out_dtype");
    cudaStream_t s = (cudaStream_t)stream;
    Scratch perm, sorted;
    Plan *P = nullptr;
    double boxes[MB_MAX_WORLD][8];
    NBR_TRY(tile_step_plan(M, xyz, dtype, n, edges_host, radii_host, n_scales, descriptor_mask, &boxes[0][0], perm, sorted, &P, s));
    if (boxes_host_out) memcpy(boxes_host_out, boxes, sizeof(double) * 8 * M->world);
    const int ncol = (descriptor_mask & NBR_DESC_EXTENDED) ? NBR_COLS_EXTENDED : NBR_COLS_REFERENCE;
    const size_t row_bytes = (size_t)ncol * n_scales * (out_dtype == NBR_F32 ? 4 : 8);
    RowDests D;
    int64_t offs[MB_MAX_WORLD + 1];
    offs[0] = 0;
    for (int r = 0; r < M->world; ++r) offs[r + 1] = offs[r] + (int64_t)boxes[r][6];
    const int64_t total = offs[M->world];
    if (row_offsets_host) std::copy(offs, offs + M->world + 1, row_offsets_host);
    // every rank sees the same box table and the staging buffers have one size, so a capacity error of the STAGING
    // buffers is raised on every rank or on none; the step's halo exchange is complete by now (tile_step_plan),
    // nothing is left half-done for the peers.  the result's capacity is the caller's own business: ranks that pass
    // differently sized results would disagree here, so it is checked last and reported the same way
    int rc = gather_dests(M, offs[M->rank], total, row_bytes, &D);
    if (!rc && (out_rows_capacity < total || (total > 0 && row_bytes > 0 && !out_all)))
        rc = fail(NBR_ERR_CAPACITY, "nbr_tile_step_gather: the result holds fewer rows than all ranks produce");
    if (!rc && P && (int64_t)boxes[M->rank][6] != n) rc = fail(NBR_ERR_INVALID, "nbr_tile_step_gather: box table and tile disagree");
    if (!rc && P)
        rc = plan_run_sorted(P, sorted.ptr, dtype, perm.as<uint32_t>(), n, (unsigned char *)out_all + (size_t)offs[M->rank] * row_bytes,
                             out_dtype, s, &D);
    if (!rc && P && n_voxels_host) rc = plan_voxel_counts(P, n_voxels_host);
    delete P;
    if (!rc) rc = gather_finish(M, s);
    if (!rc) rc = gather_unpermute(M, offs, row_bytes, out_all, s);
    return rc;
}

extern "C" int nbr_tile_step(nbr_mailbox *mailbox, const void *xyz, int dtype, int64_t n, const double *edges_host,
                             const double *radii_host, int32_t n_scales, void *out, int out_dtype,
                             int32_t descriptor_mask, double *boxes_host_out, int64_t *n_voxels_host, void *stream)
{
    if (!mailbox || n < 0 || n_scales < 0 || (n_scales > 0 && (!edges_host || !radii_host)) || (n > 0 && (!xyz || !out)))
        return fail(NBR_ERR_INVALID, "nbr_tile_step: bad argument");
    NBR_TRY(check_cloud_dtype(dtype, "nbr_tile_step"));
    if (out_dtype != NBR_F32 && out_dtype != NBR_F64) return fail(NBR_ERR_INVALID, "nbr_tile_step: bad out_dtype");
    cudaStream_t s = (cudaStream_t)stream;
    Scratch perm, sorted;
    Plan *P = nullptr;
    NBR_TRY(tile_step_plan(reinterpret_cast<Mailbox *>(mailbox), xyz, dtype, n, edges_host, radii_host, n_scales, descriptor_mask,
                           boxes_host_out, perm, sorted, &P, s));
    if (!P) return NBR_OK;
    int rc = plan_run_sorted(P, sorted.ptr, dtype, perm.as<uint32_t>(), n, out, out_dtype, s);
    if (rc == NBR_OK && n_voxels_host) rc = plan_voxel_counts(P, n_voxels_host);
    delete P;
    return rc;
}

