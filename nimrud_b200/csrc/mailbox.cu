// mailbox.cu -- multi-GPU halo exchange over peer-mapped memory (NVLink / NVSwitch), no NCCL call and no
// count hand-shake on the path.
//
// every rank owns one device allocation, its MAILBOX: a header (cursor, per-peer flags, the table of tile
// boxes) followed by rows of 3 coordinates.  the other ranks map it (CUDA IPC between processes; plain
// pointers when several tiles live in one process, which is how the single-GPU tests drive this code).
// per step (epoch):
//   1. box publish : the rank's tile box (bounding box kernel, index.cu) is stored into slot [rank] of every
//                    peer's box table, then the slot's epoch flag (system-scope release).
//   2. box wait    : one warp spins until every slot carries this epoch, copies the table to pinned host
//                    memory; the host synchronises HERE -- the only host synchronisation of the exchange (the
//                    single-GPU path has the same one: the grid parameters are derived on the host from the
//                    bounding box).
//   3. halo push   : ONE pass over the tile: every point is tested against the box of every other tile grown
//                    by the halo width (inclusive, the precedent is nested_regions, nimrud/utils/geometry.py:203-253);
//                    a warp reserves rows in the destination's mailbox with one remote atomicAdd on its cursor
//                    and stores the points straight into the peer's memory.  the last block of the grid raises
//                    this rank's "done" flag in every peer's header.
//   4. halo wait   : one warp spins until every peer's done flag carries this epoch; the cursor then is the
//                    number of rows received (kept on the device: the lattice build reads it there).
// a peer can only push epoch e+1 after it has seen this rank's box of epoch e+1, which this rank publishes
// stream-ordered after every kernel of epoch e that reads the mailbox: one buffer is enough.
// waits are bounded (NBR_MAILBOX_TIMEOUT_MS, default 120 s): a rank that never arrives raises a sticky flag
// instead of hanging the GPU.
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "mailbox.cuh"

namespace nbr {

int bbox(const void *xyz, int dtype, int64_t n, int ndim, double *lohi_dev, cudaStream_t stream);

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long global_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

struct PeerHeaders {
    MailboxHeader *h[MB_MAX_WORLD];
};

// lane d stores this rank's box into slot [rank] of peer d's table, then the slot's epoch
__global__ void box_publish_kernel(const double *__restrict__ lohi, double n_points, PeerHeaders P, int rank, int world,
                                   unsigned long long epoch)
{
    const int d = threadIdx.x;
    if (d >= world) return;
    MailboxHeader *H = P.h[d];
    volatile double *slot = H->boxes[rank];
    if (n_points > 0) {
#pragma unroll
        for (int k = 0; k < 6; ++k) slot[k] = lohi[k];
    } else {
        // an empty tile contributes the neutral box of min / max
#pragma unroll
        for (int k = 0; k < 3; ++k) { slot[k] = INFINITY; slot[3 + k] = -INFINITY; }
    }
    slot[6] = n_points;
    slot[7] = 0.0;
    __threadfence_system();
    st_release_sys(&H->box_epoch[rank], epoch);
}

// lane d waits for slot d; the table goes to pinned host memory [world][8] and status[0] = 1 on a timeout
__global__ void box_wait_kernel(MailboxHeader *H, int world, unsigned long long epoch, double *host_boxes,
                                unsigned long long *host_status, unsigned long long timeout_ns)
{
    const int d = threadIdx.x;
    bool ok = true;
    if (d < world) {
        const unsigned long long t0 = global_ns();
        while (ld_acquire_sys(&H->box_epoch[d]) < epoch) {
            if (global_ns() - t0 > timeout_ns) { ok = false; break; }
            __nanosleep(200);
        }
        if (ok) {
            const volatile double *slot = H->boxes[d];
#pragma unroll
            for (int k = 0; k < 8; ++k) host_boxes[d * 8 + k] = slot[k];
        }
    }
    const bool all_ok = __all_sync(0xffffffffu, ok);
    if (threadIdx.x == 0) {
        if (!all_ok) H->timeout = 1;
        host_status[0] = all_ok ? 0ull : 1ull;
        __threadfence_system();
    }
}

struct PushDev {
    float in_lo[3], in_hi[3];                          // a box that meets no destination's grown box: points strictly
                                                       // inside it (most of the tile) skip the per-destination tests
    double lo[MB_MAX_WORLD][3], hi[MB_MAX_WORLD][3];   // grown boxes of the destinations (inverted: no point matches)
    MailboxHeader *hdr[MB_MAX_WORLD];                  // destination headers (peer-mapped)
    long long cap[MB_MAX_WORLD];                       // rows a destination's mailbox holds
    int n_dst;
    int rank;
    unsigned long long epoch;
};

template <typename T>
__device__ __forceinline__ uint32_t push_mask(const T *__restrict__ xyz, int64_t i, int64_t n, const PushDev &P, T &px, T &py, T &pz)
{
    uint32_t m = 0;
    if (i < n) {
        px = xyz[i * 3 + 0]; py = xyz[i * 3 + 1]; pz = xyz[i * 3 + 2];
        const float fx = (float)px, fy = (float)py, fz = (float)pz;          // rounding is monotone: strictly inside stays inside or on the face
        if (fx > P.in_lo[0] && fx < P.in_hi[0] && fy > P.in_lo[1] && fy < P.in_hi[1] && fz > P.in_lo[2] && fz < P.in_hi[2]) return 0u;
        const double x = (double)px, y = (double)py, z = (double)pz;
        for (int d = 0; d < P.n_dst; ++d)
            if (x >= P.lo[d][0] && x <= P.hi[d][0] && y >= P.lo[d][1] && y <= P.hi[d][1] && z >= P.lo[d][2] && z <= P.hi[d][2])
                m |= 1u << d;
    }
    return m;
}

// persistent blocks, two sweeps over the block's share of the tile.  sweep 1 counts the block's points per
// destination; ONE remote atomic per block and destination then reserves its rows (a remote atomic per warp
// serialises on the destination's cursor: 100k round trips over NVLink took 0.37 ms for a 10M-point tile);
// sweep 2 (the share is L2-resident by then) stores the points into the reserved rows of the peers' mailboxes.
template <typename T>
__global__ void __launch_bounds__(256)
halo_push_kernel(const T *__restrict__ xyz, int64_t n, const __grid_constant__ PushDev P, MailboxHeader *own)
{
    __shared__ unsigned long long s_base[MB_MAX_WORLD];
    __shared__ unsigned int s_cnt[MB_MAX_WORLD];
    __shared__ bool last;
    const int lane = threadIdx.x & 31;
    const uint32_t lt = lanemask_lt();
    if (threadIdx.x < MB_MAX_WORLD) s_cnt[threadIdx.x] = 0;
    __syncthreads();
    const int64_t start = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) - lane, stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t first = start; first < n; first += stride) {              // whole warps: the votes need every lane
        T px, py, pz;
        const uint32_t m = push_mask<T>(xyz, first + lane, n, P, px, py, pz);
        uint32_t any = __reduce_or_sync(0xffffffffu, m);
        while (any) {
            const int d = __ffs(any) - 1;
            any &= any - 1;
            const uint32_t votes = __ballot_sync(0xffffffffu, (m >> d) & 1u);
            if (lane == 0) atomicAdd(&s_cnt[d], (unsigned int)__popc(votes));
        }
    }
    __syncthreads();
    if ((int)threadIdx.x < P.n_dst) {
        const unsigned int c = s_cnt[threadIdx.x];
        s_base[threadIdx.x] = c ? atomicAdd_system(&P.hdr[threadIdx.x]->cursor, (unsigned long long)c) : 0ull;
        s_cnt[threadIdx.x] = 0;                                             // sweep 2: running offset inside the block's rows
    }
    __syncthreads();
    bool wrote = false;
    for (int64_t first = start; first < n; first += stride) {
        T px = 0, py = 0, pz = 0;
        const uint32_t m = push_mask<T>(xyz, first + lane, n, P, px, py, pz);
        uint32_t any = __reduce_or_sync(0xffffffffu, m);
        while (any) {
            const int d = __ffs(any) - 1;
            any &= any - 1;
            const bool mine = (m >> d) & 1u;
            const uint32_t votes = __ballot_sync(0xffffffffu, mine);
            unsigned int off = 0;
            if (lane == 0) off = atomicAdd(&s_cnt[d], (unsigned int)__popc(votes));
            off = __shfl_sync(0xffffffffu, off, 0);
            if (mine) {
                const long long row = (long long)s_base[d] + off + __popc(votes & lt);
                if (row < P.cap[d]) {
                    T *dst = reinterpret_cast<T *>(reinterpret_cast<unsigned char *>(P.hdr[d]) + MB_HEADER_BYTES) + row * 3;
                    dst[0] = px; dst[1] = py; dst[2] = pz;
                    wrote = true;
                }
            }
        }
    }
    // the last block to finish raises this rank's done flag at every destination: every writer fences its remote
    // stores at system scope before its block takes a ticket, the last block fences again before the flags
    if (wrote) __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned long long ticket = atomicAdd(&own->blocks_done, 1ull);
        last = ticket == (unsigned long long)gridDim.x - 1;
    }
    __syncthreads();
    if (last) {
        __threadfence_system();
        if (threadIdx.x == 0) own->blocks_done = 0;
        if ((int)threadIdx.x < P.n_dst) st_release_sys(&P.hdr[threadIdx.x]->halo_done[P.rank], P.epoch);
    }
}

// lane d waits for peer d's done flag; then the cursor is the number of rows received
__global__ void halo_wait_kernel(MailboxHeader *H, int rank, int world, unsigned long long epoch, long long cap,
                                 unsigned long long *host_status, unsigned long long timeout_ns)
{
    const int d = threadIdx.x;
    bool ok = true;
    if (d < world && d != rank) {
        const unsigned long long t0 = global_ns();
        while (ld_acquire_sys(&H->halo_done[d]) < epoch) {
            if (global_ns() - t0 > timeout_ns) { ok = false; break; }
            __nanosleep(200);
        }
    }
    const bool all_ok = __all_sync(0xffffffffu, ok);
    if (threadIdx.x == 0) {
        const unsigned long long c = all_ok ? ld_acquire_sys(&H->cursor) : 0ull;
        H->count = c < (unsigned long long)cap ? c : (unsigned long long)cap;
        if (c > (unsigned long long)cap) { H->overflow += c - (unsigned long long)cap; host_status[1] = H->overflow; }
        if (!all_ok) { H->timeout = 1; host_status[0] = 1ull; }
        host_status[2] = c;
        H->cursor = 0;                       // nobody pushes again before this rank's next box is out
        __threadfence_system();
    }
}

// feature all-gather, last step: lane d tells peer d that this rank's rows of the epoch are in its buffer (the kernels
// that wrote them are earlier on this stream; the system-scope fence + release make them visible before the flag), then
// waits for peer d's flag in this rank's own header.  signal first, wait second: no rank waits for a rank that waits
__global__ void gather_signal_wait_kernel(MailboxHeader *H, PeerHeaders P, int rank, int world, unsigned long long epoch,
                                          unsigned long long *host_status, unsigned long long timeout_ns)
{
    const int d = threadIdx.x;
    bool ok = true;
    if (d < world && d != rank) {
        __threadfence_system();
        st_release_sys(&P.h[d]->gather_done[rank], epoch);
        const unsigned long long t0 = global_ns();
        while (ld_acquire_sys(&H->gather_done[d]) < epoch) {
            if (global_ns() - t0 > timeout_ns) { ok = false; break; }
            __nanosleep(200);
        }
    }
    const bool all_ok = __all_sync(0xffffffffu, ok);
    if (threadIdx.x == 0 && !all_ok) {
        H->timeout = 1;
        host_status[0] = 1ull;
        __threadfence_system();
    }
}

// fallback of the fused write-out (scale sets the 7x7x7 kernel does not cover alone): the rank's finished rows, in
// tile order, go to every peer's staging buffer as they are; their row numbers are the identity
__global__ void gather_push_rows_kernel(const uint4 *__restrict__ rows, size_t n_pieces, int64_t n_rows, RowDests D, size_t row_bytes)
{
    const size_t stride = (size_t)gridDim.x * blockDim.x, first = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (size_t i = first; i < n_pieces; i += stride) {
        const uint4 v = rows[i];
        for (int d = 0; d < D.n; ++d)
            if (d != D.self) __stcs(reinterpret_cast<uint4 *>(D.base[d] + (size_t)D.row_offset * row_bytes) + i, v);
    }
    for (size_t i = first; i < (size_t)n_rows; i += stride)
        for (int d = 0; d < D.n; ++d)
            if (d != D.self) D.perm[d][D.row_offset + i] = (uint32_t)i;
}

struct UnpermuteDev {
    long long off[MB_MAX_WORLD + 1];       // first row of every rank's share
    int world, self;
};

// staged rows -> their places: a warp takes 32 consecutive staged rows (one contiguous read), looks up their row
// numbers and stores each row as 16-byte pieces into its place of the result.  the stores are random 80-byte accesses:
// the kernel lives on memory-level parallelism, so all pieces of a group are loaded before the first one is stored
// (CPR = 16-byte pieces per row, compile-time for the common row sizes; measured 0.92 ms per 10M rows with one load in
// flight per warp)
template <int CPR>
__global__ void gather_unpermute_kernel(const unsigned char *__restrict__ staged, const uint32_t *__restrict__ perm, UnpermuteDev U,
                                        int cpr_arg, unsigned char *__restrict__ out)
{
    const int lane = threadIdx.x & 31;
    const int cpr = CPR > 0 ? CPR : cpr_arg;
    const long long total = U.off[U.world], row_bytes = 16ll * cpr;
    const long long n_groups = (total + 31) >> 5;
    for (long long grp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; grp < n_groups; grp += ((long long)gridDim.x * blockDim.x) >> 5) {
        const long long g = grp * 32 + lane;
        long long dst = -1;                                   // destination row of staged row g; -1: not staged (own share, tail)
        if (g < total && !(g >= U.off[U.self] && g < U.off[U.self + 1])) {
            int r = 0;
            while (g >= U.off[r + 1]) ++r;
            dst = U.off[r] + (long long)perm[g];
        }
        if (!__any_sync(0xffffffffu, dst >= 0)) continue;
        const unsigned char *src = staged + grp * 32 * row_bytes;
        if (CPR > 0) {
            constexpr int K = CPR > 0 ? CPR : 1;
            uint4 v[K];
            long long d[K];
            int c[K];
#pragma unroll
            for (int t = 0; t < CPR; ++t) {
                const int p = lane + 32 * t, row = p / K;
                c[t] = p - row * K;
                d[t] = __shfl_sync(0xffffffffu, dst, row);
                if (d[t] >= 0) v[t] = __ldcs(reinterpret_cast<const uint4 *>(src) + p);
            }
#pragma unroll
            for (int t = 0; t < CPR; ++t)
                if (d[t] >= 0) __stcs(reinterpret_cast<uint4 *>(out + d[t] * row_bytes) + c[t], v[t]);
        } else {
            for (int p = lane; p < 32 * cpr; p += 32) {
                const int row = p / cpr, c = p - row * cpr;
                const long long d = __shfl_sync(0xffffffffu, dst, row);
                if (d >= 0) __stcs(reinterpret_cast<uint4 *>(out + d * row_bytes) + c, __ldcs(reinterpret_cast<const uint4 *>(src) + p));
            }
        }
    }
}

static unsigned long long timeout_ns()
{
    static const unsigned long long ns = [] {
        const char *e = getenv("NBR_MAILBOX_TIMEOUT_MS");
        const double ms = e ? atof(e) : 120000.0;
        return (unsigned long long)((ms > 0 ? ms : 120000.0) * 1e6);
    }();
    return ns;
}

void Mailbox::gather_release()
{
    for (int d = 0; d < MB_MAX_WORLD; ++d) {
        if (gather_opened[d] && gather_peer[d]) cudaIpcCloseMemHandle(gather_peer[d]);
        gather_peer[d] = nullptr;
        gather_peer_bytes[d] = 0;
        gather_opened[d] = false;
    }
    if (gather_base) cudaFree(gather_base);
    gather_base = nullptr;
    gather_bytes = 0;
}

Mailbox::~Mailbox()
{
    int cur = 0;
    cudaGetDevice(&cur);
    cudaSetDevice(device);
    gather_release();
    for (int d = 0; d < MB_MAX_WORLD; ++d)
        if (opened[d] && peer[d]) cudaIpcCloseMemHandle(peer[d]);
    if (base) cudaFree(base);
    if (host_boxes) cudaFreeHost(host_boxes);
    if (box_dev) cudaFree(box_dev);
    cudaSetDevice(cur);
}

static size_t elem_bytes(int dtype) { return dtype == NBR_F32 ? 4 : 8; }

int halo_wait(Mailbox *M, cudaStream_t stream)
{
    PhaseTimer tw(PHASE_HALO_WAIT, stream);
    unsigned long long *status = reinterpret_cast<unsigned long long *>(M->host_boxes + 8 * MB_MAX_WORLD);
    halo_wait_kernel<<<1, 32, 0, stream>>>(reinterpret_cast<MailboxHeader *>(M->base), M->rank, M->world, M->epoch,
                                           (long long)M->capacity, status, timeout_ns());
    NBR_LAUNCHED();
    return NBR_OK;
}

int gather_finish(Mailbox *M, cudaStream_t stream)
{
    if (M->world <= 1) return NBR_OK;
    PhaseTimer tw(PHASE_HALO_WAIT, stream);
    PeerHeaders P;
    for (int d = 0; d < MB_MAX_WORLD; ++d) P.h[d] = reinterpret_cast<MailboxHeader *>(M->peer[d]);
    unsigned long long *status = reinterpret_cast<unsigned long long *>(M->host_boxes + 8 * MB_MAX_WORLD);
    gather_signal_wait_kernel<<<1, 32, 0, stream>>>(reinterpret_cast<MailboxHeader *>(M->base), P, M->rank, M->world, M->epoch, status,
                                                    timeout_ns());
    NBR_LAUNCHED();
    return NBR_OK;
}

int gather_push_rows(const RowDests *D, const void *rows, int64_t n, size_t row_bytes, cudaStream_t stream)
{
    if (n <= 0 || D->n <= 1) return NBR_OK;
    if (row_bytes % 16 != 0 || ((uintptr_t)rows & 15) != 0) return fail(NBR_ERR_UNSUPPORTED, "feature gather: rows must be multiples of 16 bytes");
    const size_t pieces = (size_t)n * row_bytes / 16;
    const unsigned blocks = (unsigned)std::min<size_t>(ceil_div(pieces, (size_t)256), (size_t)device_sm_count() * 8);
    gather_push_rows_kernel<<<blocks, 256, 0, stream>>>(reinterpret_cast<const uint4 *>(rows), pieces, n, *D, row_bytes);
    NBR_LAUNCHED();
    return NBR_OK;
}

int gather_unpermute(const Mailbox *M, const int64_t *row_offsets, size_t row_bytes, void *out_all, cudaStream_t stream)
{
    if (M->world <= 1) return NBR_OK;
    const int64_t total = row_offsets[M->world];
    if (total - (row_offsets[M->rank + 1] - row_offsets[M->rank]) <= 0 || row_bytes == 0) return NBR_OK;
    if (row_bytes % 16 != 0 || ((uintptr_t)out_all & 15) != 0) return fail(NBR_ERR_UNSUPPORTED, "feature gather: rows must be multiples of 16 bytes");
    UnpermuteDev U;
    memset(&U, 0, sizeof(U));
    for (int r = 0; r <= M->world; ++r) U.off[r] = row_offsets[r];
    U.world = M->world;
    U.self = M->rank;
    const unsigned blocks = (unsigned)std::min<int64_t>(ceil_div(ceil_div(total, (int64_t)32), (int64_t)8), (int64_t)device_sm_count() * 8);
    const uint32_t *perm = reinterpret_cast<const uint32_t *>(M->gather_base + gather_perm_offset(total, row_bytes));
    const int cpr = (int)(row_bytes / 16);
    unsigned char *dst = reinterpret_cast<unsigned char *>(out_all);
    if (cpr == 5) gather_unpermute_kernel<5><<<blocks, 256, 0, stream>>>(M->gather_base, perm, U, cpr, dst);           // 5 scales, float32
    else if (cpr == 10) gather_unpermute_kernel<10><<<blocks, 256, 0, stream>>>(M->gather_base, perm, U, cpr, dst);    // 5 scales, float64
    else gather_unpermute_kernel<0><<<blocks, 256, 0, stream>>>(M->gather_base, perm, U, cpr, dst);
    NBR_LAUNCHED();
    return NBR_OK;
}

}  // namespace nbr

using namespace nbr;

// ---- gather buffers: one per rank, sized for the rows of every rank; (re)allocation is collective on the caller's side
// (every rank allocates, exchanges the handles, connects, and only then runs the next step)
extern "C" int nbr_mailbox_gather_alloc(nbr_mailbox *mb, uint64_t bytes)
{
    Mailbox *M = reinterpret_cast<Mailbox *>(mb);
    if (!M) return fail(NBR_ERR_INVALID, "nbr_mailbox_gather_alloc: null argument");
    int cur = 0;
    NBR_CUDA(cudaGetDevice(&cur));
    NBR_CUDA(cudaSetDevice(M->device));
    cudaDeviceSynchronize();
    M->gather_release();
    cudaError_t e = bytes ? cudaMalloc(&M->gather_base, bytes) : cudaSuccess;
    cudaSetDevice(cur);
    if (e != cudaSuccess) { M->gather_base = nullptr; return fail(NBR_ERR_CUDA, std::string("nbr_mailbox_gather_alloc: ") + cudaGetErrorString(e)); }
    M->gather_bytes = bytes;
    M->gather_peer[M->rank] = M->gather_base;
    M->gather_peer_bytes[M->rank] = bytes;
    return NBR_OK;
}

extern "C" int nbr_mailbox_gather_ipc_handle(const nbr_mailbox *mb, void *handle_out_64)
{
    const Mailbox *M = reinterpret_cast<const Mailbox *>(mb);
    if (!M || !handle_out_64 || !M->gather_base) return fail(NBR_ERR_INVALID, "nbr_mailbox_gather_ipc_handle: no gather buffer");
    cudaIpcMemHandle_t h;
    NBR_CUDA(cudaIpcGetMemHandle(&h, M->gather_base));
    memcpy(handle_out_64, &h, 64);
    return NBR_OK;
}

extern "C" int nbr_mailbox_gather_connect_ipc(nbr_mailbox *mb, int32_t peer, const void *handle_64, uint64_t bytes)
{
    Mailbox *M = reinterpret_cast<Mailbox *>(mb);
    if (!M || !handle_64 || peer < 0 || peer >= M->world || peer == M->rank)
        return fail(NBR_ERR_INVALID, "nbr_mailbox_gather_connect_ipc: bad argument");
    if (M->gather_opened[peer] && M->gather_peer[peer]) cudaIpcCloseMemHandle(M->gather_peer[peer]);
    M->gather_peer[peer] = nullptr;
    M->gather_opened[peer] = false;
    cudaIpcMemHandle_t h;
    memcpy(&h, handle_64, 64);
    void *p = nullptr;
    NBR_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    M->gather_peer[peer] = reinterpret_cast<unsigned char *>(p);
    M->gather_peer_bytes[peer] = bytes;
    M->gather_opened[peer] = true;
    return NBR_OK;
}

extern "C" int nbr_mailbox_gather_connect_local(nbr_mailbox *mb, int32_t peer, const nbr_mailbox *peer_mb)
{
    Mailbox *M = reinterpret_cast<Mailbox *>(mb);
    const Mailbox *Q = reinterpret_cast<const Mailbox *>(peer_mb);
    if (!M || !Q || peer < 0 || peer >= M->world || peer == M->rank || Q->rank != peer || !Q->gather_base)
        return fail(NBR_ERR_INVALID, "nbr_mailbox_gather_connect_local: bad argument");
    M->gather_peer[peer] = Q->gather_base;           // peer access between the devices was enabled by nbr_mailbox_connect_local
    M->gather_peer_bytes[peer] = Q->gather_bytes;
    M->gather_opened[peer] = false;
    return NBR_OK;
}

extern "C" uint64_t nbr_gather_staging_bytes(int64_t total_rows, int64_t row_bytes)
{
    return total_rows < 0 || row_bytes < 0 ? 0 : (uint64_t)gather_bytes_needed(total_rows, (size_t)row_bytes);
}

extern "C" void *nbr_mailbox_gather_ptr(const nbr_mailbox *mb, uint64_t *bytes_out)
{
    const Mailbox *M = reinterpret_cast<const Mailbox *>(mb);
    if (bytes_out) *bytes_out = M ? (uint64_t)M->gather_bytes : 0;
    return M ? M->gather_base : nullptr;
}

extern "C" int nbr_mailbox_create(nbr_mailbox **out, int32_t rank, int32_t world, int dtype, int64_t capacity_rows)
{
    if (!out || world < 1 || world > MB_MAX_WORLD || rank < 0 || rank >= world || capacity_rows < 0)
        return fail(NBR_ERR_INVALID, "nbr_mailbox_create: bad argument (world <= 16)");
    if (dtype != NBR_F32 && dtype != NBR_F64) return fail(NBR_ERR_INVALID, "nbr_mailbox_create: bad dtype");
    Mailbox *M = new Mailbox();
    M->rank = rank; M->world = world; M->dtype = dtype; M->capacity = capacity_rows;
    cudaError_t e = cudaGetDevice(&M->device);
    const size_t bytes = MB_HEADER_BYTES + (size_t)std::max<int64_t>(capacity_rows, 1) * 3 * elem_bytes(dtype);
    // cudaMalloc, not the stream-ordered pool: pool memory cannot be exported through cudaIpcGetMemHandle
    if (e == cudaSuccess) e = cudaMalloc(&M->base, bytes);
    if (e == cudaSuccess) e = cudaMemset(M->base, 0, MB_HEADER_BYTES);
    if (e == cudaSuccess) e = cudaMalloc(&M->box_dev, sizeof(double) * 8);
    if (e == cudaSuccess) e = cudaHostAlloc(&M->host_boxes, sizeof(double) * 8 * MB_MAX_WORLD + 64, cudaHostAllocMapped);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
        delete M;
        return fail(NBR_ERR_CUDA, std::string("nbr_mailbox_create: ") + cudaGetErrorString(e));
    }
    memset(M->host_boxes, 0, sizeof(double) * 8 * MB_MAX_WORLD + 64);
    M->peer[rank] = M->base;
    *out = reinterpret_cast<nbr_mailbox *>(M);
    return NBR_OK;
}

extern "C" void nbr_mailbox_destroy(nbr_mailbox *mb) { delete reinterpret_cast<Mailbox *>(mb); }

// closes this rank's mappings of the PEERS' allocations (mailboxes and staging buffers; staging_only != 0: only those).
// exported memory must not be freed while another process still maps it: before a collective re-allocation or
// teardown every rank disconnects, the ranks meet at a barrier, and only then the owners free
extern "C" int nbr_mailbox_disconnect(nbr_mailbox *mb, int32_t staging_only)
{
    Mailbox *M = reinterpret_cast<Mailbox *>(mb);
    if (!M) return fail(NBR_ERR_INVALID, "nbr_mailbox_disconnect: null argument");
    int cur = 0;
    NBR_CUDA(cudaGetDevice(&cur));
    NBR_CUDA(cudaSetDevice(M->device));
    cudaDeviceSynchronize();
    for (int d = 0; d < MB_MAX_WORLD; ++d) {
        if (d == M->rank) continue;
        if (M->gather_opened[d] && M->gather_peer[d]) cudaIpcCloseMemHandle(M->gather_peer[d]);
        M->gather_peer[d] = nullptr;
        M->gather_peer_bytes[d] = 0;
        M->gather_opened[d] = false;
        if (!staging_only) {
            if (M->opened[d] && M->peer[d]) cudaIpcCloseMemHandle(M->peer[d]);
            M->peer[d] = nullptr;
            M->opened[d] = false;
        }
    }
    cudaSetDevice(cur);
    return NBR_OK;
}

extern "C" int nbr_mailbox_ipc_handle(const nbr_mailbox *mb, void *handle_out_64)
{
    const Mailbox *M = reinterpret_cast<const Mailbox *>(mb);
    if (!M || !handle_out_64) return fail(NBR_ERR_INVALID, "nbr_mailbox_ipc_handle: null argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size");
    cudaIpcMemHandle_t h;
    NBR_CUDA(cudaIpcGetMemHandle(&h, M->base));
    memcpy(handle_out_64, &h, 64);
    return NBR_OK;
}

extern "C" int nbr_mailbox_connect_ipc(nbr_mailbox *mb, int32_t peer, const void *handle_64)
{
    Mailbox *M = reinterpret_cast<Mailbox *>(mb);
    if (!M || !handle_64 || peer < 0 || peer >= M->world || peer == M->rank)
        return fail(NBR_ERR_INVALID, "nbr_mailbox_connect_ipc: bad argument");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle_64, 64);
    void *p = nullptr;
    NBR_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    M->peer[peer] = reinterpret_cast<unsigned char *>(p);
    M->opened[peer] = true;
    return NBR_OK;
}

// tiles of one process (tests, or several tiles per GPU): the peer's allocation is used directly
extern "C" int nbr_mailbox_connect_local(nbr_mailbox *mb, int32_t peer, const nbr_mailbox *peer_mb)
{
    Mailbox *M = reinterpret_cast<Mailbox *>(mb);
    const Mailbox *Q = reinterpret_cast<const Mailbox *>(peer_mb);
    if (!M || !Q || peer < 0 || peer >= M->world || peer == M->rank || Q->rank != peer || Q->world != M->world ||
        Q->dtype != M->dtype)
        return fail(NBR_ERR_INVALID, "nbr_mailbox_connect_local: bad argument");
    if (Q->device != M->device) {
        int can = 0;
        NBR_CUDA(cudaDeviceCanAccessPeer(&can, M->device, Q->device));
        if (!can) return fail(NBR_ERR_UNSUPPORTED, "nbr_mailbox_connect_local: no peer access between the devices");
        cudaError_t e = cudaDeviceEnablePeerAccess(Q->device, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
            return fail(NBR_ERR_CUDA, std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e));
        cudaGetLastError();
    }
    M->peer[peer] = Q->base;
    M->opened[peer] = false;
    M->capacity_of[peer] = Q->capacity;
    return NBR_OK;
}

extern "C" int nbr_mailbox_set_peer_capacity(nbr_mailbox *mb, int32_t peer, int64_t capacity_rows)
{
    Mailbox *M = reinterpret_cast<Mailbox *>(mb);
    if (!M || peer < 0 || peer >= M->world || capacity_rows < 0) return fail(NBR_ERR_INVALID, "nbr_mailbox_set_peer_capacity: bad argument");
    M->capacity_of[peer] = capacity_rows;
    return NBR_OK;
}

static int check_connected(const Mailbox *M, const char *who)
{
    for (int d = 0; d < M->world; ++d)
        if (!M->peer[d]) return fail(NBR_ERR_INVALID, std::string(who) + ": mailbox is not connected to every peer");
    return NBR_OK;
}

// step 1: this rank's tile box (n may be 0: the neutral box) goes to every peer.  starts a new epoch.
extern "C" int nbr_tile_box_publish(nbr_mailbox *mb, const void *xyz, int dtype, int64_t n, void *stream)
{
    Mailbox *M = reinterpret_cast<Mailbox *>(mb);
    if (!M || (n > 0 && !xyz) || n < 0) return fail(NBR_ERR_INVALID, "nbr_tile_box_publish: bad argument");
    if (dtype != M->dtype) return fail(NBR_ERR_INVALID, "nbr_tile_box_publish: dtype differs from the mailbox's");
    NBR_TRY(check_connected(M, "nbr_tile_box_publish"));
    cudaStream_t s = (cudaStream_t)stream;
    if (n > 0) {
        PhaseTimer t(PHASE_BBOX, s);
        NBR_TRY(bbox(xyz, dtype, n, 3, M->box_dev, s));
    }
    M->epoch += 1;
    PhaseTimer tb(PHASE_BOXES, s);
    PeerHeaders P;
    for (int d = 0; d < MB_MAX_WORLD; ++d) P.h[d] = reinterpret_cast<MailboxHeader *>(M->peer[d]);
    box_publish_kernel<<<1, 32, 0, s>>>(M->box_dev, (double)n, P, M->rank, M->world, M->epoch);
    NBR_LAUNCHED();
    return NBR_OK;
}

// step 2: every rank's box of this epoch -> boxes_host[world][8] = lo[3], hi[3], n_points, 0.  SYNCHRONISES the stream.
extern "C" int nbr_tile_boxes_wait(nbr_mailbox *mb, double *boxes_host, void *stream)
{
    Mailbox *M = reinterpret_cast<Mailbox *>(mb);
    if (!M || !boxes_host) return fail(NBR_ERR_INVALID, "nbr_tile_boxes_wait: null argument");
    cudaStream_t s = (cudaStream_t)stream;
    unsigned long long *status = reinterpret_cast<unsigned long long *>(M->host_boxes + 8 * MB_MAX_WORLD);
    {
        PhaseTimer tb(PHASE_BOXES, s);
        box_wait_kernel<<<1, 32, 0, s>>>(reinterpret_cast<MailboxHeader *>(M->base), M->world, M->epoch, M->host_boxes, status,
                                         timeout_ns());
        NBR_LAUNCHED();
    }
    NBR_CUDA(cudaStreamSynchronize(s));
    if (status[0]) return fail(NBR_ERR_CUDA, "nbr_tile_boxes_wait: a peer did not publish its tile box in time");
    if (status[1]) return fail(NBR_ERR_UNSUPPORTED, "halo mailbox overflow in an earlier step: " + std::to_string(status[1]) +
                                                        " rows were dropped; create the mailbox with a larger capacity");
    memcpy(boxes_host, M->host_boxes, sizeof(double) * 8 * M->world);
    return NBR_OK;
}

// step 3: one pass over the tile, points inside peer d's box grown by h go straight into d's mailbox
extern "C" int nbr_halo_push(nbr_mailbox *mb, const void *xyz, int dtype, int64_t n, const double *boxes_host, double h,
                             void *stream)
{
    Mailbox *M = reinterpret_cast<Mailbox *>(mb);
    if (!M || !boxes_host || (n > 0 && !xyz) || n < 0 || !(h >= 0)) return fail(NBR_ERR_INVALID, "nbr_halo_push: bad argument");
    if (dtype != M->dtype) return fail(NBR_ERR_INVALID, "nbr_halo_push: dtype differs from the mailbox's");
    NBR_TRY(check_connected(M, "nbr_halo_push"));
    PushDev P;
    memset(&P, 0, sizeof(P));
    P.rank = M->rank;
    P.epoch = M->epoch;
    const double *mine = boxes_host + 8 * M->rank;
    for (int d = 0; d < M->world; ++d) {
        if (d == M->rank) continue;
        const int k = P.n_dst++;
        const double *b = boxes_host + 8 * d;
        bool overlap = n > 0 && b[6] > 0;
        for (int a = 0; a < 3; ++a) {
            P.lo[k][a] = b[a] - h;
            P.hi[k][a] = b[3 + a] + h;
            overlap = overlap && !(P.lo[k][a] > mine[3 + a]) && !(P.hi[k][a] < mine[a]);
        }
        if (!overlap)
            for (int a = 0; a < 3; ++a) { P.lo[k][a] = 1.0; P.hi[k][a] = -1.0; }     // matches nothing
        P.hdr[k] = reinterpret_cast<MailboxHeader *>(M->peer[d]);
        P.cap[k] = M->capacity_of[d] > 0 ? M->capacity_of[d] : M->capacity;
    }
    // interior box: start from this tile's box and, for every destination that can receive something, move the
    // face that loses the least volume until the destination's grown box is outside.  float bounds, rounded inward
    {
        double ilo[3], ihi[3];
        for (int a = 0; a < 3; ++a) { ilo[a] = mine[a]; ihi[a] = mine[3 + a]; }
        for (int k = 0; k < P.n_dst; ++k) {
            if (P.lo[k][0] > P.hi[k][0]) continue;                       // matches nothing
            int best = -1; bool low_side = false; double best_loss = INFINITY;
            for (int a = 0; a < 3; ++a) {
                const double span = std::max(ihi[a] - ilo[a], 1e-300);
                // cut below the destination (keep [ilo, lo_k)) or above it (keep (hi_k, ihi])
                const double loss_hi = (ihi[a] - std::min(ihi[a], P.lo[k][a])) / span;     // keep the low part
                const double loss_lo = (std::max(ilo[a], P.hi[k][a]) - ilo[a]) / span;     // keep the high part
                if (loss_hi < best_loss) { best_loss = loss_hi; best = a; low_side = false; }
                if (loss_lo < best_loss) { best_loss = loss_lo; best = a; low_side = true; }
            }
            if (best < 0) continue;
            if (low_side) ilo[best] = std::max(ilo[best], P.hi[k][best]);
            else          ihi[best] = std::min(ihi[best], P.lo[k][best]);
        }
        for (int a = 0; a < 3; ++a) {
            // strict tests against bounds moved one float inward: a point that passes is outside every grown box
            P.in_lo[a] = nextafterf((float)ilo[a], INFINITY);
            if ((double)P.in_lo[a] < ilo[a]) P.in_lo[a] = nextafterf(P.in_lo[a], INFINITY);
            P.in_hi[a] = nextafterf((float)ihi[a], -INFINITY);
            if ((double)P.in_hi[a] > ihi[a]) P.in_hi[a] = nextafterf(P.in_hi[a], -INFINITY);
            // the tile's own faces are not constraints: open them up so that border points of the tile count as interior
            if (ilo[a] == mine[a]) P.in_lo[a] = -INFINITY;
            if (ihi[a] == mine[3 + a]) P.in_hi[a] = INFINITY;
        }
    }
    cudaStream_t s = (cudaStream_t)stream;
    const unsigned blocks = (unsigned)std::max<int64_t>(1, std::min<int64_t>(ceil_div(n, 256), (int64_t)device_sm_count() * 8));
    MailboxHeader *own = reinterpret_cast<MailboxHeader *>(M->base);
    PhaseTimer tp(PHASE_PUSH, s);
    if (dtype == NBR_F32) halo_push_kernel<float><<<blocks, 256, 0, s>>>((const float *)xyz, n, P, own);
    else                  halo_push_kernel<double><<<blocks, 256, 0, s>>>((const double *)xyz, n, P, own);
    NBR_LAUNCHED();
    return NBR_OK;
}

// step 4 (stream-ordered, no host synchronisation): afterwards nbr_mailbox_rows / nbr_mailbox_count_dev are valid
extern "C" int nbr_halo_wait(nbr_mailbox *mb, void *stream)
{
    Mailbox *M = reinterpret_cast<Mailbox *>(mb);
    if (!M) return fail(NBR_ERR_INVALID, "nbr_halo_wait: null argument");
    return halo_wait(M, (cudaStream_t)stream);
}

extern "C" const void *nbr_mailbox_rows(const nbr_mailbox *mb)
{
    const Mailbox *M = reinterpret_cast<const Mailbox *>(mb);
    return M ? M->base + MB_HEADER_BYTES : nullptr;
}

extern "C" const uint64_t *nbr_mailbox_count_dev(const nbr_mailbox *mb)
{
    const Mailbox *M = reinterpret_cast<const Mailbox *>(mb);
    return M ? reinterpret_cast<const uint64_t *>(&reinterpret_cast<const MailboxHeader *>(M->base)->count) : nullptr;
}

// debugging / tests: the rows of the last completed step -> dst_dev (device, room for max_rows rows); SYNCHRONISES
extern "C" int nbr_mailbox_read(const nbr_mailbox *mb, void *dst_dev, int64_t max_rows, int64_t *n_rows_host, void *stream)
{
    const Mailbox *M = reinterpret_cast<const Mailbox *>(mb);
    if (!M || !dst_dev || !n_rows_host || max_rows < 0) return fail(NBR_ERR_INVALID, "nbr_mailbox_read: bad argument");
    cudaStream_t s = (cudaStream_t)stream;
    unsigned long long count = 0;
    NBR_CUDA(cudaMemcpyAsync(&count, M->count_dev(), sizeof(count), cudaMemcpyDeviceToHost, s));
    NBR_CUDA(cudaStreamSynchronize(s));
    const int64_t rows = std::min<int64_t>((int64_t)count, max_rows);
    if (rows > 0)
        NBR_CUDA(cudaMemcpyAsync(dst_dev, M->rows(), (size_t)rows * 3 * elem_bytes(M->dtype), cudaMemcpyDeviceToDevice, s));
    NBR_CUDA(cudaStreamSynchronize(s));
    *n_rows_host = (int64_t)count;
    return NBR_OK;
}

// host view of the status words written by the wait kernels: [0] timeout, [1] rows dropped so far (overflow),
// [2] rows pushed to this rank in the last completed epoch.  valid after the stream was synchronised.
extern "C" int nbr_mailbox_status(const nbr_mailbox *mb, uint64_t *status3_host)
{
    const Mailbox *M = reinterpret_cast<const Mailbox *>(mb);
    if (!M || !status3_host) return fail(NBR_ERR_INVALID, "nbr_mailbox_status: null argument");
    const volatile unsigned long long *status = reinterpret_cast<const unsigned long long *>(M->host_boxes + 8 * MB_MAX_WORLD);
    for (int k = 0; k < 3; ++k) status3_host[k] = status[k];
    return NBR_OK;
}
