// host_path.cu -- nbr_multiscale_features_host: the call the reference-facing Python shim makes for numpy
// arguments (process_single_core, nimrud/minimal/multiscale.py:27-67, with HOST buffers in and out).
//
// the path is bound by the PCIe link, not by the kernels (10M points x 5 scales: 5.4 ms of kernels against 1.6 GB of
// float64 rows), so it is built around the wire:
//   * WIRE FORMAT float32.  the device computes float32 rows (population is an exact small integer, the other
//     columns carry a 1e-4 tolerance and float32 keeps 6e-8), they cross PCIe at half the bytes, and host threads
//     widen them into the caller's float64 array while the next batch is on the wire.  NBR_HOST_WIRE=f64 sends
//     float64 rows instead; with a pinned float64 result a batch whose turn comes while the host threads are behind is
//     widened on the DEVICE and written in place (rows_to_host: routes), which takes it off the host's memory bus.
//   * PINNED RINGS.  pageable buffers (plain numpy arrays) are staged through pinned ring buffers by a pool of host
//     threads: cloud chunks in, row batches out; pinned caller buffers are used in place.  the rings are cached
//     for the life of the process.
//   * BATCHES.  queries run in batches of n/16 (at least 256k): the device->host copy of batch b (copy stream) and
//     the host-side widening of batch b-1 (host threads) overlap the kernels of batch b+1.  device-side row
//     buffers are a ring of 3 batches, not the whole result.
#include <stdlib.h>
#include <string.h>
#include <immintrin.h>
#include <sys/mman.h>

#include <algorithm>
#include <chrono>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "common.cuh"
#include "mailbox.cuh"
#include "plan.cuh"

namespace nbr {

// ------------------------------------------------------------------------------------------------
// host threads
// ------------------------------------------------------------------------------------------------
class HostPool {
public:
    static HostPool &get()
    {
        static HostPool pool;
        return pool;
    }
    int size() const { return (int)workers_.size() + 1; }
    int workers() const { return (int)workers_.size(); }
    // fn(part, parts) on every worker and on the caller; returns when all parts are done
    void run(const std::function<void(int, int)> &fn)
    {
        std::lock_guard<std::mutex> serial(run_mutex_);            // one parallel region at a time
        const int parts = size();
        {
            std::lock_guard<std::mutex> lock(m_);
            fn_ = &fn;
            parts_ = parts;
            pending_ = parts - 1;
            ++generation_;
        }
        cv_.notify_all();
        fn(parts - 1, parts);
        std::unique_lock<std::mutex> lock(m_);
        done_.wait(lock, [&] { return pending_ == 0; });
        fn_ = nullptr;
    }
    // fn(part, parts) on the workers only (parts = workers()); the caller goes on and calls wait() later.
    // fn must stay alive until wait() returns.
    void start(const std::function<void(int, int)> &fn)
    {
        run_mutex_.lock();
        std::lock_guard<std::mutex> lock(m_);
        fn_ = &fn;
        parts_ = std::max(1, workers());
        pending_ = workers();
        ++generation_;
        cv_.notify_all();
    }
    void wait()
    {
        {
            std::unique_lock<std::mutex> lock(m_);
            done_.wait(lock, [&] { return pending_ == 0; });
            fn_ = nullptr;
        }
        run_mutex_.unlock();
    }

private:
    HostPool()
    {
        // default: this process' share of the cores (torchrun exports LOCAL_WORLD_SIZE: one process per GPU)
        const char *env = getenv("NBR_HOST_THREADS"), *lws = getenv("LOCAL_WORLD_SIZE");
        const int share = lws && atoi(lws) > 0 ? atoi(lws) : 1;
        int n = env ? atoi(env) : (int)std::thread::hardware_concurrency() / share;
        n = std::max(2, std::min(n, 64));                           // at least one worker besides the caller
        for (int i = 0; i + 1 < n; ++i) workers_.emplace_back([this, i] { loop(i); });
    }
    ~HostPool()
    {
        {
            std::lock_guard<std::mutex> lock(m_);
            stop_ = true;
        }
        cv_.notify_all();
        for (auto &t : workers_) t.join();
    }
    void loop(int id)
    {
        uint64_t seen = 0;
        for (;;) {
            const std::function<void(int, int)> *fn = nullptr;
            {
                std::unique_lock<std::mutex> lock(m_);
                cv_.wait(lock, [&] { return stop_ || generation_ != seen; });
                if (stop_) return;
                seen = generation_;
                fn = fn_;
            }
            (*fn)(id, parts_);
            {
                std::lock_guard<std::mutex> lock(m_);
                if (--pending_ == 0) done_.notify_all();
            }
        }
    }
    std::vector<std::thread> workers_;
    std::mutex m_, run_mutex_;
    std::condition_variable cv_, done_;
    const std::function<void(int, int)> *fn_ = nullptr;
    uint64_t generation_ = 0;
    int pending_ = 0, parts_ = 1;
    bool stop_ = false;
};

// dst[i] = (double)src[i], streaming stores (the destination is written once and not read here)
#if defined(__x86_64__)
__attribute__((target("avx2"))) static void widen_avx2(const float *src, double *dst, size_t n)
{
    size_t i = 0;
    while (i < n && ((uintptr_t)(dst + i) & 31)) { dst[i] = (double)src[i]; ++i; }
    for (; i + 8 <= n; i += 8) {
        const __m256 v = _mm256_loadu_ps(src + i);
        _mm256_stream_pd(dst + i, _mm256_cvtps_pd(_mm256_castps256_ps128(v)));
        _mm256_stream_pd(dst + i + 4, _mm256_cvtps_pd(_mm256_extractf128_ps(v, 1)));
    }
    for (; i < n; ++i) dst[i] = (double)src[i];
    _mm_sfence();
}
#endif
static void widen(const float *src, double *dst, size_t n)
{
#if defined(__x86_64__)
    static const bool avx2 = __builtin_cpu_supports("avx2");
    if (avx2) { widen_avx2(src, dst, n); return; }
#endif
    for (size_t i = 0; i < n; ++i) dst[i] = (double)src[i];
}

// split [0, n) elements into `parts` pieces on 64-byte boundaries
static void part_range(size_t n, size_t elem, int part, int parts, size_t *first, size_t *count)
{
    const size_t per = ((n + parts - 1) / parts + 63) & ~(size_t)63;
    const size_t a = std::min(n, per * (size_t)part), b = std::min(n, a + per);
    (void)elem;
    *first = a;
    *count = b - a;
}

// ------------------------------------------------------------------------------------------------
// pinned rings (cached per process; grown on demand)
// ------------------------------------------------------------------------------------------------
struct PinnedRing {
    std::mutex m;
    void *buf[4] = {nullptr, nullptr, nullptr, nullptr};
    size_t bytes[4] = {0, 0, 0, 0};
    int get(int slot, size_t need, void **out)
    {
        if (bytes[slot] < need) {
            if (buf[slot]) cudaFreeHost(buf[slot]);
            buf[slot] = nullptr;
            bytes[slot] = 0;
            NBR_CUDA(cudaHostAlloc(&buf[slot], need, cudaHostAllocPortable));
            bytes[slot] = need;
        }
        *out = buf[slot];
        return NBR_OK;
    }
};
static PinnedRing g_ring_out, g_ring_in;
static std::mutex g_host_call;             // the rings are shared: one host-buffer call at a time

static bool is_pinned(const void *p)
{
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeHost;
}

static size_t esize(int dtype) { return dtype == NBR_F32 ? 4 : 8; }

// host cloud -> device, through the pinned input ring when the caller's buffer is pageable
static int upload(void *dev, const void *host, size_t bytes, cudaStream_t stream)
{
    if (bytes == 0) return NBR_OK;
    if (is_pinned(host)) {
        NBR_CUDA(cudaMemcpyAsync(dev, host, bytes, cudaMemcpyHostToDevice, stream));
        return NBR_OK;
    }
    // chunks through a ring of 4 pinned slots: the first chunk's copy into the ring is the pipeline's start-up cost
    const char *chunk_env = getenv("NBR_HOST_UPLOAD_MB");
    const size_t chunk = std::max<size_t>(1u << 20, (size_t)((chunk_env && atof(chunk_env) > 0 ? atof(chunk_env) : 4.0) * (double)(1u << 20)));
    constexpr int SLOTS = 4;
    cudaEvent_t freed[SLOTS] = {nullptr, nullptr, nullptr, nullptr};
    int rc = NBR_OK;
    for (int k = 0; k < SLOTS && !rc; ++k)
        if (cudaEventCreateWithFlags(&freed[k], cudaEventDisableTiming) != cudaSuccess) rc = fail(NBR_ERR_CUDA, "upload: cudaEventCreate");
    size_t at = 0;
    for (int it = 0; !rc && at < bytes; ++it) {
        const int slot = it % SLOTS;
        const size_t len = std::min(chunk, bytes - at);
        void *pin = nullptr;
        rc = g_ring_in.get(slot, chunk, &pin);
        if (rc) break;
        if (it >= SLOTS && cudaEventSynchronize(freed[slot]) != cudaSuccess) { rc = fail(NBR_ERR_CUDA, "upload: event"); break; }
        const char *src = (const char *)host + at;
        HostPool::get().run([&](int part, int parts) {
            size_t first, count;
            part_range(len, 1, part, parts, &first, &count);
            if (count) memcpy((char *)pin + first, src + first, count);
        });
        if (cudaMemcpyAsync((char *)dev + at, pin, len, cudaMemcpyHostToDevice, stream) != cudaSuccess ||
            cudaEventRecord(freed[slot], stream) != cudaSuccess)
            rc = fail(NBR_ERR_CUDA, "upload: cudaMemcpyAsync");
        at += len;
    }
    for (int k = 0; k < SLOTS; ++k)
        if (freed[k]) { cudaEventSynchronize(freed[k]); cudaEventDestroy(freed[k]); }
    return rc;
}

}  // namespace nbr

using namespace nbr;

// host memory for clouds / result arrays.  pinned = 1: page-locked (cudaHostAlloc; used in place by the host path,
// but locking 1.6 GB takes hundreds of milliseconds); pinned = 0: 2 MB-aligned pageable memory with huge pages
// requested.  the Python shim recycles its result arrays through the pageable kind: a FRESH 1.6 GB result costs
// more in page faults than its rows cost on the wire, a recycled one costs nothing.
static std::mutex g_host_alloc_mutex;
static std::vector<std::pair<void *, int>> g_host_allocs;

extern "C" int nbr_host_alloc(size_t bytes, int pinned, void **out)
{
    if (!out) return fail(NBR_ERR_INVALID, "nbr_host_alloc: null argument");
    *out = nullptr;
    if (bytes == 0) bytes = 1;
    if (pinned) {
        NBR_CUDA(cudaHostAlloc(out, bytes, cudaHostAllocPortable));
    } else {
        const size_t align = 2u << 20, padded = (bytes + align - 1) & ~(align - 1);
        void *p = nullptr;
        if (posix_memalign(&p, align, padded) != 0 || !p) return fail(NBR_ERR_INVALID, "nbr_host_alloc: out of host memory");
        madvise(p, padded, MADV_HUGEPAGE);
        *out = p;
    }
    std::lock_guard<std::mutex> lock(g_host_alloc_mutex);
    g_host_allocs.emplace_back(*out, pinned);
    return NBR_OK;
}

extern "C" int nbr_host_free(void *ptr)
{
    if (!ptr) return NBR_OK;
    int kind = -1;
    {
        std::lock_guard<std::mutex> lock(g_host_alloc_mutex);
        for (size_t i = 0; i < g_host_allocs.size(); ++i)
            if (g_host_allocs[i].first == ptr) { kind = g_host_allocs[i].second; g_host_allocs.erase(g_host_allocs.begin() + i); break; }
    }
    if (kind < 0) return fail(NBR_ERR_INVALID, "nbr_host_free: not a pointer from nbr_host_alloc");
    if (kind) NBR_CUDA(cudaFreeHost(ptr));
    else free(ptr);
    return NBR_OK;
}

namespace nbr {

int tile_step_plan(Mailbox *M, const void *xyz, int dtype, int64_t n, const double *edges_host, const double *radii_host,
                   int32_t n_scales, int32_t descriptor_mask, double *boxes_host_out, Scratch &perm, Scratch &sorted, Plan **P_out,
                   cudaStream_t s);

// float32 rows -> float64 rows on the device (the share of a batch that crosses the wire as float64, see rows_to_host)
__global__ void widen_rows_kernel(const float *__restrict__ src, double *__restrict__ dst, size_t n)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) dst[i] = (double)src[i];
}

// features of the queries qdev[0, n_query) (device, any order) against plan P -> the caller's HOST rows, in batches:
// kernels of batch b+1 | device->host pieces of batch b | host threads moving / widening the pieces that have landed
static int rows_to_host(const Plan *P, const void *qdev_all, int q_dtype, int64_t n_query, const double *qbox, void *out_host,
                        int out_dtype, int descriptor_mask, int n_scales, cudaStream_t stream, cudaStream_t copy_stream)
{
    const bool wire_f64 = getenv("NBR_HOST_WIRE") && std::string(getenv("NBR_HOST_WIRE")) == "f64";
    const int wire = (out_dtype == NBR_F64 && wire_f64) ? NBR_F64 : NBR_F32;
    const bool widen_rows = out_dtype == NBR_F64 && wire == NBR_F32;
    const int ncol = (descriptor_mask & NBR_DESC_EXTENDED) ? NBR_COLS_EXTENDED : NBR_COLS_REFERENCE;
    const size_t cols = (size_t)ncol * n_scales;
    const size_t qrow = 3 * esize(q_dtype), wrow = cols * esize(wire), orow = cols * esize(out_dtype);
    const char *batch_env = getenv("NBR_HOST_BATCH_ROWS");
    const int64_t batch = batch_env && atoll(batch_env) > 0 ? (int64_t)atoll(batch_env) : std::max<int64_t>(262144, (n_query + 15) / 16);
    const int n_batches = (int)ceil_div(n_query, batch);
    constexpr int RING = 3;
    const bool out_pinned = is_pinned(out_host);
    const bool direct = out_pinned && !widen_rows;          // rows land in the caller's buffer straight from the device
    if (!out_pinned && orow * (size_t)n_query >= (8u << 20)) {
        // a fresh pageable result (np.empty) is faulted in by the threads that fill it; huge pages cut the fault count
        const uintptr_t a = ((uintptr_t)out_host + (2u << 20) - 1) & ~(uintptr_t)((2u << 20) - 1);
        const uintptr_t b = ((uintptr_t)out_host + orow * (size_t)n_query) & ~(uintptr_t)((2u << 20) - 1);
        if (b > a) madvise((void *)a, b - a, MADV_HUGEPAGE);
    }
    // a batch leaves the device in PIECES of a few MB, each with its own event: the host threads follow the DMA piece
    // by piece, and the tail after the last copy is one piece, not one batch
    const char *piece_env = getenv("NBR_HOST_PIECE_MB");
    const double piece_mb = piece_env && atof(piece_env) > 0 ? atof(piece_env) : 4.0;
    // ROUTES of a batch of float64 rows into a PINNED result.  "wire": float32 over the link into the pinned ring, widened
    // by the host threads (4 bytes of host DRAM traffic per wire byte: the DMA write, the read, the doubled write).
    // "device": widened on the device, float64 straight into the caller's rows (twice the link bytes, no host work).
    // which one is cheaper depends on what the box is short of -- host memory / cores (one rank: 24.2 ms all-wire, 23.2
    // with a tenth on the device route; 4 ranks on a box with fast links: 69 ms all-wire, 43 ms all-device) or link
    // bandwidth (4 ranks on a box with 65 GB/s for all links: all-device 95 ms) -- so the route is chosen per batch, when
    // its rows exist: wire if the pinned slot it needs is drained; device if that slot's rows have LANDED but the host
    // threads are still behind; if the slot's rows have not even landed the link is the bottleneck and the batch waits
    // for it.  NBR_HOST_WIDEN=host / device pins the route.  same values either way (float32 results).
    enum { ROUTE_AUTO, ROUTE_WIRE, ROUTE_DEVICE };
    const char *widen_env = getenv("NBR_HOST_WIDEN");
    int route_mode = ROUTE_WIRE;
    if (widen_rows && out_pinned)
        route_mode = !widen_env ? ROUTE_AUTO : (std::string(widen_env) == "device" ? ROUTE_DEVICE : (std::string(widen_env) == "host" ? ROUTE_WIRE : ROUTE_AUTO));
    const size_t piece_rows = direct ? (size_t)batch : std::max<size_t>(1, (size_t)(piece_mb * (1 << 20)) / wrow);
    const int pieces_per_batch = (int)ceil_div(batch, (int64_t)piece_rows);
    // computed[s]: the kernels of the batch in device slot s are done; freed[s]: the slot's rows have been read
    std::vector<cudaEvent_t> computed(RING, nullptr), freed(RING, nullptr), landed((size_t)RING * pieces_per_batch, nullptr);
    int rc = NBR_OK;
    cudaError_t e = cudaSuccess;
    for (auto &ev : computed)
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
    for (auto &ev : freed)
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
    for (auto &ev : landed)
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
    {
        Scratch o, o64;
        void *pin[RING] = {nullptr, nullptr, nullptr};
        if (e == cudaSuccess) rc = o.alloc((size_t)std::min<int64_t>(n_query, batch * RING) * wrow, stream);
        if (!rc && e == cudaSuccess && route_mode != ROUTE_WIRE) rc = o64.alloc((size_t)std::min<int64_t>(n_query, batch * RING) * orow, stream);
        if (!rc && e == cudaSuccess && !direct && route_mode != ROUTE_DEVICE)
            for (int k = 0; k < std::min(RING, n_batches) && !rc; ++k) rc = g_ring_out.get(k, (size_t)batch * wrow, &pin[k]);

        // ---- consumers: every host thread walks the batches in order; a wire batch's pieces are awaited one by one and
        // the thread's stripe of each goes from the pinned ring into the caller's rows (widening if asked)
        // route[b]: 0 not routed yet, 1 + p wire through pinned slot p, -1 device route (nothing to do for the host)
        std::vector<std::atomic<int>> route(n_batches), consumed(n_batches);
        for (int b = 0; b < n_batches; ++b) { route[b].store(0); consumed[b].store(0); }
        std::atomic<int> abort_flag{0};
        const auto nap = [] { std::this_thread::sleep_for(std::chrono::microseconds(20)); };
        const std::function<void(int, int)> consumer = [&](int part, int parts) {
            for (int b = 0; b < n_batches; ++b) {
                int r;
                while ((r = route[b].load(std::memory_order_acquire)) == 0) {
                    if (abort_flag.load(std::memory_order_relaxed)) return;
                    nap();
                }
                if (r > 0) {
                    const int slot = r - 1;
                    const int64_t first = (int64_t)b * batch, n = std::min(batch, n_query - first);
                    const char *src = (const char *)pin[slot];
                    char *dst = (char *)out_host + (size_t)first * orow;
                    for (int64_t r0 = 0, pc = 0; r0 < n; r0 += (int64_t)piece_rows, ++pc) {
                        const size_t rows = (size_t)std::min<int64_t>((int64_t)piece_rows, n - r0);
                        if (cudaEventSynchronize(landed[(size_t)slot * pieces_per_batch + pc]) != cudaSuccess) { abort_flag.store(1); return; }
                        size_t a, c;
                        part_range(rows * cols, 1, part, parts, &a, &c);
                        if (!c) continue;
                        a += (size_t)r0 * cols;
                        if (widen_rows) widen((const float *)src + a, (double *)dst + a, c);
                        else memcpy(dst + a * esize(out_dtype), src + a * esize(out_dtype), c * esize(out_dtype));
                    }
                }
                consumed[b].fetch_add(1, std::memory_order_release);
            }
        };
        const bool use_consumers = !rc && e == cudaSuccess && !direct && route_mode != ROUTE_DEVICE;
        const int n_consumers = std::max(1, HostPool::get().workers());
        if (use_consumers) HostPool::get().start(consumer);

        int wire_count = 0, device_count = 0;
        int pin_user[RING] = {-1, -1, -1};                      // the batch that went through pinned slot p last
        // hands batch j (computed into device slot j % RING) to the copy stream
        const auto ship = [&](int j) {
            const int slot = j % RING;
            const int64_t first = (int64_t)j * batch, n = std::min(batch, n_query - first);
            char *odev = (char *)o.ptr + (size_t)slot * batch * wrow;
            if (direct) {
                e = cudaStreamWaitEvent(copy_stream, computed[slot], 0);
                if (e == cudaSuccess) e = cudaMemcpyAsync((char *)out_host + (size_t)first * orow, odev, (size_t)n * wrow, cudaMemcpyDeviceToHost, copy_stream);
                if (e == cudaSuccess) e = cudaEventRecord(freed[slot], copy_stream);
                return;
            }
            const int p = wire_count % RING;
            bool to_device = route_mode == ROUTE_DEVICE;
            if (route_mode == ROUTE_AUTO) {
                e = cudaEventSynchronize(computed[slot]);          // decide when the rows exist, not when they are queued
                while (e == cudaSuccess && !abort_flag.load(std::memory_order_relaxed)) {
                    const int w = pin_user[p];
                    if (w < 0 || consumed[w].load(std::memory_order_acquire) >= n_consumers) break;
                    const int64_t n_w = std::min(batch, n_query - (int64_t)w * batch);
                    const cudaError_t q = cudaEventQuery(landed[(size_t)p * pieces_per_batch + ceil_div(n_w, (int64_t)piece_rows) - 1]);
                    if (q == cudaSuccess) { to_device = true; break; }        // landed, not drained: the host is behind
                    if (q != cudaErrorNotReady) { e = q; break; }
                    nap();                                                     // not landed: the link is behind
                }
            } else if (route_mode == ROUTE_WIRE) {
                const int w = pin_user[p];
                if (w >= 0 && use_consumers)
                    while (consumed[w].load(std::memory_order_acquire) < n_consumers && !abort_flag.load(std::memory_order_relaxed)) nap();
            }
            if (e != cudaSuccess) return;
            e = cudaStreamWaitEvent(copy_stream, computed[slot], 0);
            if (to_device) {
                // everything of this route is ordered on the copy stream: the float64 slot's previous copy is ahead of
                // the kernel that overwrites it
                char *odev64 = (char *)o64.ptr + (size_t)(device_count % RING) * batch * orow;
                ++device_count;
                const size_t elems = (size_t)n * cols;
                if (e == cudaSuccess) {
                    widen_rows_kernel<<<(unsigned)std::min<size_t>(ceil_div(elems, (size_t)256), (size_t)device_sm_count() * 8), 256, 0, copy_stream>>>(
                        (const float *)odev, (double *)odev64, elems);
                    g_launches.fetch_add(1, std::memory_order_relaxed);
                    e = cudaGetLastError();                 // no early return here: the consumers are running
                }
                if (e == cudaSuccess) e = cudaEventRecord(freed[slot], copy_stream);
                if (e == cudaSuccess) e = cudaMemcpyAsync((char *)out_host + (size_t)first * orow, odev64, (size_t)n * orow, cudaMemcpyDeviceToHost, copy_stream);
                if (e == cudaSuccess) route[j].store(-1, std::memory_order_release);
            } else {
                ++wire_count;
                pin_user[p] = j;
                for (int64_t r0 = 0, pc = 0; e == cudaSuccess && r0 < n; r0 += (int64_t)piece_rows, ++pc) {
                    const size_t rows = (size_t)std::min<int64_t>((int64_t)piece_rows, n - r0);
                    e = cudaMemcpyAsync((char *)pin[p] + (size_t)r0 * wrow, odev + (size_t)r0 * wrow, rows * wrow, cudaMemcpyDeviceToHost, copy_stream);
                    if (e == cudaSuccess) e = cudaEventRecord(landed[(size_t)p * pieces_per_batch + pc], copy_stream);
                }
                if (e == cudaSuccess) e = cudaEventRecord(freed[slot], copy_stream);
                if (e == cudaSuccess) route[j].store(1 + p, std::memory_order_release);
            }
        };

        // trip b launches the kernels of batch b, then ships batch b - 1 (whose route may have to wait for its rows)
        for (int b = 0; !rc && e == cudaSuccess && b <= n_batches; ++b) {
            if (b < n_batches) {
                const int slot = b % RING;
                const int64_t first = (int64_t)b * batch, n = std::min(batch, n_query - first);
                if (b >= RING) e = cudaStreamWaitEvent(stream, freed[slot], 0);     // recorded when batch b - RING was shipped
                if (e != cudaSuccess) break;
                rc = plan_run(P, (const char *)qdev_all + (size_t)first * qrow, q_dtype, n, qbox, (char *)o.ptr + (size_t)slot * batch * wrow, wire, stream);
                if (rc) break;
                e = cudaEventRecord(computed[slot], stream);
            }
            if (b >= 1 && e == cudaSuccess) ship(b - 1);
        }
        if (rc || e != cudaSuccess) abort_flag.store(1);
        if (use_consumers) HostPool::get().wait();
        if (!rc && abort_flag.load() && e == cudaSuccess) rc = fail(NBR_ERR_CUDA, "host path: a device->host copy failed");
        if (!rc && e == cudaSuccess) e = cudaStreamSynchronize(stream);
        if (!rc && e == cudaSuccess) e = cudaStreamSynchronize(copy_stream);
        if (!rc && e != cudaSuccess) rc = fail(NBR_ERR_CUDA, std::string("host path: ") + cudaGetErrorString(e));
        cudaStreamSynchronize(copy_stream);
        cudaStreamSynchronize(stream);
        if (getenv("NBR_HOST_STATS")) fprintf(stderr, "[nbr host path] %d batches: %d wire, %d device route\n", n_batches, wire_count, device_count);
    }
    for (auto ev : freed) if (ev) cudaEventDestroy(ev);
    for (auto ev : computed) if (ev) cudaEventDestroy(ev);
    for (auto ev : landed) if (ev) cudaEventDestroy(ev);
    return rc;
}

struct TwoStreams {
    cudaStream_t stream = nullptr, copy_stream = nullptr;
    int create()
    {
        NBR_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
        NBR_CUDA(cudaStreamCreateWithFlags(&copy_stream, cudaStreamNonBlocking));
        return NBR_OK;
    }
    ~TwoStreams()
    {
        if (stream) { cudaStreamSynchronize(stream); cudaStreamDestroy(stream); }
        if (copy_stream) { cudaStreamSynchronize(copy_stream); cudaStreamDestroy(copy_stream); }
    }
};

}  // namespace nbr

using namespace nbr;

extern "C" int nbr_multiscale_features_host(const void *query_host, int q_dtype, int64_t n_query,
                                            const void *search_host, int s_dtype, int64_t n_search,
                                            const double *edges_host, const double *radii_host, int32_t n_scales,
                                            void *out_host, int out_dtype, int32_t descriptor_mask,
                                            int64_t *n_voxels_host)
{
    if ((q_dtype != NBR_F32 && q_dtype != NBR_F64) || (s_dtype != NBR_F32 && s_dtype != NBR_F64))
        return fail(NBR_ERR_INVALID, "nbr_multiscale_features_host: dtype must be NBR_F32 or NBR_F64");
    if (out_dtype != NBR_F32 && out_dtype != NBR_F64) return fail(NBR_ERR_INVALID, "nbr_multiscale_features_host: bad out_dtype");
    if (n_search < 2) return fail(NBR_ERR_TOO_FEW_POINTS, "need at least 2 points to define a voxel grid");
    if (n_query <= 0 || n_scales <= 0) return NBR_OK;
    if (!query_host || !search_host || !out_host) return fail(NBR_ERR_INVALID, "nbr_multiscale_features_host: null argument");
    std::lock_guard<std::mutex> one_call(g_host_call);
    const size_t qrow = 3 * esize(q_dtype), sbytes = (size_t)n_search * 3 * esize(s_dtype);
    const bool same = query_host == search_host && q_dtype == s_dtype && n_query == n_search;
    TwoStreams S;
    NBR_TRY(S.create());
    Plan *P = nullptr;
    int rc = NBR_OK;
    {
        Scratch q, s;
        rc = s.alloc(sbytes, S.stream);
        if (!rc && !same) rc = q.alloc((size_t)n_query * qrow, S.stream);
        if (!rc) rc = upload(s.ptr, search_host, sbytes, S.stream);
        if (!rc) rc = plan_create(&P, s.ptr, s_dtype, n_search, edges_host, radii_host, n_scales, descriptor_mask, nullptr, nullptr, S.stream);
        if (!rc && !same) rc = upload(q.ptr, query_host, (size_t)n_query * qrow, S.stream);
        if (!rc) rc = rows_to_host(P, same ? s.ptr : q.ptr, q_dtype, n_query, same ? P->local_box : nullptr, out_host, out_dtype,
                                   descriptor_mask, n_scales, S.stream, S.copy_stream);
        if (!rc && n_voxels_host) rc = plan_voxel_counts(P, n_voxels_host);
        cudaStreamSynchronize(S.stream);
        delete P;
    }
    return rc;
}

// one step of a rank with HOST buffers: tile up, box table / halo push / lattices as nbr_tile_step, rows down in
// batches (float32 on the wire, widened by host threads for out_dtype NBR_F64).  collective over the ranks.
extern "C" int nbr_tile_step_host(nbr_mailbox *mailbox, const void *xyz_host, int dtype, int64_t n, const double *edges_host,
                                  const double *radii_host, int32_t n_scales, void *out_host, int out_dtype,
                                  int32_t descriptor_mask, double *boxes_host_out)
{
    if (!mailbox || n < 0 || n_scales < 0 || (n > 0 && (!xyz_host || !out_host)))
        return fail(NBR_ERR_INVALID, "nbr_tile_step_host: bad argument");
    if (dtype != NBR_F32 && dtype != NBR_F64) return fail(NBR_ERR_INVALID, "nbr_tile_step_host: dtype must be NBR_F32 or NBR_F64");
    if (out_dtype != NBR_F32 && out_dtype != NBR_F64) return fail(NBR_ERR_INVALID, "nbr_tile_step_host: bad out_dtype");
    std::lock_guard<std::mutex> one_call(g_host_call);
    TwoStreams S;
    NBR_TRY(S.create());
    Plan *P = nullptr;
    int rc = NBR_OK;
    {
        Scratch dev, perm, sorted;
        rc = dev.alloc((size_t)std::max<int64_t>(n, 1) * 3 * esize(dtype), S.stream);
        if (!rc) rc = upload(dev.ptr, xyz_host, (size_t)n * 3 * esize(dtype), S.stream);
        if (!rc) rc = tile_step_plan(reinterpret_cast<Mailbox *>(mailbox), dev.ptr, dtype, n, edges_host, radii_host, n_scales,
                                     descriptor_mask, boxes_host_out, perm, sorted, &P, S.stream);
        if (!rc && P && n > 0 && n_scales > 0)
            rc = rows_to_host(P, dev.ptr, dtype, n, P->local_box, out_host, out_dtype, descriptor_mask, n_scales, S.stream,
                              S.copy_stream);             // the plan's box (tile grown by the halo width) bounds every batch
        cudaStreamSynchronize(S.stream);
        delete P;
    }
    return rc;
}
