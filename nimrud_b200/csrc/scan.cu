// scan.cu -- device-wide exclusive prefix sums.
#include "scan.cuh"

#include <stdlib.h>

#include <type_traits>

namespace nbr {

template <typename TIn, typename TOut>
__global__ void __launch_bounds__(SCAN_THREADS)
scan_reduce_kernel(const TIn *__restrict__ in, TOut *__restrict__ tile_sums, int64_t n)
{
    __shared__ TOut smem[33];
    const int64_t base = (int64_t)blockIdx.x * SCAN_TILE;
    TOut acc = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        int64_t i = base + k * SCAN_THREADS + threadIdx.x;   // coalesced
        if (i < n) acc += (TOut)in[i];
    }
    TOut total;
    block_exclusive_scan<TOut>(acc, smem, &total);
    if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}

// each thread owns SCAN_ITEMS consecutive elements (blocked arrangement) so the scan is a plain
// thread-local prefix plus one block scan.
template <typename TIn, typename TOut, bool FLAGS>
__global__ void __launch_bounds__(SCAN_THREADS)
scan_apply_kernel(const TIn *in, TOut *out, const TOut *__restrict__ tile_offsets, int64_t n)
{
    __shared__ TOut smem[33];
    const int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
    TOut v[SCAN_ITEMS];
    TOut acc = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        int64_t i = base + k;
        TOut x = i < n ? (TOut)in[i] : TOut(0);
        if (FLAGS) x = x ? TOut(1) : TOut(0);
        v[k] = x;
        acc += x;
    }
    TOut total;
    TOut prefix = block_exclusive_scan<TOut>(acc, smem, &total);
    prefix += tile_offsets ? tile_offsets[blockIdx.x] : TOut(0);
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        int64_t i = base + k;
        if (i < n) out[i] = FLAGS ? (v[k] ? prefix + 1 : TOut(0)) : prefix;
        prefix += v[k];
    }
}

template <typename T>
__global__ void scan_flag_reduce_kernel(const T *__restrict__ in, T *__restrict__ tile_sums, int64_t n)
{
    __shared__ T smem[33];
    const int64_t base = (int64_t)blockIdx.x * SCAN_TILE;
    T acc = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        int64_t i = base + k * SCAN_THREADS + threadIdx.x;
        if (i < n) acc += in[i] ? T(1) : T(0);
    }
    T total;
    block_exclusive_scan<T>(acc, smem, &total);
    if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}

// ------------------------------------------------------------------------------------------------
// single-pass scan of 32-bit counters with decoupled look-back: every tile publishes its aggregate, then
// its inclusive prefix, in one 64-bit word (status << 32 | value); a tile sums the aggregates of its
// predecessors (a warp reads 32 of them at a time) until it meets a published prefix.  tiles are handed out
// by an atomic ticket, so a tile only ever waits for tiles that are already running.  one read and one
// write of the data instead of two reads and one write, one launch instead of five.
//   FLAGS: out[i] = in[i] ? (number of non-zero entries up to and including i) : 0   (flags -> 1-based slots)
//   else : out[i] = sum of in[0..i)                                                  (exclusive scan)
// in == out is allowed (a thread reads its items before any thread of the tile writes).
// ------------------------------------------------------------------------------------------------
static bool lookback_disabled()
{
    static const bool off = getenv("NBR_NO_LOOKBACK_SCAN") != nullptr;
    return off;
}

constexpr unsigned long long LB_AGGREGATE = 1ull << 32, LB_PREFIX = 2ull << 32;

template <bool FLAGS>
__global__ void __launch_bounds__(SCAN_THREADS)
scan_lookback_kernel(const uint32_t *in, uint32_t *out, int64_t n, unsigned long long *state, unsigned int *ticket,
                     uint32_t *total_out, int64_t tiles)
{
    __shared__ uint32_t smem[33];
    __shared__ unsigned int s_tile;
    __shared__ uint32_t s_prefix;
    if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1u);
    __syncthreads();
    const int64_t tile = s_tile;
    const int64_t base = tile * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
    uint32_t v[SCAN_ITEMS];
    uint32_t acc = 0;
    // a thread owns SCAN_ITEMS consecutive items: 16-byte loads / stores when the arrays are aligned and the
    // thread's chunk is whole (scalar accesses touch every 32-byte sector SCAN_ITEMS times)
    static_assert(SCAN_ITEMS % 4 == 0, "vector path");
    const bool vec = (((uintptr_t)in | (uintptr_t)out) & 15) == 0 && base + SCAN_ITEMS <= n;
    if (vec) {
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS; k += 4) {
            const uint4 q = *reinterpret_cast<const uint4 *>(in + base + k);
            v[k] = q.x; v[k + 1] = q.y; v[k + 2] = q.z; v[k + 3] = q.w;
        }
    } else {
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS; ++k) {
            const int64_t i = base + k;
            v[k] = i < n ? in[i] : 0u;
        }
    }
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        if (FLAGS) v[k] = v[k] ? 1u : 0u;
        acc += v[k];
    }
    uint32_t total;
    uint32_t prefix = block_exclusive_scan<uint32_t>(acc, smem, &total);
    if (threadIdx.x < 32) {
        const int lane = threadIdx.x;
        uint32_t run = 0;
        if (tile == 0) {
            if (lane == 0) atomicExch(&state[0], LB_PREFIX | total);
        } else {
            if (lane == 0) atomicExch(&state[tile], LB_AGGREGATE | total);
            int64_t hi = tile - 1;                            // the window is tiles [hi - 31, hi], lane l reads hi - l
            for (;;) {
                const int64_t t = hi - lane;
                unsigned long long st = LB_PREFIX;                // tiles before the first count as an empty prefix
                if (t >= 0) {
                    do { st = *reinterpret_cast<volatile unsigned long long *>(&state[t]); } while ((st >> 32) == 0);
                }
                const uint32_t is_prefix = __ballot_sync(0xffffffffu, (st >> 32) == 2);
                // the nearest published prefix ends the look-back: add the lanes up to and including it
                const int stop = is_prefix ? __ffs(is_prefix) - 1 : 31;
                const uint32_t mine = (lane <= stop && t >= 0) ? (uint32_t)st : 0u;
                run += __reduce_add_sync(0xffffffffu, mine);
                if (is_prefix) break;
                hi -= 32;
            }
            if (lane == 0) atomicExch(&state[tile], LB_PREFIX | (unsigned long long)(run + total));
        }
        if (lane == 0) {
            s_prefix = run;
            if (tile == tiles - 1 && total_out) *total_out = run + total;
        }
    }
    __syncthreads();
    prefix += s_prefix;
    uint32_t r[SCAN_ITEMS];
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        r[k] = FLAGS ? (v[k] ? prefix + 1 : 0u) : prefix;
        prefix += v[k];
    }
    if (vec) {
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS; k += 4)
            *reinterpret_cast<uint4 *>(out + base + k) = make_uint4(r[k], r[k + 1], r[k + 2], r[k + 3]);
    } else {
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS; ++k)
            if (base + k < n) out[base + k] = r[k];
    }
}

template <bool FLAGS>
static int scan_lookback(const uint32_t *in, uint32_t *out, int64_t n, uint32_t *total_out, cudaStream_t stream)
{
    const int64_t tiles = ceil_div(n, SCAN_TILE);
    Scratch state;
    NBR_TRY(state.alloc(sizeof(unsigned long long) * (tiles + 1), stream));
    NBR_CUDA(cudaMemsetAsync(state.ptr, 0, sizeof(unsigned long long) * (tiles + 1), stream));
    unsigned long long *st = state.as<unsigned long long>();
    scan_lookback_kernel<FLAGS><<<(unsigned)tiles, SCAN_THREADS, 0, stream>>>(in, out, n, st, reinterpret_cast<unsigned int *>(st + tiles),
                                                                              total_out, tiles);
    NBR_LAUNCHED();
    return NBR_OK;
}

template <typename TIn, typename TOut>
int exclusive_scan(const TIn *in, TOut *out, int64_t n, cudaStream_t stream)
{
    if (n <= 0) return NBR_OK;
    const int64_t tiles = ceil_div(n, SCAN_TILE);
    if (tiles == 1) {
        scan_apply_kernel<TIn, TOut, false><<<1, SCAN_THREADS, 0, stream>>>(in, out, nullptr, n);
        NBR_LAUNCHED();
        return NBR_OK;
    }
    if (std::is_same<TIn, uint32_t>::value && std::is_same<TOut, uint32_t>::value && tiles < ((int64_t)1 << 31) && !lookback_disabled())
        return scan_lookback<false>(reinterpret_cast<const uint32_t *>(in), reinterpret_cast<uint32_t *>(out), n, nullptr, stream);
    Scratch sums;
    NBR_TRY(sums.alloc(sizeof(TOut) * tiles, stream));
    scan_reduce_kernel<TIn, TOut><<<(unsigned)tiles, SCAN_THREADS, 0, stream>>>(in, sums.as<TOut>(), n);
    NBR_LAUNCHED();
    NBR_TRY((exclusive_scan<TOut, TOut>(sums.as<TOut>(), sums.as<TOut>(), tiles, stream)));
    scan_apply_kernel<TIn, TOut, false><<<(unsigned)tiles, SCAN_THREADS, 0, stream>>>(in, out, sums.as<TOut>(), n);
    NBR_LAUNCHED();
    return NBR_OK;
}

template int exclusive_scan<uint32_t, uint32_t>(const uint32_t *, uint32_t *, int64_t, cudaStream_t);
template int exclusive_scan<int32_t, int64_t>(const int32_t *, int64_t *, int64_t, cudaStream_t);
template int exclusive_scan<int64_t, int64_t>(const int64_t *, int64_t *, int64_t, cudaStream_t);
template int exclusive_scan<uint64_t, uint64_t>(const uint64_t *, uint64_t *, int64_t, cudaStream_t);

__global__ void slots_total_kernel(const uint32_t *tile_offsets, const uint32_t *last_tile_count, int64_t tiles,
                                   uint32_t *count)
{
    *count = tile_offsets[tiles - 1] + *last_tile_count;
}

int flags_to_slots(uint32_t *flags, int64_t n, uint32_t *count_dev, cudaStream_t stream)
{
    if (n <= 0) {
        NBR_CUDA(cudaMemsetAsync(count_dev, 0, sizeof(uint32_t), stream));
        return NBR_OK;
    }
    const int64_t tiles = ceil_div(n, SCAN_TILE);
    if (!lookback_disabled()) return scan_lookback<true>(flags, flags, n, count_dev, stream);
    Scratch sums;
    NBR_TRY(sums.alloc(sizeof(uint32_t) * (tiles + 1), stream));
    uint32_t *s = sums.as<uint32_t>();
    scan_flag_reduce_kernel<uint32_t><<<(unsigned)tiles, SCAN_THREADS, 0, stream>>>(flags, s, n);
    NBR_LAUNCHED();
    // keep the last tile's own count before the in-place scan overwrites it
    NBR_CUDA(cudaMemcpyAsync(s + tiles, s + tiles - 1, sizeof(uint32_t), cudaMemcpyDeviceToDevice, stream));
    NBR_TRY((exclusive_scan<uint32_t, uint32_t>(s, s, tiles, stream)));
    slots_total_kernel<<<1, 1, 0, stream>>>(s, s + tiles, tiles, count_dev);
    NBR_LAUNCHED();
    scan_apply_kernel<uint32_t, uint32_t, true><<<(unsigned)tiles, SCAN_THREADS, 0, stream>>>(flags, flags, s, n);
    NBR_LAUNCHED();
    return NBR_OK;
}

}  // namespace nbr

extern "C" int nbr_exclusive_scan_u32(const uint32_t *in, uint32_t *out, int64_t n, void *stream)
{
    return nbr::exclusive_scan<uint32_t, uint32_t>(in, out, n, (cudaStream_t)stream);
}

extern "C" int nbr_exclusive_scan_i64(const int64_t *in, int64_t *out, int64_t n, void *stream)
{
    return nbr::exclusive_scan<int64_t, int64_t>(in, out, n, (cudaStream_t)stream);
}
