// scan.cu -- device-wide exclusive prefix sums.
#include "scan.cuh"

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
