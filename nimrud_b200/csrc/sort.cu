// sort.cu -- hand-written LSD radix sort (8-bit digits) of 64-bit keys, optional 32-bit payload.
// This is the sort inside np.unique (utils/geometry.py:150) and the query ordering sort.
//
// per pass:   histogram (tile x digit counts, digit-major)  ->  exclusive scan  ->  stable scatter.
// the scatter ranks keys with warp match_any (keys of a warp are a contiguous run of the tile, so
// (warp, round, lane) order is input order and the sort is stable).
#include "common.cuh"
#include "scan.cuh"

namespace nbr {

constexpr int SORT_THREADS = 256;
constexpr int SORT_WARPS = SORT_THREADS / 32;
constexpr int SORT_ROUNDS = 16;                               // keys per thread
constexpr int SORT_TILE = SORT_THREADS * SORT_ROUNDS;         // 4096 keys per block
constexpr int RADIX_BITS = 8;
constexpr int RADIX = 1 << RADIX_BITS;

__global__ void __launch_bounds__(SORT_THREADS)
radix_hist_kernel(const uint64_t *__restrict__ keys, int64_t n, int shift, uint32_t *__restrict__ counts,
                  int64_t tiles)
{
    __shared__ uint32_t hist[RADIX];
    hist[threadIdx.x] = 0;
    __syncthreads();
    const int64_t base = (int64_t)blockIdx.x * SORT_TILE;
#pragma unroll 4
    for (int k = 0; k < SORT_ROUNDS; ++k) {
        int64_t i = base + k * SORT_THREADS + threadIdx.x;
        if (i < n) atomicAdd(&hist[(keys[i] >> shift) & (RADIX - 1)], 1u);
    }
    __syncthreads();
    counts[(int64_t)threadIdx.x * tiles + blockIdx.x] = hist[threadIdx.x];
}

template <bool HAS_VALUES>
__global__ void __launch_bounds__(SORT_THREADS)
radix_scatter_kernel(const uint64_t *__restrict__ keys_in, uint64_t *__restrict__ keys_out,
                     const uint32_t *__restrict__ vals_in, uint32_t *__restrict__ vals_out, int64_t n,
                     int shift, const uint32_t *__restrict__ offsets, int64_t tiles)
{
    __shared__ uint32_t warp_hist[SORT_WARPS][RADIX];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < SORT_WARPS * RADIX; i += SORT_THREADS) (&warp_hist[0][0])[i] = 0;
    __syncthreads();

    const int64_t warp_base = (int64_t)blockIdx.x * SORT_TILE + (int64_t)warp * (SORT_ROUNDS * 32);
    uint64_t key[SORT_ROUNDS];
    uint16_t rank[SORT_ROUNDS];
    const uint32_t lt = lanemask_lt();
#pragma unroll
    for (int r = 0; r < SORT_ROUNDS; ++r) {
        const int64_t i = warp_base + r * 32 + lane;
        const bool valid = i < n;
        key[r] = valid ? keys_in[i] : ~0ull;
        const uint32_t d = (uint32_t)(key[r] >> shift) & (RADIX - 1);
        const uint32_t peers = __match_any_sync(0xffffffffu, valid ? d : 0x100u);
        const uint32_t prev = valid ? warp_hist[warp][d] : 0u;
        __syncwarp();
        if (valid && (peers & lt) == 0) warp_hist[warp][d] = prev + __popc(peers);
        __syncwarp();
        rank[r] = (uint16_t)(prev + __popc(peers & lt));
    }
    __syncthreads();
    {
        const int d = threadIdx.x;   // one digit per thread
        uint32_t running = offsets[(int64_t)d * tiles + blockIdx.x];
#pragma unroll
        for (int w = 0; w < SORT_WARPS; ++w) {
            uint32_t c = warp_hist[w][d];
            warp_hist[w][d] = running;
            running += c;
        }
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < SORT_ROUNDS; ++r) {
        const int64_t i = warp_base + r * 32 + lane;
        if (i < n) {
            const uint32_t d = (uint32_t)(key[r] >> shift) & (RADIX - 1);
            const uint32_t pos = warp_hist[warp][d] + rank[r];
            keys_out[pos] = key[r];
            if (HAS_VALUES) vals_out[pos] = vals_in[i];
        }
    }
}

static int radix_sort(uint64_t *keys, uint64_t *keys_tmp, uint32_t *vals, uint32_t *vals_tmp, int64_t n,
                      int begin_bit, int end_bit, cudaStream_t stream)
{
    if (n < 0 || begin_bit < 0 || end_bit > 64 || begin_bit > end_bit)
        return fail(NBR_ERR_INVALID, "radix_sort: bad arguments");
    if (n >= (int64_t)1 << 32) return fail(NBR_ERR_UNSUPPORTED, "radix_sort: n >= 2^32");
    if (n <= 1 || begin_bit == end_bit) return NBR_OK;
    const int64_t tiles = ceil_div(n, SORT_TILE);
    Scratch counts;
    NBR_TRY(counts.alloc(sizeof(uint32_t) * RADIX * tiles, stream));
    uint32_t *c = counts.as<uint32_t>();
    uint64_t *src = keys, *dst = keys_tmp;
    uint32_t *vsrc = vals, *vdst = vals_tmp;
    for (int shift = begin_bit; shift < end_bit; shift += RADIX_BITS) {
        radix_hist_kernel<<<(unsigned)tiles, SORT_THREADS, 0, stream>>>(src, n, shift, c, tiles);
        NBR_LAUNCHED();
        NBR_TRY((exclusive_scan<uint32_t, uint32_t>(c, c, RADIX * tiles, stream)));
        if (vals)
            radix_scatter_kernel<true><<<(unsigned)tiles, SORT_THREADS, 0, stream>>>(src, dst, vsrc, vdst, n, shift, c, tiles);
        else
            radix_scatter_kernel<false><<<(unsigned)tiles, SORT_THREADS, 0, stream>>>(src, dst, nullptr, nullptr, n, shift, c, tiles);
        NBR_LAUNCHED();
        uint64_t *t = src; src = dst; dst = t;
        uint32_t *vt = vsrc; vsrc = vdst; vdst = vt;
    }
    if (src != keys) {
        NBR_CUDA(cudaMemcpyAsync(keys, src, sizeof(uint64_t) * n, cudaMemcpyDeviceToDevice, stream));
        if (vals) NBR_CUDA(cudaMemcpyAsync(vals, vsrc, sizeof(uint32_t) * n, cudaMemcpyDeviceToDevice, stream));
    }
    return NBR_OK;
}

int sort_keys(uint64_t *keys, uint64_t *tmp, int64_t n, int begin_bit, int end_bit, cudaStream_t stream)
{
    return radix_sort(keys, tmp, nullptr, nullptr, n, begin_bit, end_bit, stream);
}

int sort_pairs(uint64_t *keys, uint64_t *keys_tmp, uint32_t *vals, uint32_t *vals_tmp, int64_t n,
               int begin_bit, int end_bit, cudaStream_t stream)
{
    return radix_sort(keys, keys_tmp, vals, vals_tmp, n, begin_bit, end_bit, stream);
}

}  // namespace nbr

extern "C" int nbr_sort_u64(uint64_t *keys, uint64_t *tmp, int64_t n, int begin_bit, int end_bit, void *stream)
{
    if (n <= 1) return n < 0 ? nbr::fail(NBR_ERR_INVALID, "nbr_sort_u64: negative n") : NBR_OK;
    if (!keys || !tmp) return nbr::fail(NBR_ERR_INVALID, "nbr_sort_u64: null buffer");
    return nbr::sort_keys(keys, tmp, n, begin_bit, end_bit, (cudaStream_t)stream);
}

extern "C" int nbr_sort_pairs_u64_u32(uint64_t *keys, uint64_t *keys_tmp, uint32_t *vals, uint32_t *vals_tmp,
                                      int64_t n, int begin_bit, int end_bit, void *stream)
{
    if (n <= 1) return n < 0 ? nbr::fail(NBR_ERR_INVALID, "nbr_sort_pairs: negative n") : NBR_OK;
    if (!keys || !keys_tmp || !vals || !vals_tmp) return nbr::fail(NBR_ERR_INVALID, "nbr_sort_pairs: null buffer");
    return nbr::sort_pairs(keys, keys_tmp, vals, vals_tmp, n, begin_bit, end_bit, (cudaStream_t)stream);
}
