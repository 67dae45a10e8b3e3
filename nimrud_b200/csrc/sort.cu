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

// stable scatter of one tile.  ranks come from warp match_any; the tile is first reordered by digit in
// shared memory, so that the global writes of each digit are contiguous runs (coalesced).
// dynamic shared memory: keys[SORT_TILE] (+ vals[SORT_TILE]) + warp_hist[SORT_WARPS][RADIX] + digit tables.
template <bool HAS_VALUES>
__global__ void __launch_bounds__(SORT_THREADS)
radix_scatter_kernel(const uint64_t *__restrict__ keys_in, uint64_t *__restrict__ keys_out,
                     const uint32_t *__restrict__ vals_in, uint32_t *__restrict__ vals_out, int64_t n,
                     int shift, const uint32_t *__restrict__ offsets, int64_t tiles)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint64_t *s_keys = reinterpret_cast<uint64_t *>(smem_raw);
    uint32_t *s_vals = reinterpret_cast<uint32_t *>(s_keys + SORT_TILE);
    uint32_t *s_hist = HAS_VALUES ? s_vals + SORT_TILE : s_vals;            // [SORT_WARPS][RADIX]
    uint32_t *s_start = s_hist + SORT_WARPS * RADIX;                        // [RADIX] first local slot of a digit
    uint32_t *s_gbase = s_start + RADIX;                                    // [RADIX] global slot of that first one
    __shared__ uint32_t s_scan[33];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < SORT_WARPS * RADIX; i += SORT_THREADS) s_hist[i] = 0;
    __syncthreads();

    const int64_t tile_base = (int64_t)blockIdx.x * SORT_TILE;
    const int64_t warp_base = tile_base + (int64_t)warp * (SORT_ROUNDS * 32);
    const int tile_n = (int)min((int64_t)SORT_TILE, n - tile_base);
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
        const uint32_t prev = valid ? s_hist[warp * RADIX + d] : 0u;
        __syncwarp();
        if (valid && (peers & lt) == 0) s_hist[warp * RADIX + d] = prev + __popc(peers);
        __syncwarp();
        rank[r] = (uint16_t)(prev + __popc(peers & lt));
    }
    __syncthreads();
    {
        // one digit per thread: exclusive prefix over the warps, the tile total, then the local start
        const int d = threadIdx.x;
        uint32_t running = 0;
#pragma unroll
        for (int w = 0; w < SORT_WARPS; ++w) {
            const uint32_t c = s_hist[w * RADIX + d];
            s_hist[w * RADIX + d] = running;
            running += c;
        }
        uint32_t total;
        const uint32_t start = block_exclusive_scan<uint32_t>(running, s_scan, &total);
        s_start[d] = start;
        s_gbase[d] = offsets[(int64_t)d * tiles + blockIdx.x];
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < SORT_ROUNDS; ++r) {
        const int64_t i = warp_base + r * 32 + lane;
        if (i < n) {
            const uint32_t d = (uint32_t)(key[r] >> shift) & (RADIX - 1);
            const uint32_t pos = s_start[d] + s_hist[warp * RADIX + d] + rank[r];
            s_keys[pos] = key[r];
            if (HAS_VALUES) s_vals[pos] = vals_in[i];
        }
    }
    __syncthreads();
    for (int pos = threadIdx.x; pos < tile_n; pos += SORT_THREADS) {
        const uint64_t k = s_keys[pos];
        const uint32_t d = (uint32_t)(k >> shift) & (RADIX - 1);
        const uint32_t g = s_gbase[d] + ((uint32_t)pos - s_start[d]);
        keys_out[g] = k;
        if (HAS_VALUES) vals_out[g] = s_vals[pos];
    }
}

static size_t scatter_smem(bool has_values)
{
    return sizeof(uint64_t) * SORT_TILE + (has_values ? sizeof(uint32_t) * SORT_TILE : 0) +
           sizeof(uint32_t) * (SORT_WARPS * RADIX + 2 * RADIX);
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
    const size_t smem = scatter_smem(vals != nullptr);
    static std::atomic<uint64_t> configured{0};
    if (first_use_on_device(configured)) {
        NBR_CUDA(cudaFuncSetAttribute(radix_scatter_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)scatter_smem(true)));
        NBR_CUDA(cudaFuncSetAttribute(radix_scatter_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)scatter_smem(false)));
    }
    for (int shift = begin_bit; shift < end_bit; shift += RADIX_BITS) {
        radix_hist_kernel<<<(unsigned)tiles, SORT_THREADS, 0, stream>>>(src, n, shift, c, tiles);
        NBR_LAUNCHED();
        NBR_TRY((exclusive_scan<uint32_t, uint32_t>(c, c, RADIX * tiles, stream)));
        if (vals)
            radix_scatter_kernel<true><<<(unsigned)tiles, SORT_THREADS, smem, stream>>>(src, dst, vsrc, vdst, n, shift, c, tiles);
        else
            radix_scatter_kernel<false><<<(unsigned)tiles, SORT_THREADS, smem, stream>>>(src, dst, nullptr, nullptr, n, shift, c, tiles);
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
