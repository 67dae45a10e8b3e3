// scan.cuh -- device-wide exclusive prefix sums (the "cell-offset scan" of the spatial index).
// three-phase: per-tile reduce -> scan of tile sums (recursive) -> per-tile scan + offset.
#pragma once
#include "common.cuh"

namespace nbr {

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

#ifdef __CUDACC__
// exclusive scan of one value per thread across the block; returns the exclusive prefix and the
// block total through *total.  `smem` holds 33 T.
template <typename T>
__device__ __forceinline__ T block_exclusive_scan(T v, T *smem, T *total)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    T inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        T t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) smem[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        const int nw = (blockDim.x + 31) >> 5;
        T w = lane < nw ? smem[lane] : T(0);
        T winc = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            T t = __shfl_up_sync(0xffffffffu, winc, o);
            if (lane >= o) winc += t;
        }
        if (lane < nw) smem[lane] = winc - w;
        if (lane == nw - 1) smem[32] = winc;
    }
    __syncthreads();
    T res = smem[warp] + inc - v;
    *total = smem[32];
    __syncthreads();
    return res;
}
#endif

template <typename TIn, typename TOut>
int exclusive_scan(const TIn *in, TOut *out, int64_t n, cudaStream_t stream);

// out[i] = in[i] ? (number of non-zero entries before i) + 1 : 0 ; total non-zero count -> *count_dev
int flags_to_slots(uint32_t *flags_inout, int64_t n, uint32_t *count_dev, cudaStream_t stream);

}  // namespace nbr
