// order.cu -- spatially coherent processing order of the query cloud: Morton (Z-curve) keys at a
// granularity of a few finest voxels, sorted with the hand-written radix sort.  consecutive queries
// of the order are close in space at every scale, which is what lets one warp of the feature kernel
// share a staged occupancy window.  results never depend on the order (each query is independent).
#include "common.cuh"

namespace nbr {

int sort_pairs(uint64_t *keys, uint64_t *keys_tmp, uint32_t *vals, uint32_t *vals_tmp, int64_t n, int begin_bit,
               int end_bit, cudaStream_t stream);

struct MortonParam {
    double lo[3];
    double inv_cell;
    int bits[3];
    int levels;
    int common;   // levels present on all three axes
    // position of bit l of axis a in the key (axes with fewer bits drop out of the upper levels)
    unsigned char pos[3][21];
};

// spread the low 21 bits of x so that bit i lands at 3*i
__device__ __forceinline__ uint64_t spread3(uint32_t v)
{
    uint64_t x = v & 0x1fffffu;
    x = (x | x << 32) & 0x1f00000000ffffull;
    x = (x | x << 16) & 0x1f0000ff0000ffull;
    x = (x | x << 8) & 0x100f00f00f00f00full;
    x = (x | x << 4) & 0x10c30c30c30c30c3ull;
    x = (x | x << 2) & 0x1249249249249249ull;
    return x;
}

__global__ void __launch_bounds__(256)
morton_kernel(const void *__restrict__ xyz, int dtype, int64_t n, MortonParam P, uint64_t *__restrict__ keys,
              uint32_t *__restrict__ vals)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    // the order only has to be spatially coherent, so float32 quantisation is fine here
    uint32_t c[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        float v = (float)(load_coord(xyz, dtype, i, 3, a) - P.lo[a]) * (float)P.inv_cell;
        const float top = (float)((1u << P.bits[a]) - 1u);
        v = fminf(fmaxf(v, 0.0f), top);
        c[a] = (uint32_t)v;
    }
    // levels shared by all three axes: classic 3-way bit interleave; the rest (axes with more bits) generically
    const uint32_t low = (1u << P.common) - 1u;
    uint64_t key = spread3(c[0] & low) | (spread3(c[1] & low) << 1) | (spread3(c[2] & low) << 2);
#pragma unroll
    for (int a = 0; a < 3; ++a)
        for (int l = P.common; l < P.bits[a]; ++l) key |= (uint64_t)((c[a] >> l) & 1u) << P.pos[a][l];
    keys[i] = key;
    vals[i] = (uint32_t)i;
}

// sorted_out[i] = xyz[perm[i]] (same dtype as the input)
__global__ void __launch_bounds__(256)
gather_kernel(const void *__restrict__ xyz, int dtype, int64_t n, const uint32_t *__restrict__ perm,
              void *__restrict__ sorted_out)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int64_t j = perm[i];
    if (dtype == NBR_F32) {
        const float *src = reinterpret_cast<const float *>(xyz) + j * 3;
        float *dst = reinterpret_cast<float *>(sorted_out) + i * 3;
        dst[0] = src[0]; dst[1] = src[1]; dst[2] = src[2];
    } else {
        const double *src = reinterpret_cast<const double *>(xyz) + j * 3;
        double *dst = reinterpret_cast<double *>(sorted_out) + i * 3;
        dst[0] = src[0]; dst[1] = src[1]; dst[2] = src[2];
    }
}

// perm_out[n]: query indices in Morton order; sorted_xyz_out (optional): the cloud in that order.
// lohi = bounding box of the query cloud (host).
int morton_order(const void *xyz, int dtype, int64_t n, const double lohi[6], double cell, uint32_t *perm_out,
                 void *sorted_xyz_out, cudaStream_t stream)
{
    if (n <= 0) return NBR_OK;
    if (n >= (int64_t)1 << 32) return fail(NBR_ERR_UNSUPPORTED, "morton_order: more than 2^32 queries");
    MortonParam P;
    P.inv_cell = 1.0 / cell;
    P.levels = 0;
    int total = 0;
    for (int a = 0; a < 3; ++a) {
        P.lo[a] = lohi[a];
        double cells = (lohi[3 + a] - lohi[a]) / cell + 1.0;
        int b = 1;
        while (b < 21 && (double)(1u << b) < cells) ++b;
        P.bits[a] = b;
        P.levels = std::max(P.levels, b);
        total += b;
    }
    P.common = std::min(P.bits[0], std::min(P.bits[1], P.bits[2]));
    {
        int pos = 0;
        for (int l = 0; l < P.levels; ++l)
            for (int a = 0; a < 3; ++a)
                if (l < P.bits[a]) P.pos[a][l] = (unsigned char)pos++;
    }
    Scratch keys, keys_tmp, vals_tmp;
    NBR_TRY(keys.alloc(sizeof(uint64_t) * n, stream));
    NBR_TRY(keys_tmp.alloc(sizeof(uint64_t) * n, stream));
    NBR_TRY(vals_tmp.alloc(sizeof(uint32_t) * n, stream));
    morton_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, stream>>>(xyz, dtype, n, P, keys.as<uint64_t>(), perm_out);
    NBR_LAUNCHED();
    NBR_TRY(sort_pairs(keys.as<uint64_t>(), keys_tmp.as<uint64_t>(), perm_out, vals_tmp.as<uint32_t>(), n, 0, total, stream));
    if (sorted_xyz_out) {
        gather_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, stream>>>(xyz, dtype, n, perm_out, sorted_xyz_out);
        NBR_LAUNCHED();
    }
    return NBR_OK;
}

}  // namespace nbr
