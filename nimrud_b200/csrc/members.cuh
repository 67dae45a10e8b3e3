// members.cuh -- per-candidate visitor of the voxels within a radius of a query, every membership test in the
// reference's exact float64 arithmetic (shared by radius_exact.cu and vector_field.cu).
#pragma once
#include "common.cuh"

namespace nbr {

// visit every occupied cell within `radius` of q.  fn(jx, jy, jz, slot, word, bit_in_word)
// cells are produced in ascending address order (z, then y, then x).
template <typename F>
__device__ __forceinline__ void for_each_member(const LatticeDev &L, const double q[3], const int c[3], double radius,
                                                F fn)
{
    const GridDev &g = L.g;
    const double r2 = __dmul_rn(radius, radius);
    const int W = (int)fmin(ceil(radius / g.edge) + 1.0, 5.0e8);
    int lo[3], hi[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        lo[a] = c[a] - W < 0 ? 0 : c[a] - W;
        hi[a] = (long long)c[a] + W > g.ncell[a] - 1 ? g.ncell[a] - 1 : c[a] + W;
        if (c[a] < -W - 1 || lo[a] > hi[a]) return;
    }
    for (int kz = lo[2]; kz <= hi[2]; ++kz) {
        const double dz2 = sqdiff(q[2], grid_centre(g, kz, 2));
        if (dz2 > r2) continue;
        for (int ky = lo[1]; ky <= hi[1]; ++ky) {
            const double dy2 = sqdiff(q[1], grid_centre(g, ky, 1));
            if (dy2 > r2) continue;
            const int word = ((kz & (BRICK_Z - 1)) << BRICK_YS) | (ky & (BRICK_Y - 1));
            const int64_t rowb = ((int64_t)(kz >> BRICK_ZS) * L.nby + (ky >> BRICK_YS)) * L.nbx;
            for (int bx = lo[0] >> BRICK_XS; bx <= hi[0] >> BRICK_XS; ++bx) {
                const uint32_t slot = L.dir[rowb + bx];
                if (!slot) continue;
                uint32_t w = L.pool[(int64_t)slot * BRICK_WORDS + word];
                const int x0 = bx << BRICK_XS;
                if (lo[0] > x0) w &= ~0u << (lo[0] - x0);
                if (hi[0] < x0 + 31) w &= ~0u >> (x0 + 31 - hi[0]);
                while (w) {
                    const int b = __ffs(w) - 1;
                    w &= w - 1;
                    const int kx = x0 + b;
                    // ((dx*dx + dy*dy) + dz*dz) <= r*r, float64, inclusive
                    double s = sqdiff(q[0], grid_centre(g, kx, 0));
                    s = __dadd_rn(s, dy2);
                    s = __dadd_rn(s, dz2);
                    if (s <= r2) fn(kx - c[0], ky - c[1], kz - c[2], slot, word, b);
                }
            }
        }
    }
}

}  // namespace nbr
