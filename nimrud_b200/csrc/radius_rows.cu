// radius_rows.cu -- the fused radius query + covariance + eigensolve + feature kernel (hot path).
//
// replaces nimrud/minimal/multiscale.py:94-122 (chunk kd-tree, query_ball_tree, take, population,
// centroid, pca) for one lattice (one voxel edge) and all radii that share it.
//
// the search set is a voxel LATTICE, so a ball is a stack of x-intervals, one per (y,z) row.
//   * one warp owns 32 consecutive queries (callers pass a spatially coherent order).  it stages the
//     occupancy bricks covering the warp's bounding box + halo into shared memory once, coalesced.
//   * each lane then walks the (2W+1)^2 rows of its own window.  per row the x-interval is found in
//     float32 from sqrt(rho^2 - dy^2 - dz^2); an endpoint closer than a rounding bound to a cell
//     boundary sends that row to the exact float64 test of the reference
//     ((dx*dx + dy*dy) + dz*dz <= r*r, no fma), so neighbor SETS are bit-exact by construction.
//   * occupancy & interval mask -> count / sum x / sum x^2 through a 256-entry byte table; y and z
//     are row constants.  all moments are exact integers; finalize.cuh turns them into features in
//     the same kernel.  no neighbor list ever exists in memory.
// windows too large for the staging buffer (outlier queries far from the cloud) fall back to the
// exact per-candidate kernel's traversal for that warp.
#include "common.cuh"
#include "finalize.cuh"
#include "lattice.cuh"

namespace nbr {

constexpr int RW_WARPS = 4;
constexpr int RW_CAP_BRICKS = 48;                 // staged bricks per warp (6 KB)
constexpr int RW_MAX_W = 15;                      // 2W+1 <= 31 bits per row
constexpr int RW_MAX_RADII = 8;

struct RowsParam {
    double r[RW_MAX_RADII];      // exact radii (fallback test)
    float rho2[RW_MAX_RADII];    // (r/e)^2
    float eps_a[RW_MAX_RADII];   // rounding bound: delta = eps_a * min(rsqrt(T), 1e3) + eps_b
    float t_min[RW_MAX_RADII];   // rows with T < t_min are certainly empty
    int w[RW_MAX_RADII];         // window half-width of each radius
    float eps_b;
    int n;
    int wmax;
};

// byte -> count | sum(pos) << 8 | sum(pos^2) << 16     (positions 0..7)
__device__ __forceinline__ uint32_t byte_moments(uint32_t b)
{
    uint32_t cnt = 0, s1 = 0, s2 = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i)
        if (b & (1u << i)) { cnt += 1; s1 += i; s2 += i * i; }
    return cnt | (s1 << 8) | (s2 << 16);
}

struct Acc {
    int n, sx, sy, sz, sxx, sxy, sxz, syy, syz, szz;
};

// exact membership mask of one row: bit t <-> cell kx = cx - W + t
__device__ __noinline__ uint32_t exact_row_mask(const GridDev &g, double qx, double qy, double qz, int cx, int ky,
                                                int kz, int W, double radius)
{
    const double r2 = __dmul_rn(radius, radius);
    const double dy2 = sqdiff(qy, cell_centre(ky, g.minc[1], g.edge));
    const double dz2 = sqdiff(qz, cell_centre(kz, g.minc[2], g.edge));
    uint32_t m = 0;
    for (int t = 0; t <= 2 * W; ++t) {
        double s = sqdiff(qx, cell_centre((long long)cx - W + t, g.minc[0], g.edge));
        s = __dadd_rn(s, dy2);
        s = __dadd_rn(s, dz2);
        if (s <= r2) m |= 1u << t;
    }
    return m;
}

template <typename OutT>
__global__ void __launch_bounds__(RW_WARPS * 32)
radius_rows_kernel(LatticeDev L, const void *__restrict__ query, int dtype, const uint32_t *__restrict__ perm,
                   int64_t nq, RowsParam P, OutT *__restrict__ out, int64_t row_stride, int col_offset,
                   int descriptor_mask)
{
    __shared__ uint32_t s_lut[256];
    __shared__ uint32_t s_win[RW_WARPS][RW_CAP_BRICKS * BRICK_WORDS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 256; i += blockDim.x) s_lut[i] = byte_moments(i);
    __syncthreads();

    const GridDev &g = L.g;
    const int ncol = (descriptor_mask & NBR_DESC_EXTENDED) ? NBR_COLS_EXTENDED : NBR_COLS_REFERENCE;
    uint32_t *win = s_win[warp];
    const int64_t n_groups = (nq + 31) >> 5;

    for (int64_t grp = (int64_t)blockIdx.x * RW_WARPS + warp; grp < n_groups; grp += (int64_t)gridDim.x * RW_WARPS) {
        const int64_t slot_i = grp * 32 + lane;
        const bool active = slot_i < nq;
        const int64_t qi = perm ? (int64_t)perm[active ? slot_i : grp * 32] : (active ? slot_i : grp * 32);
        double q[3], f[3];
        int c[3];
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            q[a] = load_coord(query, dtype, qi, 3, a);
            query_anchor(q[a], g.minc[a], g.edge, c[a], f[a]);
        }
        // warp bounding box of the anchor cells -> brick window
        const int W = P.wmax;
        int lo[3], nb[3];
        bool fits = true;
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const int mn = __reduce_min_sync(0xffffffffu, c[a]) - W;
            const int mx = __reduce_max_sync(0xffffffffu, c[a]) + W;
            const int sh = a == 0 ? BRICK_XS : (a == 1 ? BRICK_YS : BRICK_ZS);
            lo[a] = mn >> sh;                       // arithmetic shift = floor division
            const long long cnt = (long long)(mx >> sh) - lo[a] + 1;
            fits &= cnt <= RW_CAP_BRICKS;
            nb[a] = (int)cnt;
        }
        fits = fits && (long long)nb[0] * nb[1] * nb[2] <= RW_CAP_BRICKS;

        OutT *dst = out + qi * row_stride + col_offset;

        if (!fits) {
            // outlier warp: exact per-candidate traversal straight from global memory
            for (int ri = 0; ri < P.n; ++ri) {
                Moments m;
                m.n = 0;
#pragma unroll
                for (int k = 0; k < 3; ++k) m.s1[k] = 0;
#pragma unroll
                for (int k = 0; k < 6; ++k) m.s2[k] = 0;
                const double radius = P.r[ri];
                const double r2 = __dmul_rn(radius, radius);
                const int Wr = P.w[ri];
                int l3[3], h3[3];
                bool none = false;
#pragma unroll
                for (int a = 0; a < 3; ++a) {
                    l3[a] = c[a] - Wr < 0 ? 0 : c[a] - Wr;
                    h3[a] = (long long)c[a] + Wr > g.ncell[a] - 1 ? g.ncell[a] - 1 : c[a] + Wr;
                    none |= (c[a] < -Wr - 1) | (l3[a] > h3[a]);
                }
                if (!none)
                    for (int kz = l3[2]; kz <= h3[2]; ++kz) {
                        const double dz2 = sqdiff(q[2], cell_centre(kz, g.minc[2], g.edge));
                        if (dz2 > r2) continue;
                        for (int ky = l3[1]; ky <= h3[1]; ++ky) {
                            const double dy2 = sqdiff(q[1], cell_centre(ky, g.minc[1], g.edge));
                            if (dy2 > r2) continue;
                            const int word = ((kz & (BRICK_Z - 1)) << BRICK_YS) | (ky & (BRICK_Y - 1));
                            const int64_t rowb = ((int64_t)(kz >> BRICK_ZS) * L.nby + (ky >> BRICK_YS)) * L.nbx;
                            for (int bx = l3[0] >> BRICK_XS; bx <= h3[0] >> BRICK_XS; ++bx) {
                                const uint32_t slot = L.dir[rowb + bx];
                                if (!slot) continue;
                                uint32_t w = L.pool[(int64_t)slot * BRICK_WORDS + word];
                                const int x0 = bx << BRICK_XS;
                                if (l3[0] > x0) w &= ~0u << (l3[0] - x0);
                                if (h3[0] < x0 + 31) w &= ~0u >> (x0 + 31 - h3[0]);
                                while (w) {
                                    const int b = __ffs(w) - 1;
                                    w &= w - 1;
                                    double s = sqdiff(q[0], cell_centre(x0 + b, g.minc[0], g.edge));
                                    s = __dadd_rn(s, dy2);
                                    s = __dadd_rn(s, dz2);
                                    if (s <= r2) {
                                        const long long jx = x0 + b - c[0], jy = ky - c[1], jz = kz - c[2];
                                        m.n += 1;
                                        m.s1[0] += jx; m.s1[1] += jy; m.s1[2] += jz;
                                        m.s2[0] += jx * jx; m.s2[1] += jx * jy; m.s2[2] += jx * jz;
                                        m.s2[3] += jy * jy; m.s2[4] += jy * jz; m.s2[5] += jz * jz;
                                    }
                                }
                            }
                        }
                    }
                if (active) emit_features<OutT>(m, f, g.edge, dst + ri * ncol, descriptor_mask);
            }
            continue;
        }

        // ---- stage the window: brick (ix,iy,iz) of the window -> win[((iz*nb1)+iy)*nb0+ix][32]
        const int total = nb[0] * nb[1] * nb[2];
        __syncwarp();
        {
            // each lane resolves the slot of bricks lane, lane+32; then the warp copies brick by brick
            uint32_t slot0 = 0, slot1 = 0;
#pragma unroll
            for (int t = 0; t < 2; ++t) {
                const int b = lane + 32 * t;
                uint32_t s = 0;
                if (b < total) {
                    const int ix = b % nb[0], iy = (b / nb[0]) % nb[1], iz = b / (nb[0] * nb[1]);
                    const int gx = lo[0] + ix, gy = lo[1] + iy, gz = lo[2] + iz;
                    if (gx >= 0 && gx < L.nbx && gy >= 0 && gy < L.nby && gz >= 0 && gz < L.nbz)
                        s = L.dir[((int64_t)gz * L.nby + gy) * L.nbx + gx];
                }
                if (t == 0) slot0 = s; else slot1 = s;
            }
            for (int b = 0; b < total; ++b) {
                const uint32_t s = __shfl_sync(0xffffffffu, b < 32 ? slot0 : slot1, b & 31);
                win[b * BRICK_WORDS + lane] = s ? L.pool[(int64_t)s * BRICK_WORDS + lane] : 0u;
            }
        }
        __syncwarp();

        // ---- per-lane row walk
        const int x0 = c[0] - W - lo[0] * BRICK_X;                 // >= 0
        const int ix = x0 >> 5, sh = x0 & 31;
        const bool two = sh + 2 * W + 1 > 32;
        const int y0 = c[1] - W - lo[1] * BRICK_Y;             // window-local cell of row jy' = 0
        const int z0 = c[2] - W - lo[2] * BRICK_Z;
        const float fxm = (float)f[0] - 0.5f + (float)W;         // cell j' has its centre at x = j'
        const float fym = (float)f[1] - 0.5f + (float)W;
        const float fzm = (float)f[2] - 0.5f + (float)W;
        const uint32_t rowmask = (2u << (2 * W)) - 1u;

        for (int ri = 0; ri < P.n; ++ri) {
            const float rho2 = P.rho2[ri], eps_a = P.eps_a[ri], t_min = P.t_min[ri];
            const int Wr = P.w[ri];
            Acc A = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
            for (int jz = W - Wr; jz <= W + Wr; ++jz) {
                const float dz = fzm - (float)jz;
                const float Tz = rho2 - dz * dz;
                if (Tz < t_min) continue;
                const int lz = z0 + jz;
                const int zoff = ((lz >> BRICK_ZS) * nb[1]) * nb[0] * BRICK_WORDS + ((lz & (BRICK_Z - 1)) << BRICK_YS);
                int C = 0, SX = 0, SXX = 0, SY = 0, SYY = 0, SXY = 0;
                for (int jy = W - Wr; jy <= W + Wr; ++jy) {
                    const float dy = fym - (float)jy;
                    const float T = Tz - dy * dy;
                    if (T < t_min) continue;
                    const int ly = y0 + jy;
                    const int off = zoff + ((ly >> BRICK_YS) * nb[0] + ix) * BRICK_WORDS + (ly & (BRICK_Y - 1));
                    const uint32_t w0 = win[off];
                    const uint32_t w1 = two ? win[off + BRICK_WORDS] : 0u;
                    const uint32_t bits = __funnelshift_r(w0, w1, sh) & rowmask;
                    if (bits == 0) continue;
                    // float32 interval [fxm - s, fxm + s]
                    const float Tc = fmaxf(T, 0.0f);
                    const float rs = fminf(rsqrtf(Tc), 1.0e3f);
                    const float s = Tc * rs;
                    const float a = fxm - s, b = fxm + s;
                    const float ca = ceilf(a), fb = floorf(b);
                    const float delta = eps_a * rs + P.eps_b;
                    const bool unsure = (ca - a < delta) | (a - (ca - 1.0f) < delta) | (b - fb < delta) |
                                        (fb + 1.0f - b < delta) | (T < -t_min);
                    uint32_t m;
                    if (unsure) {
                        m = exact_row_mask(g, q[0], q[1], q[2], c[0], c[1] - W + jy, c[2] - W + jz, W, P.r[ri]);
                    } else {
                        const int il = max((int)ca, 0), ih = min((int)fb, 2 * W);
                        m = ih >= il ? (((2u << ih) - 1u) & ~((1u << il) - 1u)) : 0u;
                    }
                    m &= bits;
                    if (m == 0) continue;
                    // moments of the row through the byte table
                    int cnt = 0, sx = 0, sxx = 0;
#pragma unroll
                    for (int byte = 0; byte < 4; ++byte) {
                        if (byte * 8 <= 2 * RW_MAX_W && byte * 8 <= 2 * W) {
                            const uint32_t e = s_lut[(m >> (8 * byte)) & 255u];
                            const int bc = e & 255, b1 = (e >> 8) & 255, b2 = e >> 16;
                            cnt += bc;
                            sx += b1 + 8 * byte * bc;
                            sxx += b2 + 16 * byte * b1 + 64 * byte * byte * bc;
                        }
                    }
                    C += cnt; SX += sx; SXX += sxx;
                    SY += jy * cnt; SYY += jy * jy * cnt; SXY += jy * sx;
                }
                A.n += C; A.sx += SX; A.sxx += SXX; A.sy += SY; A.syy += SYY; A.sxy += SXY;
                A.sz += jz * C; A.szz += jz * jz * C; A.sxz += jz * SX; A.syz += jz * SY;
            }
            if (active) {
                // shift the unsigned window coordinates j' = j + W back to offsets from the anchor cell
                Moments m;
                const long long n = A.n, w = W;
                m.n = n;
                m.s1[0] = A.sx - w * n; m.s1[1] = A.sy - w * n; m.s1[2] = A.sz - w * n;
                m.s2[0] = A.sxx - 2 * w * A.sx + w * w * n;
                m.s2[3] = A.syy - 2 * w * A.sy + w * w * n;
                m.s2[5] = A.szz - 2 * w * A.sz + w * w * n;
                m.s2[1] = A.sxy - w * A.sx - w * A.sy + w * w * n;
                m.s2[2] = A.sxz - w * A.sx - w * A.sz + w * w * n;
                m.s2[4] = A.syz - w * A.sy - w * A.sz + w * w * n;
                emit_features<OutT>(m, f, g.edge, dst + ri * ncol, descriptor_mask);
            }
        }
        __syncwarp();
    }
}

int radius_features_rows(const Lattice *lat, const void *query, int dtype, const uint32_t *perm, int64_t nq,
                         const double *radii, int nr, void *out, int out_dtype, int64_t row_stride, int col_offset,
                         int descriptor_mask, cudaStream_t stream, bool *handled)
{
    *handled = false;
    if (nq <= 0 || nr <= 0) { *handled = true; return NBR_OK; }
    const double e = lat->grid.edge;
    const int ncol = (descriptor_mask & NBR_DESC_EXTENDED) ? NBR_COLS_EXTENDED : NBR_COLS_REFERENCE;
    for (int k = 0; k < nr; ++k)
        if (!(radii[k] / e + 0.5 + 1e-6 < RW_MAX_W + 1)) return NBR_OK;       // window too wide: caller falls back
    const int blocks = (int)std::min<int64_t>(ceil_div(ceil_div(nq, 32), RW_WARPS), (int64_t)device_sm_count() * 16);
    for (int base = 0; base < nr; base += RW_MAX_RADII) {
        RowsParam P;
        P.n = std::min(RW_MAX_RADII, nr - base);
        P.wmax = 0;
        for (int k = 0; k < P.n; ++k) {
            const double rho = radii[base + k] / e;
            P.r[k] = radii[base + k];
            P.rho2[k] = (float)(rho * rho);
            P.w[k] = (int)floor(rho + 0.5 + 1e-6);
            P.wmax = std::max(P.wmax, P.w[k]);
        }
        for (int k = 0; k < P.n; ++k) {
            const double mag = (double)P.rho2[k] + (P.wmax + 1.0) * (P.wmax + 1.0);
            P.eps_a[k] = (float)(8.0 * 5.96e-8 * mag);
            P.t_min[k] = (float)(-16.0 * 5.96e-8 * mag);
        }
        P.eps_b = (float)(4.77e-7 * (P.wmax + 1.0));
        if (out_dtype == NBR_F32)
            radius_rows_kernel<float><<<blocks, RW_WARPS * 32, 0, stream>>>(lat->dev(), query, dtype, perm, nq, P,
                                                                            (float *)out, row_stride,
                                                                            col_offset + base * ncol, descriptor_mask);
        else
            radius_rows_kernel<double><<<blocks, RW_WARPS * 32, 0, stream>>>(lat->dev(), query, dtype, perm, nq, P,
                                                                             (double *)out, row_stride,
                                                                             col_offset + base * ncol, descriptor_mask);
        NBR_LAUNCHED();
    }
    *handled = true;
    return NBR_OK;
}

}  // namespace nbr
