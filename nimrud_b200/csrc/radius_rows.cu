// radius_rows.cu -- the row-interval fused radius query + covariance + eigensolve + feature kernel.
// since rows3.cu / rows5.cu (shell tables) took over every scale with r/e < 5.5 this kernel serves
// 5.5 <= r/e < 9.5, and everything below when the table kernels are switched off (NBR_NO_ROWS3 / NBR_NO_ROWS5).
//
// replaces nimrud/minimal/multiscale.py:94-122 (chunk kd-tree, query_ball_tree, take, population,
// centroid, pca) for EVERY such scale of a call in one launch.
//
// the search set of a scale is a voxel LATTICE, so a ball is a stack of x-intervals, one per (y,z) row.
//   * one warp owns 32 consecutive queries of a spatially coherent (Morton) order and walks through
//     the lattices (one per distinct voxel edge) in turn.  for each lattice it stages the occupancy
//     bricks covering the warp's bounding box + halo into shared memory with 16-byte cp.async copies.
//     if the box does not fit the staging buffer (sparse points, or the Morton curve jumped inside the
//     warp) the lanes read their rows straight from global memory, skipping empty bricks through the
//     directory.
//   * each lane then walks the (2W+1)^2 rows of its own window.  per row the x-interval is found in
//     float32 from sqrt(rho^2 - dy^2 - dz^2); an endpoint closer than a rounding bound to a cell
//     boundary sends that row to the exact float64 test of the reference
//     ((dx*dx + dy*dy) + dz*dz <= r*r, no fma), so neighbor SETS are bit-exact by construction.
//   * occupancy & interval mask -> count / sum x / sum x^2 through a 256-entry byte table; y and z
//     are row constants.  all moments are exact integers; finalize.cuh turns them into features in
//     the same kernel.  no neighbor list ever exists in memory.
#include "common.cuh"
#include "finalize.cuh"
#include "lattice.cuh"
#include "radius_rows.cuh"

#include <stdlib.h>

namespace nbr {

// byte -> count | sum(pos) << 8 | sum(pos^2) << 16     (positions 0..7)
__device__ __forceinline__ uint32_t byte_moments(uint32_t b)
{
    uint32_t cnt = 0, s1 = 0, s2 = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i)
        if (b & (1u << i)) { cnt += 1; s1 += i; s2 += i * i; }
    return cnt | (s1 << 8) | (s2 << 16);
}

// 7-bit row -> count | sum(pos) << 10 | sum(pos^2) << 20: fields wide enough to add up a whole z-slab
// of a W <= 3 window (49 cells) and its jy-weighted sums without unpacking
__device__ __forceinline__ uint32_t row7_moments(uint32_t b)
{
    uint32_t cnt = 0, s1 = 0, s2 = 0;
#pragma unroll
    for (int i = 0; i < 7; ++i)
        if (b & (1u << i)) { cnt += 1; s1 += i; s2 += i * i; }
    return cnt | (s1 << 10) | (s2 << 20);
}

struct Acc {
    int n, sx, sy, sz, sxx, sxy, sxz, syy, syz, szz;
};

// exact membership mask of one row: bit t <-> cell kx = cx - W + t
__device__ __noinline__ uint32_t exact_row_mask(const GridDev &g, double qx, double qy, double qz, int cx, int ky,
                                                int kz, int W, double radius)
{
    const double r2 = __dmul_rn(radius, radius);
    const double dy2 = sqdiff(qy, grid_centre(g, ky, 1));
    const double dz2 = sqdiff(qz, grid_centre(g, kz, 2));
    uint32_t m = 0;
    for (int t = 0; t <= 2 * W; ++t) {
        double s = sqdiff(qx, grid_centre(g, (long long)cx - W + t, 0));
        s = __dadd_rn(s, dy2);
        s = __dadd_rn(s, dz2);
        if (s <= r2) m |= 1u << t;
    }
    return m;
}

// everything a lane needs to know about one (query, lattice)
struct LaneCtx {
    double q[3];
    int c[3];
    float fxm, fym, fzm;      // query position in window cell units (cell j' has its centre at j')
    int W;
};

// x-interval of one row in float32, exact float64 fallback near cell boundaries; then the byte table.
// returns count | sum x | sum x^2 of (bits & ball) in cnt, sx, sxx; x in window units 0..2W
__device__ __forceinline__ void row_moments(const GridDev &g, const RowsParam &P, int ri, const LaneCtx &X, int jy,
                                            int jz, float T, uint32_t bits, const uint32_t *s_lut, int &cnt, int &sx,
                                            int &sxx)
{
    const int W = X.W;
    const float Tc = fmaxf(T, 0.0f);
    const float rs = fminf(rsqrtf(Tc), 1.0e3f);
    const float s = Tc * rs;
    const float a = X.fxm - s, b = X.fxm + s;
    const float ca = ceilf(a), fb = floorf(b);
    const float delta = P.eps_a[ri] * rs + P.eps_b;
    const bool unsure = (ca - a < delta) | (a - (ca - 1.0f) < delta) | (b - fb < delta) | (fb + 1.0f - b < delta) |
                        (T < -P.t_min[ri]);
    uint32_t m;
    if (unsure) {
        m = exact_row_mask(g, X.q[0], X.q[1], X.q[2], X.c[0], X.c[1] - W + jy, X.c[2] - W + jz, W, P.r[ri]);
    } else {
        const int il = max((int)ca, 0), ih = min((int)fb, 2 * W);
        m = ih >= il ? (((2u << ih) - 1u) & ~((1u << il) - 1u)) : 0u;
    }
    m &= bits;
    cnt = 0; sx = 0; sxx = 0;
    if (m == 0) return;
#pragma unroll
    for (int byte = 0; byte < 3; ++byte) {
        if (byte * 8 <= 2 * W) {
            const uint32_t e = s_lut[(m >> (8 * byte)) & 255u];
            const int bc = e & 255, b1 = (e >> 8) & 255, b2 = e >> 16;
            cnt += bc;
            sx += b1 + 8 * byte * bc;
            sxx += b2 + 16 * byte * b1 + 64 * byte * byte * bc;
        }
    }
}

// one lane, one radius: walk the (2Wr+1)^2 rows of the window.
// STAGED: rows come from the warp's shared-memory window (bricks [iz][iy][ix], origin lo[] in bricks).
// !STAGED: rows come straight from the directory + pool in global memory (incoherent warps); empty
//          bricks are skipped through the directory.
template <bool STAGED>
__device__ __noinline__ void lane_rows(const LatticeDev &L, const RowsParam &P, int ri, const LaneCtx &X,
                                          const uint32_t *win, const int lo[3], int nb0, int nb1,
                                          const uint32_t *s_lut, Acc &A)
{
    const GridDev &g = L.g;
    const int W = X.W, Wr = P.w[ri];
    const float rho2 = P.rho2[ri], t_min = P.t_min[ri];
    const uint32_t rowmask = (2u << (2 * W)) - 1u;
    // x alignment: window bit 0 <-> absolute cell c0 - W
    const int xa = X.c[0] - W;                                    // absolute cell of window bit 0
    const int bx0 = xa >> BRICK_XS;                               // floor
    const int sh = xa & 31;
    const bool two = sh + 2 * W + 1 > 32;
    const int ixs = STAGED ? bx0 - lo[0] : 0;
    const int ya = X.c[1] - W, za = X.c[2] - W;                   // absolute cells of rows jy' = 0, jz' = 0
    int cached_gy = 0x7fffffff, cached_gz = 0x7fffffff;
    uint32_t slot_a = 0, slot_b = 0;

    for (int jz = W - Wr; jz <= W + Wr; ++jz) {
        const float dz = X.fzm - (float)jz;
        const float Tz = rho2 - dz * dz;
        if (Tz < t_min) continue;
        const int az = za + jz;
        const int gz = az >> BRICK_ZS;
        const int wz = (az & (BRICK_Z - 1)) << BRICK_YS;
        if (!STAGED && (gz < 0 || gz >= L.nbz)) continue;
        const int zoff = STAGED ? ((gz - lo[2]) * nb1) * nb0 * BRICK_WORDS + wz : 0;
        int C = 0, SX = 0, SXX = 0, SY = 0, SYY = 0, SXY = 0;
        for (int jy = W - Wr; jy <= W + Wr; ++jy) {
            const float dy = X.fym - (float)jy;
            const float T = Tz - dy * dy;
            if (T < t_min) continue;
            const int ay = ya + jy;
            const int gy = ay >> BRICK_YS;
            uint32_t w0, w1;
            if (STAGED) {
                const int off = zoff + ((gy - lo[1]) * nb0 + ixs) * BRICK_WORDS + (ay & (BRICK_Y - 1));
                w0 = win[off];
                w1 = two ? win[off + BRICK_WORDS] : 0u;
            } else {
                if (gy < 0 || gy >= L.nby) continue;
                if (gy != cached_gy || gz != cached_gz) {
                    cached_gy = gy; cached_gz = gz;
                    const int64_t rowb = ((int64_t)gz * L.nby + gy) * L.nbx;
                    slot_a = (bx0 >= 0 && bx0 < L.nbx) ? L.dir[rowb + bx0] : 0u;
                    slot_b = (two && bx0 + 1 >= 0 && bx0 + 1 < L.nbx) ? L.dir[rowb + bx0 + 1] : 0u;
                }
                if ((slot_a | slot_b) == 0) continue;
                const int word = wz | (ay & (BRICK_Y - 1));
                w0 = slot_a ? L.pool[(int64_t)slot_a * BRICK_WORDS + word] : 0u;
                w1 = slot_b ? L.pool[(int64_t)slot_b * BRICK_WORDS + word] : 0u;
            }
            const uint32_t bits = __funnelshift_r(w0, w1, sh) & rowmask;
            if (bits == 0) continue;
            int cnt, sx, sxx;
            row_moments(g, P, ri, X, jy, jz, T, bits, s_lut, cnt, sx, sxx);
            C += cnt; SX += sx; SXX += sxx;
            SY += jy * cnt; SYY += jy * jy * cnt; SXY += jy * sx;
        }
        A.n += C; A.sx += SX; A.sxx += SXX; A.sy += SY; A.syy += SYY; A.sxy += SXY;
        A.sz += jz * C; A.szz += jz * jz * C; A.sxz += jz * SX; A.syz += jz * SY;
    }
}

// ---- W = 3 window (every r/e < 3.5) --------------------------------------------------------------
// per z-slab the 7 rows x 7 bits of the lane's window are gathered into one 49-bit word by a short
// unrolled block (all loads in flight together: shared memory when the warp's window is staged,
// directory + pool in global memory otherwise), then a ROLLED loop walks the non-empty rows.  the hot
// loop is ~60 instructions, so it lives in the L0 instruction cache; unrolling it made the kernel
// instruction-fetch bound (ncu: 31% of stall samples were no_instruction, profiles/).
//   staged: bricks [iz][iy][ix] at win, origin lo[] (bricks), nb0 x nb1 bricks per slab.
//   direct: the 2 x 2 x 3 bricks of the lane's own window are resolved through the directory first;
//           slabs whose bricks are all empty never touch the pool.
__device__ __forceinline__ void lane_rows_w3(const LatticeDev &L, const RowsParam &P, int ri, const LaneCtx &X,
                                             bool staged, const uint32_t *win, const int lo[3], int nb0, int nb1,
                                             const uint32_t *s_lut10, Acc &A)
{
    constexpr int W = 3, N = 7;
    const GridDev &g = L.g;
    const float rho2 = P.rho2[ri], t_min = P.t_min[ri], eps_a = P.eps_a[ri], eps_b = P.eps_b;
    const uint32_t rowmask = (1u << N) - 1u;
    const int xa = X.c[0] - W;
    const int sh = xa & 31;
    const bool two = sh + N > 32;
    const int bx0 = xa >> BRICK_XS;
    const int ya = X.c[1] - W, za = X.c[2] - W;
    const int by0 = ya >> BRICK_YS, bz0 = za >> BRICK_ZS;
    const int ycross = BRICK_Y - (ya & (BRICK_Y - 1));          // rows jy >= ycross live in the next y-brick
    int ybase = 0, ystep = 0, zstride = 0;
    uint32_t slot[3][2][2];
    if (staged) {
        ybase = ((by0 - lo[1]) * nb0 + (bx0 - lo[0])) * BRICK_WORDS + (ya & (BRICK_Y - 1));
        ystep = nb0 * BRICK_WORDS - BRICK_Y;                      // extra offset once the row crosses into the next brick
        zstride = nb1 * nb0 * BRICK_WORDS;
    } else {
#pragma unroll
        for (int iz = 0; iz < 3; ++iz)
#pragma unroll
            for (int iy = 0; iy < 2; ++iy)
#pragma unroll
                for (int ix = 0; ix < 2; ++ix) {
                    const int gx = bx0 + ix, gy = by0 + iy, gz = bz0 + iz;
                    const bool ok = (ix == 0 || two) && gx >= 0 && gx < L.nbx && gy >= 0 && gy < L.nby && gz >= 0 &&
                                    gz < L.nbz;
                    slot[iz][iy][ix] = ok ? L.dir[((int64_t)gz * L.nby + gy) * L.nbx + gx] : 0u;
                }
    }
    for (int jz = 0; jz < N; ++jz) {
        const float dz = X.fzm - (float)jz;
        const float Tz = rho2 - dz * dz;
        if (Tz < t_min) continue;
        const int az = za + jz;
        const int wz = (az & (BRICK_Z - 1)) << BRICK_YS;
        // ---- gather the slab: bits of row jy at [7*jy, 7*jy+7)
        unsigned long long slab = 0;
        if (staged) {
            const int zoff = ((az >> BRICK_ZS) - lo[2]) * zstride + wz + ybase;
#pragma unroll
            for (int jy = 0; jy < N; ++jy) {
                const int off = zoff + jy + (jy >= ycross ? ystep : 0);
                const uint32_t w0 = win[off];
                const uint32_t w1 = two ? win[off + BRICK_WORDS] : 0u;
                slab |= (unsigned long long)(__funnelshift_r(w0, w1, sh) & rowmask) << (N * jy);
            }
        } else {
            const int iz = (az >> BRICK_ZS) - bz0;                     // 0..2
            const uint32_t s00 = iz == 0 ? slot[0][0][0] : (iz == 1 ? slot[1][0][0] : slot[2][0][0]);
            const uint32_t s01 = iz == 0 ? slot[0][0][1] : (iz == 1 ? slot[1][0][1] : slot[2][0][1]);
            const uint32_t s10 = iz == 0 ? slot[0][1][0] : (iz == 1 ? slot[1][1][0] : slot[2][1][0]);
            const uint32_t s11 = iz == 0 ? slot[0][1][1] : (iz == 1 ? slot[1][1][1] : slot[2][1][1]);
            if ((s00 | s01 | s10 | s11) == 0) continue;
#pragma unroll
            for (int jy = 0; jy < N; ++jy) {
                const bool up = jy >= ycross;
                const uint32_t sa = up ? s10 : s00, sb = up ? s11 : s01;
                const int word = wz | ((ya + jy) & (BRICK_Y - 1));
                const uint32_t w0 = sa ? L.pool[(int64_t)sa * BRICK_WORDS + word] : 0u;
                const uint32_t w1 = sb ? L.pool[(int64_t)sb * BRICK_WORDS + word] : 0u;
                slab |= (unsigned long long)(__funnelshift_r(w0, w1, sh) & rowmask) << (N * jy);
            }
        }
        if (slab == 0) continue;
        // ---- rolled walk over the rows
        uint32_t Pk = 0, Qk = 0;
        int R = 0;
#pragma unroll 1
        for (int jy = 0; jy < N; ++jy) {
            const uint32_t bits = (uint32_t)(slab >> (N * jy)) & rowmask;
            const float dy = X.fym - (float)jy;
            const float T = Tz - dy * dy;
            if (bits == 0 || T < t_min) continue;
            const float Tc = fmaxf(T, 0.0f);
            const float rs = fminf(rsqrtf(Tc), 1.0e3f);
            const float s = Tc * rs;
            const float a = X.fxm - s, b = X.fxm + s;
            const float ca = ceilf(a), fb = floorf(b);
            const float delta = eps_a * rs + eps_b;
            const float da = ca - a, db = b - fb;                      // both in [0, 1)
            const bool unsure = (fminf(da, db) < delta) | (fmaxf(da, db) > 1.0f - delta) | (T < -t_min);
            uint32_t m;
            if (unsure) {
                m = exact_row_mask(g, X.q[0], X.q[1], X.q[2], X.c[0], ya + jy, az, W, P.r[ri]);
            } else {
                const int il = max((int)ca, 0), ih = (int)fb;          // ih in [0, 2W+1] by construction
                m = ((2u << ih) - 1u) & ~((1u << il) - 1u);
            }
            const uint32_t e = s_lut10[m & bits];
            Pk += e;
            Qk += jy * e;
            R += jy * jy * (int)(e & 1023u);
        }
        const int C = Pk & 1023, SX = (Pk >> 10) & 1023, SXX = Pk >> 20;
        const int SY = Qk & 1023, SXY = (Qk >> 10) & 1023;
        A.n += C; A.sx += SX; A.sxx += SXX; A.sy += SY; A.syy += R; A.sxy += SXY;
        A.sz += jz * C; A.szz += jz * jz * C; A.sxz += jz * SX; A.syz += jz * SY;
    }
}

__device__ __forceinline__ void cp_async16(uint32_t smem_addr, const void *gptr, bool valid)
{
    const int src_bytes = valid ? 16 : 0;      // 0 -> the 16 bytes are zero-filled
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_addr), "l"(gptr), "r"(src_bytes));
}

template <typename OutT>
__global__ void __launch_bounds__(RW_WARPS * 32, 4)
radius_rows_kernel(const RowsLaunch *__restrict__ launch, const void *__restrict__ query, int dtype,
                   const uint32_t *__restrict__ perm, int64_t nq, OutT *__restrict__ out, int64_t row_stride,
                   int descriptor_mask)
{
    __shared__ uint32_t s_lut[256];
    __shared__ uint32_t s_lut10[128];
    __shared__ __align__(16) uint32_t s_win[RW_WARPS][RW_CAP_BRICKS * BRICK_WORDS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 256; i += blockDim.x) s_lut[i] = byte_moments(i);
    for (int i = threadIdx.x; i < 128; i += blockDim.x) s_lut10[i] = row7_moments(i);
    __syncthreads();

    uint32_t *win = s_win[warp];
    const uint32_t win_addr = (uint32_t)__cvta_generic_to_shared(win);
    const int64_t n_groups = (nq + 31) >> 5;
    const int n_lat = launch->n_lat;

    for (int64_t grp = (int64_t)blockIdx.x * RW_WARPS + warp; grp < n_groups; grp += (int64_t)gridDim.x * RW_WARPS) {
        const int64_t slot_i = grp * 32 + lane;
        const bool active = slot_i < nq;
        const int64_t src = active ? slot_i : grp * 32;             // inactive lanes shadow lane 0
        const int64_t qi = perm ? (int64_t)perm[src] : src;         // row of the output
        LaneCtx X;
#pragma unroll
        for (int a = 0; a < 3; ++a) X.q[a] = load_coord(query, dtype, src, 3, a);
        OutT *dst_row = out + qi * row_stride;

        for (int li = 0; li < n_lat; ++li) {
            const LatticeDev &L = launch->lat[li];
            const RowsParam &P = launch->rows[li];
            const GridDev &g = L.g;
            double f[3];
#pragma unroll
            for (int a = 0; a < 3; ++a) query_anchor(X.q[a], g, a, X.c[a], f[a]);
            // every window of half-width <= 3 runs through the one W = 3 code path (rows beyond the ball just
            // fail the disc test)
            const int W = P.wmax <= 3 ? 3 : P.wmax;
            X.W = W;
            X.fxm = (float)f[0] - 0.5f + (float)W;
            X.fym = (float)f[1] - 0.5f + (float)W;
            X.fzm = (float)f[2] - 0.5f + (float)W;

            // ---- brick window of the whole warp
            int lo[3], nb[3];
            long long vol = 1;
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                const int mn = __reduce_min_sync(0xffffffffu, X.c[a]) - W;
                const int mx = __reduce_max_sync(0xffffffffu, X.c[a]) + W;
                const int sh = a == 0 ? BRICK_XS : (a == 1 ? BRICK_YS : BRICK_ZS);
                lo[a] = mn >> sh;                       // arithmetic shift = floor division
                const long long cnt = (long long)(mx >> sh) - lo[a] + 1;
                nb[a] = (int)(cnt < 1000 ? cnt : 1000);
                vol *= nb[a];
            }
            const bool staged = vol <= RW_CAP_BRICKS;
            if (launch->stats && lane == 0) {
                unsigned long long *st = launch->stats + li * 8;
                atomicAdd(st + (staged ? 0 : 1), 1ull);
                if (staged) atomicAdd(st + 2, (unsigned long long)vol);
            }

            if (staged) {
                // brick (ix,iy,iz) -> win[((iz*nb1)+iy)*nb0+ix][32]; 16-byte async copies, 4 bricks per step
                const int total = nb[0] * nb[1] * nb[2];
                __syncwarp();
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
                const int sub = lane >> 3, chunk = lane & 7;
                for (int b0 = 0; b0 < total; b0 += 4) {
                    const int b = b0 + sub;
                    const uint32_t s = __shfl_sync(0xffffffffu, b < 32 ? slot0 : slot1, b & 31);
                    if (b < total)
                        cp_async16(win_addr + (uint32_t)(b * BRICK_WORDS + chunk * 4) * 4u,
                                   L.pool + (int64_t)s * BRICK_WORDS + chunk * 4, s != 0);
                }
                asm volatile("cp.async.wait_all;" ::: "memory");
                __syncwarp();
            }

            for (int ri = 0; ri < P.n; ++ri) {
                Acc A = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
                if (W == 3)      lane_rows_w3(L, P, ri, X, staged, win, lo, nb[0], nb[1], s_lut10, A);
                else if (staged) lane_rows<true>(L, P, ri, X, win, lo, nb[0], nb[1], s_lut, A);
                else             lane_rows<false>(L, P, ri, X, win, lo, nb[0], nb[1], s_lut, A);
                if (active)
                    emit_features_window<OutT>(A.n, A.sx, A.sy, A.sz, A.sxx, A.sxy, A.sxz, A.syy, A.syz, A.szz, X.fxm,
                                               X.fym, X.fzm, W <= 6, g.edge, dst_row + P.col[ri], descriptor_mask);
            }
            __syncwarp();
        }
    }
}

int rows_param(const Lattice *lat, const double *radii, const int *cols, int nr, RowsParam *P, cudaStream_t)
{
    const double e = lat->grid.edge;
    if (nr > RW_MAX_RADII) return fail(NBR_ERR_INVALID, "rows_param: too many radii in one group");
    P->n = nr;
    P->wmax = 0;
    for (int k = 0; k < nr; ++k) {
        const double rho = radii[k] / e;
        P->r[k] = radii[k];
        P->rho2[k] = (float)(rho * rho);
        P->w[k] = (int)floor(rho + 0.5 + 1e-6);
        P->col[k] = cols[k];
        P->wmax = std::max(P->wmax, P->w[k]);
    }
    const int weff = std::max(P->wmax, 3);      // the kernel runs every window <= 3 as a W = 3 window
    for (int k = 0; k < nr; ++k) {
        const double mag = (double)P->rho2[k] + (weff + 1.0) * (weff + 1.0);
        P->eps_a[k] = (float)(8.0 * 5.96e-8 * mag);
        P->t_min[k] = (float)(-16.0 * 5.96e-8 * mag);
    }
    P->eps_b = (float)(4.77e-7 * (weff + 1.0));
    return NBR_OK;
}

bool rows_supported(double edge, const double *radii, int nr)
{
    for (int k = 0; k < nr; ++k)
        if (!(radii[k] / edge + 0.5 + 1e-6 < RW_MAX_W + 1)) return false;
    return true;
}

// launch->lat / rows / n_lat filled by the caller (host copy); one launch covers every lattice in it
int radius_rows_launch(const RowsLaunch *launch_host, const void *query, int dtype, const uint32_t *perm, int64_t nq,
                       void *out, int out_dtype, int64_t row_stride, int descriptor_mask, cudaStream_t stream)
{
    if (nq <= 0 || launch_host->n_lat <= 0) return NBR_OK;
    Scratch dev, stats;
    NBR_TRY(dev.alloc(sizeof(RowsLaunch), stream));
    RowsLaunch copy = *launch_host;
    const bool want_stats = getenv("NBR_ROW_STATS") != nullptr;
    if (want_stats) {
        NBR_TRY(stats.alloc(sizeof(unsigned long long) * 8 * RW_MAX_LATTICES, stream));
        NBR_CUDA(cudaMemsetAsync(stats.ptr, 0, sizeof(unsigned long long) * 8 * RW_MAX_LATTICES, stream));
        copy.stats = stats.as<unsigned long long>();
    } else {
        copy.stats = nullptr;
    }
    NBR_CUDA(cudaMemcpyAsync(dev.ptr, &copy, sizeof(RowsLaunch), cudaMemcpyHostToDevice, stream));
    const int blocks = (int)std::min<int64_t>(ceil_div(ceil_div(nq, 32), RW_WARPS), (int64_t)device_sm_count() * 16);
    if (out_dtype == NBR_F32)
        radius_rows_kernel<float><<<blocks, RW_WARPS * 32, 0, stream>>>(dev.as<RowsLaunch>(), query, dtype, perm, nq,
                                                                        (float *)out, row_stride, descriptor_mask);
    else
        radius_rows_kernel<double><<<blocks, RW_WARPS * 32, 0, stream>>>(dev.as<RowsLaunch>(), query, dtype, perm, nq,
                                                                         (double *)out, row_stride, descriptor_mask);
    NBR_LAUNCHED();
    if (want_stats) {
        unsigned long long h[8 * RW_MAX_LATTICES];
        NBR_CUDA(cudaMemcpyAsync(h, stats.ptr, sizeof(h), cudaMemcpyDeviceToHost, stream));
        NBR_CUDA(cudaStreamSynchronize(stream));
        for (int l = 0; l < launch_host->n_lat; ++l)
            fprintf(stderr, "[nbr row stats] lattice %d edge %.3g W %d: staged warps %llu (avg %.1f bricks), direct warps %llu\n",
                    l, launch_host->lat[l].g.edge, launch_host->rows[l].wmax, h[l * 8], h[l * 8] ? (double)h[l * 8 + 2] / h[l * 8] : 0.0,
                    h[l * 8 + 1]);
    }
    return NBR_OK;
}

}  // namespace nbr
