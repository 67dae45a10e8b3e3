// rows5.cu -- the fused feature kernel for 11x11x11 windows (3.5 <= r/e + 0.5 < 6, e.g. the reference's
// own example: edges (0.1, 0.2, 0.4), radii (0.5, 1.0, 2.0)).
//
// same scheme as rows3.cu (nimrud/minimal/multiscale.py:94-122 for every such scale of a call in one launch),
// with rows instead of slabs as the unit: per (row, table bin) one 32-bit table word holds the cells that are
// inside the ball for every fractional position of the bin (low 16 bits) and the cells that may be on either
// side (high 16 bits).  occupancy & inside -> moments through a 2048-entry row table; occupied uncertain cells
// are parked and decided after the rows, float32 first, the reference's float64 expression inside the rounding
// band: neighbor sets stay bit-exact.
#include "common.cuh"
#include "finalize.cuh"
#include "lattice.cuh"
#include "radius_rows.cuh"

#include <map>
#include <mutex>
#include <tuple>

#include <stdlib.h>

namespace nbr {

constexpr int R5_WARPS = 4;
#ifndef R5_BLOCKS_N
#define R5_BLOCKS_N 6
#endif
#ifndef R5_CAP_N
#define R5_CAP_N 64
#endif
constexpr int R5_CAP = R5_CAP_N;           // staged bricks per warp (128 bytes each)
constexpr int N11 = 11, W5 = 5;
constexpr int R5_SLAB_WORDS = 12;          // 11 row words + 1 padding word = 3 x uint4
constexpr int R5_BIN_WORDS = N11 * R5_SLAB_WORDS;
constexpr int R5_WIN_BYTES = R5_CAP * BRICK_WORDS * 4;
constexpr int R5_ULIST_BYTES = 0;          // rows with occupied uncertain cells are only FLAGGED (121 bits per lane) and
                                           // gathered again when they are decided: parking their masks took 7.8 KB of
                                           // shared memory per warp and a store per row

// ---- shell table for the 11-wide window: table[bin][slab][12] uint32 = in | unc << 16 per row
__global__ void __launch_bounds__(256)
ball_table5_kernel(uint32_t *__restrict__ table, int Q, double rho2, double margin)
{
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= Q * Q * Q * N11) return;
    const int jz = idx % N11, bin = idx / N11;
    const int b[3] = {bin % Q, (bin / Q) % Q, bin / (Q * Q)};
    double lo[3], hi[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        lo[a] = (double)b[a] / Q - 0.5 + (double)W5;       // window units: cell t has its centre at t
        hi[a] = (double)(b[a] + 1) / Q - 0.5 + (double)W5;
    }
    auto axis = [&](int a, int t, double &dmin, double &dmax) {
        const double x = (double)t;
        dmin = fmax(0.0, fmax(lo[a] - x, x - hi[a]));
        dmax = fmax(fabs(lo[a] - x), fabs(hi[a] - x));
    };
    double zmin, zmax;
    axis(2, jz, zmin, zmax);
    uint32_t *dst = table + (size_t)bin * R5_BIN_WORDS + jz * R5_SLAB_WORDS;
    for (int jy = 0; jy < N11; ++jy) {
        double ymin, ymax;
        axis(1, jy, ymin, ymax);
        uint32_t in = 0, unc = 0;
        for (int t = 0; t < N11; ++t) {
            double xmin, xmax;
            axis(0, t, xmin, xmax);
            const double d2min = xmin * xmin + ymin * ymin + zmin * zmin;
            const double d2max = xmax * xmax + ymax * ymax + zmax * zmax;
            if (d2max <= rho2 - margin) in |= 1u << t;
            else if (!(d2min > rho2 + margin)) unc |= 1u << t;
        }
        dst[jy] = in | (unc << 16);
    }
    dst[N11] = 0;
}

static std::mutex g_table5_mutex;
static std::map<std::tuple<int, int, uint64_t, uint64_t>, const uint32_t *> g_tables5;

static int ball_table5_get(double rho2, double margin, int Q, const uint32_t **out, cudaStream_t stream)
{
    int dev = 0;
    NBR_CUDA(cudaGetDevice(&dev));
    int ex = 0;
    frexp(margin, &ex);
    margin = ldexp(1.0, ex);
    uint64_t kr, km;
    memcpy(&kr, &rho2, 8);
    memcpy(&km, &margin, 8);
    const auto key = std::make_tuple(dev, Q, kr, km);
    std::lock_guard<std::mutex> lock(g_table5_mutex);
    auto it = g_tables5.find(key);
    if (it != g_tables5.end()) { *out = it->second; return NBR_OK; }
    uint32_t *t = nullptr;
    const size_t bins = (size_t)Q * Q * Q;
    NBR_CUDA(cudaMalloc(&t, bins * R5_BIN_WORDS * sizeof(uint32_t)));
    ball_table5_kernel<<<(unsigned)ceil_div((int64_t)(bins * N11), 256), 256, 0, stream>>>(t, Q, rho2, margin);
    NBR_LAUNCHED();
    NBR_CUDA(cudaStreamSynchronize(stream));       // once per distinct (rho, margin): later callers may be on other streams
    g_tables5[key] = t;
    *out = t;
    return NBR_OK;
}

int ball_tables5_trim()
{
    // the current device's tables only (see ball_tables_trim)
    int dev = 0;
    NBR_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lock(g_table5_mutex);
    size_t mine = 0;
    for (auto &kv : g_tables5) mine += std::get<0>(kv.first) == dev;
    if (mine < 128) return NBR_OK;
    NBR_CUDA(cudaDeviceSynchronize());
    for (auto it = g_tables5.begin(); it != g_tables5.end();) {
        if (std::get<0>(it->first) == dev) { cudaFree(const_cast<uint32_t *>(it->second)); it = g_tables5.erase(it); }
        else ++it;
    }
    return NBR_OK;
}

// 11-bit row -> {count | sum(pos) << 8 | sum(pos^2) << 19,  count | sum(pos) << 12}: the first word adds up a
// whole slab (121 cells) without overflow, the second its jy-weighted sums.  the table is split: entry(row) =
// lo[row & 63] + hi[row >> 6] (the fields are sums over the set bits, the high part carries its positions 6..10):
// 96 entries (768 bytes) instead of 2048 (16 KB of shared memory per block, which capped the kernel at 12 warps per SM)
#ifndef R5_SPLIT_LUT
#define R5_SPLIT_LUT 1
#endif
__device__ __forceinline__ uint2 row11_entry(uint32_t b, int first_pos = 0)
{
    uint32_t cnt = 0, s1 = 0, s2 = 0;
#pragma unroll
    for (int i = 0; i < N11; ++i)
        if (b & (1u << i)) { cnt += 1; s1 += i + first_pos; s2 += (i + first_pos) * (i + first_pos); }
    return make_uint2(cnt | (s1 << 8) | (s2 << 19), cnt | (s1 << 12));
}

__device__ __forceinline__ void cp_async16_r5(uint32_t smem_addr, const void *gptr)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_addr), "l"(gptr) : "memory");
}

__device__ __forceinline__ double r5_centre(const R3Entry &E, long long k, int a)
{
    return cell_centre(k + (long long)E.cell_lo[a], E.minc[a], E.edge);
}

// one cell, the reference's float64 expression ((dx^2 + dy^2) + dz^2 <= r*r, no fma)
__device__ __noinline__ bool r5_exact_in(const R3Entry &E, double qx, double qy, double qz, int kx, int ky, int kz)
{
    double s = sqdiff(qx, r5_centre(E, kx, 0));
    s = __dadd_rn(s, sqdiff(qy, r5_centre(E, ky, 1)));
    s = __dadd_rn(s, sqdiff(qz, r5_centre(E, kz, 2)));
    return s <= __dmul_rn(E.r, E.r);
}

__device__ __forceinline__ uint2 r5_lut2(const uint2 *lut, uint32_t M)
{
    const uint2 a = lut[M & 63u], b = lut[64u + (M >> 6)];
    return make_uint2(a.x + b.x, a.y + b.y);
}

struct Acc10 {
    int n, sx, sy, sz, sxx, sxy, sxz, syy, syz, szz;
};

// decide the uncertain occupied cells `u` (bit t <-> window cell t) of window row (jy, jz) and add the accepted ones
__device__ __forceinline__ void r5_decide(const R3Entry &E, const double q[3], float fxm, float fym, float fzm, int xa,
                                          int ya, int za, uint32_t u, int jy, int jz, Acc10 &A)
{
    const float dy = fym - (float)jy, dz = fzm - (float)jz;
    const float row2 = fmaf(dy, dy, dz * dz);
    while (u) {
        const int t = __ffs(u) - 1;
        u &= u - 1;
        const float dx = fxm - (float)t;
        const float d2 = fmaf(dx, dx, row2);
        bool in = d2 < E.rho2;
        if (fabsf(d2 - E.rho2) < 1.0e-4f) in = r5_exact_in(E, q[0], q[1], q[2], xa + t, ya + jy, za + jz);
        if (in) {
            A.n += 1; A.sx += t; A.sy += jy; A.sz += jz;
            A.sxx += t * t; A.sxy += t * jy; A.sxz += t * jz;
            A.syy += jy * jy; A.syz += jy * jz; A.szz += jz * jz;
        }
    }
}

template <typename OutT, bool EXT>
__global__ void __launch_bounds__(R5_WARPS * 32, R5_BLOCKS_N)
rows5_kernel(const __grid_constant__ R3Launch P, const void *__restrict__ query, int dtype,
             const uint32_t *__restrict__ perm, int64_t nq, OutT *__restrict__ out, int64_t row_stride)
{
    // dynamic shared memory, per warp: brick window | parked rows
    extern __shared__ __align__(128) unsigned char smem_raw[];
#if R5_SPLIT_LUT
    __shared__ uint2 s_lut[64 + 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 96; i += blockDim.x) s_lut[i] = i < 64 ? row11_entry(i) : row11_entry(i - 64, 6);
#define R5_LUT(M) r5_lut2(s_lut, (M))
#else
    __shared__ uint2 s_lut[1 << N11];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < (1 << N11); i += blockDim.x) s_lut[i] = row11_entry(i);
#define R5_LUT(M) s_lut[(M)]
#endif
    unsigned char *warp_base = smem_raw + (size_t)warp * (R5_WIN_BYTES + R5_ULIST_BYTES);
    const uint32_t *win = reinterpret_cast<const uint32_t *>(warp_base);
    const uint32_t win_addr = (uint32_t)__cvta_generic_to_shared(win);
    __syncthreads();
    const int64_t n_groups = (nq + 31) >> 5;
    constexpr uint32_t rowmask = (1u << N11) - 1u;

    for (int64_t grp = (int64_t)blockIdx.x * R5_WARPS + warp; grp < n_groups; grp += (int64_t)gridDim.x * R5_WARPS) {
        const int64_t slot_i = grp * 32 + lane;
        const bool active = slot_i < nq;
        const int64_t src = active ? slot_i : grp * 32;             // inactive lanes shadow lane 0
        const int64_t qi = perm ? (int64_t)perm[src] : src;         // row of the output
        double q[3];
#pragma unroll
        for (int a = 0; a < 3; ++a) q[a] = load_coord(query, dtype, src, 3, a);
        OutT *dst_row = out + qi * row_stride;

        int c0 = 0, c1 = 0, c2 = 0, tbin = 0;
        float fxm = 0.f, fym = 0.f, fzm = 0.f;
        int lo0 = 0, lo1 = 0, lo2 = 0, nb0 = 1, nb1 = 1;
        bool staged = false;

        for (int li = 0; li < P.n; ++li) {
            const R3Entry &E = P.e[li];
            if (!E.reuse) {
                // ---- anchor cell, fractional position, table bin
                double f[3];
                int c[3];
#pragma unroll
                for (int a = 0; a < 3; ++a) {
                    const double u = (q[a] - E.minc[a]) * E.inv_edge;
                    const double cf = floor(u);
                    f[a] = u - cf;
                    // saturating conversion + integer clamp (see rows3.cu)
                    c[a] = max(min(__double2int_rn(cf - (double)E.cell_lo[a]), 1000000000), -1000000000);
                }
                c0 = c[0]; c1 = c[1]; c2 = c[2];
                fxm = (float)f[0] + 4.5f; fym = (float)f[1] + 4.5f; fzm = (float)f[2] + 4.5f;
                const int tq = P.tq;
                tbin = (min((int)(f[2] * tq), tq - 1) * tq + min((int)(f[1] * tq), tq - 1)) * tq +
                       min((int)(f[0] * tq), tq - 1);

                // ---- brick window of the whole warp
                lo0 = (__reduce_min_sync(0xffffffffu, c0) - W5) >> BRICK_XS;
                lo1 = (__reduce_min_sync(0xffffffffu, c1) - W5) >> BRICK_YS;
                lo2 = (__reduce_min_sync(0xffffffffu, c2) - W5) >> BRICK_ZS;
                const long long n0 = (long long)((__reduce_max_sync(0xffffffffu, c0) + W5) >> BRICK_XS) - lo0 + 1;
                const long long n1 = (long long)((__reduce_max_sync(0xffffffffu, c1) + W5) >> BRICK_YS) - lo1 + 1;
                const long long n2 = (long long)((__reduce_max_sync(0xffffffffu, c2) + W5) >> BRICK_ZS) - lo2 + 1;
                staged = n0 <= R5_CAP && n1 <= R5_CAP && n2 <= R5_CAP && n0 * n1 * n2 <= R5_CAP;
                nb0 = (int)n0; nb1 = (int)n1;
                if (P.stats && lane == 0) atomicAdd(P.stats + li * 8 + (staged ? 0 : 1), 1ull);
                __syncwarp();                                      // every lane is done reading the previous window
                if (staged) {
                    // brick (ix,iy,iz) -> win[((iz*nb1)+iy)*nb0+ix][32]; empty and out-of-range bricks copy slot 0 (zeros)
                    const int total = nb0 * nb1 * (int)n2;
                    const float inv0 = rcp_fast((float)nb0), inv01 = rcp_fast((float)(nb0 * nb1));
                    uint32_t slot0 = 0, slot1 = 0;
#pragma unroll
                    for (int t = 0; t < 2; ++t) {
                        const int b = lane + 32 * t;
                        uint32_t sl = 0;
                        if (b < total) {
                            const int iz = (int)(((float)b + 0.5f) * inv01);
                            const int rem = b - iz * nb0 * nb1;
                            const int iy = (int)(((float)rem + 0.5f) * inv0), ix = rem - iy * nb0;
                            const int gx = lo0 + ix, gy = lo1 + iy, gz = lo2 + iz;
                            if (gx >= 0 && gx < E.nbx && gy >= 0 && gy < E.nby && gz >= 0 && gz < E.nbz)
                                sl = E.dir[((int64_t)gz * E.nby + gy) * E.nbx + gx];
                        }
                        if (t == 0) slot0 = sl; else slot1 = sl;
                    }
                    const int sub = lane >> 3, chunk = lane & 7;
                    for (int b0 = 0; b0 < total; b0 += 4) {
                        const int b = b0 + sub;
                        const uint32_t sl = __shfl_sync(0xffffffffu, b < 32 ? slot0 : slot1, b & 31);
                        if (b < total)
                            cp_async16_r5(win_addr + (uint32_t)(b * BRICK_WORDS + chunk * 4) * 4u,
                                          E.pool + (int64_t)sl * BRICK_WORDS + chunk * 4);
                    }
                    asm volatile("cp.async.wait_all;" ::: "memory");
                    __syncwarp();
                }
            }

            // ---- per lane: 11 slabs of 11 rows of 11 bits
            const uint4 *tab = reinterpret_cast<const uint4 *>(reinterpret_cast<const uint32_t *>(E.table) + (size_t)tbin * R5_BIN_WORDS);
            const int xa = c0 - W5, ya = c1 - W5, za = c2 - W5;
            const int sh = xa & 31;
            const bool two = sh + N11 > 32;
            const int bx0 = xa >> BRICK_XS, by0 = ya >> BRICK_YS;
            const int ycross = BRICK_Y - (ya & (BRICK_Y - 1));      // rows jy >= ycross live in the next y-brick, >= ycross + 8 one further
            int ybase = 0, ystep = 0, zstride = 0;
            if (staged) {
                ybase = ((by0 - lo1) * nb0 + (bx0 - lo0)) * BRICK_WORDS + (ya & (BRICK_Y - 1));
                ystep = nb0 * BRICK_WORDS - BRICK_Y;
                zstride = nb1 * nb0 * BRICK_WORDS;
            }
            Acc10 A = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
            unsigned long long uf0 = 0, uf1 = 0, uf2 = 0;      // rows with parked cells: slabs 0..3 | 4..7 | 8..10
            uint4 t0 = tab[0], t1 = tab[1], t2 = tab[2];
#pragma unroll 1
            for (int jz = 0; jz < N11; ++jz) {
                const uint32_t T[12] = {t0.x, t0.y, t0.z, t0.w, t1.x, t1.y, t1.z, t1.w, t2.x, t2.y, t2.z, t2.w};
                if (jz + 1 < N11) { t0 = tab[3 * (jz + 1)]; t1 = tab[3 * (jz + 1) + 1]; t2 = tab[3 * (jz + 1) + 2]; }
                if ((T[0] | T[1] | T[2] | T[3] | T[4] | T[5] | T[6] | T[7] | T[8] | T[9] | T[10]) == 0) continue;
                const int az = za + jz;
                const int wz = (az & (BRICK_Z - 1)) << BRICK_YS;
                const int gz = az >> BRICK_ZS;
                uint32_t Pk = 0, Qk = 0, uflags = 0;
                int R = 0;
                if (staged) {
                    const int zoff = (gz - lo2) * zstride + wz + ybase;
#pragma unroll
                    for (int jy = 0; jy < N11; ++jy) {
                        const int off = zoff + jy + (jy >= ycross ? ystep : 0) + (jy >= ycross + BRICK_Y ? ystep : 0);
                        const uint32_t w0 = win[off];
                        const uint32_t w1 = two ? win[off + BRICK_WORDS] : 0u;
                        const uint32_t bits = __funnelshift_r(w0, w1, sh) & rowmask;
                        const uint32_t M = bits & T[jy], U = bits & (T[jy] >> 16);
                        // rows with uncertain occupied cells are flagged and decided after the rows
                        uflags |= (U ? 1u : 0u) << jy;
                        const uint2 e = R5_LUT(M);
                        Pk += e.x;
                        Qk += jy * e.y;
                        R += jy * jy * (int)(e.x & 255u);
                    }
                } else {
                    // the warp's window does not fit: rows straight from the directory + pool (one brick pair cached)
                    if (gz < 0 || gz >= E.nbz) continue;
                    int cached_gy = 0x7fffffff;
                    uint32_t slot_a = 0, slot_b = 0;
#pragma unroll 1
                    for (int jy = 0; jy < N11; ++jy) {
                        const uint32_t tw = jy < 4 ? (jy < 2 ? (jy == 0 ? T[0] : T[1]) : (jy == 2 ? T[2] : T[3]))
                                          : jy < 8 ? (jy < 6 ? (jy == 4 ? T[4] : T[5]) : (jy == 6 ? T[6] : T[7]))
                                                   : (jy == 8 ? T[8] : (jy == 9 ? T[9] : T[10]));
                        if (tw == 0) continue;
                        const int ay = ya + jy;
                        const int gy = ay >> BRICK_YS;
                        if (gy < 0 || gy >= E.nby) continue;
                        if (gy != cached_gy) {
                            cached_gy = gy;
                            const int64_t rowb = ((int64_t)gz * E.nby + gy) * E.nbx;
                            slot_a = (bx0 >= 0 && bx0 < E.nbx) ? E.dir[rowb + bx0] : 0u;
                            slot_b = (two && bx0 + 1 >= 0 && bx0 + 1 < E.nbx) ? E.dir[rowb + bx0 + 1] : 0u;
                        }
                        if ((slot_a | slot_b) == 0) continue;
                        const int word = wz | (ay & (BRICK_Y - 1));
                        const uint32_t w0 = slot_a ? E.pool[(int64_t)slot_a * BRICK_WORDS + word] : 0u;
                        const uint32_t w1 = slot_b ? E.pool[(int64_t)slot_b * BRICK_WORDS + word] : 0u;
                        const uint32_t bits = __funnelshift_r(w0, w1, sh) & rowmask;
                        const uint32_t M = bits & tw, U = bits & (tw >> 16);
                        uflags |= (U ? 1u : 0u) << jy;
                        const uint2 e = R5_LUT(M);
                        Pk += e.x;
                        Qk += jy * e.y;
                        R += jy * jy * (int)(e.x & 255u);
                    }
                }
                {
                    const unsigned long long fl = (unsigned long long)uflags << (N11 * (jz & 3));
                    if (jz < 4) uf0 |= fl; else if (jz < 8) uf1 |= fl; else uf2 |= fl;
                }
                const int C = Pk & 255, SX = (Pk >> 8) & 2047, SXX = Pk >> 19;
                const int SY = Qk & 4095, SXY = Qk >> 12;
                A.n += C; A.sx += SX; A.sxx += SXX; A.sy += SY; A.syy += R; A.sxy += SXY;
                A.sz += jz * C; A.szz += jz * jz * C; A.sxz += jz * SX; A.syz += jz * SY;
            }
            // ---- parked rows: occupied cells of the uncertain shell
#pragma unroll 1
            for (int part = 0; part < 3; ++part) {
                unsigned long long fl = part == 0 ? uf0 : (part == 1 ? uf1 : uf2);
                while (fl) {
                    const int b = __ffsll((long long)fl) - 1;
                    fl &= fl - 1;
                    const int rid = b + part * 4 * N11;
                    const int jz = (rid * 373) >> 12, jy = rid - N11 * jz;          // rid / 11 for rid < 121
                    // the row's occupancy again (a flagged row is rare: a handful per query), and its uncertain mask
                    const int az = za + jz, ay = ya + jy;
                    const int wz = (az & (BRICK_Z - 1)) << BRICK_YS, gz = az >> BRICK_ZS;
                    uint32_t w0 = 0, w1 = 0;
                    if (staged) {
                        const int off = (gz - lo2) * zstride + wz + ybase + jy + (jy >= ycross ? ystep : 0) + (jy >= ycross + BRICK_Y ? ystep : 0);
                        w0 = win[off];
                        w1 = two ? win[off + BRICK_WORDS] : 0u;
                    } else {
                        const int gy = ay >> BRICK_YS;
                        const int64_t rowb = ((int64_t)gz * E.nby + gy) * E.nbx;             // flagged rows are inside the directory
                        const uint32_t slot_a = (bx0 >= 0 && bx0 < E.nbx) ? E.dir[rowb + bx0] : 0u;
                        const uint32_t slot_b = (two && bx0 + 1 >= 0 && bx0 + 1 < E.nbx) ? E.dir[rowb + bx0 + 1] : 0u;
                        const int word = wz | (ay & (BRICK_Y - 1));
                        w0 = slot_a ? E.pool[(int64_t)slot_a * BRICK_WORDS + word] : 0u;
                        w1 = slot_b ? E.pool[(int64_t)slot_b * BRICK_WORDS + word] : 0u;
                    }
                    const uint32_t tw = reinterpret_cast<const uint32_t *>(tab)[jz * R5_SLAB_WORDS + jy];
                    const uint32_t U = __funnelshift_r(w0, w1, sh) & rowmask & (tw >> 16);
                    r5_decide(E, q, fxm, fym, fzm, xa, ya, za, U, jy, jz, A);
                }
            }
            if (active)
                emit_features_window<OutT>(A.n, A.sx, A.sy, A.sz, A.sxx, A.sxy, A.sxz, A.syy, A.syz, A.szz, fxm, fym, fzm,
                                           true, E.edge, dst_row + E.col, EXT ? NBR_DESC_EXTENDED : 0);
        }
    }
}

// fills one entry; false if this (lattice, radius) does not fit an 11x11x11 window or tables are disabled
bool rows5_entry(const Lattice *lat, double radius, int col, const R3Entry *prev, R3Entry *E, int *tq_io,
                 cudaStream_t stream, int *rc)
{
    *rc = NBR_OK;
    static const bool disabled = getenv("NBR_NO_ROWS5") != nullptr || getenv("NBR_NO_BALL_TABLE") != nullptr;
    if (disabled) return false;
    const double e = lat->grid.edge;
    const double rho = radius / e;
    if (!(rho + 0.5 + 1e-6 < 6.0)) return false;
    static const int q_env = getenv("NBR_BALL_Q5") ? atoi(getenv("NBR_BALL_Q5")) : 0;
    const int tq = q_env >= 1 && q_env <= 32 ? q_env : 16;
    *tq_io = tq;
    const LatticeDev d = lat->dev();
    for (int a = 0; a < 3; ++a) { E->minc[a] = d.g.minc[a]; E->cell_lo[a] = d.g.cell_lo[a]; }
    E->edge = d.g.edge; E->inv_edge = d.g.inv_edge; E->r = radius;
    E->dir = d.dir; E->pool = d.pool;
    E->nbx = d.nbx; E->nby = d.nby; E->nbz = d.nbz;
    E->rho2 = (float)(rho * rho);
    E->col = col;
    E->reuse = prev && prev->pool == d.pool && prev->dir == d.dir;
    double maxabs = 0.0;
    for (int a = 0; a < 3; ++a)
        maxabs = std::max(maxabs, std::max(fabs(lat->grid.min_corner[a]), fabs(lat->grid.max_corner[a])));
    const double margin = std::max(1e-11, 64.0 * 2.3e-16 * (maxabs / e + 12.0));
    const uint32_t *table = nullptr;
    *rc = ball_table5_get(rho * rho, margin, tq, &table, stream);
    E->table = reinterpret_cast<const uint4 *>(table);
    return *rc == NBR_OK;
}

int rows5_launch(const R3Launch *L, const void *query, int dtype, const uint32_t *perm, int64_t nq, void *out,
                 int out_dtype, int64_t row_stride, int descriptor_mask, cudaStream_t stream)
{
    if (nq <= 0 || L->n <= 0) return NBR_OK;
    R3Launch copy = *L;
    copy.stats = nullptr;
    const int blocks = (int)std::min<int64_t>(ceil_div(ceil_div(nq, 32), R5_WARPS), (int64_t)device_sm_count() * 12);
    const bool ext = (descriptor_mask & NBR_DESC_EXTENDED) != 0;
    const size_t smem = (size_t)R5_WARPS * (R5_WIN_BYTES + R5_ULIST_BYTES);
    static std::atomic<uint64_t> configured{0};
    if (first_use_on_device(configured)) {
        NBR_CUDA(cudaFuncSetAttribute(rows5_kernel<float, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        NBR_CUDA(cudaFuncSetAttribute(rows5_kernel<float, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        NBR_CUDA(cudaFuncSetAttribute(rows5_kernel<double, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        NBR_CUDA(cudaFuncSetAttribute(rows5_kernel<double, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
#define R5_GO(T, X) rows5_kernel<T, X><<<blocks, R5_WARPS * 32, smem, stream>>>(copy, query, dtype, perm, nq, (T *)out, row_stride)
    if (out_dtype == NBR_F32) { if (ext) R5_GO(float, true); else R5_GO(float, false); }
    else                      { if (ext) R5_GO(double, true); else R5_GO(double, false); }
#undef R5_GO
    NBR_LAUNCHED();
    return NBR_OK;
}

}  // namespace nbr
