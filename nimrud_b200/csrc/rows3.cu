// rows3.cu -- the lean fused feature kernel for 7x7x7 windows (every scale with r/e < 3.5).
//
// same job as radius_rows.cu (nimrud/minimal/multiscale.py:94-122 for every scale of a call in one
// launch: neighbor search, take, population, centroid, pca) but only the common case, so that the code
// is small enough for the instruction cache and light enough on registers for 20+ warps per SM:
//   * launch description in kernel parameters (constant bank), one entry per (lattice, radius);
//   * the warp's brick window is staged into shared memory by 16-byte asynchronous copies (LDGSTS, 8 lanes
//     per 128-byte brick); warps whose window does not fit read their rows from global memory through the
//     directory.  measured and dropped: 128-byte TMA bulk copies on an mbarrier (69M tiny copies per step are
//     bound by the small-copy rate: 5.90 vs 5.79 ms), and a row-major window layout whose scattered LDGSTS
//     destinations (one 32-byte piece per z of a brick) congest the return path of every global load
//     (long-scoreboard stalls x4: 6.3 vs 4.4 ms, profiles/r01_rows3_ncu_summary.md);
//   * membership comes from the shell tables (ball_table.cu): cells inside for the whole bin of the
//     query's fractional position are a mask, occupied cells of the uncertain shell are tested in
//     float32 and, inside the rounding band, with the reference's float64 expression -- neighbor sets
//     stay bit-exact;
//   * moments are exact integers (7-bit row table), features come out of finalize.cuh in the same
//     kernel.
#include "common.cuh"
#include "finalize.cuh"
#include "lattice.cuh"
#include "radius_rows.cuh"
#include "mailbox.cuh"

#include <stdlib.h>

namespace nbr {

#ifndef R3_WARPS_N
#define R3_WARPS_N 4
#endif
#ifndef R3_CAP_N
#define R3_CAP_N 48
#endif
#ifndef R3_W1_ALWAYS
#define R3_W1_ALWAYS 1
#endif
#ifndef R3_ROW_BITS
#define R3_ROW_BITS 8               // bits per window row in the slab words and the shell tables (7: packed, 8: one byte per row)
#endif
#ifndef R3_UNROLL_N
#define R3_UNROLL_N 1
#endif
#ifndef R3_BLOCKS_N
#define R3_BLOCKS_N 4
#endif
#ifndef R3_TMA_ROWS
#define R3_TMA_ROWS 0               // 1: finished rows leave shared memory as 80-byte bulk copies (TMA, UBLKCP) instead of through registers. measured: 4.128 vs 4.091 ms (10M tiny copies per step are bound by the engine's small-copy rate)
#endif
#ifndef R3_TMA_PEER
#define R3_TMA_PEER 1               // MULTI launches: a warp's 32 finished rows leave for every peer as ONE bulk copy (TMA, UBLKCP: 2560 contiguous bytes) instead of 5 16-byte stores per lane and peer
#endif
#ifndef R3_PIPELINE
#define R3_PIPELINE 0               // stage entry li + 1 before the eigen-solve of entry li
#endif
#ifndef R3_UNC2
#define R3_UNC2 0                   // two uncertain cells per trip of the shell loop
#endif
#ifndef R3_TAB_PREFETCH
#define R3_TAB_PREFETCH 0           // the next slab's table masks are loaded one slab ahead
#endif
// NBR_BOUNDS_CHECK=1 (debug builds; compute-sanitizer is closed on the pool this was developed on): every shared-memory
// index of the window gather, the parked-cell list and the row buffer is checked against the warp's own buffers and
// violations are counted (nbr_debug_bounds_violations()).  scripts/bounds_check.sh builds and runs such a variant.
#ifndef NBR_BOUNDS_CHECK
#define NBR_BOUNDS_CHECK 0
#endif
__device__ unsigned long long g_r3_violations = 0;
#if NBR_BOUNDS_CHECK
#define R3_CHECK(cond) do { if (!(cond)) atomicAdd(&g_r3_violations, 1ull); } while (0)
#else
#define R3_CHECK(cond) do { } while (0)
#endif
constexpr int R3_WARPS = R3_WARPS_N;
constexpr int R3_UNROLL = R3_UNROLL_N;   // slabs per trip of the slab loop
constexpr int R3_CAP = R3_CAP_N;          // staged bricks per warp (128 bytes each)
constexpr int N7 = 7, W3 = 3;
constexpr int R3_TAB_STRIDE = 28;        // words per lane: 7 slabs x 16 bytes; conflict-free for LDS.128
constexpr int R3_WIN_BYTES = R3_CAP * BRICK_WORDS * 4;
constexpr int R3_TAB_BYTES = 32 * R3_TAB_STRIDE * 4;
constexpr int R3_MAX_ROW_BYTES = 256;   // output rows up to this size are assembled in shared memory

__device__ __forceinline__ uint32_t row7_entry(uint32_t b)
{
    uint32_t cnt = 0, s1 = 0, s2 = 0;
#pragma unroll
    for (int i = 0; i < 7; ++i)
        if (b & (1u << i)) { cnt += 1; s1 += i; s2 += i * i; }
    return cnt | (s1 << 10) | (s2 << 20);
}

__device__ __forceinline__ void cp_async16(uint32_t smem_addr, const void *gptr)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_addr), "l"(gptr) : "memory");
}

// slot of brick (gx, gy, gz); 0 (the all-zero brick) outside the directory.  directories hold < 2^32 entries
// (lattice_create refuses larger ones), so the index arithmetic is 32-bit
__device__ __forceinline__ uint32_t dir_slot(const R3Entry &E, int gx, int gy, int gz, bool wanted = true)
{
    const bool ok = wanted && (uint32_t)gx < (uint32_t)E.nbx && (uint32_t)gy < (uint32_t)E.nby && (uint32_t)gz < (uint32_t)E.nbz;
    const uint32_t idx = ((uint32_t)gz * (uint32_t)E.nby + (uint32_t)gy) * (uint32_t)E.nbx + (uint32_t)gx;
    return ok ? E.dir[idx] : 0u;
}

__device__ __forceinline__ double r3_centre(const R3Entry &E, long long k, int a)
{
    return cell_centre(k + (long long)E.cell_lo[a], E.minc[a], E.edge);
}

// one cell, the reference's float64 expression ((dx^2 + dy^2) + dz^2 <= r*r, no fma)
__device__ __noinline__ bool r3_exact_in(const R3Entry &E, double qx, double qy, double qz, int kx, int ky, int kz)
{
    double s = sqdiff(qx, r3_centre(E, kx, 0));
    s = __dadd_rn(s, sqdiff(qy, r3_centre(E, ky, 1)));
    s = __dadd_rn(s, sqdiff(qz, r3_centre(E, kz, 2)));
    return s <= __dmul_rn(E.r, E.r);
}

// MULTI: the finished rows are the rank's share of a feature all-gather -- besides its place in the rank's own result,
// every row is stored into the staging buffer of every OTHER rank (D.base[d], peer-mapped over NVLink) in processing
// order (a warp's 32 rows are one contiguous piece) together with its row number, so the exchange rides on the
// kernel's own write-out instead of following it as a collective; the receivers put the rows in place (mailbox.cu)
template <typename OutT, bool EXT, bool MULTI>
__global__ void __launch_bounds__(R3_WARPS * 32, R3_BLOCKS_N)
rows3_kernel(const __grid_constant__ R3Launch P, const void *__restrict__ query, int dtype,
             const uint32_t *__restrict__ perm, int64_t nq, OutT *__restrict__ out, int64_t row_stride, int stage_rows,
             const __grid_constant__ RowDests D)
{
    // dynamic shared memory, per warp: brick window | table lines | output rows (when the launch owns whole rows)
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ uint32_t s_lut[128];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 128; i += blockDim.x) s_lut[i] = row7_entry(i);
    const int row_bytes = stage_rows ? (int)row_stride * (int)sizeof(OutT) : 0;
    unsigned char *warp_base = smem_raw + (size_t)warp * (R3_WIN_BYTES + R3_TAB_BYTES + 32 * row_bytes);
    const uint32_t *win = reinterpret_cast<const uint32_t *>(warp_base);
    const uint32_t win_addr = (uint32_t)__cvta_generic_to_shared(win);
    const uint4 *tab = reinterpret_cast<const uint4 *>(warp_base + R3_WIN_BYTES + lane * R3_TAB_STRIDE * 4);
    const uint32_t tab_addr = (uint32_t)__cvta_generic_to_shared(tab);
    unsigned char *rows = warp_base + R3_WIN_BYTES + R3_TAB_BYTES;       // [32][row_bytes]
    __syncthreads();
    // row buffer -> global: 16-byte pieces per row, and where lane's first piece sits
    const int cpr = max(row_bytes >> 4, 1);
    const int r_step = 32 / cpr, c_step = 32 - r_step * cpr;
    const int r_first = lane / cpr, c_first = lane - r_first * cpr;
    const int64_t n_groups = (nq + 31) >> 5;
    constexpr uint32_t rowmask = 127u;

    for (int64_t grp = (int64_t)blockIdx.x * R3_WARPS + warp; grp < n_groups; grp += (int64_t)gridDim.x * R3_WARPS) {
        const int64_t slot_i = grp * 32 + lane;
        const bool active = slot_i < nq;
        const int64_t src = active ? slot_i : grp * 32;             // inactive lanes shadow lane 0
        const int64_t qi = perm ? (int64_t)perm[src] : src;         // row of the output
        double q[3];
#pragma unroll
        for (int a = 0; a < 3; ++a) q[a] = load_coord(query, dtype, src, 3, a);
        // features go straight to the query's row, or (the launch covers whole rows) to the warp's row buffer,
        // which is written out as contiguous rows after the last lattice: the query order scatters the rows,
        // and 16-byte pieces of scattered rows written lattice by lattice cost a partial-sector fill each
        OutT *dst_row = stage_rows ? reinterpret_cast<OutT *>(rows + lane * row_bytes) : out + qi * row_stride;
#if R3_TMA_ROWS
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // the previous group's rows have been read out of the buffer
#elif R3_TMA_PEER
        if (MULTI) {
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // the copies to the peers have read the previous group's rows
            __syncwarp();
        }
#endif

        // state of the current lattice (kept across entries that share it)
        int c0 = 0, c1 = 0, c2 = 0, tbin = 0;
        float fxm = 0.f, fym = 0.f, fzm = 0.f;
        int lo0 = 0, lo1 = 0, lo2 = 0, nb0 = 1, nb1 = 1;
        bool staged = false;

        // stage entry li: anchor cell, table bin, the warp's brick window and every lane's table line (asynchronous
        // copies; the caller waits for them).  called for entry li + 1 BEFORE entry li's eigen-solve, so that the
        // directory look-ups and the copies are in flight while the warp does arithmetic
        auto stage_entry = [&](int li) {
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
                    // local cell coordinate: the conversion saturates, the integer clamp keeps the window arithmetic
                    // of far-away queries inside int32 (float64 min/max cost 16 instructions per axis)
                    c[a] = max(min(__double2int_rn(cf - (double)E.cell_lo[a]), 1000000000), -1000000000);
                }
                c0 = c[0]; c1 = c[1]; c2 = c[2];
                fxm = (float)f[0] + 2.5f; fym = (float)f[1] + 2.5f; fzm = (float)f[2] + 2.5f;
                const int tq = P.tq;
                tbin = (min((int)(f[2] * tq), tq - 1) * tq + min((int)(f[1] * tq), tq - 1)) * tq +
                       min((int)(f[0] * tq), tq - 1);

                // ---- brick window of the whole warp
                lo0 = (__reduce_min_sync(0xffffffffu, c0) - W3) >> BRICK_XS;
                lo1 = (__reduce_min_sync(0xffffffffu, c1) - W3) >> BRICK_YS;
                lo2 = (__reduce_min_sync(0xffffffffu, c2) - W3) >> BRICK_ZS;
                const long long n0 = (long long)((__reduce_max_sync(0xffffffffu, c0) + W3) >> BRICK_XS) - lo0 + 1;
                const long long n1 = (long long)((__reduce_max_sync(0xffffffffu, c1) + W3) >> BRICK_YS) - lo1 + 1;
                const long long n2 = (long long)((__reduce_max_sync(0xffffffffu, c2) + W3) >> BRICK_ZS) - lo2 + 1;
                staged = n0 <= R3_CAP && n1 <= R3_CAP && n2 <= R3_CAP && n0 * n1 * n2 <= R3_CAP;
                nb0 = (int)n0; nb1 = (int)n1;
                if (P.stats) {
                    // diagnostics: staged / direct warps, and the number of distinct anchor cells among the warp's queries
                    const unsigned long long key = ((unsigned long long)(uint32_t)c0 << 42) ^ ((unsigned long long)(uint32_t)c1 << 21) ^ (uint32_t)c2;
                    const uint32_t peers = __match_any_sync(0xffffffffu, key);
                    const int leaders = __popc(__ballot_sync(0xffffffffu, (peers & lanemask_lt()) == 0));
                    if (lane == 0) {
                        atomicAdd(P.stats + li * 8 + (staged ? 0 : 1), 1ull);
                        atomicAdd(P.stats + li * 8 + 2, (unsigned long long)leaders);
                        if (leaders <= 8) atomicAdd(P.stats + li * 8 + 3, 1ull);
                    }
                }
                // stage every lane's table line and (window fits) the bricks of the warp's window:
                // brick (ix,iy,iz) -> win[((iz*nb1)+iy)*nb0+ix][32]; empty and out-of-range bricks copy slot 0 (zeros)
                const int total = staged ? nb0 * nb1 * (int)n2 : 0;
                __syncwarp();                                      // every lane is done reading the previous window / lines
                // 16-byte asynchronous copies (LDGSTS): 7 per lane for its table line, then 4 bricks per step
                // (8 lanes x 16 bytes per brick).  measured: ~5x the small-copy throughput of the TMA path
                {
                    const char *line = reinterpret_cast<const char *>(E.table + (size_t)tbin * 8);
#pragma unroll
                    for (int j = 0; j < N7; ++j) cp_async16(tab_addr + 16u * j, line + 16 * j);
                }
                if (total) {
                    const float inv0 = rcp_fast((float)nb0), inv01 = rcp_fast((float)(nb0 * nb1));
                    uint32_t slot0 = 0, slot1 = 0;
#pragma unroll
                    for (int t = 0; t < 2; ++t) {
                        const int b = lane + 32 * t;
                        uint32_t sl = 0;
                        if (b < total) {
                            // b < 64 and nb0 * nb1 <= 48: (b + 0.5) / n is at least 0.01 away from an integer, so the float
                            // quotients (1-ulp reciprocal) are exact after truncation
                            const int iz = (int)(((float)b + 0.5f) * inv01);
                            const int rem = b - iz * nb0 * nb1;
                            const int iy = (int)(((float)rem + 0.5f) * inv0), ix = rem - iy * nb0;
                            sl = dir_slot(E, lo0 + ix, lo1 + iy, lo2 + iz);
                        }
                        if (t == 0) slot0 = sl; else slot1 = sl;
                    }
                    const int sub = lane >> 3, chunk = lane & 7;
                    for (int b0 = 0; b0 < total; b0 += 4) {
                        const int b = b0 + sub;
                        const uint32_t sl = __shfl_sync(0xffffffffu, b < 32 ? slot0 : slot1, b & 31);
                        if (b < total)
                            cp_async16(win_addr + (uint32_t)(b * BRICK_WORDS + chunk * 4) * 4u,
                                       E.pool + (size_t)sl * BRICK_WORDS + chunk * 4);
                    }
                }
            } else {
                // same lattice, another radius: only the table lines change
                __syncwarp();
                const char *line = reinterpret_cast<const char *>(E.table + (size_t)tbin * 8);
#pragma unroll
                for (int j = 0; j < N7; ++j) cp_async16(tab_addr + 16u * j, line + 16 * j);
            }
        };
        // trip li = -1 only stages entry 0; trip li computes entry li, stages entry li + 1, then solves entry li
        // (one copy of every phase in the instruction stream: the kernel is sensitive to its code size)
        for (int li = R3_PIPELINE ? -1 : 0; li < P.n; ++li) {
            int An = 0, Asx = 0, Asy = 0, Asz = 0, Asxx = 0, Asxy = 0, Asxz = 0, Asyy = 0, Asyz = 0, Aszz = 0;
            const R3Entry &E = P.e[li < 0 ? 0 : li];
            if (!R3_PIPELINE) stage_entry(li);
            if (li >= 0) {
            // ---- per lane: 7 slabs of 7 rows of 7 bits
            const int xa = c0 - W3, ya = c1 - W3, za = c2 - W3;
            const int sh = xa & 31;
            const bool two = sh + N7 > 32;
            const int bx0 = xa >> BRICK_XS, by0 = ya >> BRICK_YS, bz0 = za >> BRICK_ZS;
            const int ycross = BRICK_Y - (ya & (BRICK_Y - 1));      // rows jy >= ycross live in the next y-brick
            int ybase = 0, ystep = 0, zstride = 0;
            uint32_t slot[3][2][2];
            if (staged) {
                ybase = ((by0 - lo1) * nb0 + (bx0 - lo0)) * BRICK_WORDS + (ya & (BRICK_Y - 1));
                ystep = nb0 * BRICK_WORDS - BRICK_Y;
                zstride = nb1 * nb0 * BRICK_WORDS;
            } else {
#pragma unroll
                for (int iz = 0; iz < 3; ++iz)
#pragma unroll
                    for (int iy = 0; iy < 2; ++iy)
#pragma unroll
                        for (int ix = 0; ix < 2; ++ix)
                            slot[iz][iy][ix] = dir_slot(E, bx0 + ix, by0 + iy, bz0 + iz, ix == 0 || two);
            }
            asm volatile("cp.async.wait_all;" ::: "memory");
            __syncwarp();
            uint2 *ulist = reinterpret_cast<uint2 *>(const_cast<uint4 *>(tab));
            int n_u = 0;
#if R3_TAB_PREFETCH
            uint4 tnext = tab[0];
#endif
#pragma unroll R3_UNROLL
            for (int jz = 0; jz < N7; ++jz) {
#if R3_TAB_PREFETCH
                const uint4 tcur = tnext;
                tnext = tab[min(jz + 1, N7 - 1)];
#else
                const uint4 tcur = tab[jz];
#endif
                // every skip of this loop is warp-uniform (the body synchronises the warp): no lane's cells of this slab
                // can be in the ball
                if (!__any_sync(0xffffffffu, (tcur.x | tcur.y | tcur.z | tcur.w) != 0)) continue;
                const int az = za + jz;
                const int wz = (az & (BRICK_Z - 1)) << BRICK_YS;
                // ---- gather the slab: bits of row jy at [R3_ROW_BITS * jy, + 7)
#if R3_ROW_BITS == 8
                uint32_t v[N7];                                              // row jy in the low 7 bits, neighbours' bits above
#pragma unroll
                for (int jy = 0; jy < N7; ++jy) v[jy] = 0u;
#define R3_PUT_ROW(jy, bits) v[jy] = (bits)
#else
                unsigned long long slab = 0;
#define R3_PUT_ROW(jy, bits) slab |= (unsigned long long)((bits) & rowmask) << (N7 * (jy))
#endif
                if (staged) {
                    const int zoff = ((az >> BRICK_ZS) - lo2) * zstride + wz + ybase;
#pragma unroll
                    for (int jy = 0; jy < N7; ++jy) {
                        const int off = zoff + jy + (jy >= ycross ? ystep : 0);
                        // w0 inside the staged bricks; w1 (read unconditionally) at most one brick further: inside the
                        // window + table area of this warp
                        R3_CHECK(off >= 0 && off < R3_CAP * BRICK_WORDS && off + BRICK_WORDS < (R3_WIN_BYTES + R3_TAB_BYTES) / 4);
                        const uint32_t w0 = win[off];
#if R3_W1_ALWAYS
                        const uint32_t w1 = win[off + BRICK_WORDS];   // unused bits when !two; stays inside the warp's buffers
#else
                        const uint32_t w1 = two ? win[off + BRICK_WORDS] : 0u;
#endif
                        R3_PUT_ROW(jy, __funnelshift_r(w0, w1, sh));
                    }
                } else {
                    const int iz = (az >> BRICK_ZS) - bz0;                     // 0..2
                    const uint32_t s00 = iz == 0 ? slot[0][0][0] : (iz == 1 ? slot[1][0][0] : slot[2][0][0]);
                    const uint32_t s01 = iz == 0 ? slot[0][0][1] : (iz == 1 ? slot[1][0][1] : slot[2][0][1]);
                    const uint32_t s10 = iz == 0 ? slot[0][1][0] : (iz == 1 ? slot[1][1][0] : slot[2][1][0]);
                    const uint32_t s11 = iz == 0 ? slot[0][1][1] : (iz == 1 ? slot[1][1][1] : slot[2][1][1]);
                    if ((s00 | s01 | s10 | s11) != 0) {
#pragma unroll
                        for (int jy = 0; jy < N7; ++jy) {
                            const bool up = jy >= ycross;
                            const uint32_t sa = up ? s10 : s00, sb = up ? s11 : s01;
                            const int word = wz | ((ya + jy) & (BRICK_Y - 1));
                            const uint32_t w0 = sa ? E.pool[(int64_t)sa * BRICK_WORDS + word] : 0u;
                            const uint32_t w1 = sb ? E.pool[(int64_t)sb * BRICK_WORDS + word] : 0u;
                            R3_PUT_ROW(jy, __funnelshift_r(w0, w1, sh));
                        }
                    }
                }
#undef R3_PUT_ROW
                // ---- membership: sure cells + occupied cells of the uncertain shell (low word: rows 0..3, high word: 4..6)
#if R3_ROW_BITS == 8
                // one byte per row: 5 byte permutes assemble the slab; bit 7 of every byte and byte 3 of the high word
                // carry other cells' bits, the table masks are zero there
                const uint32_t slab_lo = __byte_perm(__byte_perm(v[0], v[1], 0x0040), __byte_perm(v[2], v[3], 0x0040), 0x5410);
                const uint32_t slab_hi = __byte_perm(__byte_perm(v[4], v[5], 0x0040), v[6], 0x0410);
#else
                const uint32_t slab_lo = (uint32_t)slab, slab_hi = (uint32_t)(slab >> 32);
#endif
                const uint32_t Mlo = slab_lo & tcur.x, Mhi = slab_hi & tcur.y, Ulo = slab_lo & tcur.z, Uhi = slab_hi & tcur.w;
                // warp-uniform skip only: the body below is straight-line code for every lane (an empty slab adds
                // zeros), which lets the loads of the table lookups overlap instead of ending at divergent branches
                if (!__any_sync(0xffffffffu, (Mlo | Mhi | Ulo | Uhi) != 0)) continue;
                // the uncertain cells are parked in the lane's table line (slots <= jz are consumed) and decided after
                // the slab loop in ONE loop: deciding them slab by slab made every slab wait for its slowest lane
                R3_CHECK(n_u <= 2 * jz && n_u + 1 < 2 * N7);                             // both stores land in consumed slots of the lane's line
                ulist[n_u] = make_uint2(Ulo, (uint32_t)jz);                               // slot n_u <= 2 jz + 1: consumed
                n_u += Ulo != 0;
                ulist[n_u] = make_uint2(Uhi, (uint32_t)jz | 256u);
                n_u += Uhi != 0;
                // ---- moments of the slab
                // packed sums: sum e, sum jy*e, sum jy^2*e (of the last only the count field is read: it cannot be
                // reached by carries from above).  the row's 7 bits are extracted as a byte offset into the table
                uint32_t Pk = 0, Qk = 0, Rk = 0;
                const char *lut_bytes = reinterpret_cast<const char *>(s_lut);
#if R3_ROW_BITS != 8
                const unsigned long long M4 = ((unsigned long long)Mlo | ((unsigned long long)Mhi << 32)) << 2;
#endif
#pragma unroll
                for (int jy = 0; jy < N7; ++jy) {
#if R3_ROW_BITS == 8
                    const uint32_t mw = jy < 4 ? Mlo : Mhi;
                    const int sb = 8 * (jy & 3) - 2;                     // byte jy & 3 of the word, times 4
                    const uint32_t boff = (sb < 0 ? mw << 2 : mw >> sb) & (rowmask << 2);
#else
                    const uint32_t boff = (uint32_t)(M4 >> (N7 * jy)) & (rowmask << 2);
#endif
                    const uint32_t e = *reinterpret_cast<const uint32_t *>(lut_bytes + boff);
                    Pk += e;
                    Qk += jy * e;
                    Rk += jy * jy * e;
                }
                const int C = Pk & 1023, SX = (Pk >> 10) & 1023, SXX = Pk >> 20;
                const int SY = Qk & 1023, SXY = (Qk >> 10) & 1023, R = Rk & 1023;
                An += C; Asx += SX; Asxx += SXX; Asy += SY; Asyy += R; Asxy += SXY;
                Asz += jz * C; Aszz += jz * jz * C; Asxz += jz * SX; Asyz += jz * SY;
            }
            // ---- occupied cells of the uncertain shell: float32 first, |d^2 - rho^2| > band decides; inside the band
            // the reference's float64 expression does.  accepted cells are added to the moments one by one
#if R3_UNC2 && R3_ROW_BITS == 8
            {
                // two cells of the same word per trip: their distance arithmetic overlaps, and the trip count (the
                // maximum over the lanes) halves.  a rejected cell is added with weight 0
                uint32_t cur = 0;
                int k = 0, jzc = 0, half = 0;
                float dz2f = 0.0f;
                for (;;) {
                    if (cur == 0) {
                        if (k >= n_u) break;
                        const uint2 e = ulist[k++];
                        cur = e.x;
                        jzc = (int)(e.y & 255u);
                        half = (int)(e.y >> 8) * 4;
                        const float dzf = fzm - (float)jzc;
                        dz2f = dzf * dzf;
                    }
                    const int i1 = __ffs(cur) - 1;
                    cur &= cur - 1;
                    const bool has2 = cur != 0;
                    const int i2 = has2 ? __ffs(cur) - 1 : i1;
                    cur &= cur - 1;                                        // 0 stays 0
                    const int jy1 = (i1 >> 3) + half, t1 = i1 & 7, jy2 = (i2 >> 3) + half, t2 = i2 & 7;
                    const float dx1 = fxm - (float)t1, dy1 = fym - (float)jy1, dx2 = fxm - (float)t2, dy2 = fym - (float)jy2;
                    const float d21 = fmaf(dx1, dx1, fmaf(dy1, dy1, dz2f)), d22 = fmaf(dx2, dx2, fmaf(dy2, dy2, dz2f));
                    bool in1 = d21 < E.rho2, in2 = has2 && d22 < E.rho2;
                    if (fabsf(d21 - E.rho2) < 4.0e-5f) in1 = r3_exact_in(E, q[0], q[1], q[2], xa + t1, ya + jy1, za + jzc);
                    if (has2 && fabsf(d22 - E.rho2) < 4.0e-5f) in2 = r3_exact_in(E, q[0], q[1], q[2], xa + t2, ya + jy2, za + jzc);
                    const int w1 = in1 ? 1 : 0, w2 = in2 ? 1 : 0;
                    const int x1 = w1 * t1, x2 = w2 * t2, y1 = w1 * jy1, y2 = w2 * jy2, nz = w1 + w2;
                    An += nz; Asx += x1 + x2; Asy += y1 + y2; Asz += nz * jzc;
                    Asxx += x1 * t1 + x2 * t2; Asxy += x1 * jy1 + x2 * jy2; Asxz += (x1 + x2) * jzc;
                    Asyy += y1 * jy1 + y2 * jy2; Asyz += (y1 + y2) * jzc; Aszz += nz * jzc * jzc;
                }
            }
#else
            {
                uint32_t cur = 0;
                int k = 0, jzc = 0, half = 0;                               // half: first row (8-bit rows) / first bit (7-bit rows) of the word
                float dz2f = 0.0f;
                for (;;) {
                    if (cur == 0) {
                        if (k >= n_u) break;
                        const uint2 e = ulist[k++];
                        cur = e.x;
                        jzc = (int)(e.y & 255u);
                        half = (int)(e.y >> 8) * (R3_ROW_BITS == 8 ? 4 : 32);
                        const float dzf = fzm - (float)jzc;
                        dz2f = dzf * dzf;
                    }
#if R3_ROW_BITS == 8
                    const int i = __ffs(cur) - 1;
                    cur &= cur - 1;
                    const int jy = (i >> 3) + half, t = i & 7;
#else
                    const int i = __ffs(cur) - 1 + half;
                    cur &= cur - 1;
                    const int jy = (i * 37) >> 8, t = i - 7 * jy;
#endif
                    const float dx = fxm - (float)t, dy = fym - (float)jy;
                    const float d2 = fmaf(dx, dx, fmaf(dy, dy, dz2f));
                    bool in = d2 < E.rho2;
                    if (fabsf(d2 - E.rho2) < 4.0e-5f) in = r3_exact_in(E, q[0], q[1], q[2], xa + t, ya + jy, za + jzc);
                    if (in) {
                        An += 1; Asx += t; Asy += jy; Asz += jzc;
                        Asxx += t * t; Asxy += t * jy; Asxz += t * jzc;
                        Asyy += jy * jy; Asyz += jy * jzc; Aszz += jzc * jzc;
                    }
                }
            }
#endif
            }
            const float exm = fxm, eym = fym, ezm = fzm;
#if R3_PIPELINE
            if (li + 1 < P.n) stage_entry(li + 1);              // window and table lines of entry li are consumed
#endif
            if (active && li >= 0)
                emit_features_window<OutT>(An, Asx, Asy, Asz, Asxx, Asxy, Asxz, Asyy, Asyz, Aszz, exm, eym, ezm, true,
                                           E.edge, dst_row + E.col, EXT ? NBR_DESC_EXTENDED : 0);
        }
#if R3_TMA_ROWS
        if (stage_rows) {
            // row buffer -> global with the bulk-copy engine (TMA, cp.async.bulk shared -> global): every lane hands over
            // its own finished row as ONE asynchronous copy of row_bytes (a multiple of 16), instead of the warp moving
            // 16-byte pieces through registers.  the generic-proxy writes of the row are fenced for the async proxy first;
            // the buffer is reused only after the copies have read it (wait_group.read at the top of the next group)
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (active) {
                const uint32_t src_addr = (uint32_t)__cvta_generic_to_shared(rows + lane * row_bytes);
                unsigned char *gdst = reinterpret_cast<unsigned char *>(out) + qi * (long long)row_bytes;
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(src_addr), "r"(row_bytes) : "memory");
            }
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
#else
        if (stage_rows) {
            // row buffer -> global: consecutive lanes write consecutive 16-byte pieces of a row
#if R3_TMA_PEER
            if (MULTI) {
                // the warp's rows are one contiguous piece of shared memory and one contiguous piece of every peer's
                // staging buffer: lane d hands them to the bulk-copy engine for peer d (the right granularity for it:
                // 2.5 KB per copy, not one 80-byte row).  generic-proxy writes are fenced for the async proxy first
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                const long long left = (long long)nq - (long long)grp * 32, n_rows = left < 32 ? left : 32;
                if (lane < D.n && lane != D.self) {
                    const uint32_t src_addr = (uint32_t)__cvta_generic_to_shared(rows);
                    unsigned char *gdst = D.base[lane] + (D.row_offset + grp * 32) * (long long)row_bytes;
                    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(src_addr), "r"((uint32_t)(n_rows * row_bytes)) : "memory");
                }
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
#endif
            __syncwarp();
            // piece gp = lane + 32 t is piece cidx of row r; (r, cidx) advance by (32 / cpr, 32 % cpr) per trip
            const int pieces = 32 * cpr;
            int r = r_first, cidx = c_first;
            for (int gp = lane; gp < pieces; gp += 32) {
                const long long row_q = __shfl_sync(0xffffffffu, (long long)qi, r);
                const bool row_active = grp * 32 + r < nq;
                if (row_active) {
                    const uint4 v = *reinterpret_cast<const uint4 *>(rows + r * row_bytes + cidx * 16);
                    // streaming store: the rows are never read again, they should not push bricks and tables out of L2
                    __stcs(reinterpret_cast<uint4 *>(reinterpret_cast<unsigned char *>(out) + row_q * (long long)row_bytes + cidx * 16), v);
                    if (MULTI && !R3_TMA_PEER) {
                        const long long at = (D.row_offset + grp * 32) * (long long)row_bytes + (long long)gp * 16;
                        for (int d = 0; d < D.n; ++d)
                            if (d != D.self) __stcs(reinterpret_cast<uint4 *>(D.base[d] + at), v);
                    }
                }
                r += r_step; cidx += c_step;
                if (cidx >= cpr) { cidx -= cpr; r += 1; }
            }
            if (MULTI && active)
                for (int d = 0; d < D.n; ++d)
                    if (d != D.self) D.perm[d][D.row_offset + slot_i] = (uint32_t)qi;
            __syncwarp();
        }
#endif
    }
#if R3_TMA_ROWS
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");          // the last rows have left before the block retires
#elif R3_TMA_PEER
    if (MULTI) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
#endif
}

// fills one entry; false if this (lattice, radius) is not a 7x7x7 window or tables are disabled
bool rows3_entry(const Lattice *lat, double radius, int col, const R3Entry *prev, R3Entry *E, int *tq_io,
                 cudaStream_t stream, int *rc)
{
    *rc = NBR_OK;
    static const bool disabled = getenv("NBR_NO_ROWS3") != nullptr || getenv("NBR_NO_BALL_TABLE") != nullptr;
    if (disabled) return false;
    const double e = lat->grid.edge;
    const double rho = radius / e;
    if (!(rho + 0.5 + 1e-6 < 4.0)) return false;
    static const int q_env = getenv("NBR_BALL_Q") ? atoi(getenv("NBR_BALL_Q")) : 0;
    const int tq = q_env >= 1 && q_env <= 64 ? q_env : 32;
    *tq_io = tq;
    const LatticeDev d = lat->dev();
    for (int a = 0; a < 3; ++a) { E->minc[a] = d.g.minc[a]; E->cell_lo[a] = d.g.cell_lo[a]; }
    E->edge = d.g.edge; E->inv_edge = d.g.inv_edge; E->r = radius;
    E->dir = d.dir; E->pool = d.pool;
    E->nbx = d.nbx; E->nby = d.nby; E->nbz = d.nbz;
    E->rho2 = (float)(rho * rho);
    E->col = col;
    E->reuse = prev && prev->pool == d.pool && prev->dir == d.dir;
    // margin (squared distance, cell units): the reference expression and the kernel's f are each off by a
    // few ulp of the largest coordinate magnitude, times 2*(W+1) cells
    double maxabs = 0.0;
    for (int a = 0; a < 3; ++a)
        maxabs = std::max(maxabs, std::max(fabs(lat->grid.min_corner[a]), fabs(lat->grid.max_corner[a])));
    const double margin = std::max(1e-11, 64.0 * 2.3e-16 * (maxabs / e + 8.0));
    *rc = ball_table_get(rho * rho, margin, tq, R3_ROW_BITS, &E->table, stream);
    return *rc == NBR_OK;
}

// true if a launch of n_entries entries assembles whole rows in shared memory (the condition of the MULTI write-out)
bool rows3_owns_rows(int n_entries, int64_t row_stride, int out_dtype, int descriptor_mask)
{
    static const bool no_stage = getenv("NBR_NO_ROW_STAGING") != nullptr;
    const bool ext = (descriptor_mask & NBR_DESC_EXTENDED) != 0;
    const int ncol = ext ? NBR_COLS_EXTENDED : NBR_COLS_REFERENCE;
    const size_t row_bytes = (size_t)row_stride * (out_dtype == NBR_F32 ? 4 : 8);
    return !no_stage && !ext && n_entries > 0 && n_entries <= R3_MAX_ENTRIES && (int64_t)n_entries * ncol == row_stride &&
           row_bytes <= R3_MAX_ROW_BYTES && row_bytes % 16 == 0;
}

int rows3_launch(const R3Launch *L, const void *query, int dtype, const uint32_t *perm, int64_t nq, void *out,
                 int out_dtype, int64_t row_stride, int descriptor_mask, cudaStream_t stream, const RowDests *dests)
{
    if (nq <= 0 || L->n <= 0) return NBR_OK;
    R3Launch copy = *L;
    Scratch stats;
    const bool want_stats = getenv("NBR_ROW_STATS") != nullptr;
    copy.stats = nullptr;
    if (want_stats) {
        NBR_TRY(stats.alloc(sizeof(unsigned long long) * 8 * R3_MAX_ENTRIES, stream));
        NBR_CUDA(cudaMemsetAsync(stats.ptr, 0, sizeof(unsigned long long) * 8 * R3_MAX_ENTRIES, stream));
        copy.stats = stats.as<unsigned long long>();
    }
    const int blocks = (int)std::min<int64_t>(ceil_div(ceil_div(nq, 32), R3_WARPS), (int64_t)device_sm_count() * 16);
    const bool ext = (descriptor_mask & NBR_DESC_EXTENDED) != 0;
    // whole rows are assembled in shared memory when this launch writes every column of the rows
    const int ncol = ext ? NBR_COLS_EXTENDED : NBR_COLS_REFERENCE;
    const size_t row_bytes = (size_t)row_stride * (out_dtype == NBR_F32 ? 4 : 8);
    static const bool no_stage = getenv("NBR_NO_ROW_STAGING") != nullptr;
    const int stage_rows = !no_stage && (int64_t)L->n * ncol == row_stride && row_bytes <= R3_MAX_ROW_BYTES && row_bytes % 16 == 0 &&
                           ((uintptr_t)out & 15) == 0;
    if (dests && (!stage_rows || ext)) return fail(NBR_ERR_UNSUPPORTED, "rows3_launch: rows for several destinations need a launch that owns whole rows");
    RowDests D;
    memset(&D, 0, sizeof(D));
    if (dests) {
        D = *dests;
        for (int d = 0; d < D.n; ++d)
            if (d != D.self && (((uintptr_t)D.base[d] & 15) != 0 || !D.base[d] || !D.perm[d]))
                return fail(NBR_ERR_INVALID, "rows3_launch: destination rows must be 16-byte aligned");
    }
    const size_t smem = (size_t)R3_WARPS * (R3_WIN_BYTES + R3_TAB_BYTES + (stage_rows ? 32 * row_bytes : 0));
    static std::atomic<uint64_t> configured{0};
    if (first_use_on_device(configured)) {
        const int max_smem = R3_WARPS * (R3_WIN_BYTES + R3_TAB_BYTES + 32 * R3_MAX_ROW_BYTES);
        NBR_CUDA(cudaFuncSetAttribute(rows3_kernel<float, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
        NBR_CUDA(cudaFuncSetAttribute(rows3_kernel<float, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
        NBR_CUDA(cudaFuncSetAttribute(rows3_kernel<double, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
        NBR_CUDA(cudaFuncSetAttribute(rows3_kernel<double, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
        NBR_CUDA(cudaFuncSetAttribute(rows3_kernel<float, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
        NBR_CUDA(cudaFuncSetAttribute(rows3_kernel<double, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
    }
#define R3_GO(T, X, M) rows3_kernel<T, X, M><<<blocks, R3_WARPS * 32, smem, stream>>>(copy, query, dtype, perm, nq, (T *)out, row_stride, stage_rows, D)
    if (dests) { if (out_dtype == NBR_F32) R3_GO(float, false, true); else R3_GO(double, false, true); }
    else if (out_dtype == NBR_F32) { if (ext) R3_GO(float, true, false); else R3_GO(float, false, false); }
    else                           { if (ext) R3_GO(double, true, false); else R3_GO(double, false, false); }
#undef R3_GO
    NBR_LAUNCHED();
    if (want_stats) {
        unsigned long long h[8 * R3_MAX_ENTRIES];
        NBR_CUDA(cudaMemcpyAsync(h, stats.ptr, sizeof(h), cudaMemcpyDeviceToHost, stream));
        NBR_CUDA(cudaStreamSynchronize(stream));
        for (int l = 0; l < L->n; ++l)
            fprintf(stderr, "[nbr rows3 stats] entry %d edge %.3g r %.3g: staged warps %llu, direct warps %llu, distinct anchor cells per warp %.2f, warps with <= 8: %.1f %%\n", l,
                    L->e[l].edge, L->e[l].r, h[l * 8], h[l * 8 + 1], (double)h[l * 8 + 2] / (double)(h[l * 8] + h[l * 8 + 1]),
                    100.0 * (double)h[l * 8 + 3] / (double)(h[l * 8] + h[l * 8 + 1]));
    }
    return NBR_OK;
}

}  // namespace nbr

extern "C" int64_t nbr_debug_bounds_violations(void)
{
    unsigned long long v = 0;
    if (cudaMemcpyFromSymbol(&v, nbr::g_r3_violations, sizeof(v)) != cudaSuccess) return -1;
    return (int64_t)v;
}
