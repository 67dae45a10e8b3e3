// ball_table.cu -- "shell tables" for the fused radius kernel (W = 3 windows, r/e < 3.5).
//
// a query sits at fractional position f in [0,1)^3 of its anchor cell; the 7x7x7 cells of its window
// are at integer offsets.  which of them lie in the ball of radius rho = r/e depends on f only, and for
// a small BIN of f (1/Q of a cell per axis) almost every cell is either inside for the whole bin or
// outside for the whole bin.  per (rho, bin) the table stores, per z-slab of the window, two masks of 49
// cells (bit row_bits*jy + t <-> cell (t, jy); row_bits = 8 puts every row in its own byte, rows 0..3 in
// the low word, 4..6 in the high word; 7 packs them):
//     in   cells that are inside the ball for EVERY f of the bin (with a safety margin)
//     unc  cells that may be on either side -- the kernel evaluates those, and only those that are
//          occupied, with the reference's own float64 expression (nimrud/minimal/multiscale.py:103,
//          scipy's inclusive sum((q-v)^2) <= r^2), so neighbor sets stay bit-exact by construction.
// layout: table[bin][8] uint4 = {in_lo, in_hi, unc_lo, unc_hi}, slab 7 is padding (one 128-byte line
// per bin).  tables are built on the device once per (device, rho^2, margin) and cached for the life of
// the process.
#include <map>
#include <mutex>
#include <tuple>

#include "common.cuh"
#include "radius_rows.cuh"

namespace nbr {

__global__ void __launch_bounds__(256)
ball_table_kernel(uint4 *__restrict__ table, int Q, double rho2, double margin, int row_bits)
{
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= Q * Q * Q * 8) return;
    const int jz = idx & 7, bin = idx >> 3;
    if (jz == 7) { table[idx] = make_uint4(0, 0, 0, 0); return; }
    const int b[3] = {bin % Q, (bin / Q) % Q, bin / (Q * Q)};
    double lo[3], hi[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        lo[a] = (double)b[a] / Q - 0.5 + 3.0;       // window units: cell t has its centre at t
        hi[a] = (double)(b[a] + 1) / Q - 0.5 + 3.0;
    }
    auto axis = [&](int a, int t, double &dmin, double &dmax) {
        const double x = (double)t;
        dmin = fmax(0.0, fmax(lo[a] - x, x - hi[a]));
        dmax = fmax(fabs(lo[a] - x), fabs(hi[a] - x));
    };
    double zmin, zmax;
    axis(2, jz, zmin, zmax);
    unsigned long long in = 0, unc = 0;
    for (int jy = 0; jy < 7; ++jy) {
        double ymin, ymax;
        axis(1, jy, ymin, ymax);
        for (int t = 0; t < 7; ++t) {
            double xmin, xmax;
            axis(0, t, xmin, xmax);
            const double d2min = xmin * xmin + ymin * ymin + zmin * zmin;
            const double d2max = xmax * xmax + ymax * ymax + zmax * zmax;
            const unsigned long long bit = 1ull << (row_bits * jy + t);
            if (d2max <= rho2 - margin) in |= bit;
            else if (!(d2min > rho2 + margin)) unc |= bit;
        }
    }
    table[idx] = make_uint4((uint32_t)in, (uint32_t)(in >> 32), (uint32_t)unc, (uint32_t)(unc >> 32));
}

static std::mutex g_table_mutex;
static std::map<std::tuple<int, int, int, uint64_t, uint64_t>, const uint4 *> g_tables;

// margin: absolute slack on squared distances in cell units (covers the rounding of the reference
// expression and of the kernel's own f); rounded up to a power of two so that the cache key is stable
int ball_table_get(double rho2, double margin, int Q, int row_bits, const uint4 **out, cudaStream_t stream)
{
    int dev = 0;
    NBR_CUDA(cudaGetDevice(&dev));
    int ex = 0;
    frexp(margin, &ex);
    margin = ldexp(1.0, ex);
    uint64_t kr, km;
    memcpy(&kr, &rho2, 8);
    memcpy(&km, &margin, 8);
    const auto key = std::make_tuple(dev, Q, row_bits, kr, km);
    std::lock_guard<std::mutex> lock(g_table_mutex);
    auto it = g_tables.find(key);
    if (it != g_tables.end()) { *out = it->second; return NBR_OK; }
    uint4 *t = nullptr;
    const size_t n = (size_t)Q * Q * Q * 8;
    NBR_CUDA(cudaMalloc(&t, n * sizeof(uint4)));
    ball_table_kernel<<<(unsigned)ceil_div((int64_t)n, 256), 256, 0, stream>>>(t, Q, rho2, margin, row_bits);
    NBR_LAUNCHED();
    NBR_CUDA(cudaStreamSynchronize(stream));       // once per distinct (rho, margin): later callers may be on other streams
    g_tables[key] = t;
    *out = t;
    return NBR_OK;
}

// called at the start of a feature call (never between building a launch description and launching it): a
// process that keeps asking for new r/e ratios drops every table once no kernel can be reading them
int ball_tables_trim()
{
    // only the CURRENT device's tables, and only after that device has drained: a table of another device may be in
    // use by a kernel this thread cannot wait for.  (a process that drives one device from several host threads
    // must not interleave feature calls with more than 128 distinct r/e ratios in flight)
    int dev = 0;
    NBR_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lock(g_table_mutex);
    size_t mine = 0;
    for (auto &kv : g_tables) mine += std::get<0>(kv.first) == dev;
    if (mine < 128) return NBR_OK;
    NBR_CUDA(cudaDeviceSynchronize());
    for (auto it = g_tables.begin(); it != g_tables.end();) {
        if (std::get<0>(it->first) == dev) { cudaFree(const_cast<uint4 *>(it->second)); it = g_tables.erase(it); }
        else ++it;
    }
    return NBR_OK;
}

}  // namespace nbr
