// radius_rows.cuh -- launch description of the fused row-interval feature kernel.
#pragma once
#include "common.cuh"
#include "lattice.cuh"

namespace nbr {

constexpr int RW_WARPS = 4;
constexpr int RW_CAP_BRICKS = 48;                 // staged bricks per warp (6 KB)
constexpr int RW_MAX_W = 9;                       // a single query's window (2*4*6 bricks) must fit the buffer
constexpr int RW_MAX_RADII = 8;                   // radii per lattice in one launch
constexpr int RW_MAX_LATTICES = 8;                // lattices per launch

struct RowsParam {
    double r[RW_MAX_RADII];      // exact radii (float64 test)
    float rho2[RW_MAX_RADII];    // (r/e)^2
    float eps_a[RW_MAX_RADII];   // rounding bound: delta = eps_a * min(rsqrt(T), 1e3) + eps_b
    float t_min[RW_MAX_RADII];   // rows with T < t_min are certainly empty
    int w[RW_MAX_RADII];         // window half-width of each radius
    int col[RW_MAX_RADII];       // first output column of each radius
    float eps_b;
    int n;
    int wmax;
    int pad;
};

struct RowsLaunch {
    LatticeDev lat[RW_MAX_LATTICES];
    RowsParam rows[RW_MAX_LATTICES];
    unsigned long long *stats;   // optional diagnostics (NBR_ROW_STATS=1): per lattice 8 counters
    int n_lat;
    int pad;
};

// ---- lean kernel for 7x7x7 windows (rows3.cu): one entry per (lattice, radius), passed in kernel parameters
constexpr int R3_MAX_ENTRIES = 16;
struct R3Entry {
    double minc[3];
    double edge, inv_edge, r;
    const uint32_t *dir, *pool;
    const uint4 *table;          // shell table of r/e (ball_table.cu)
    int32_t cell_lo[3];
    int32_t nbx, nby, nbz;
    float rho2;
    int32_t col;                 // first output column
    int32_t reuse;               // same lattice as the previous entry: its window is still staged
    int32_t pad;
};
struct R3Launch {
    R3Entry e[R3_MAX_ENTRIES];
    unsigned long long *stats;
    int32_t n;
    int32_t tq;                  // bins per axis of the shell tables
};
bool rows3_entry(const Lattice *lat, double radius, int col, const R3Entry *prev, R3Entry *E, int *tq_io,
                 cudaStream_t stream, int *rc);
struct RowDests;
int rows3_launch(const R3Launch *L, const void *query, int dtype, const uint32_t *perm, int64_t nq, void *out,
                 int out_dtype, int64_t row_stride, int descriptor_mask, cudaStream_t stream, const RowDests *dests = nullptr);
bool rows3_owns_rows(int n_entries, int64_t row_stride, int out_dtype, int descriptor_mask);

// same for 11x11x11 windows (rows5.cu): r/e + 0.5 < 6
bool rows5_entry(const Lattice *lat, double radius, int col, const R3Entry *prev, R3Entry *E, int *tq_io,
                 cudaStream_t stream, int *rc);
int rows5_launch(const R3Launch *L, const void *query, int dtype, const uint32_t *perm, int64_t nq, void *out,
                 int out_dtype, int64_t row_stride, int descriptor_mask, cudaStream_t stream);

int rows_param(const Lattice *lat, const double *radii, const int *cols, int nr, RowsParam *P, cudaStream_t stream);
int ball_table_get(double rho2, double margin, int Q, int row_bits, const uint4 **out, cudaStream_t stream);
int ball_tables_trim();      // bound the table caches; only at the start of a feature call
int ball_tables5_trim();
bool rows_supported(double edge, const double *radii, int nr);
int radius_rows_launch(const RowsLaunch *launch_host, const void *query, int dtype, const uint32_t *perm, int64_t nq,
                       void *out, int out_dtype, int64_t row_stride, int descriptor_mask, cudaStream_t stream);

}  // namespace nbr
