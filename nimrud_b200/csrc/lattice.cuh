// lattice.cuh -- host-side handle of one voxel-filtered search cloud (one edge length).
#pragma once
#include <memory>

#include "common.cuh"

namespace nbr {

struct Lattice {
    cudaStream_t stream = nullptr;
    nbr_grid grid;
    GridDev gdev;
    int64_t n_search = 0;
    int32_t nbx = 0, nby = 0, nbz = 0;
    int64_t n_dir = 0;
    int64_t pool_slots = 0;        // allocated slots (upper bound), slot 0 = empty brick
    uint32_t *dir = nullptr;
    uint32_t *pool = nullptr;
    uint32_t *rowbase = nullptr;   // INDEXED only
    uint64_t *ukeys = nullptr;     // INDEXED only: sorted unique addresses (np.unique order)
    unsigned char *counters = nullptr;   // device: [0] u32 n_bricks, [8] u64 n_voxels, [16] i64 n_unique
    bool indexed = false;
    // lattices of one batch (lattices_create_batch) share one directory / pool / counter allocation; the
    // pointers above then point into it and the last lattice of the batch releases it
    struct SharedBuffers {
        cudaStream_t stream = nullptr;
        void *dir = nullptr, *pool = nullptr, *counters = nullptr;
        ~SharedBuffers();
    };
    std::shared_ptr<SharedBuffers> shared;

    ~Lattice();
    LatticeDev dev() const;
    const uint32_t *n_bricks_dev() const { return reinterpret_cast<const uint32_t *>(counters); }
    const int64_t *n_unique_dev() const { return reinterpret_cast<const int64_t *>(counters + 16); }
};

int lattice_create(Lattice **out, const void *xyz, int dtype, int64_t n, const nbr_grid *grid, int flags,
                   cudaStream_t stream, const double *local_lohi = nullptr);
constexpr int LATTICE_BATCH = 8;
// all lattices of `grids` over the same search cloud in one pass over the points (not INDEXED)
// xyz2 / n2: optional second part of the search cloud (multi-GPU: the halo points next to the ordered tile)
struct Mailbox;
int lattices_create_batch(Lattice **out, int n_lat, const nbr_grid *grids, const void *xyz, int dtype, int64_t n,
                          cudaStream_t stream, const double *local_lohi = nullptr, const void *xyz2 = nullptr,
                          int64_t n2 = 0, Mailbox *mailbox = nullptr, const CellOrderInfo *order = nullptr);
int lattice_counts(const Lattice *L, int64_t *n_voxels, int64_t *n_bricks);
int bbox(const void *xyz, int dtype, int64_t n, int ndim, double *lohi_dev, cudaStream_t stream);
int grid_from_bbox(const double lo[3], const double hi[3], double edge, int ndim, nbr_grid *out);
int grid_to_dev(const nbr_grid *g, GridDev *d, const double *local_lohi);

}  // namespace nbr
