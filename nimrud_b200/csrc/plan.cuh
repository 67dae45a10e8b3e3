// plan.cuh -- the whole-path driver's state (api.cu): what depends on the SEARCH cloud only (one lattice per distinct
// edge), and the calls that run batches of queries against it.
#pragma once
#include <vector>

#include "common.cuh"
#include "lattice.cuh"

namespace nbr {

struct Mailbox;

struct Plan {
    struct Group { double edge; Lattice *lat; std::vector<int> scales; };
    std::vector<Group> groups;
    std::vector<double> edges, radii;
    int n_scales = 0, ncol = 4, descriptor_mask = 0;
    double finest = 0.0;
    double local_box[6];
    double order_origin[3];   // brick corner of the finest lattice: the query order is aligned with it
    ~Plan() { for (auto &g : groups) delete g.lat; }
};

int plan_create(Plan **out, const void *search, int s_dtype, int64_t ns, const double *edges, const double *radii,
                int n_scales, int descriptor_mask, const double *global_lohi, const double *known_local_box,
                cudaStream_t stream, const void *search2 = nullptr, int64_t ns2 = 0, Mailbox *mailbox = nullptr,
                const CellOrderInfo *order = nullptr);
// features of one batch of queries in arbitrary order -> rows [0, nq) of `out` (device)
int plan_run(const Plan *P, const void *query, int q_dtype, int64_t nq, const double *qbox_known, void *out,
             int out_dtype, cudaStream_t stream);
int plan_voxel_counts(const Plan *P, int64_t *n_voxels_host);

}  // namespace nbr
