// radius_rows.cu -- row-interval radius feature kernel (placeholder until the first GPU parity run).
#include "common.cuh"
#include "lattice.cuh"

namespace nbr {

int radius_features_rows(const Lattice *, const void *, int, int64_t, const double *, int, void *, int, int64_t, int,
                         int, cudaStream_t, bool *handled)
{
    *handled = false;
    return NBR_OK;
}

}  // namespace nbr
