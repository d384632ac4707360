// Shared-memory / TMA-staged SetConv for the induced -> target direction (regular key grid).
// Placeholder: reports NPF_ENOTSUP so callers use the generic kernels in setconv.cu.
#include "common.cuh"

namespace npf {

int setconv_tile_fwd(const float*, long, const float*, long, const float*, const float*, float*, float*, float*, int,
                     int, int, int, cudaStream_t) { return NPF_ENOTSUP; }
int setconv_tile_bwd(const float*, long, const float*, long, const float*, const float*, const float*, const float*,
                     const float*, const float*, float*, float*, int, int, int, int, cudaStream_t) { return NPF_ENOTSUP; }

}  // namespace npf
