// Tensor-core (tcgen05 / TMEM) path for the 128-wide linear layers.  Placeholder until the UMMA kernel
// lands: every entry reports NPF_ENOTSUP so callers fall back to the fp32 FFMA kernel in gemm.cu.
#include "common.cuh"

namespace npf {

int linear_fwd_tc(const float*, int, const float*, int, const float*, float*, int, int, int, int, int, const float*,
                  const float*, int, int, cudaStream_t) { return NPF_ENOTSUP; }
int linear_bwd_data_tc(const float*, int, const float*, int, float*, int, int, int, int, const float*, int, int, int,
                       cudaStream_t) { return NPF_ENOTSUP; }
int linear_bwd_weight_tc(const float*, int, const float*, int, float*, int, int, int, int, int, int, cudaStream_t) {
    return NPF_ENOTSUP;
}

}  // namespace npf
