// Parameter blocks of the thin (<= 8-wide) linear kernels, shared by gemm.cu (dispatch) and gemm_thin.cu (kernels).
#pragma once
#include <cuda_runtime.h>

namespace npf {

struct ThinRedParams {
    const float* A; long lda; int relu_a;
    const float* B; long sb_o, sb_r;
    const float* bias;
    const float* u; const float* w2; long ldw2;
    const float* mask; long ldm;
    float* out; long ldo;
    long M; int R, O;
    int relu_out, accum;
};
struct RowDotParams {
    const float* A; long lda; int relu_a;
    const float* B; long sb_j, sb_i;
    const float* bias;
    const float* mask; long ldm;
    float* out; long ldo;
    long M; int I, J;
    int relu_out, accum;
};
struct ThinOuterParams {
    const float* S; long lds; int relu_s;
    const float* T; long ldt; int relu_t;
    float* out; long so_j, so_c;
    float* out_ones;
    const float* u; float* out_u; long so_u;
    long M; int J, C;
    long rows_per_block;
};
int thin_red(ThinRedParams& p, cudaStream_t st);
int rowdot(RowDotParams& p, cudaStream_t st);
int thin_outer(ThinOuterParams& p, cudaStream_t st);
constexpr int kThinMax = 8;

}  // namespace npf
