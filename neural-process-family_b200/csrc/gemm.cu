// fp32 GEMM family behind npf_linear_{fwd,bwd_data,bwd_weight}: one 128x128x16 register-tiled FFMA kernel,
// instantiated for the three operand orientations.  This is the NPF_PREC_FP32 (1e-4 parity) path; the
// tensor-core (tcgen05) path for 128-wide layers lives in gemm_tc.cu.
//
//   C[P,Q] (+)= sum_r A(p,r) * B(r,q)
//     fwd        (NT): P=M rows, Q=N, R=K   A = X  (reduction-contiguous), B = W   (reduction-contiguous)
//     bwd data   (NN): P=M rows, Q=K, R=N   A = dY (reduction-contiguous), B = W   (output-contiguous)
//     bwd weight (TN): P=N,      Q=K, R=M   A = dY (output-contiguous),    B = X   (output-contiguous), split-R
#include <stdarg.h>

#include "common.cuh"
#include "gemm_thin.cuh"

namespace npf {

constexpr int BM = 128, BN = 128, BK = 16, PAD = 4;

struct GemmParams {
    const float* A; long lda;
    const float* B; long ldb;
    float* C; long ldc;
    int P, Q, R;
    int r_per_split;
    int relu_a, relu_b;
    const float* bias_q;
    const float* u_p; const float* w_q; long ldw2;
    const float* mask; long ldm;
    int relu_out, accum, atomic;
    int a_vec, b_vec, c_vec;
};

// Load one [128 x 16] operand tile into 8 registers per thread.
// RC (reduction-contiguous): element (o, r) at ptr[o*ld + r];  OC (output-contiguous): ptr[r*ld + o].
template <bool RC>
__device__ __forceinline__ void load_tile(const float* __restrict__ ptr, long ld, int o0, int O, int r0, int r_end,
                                          int vec, int relu, float (&reg)[8]) {
    const int t = threadIdx.x;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const int idx = t + i * 256;
        int o, r;
        if (RC) { o = o0 + (idx >> 2); r = r0 + ((idx & 3) << 2); }
        else    { r = r0 + (idx >> 5); o = o0 + ((idx & 31) << 2); }
        float v0 = 0.f, v1 = 0.f, v2 = 0.f, v3 = 0.f;
        if (RC) {
            if (o < O) {
                const float* g = ptr + (long)o * ld + r;
                if (vec && r + 4 <= r_end) {
                    float4 v = __ldg(reinterpret_cast<const float4*>(g));
                    v0 = v.x; v1 = v.y; v2 = v.z; v3 = v.w;
                } else {
                    if (r + 0 < r_end) v0 = __ldg(g + 0);
                    if (r + 1 < r_end) v1 = __ldg(g + 1);
                    if (r + 2 < r_end) v2 = __ldg(g + 2);
                    if (r + 3 < r_end) v3 = __ldg(g + 3);
                }
            }
        } else {
            if (r < r_end) {
                const float* g = ptr + (long)r * ld + o;
                if (vec && o + 4 <= O) {
                    float4 v = __ldg(reinterpret_cast<const float4*>(g));
                    v0 = v.x; v1 = v.y; v2 = v.z; v3 = v.w;
                } else {
                    if (o + 0 < O) v0 = __ldg(g + 0);
                    if (o + 1 < O) v1 = __ldg(g + 1);
                    if (o + 2 < O) v2 = __ldg(g + 2);
                    if (o + 3 < O) v3 = __ldg(g + 3);
                }
            }
        }
        if (relu) { v0 = fmaxf(v0, 0.f); v1 = fmaxf(v1, 0.f); v2 = fmaxf(v2, 0.f); v3 = fmaxf(v3, 0.f); }
        reg[i * 4 + 0] = v0; reg[i * 4 + 1] = v1; reg[i * 4 + 2] = v2; reg[i * 4 + 3] = v3;
    }
}

// Store the 8 registers into the [BK][128+PAD] shared tile (reduction-major).
template <bool RC>
__device__ __forceinline__ void store_tile(float (*S)[BM + PAD], const float (&reg)[8]) {
    const int t = threadIdx.x;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const int idx = t + i * 256;
        if (RC) {
            const int o = idx >> 2, r = (idx & 3) << 2;
            S[r + 0][o] = reg[i * 4 + 0];
            S[r + 1][o] = reg[i * 4 + 1];
            S[r + 2][o] = reg[i * 4 + 2];
            S[r + 3][o] = reg[i * 4 + 3];
        } else {
            const int r = idx >> 5, o = (idx & 31) << 2;
            *reinterpret_cast<float4*>(&S[r][o]) = make_float4(reg[i * 4 + 0], reg[i * 4 + 1], reg[i * 4 + 2], reg[i * 4 + 3]);
        }
    }
}

template <bool A_RC, bool B_RC>
__global__ void __launch_bounds__(256) gemm_f32_kernel(GemmParams p) {
    __shared__ __align__(16) float As[2][BK][BM + PAD];
    __shared__ __align__(16) float Bs[2][BK][BN + PAD];

    const int q0 = blockIdx.x * BN;
    const int p0 = blockIdx.y * BM;
    const int r_begin = blockIdx.z * p.r_per_split;
    const int r_end = min(p.R, r_begin + p.r_per_split);
    const int n_iter = (r_end - r_begin + BK - 1) / BK;

    const int t = threadIdx.x;
    const int ty = t >> 4, tx = t & 15;

    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

    float ra[8], rb[8];
    if (n_iter > 0) {
        load_tile<A_RC>(p.A, p.lda, p0, p.P, r_begin, r_end, p.a_vec, p.relu_a, ra);
        load_tile<B_RC>(p.B, p.ldb, q0, p.Q, r_begin, r_end, p.b_vec, p.relu_b, rb);
        store_tile<A_RC>(As[0], ra);
        store_tile<B_RC>(Bs[0], rb);
    }
    __syncthreads();

    int buf = 0;
    for (int it = 0; it < n_iter; ++it) {
        const bool more = (it + 1 < n_iter);
        if (more) {
            const int r0 = r_begin + (it + 1) * BK;
            load_tile<A_RC>(p.A, p.lda, p0, p.P, r0, r_end, p.a_vec, p.relu_a, ra);
            load_tile<B_RC>(p.B, p.ldb, q0, p.Q, r0, r_end, p.b_vec, p.relu_b, rb);
        }
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 4]);
            const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][kk][64 + ty * 4]);
            const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][kk][tx * 4]);
            const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][kk][64 + tx * 4]);
            const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        if (more) {
            store_tile<A_RC>(As[buf ^ 1], ra);
            store_tile<B_RC>(Bs[buf ^ 1], rb);
        }
        __syncthreads();
        buf ^= 1;
    }

    // epilogue
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int pr = p0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
        if (pr >= p.P) continue;
        const float up = p.u_p ? __ldg(p.u_p + pr) : 0.f;
#pragma unroll
        for (int jg = 0; jg < 2; ++jg) {
            const int qc = q0 + jg * 64 + tx * 4;
            if (qc >= p.Q) continue;
            float v[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float x = acc[i][jg * 4 + j];
                const int q = qc + j;
                if (q < p.Q) {
                    if (p.bias_q) x += __ldg(p.bias_q + q);
                    if (p.u_p) x = fmaf(up, __ldg(p.w_q + (long)q * p.ldw2), x);
                    if (p.relu_out) x = fmaxf(x, 0.f);
                    if (p.mask) x = (__ldg(p.mask + (long)pr * p.ldm + q) > 0.f) ? x : 0.f;
                }
                v[j] = x;
            }
            float* c = p.C + (long)pr * p.ldc + qc;
            if (p.atomic) {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (qc + j < p.Q) atomicAdd(c + j, v[j]);
            } else if (p.c_vec && qc + 4 <= p.Q) {
                float4 o = make_float4(v[0], v[1], v[2], v[3]);
                if (p.accum) {
                    const float4 old = *reinterpret_cast<const float4*>(c);
                    o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w;
                }
                *reinterpret_cast<float4*>(c) = o;
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (qc + j < p.Q) c[j] = p.accum ? c[j] + v[j] : v[j];
            }
        }
    }
}

// column sums over rows: db[n] += sum_m dY[m,n] ; dw2[n*ldw2] += sum_m dY[m,n] * u[m]
__global__ void __launch_bounds__(256) colsum_kernel(const float* __restrict__ dY, long lddy, const float* __restrict__ u,
                                                     float* db, float* dw2, long ldw2, long M, int N, long rows_per_block) {
    __shared__ float s1[8][33], s2[8][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int n = blockIdx.y * 32 + tx;
    const long m0 = (long)blockIdx.x * rows_per_block;
    const long m1 = min(M, m0 + rows_per_block);
    float a = 0.f, b = 0.f;
    if (n < N) {
        for (long m = m0 + ty; m < m1; m += 8) {
            const float g = __ldg(dY + m * lddy + n);
            a += g;
            if (u) b = fmaf(g, __ldg(u + m), b);
        }
    }
    s1[ty][tx] = a; s2[ty][tx] = b;
    __syncthreads();
    if (ty == 0 && n < N) {
#pragma unroll
        for (int i = 1; i < 8; ++i) { a += s1[i][tx]; b += s2[i][tx]; }
        if (db) atomicAdd(db + n, a);
        if (dw2) atomicAdd(dw2 + (long)n * ldw2, b);
    }
}

__global__ void relu_bwd_kernel(const float* __restrict__ dH, const float* __restrict__ H, float* __restrict__ dZ, long n) {
    long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const long stride = (long)gridDim.x * blockDim.x;
    for (; i < n; i += stride) dZ[i] = H[i] > 0.f ? dH[i] : 0.f;
}

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

template <bool A_RC, bool B_RC>
static int launch_gemm(GemmParams& p, int splits, cudaStream_t st) {
    dim3 grid((unsigned)cdiv(p.Q, BN), (unsigned)cdiv(p.P, BM), (unsigned)splits);
    gemm_f32_kernel<A_RC, B_RC><<<grid, 256, 0, st>>>(p);
    count_launch();
    return check_launch("gemm_f32_kernel");
}

// implemented in gemm_tc.cu: returns NPF_ENOTSUP when the shape is not covered by the tensor-core path
int linear_fwd_tc(const float* X, int ldx, const float* W, int ldw, const float* b, float* Y, int ldy, int M, int K,
                  int N, int flags, const float* u, const float* w2, int ldw2, int precision, cudaStream_t st);
int linear_bwd_data_tc(const float* dY, int lddy, const float* W, int ldw, float* dX, int lddx, int M, int K, int N,
                       const float* mask_src, int ldm, int flags, int precision, cudaStream_t st);
int thin128_in_fwd(const float* X, long ldx, const float* W, long ldw, const float* b, float* Y, long ldy, long M, int R, int relu_in,
                   int relu_out, cudaStream_t st);
int thin128_in_bwd(const float* dY, long lddy, const float* X, long ldx, const float* W, long ldw, float* dX, long lddx, float* dW, long lddw,
                   float* db, long M, int R, int relu_in, cudaStream_t st);
int thin128_out_fwd(const float* X, long ldx, const float* W, long ldw, const float* b, float* Y, long ldy, long M, int J, int relu_in,
                    int relu_out, cudaStream_t st);
int thin128_out_bwd(const float* dY, long lddy, const float* X, long ldx, const float* W, long ldw, float* dX, long lddx, float* dW, long lddw,
                    float* db, long M, int J, int relu_in, int use_mask, cudaStream_t st);
int mlp_chain_fwd_tc(const float* X, int ldx, const float* const* W, const int* ldw, const float* const* b, float* const* Y, const int* ldy, int L, int M,
                     int relu_in, unsigned relu_mask, int precision, cudaStream_t st);
int mlp_chain_bwd_tc(const float* dY, int lddy, const float* const* X, const float* const* W, float* dX, int lddx, float* const* dW, float* const* db,
                     int L, int M, int mask0, int precision, cudaStream_t st);
int linear_bwd_fused_tc(const float* dY, int lddy, const float* X, int ldx, const float* W, int ldw, float* dX, int lddx, float* dW,
                        int lddw, float* db, int M, int K, int N, int flags, int precision, cudaStream_t st);
int linear_bwd_weight_tc(const float* dY, int lddy, const float* X, int ldx, float* dW, int lddw, float* db, int* db_done, int M,
                         int K, int N, int flags, int precision, cudaStream_t st);


}  // namespace npf

using namespace npf;

extern "C" int npf_linear_fwd(const float* X, int ldx, const float* W, int ldw, const float* b, float* Y, int ldy,
                              int M, int K, int N, int flags, const float* u, const float* w2, int ldw2,
                              int precision, npf_stream_t stream) {
    if (M == 0) return NPF_OK;
    NPF_REQUIRE(X && W && Y, "npf_linear_fwd: null pointer");
    NPF_REQUIRE(M >= 0 && K >= 1 && N >= 1, "npf_linear_fwd: bad shape M=%d K=%d N=%d", M, K, N);
    NPF_REQUIRE(ldx >= K && ldw >= K && ldy >= N, "npf_linear_fwd: leading dimension too small");
    NPF_REQUIRE((u == nullptr) == (w2 == nullptr), "npf_linear_fwd: u and w2 must be given together");
    if (M == 0) return NPF_OK;
    cudaStream_t st = as_stream(stream);
    if (!u && !(flags & NPF_ACCUM)) {      // wide side exactly 128: the specialised one-pass kernels
        int rc = NPF_ENOTSUP;
        if (K <= kThinMax && N == 128) rc = thin128_in_fwd(X, ldx, W, ldw, b, Y, ldy, M, K, (flags & NPF_RELU_IN) != 0, (flags & NPF_RELU_OUT) != 0, st);
        else if (N <= kThinMax && K == 128) rc = thin128_out_fwd(X, ldx, W, ldw, b, Y, ldy, M, N, (flags & NPF_RELU_IN) != 0, (flags & NPF_RELU_OUT) != 0, st);
        if (rc != NPF_ENOTSUP) return rc;
    }
    if (K <= kThinMax) {
        ThinRedParams t{};
        t.A = X; t.lda = ldx; t.relu_a = (flags & NPF_RELU_IN) ? 1 : 0;
        t.B = W; t.sb_o = ldw; t.sb_r = 1; t.bias = b; t.u = u; t.w2 = w2; t.ldw2 = ldw2;
        t.out = Y; t.ldo = ldy; t.M = M; t.R = K; t.O = N;
        t.relu_out = (flags & NPF_RELU_OUT) ? 1 : 0; t.accum = (flags & NPF_ACCUM) ? 1 : 0;
        int rc = thin_red(t, st);
        if (rc != NPF_ENOTSUP) return rc;
    } else if (N <= kThinMax && !u) {
        RowDotParams t{};
        t.A = X; t.lda = ldx; t.relu_a = (flags & NPF_RELU_IN) ? 1 : 0;
        t.B = W; t.sb_j = ldw; t.sb_i = 1; t.bias = b;
        t.out = Y; t.ldo = ldy; t.M = M; t.I = K; t.J = N;
        t.relu_out = (flags & NPF_RELU_OUT) ? 1 : 0; t.accum = (flags & NPF_ACCUM) ? 1 : 0;
        int rc = rowdot(t, st);
        if (rc != NPF_ENOTSUP) return rc;
    }
    if (precision != NPF_PREC_FP32) {
        int rc = linear_fwd_tc(X, ldx, W, ldw, b, Y, ldy, M, K, N, flags, u, w2, ldw2, precision, st);
        if (rc != NPF_ENOTSUP) return rc;
    }
    GemmParams p{};
    p.A = X; p.lda = ldx; p.B = W; p.ldb = ldw; p.C = Y; p.ldc = ldy;
    p.P = M; p.Q = N; p.R = K; p.r_per_split = K;
    p.relu_a = (flags & NPF_RELU_IN) ? 1 : 0;
    p.bias_q = b; p.u_p = u; p.w_q = w2; p.ldw2 = ldw2;
    p.relu_out = (flags & NPF_RELU_OUT) ? 1 : 0;
    p.accum = (flags & NPF_ACCUM) ? 1 : 0;
    p.a_vec = (ldx % 4 == 0) && aligned16(X);
    p.b_vec = (ldw % 4 == 0) && aligned16(W);
    p.c_vec = (ldy % 4 == 0) && aligned16(Y);
    return launch_gemm<true, true>(p, 1, st);
}

extern "C" int npf_linear_bwd_data(const float* dY, int lddy, const float* W, int ldw, float* dX, int lddx, int M,
                                   int K, int N, const float* mask_src, int ldm, int flags, int precision,
                                   npf_stream_t stream) {
    if (M == 0) return NPF_OK;
    NPF_REQUIRE(dY && W && dX, "npf_linear_bwd_data: null pointer");
    NPF_REQUIRE(M >= 0 && K >= 1 && N >= 1, "npf_linear_bwd_data: bad shape");
    NPF_REQUIRE(lddy >= N && ldw >= K && lddx >= K, "npf_linear_bwd_data: leading dimension too small");
    if (M == 0) return NPF_OK;
    cudaStream_t st = as_stream(stream);
    if (N <= kThinMax) {
        ThinRedParams t{};
        t.A = dY; t.lda = lddy; t.B = W; t.sb_o = 1; t.sb_r = ldw;
        t.mask = mask_src; t.ldm = ldm;
        t.out = dX; t.ldo = lddx; t.M = M; t.R = N; t.O = K; t.accum = (flags & NPF_ACCUM) ? 1 : 0;
        int rc = thin_red(t, st);
        if (rc != NPF_ENOTSUP) return rc;
    } else if (K <= kThinMax) {
        RowDotParams t{};
        t.A = dY; t.lda = lddy; t.B = W; t.sb_j = 1; t.sb_i = ldw;
        t.mask = mask_src; t.ldm = ldm;
        t.out = dX; t.ldo = lddx; t.M = M; t.I = N; t.J = K; t.accum = (flags & NPF_ACCUM) ? 1 : 0;
        int rc = rowdot(t, st);
        if (rc != NPF_ENOTSUP) return rc;
    }
    if (precision != NPF_PREC_FP32) {
        int rc = linear_bwd_data_tc(dY, lddy, W, ldw, dX, lddx, M, K, N, mask_src, ldm, flags, precision, st);
        if (rc != NPF_ENOTSUP) return rc;
    }
    GemmParams p{};
    p.A = dY; p.lda = lddy; p.B = W; p.ldb = ldw; p.C = dX; p.ldc = lddx;
    p.P = M; p.Q = K; p.R = N; p.r_per_split = N;
    p.mask = mask_src; p.ldm = ldm;
    p.accum = (flags & NPF_ACCUM) ? 1 : 0;
    p.a_vec = (lddy % 4 == 0) && aligned16(dY);
    p.b_vec = (ldw % 4 == 0) && aligned16(W);
    p.c_vec = (lddx % 4 == 0) && aligned16(dX);
    return launch_gemm<true, false>(p, 1, st);
}

extern "C" int npf_linear_bwd_weight(const float* dY, int lddy, const float* X, int ldx, float* dW, int lddw,
                                     float* db, int M, int K, int N, int flags, const float* u, float* dw2, int ldw2,
                                     int precision, npf_stream_t stream) {
    if (M == 0) return NPF_OK;
    NPF_REQUIRE(dY && X && dW, "npf_linear_bwd_weight: null pointer");
    NPF_REQUIRE(M >= 0 && K >= 1 && N >= 1, "npf_linear_bwd_weight: bad shape");
    NPF_REQUIRE(lddy >= N && ldx >= K && lddw >= K, "npf_linear_bwd_weight: leading dimension too small");
    NPF_REQUIRE((u == nullptr) == (dw2 == nullptr), "npf_linear_bwd_weight: u and dw2 must be given together");
    if (M == 0) return NPF_OK;
    cudaStream_t st = as_stream(stream);
    if (!u) {
        int rc128 = NPF_ENOTSUP;
        if (K <= kThinMax && N == 128) rc128 = thin128_in_bwd(dY, lddy, X, ldx, nullptr, 0, nullptr, 0, dW, lddw, db, M, K, (flags & NPF_RELU_IN) != 0, st);
        else if (N <= kThinMax && K == 128)
            rc128 = thin128_out_bwd(dY, lddy, X, ldx, nullptr, 0, nullptr, 0, dW, lddw, db, M, N, (flags & NPF_RELU_IN) != 0, 0, st);
        if (rc128 != NPF_ENOTSUP) return rc128;
    }
    if (K <= kThinMax) {   // one pass over dY: dW, db and the rank-1 column together
        ThinOuterParams t{};
        t.S = X; t.lds = ldx; t.relu_s = (flags & NPF_RELU_IN) ? 1 : 0;
        t.T = dY; t.ldt = lddy;
        t.out = dW; t.so_j = 1; t.so_c = lddw; t.out_ones = db; t.u = u; t.out_u = dw2; t.so_u = ldw2;
        t.M = M; t.J = K; t.C = N;
        return thin_outer(t, st);
    }
    int rc = NPF_ENOTSUP;
    if (N <= kThinMax) {
        ThinOuterParams t{};
        t.S = dY; t.lds = lddy; t.T = X; t.ldt = ldx; t.relu_t = (flags & NPF_RELU_IN) ? 1 : 0;
        t.out = dW; t.so_j = lddw; t.so_c = 1;
        t.M = M; t.J = N; t.C = K;
        rc = thin_outer(t, st);
    }
    int db_done = 0;
    if (rc == NPF_ENOTSUP && precision != NPF_PREC_FP32) {
        rc = linear_bwd_weight_tc(dY, lddy, X, ldx, dW, lddw, db, &db_done, M, K, N, flags, precision, st);
        if (rc != NPF_OK) db_done = 0;
    }
    if (rc == NPF_ENOTSUP) {
        GemmParams p{};
        p.A = dY; p.lda = lddy; p.B = X; p.ldb = ldx; p.C = dW; p.ldc = lddw;
        p.P = N; p.Q = K; p.R = M;
        p.relu_b = (flags & NPF_RELU_IN) ? 1 : 0;
        const long tiles = cdiv(N, BM) * cdiv(K, BN);
        long splits = cdiv(4L * kNumSMs, tiles);
        const long max_splits = cdiv(M, 8 * BK);
        if (splits > max_splits) splits = max_splits;
        if (splits < 1) splits = 1;
        long rps = cdiv(cdiv(M, splits), BK) * BK;
        splits = cdiv(M, rps);
        p.r_per_split = (int)rps;
        p.atomic = 1;
        p.a_vec = (lddy % 4 == 0) && aligned16(dY);
        p.b_vec = (ldx % 4 == 0) && aligned16(X);
        rc = launch_gemm<false, false>(p, (int)splits, st);
    }
    if (rc != NPF_OK) return rc;
    if (db_done) db = nullptr;
    if (db || dw2) {
        long rows_per_block = cdiv(M, 2L * kNumSMs);
        if (rows_per_block < 64) rows_per_block = 64;
        dim3 grid((unsigned)cdiv(M, rows_per_block), (unsigned)cdiv(N, 32));
        colsum_kernel<<<grid, 256, 0, st>>>(dY, lddy, u, db, dw2, ldw2, M, N, rows_per_block);
        count_launch();
        return check_launch("colsum_kernel");
    }
    return NPF_OK;
}

extern "C" int npf_relu_bwd(const float* dH, const float* H, float* dZ, long n, npf_stream_t stream) {
    NPF_REQUIRE(dH && H && dZ, "npf_relu_bwd: null pointer");
    if (n == 0) return NPF_OK;
    long blocks = cdiv(n, 256);
    if (blocks > 8L * kNumSMs) blocks = 8L * kNumSMs;
    relu_bwd_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(dH, H, dZ, n);
    count_launch();
    return check_launch("relu_bwd_kernel");
}

extern "C" int npf_linear_bwd(const float* dY, int lddy, const float* X, int ldx, const float* W, int ldw, float* dX, int lddx,
                              float* dW, int lddw, float* db, int M, int K, int N, int flags, int precision, npf_stream_t stream) {
    if (M == 0) return NPF_OK;
    NPF_REQUIRE(dY && X && W && dW, "npf_linear_bwd: null pointer");
    NPF_REQUIRE(M >= 0 && K >= 1 && N >= 1, "npf_linear_bwd: bad shape");
    NPF_REQUIRE(lddy >= N && ldx >= K && ldw >= K && (!dX || lddx >= K) && lddw >= K, "npf_linear_bwd: leading dimension too small");
    NPF_REQUIRE(!(flags & ~(NPF_RELU_IN | NPF_MASK_X)), "npf_linear_bwd: unsupported flag");
    {
        cudaStream_t st = npf::as_stream(stream);
        int rc = NPF_ENOTSUP;
        if (K <= npf::kThinMax && N == 128 && !(flags & NPF_MASK_X))
            rc = npf::thin128_in_bwd(dY, lddy, X, ldx, W, ldw, dX, lddx, dW, lddw, db, M, K, (flags & NPF_RELU_IN) != 0, st);
        else if (N <= npf::kThinMax && K == 128)
            rc = npf::thin128_out_bwd(dY, lddy, X, ldx, W, ldw, dX, lddx, dW, lddw, db, M, N, (flags & NPF_RELU_IN) != 0, (flags & NPF_MASK_X) != 0, st);
        if (rc != NPF_ENOTSUP) return rc;
    }
    if (precision != NPF_PREC_FP32 && dX) {
        const int rc = npf::linear_bwd_fused_tc(dY, lddy, X, ldx, W, ldw, dX, lddx, dW, lddw, db, M, K, N, flags, precision, npf::as_stream(stream));
        if (rc != NPF_ENOTSUP) return rc;
    }
    int rc = npf_linear_bwd_weight(dY, lddy, X, ldx, dW, lddw, db, M, K, N, flags & NPF_RELU_IN, nullptr, nullptr, 0, precision, stream);
    if (rc != NPF_OK || !dX) return rc;
    return npf_linear_bwd_data(dY, lddy, W, ldw, dX, lddx, M, K, N, (flags & NPF_MASK_X) ? X : nullptr, ldx, 0, precision, stream);
}

extern "C" int npf_mlp_chain_bwd(const float* dY, int lddy, const float* const* X, const float* const* W, float* dX, int lddx, float* const* dW,
                                 float* const* db, int L, int M, int width, int flags, int precision, npf_stream_t stream) {
    if (M == 0) return NPF_OK;
    NPF_REQUIRE(dY && X && W && dW && L >= 1 && L <= 8, "npf_mlp_chain_bwd: null pointer or bad layer count");
    NPF_REQUIRE(M >= 0 && width >= 1 && lddy >= width && (!dX || lddx >= width), "npf_mlp_chain_bwd: bad shape");
    for (int l = 0; l < L; ++l) NPF_REQUIRE(X[l] && W[l] && dW[l], "npf_mlp_chain_bwd: null layer pointer");
    if (width == 128) {
        const int rc = npf::mlp_chain_bwd_tc(dY, lddy, X, W, dX, lddx, dW, db, L, M, (flags & NPF_MASK_X) ? 1 : 0, precision, npf::as_stream(stream));
        if (rc != NPF_ENOTSUP) return rc;
    }
    // layer by layer: needs a scratch gradient per intermediate layer, which this allocation-free entry point does not own
    npf::set_error("npf_mlp_chain_bwd: shape / precision not covered by the on-chip chain (width 128, tensor-core precision, M >= 64, "
                   "16-byte aligned rows); call npf_linear_bwd per layer");
    return NPF_ENOTSUP;
}

extern "C" int npf_mlp_chain_fwd(const float* X, int ldx, const float* const* W, const float* const* b, float* const* Y, int L, int M, int width,
                                 int relu_in, unsigned relu_mask, int precision, npf_stream_t stream) {
    if (M == 0) return NPF_OK;
    NPF_REQUIRE(X && W && Y && L >= 1 && L <= 8, "npf_mlp_chain_fwd: null pointer or bad layer count");
    NPF_REQUIRE(M >= 0 && width >= 1 && ldx >= width, "npf_mlp_chain_fwd: bad shape");
    for (int l = 0; l < L; ++l) NPF_REQUIRE(W[l] && Y[l], "npf_mlp_chain_fwd: null layer pointer");
    int ld[8];
    for (int l = 0; l < L; ++l) ld[l] = width;
    if (width == 128) {
        const int rc = npf::mlp_chain_fwd_tc(X, ldx, W, ld, b, Y, ld, L, M, relu_in, relu_mask, precision, npf::as_stream(stream));
        if (rc != NPF_ENOTSUP) return rc;
    }
    // layer by layer (any width, any precision)
    const float* in = X;
    int ldin = ldx;
    for (int l = 0; l < L; ++l) {
        const int flags = ((l == 0 && relu_in) ? NPF_RELU_IN : 0) | (((relu_mask >> l) & 1u) ? NPF_RELU_OUT : 0);
        const int rc = npf_linear_fwd(in, ldin, W[l], width, b ? b[l] : nullptr, Y[l], width, M, width, width, flags, nullptr, nullptr, 0, precision, stream);
        if (rc != NPF_OK) return rc;
        in = Y[l]; ldin = width;
    }
    return NPF_OK;
}

