// Gaussian-process prior sampler on the device: the synthetic-task generator in front of the hot path
// (second half).
//
// Upstream draws every epoch's 50 000 tasks on the host with scikit-learn: GPDataset._sample_targets
// (utils/data/gaussian_process.py:201-231) -> GaussianProcessRegressor.sample_y on the un-fitted regressor, i.e.
// y ~ N(0, k(X, X)) by numpy's SVD-based multivariate_normal, for the kernels of utils/ntbks_helpers.py:76-108:
//   RBF(l)                       k = exp(-d^2 / (2 l^2))
//   Matern(l, nu = 1.5)          k = (1 + sqrt(3) d / l) exp(-sqrt(3) d / l)
//   ExpSineSquared(l, p)         k = exp(-2 sin^2(pi d / p) / l^2)
//   WhiteKernel(s) + .           + s on the diagonal
// These covariance matrices are numerically rank-deficient (RBF, 128 points on [-2, 2], l = 0.2: rank ~ 40 at 1e-6), so a
// plain Cholesky fails and a jittered one changes the law.  Here: DIAGONALLY PIVOTED Cholesky with early termination,
// K = L L^T + E with max(diag E) <= tol, run entirely in one CTA's shared memory with K generated on the fly from x (K is
// never stored, HBM sees x, eps and y only); then y_s = L eps_s for S samples sharing the same x (upstream's
// n_same_samples).  fp32 throughout: the pivoted factorisation is backward stable for PSD matrices, the residual
// |L L^T - K| is ~1e-6 (tests/test_gpu_gp.py checks it against scikit-learn's own K).
// Work per task: N^3/3 MACs out of shared memory (0.7 MF at N = 128) -- LDS/FFMA-bound, no tensor cores at this size.
#include "common.cuh"

namespace npf {

constexpr int kGpThreads = 256;   // one thread per point (N <= 256); the smem factor caps N at 232 anyway
constexpr int kGpMaxN = 232;      // N * N * 4 + small <= 227 KB

struct GpKernel {
    int kind;          // 0 RBF, 1 Matern-1.5, 2 ExpSineSquared
    float length_scale, periodicity, noise;
};

__device__ __forceinline__ float gp_cov(const GpKernel& g, float xi, float xj) {
    const float d = fabsf(xi - xj);
    if (g.kind == 0) {
        const float r = d / g.length_scale;
        return expf(-0.5f * r * r);
    }
    if (g.kind == 1) {
        const float r = 1.7320508075688772f * d / g.length_scale;
        return (1.f + r) * expf(-r);
    }
    const float s = sinpif(d / g.periodicity) / g.length_scale;
    return expf(-2.f * s * s);
}

// One CTA per task b.  Lc[k * N + i] = L[i, k] (column-major: threads i of a warp hit consecutive banks, L[p, k] is a
// broadcast).  Step k: p = argmax_i d_i (smallest index among ties); stop when d_p <= tol; column k from
// K[:, p] - L[:, :k] L[p, :k]^T; d_i -= L[i, k]^2.
__global__ void __launch_bounds__(kGpThreads, 1)
gp_sample_kernel(const float* __restrict__ X, const float* __restrict__ eps, float* __restrict__ Y, float* __restrict__ Lout,
                 int32_t* __restrict__ rank_out, int N, int S, GpKernel g, const float* __restrict__ hyp, float tol) {
    extern __shared__ float gp_smem[];
    float* Lc = gp_smem;                 // [N][N]
    float* xs = Lc + (size_t)N * N;      // [N]
    float* es = xs + N;                  // [N]   eps of the current sample
    __shared__ float s_val[kGpThreads / 32];
    __shared__ int s_idx[kGpThreads / 32];
    __shared__ float s_piv;
    __shared__ int s_p;

    const int b = blockIdx.x, i = threadIdx.x, lane = i & 31, warp = i >> 5;
    const bool active = i < N;
    if (hyp) {      // per-task hyper-parameters (upstream `is_vary_kernel_hyp`): (length_scale, periodicity, noise_level) of task b
        g.length_scale = __ldg(hyp + 3 * (long)b);
        g.periodicity = __ldg(hyp + 3 * (long)b + 1);
        g.noise = __ldg(hyp + 3 * (long)b + 2);
    }
    if (active) xs[i] = X[(long)b * N + i];
    __syncthreads();
    const float xi = active ? xs[i] : 0.f;
    float d = active ? 1.f + g.noise : -1.f;   // k(x, x) = 1 for the three stationary kernels
    bool done = !active;
    int rank = 0;
    for (int k = 0; k < N; ++k) {
        // ---- pivot: block arg-max of the residual diagonal
        float v = done ? -1.f : d;
        int idx = i;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float v2 = __shfl_xor_sync(0xffffffffu, v, o);
            const int i2 = __shfl_xor_sync(0xffffffffu, idx, o);
            if (v2 > v || (v2 == v && i2 < idx)) { v = v2; idx = i2; }
        }
        if (lane == 0) { s_val[warp] = v; s_idx[warp] = idx; }
        __syncthreads();
        if (i == 0) {
            float bv = s_val[0];
            int bi = s_idx[0];
            for (int w = 1; w < kGpThreads / 32; ++w)
                if (s_val[w] > bv || (s_val[w] == bv && s_idx[w] < bi)) { bv = s_val[w]; bi = s_idx[w]; }
            s_piv = bv;
            s_p = bi;
        }
        __syncthreads();
        const float piv = s_piv;
        const int p = s_p;
        if (!(piv > tol)) break;           // uniform: the remaining variance of every point is <= tol
        rank = k + 1;
        // ---- column k
        float lik = 0.f;
        if (active) {
            if (i == p) {
                lik = sqrtf(piv);
            } else if (!done) {
                float acc = gp_cov(g, xi, xs[p]);
                for (int j = 0; j < k; ++j) acc = fmaf(-Lc[j * N + i], Lc[j * N + p], acc);
                lik = acc * rsqrtf(piv);
            }
            Lc[k * N + i] = lik;
            d = fmaf(-lik, lik, d);
            if (i == p) done = true;
        }
        __syncthreads();
    }
    // ---- samples: y_s = L eps_s  (eps indexed by factorisation step)
    for (int s = 0; s < S; ++s) {
        __syncthreads();
        if (active) es[i] = eps[((long)b * S + s) * N + i];
        __syncthreads();
        if (active) {
            float acc = 0.f;
            for (int j = 0; j < rank; ++j) acc = fmaf(Lc[j * N + i], es[j], acc);
            Y[((long)b * S + s) * N + i] = acc;
        }
    }
    if (Lout && active)
        for (int j = 0; j < N; ++j) Lout[((long)b * N + i) * N + j] = j < rank ? Lc[j * N + i] : 0.f;
    if (rank_out && i == 0) rank_out[b] = rank;
}

}  // namespace npf

using namespace npf;

static int gp_launch(const float* X, const float* eps, float* Y, float* L, int32_t* rank, const float* hyp, int B, int N, int S, GpKernel g,
                     float tol, npf_stream_t stream) {
    if (N > kGpMaxN) {
        set_error("npf_gp_sample: N=%d exceeds the %d points whose factor fits one CTA's shared memory", N, kGpMaxN);
        return NPF_ENOTSUP;
    }
    if (B == 0) return NPF_OK;
    NPF_REQUIRE(X && (S == 0 || (eps && Y)), "npf_gp_sample: null pointer");
    const size_t smem = ((size_t)N * N + 2 * (size_t)N) * sizeof(float);
    static size_t configured = 0;   // grows monotonically; the attribute is per function, not per stream
    if (smem > 48 * 1024 && smem > configured) {
        if (cudaFuncSetAttribute(gp_sample_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
            return check_launch("npf_gp_sample: cudaFuncSetAttribute");
        configured = smem;
    }
    gp_sample_kernel<<<B, kGpThreads, smem, as_stream(stream)>>>(X, eps, Y, L, rank, N, S, g, hyp, tol);
    count_launch();
    return check_launch("gp_sample_kernel");
}

extern "C" int npf_gp_sample(const float* X, const float* eps, float* Y, float* L, int32_t* rank, int B, int N, int S, int kernel,
                             float length_scale, float periodicity, float noise_level, float tol, npf_stream_t stream) {
    NPF_REQUIRE(B >= 0 && N >= 1 && S >= 0, "npf_gp_sample: bad shape");
    NPF_REQUIRE(kernel >= 0 && kernel <= 2, "npf_gp_sample: kernel must be 0 (RBF), 1 (Matern-1.5) or 2 (ExpSineSquared)");
    NPF_REQUIRE(length_scale > 0.f && (kernel != 2 || periodicity > 0.f) && noise_level >= 0.f && tol >= 0.f,
                "npf_gp_sample: bad hyper-parameter");
    return gp_launch(X, eps, Y, L, rank, nullptr, B, N, S, GpKernel{kernel, length_scale, periodicity, noise_level}, tol, stream);
}

extern "C" int npf_gp_sample_hyp(const float* X, const float* eps, float* Y, float* L, int32_t* rank, const float* hyp, int B, int N, int S,
                                 int kernel, float tol, npf_stream_t stream) {
    NPF_REQUIRE(B >= 0 && N >= 1 && S >= 0, "npf_gp_sample_hyp: bad shape");
    NPF_REQUIRE(kernel >= 0 && kernel <= 2, "npf_gp_sample_hyp: kernel must be 0 (RBF), 1 (Matern-1.5) or 2 (ExpSineSquared)");
    NPF_REQUIRE(tol >= 0.f && (B == 0 || hyp), "npf_gp_sample_hyp: bad tolerance / null hyper-parameter table");
    return gp_launch(X, eps, Y, L, rank, hyp, B, N, S, GpKernel{kernel, 1.f, 1.f, 0.f}, tol, stream);
}
