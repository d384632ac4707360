// Small fused element-wise / reduction kernels around the three hot kernel groups: sum-merge, context mean
// pooling, add+LayerNorm, predictive head, Gaussian log-likelihood, latent sampling, global latent, input check.
// All HBM-bound streaming kernels: grid-stride, coalesced on the channel (last) axis.
#include <cmath>

#include "common.cuh"

namespace npf {

constexpr float kHalfLog2Pi = 0.91893853320467274178f;  // log(sqrt(2 pi)), torch Normal.log_prob

static inline unsigned grid_for(long n, int block = 256) {
    long g = cdiv(n, block);
    const long cap = 64L * kNumSMs;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (unsigned)g;
}

// ------------------------------------------------------------------------------------------------ sum-merge
__global__ void merge_relu_fwd_kernel(const float* __restrict__ x1, const float* __restrict__ x2, float* __restrict__ out,
                                      int Z, int B, int T, int C, int x2_has_t) {
    const long n = (long)Z * B * T * C;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        long r = i / C;
        const int t = (int)(r % T); r /= T;
        const int b = (int)(r % B);
        const int z = (int)(r / B);
        const float a = __ldg(x1 + ((long)b * T + t) * C + c);
        const float v = x2_has_t ? __ldg(x2 + (((long)z * B + b) * T + t) * C + c) : __ldg(x2 + ((long)z * B + b) * C + c);
        out[i] = fmaxf(a + v, 0.f);
    }
}

// dx1[b,t,c] = sum_z dpre[z,b,t,c]
__global__ void merge_relu_bwd_x1_kernel(const float* __restrict__ dout, const float* __restrict__ out, float* __restrict__ dx1,
                                         int Z, long BTC) {
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < BTC; i += (long)gridDim.x * blockDim.x) {
        float s = 0.f;
        for (int z = 0; z < Z; ++z) {
            const long o = (long)z * BTC + i;
            s += __ldg(out + o) > 0.f ? __ldg(dout + o) : 0.f;
        }
        dx1[i] = s;
    }
}

// dx2[z,b,(t),c]: has_t -> dpre ; else sum_t dpre
__global__ void merge_relu_bwd_x2_kernel(const float* __restrict__ dout, const float* __restrict__ out, float* __restrict__ dx2,
                                         long ZB, int T, int C, int x2_has_t) {
    if (x2_has_t) {
        const long n = ZB * T * C;
        for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x)
            dx2[i] = __ldg(out + i) > 0.f ? __ldg(dout + i) : 0.f;
    } else {
        const long n = ZB * C;
        for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
            const int c = (int)(i % C);
            const long zb = i / C;
            float s = 0.f;
            for (int t = 0; t < T; ++t) {
                const long o = (zb * T + t) * C + c;
                s += __ldg(out + o) > 0.f ? __ldg(dout + o) : 0.f;
            }
            dx2[i] = s;
        }
    }
}

// ------------------------------------------------------------------------------------------------ mean pool
__global__ void mean_pool_fwd_kernel(const float* __restrict__ X, float* __restrict__ R, int B, int N, int C) {
    const long n = (long)B * C;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        const long b = i / C;
        float s = 0.f;
        for (int k = 0; k < N; ++k) s += __ldg(X + (b * N + k) * C + c);
        R[i] = s / (float)N;
    }
}
__global__ void mean_pool_bwd_kernel(const float* __restrict__ dR, float* __restrict__ dX, int B, int N, int C) {
    const long n = (long)B * N * C;
    const float inv = 1.f / (float)N;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        const long b = i / ((long)N * C);
        dX[i] = __ldg(dR + b * C + c) * inv;
    }
}

// ------------------------------------------------------------------------------------------------ add + LayerNorm
constexpr int kLnChunks = 8;  // C <= 256

__global__ void __launch_bounds__(256) add_layernorm_fwd_kernel(const float* __restrict__ A, const float* __restrict__ Bm,
                                                                const float* __restrict__ gamma, const float* __restrict__ beta,
                                                                float* __restrict__ Y, float* __restrict__ rstat, long M, int C) {
    const int lane = threadIdx.x & 31;
    const long warp = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long nwarps = ((long)gridDim.x * blockDim.x) >> 5;
    const int nch = (C + 31) >> 5;
    for (long m = warp; m < M; m += nwarps) {
        float v[kLnChunks];
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < kLnChunks; ++i) {
            const int c = lane + 32 * i;
            v[i] = (i < nch && c < C) ? __ldg(A + m * C + c) + __ldg(Bm + m * C + c) : 0.f;
            s += v[i];
        }
        const float mean = warp_sum(s) / (float)C;
        float q = 0.f;
#pragma unroll
        for (int i = 0; i < kLnChunks; ++i) {
            const int c = lane + 32 * i;
            if (i < nch && c < C) { const float d = v[i] - mean; q = fmaf(d, d, q); }
        }
        const float rstd = rsqrtf(warp_sum(q) / (float)C + 1e-5f);
#pragma unroll
        for (int i = 0; i < kLnChunks; ++i) {
            const int c = lane + 32 * i;
            if (i < nch && c < C) Y[m * C + c] = (v[i] - mean) * rstd * __ldg(gamma + c) + __ldg(beta + c);
        }
        if (lane == 0) { rstat[m * 2] = mean; rstat[m * 2 + 1] = rstd; }
    }
}

__global__ void __launch_bounds__(256) add_layernorm_bwd_kernel(const float* __restrict__ dY, const float* __restrict__ A,
                                                                const float* __restrict__ Bm, const float* __restrict__ gamma,
                                                                const float* __restrict__ rstat, float* __restrict__ dS,
                                                                float* __restrict__ dgamma, float* __restrict__ dbeta, long M, int C) {
    __shared__ float sg[32 * kLnChunks], sb[32 * kLnChunks];
    const int lane = threadIdx.x & 31;
    const long warp = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long nwarps = ((long)gridDim.x * blockDim.x) >> 5;
    const int nch = (C + 31) >> 5;
    for (int i = threadIdx.x; i < 32 * kLnChunks; i += blockDim.x) { sg[i] = 0.f; sb[i] = 0.f; }
    __syncthreads();
    float g_acc[kLnChunks], b_acc[kLnChunks], gam[kLnChunks];
#pragma unroll
    for (int i = 0; i < kLnChunks; ++i) {
        const int c = lane + 32 * i;
        g_acc[i] = 0.f; b_acc[i] = 0.f;
        gam[i] = (i < nch && c < C) ? __ldg(gamma + c) : 0.f;
    }
    for (long m = warp; m < M; m += nwarps) {
        const float mean = __ldg(rstat + m * 2), rstd = __ldg(rstat + m * 2 + 1);
        float xh[kLnChunks], dxh[kLnChunks];
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int i = 0; i < kLnChunks; ++i) {
            const int c = lane + 32 * i;
            if (i < nch && c < C) {
                const float dy = __ldg(dY + m * C + c);
                xh[i] = (__ldg(A + m * C + c) + __ldg(Bm + m * C + c) - mean) * rstd;
                dxh[i] = dy * gam[i];
                s1 += dxh[i];
                s2 = fmaf(dxh[i], xh[i], s2);
                g_acc[i] = fmaf(dy, xh[i], g_acc[i]);
                b_acc[i] += dy;
            } else { xh[i] = 0.f; dxh[i] = 0.f; }
        }
        s1 = warp_sum(s1) / (float)C;
        s2 = warp_sum(s2) / (float)C;
#pragma unroll
        for (int i = 0; i < kLnChunks; ++i) {
            const int c = lane + 32 * i;
            if (i < nch && c < C) dS[m * C + c] = rstd * (dxh[i] - s1 - xh[i] * s2);
        }
    }
#pragma unroll
    for (int i = 0; i < kLnChunks; ++i) {
        const int c = lane + 32 * i;
        if (i < nch && c < C) { atomicAdd(sg + c, g_acc[i]); atomicAdd(sb + c, b_acc[i]); }
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        atomicAdd(dgamma + c, sg[c]);
        atomicAdd(dbeta + c, sb[c]);
    }
}

// ------------------------------------------------------------------------------------------------ predictive head
__global__ void gauss_head_fwd_kernel(const float* __restrict__ suff, float* __restrict__ loc, float* __restrict__ scale, long M,
                                      int y, float min_scale) {
    const long n = M * y;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        const long m = i / y;
        const int j = (int)(i % y);
        loc[i] = __ldg(suff + m * 2 * y + j);
        scale[i] = min_scale + (1.f - min_scale) * softplus_f(__ldg(suff + m * 2 * y + y + j));
    }
}
__global__ void gauss_head_bwd_kernel(const float* __restrict__ suff, const float* __restrict__ dloc, const float* __restrict__ dscale,
                                      float* __restrict__ dsuff, long M, int y, float min_scale) {
    const long n = M * y;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        const long m = i / y;
        const int j = (int)(i % y);
        const float s = __ldg(suff + m * 2 * y + y + j);
        const float dsp = s > 20.f ? 1.f : sigmoid_f(s);
        dsuff[m * 2 * y + j] = dloc ? __ldg(dloc + i) : 0.f;
        dsuff[m * 2 * y + y + j] = dscale ? __ldg(dscale + i) * (1.f - min_scale) * dsp : 0.f;
    }
}

// ------------------------------------------------------------------------------------------------ Gaussian log-lik
// one CTA per (z, b): slp[z,b] = sum_i -d^2/(2 s^2) - log s - log sqrt(2 pi)
__global__ void __launch_bounds__(256) gauss_nll_fwd_kernel(const float* __restrict__ loc, const float* __restrict__ scale,
                                                            const float* __restrict__ Y, float* __restrict__ slp, int B, long n) {
    __shared__ float part[8];
    const long zb = blockIdx.x;
    const int b = (int)(zb % B);
    const float* mu = loc + zb * n;
    const float* sg = scale + zb * n;
    const float* y = Y + (long)b * n;
    float s = 0.f;
    for (long i = threadIdx.x; i < n; i += blockDim.x) {
        const float sc = __ldg(sg + i), d = __ldg(y + i) - __ldg(mu + i);
        s += -(d * d) / (2.f * sc * sc) - logf(sc) - kHalfLog2Pi;
    }
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) t += part[i];
        slp[zb] = t;
    }
}
__global__ void gauss_nll_bwd_kernel(const float* __restrict__ loc, const float* __restrict__ scale, const float* __restrict__ Y,
                                     const float* __restrict__ g, float* __restrict__ dloc, float* __restrict__ dscale, int B, long n,
                                     long total) {
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const long zb = i / n, k = i % n;
        const int b = (int)(zb % B);
        const float sc = __ldg(scale + i), d = __ldg(Y + (long)b * n + k) - __ldg(loc + i);
        const float gz = __ldg(g + zb);
        const float inv = 1.f / sc;
        dloc[i] = gz * d * inv * inv;
        dscale[i] = gz * (d * d * inv * inv * inv - inv);
    }
}

// ------------------------------------------------------------------------------------------------ latent path
__global__ void latent_sample_fwd_kernel(const float* __restrict__ suff, const float* __restrict__ eps, float* __restrict__ q_loc,
                                         float* __restrict__ q_scale, float* __restrict__ z, int S, long M, int zd) {
    const long n = M * zd;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        const long m = i / zd;
        const int j = (int)(i % zd);
        const float lo = __ldg(suff + m * 2 * zd + j);
        const float sc = 0.1f + 0.9f * sigmoid_f(__ldg(suff + m * 2 * zd + zd + j));
        q_loc[i] = lo;
        q_scale[i] = sc;
        for (int s = 0; s < S; ++s) z[(long)s * n + i] = fmaf(sc, __ldg(eps + (long)s * n + i), lo);
    }
}
__global__ void latent_sample_bwd_kernel(const float* __restrict__ suff, const float* __restrict__ eps, const float* __restrict__ dz,
                                         const float* __restrict__ dq_loc, const float* __restrict__ dq_scale, float* __restrict__ dsuff,
                                         int S, long M, int zd) {
    const long n = M * zd;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        const long m = i / zd;
        const int j = (int)(i % zd);
        float dl = dq_loc ? __ldg(dq_loc + i) : 0.f;
        float ds = dq_scale ? __ldg(dq_scale + i) : 0.f;
        if (dz)
            for (int s = 0; s < S; ++s) {
                const float g = __ldg(dz + (long)s * n + i);
                dl += g;
                ds = fmaf(g, __ldg(eps + (long)s * n + i), ds);
            }
        const float sg = sigmoid_f(__ldg(suff + m * 2 * zd + zd + j));
        dsuff[m * 2 * zd + j] = dl;
        dsuff[m * 2 * zd + zd + j] = ds * 0.9f * sg * (1.f - sg);
    }
}

// out[n,p,c] = c < C/2 ? zin[n,p,c] : mean_p zin[n,p,c]       (grid.x = N, threads over channels)
__global__ void global_latent_kernel(const float* __restrict__ in, float* __restrict__ out, int P, int C) {
    const long base = (long)blockIdx.x * P * C;
    const int h = C / 2;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        if (c < h) {
            for (int p = 0; p < P; ++p) out[base + (long)p * C + c] = __ldg(in + base + (long)p * C + c);
        } else {
            float s = 0.f;
            for (int p = 0; p < P; ++p) s += __ldg(in + base + (long)p * C + c);
            s /= (float)P;
            for (int p = 0; p < P; ++p) out[base + (long)p * C + c] = s;
        }
    }
}

// ------------------------------------------------------------------------------------------------ input check
__global__ void range_check_kernel(const float* __restrict__ X, long n, float lo, float hi, int* flag) {
    bool bad = false;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        const float v = __ldg(X + i);
        bad |= !(v >= lo && v <= hi);
    }
    if (__any_sync(0xffffffffu, bad) && (threadIdx.x & 31) == 0) atomicOr(flag, 1);
}

}  // namespace npf

using namespace npf;


// ------------------------------------------------------------------------------------------------ optimizer (next row)
// torch.optim.Adam (utils/train.py:50, the optimizer every notebook passes to skorch) on flat fp32 buffers: one elementwise pass
// over (param, grad, exp_avg, exp_avg_sq) instead of ~40 small foreach launches.  Same update order as torch's single-tensor path:
//   g' = g * grad_scale + wd * p;  m = b1 m + (1 - b1) g';  v = b2 v + (1 - b2) g'^2;  p -= step_size * m / (sqrt(v) / sqrt(bc2) + eps)
__global__ void adam_kernel(float* __restrict__ P, const float* __restrict__ G, float* __restrict__ Mo, float* __restrict__ V, long n, float step_size,
                            float inv_sqrt_bc2, float b1, float b2, float eps, float wd, float grad_scale) {
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        const float p = P[i];
        const float g = fmaf(wd, p, G[i] * grad_scale);
        const float m = fmaf(b1, Mo[i], (1.f - b1) * g);
        const float v = fmaf(b2, V[i], (1.f - b2) * g * g);
        Mo[i] = m;
        V[i] = v;
        P[i] = p - step_size * (m / (sqrtf(v) * inv_sqrt_bc2 + eps));
    }
}

// sum of squares of a flat buffer (global gradient norm for clipping): warp shuffle + one atomic per block
__global__ void sqnorm_kernel(const float* __restrict__ X, long n, float* __restrict__ out) {
    float s = 0.f;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        const float v = __ldg(X + i);
        s = fmaf(v, v, s);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    __shared__ float part[8];
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += part[w];
        atomicAdd(out, t);
    }
}

// Adam with the clipping factor of torch.nn.utils.clip_grad_norm_ taken from a DEVICE scalar (no host sync):
//   coef = min(1, max_norm / (grad_scale * sqrt(sqnorm) + 1e-6))
__global__ void adam_clipped_kernel(float* __restrict__ P, const float* __restrict__ G, float* __restrict__ Mo, float* __restrict__ V, long n,
                                    float step_size, float inv_sqrt_bc2, float b1, float b2, float eps, float wd, float grad_scale,
                                    const float* __restrict__ sqnorm, float max_norm) {
    const float coef = fminf(1.f, max_norm / (grad_scale * sqrtf(__ldg(sqnorm)) + 1e-6f));
    const float gs = grad_scale * coef;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        const float p = P[i];
        const float g = fmaf(wd, p, G[i] * gs);
        const float m = fmaf(b1, Mo[i], (1.f - b1) * g);
        const float v = fmaf(b2, V[i], (1.f - b2) * g * g);
        Mo[i] = m;
        V[i] = v;
        P[i] = p - step_size * (m / (sqrtf(v) * inv_sqrt_bc2 + eps));
    }
}

#define LAUNCH_1D(kernel, n, st, ...)                                    \
    do {                                                                 \
        kernel<<<grid_for(n), 256, 0, st>>>(__VA_ARGS__);                \
        count_launch();                                                  \
        return check_launch(#kernel);                                    \
    } while (0)

extern "C" int npf_merge_relu_fwd(const float* x1, const float* x2, float* out, int Z, int B, int T, int C, int x2_has_t,
                                  npf_stream_t stream) {
    NPF_REQUIRE(x1 && x2 && out, "npf_merge_relu_fwd: null pointer");
    NPF_REQUIRE(Z >= 1 && B >= 0 && T >= 0 && C >= 1, "npf_merge_relu_fwd: bad shape");
    const long n = (long)Z * B * T * C;
    if (n == 0) return NPF_OK;
    LAUNCH_1D(merge_relu_fwd_kernel, n, as_stream(stream), x1, x2, out, Z, B, T, C, x2_has_t);
}

extern "C" int npf_merge_relu_bwd(const float* dout, const float* out, float* dx1, float* dx2, int Z, int B, int T, int C,
                                  int x2_has_t, npf_stream_t stream) {
    NPF_REQUIRE(dout && out, "npf_merge_relu_bwd: null pointer");
    NPF_REQUIRE(Z >= 1 && B >= 0 && T >= 0 && C >= 1, "npf_merge_relu_bwd: bad shape");
    cudaStream_t st = as_stream(stream);
    const long btc = (long)B * T * C;
    if ((long)Z * B * C == 0) return NPF_OK;
    if (dx1 && btc > 0) {
        merge_relu_bwd_x1_kernel<<<grid_for(btc), 256, 0, st>>>(dout, out, dx1, Z, btc);
        count_launch();
        int rc = check_launch("merge_relu_bwd_x1_kernel");
        if (rc != NPF_OK) return rc;
    }
    if (dx2) {
        const long n = x2_has_t ? (long)Z * btc : (long)Z * B * C;
        if (n == 0) return NPF_OK;
        merge_relu_bwd_x2_kernel<<<grid_for(n), 256, 0, st>>>(dout, out, dx2, (long)Z * B, T, C, x2_has_t);
        count_launch();
        return check_launch("merge_relu_bwd_x2_kernel");
    }
    return NPF_OK;
}

extern "C" int npf_mean_pool_fwd(const float* X, float* R, int B, int N, int C, npf_stream_t stream) {
    NPF_REQUIRE(X && R, "npf_mean_pool_fwd: null pointer");
    NPF_REQUIRE(B >= 0 && N >= 1 && C >= 1, "npf_mean_pool_fwd: bad shape (N must be >= 1)");
    if (B == 0) return NPF_OK;
    LAUNCH_1D(mean_pool_fwd_kernel, (long)B * C, as_stream(stream), X, R, B, N, C);
}
extern "C" int npf_mean_pool_bwd(const float* dR, float* dX, int B, int N, int C, npf_stream_t stream) {
    NPF_REQUIRE(dR && dX, "npf_mean_pool_bwd: null pointer");
    NPF_REQUIRE(B >= 0 && N >= 1 && C >= 1, "npf_mean_pool_bwd: bad shape");
    if (B == 0) return NPF_OK;
    LAUNCH_1D(mean_pool_bwd_kernel, (long)B * N * C, as_stream(stream), dR, dX, B, N, C);
}

extern "C" int npf_add_layernorm_fwd(const float* A, const float* Bm, const float* gamma, const float* beta, float* Y,
                                     float* rstat, long M, int C, npf_stream_t stream) {
    NPF_REQUIRE(A && Bm && gamma && beta && Y && rstat, "npf_add_layernorm_fwd: null pointer");
    NPF_REQUIRE(M >= 0 && C >= 1 && C <= 32 * kLnChunks, "npf_add_layernorm_fwd: C must be in [1, %d]", 32 * kLnChunks);
    if (M == 0) return NPF_OK;
    LAUNCH_1D(add_layernorm_fwd_kernel, M * 32, as_stream(stream), A, Bm, gamma, beta, Y, rstat, M, C);
}
extern "C" int npf_add_layernorm_bwd(const float* dY, const float* A, const float* Bm, const float* gamma,
                                     const float* rstat, float* dS, float* dgamma, float* dbeta, long M, int C,
                                     npf_stream_t stream) {
    NPF_REQUIRE(dY && A && Bm && gamma && rstat && dS && dgamma && dbeta, "npf_add_layernorm_bwd: null pointer");
    NPF_REQUIRE(M >= 0 && C >= 1 && C <= 32 * kLnChunks, "npf_add_layernorm_bwd: bad shape");
    if (M == 0) return NPF_OK;
    long blocks = cdiv(M * 32, 256);
    if (blocks > 4L * kNumSMs) blocks = 4L * kNumSMs;
    add_layernorm_bwd_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(dY, A, Bm, gamma, rstat, dS, dgamma, dbeta, M, C);
    count_launch();
    return check_launch("add_layernorm_bwd_kernel");
}

extern "C" int npf_gauss_head_fwd(const float* suff, float* loc, float* scale, long M, int y, float min_scale,
                                  npf_stream_t stream) {
    NPF_REQUIRE(suff && loc && scale, "npf_gauss_head_fwd: null pointer");
    NPF_REQUIRE(M >= 0 && y >= 1, "npf_gauss_head_fwd: bad shape");
    if (M == 0) return NPF_OK;
    LAUNCH_1D(gauss_head_fwd_kernel, M * y, as_stream(stream), suff, loc, scale, M, y, min_scale);
}
extern "C" int npf_gauss_head_bwd(const float* suff, const float* dloc, const float* dscale, float* dsuff, long M, int y,
                                  float min_scale, npf_stream_t stream) {
    NPF_REQUIRE(suff && dsuff, "npf_gauss_head_bwd: null pointer");
    NPF_REQUIRE(M >= 0 && y >= 1, "npf_gauss_head_bwd: bad shape");
    if (M == 0) return NPF_OK;
    LAUNCH_1D(gauss_head_bwd_kernel, M * y, as_stream(stream), suff, dloc, dscale, dsuff, M, y, min_scale);
}

extern "C" int npf_gauss_nll_fwd(const float* loc, const float* scale, const float* Y, float* slp, int Z, int B, long n,
                                 npf_stream_t stream) {
    NPF_REQUIRE(loc && scale && Y && slp, "npf_gauss_nll_fwd: null pointer");
    NPF_REQUIRE(Z >= 1 && B >= 0 && n >= 0, "npf_gauss_nll_fwd: bad shape");
    if (B == 0) return NPF_OK;
    gauss_nll_fwd_kernel<<<(unsigned)((long)Z * B), 256, 0, as_stream(stream)>>>(loc, scale, Y, slp, B, n);
    count_launch();
    return check_launch("gauss_nll_fwd_kernel");
}
extern "C" int npf_gauss_nll_bwd(const float* loc, const float* scale, const float* Y, const float* g, float* dloc,
                                 float* dscale, int Z, int B, long n, npf_stream_t stream) {
    NPF_REQUIRE(loc && scale && Y && g && dloc && dscale, "npf_gauss_nll_bwd: null pointer");
    NPF_REQUIRE(Z >= 1 && B >= 0 && n >= 0, "npf_gauss_nll_bwd: bad shape");
    const long total = (long)Z * B * n;
    if (total == 0) return NPF_OK;
    LAUNCH_1D(gauss_nll_bwd_kernel, total, as_stream(stream), loc, scale, Y, g, dloc, dscale, B, n, total);
}

extern "C" int npf_latent_sample_fwd(const float* suff, const float* eps, float* q_loc, float* q_scale, float* z, int S,
                                     long M, int zd, npf_stream_t stream) {
    NPF_REQUIRE(suff && q_loc && q_scale && (S == 0 || (eps && z)), "npf_latent_sample_fwd: null pointer");
    NPF_REQUIRE(S >= 0 && M >= 0 && zd >= 1, "npf_latent_sample_fwd: bad shape");
    if (M == 0) return NPF_OK;
    LAUNCH_1D(latent_sample_fwd_kernel, M * zd, as_stream(stream), suff, eps, q_loc, q_scale, z, S, M, zd);
}
extern "C" int npf_latent_sample_bwd(const float* suff, const float* eps, const float* dz, const float* dq_loc,
                                     const float* dq_scale, float* dsuff, int S, long M, int zd, npf_stream_t stream) {
    NPF_REQUIRE(suff && dsuff && (S == 0 || !dz || eps), "npf_latent_sample_bwd: null pointer");
    NPF_REQUIRE(S >= 0 && M >= 0 && zd >= 1, "npf_latent_sample_bwd: bad shape");
    if (M == 0) return NPF_OK;
    LAUNCH_1D(latent_sample_bwd_kernel, M * zd, as_stream(stream), suff, eps, dz, dq_loc, dq_scale, dsuff, S, M, zd);
}

extern "C" int npf_global_latent_fwd(const float* zin, float* out, int N, int P, int C, npf_stream_t stream) {
    NPF_REQUIRE(zin && out, "npf_global_latent_fwd: null pointer");
    NPF_REQUIRE(N >= 0 && P >= 1 && C >= 2 && C % 2 == 0, "npf_global_latent_fwd: bad shape");
    if (N == 0) return NPF_OK;
    global_latent_kernel<<<(unsigned)N, 128, 0, as_stream(stream)>>>(zin, out, P, C);
    count_launch();
    return check_launch("global_latent_kernel");
}
extern "C" int npf_global_latent_bwd(const float* dout, float* dzin, int N, int P, int C, npf_stream_t stream) {
    // the map is linear and self-adjoint: identity on the first half, mean-and-broadcast on the second
    return npf_global_latent_fwd(dout, dzin, N, P, C, stream);
}

extern "C" int npf_range_check(const float* X, long n, float lo, float hi, int* flag, npf_stream_t stream) {
    NPF_REQUIRE(X && flag, "npf_range_check: null pointer");
    if (n <= 0) return NPF_OK;
    LAUNCH_1D(range_check_kernel, n, as_stream(stream), X, n, lo, hi, flag);
}

extern "C" int npf_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, long n, int step, float lr, float beta1,
                             float beta2, float eps, float weight_decay, float grad_scale, npf_stream_t stream) {
    NPF_REQUIRE(param && grad && exp_avg && exp_avg_sq, "npf_adam_step: null pointer");
    NPF_REQUIRE(n >= 0 && step >= 1 && lr >= 0.f && beta1 >= 0.f && beta1 < 1.f && beta2 >= 0.f && beta2 < 1.f, "npf_adam_step: bad hyper-parameter");
    if (n == 0) return NPF_OK;
    const double bc1 = 1.0 - pow((double)beta1, (double)step), bc2 = 1.0 - pow((double)beta2, (double)step);
    LAUNCH_1D(adam_kernel, n, as_stream(stream), param, grad, exp_avg, exp_avg_sq, n, (float)(lr / bc1), (float)(1.0 / sqrt(bc2)), beta1, beta2, eps,
              weight_decay, grad_scale);
}

extern "C" int npf_sqnorm(const float* x, long n, float* sqnorm, npf_stream_t stream) {
    NPF_REQUIRE(sqnorm && (x || n == 0), "npf_sqnorm: null pointer");
    NPF_REQUIRE(n >= 0, "npf_sqnorm: bad size");
    cudaStream_t st = as_stream(stream);
    if (cudaMemsetAsync(sqnorm, 0, sizeof(float), st) != cudaSuccess) return check_launch("npf_sqnorm memset");
    if (n == 0) return NPF_OK;
    LAUNCH_1D(sqnorm_kernel, n, st, x, n, sqnorm);
}

extern "C" int npf_adam_step_clipped(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, long n, int step, float lr,
                                     float beta1, float beta2, float eps, float weight_decay, float grad_scale, const float* sqnorm,
                                     float max_norm, npf_stream_t stream) {
    NPF_REQUIRE(param && grad && exp_avg && exp_avg_sq && sqnorm, "npf_adam_step_clipped: null pointer");
    NPF_REQUIRE(n >= 0 && step >= 1 && lr >= 0.f && beta1 >= 0.f && beta1 < 1.f && beta2 >= 0.f && beta2 < 1.f && max_norm > 0.f,
                "npf_adam_step_clipped: bad hyper-parameter");
    if (n == 0) return NPF_OK;
    const double bc1 = 1.0 - pow((double)beta1, (double)step), bc2 = 1.0 - pow((double)beta2, (double)step);
    LAUNCH_1D(adam_clipped_kernel, n, as_stream(stream), param, grad, exp_avg, exp_avg_sq, n, (float)(lr / bc1), (float)(1.0 / sqrt(bc2)), beta1,
              beta2, eps, weight_decay, grad_scale, sqnorm, max_norm);
}
