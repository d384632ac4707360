// Multi-head scaled-dot cross-attention of targets over context, flash-style: the [Tq,Tk] logits never reach
// memory (upstream materialises them: npf/architectures/attention.py:129-161, 204-220; head split :507-527).
//
// This file is the fp32 (NPF_PREC_FP32, 1e-4 parity) path: one warp owns one query (forward, dQ) or one key
// (dK/dV); K/V (resp. Q/dO) tiles are staged in shared memory and shared by the 8 warps of the CTA; the softmax is
// online (running max / sum) so every pass is a single sweep.  Head h owns channels [h*D, (h+1)*D): no permute.
// The bf16 tensor-core path is in attention_tc.cu.
#include "common.cuh"

namespace npf {

constexpr int KT = 64;        // keys (or queries) per shared tile
constexpr int kMaxD = 128;    // head dim limit of this path

struct AttnParams {
    const float* Q; const float* K; const float* V; const float* O; const float* LSE; const float* dO;
    float* Oo; float* LSEo; float* dQ; float* dK; float* dV;
    int Tq, Tk, H, D, Dv;
    float scale;
};

// smem layout helper: rows padded by +1 float to keep column walks conflict-free
__device__ __forceinline__ void load_tile(float* S, int ld, const float* __restrict__ G, long g_ld, int row0, int nrows_total,
                                          int width, int h_off) {
    // S[r][c] = G[(row0 + r) * g_ld + h_off + c] for r < KT, zero beyond nrows_total
    for (int idx = threadIdx.x; idx < KT * width; idx += blockDim.x) {
        const int r = idx / width, c = idx % width;
        const int gr = row0 + r;
        S[r * ld + c] = (gr < nrows_total) ? __ldg(G + (long)gr * g_ld + h_off + c) : 0.f;
    }
}

__global__ void __launch_bounds__(256) xattn_fwd_kernel(AttnParams p) {
    extern __shared__ float smem[];
    const int D = p.D, Dv = p.Dv, ldk = D + 1;
    float* Ks = smem;                 // [KT][D+1]
    float* Vs = Ks + KT * ldk;        // [KT][Dv]
    float* Qs = Vs + KT * Dv;         // [8][D]
    float* Ps = Qs + 8 * D;           // [8][KT]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int h = blockIdx.y, b = blockIdx.z;
    const int q = blockIdx.x * 8 + warp;
    const bool active = q < p.Tq;
    const long ldq = (long)p.H * D, ldv = (long)p.H * Dv;
    const float* Qb = p.Q + (long)b * p.Tq * ldq;
    const float* Kb = p.K + (long)b * p.Tk * ldq;
    const float* Vb = p.V + (long)b * p.Tk * ldv;

    for (int d = lane; d < D; d += 32) Qs[warp * D + d] = active ? __ldg(Qb + (long)q * ldq + h * D + d) : 0.f;

    float m = -INFINITY, l = 0.f;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int k0 = 0; k0 < p.Tk; k0 += KT) {
        __syncthreads();
        load_tile(Ks, ldk, Kb, ldq, k0, p.Tk, D, h * D);
        load_tile(Vs, Dv, Vb, ldv, k0, p.Tk, Dv, h * Dv);
        __syncthreads();
        float s0 = 0.f, s1 = 0.f;
        const float* qv = Qs + warp * D;
        const float* ka = Ks + lane * ldk;
        const float* kb = Ks + (lane + 32) * ldk;
        for (int d = 0; d < D; ++d) {
            const float x = qv[d];
            s0 = fmaf(x, ka[d], s0);
            s1 = fmaf(x, kb[d], s1);
        }
        s0 = (k0 + lane < p.Tk) ? s0 * p.scale : -INFINITY;
        s1 = (k0 + lane + 32 < p.Tk) ? s1 * p.scale : -INFINITY;
        const float m_new = fmaxf(m, warp_max(fmaxf(s0, s1)));
        const float alpha = expf(m - m_new);  // m = -inf on the first tile -> 0
        const float p0 = expf(s0 - m_new), p1 = expf(s1 - m_new);
        l = l * alpha + warp_sum(p0 + p1);
        Ps[warp * KT + lane] = p0;
        Ps[warp * KT + lane + 32] = p1;
        __syncwarp();
        const float* pr = Ps + warp * KT;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int c = lane + 32 * i;
            if (c < Dv) {
                float a = acc[i] * alpha;
                for (int kk = 0; kk < KT; ++kk) a = fmaf(pr[kk], Vs[kk * Dv + c], a);
                acc[i] = a;
            }
        }
        m = m_new;
        __syncwarp();
    }
    if (active) {
        const float inv = 1.f / l;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int c = lane + 32 * i;
            if (c < Dv) p.Oo[((long)b * p.Tq + q) * ldv + h * Dv + c] = acc[i] * inv;
        }
        if (lane == 0) p.LSEo[((long)b * p.H + h) * p.Tq + q] = m + logf(l);
    }
}

// dQ: one warp per query.  dS = P (.) (dP - Di) * scale ; dQ = dS K
__global__ void __launch_bounds__(256) xattn_bwd_dq_kernel(AttnParams p) {
    extern __shared__ float smem[];
    const int D = p.D, Dv = p.Dv, ldk = D + 1, ldvv = Dv + 1;
    float* Ks = smem;                  // [KT][D+1]
    float* Vs = Ks + KT * ldk;         // [KT][Dv+1]
    float* Qs = Vs + KT * ldvv;        // [8][D]
    float* Gs = Qs + 8 * D;            // [8][Dv]   dO rows
    float* Os = Gs + 8 * Dv;           // [8][Dv]   O rows
    float* Ps = Os + 8 * Dv;           // [8][KT]   dS
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int h = blockIdx.y, b = blockIdx.z;
    const int q = blockIdx.x * 8 + warp;
    const bool active = q < p.Tq;
    const long ldq = (long)p.H * D, ldv = (long)p.H * Dv;
    const float* Kb = p.K + (long)b * p.Tk * ldq;
    const float* Vb = p.V + (long)b * p.Tk * ldv;
    const long qrow = (long)b * p.Tq + q;

    for (int d = lane; d < D; d += 32) Qs[warp * D + d] = active ? __ldg(p.Q + qrow * ldq + h * D + d) : 0.f;
    for (int c = lane; c < Dv; c += 32) {
        Gs[warp * Dv + c] = active ? __ldg(p.dO + qrow * ldv + h * Dv + c) : 0.f;
        Os[warp * Dv + c] = active ? __ldg(p.O + qrow * ldv + h * Dv + c) : 0.f;
    }
    __syncwarp();
    // Di = dO . O accumulated in the SAME order as dP = dO . V below, so that for a single key (O == V) the
    // softmax gradient P (dP - Di) cancels exactly, as it does in a materialised softmax backward
    float di = 0.f;
    for (int c = 0; c < Dv; ++c) di = fmaf(Gs[warp * Dv + c], Os[warp * Dv + c], di);
    const float lse = active ? __ldg(p.LSE + ((long)b * p.H + h) * p.Tq + q) : 0.f;

    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int k0 = 0; k0 < p.Tk; k0 += KT) {
        __syncthreads();
        load_tile(Ks, ldk, Kb, ldq, k0, p.Tk, D, h * D);
        load_tile(Vs, ldvv, Vb, ldv, k0, p.Tk, Dv, h * Dv);
        __syncthreads();
        float s0 = 0.f, s1 = 0.f, g0 = 0.f, g1 = 0.f;
        const float* qv = Qs + warp * D;
        const float* gv = Gs + warp * Dv;
        for (int d = 0; d < D; ++d) {
            const float x = qv[d];
            s0 = fmaf(x, Ks[lane * ldk + d], s0);
            s1 = fmaf(x, Ks[(lane + 32) * ldk + d], s1);
        }
        for (int c = 0; c < Dv; ++c) {
            const float x = gv[c];
            g0 = fmaf(x, Vs[lane * ldvv + c], g0);
            g1 = fmaf(x, Vs[(lane + 32) * ldvv + c], g1);
        }
        const float p0 = (k0 + lane < p.Tk) ? expf(s0 * p.scale - lse) : 0.f;
        const float p1 = (k0 + lane + 32 < p.Tk) ? expf(s1 * p.scale - lse) : 0.f;
        Ps[warp * KT + lane] = p0 * (g0 - di) * p.scale;
        Ps[warp * KT + lane + 32] = p1 * (g1 - di) * p.scale;
        __syncwarp();
        const float* pr = Ps + warp * KT;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int d = lane + 32 * i;
            if (d < D) {
                float a = acc[i];
                for (int kk = 0; kk < KT; ++kk) a = fmaf(pr[kk], Ks[kk * ldk + d], a);
                acc[i] = a;
            }
        }
        __syncwarp();
    }
    if (active) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int d = lane + 32 * i;
            if (d < D) p.dQ[qrow * ldq + h * D + d] = acc[i];
        }
    }
}

// dK, dV: one warp per key; sweeps query tiles.  dV = P^T dO ; dK = dS^T Q
__global__ void __launch_bounds__(256) xattn_bwd_dkv_kernel(AttnParams p) {
    extern __shared__ float smem[];
    const int D = p.D, Dv = p.Dv, ldq_s = D + 1, ldg_s = Dv + 1;
    float* Qs = smem;                    // [KT][D+1]
    float* Gs = Qs + KT * ldq_s;         // [KT][Dv+1]  dO
    float* Ls = Gs + KT * ldg_s;         // [KT] lse
    float* Ds = Ls + KT;                 // [KT] Di
    float* Kr = Ds + KT;                 // [8][D]
    float* Vr = Kr + 8 * D;              // [8][Dv]
    float* Ps = Vr + 8 * Dv;             // [8][KT]  P
    float* Ss = Ps + 8 * KT;             // [8][KT]  dS
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int h = blockIdx.y, b = blockIdx.z;
    const int k = blockIdx.x * 8 + warp;
    const bool active = k < p.Tk;
    const long ldq = (long)p.H * D, ldv = (long)p.H * Dv;
    const float* Qb = p.Q + (long)b * p.Tq * ldq;
    const float* Gb = p.dO + (long)b * p.Tq * ldv;
    const float* Ob = p.O + (long)b * p.Tq * ldv;
    const long krow = (long)b * p.Tk + k;

    for (int d = lane; d < D; d += 32) Kr[warp * D + d] = active ? __ldg(p.K + krow * ldq + h * D + d) : 0.f;
    for (int c = lane; c < Dv; c += 32) Vr[warp * Dv + c] = active ? __ldg(p.V + krow * ldv + h * Dv + c) : 0.f;

    float accK[4] = {0.f, 0.f, 0.f, 0.f}, accV[4] = {0.f, 0.f, 0.f, 0.f};
    for (int q0 = 0; q0 < p.Tq; q0 += KT) {
        __syncthreads();
        load_tile(Qs, ldq_s, Qb, ldq, q0, p.Tq, D, h * D);
        load_tile(Gs, ldg_s, Gb, ldv, q0, p.Tq, Dv, h * Dv);
        __syncthreads();
        // per-query row statistics of this tile: lse and Di = dO . O, summed sequentially over channels (same order
        // as dP = V . dO below; see the dQ kernel)
        if (threadIdx.x < KT) {
            const int r = threadIdx.x, q = q0 + r;
            float di = 0.f;
            if (q < p.Tq)
                for (int c = 0; c < Dv; ++c) di = fmaf(Gs[r * ldg_s + c], __ldg(Ob + (long)q * ldv + h * Dv + c), di);
            Ds[r] = di;
            Ls[r] = (q < p.Tq) ? __ldg(p.LSE + ((long)b * p.H + h) * p.Tq + q) : 0.f;
        }
        __syncthreads();
        float s0 = 0.f, s1 = 0.f, g0 = 0.f, g1 = 0.f;
        const float* kv = Kr + warp * D;
        const float* vv = Vr + warp * Dv;
        for (int d = 0; d < D; ++d) {
            const float x = kv[d];
            s0 = fmaf(x, Qs[lane * ldq_s + d], s0);
            s1 = fmaf(x, Qs[(lane + 32) * ldq_s + d], s1);
        }
        for (int c = 0; c < Dv; ++c) {
            const float x = vv[c];
            g0 = fmaf(x, Gs[lane * ldg_s + c], g0);
            g1 = fmaf(x, Gs[(lane + 32) * ldg_s + c], g1);
        }
        const float p0 = (q0 + lane < p.Tq) ? expf(s0 * p.scale - Ls[lane]) : 0.f;
        const float p1 = (q0 + lane + 32 < p.Tq) ? expf(s1 * p.scale - Ls[lane + 32]) : 0.f;
        Ps[warp * KT + lane] = p0;
        Ps[warp * KT + lane + 32] = p1;
        Ss[warp * KT + lane] = p0 * (g0 - Ds[lane]) * p.scale;
        Ss[warp * KT + lane + 32] = p1 * (g1 - Ds[lane + 32]) * p.scale;
        __syncwarp();
        const float* pr = Ps + warp * KT;
        const float* sr = Ss + warp * KT;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int c = lane + 32 * i;
            if (c < Dv) {
                float a = accV[i];
                for (int qq = 0; qq < KT; ++qq) a = fmaf(pr[qq], Gs[qq * ldg_s + c], a);
                accV[i] = a;
            }
            if (c < D) {
                float a = accK[i];
                for (int qq = 0; qq < KT; ++qq) a = fmaf(sr[qq], Qs[qq * ldq_s + c], a);
                accK[i] = a;
            }
        }
        __syncwarp();
    }
    if (active) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int c = lane + 32 * i;
            if (c < Dv) p.dV[krow * ldv + h * Dv + c] = accV[i];
            if (c < D) p.dK[krow * ldq + h * D + c] = accK[i];
        }
    }
}

// attention_tc.cu: tcgen05 path (head dim 16 / 32); NPF_ENOTSUP for other shapes
int xattn_fwd_tc(const float* Q, const float* K, const float* V, float* O, float* LSE, int B, int Tq, int Tk, int H, int D, int Dv,
                 float scale, int precision, cudaStream_t st);
int xattn_bwd_tc(const float* Q, const float* K, const float* V, const float* O, const float* LSE, const float* dO, float* dQ, float* dK,
                 float* dV, int B, int Tq, int Tk, int H, int D, int Dv, float scale, int precision, cudaStream_t st);

static int set_smem(const void* fn, size_t bytes) {
    if (bytes > 48 * 1024) {
        if (cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes) != cudaSuccess) {
            set_error("xattn: cannot reserve %zu B of shared memory", bytes);
            cudaGetLastError();
            return NPF_ECUDA;
        }
    }
    return NPF_OK;
}

}  // namespace npf

using namespace npf;

extern "C" int npf_xattn_fwd(const float* Q, const float* K, const float* V, float* O, float* LSE, int B, int Tq, int Tk,
                             int H, int D, int Dv, float scale, int precision, npf_stream_t stream) {
    NPF_REQUIRE(Q && K && V && O && LSE, "npf_xattn_fwd: null pointer");
    NPF_REQUIRE(B >= 0 && Tq >= 0 && Tk >= 1 && H >= 1 && D >= 1 && Dv >= 1, "npf_xattn_fwd: bad shape (Tk must be >= 1)");
    NPF_REQUIRE(D <= kMaxD && Dv <= kMaxD, "npf_xattn_fwd: head dim > %d", kMaxD);
    NPF_REQUIRE(B <= 65535 && H <= 65535, "npf_xattn_fwd: batch/heads > 65535");
    if (B == 0 || Tq == 0) return NPF_OK;
    cudaStream_t st = as_stream(stream);
    if (precision != NPF_PREC_FP32) {
        int rc = xattn_fwd_tc(Q, K, V, O, LSE, B, Tq, Tk, H, D, Dv, scale, precision, st);
        if (rc != NPF_ENOTSUP) return rc;
    }
    AttnParams p{};
    p.Q = Q; p.K = K; p.V = V; p.Oo = O; p.LSEo = LSE;
    p.Tq = Tq; p.Tk = Tk; p.H = H; p.D = D; p.Dv = Dv; p.scale = scale;
    const size_t smem = sizeof(float) * ((size_t)KT * (D + 1) + (size_t)KT * Dv + 8 * D + 8 * KT);
    int rc = set_smem((const void*)xattn_fwd_kernel, smem);
    if (rc != NPF_OK) return rc;
    dim3 grid((unsigned)cdiv(Tq, 8), (unsigned)H, (unsigned)B);
    xattn_fwd_kernel<<<grid, 256, smem, st>>>(p);
    count_launch();
    return check_launch("xattn_fwd_kernel");
}

extern "C" int npf_xattn_bwd(const float* Q, const float* K, const float* V, const float* O, const float* LSE,
                             const float* dO, float* dQ, float* dK, float* dV, int B, int Tq, int Tk, int H, int D, int Dv,
                             float scale, int precision, npf_stream_t stream) {
    NPF_REQUIRE(Q && K && V && O && LSE && dO && dQ && dK && dV, "npf_xattn_bwd: null pointer");
    NPF_REQUIRE(B >= 0 && Tq >= 0 && Tk >= 1 && H >= 1 && D >= 1 && Dv >= 1, "npf_xattn_bwd: bad shape");
    NPF_REQUIRE(D <= kMaxD && Dv <= kMaxD, "npf_xattn_bwd: head dim > %d", kMaxD);
    NPF_REQUIRE(B <= 65535 && H <= 65535, "npf_xattn_bwd: batch/heads > 65535");
    if (B == 0) return NPF_OK;
    cudaStream_t st = as_stream(stream);
    if (Tq == 0) {
        cudaMemsetAsync(dK, 0, sizeof(float) * (size_t)B * Tk * H * D, st);
        cudaMemsetAsync(dV, 0, sizeof(float) * (size_t)B * Tk * H * Dv, st);
        return NPF_OK;
    }
    if (precision != NPF_PREC_FP32) {
        int rc = xattn_bwd_tc(Q, K, V, O, LSE, dO, dQ, dK, dV, B, Tq, Tk, H, D, Dv, scale, precision, st);
        if (rc != NPF_ENOTSUP) return rc;
    }
    AttnParams p{};
    p.Q = Q; p.K = K; p.V = V; p.O = O; p.LSE = LSE; p.dO = dO; p.dQ = dQ; p.dK = dK; p.dV = dV;
    p.Tq = Tq; p.Tk = Tk; p.H = H; p.D = D; p.Dv = Dv; p.scale = scale;
    {
        const size_t smem = sizeof(float) * ((size_t)KT * (D + 1) + (size_t)KT * (Dv + 1) + 8 * D + 16 * Dv + 8 * KT);
        int rc = set_smem((const void*)xattn_bwd_dq_kernel, smem);
        if (rc != NPF_OK) return rc;
        dim3 grid((unsigned)cdiv(Tq, 8), (unsigned)H, (unsigned)B);
        xattn_bwd_dq_kernel<<<grid, 256, smem, st>>>(p);
        count_launch();
        rc = check_launch("xattn_bwd_dq_kernel");
        if (rc != NPF_OK) return rc;
    }
    {
        const size_t smem = sizeof(float) * ((size_t)KT * (D + 1) + (size_t)KT * (Dv + 1) + 2 * KT + 8 * D + 8 * Dv + 16 * KT);
        int rc = set_smem((const void*)xattn_bwd_dkv_kernel, smem);
        if (rc != NPF_OK) return rc;
        dim3 grid((unsigned)cdiv(Tk, 8), (unsigned)H, (unsigned)B);
        xattn_bwd_dkv_kernel<<<grid, 256, smem, st>>>(p);
        count_launch();
        return check_launch("xattn_bwd_dkv_kernel");
    }
}
