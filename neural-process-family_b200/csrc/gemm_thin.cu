// "Thin" linear layers: one side of the weight matrix has <= 8 entries (x-encoder input layer K = x_dim, y-resizer
// K = y_dim, SetConv's Linear(y+1 -> r) and its density column, the 2y-wide predictive head).  A 128x128 GEMM tile wastes
// 16-128x of its work on these; they are pure streaming problems, so each gets a one-pass HBM-bound kernel.
//
//   thin_red   : out[m, o] = epi( sum_{r < R} A[m, r] * B(o, r) )          R <= 8, O wide   (elementwise over [M, O])
//   rowdot     : out[m, j] = epi( sum_{i < I} A[m, i] * B(j, i) )          J <= 8, I wide   (one warp per row)
//   thin_outer : out(j, c) += sum_m S[m, j] * T[m, c]                       J <= 9, C wide   (column reduction over rows)
#include "common.cuh"
#include "gemm_thin.cuh"

namespace npf {

constexpr int kThin = 8;


__global__ void __launch_bounds__(256) thin_red_kernel(ThinRedParams p) {
    extern __shared__ float sB[];            // [R][O] + bias[O] + w2[O]
    float* sbias = sB + p.R * p.O;
    float* sw2 = sbias + p.O;
    for (int i = threadIdx.x; i < p.R * p.O; i += blockDim.x) {
        const int r = i / p.O, o = i % p.O;
        sB[i] = __ldg(p.B + (long)o * p.sb_o + (long)r * p.sb_r);
    }
    for (int o = threadIdx.x; o < p.O; o += blockDim.x) {
        sbias[o] = p.bias ? __ldg(p.bias + o) : 0.f;
        sw2[o] = p.w2 ? __ldg(p.w2 + (long)o * p.ldw2) : 0.f;
    }
    __syncthreads();
    const int oq = (p.O + 3) >> 2;            // groups of 4 outputs
    const long total = p.M * oq;
    for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
        const long m = idx / oq;
        const int o0 = (int)(idx % oq) * 4;
        float a[kThin];
#pragma unroll
        for (int r = 0; r < kThin; ++r) {
            a[r] = (r < p.R) ? __ldg(p.A + m * p.lda + r) : 0.f;
            if (p.relu_a) a[r] = fmaxf(a[r], 0.f);
        }
        const float um = p.u ? __ldg(p.u + m) : 0.f;
        float v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int o = o0 + j;
            float x = 0.f;
            if (o < p.O) {
#pragma unroll
                for (int r = 0; r < kThin; ++r)
                    if (r < p.R) x = fmaf(a[r], sB[r * p.O + o], x);
                x += sbias[o];
                if (p.u) x = fmaf(um, sw2[o], x);
                if (p.relu_out) x = fmaxf(x, 0.f);
                if (p.mask) x = (__ldg(p.mask + m * p.ldm + o) > 0.f) ? x : 0.f;
            }
            v[j] = x;
        }
        float* out = p.out + m * p.ldo + o0;
        if (o0 + 4 <= p.O && (p.ldo & 3) == 0 && ((reinterpret_cast<uintptr_t>(p.out) & 15) == 0)) {
            float4 w = make_float4(v[0], v[1], v[2], v[3]);
            if (p.accum) { const float4 old = *reinterpret_cast<const float4*>(out); w.x += old.x; w.y += old.y; w.z += old.z; w.w += old.w; }
            *reinterpret_cast<float4*>(out) = w;
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (o0 + j < p.O) out[j] = p.accum ? out[j] + v[j] : v[j];
        }
    }
}


__global__ void __launch_bounds__(256) rowdot_kernel(RowDotParams p) {
    extern __shared__ float sB[];            // [J][I]
    for (int i = threadIdx.x; i < p.J * p.I; i += blockDim.x) {
        const int j = i / p.I, ii = i % p.I;
        sB[i] = __ldg(p.B + (long)j * p.sb_j + (long)ii * p.sb_i);
    }
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const long warp = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long nwarps = ((long)gridDim.x * blockDim.x) >> 5;
    constexpr int RU = 4;                    // rows per warp pass: RU independent row loads in flight
    for (long m0 = warp * RU; m0 < p.M; m0 += nwarps * RU) {
        float acc[RU][kThin];
#pragma unroll
        for (int u = 0; u < RU; ++u)
#pragma unroll
            for (int j = 0; j < kThin; ++j) acc[u][j] = 0.f;
        for (int i = lane; i < p.I; i += 32) {
            float a[RU];
#pragma unroll
            for (int u = 0; u < RU; ++u) {
                a[u] = (m0 + u < p.M) ? __ldg(p.A + (m0 + u) * p.lda + i) : 0.f;
                if (p.relu_a) a[u] = fmaxf(a[u], 0.f);
            }
#pragma unroll
            for (int j = 0; j < kThin; ++j)
                if (j < p.J) {
                    const float w = sB[j * p.I + i];
#pragma unroll
                    for (int u = 0; u < RU; ++u) acc[u][j] = fmaf(a[u], w, acc[u][j]);
                }
        }
#pragma unroll
        for (int u = 0; u < RU; ++u) {
#pragma unroll
            for (int j = 0; j < kThin; ++j)
                if (j < p.J) acc[u][j] = warp_sum(acc[u][j]);
            const long m = m0 + u;
            if (lane < p.J && m < p.M) {
                float x = 0.f;
#pragma unroll
                for (int j = 0; j < kThin; ++j)
                    if (j == lane) x = acc[u][j];
                if (p.bias) x += __ldg(p.bias + lane);
                if (p.relu_out) x = fmaxf(x, 0.f);
                if (p.mask) x = (__ldg(p.mask + m * p.ldm + lane) > 0.f) ? x : 0.f;
                float* o = p.out + m * p.ldo + lane;
                *o = p.accum ? *o + x : x;
            }
        }
    }
}


// float4 columns, 4 rows in flight per thread; block-level reduction through shared-memory atomics
__global__ void __launch_bounds__(256) thin_outer_vec_kernel(ThinOuterParams p) {
    extern __shared__ float sacc[];          // [(J + 2)][C]
    const int nacc = (p.J + 2) * p.C;
    for (int i = threadIdx.x; i < nacc; i += blockDim.x) sacc[i] = 0.f;
    __syncthreads();
    const int CQ = p.C >> 2;
    const int q = threadIdx.x % CQ, ry = threadIdx.x / CQ, RY = blockDim.x / CQ;
    const long m0 = (long)blockIdx.x * p.rows_per_block, m1 = min(p.M, m0 + p.rows_per_block);
    float4 acc[kThin + 2];
#pragma unroll
    for (int j = 0; j < kThin + 2; ++j) acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    constexpr int RU = 4;
    for (long mb = m0 + ry * RU; mb < m1; mb += (long)RY * RU) {
        float4 t[RU];
#pragma unroll
        for (int u = 0; u < RU; ++u) {
            t[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (mb + u < m1) t[u] = __ldg(reinterpret_cast<const float4*>(p.T + (mb + u) * p.ldt + 4 * q));
            if (p.relu_t) { t[u].x = fmaxf(t[u].x, 0.f); t[u].y = fmaxf(t[u].y, 0.f); t[u].z = fmaxf(t[u].z, 0.f); t[u].w = fmaxf(t[u].w, 0.f); }
        }
#pragma unroll
        for (int u = 0; u < RU; ++u) {
            if (mb + u >= m1) continue;
#pragma unroll
            for (int j = 0; j < kThin; ++j)
                if (j < p.J) {
                    float s = __ldg(p.S + (mb + u) * p.lds + j);
                    if (p.relu_s) s = fmaxf(s, 0.f);
                    acc[j].x = fmaf(s, t[u].x, acc[j].x); acc[j].y = fmaf(s, t[u].y, acc[j].y);
                    acc[j].z = fmaf(s, t[u].z, acc[j].z); acc[j].w = fmaf(s, t[u].w, acc[j].w);
                }
            acc[kThin].x += t[u].x; acc[kThin].y += t[u].y; acc[kThin].z += t[u].z; acc[kThin].w += t[u].w;
            if (p.u) {
                const float uu = __ldg(p.u + mb + u);
                acc[kThin + 1].x = fmaf(uu, t[u].x, acc[kThin + 1].x); acc[kThin + 1].y = fmaf(uu, t[u].y, acc[kThin + 1].y);
                acc[kThin + 1].z = fmaf(uu, t[u].z, acc[kThin + 1].z); acc[kThin + 1].w = fmaf(uu, t[u].w, acc[kThin + 1].w);
            }
        }
    }
#pragma unroll
    for (int j = 0; j < kThin + 2; ++j) {
        const int slot = j < kThin ? j : p.J + (j - kThin);
        if (j < kThin && j >= p.J) continue;
        float* d = sacc + slot * p.C + 4 * q;
        atomicAdd(d + 0, acc[j].x); atomicAdd(d + 1, acc[j].y); atomicAdd(d + 2, acc[j].z); atomicAdd(d + 3, acc[j].w);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < nacc; i += blockDim.x) {
        const int slot = i / p.C, c = i % p.C;
        const float v = sacc[i];
        if (slot < p.J) atomicAdd(p.out + (long)slot * p.so_j + (long)c * p.so_c, v);
        else if (slot == p.J) { if (p.out_ones) atomicAdd(p.out_ones + c, v); }
        else if (p.out_u) atomicAdd(p.out_u + (long)c * p.so_u, v);
    }
}


__global__ void __launch_bounds__(256) thin_outer_kernel(ThinOuterParams p) {
    __shared__ float red[8][32][kThin + 2];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int c = blockIdx.y * 32 + tx;
    const long m0 = (long)blockIdx.x * p.rows_per_block, m1 = min(p.M, m0 + p.rows_per_block);
    float acc[kThin + 2];
#pragma unroll
    for (int j = 0; j < kThin + 2; ++j) acc[j] = 0.f;
    if (c < p.C) {
        for (long m = m0 + ty; m < m1; m += 8) {
            float t = __ldg(p.T + m * p.ldt + c);
            if (p.relu_t) t = fmaxf(t, 0.f);
#pragma unroll
            for (int j = 0; j < kThin; ++j)
                if (j < p.J) {
                    float s = __ldg(p.S + m * p.lds + j);
                    if (p.relu_s) s = fmaxf(s, 0.f);
                    acc[j] = fmaf(s, t, acc[j]);
                }
            acc[kThin] += t;
            if (p.u) acc[kThin + 1] = fmaf(__ldg(p.u + m), t, acc[kThin + 1]);
        }
    }
#pragma unroll
    for (int j = 0; j < kThin + 2; ++j) red[ty][tx][j] = acc[j];
    __syncthreads();
    if (ty == 0 && c < p.C) {
#pragma unroll
        for (int j = 0; j < kThin + 2; ++j) {
            float s = 0.f;
#pragma unroll
            for (int i = 0; i < 8; ++i) s += red[i][tx][j];
            if (j < p.J) atomicAdd(p.out + (long)j * p.so_j + (long)c * p.so_c, s);
            else if (j == kThin && p.out_ones) atomicAdd(p.out_ones + c, s);
            else if (j == kThin + 1 && p.out_u) atomicAdd(p.out_u + (long)c * p.so_u, s);
        }
    }
}

static inline unsigned stream_grid(long work_items) {
    long g = cdiv(work_items, 256);
    if (g > 64L * kNumSMs) g = 64L * kNumSMs;
    if (g < 1) g = 1;
    return (unsigned)g;
}

int thin_red(ThinRedParams& p, cudaStream_t st) {
    const size_t smem = sizeof(float) * ((size_t)p.R * p.O + 2 * (size_t)p.O);
    if (smem > 48 * 1024) return NPF_ENOTSUP;
    thin_red_kernel<<<stream_grid(p.M * ((p.O + 3) / 4)), 256, smem, st>>>(p);
    count_launch();
    return check_launch("thin_red_kernel");
}

int rowdot(RowDotParams& p, cudaStream_t st) {
    const size_t smem = sizeof(float) * (size_t)p.J * p.I;
    if (smem > 48 * 1024) return NPF_ENOTSUP;
    rowdot_kernel<<<stream_grid(p.M * 32), 256, smem, st>>>(p);
    count_launch();
    return check_launch("rowdot_kernel");
}

int thin_outer(ThinOuterParams& p, cudaStream_t st) {
    const bool vec = (p.C % 4 == 0) && (p.ldt % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.T) & 15) == 0) && p.C <= 1024 &&
                     256 % (p.C / 4) == 0 && (size_t)(p.J + 2) * p.C * sizeof(float) <= 40 * 1024;
    if (vec) {
        long rows = cdiv(p.M, 8L * kNumSMs);
        if (rows < 32) rows = 32;
        p.rows_per_block = rows;
        thin_outer_vec_kernel<<<(unsigned)cdiv(p.M, rows), 256, sizeof(float) * (size_t)(p.J + 2) * p.C, st>>>(p);
        count_launch();
        return check_launch("thin_outer_vec_kernel");
    }
    long rows = cdiv(p.M, 4L * kNumSMs);
    if (rows < 64) rows = 64;
    p.rows_per_block = rows;
    dim3 grid((unsigned)cdiv(p.M, rows), (unsigned)cdiv(p.C, 32));
    thin_outer_kernel<<<grid, 256, 0, st>>>(p);
    count_launch();
    return check_launch("thin_outer_kernel");
}

}  // namespace npf
