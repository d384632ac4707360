// Tensor-core (tcgen05 / TMEM) path of the multi-head cross-attention, head dim 16 or 32 (the "multihead" /
// "transformer" attenders: 8 heads x 16).   NPF_PREC_BF16: operands rounded to bf16; NPF_PREC_BF16X3: hi+lo split,
// 3 MMAs per product (fp32-level logits and P.V).
//
// Forward, one CTA per (task, head, 128-query block), 128 threads = 128 TMEM lanes = 128 query rows:
//   per 128-key chunk:  S = Q K^T          tcgen05.mma  M=128 N=128 K=D        (accumulator: TMEM cols [0,128))
//                       thread r reads row r of S from TMEM (tcgen05.ld), online softmax (running max / sum, exp2),
//                       writes P (bf16 hi/lo) straight into the UMMA K-major operand layout in shared memory
//                       O_c = P V          tcgen05.mma  M=128 N=Dv  K=128      (accumulator: TMEM cols [128,128+Dv))
//                       thread r: O_r = O_r * alpha + O_c[r, :]   (registers)
//   The [Tq, Tk] logits / probabilities never leave the SM.
//
// Operand layouts (no swizzle, 8x8 core matrices of 128 B; see tc_common.cuh / gemm_tc.cu):
//   Q, K   K-major over d :  (d/8)*LBO + (row/8)*128 + (row%8)*16 + (d%8)*2,   LBO = rows*16
//   P      K-major over key: (key/8)*2048 + (q/8)*128 + (q%8)*16 + (key%8)*2
//   V      MN-major (N = channel c, K = key): (key/8)*LBO + (c/8)*128 + (key%8)*16 + (c%8)*2,  LBO = (Dv/8)*128
#include "tc_common.cuh"

namespace npf {

struct AttnTcParams {
    const float* Q; const float* K; const float* V; const float* O; const float* LSE; const float* dO;
    float* Oo; float* LSEo; float* dQ; float* dK; float* dV;
    int Tq, Tk, H, D, Dv;
    float scale;
};

constexpr int kQB = 128;   // query rows per CTA (= TMEM lanes)
constexpr int kKC = 128;   // keys per chunk

__device__ __forceinline__ uint4 pack8(const float (&v)[8]) {
    return make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
}
__device__ __forceinline__ void split8(const float (&v)[8], float (&lo)[8]) {
#pragma unroll
    for (int i = 0; i < 8; ++i) lo[i] = v[i] - bf16_round(v[i]);
}

// Stage `rows_valid` rows (thread = row) of a [rows, width] fp32 slice (row stride ld) as a K-major operand with `R`
// rows in the tile: chunk (w/8) at (w/8)*R*16 + (row/8)*128 + (row%8)*16.
template <int NSPLIT>
__device__ __forceinline__ void stage_rows_kmajor(uint8_t* hi, uint8_t* lo, const float* __restrict__ src, long ld, int row, bool valid, int width,
                                                  int R) {
    const uint32_t base = (uint32_t)(row >> 3) * 128u + (uint32_t)(row & 7) * 16u;
    for (int ch = 0; ch < (width >> 3); ++ch) {
        float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        if (valid) {
            const float4 a = __ldg(reinterpret_cast<const float4*>(src + (long)row * ld + ch * 8));
            const float4 b = __ldg(reinterpret_cast<const float4*>(src + (long)row * ld + ch * 8 + 4));
            v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
        }
        const uint32_t off = (uint32_t)ch * (uint32_t)R * 16u + base;
        *reinterpret_cast<uint4*>(hi + off) = pack8(v);
        if (NSPLIT == 3) {
            float l[8];
            split8(v, l);
            *reinterpret_cast<uint4*>(lo + off) = pack8(l);
        }
    }
}

// V chunk as MN-major B operand (N = channel, K = key): thread = key row.
template <int NSPLIT>
__device__ __forceinline__ void stage_rows_mnmajor(uint8_t* hi, uint8_t* lo, const float* __restrict__ src, long ld, int row, bool valid, int width) {
    const uint32_t lbo = (uint32_t)(width >> 3) * 128u;
    const uint32_t base = (uint32_t)(row >> 3) * lbo + (uint32_t)(row & 7) * 16u;
    for (int ch = 0; ch < (width >> 3); ++ch) {
        float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        if (valid) {
            const float4 a = __ldg(reinterpret_cast<const float4*>(src + (long)row * ld + ch * 8));
            const float4 b = __ldg(reinterpret_cast<const float4*>(src + (long)row * ld + ch * 8 + 4));
            v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
        }
        const uint32_t off = base + (uint32_t)ch * 128u;
        *reinterpret_cast<uint4*>(hi + off) = pack8(v);
        if (NSPLIT == 3) {
            float l[8];
            split8(v, l);
            *reinterpret_cast<uint4*>(lo + off) = pack8(l);
        }
    }
}

template <int NSPLIT, int DV>
__global__ void __launch_bounds__(128) xattn_fwd_tc_kernel(AttnTcParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar_s, bar_o;
    __shared__ uint32_t tmem_slot;
    const int D = p.D;
    const uint32_t q_bytes = kQB * D * 2u, k_bytes = kKC * D * 2u, v_bytes = kKC * DV * 2u, p_bytes = kQB * kKC * 2u;
    uint8_t* q_hi = smem_raw;
    uint8_t* k_hi = q_hi + q_bytes;
    uint8_t* v_hi = k_hi + k_bytes;
    uint8_t* p_hi = v_hi + v_bytes;
    const uint32_t half = q_bytes + k_bytes + v_bytes + p_bytes;
    uint8_t* q_lo = q_hi + half; uint8_t* k_lo = k_hi + half; uint8_t* v_lo = v_hi + half; uint8_t* p_lo = p_hi + half;   // NSPLIT == 3 only

    const int tid = threadIdx.x, warp = tid >> 5;
    const int qb = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
    const int q_row = qb * kQB + tid;
    const bool q_ok = q_row < p.Tq;
    const long ldq = (long)p.H * D, ldv = (long)p.H * DV;
    const float* Qb = p.Q + ((long)b * p.Tq) * ldq + h * D;
    const float* Kb = p.K + ((long)b * p.Tk) * ldq + h * D;
    const float* Vb = p.V + ((long)b * p.Tk) * ldv + h * DV;

    constexpr uint32_t kCols = 256;   // S: [0,128), O chunk: [128, 128 + DV)
    if (warp == 0) tmem_alloc(&tmem_slot, kCols);
    if (tid == 0) { mbar_init(&bar_s, 1); mbar_init(&bar_o, 1); }
    stage_rows_kmajor<NSPLIT>(q_hi, q_lo, Qb + (long)qb * kQB * ldq, ldq, tid, q_ok, D, kQB);   // tile row tid <- query row q_row
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    const uint32_t t_s = tmem + ((uint32_t)(32 * warp) << 16);
    const uint32_t t_o = t_s + 128u;
    const uint32_t idesc_s = make_idesc(128, kKC, 0, 0);
    const uint32_t idesc_o = make_idesc(128, DV, 0, 1);
    const uint32_t q_lbo = kQB * 16u, k_lbo = kKC * 16u, p_lbo = kQB * 16u, v_lbo = (uint32_t)(DV >> 3) * 128u;

    const float sl2 = p.scale * 1.4426950408889634f;      // logits in log2 units: exp(x) = exp2(x * log2 e)
    float m_run = -INFINITY, l_run = 0.f;
    float o_run[DV];
#pragma unroll
    for (int c = 0; c < DV; ++c) o_run[c] = 0.f;

    uint32_t ph = 0;
    for (int k0 = 0; k0 < p.Tk; k0 += kKC) {
        // stage this chunk of K and V (thread = key row); the previous chunk's MMAs have completed (bar_o waited below)
        const int key = k0 + tid;
        const bool k_ok = key < p.Tk;
        stage_rows_kmajor<NSPLIT>(k_hi, k_lo, Kb + (long)k0 * ldq, ldq, tid, k_ok, D, kKC);
        stage_rows_mnmajor<NSPLIT>(v_hi, v_lo, Vb + (long)k0 * ldv, ldv, tid, k_ok, DV);
        fence_async_smem();
        tc_fence_before();
        __syncthreads();
        if (tid == 0) {
            tc_fence_after();
            uint32_t acc = 0;
            for (int ks = 0; ks < D / 16; ++ks) {
                const uint32_t qo = (uint32_t)ks * 2u * q_lbo, ko = (uint32_t)ks * 2u * k_lbo;
                umma_bf16(tmem, make_desc(smem_u32(q_hi) + qo, q_lbo, 128), make_desc(smem_u32(k_hi) + ko, k_lbo, 128), idesc_s, acc);
                acc = 1;
                if (NSPLIT == 3) {
                    umma_bf16(tmem, make_desc(smem_u32(q_hi) + qo, q_lbo, 128), make_desc(smem_u32(k_lo) + ko, k_lbo, 128), idesc_s, 1);
                    umma_bf16(tmem, make_desc(smem_u32(q_lo) + qo, q_lbo, 128), make_desc(smem_u32(k_hi) + ko, k_lbo, 128), idesc_s, 1);
                }
            }
            umma_commit(&bar_s);
        }
        mbar_wait(&bar_s, ph);
        tc_fence_after();

        // ---- online softmax over this chunk's 128 logits of row `tid` ----
        const int n_valid = min(kKC, p.Tk - k0);
        float mx = -INFINITY;
#pragma unroll 1
        for (int c0 = 0; c0 < kKC; c0 += 32) {
            float v[32];
            tmem_ld32(t_s + (uint32_t)c0, v);
#pragma unroll
            for (int j = 0; j < 32; ++j)
                if (c0 + j < n_valid) mx = fmaxf(mx, v[j] * sl2);
        }
        const float m_new = fmaxf(m_run, mx);               // finite: every chunk holds >= 1 valid key
        const float alpha = exp2f(m_run - m_new);           // first chunk: exp2(-inf) = 0
        float lsum = 0.f;
#pragma unroll 1
        for (int c0 = 0; c0 < kKC; c0 += 32) {
            float v[32];
            tmem_ld32(t_s + (uint32_t)c0, v);
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                const float e = (c0 + j < n_valid) ? exp2f(fmaf(v[j], sl2, -m_new)) : 0.f;
                v[j] = e;
                lsum += e;
            }
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                float w8[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) w8[i] = v[g * 8 + i];
                const uint32_t off = (uint32_t)((c0 >> 3) + g) * p_lbo + (uint32_t)(tid >> 3) * 128u + (uint32_t)(tid & 7) * 16u;
                *reinterpret_cast<uint4*>(p_hi + off) = pack8(w8);
                if (NSPLIT == 3) {
                    float l8[8];
                    split8(w8, l8);
                    *reinterpret_cast<uint4*>(p_lo + off) = pack8(l8);
                }
            }
        }
        l_run = l_run * alpha + lsum;
        m_run = m_new;

        fence_async_smem();
        tc_fence_before();
        __syncthreads();
        if (tid == 0) {
            tc_fence_after();
            uint32_t acc = 0;
            for (int ks = 0; ks < kKC / 16; ++ks) {
                const uint32_t po = (uint32_t)ks * 2u * p_lbo, vo = (uint32_t)ks * 2u * v_lbo;
                umma_bf16(tmem + 128u, make_desc(smem_u32(p_hi) + po, p_lbo, 128), make_desc(smem_u32(v_hi) + vo, v_lbo, 128), idesc_o, acc);
                acc = 1;
                if (NSPLIT == 3) {
                    umma_bf16(tmem + 128u, make_desc(smem_u32(p_hi) + po, p_lbo, 128), make_desc(smem_u32(v_lo) + vo, v_lbo, 128), idesc_o, 1);
                    umma_bf16(tmem + 128u, make_desc(smem_u32(p_lo) + po, p_lbo, 128), make_desc(smem_u32(v_hi) + vo, v_lbo, 128), idesc_o, 1);
                }
            }
            umma_commit(&bar_o);
        }
        mbar_wait(&bar_o, ph);
        ph ^= 1;
        tc_fence_after();
#pragma unroll
        for (int c0 = 0; c0 < DV; c0 += 16) {
            float oc[16];
            tmem_ld16(t_o + (uint32_t)c0, oc);
#pragma unroll
            for (int j = 0; j < 16; ++j) o_run[c0 + j] = fmaf(o_run[c0 + j], alpha, oc[j]);
        }
        tc_fence_before();
        __syncthreads();       // all TMEM reads of this chunk done before the next chunk's MMAs / staging
    }
    if (q_ok) {
        const float inv = 1.f / l_run;
        float* out = p.Oo + ((long)b * p.Tq + q_row) * ldv + h * DV;
#pragma unroll
        for (int c = 0; c < DV; c += 4)
            *reinterpret_cast<float4*>(out + c) = make_float4(o_run[c] * inv, o_run[c + 1] * inv, o_run[c + 2] * inv, o_run[c + 3] * inv);
        p.LSEo[((long)b * p.H + h) * p.Tq + q_row] = (m_run + log2f(l_run)) * 0.6931471805599453f;   // back to natural log
    }
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, kCols);
}

template <int NSPLIT, int DV>
static int launch_fwd(AttnTcParams& p, int B, cudaStream_t st) {
    const size_t half = (size_t)(kQB * p.D + kKC * p.D + kKC * DV + kQB * kKC) * 2;
    const size_t smem = half * (NSPLIT == 3 ? 2 : 1);
    static bool attr = false;
    if (!attr) {
        if (cudaFuncSetAttribute(xattn_fwd_tc_kernel<NSPLIT, DV>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess) {
            cudaGetLastError();
            return NPF_ENOTSUP;
        }
        attr = true;
    }
    dim3 grid((unsigned)cdiv(p.Tq, kQB), (unsigned)p.H, (unsigned)B);
    xattn_fwd_tc_kernel<NSPLIT, DV><<<grid, 128, smem, st>>>(p);
    count_launch();
    return check_launch("xattn_fwd_tc_kernel");
}

static bool attn_tc_ok(const AttnTcParams& p) {
    auto al = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
    return (p.D == 16 || p.D == 32) && (p.Dv == 16 || p.Dv == 32) && al(p.Q) && al(p.K) && al(p.V) && p.Tk >= 1;
}

int xattn_fwd_tc(const float* Q, const float* K, const float* V, float* O, float* LSE, int B, int Tq, int Tk, int H, int D, int Dv,
                 float scale, int precision, cudaStream_t st) {
    AttnTcParams p{};
    p.Q = Q; p.K = K; p.V = V; p.Oo = O; p.LSEo = LSE;
    p.Tq = Tq; p.Tk = Tk; p.H = H; p.D = D; p.Dv = Dv; p.scale = scale;
    if (!attn_tc_ok(p) || (reinterpret_cast<uintptr_t>(O) & 15)) return NPF_ENOTSUP;
    if (precision == NPF_PREC_BF16X3) return Dv == 16 ? launch_fwd<3, 16>(p, B, st) : launch_fwd<3, 32>(p, B, st);
    return Dv == 16 ? launch_fwd<1, 16>(p, B, st) : launch_fwd<1, 32>(p, B, st);
}

}  // namespace npf
