// Tensor-core path of the 128-wide linear layers: tcgen05.mma (UMMA) with the accumulator in TMEM.
//
//   NPF_PREC_BF16   : operands rounded to bf16, fp32 accumulate                        (1e-2 parity)
//   NPF_PREC_BF16X3 : x = hi + lo (both bf16); x.w ~ hi.hi + hi.lo + lo.hi, fp32 accumulate (1e-4 parity; 3 MMAs)
//
// Activations stay fp32 in HBM (layout contract of the library), so operands are converted on the fly: the CTA's
// threads load the fp32 tile with coalesced 16-byte loads, apply the prologue (relu), split/round to bf16 and write
// the UMMA canonical *no-swizzle* layout straight into shared memory (8x8 "core matrices" of 128 contiguous
// bytes; LBO = stride between core matrices along the reduction dim, SBO = along rows).  One elected thread then
// issues the K/16 MMAs (x1 or x3), commits to an mbarrier, and all 8 warps drain the 128xN fp32 accumulator from
// TMEM (tcgen05.ld 32x32b) through the bias / relu / rank-1 / relu-mask epilogue into global memory.
// The weight matrix is converted and staged ONCE per CTA (persistent over row tiles).
//
//   linear_tc_kernel : fwd  Y = act(X) W^T            (A = X rows K-major,  B = W   [N x K] K-major)
//                      bwd  dX = dY W                 (A = dY rows K-major, B = W^T [K x N], staged transposed)
//   wgrad_tc_kernel  : dW += dY^T X                   (A = dY, B = X, both MN-major: the reduction runs over rows)
//
// Shapes covered: reduction dim <= 128 and output dim <= 256, both multiples of 16.  Anything else reports
// NPF_ENOTSUP and the caller uses the fp32 FFMA kernel.
#include "tc_common.cuh"

namespace npf {

// ------------------------------------------------------------------------------------------------ K-major staging
// Tile of R rows x KR reduction elements, element (row, k) at byte
//     (k/8) * LBO + (row/8) * 128 + (row%8) * 16 + (k%8) * 2,      LBO = R * 16
// (core matrix = 8 rows x 8 k = 128 contiguous bytes).  Source: fp32 row-major with leading dimension ld.
// A half-warp owns one core matrix per step (conflict-free 8-byte stores; 8 full 32-byte sectors per load).
// Loads are split from the convert+store so that a whole tile (16 x 16 B per thread = 64 KB per CTA) is in
// flight at once, and so that the NEXT tile can be prefetched into registers under the current tile's MMA + epilogue.
constexpr int kLinThreads = 512;            // linear_tc_kernel: 16 warps (32 half-warps stage 32 core matrices per step)
constexpr int kHW = kLinThreads / 16;
constexpr int kPre = 8;                     // float4 registers per thread per batch: 128 x 128 tile / 4 / 512 threads

__device__ __forceinline__ int ilog2(int x) { return 31 - __clz(x); }

// All tile extents on this path are powers of two (checked on the host), so the (core-matrix -> row group, k chunk)
// maps are shifts, and because 32 half-warps step through core matrices 32 at a time the k chunk of a thread is
// CONSTANT: only the row group advances -> one pointer increment and one smem-offset increment per step.
__device__ __forceinline__ void load_kmajor(float4 (&pre)[kPre], const float* __restrict__ src, long ld, long row0, int rows_valid,
                                            int R, int KR, int cm_base, int vec_ok) {
    const int hw = threadIdx.x >> 4, l16 = threadIdx.x & 15;
    const int r = l16 & 7, half = l16 >> 3;
    const int n_kc = KR >> 3, lg = ilog2(n_kc);
    const int cm0 = cm_base + hw;
    const int kc = cm0 & (n_kc - 1);
    int rg = cm0 >> lg;
    const int rg_step = kHW >> lg, n_rg = R >> 3;   // n_kc <= 16 divides kHW = 32
    const float* g = src + (row0 + rg * 8 + r) * ld + kc * 8 + half * 4;
    const long gstep = (long)rg_step * 8 * ld;
#pragma unroll
    for (int i = 0; i < kPre; ++i) {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (rg < n_rg && rg * 8 + r < rows_valid) {
            if (vec_ok) v = __ldg(reinterpret_cast<const float4*>(g));
            else { v.x = __ldg(g); v.y = __ldg(g + 1); v.z = __ldg(g + 2); v.w = __ldg(g + 3); }
        }
        pre[i] = v;
        rg += rg_step;
        g += gstep;
    }
}

template <int NSPLIT>
__device__ __forceinline__ void store_kmajor(const float4 (&pre)[kPre], uint8_t* hi, uint8_t* lo, int R, int KR, int cm_base, int relu) {
    const int hw = threadIdx.x >> 4, l16 = threadIdx.x & 15;
    const int r = l16 & 7, half = l16 >> 3;
    const int n_kc = KR >> 3, lg = ilog2(n_kc);
    const int cm0 = cm_base + hw;
    const int kc = cm0 & (n_kc - 1);
    int rg = cm0 >> lg;
    const int rg_step = kHW >> lg, n_rg = R >> 3;
    uint32_t off = (uint32_t)kc * ((uint32_t)R * 16u) + (uint32_t)rg * 128u + (uint32_t)r * 16u + (uint32_t)half * 8u;
#pragma unroll
    for (int i = 0; i < kPre; ++i) {
        if (rg < n_rg) {
            float4 v = pre[i];
            if (relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
            const uint32_t h01 = pack_bf16(v.x, v.y), h23 = pack_bf16(v.z, v.w);
            *reinterpret_cast<uint2*>(hi + off) = make_uint2(h01, h23);
            if (NSPLIT == 3) {   // residuals against the packed hi halves (bf16 -> fp32 is a 16-bit shift)
                const float rx = v.x - __uint_as_float(h01 << 16), ry = v.y - __uint_as_float(h01 & 0xFFFF0000u);
                const float rz = v.z - __uint_as_float(h23 << 16), rw = v.w - __uint_as_float(h23 & 0xFFFF0000u);
                *reinterpret_cast<uint2*>(lo + off) = make_uint2(pack_bf16(rx, ry), pack_bf16(rz, rw));
            }
        }
        rg += rg_step;
        off += (uint32_t)rg_step * 128u;
    }
}

// Transposed staging of the weights for the data gradient: B'(row = k_out, red = n) = W[n, k_out], K-major in n.
// Batched float4 loads along k_out (coalesced rows of W), four 2-byte scatter stores each (once per CTA).
template <int NSPLIT>
__device__ __forceinline__ void stage_kmajor_transposed(uint8_t* hi, uint8_t* lo, const float* __restrict__ W, long ldw, int R /*rows = K_out*/,
                                                        int KR /*reduction = N*/, int vec_ok) {
    const uint32_t lbo = (uint32_t)R * 16u;
    const int rq = R >> 2, lgq = ilog2(rq), total = rq * KR;      // float4 count; rq is a power of two <= 64
    const int row0 = (threadIdx.x & (rq - 1)) * 4;                // constant per thread (512 is a multiple of rq)
    for (int base = 0; base < total; base += kLinThreads * kPre) {
        float4 pre[kPre];
        const int n0 = (base + threadIdx.x) >> lgq, n_step = kLinThreads >> lgq;
#pragma unroll
        for (int i = 0; i < kPre; ++i) {
            const int n = n0 + i * n_step;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (n < KR) {
                const float* g = W + (long)n * ldw + row0;
                if (vec_ok) v = __ldg(reinterpret_cast<const float4*>(g));
                else { v.x = __ldg(g); v.y = __ldg(g + 1); v.z = __ldg(g + 2); v.w = __ldg(g + 3); }
            }
            pre[i] = v;
        }
#pragma unroll
        for (int i = 0; i < kPre; ++i) {
            const int n = n0 + i * n_step;
            if (n >= KR) continue;
            const float vv[4] = {pre[i].x, pre[i].y, pre[i].z, pre[i].w};
            const uint32_t off0 = (uint32_t)(n >> 3) * lbo + (uint32_t)(row0 >> 3) * 128u + (uint32_t)(row0 & 7) * 16u + (uint32_t)(n & 7) * 2u;
#pragma unroll
            for (int e = 0; e < 4; ++e) {      // rows row0 .. row0+3 stay inside one 8-row group (row0 is a multiple of 4)
                const uint32_t off = off0 + (uint32_t)e * 16u;
                const __nv_bfloat16 h = __float2bfloat16_rn(vv[e]);
                *reinterpret_cast<__nv_bfloat16*>(hi + off) = h;
                if (NSPLIT == 3) *reinterpret_cast<__nv_bfloat16*>(lo + off) = __float2bfloat16_rn(vv[e] - __bfloat162float(h));
            }
        }
    }
}

struct TcLinParams {
    const float* A; long lda;       // [M, KR] activations (X or dY)
    const float* W; long ldw;       // fwd: [NO, KR]; bwd-data: [KR, NO]
    float* C; long ldc;             // [M, NO]
    const float* bias;              // [NO] or null
    const float* u; const float* w2; long ldw2;   // rank-1 epilogue
    const float* mask; long ldm;    // relu mask source [M, NO]
    int M, KR, NO;
    int relu_in, relu_out, transposed_w, a_vec, w_vec, c_vec;
    int n_tiles;
};

// ------------------------------------------------------------------------------------------------ fwd / bwd-data
constexpr int kScratchLd = 36;     // floats per staged row: 32 + 4 keeps both the row-wise STS.128 and the LDS.128 conflict-free

// KR_T / NO_T: compile-time reduction / output extents (0 = run-time values from the parameter block); HAS_U / HAS_MASK:
// rank-1 epilogue term / relu-mask epilogue compiled in.  The 128 x 128 instantiations are the hot ones: with the
// extents known every staging predicate and index computation folds away.
template <int NSPLIT, int KR_T, int NO_T, bool HAS_U, bool HAS_MASK>
__global__ void __launch_bounds__(kLinThreads, 1) linear_tc_kernel(TcLinParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t mma_bar;
    __shared__ uint32_t tmem_slot;
    __shared__ __align__(16) float s_bias[256], s_w2[256];

    const int KR = KR_T ? KR_T : p.KR, NO = NO_T ? NO_T : p.NO;
    const uint32_t a_bytes = 128u * KR * 2u, b_bytes = (uint32_t)NO * KR * 2u;
    uint8_t* a_hi = smem_raw;
    uint8_t* a_lo = a_hi + a_bytes;                               // only touched when NSPLIT == 3
    uint8_t* b_hi = smem_raw + (NSPLIT == 3 ? 2 : 1) * a_bytes;
    uint8_t* b_lo = b_hi + b_bytes;
    float* scratch_all = reinterpret_cast<float*>(smem_raw + (NSPLIT == 3 ? 2 : 1) * (a_bytes + b_bytes));

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    float* scratch = scratch_all + warp * (32 * kScratchLd);
    const uint32_t ncols = NO <= 32 ? 32u : (NO <= 64 ? 64u : (NO <= 128 ? 128u : 256u));
    if (warp == 0) tmem_alloc(&tmem_slot, ncols);
    if (tid == 0) mbar_init(&mma_bar, 1);
    for (int c = tid; c < 256; c += kLinThreads) {
        s_bias[c] = (p.bias && c < NO) ? __ldg(p.bias + c) : 0.f;
        s_w2[c] = (p.w2 && c < NO) ? __ldg(p.w2 + (long)c * p.ldw2) : 0.f;
    }

    float4 pre[kPre];
    int tile = blockIdx.x;
    // first activation tile: in flight while the weights are staged
    if (tile < p.n_tiles) load_kmajor(pre, p.A, p.lda, (long)tile * 128, min(128, p.M - tile * 128), 128, KR, 0, p.a_vec);
    // weights: converted and staged once per CTA
    if (p.transposed_w) {
        stage_kmajor_transposed<NSPLIT>(b_hi, b_lo, p.W, p.ldw, NO, KR, p.w_vec);
    } else {
        const int n_cm = (NO >> 3) * (KR >> 3);
        for (int base = 0; base < n_cm; base += kHW * kPre) {
            float4 wpre[kPre];
            load_kmajor(wpre, p.W, p.ldw, 0, NO, NO, KR, base, p.w_vec);
            store_kmajor<NSPLIT>(wpre, b_hi, b_lo, NO, KR, base, 0);
        }
    }
    if (tile < p.n_tiles) store_kmajor<NSPLIT>(pre, a_hi, a_lo, 128, KR, 0, p.relu_in);
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    const uint32_t idesc = make_idesc(128, NO, 0, 0);
    const uint32_t a_lbo = 128u * 16u, b_lbo = (uint32_t)NO * 16u;

    // epilogue geometry: warp w drains TMEM lanes 32*(w&3)..+31 (its 32 rows) for the column chunks w>>2, w>>2 + 4, ...
    const int CW = NO < 32 ? NO : 32;            // chunk width (16 or 32 columns)
    const int n_chunks = NO / CW;
    const int lane_base = 32 * (warp & 3);
    const int qpr = CW >> 2, rpi = 32 / qpr;     // float4 per staged row, rows per coalesced instruction
    const int r_in = lane / qpr, c4 = (lane % qpr) * 4;

    uint32_t phase = 0;
    for (; tile < p.n_tiles; tile += gridDim.x) {
        const int m0 = tile * 128;
        if (tid == 0) {
            const uint32_t sa_hi = smem_u32(a_hi), sa_lo = smem_u32(a_lo), sb_hi = smem_u32(b_hi), sb_lo = smem_u32(b_lo);
            uint32_t acc = 0;
            for (int ks = 0; ks < KR / 16; ++ks) {
                const uint32_t ao = (uint32_t)ks * 2u * a_lbo, bo = (uint32_t)ks * 2u * b_lbo;
                umma_bf16(tmem, make_desc(sa_hi + ao, a_lbo, 128), make_desc(sb_hi + bo, b_lbo, 128), idesc, acc);
                acc = 1;
                if (NSPLIT == 3) {
                    umma_bf16(tmem, make_desc(sa_hi + ao, a_lbo, 128), make_desc(sb_lo + bo, b_lbo, 128), idesc, 1);
                    umma_bf16(tmem, make_desc(sa_lo + ao, a_lbo, 128), make_desc(sb_hi + bo, b_lbo, 128), idesc, 1);
                }
            }
            umma_commit(&mma_bar);   // implies tcgen05.fence::before_thread_sync
        }
        // prefetch the next tile into registers: in flight under this tile's MMA and epilogue
        const int next = tile + gridDim.x;
        if (next < p.n_tiles) load_kmajor(pre, p.A, p.lda, (long)next * 128, min(128, p.M - next * 128), 128, KR, 0, p.a_vec);

        mbar_wait(&mma_bar, phase);
        phase ^= 1;
        tc_fence_after();

        for (int ch = warp >> 2; ch < n_chunks; ch += 4) {
            const int c0 = ch * CW;
            // TMEM -> registers (thread = row) -> per-warp smem tile
            if (CW == 32) {
                float v[32];
                tmem_ld32(tmem + ((uint32_t)lane_base << 16) + (uint32_t)c0, v);
#pragma unroll
                for (int j = 0; j < 32; j += 4)
                    *reinterpret_cast<float4*>(scratch + lane * kScratchLd + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
            } else {
                float v[16];
                tmem_ld16(tmem + ((uint32_t)lane_base << 16) + (uint32_t)c0, v);
#pragma unroll
                for (int j = 0; j < 16; j += 4)
                    *reinterpret_cast<float4*>(scratch + lane * kScratchLd + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
            }
            __syncwarp();
            // smem -> global, row-contiguous: a warp instruction covers `rpi` rows x CW columns (full 32-byte sectors)
            const float4 bb = *reinterpret_cast<const float4*>(&s_bias[c0 + c4]);
            const float4 ww = *reinterpret_cast<const float4*>(&s_w2[c0 + c4]);
            for (int r0 = 0; r0 < 32; r0 += rpi) {
                const int r = r0 + r_in;
                const int row = m0 + lane_base + r;
                if (row < p.M) {
                    float4 x = *reinterpret_cast<const float4*>(scratch + r * kScratchLd + c4);
                    x.x += bb.x; x.y += bb.y; x.z += bb.z; x.w += bb.w;
                    if (HAS_U && p.u) {
                        const float up = __ldg(p.u + row);
                        x.x = fmaf(up, ww.x, x.x); x.y = fmaf(up, ww.y, x.y); x.z = fmaf(up, ww.z, x.z); x.w = fmaf(up, ww.w, x.w);
                    }
                    if (p.relu_out) { x.x = fmaxf(x.x, 0.f); x.y = fmaxf(x.y, 0.f); x.z = fmaxf(x.z, 0.f); x.w = fmaxf(x.w, 0.f); }
                    if (HAS_MASK && p.mask) {
                        const float* mk = p.mask + (long)row * p.ldm + c0 + c4;
                        float4 mv;
                        if ((p.ldm & 3) == 0 && (reinterpret_cast<uintptr_t>(p.mask) & 15) == 0) mv = __ldg(reinterpret_cast<const float4*>(mk));
                        else mv = make_float4(__ldg(mk), __ldg(mk + 1), __ldg(mk + 2), __ldg(mk + 3));
                        x.x = mv.x > 0.f ? x.x : 0.f; x.y = mv.y > 0.f ? x.y : 0.f; x.z = mv.z > 0.f ? x.z : 0.f; x.w = mv.w > 0.f ? x.w : 0.f;
                    }
                    float* out = p.C + (long)row * p.ldc + c0 + c4;
                    if (p.c_vec) *reinterpret_cast<float4*>(out) = x;
                    else { out[0] = x.x; out[1] = x.y; out[2] = x.z; out[3] = x.w; }
                }
            }
            __syncwarp();
        }
        tc_fence_before();   // TMEM reads done (and the MMAs have consumed the A buffer) before it is overwritten
        __syncthreads();
        if (next < p.n_tiles) {
            store_kmajor<NSPLIT>(pre, a_hi, a_lo, 128, KR, 0, p.relu_in);
            fence_async_smem();   // generic-proxy smem writes -> visible to the tensor core (async proxy)
        }
        __syncthreads();
        tc_fence_after();
    }
    if (warp == 0) tmem_dealloc(tmem, ncols);
}

// ------------------------------------------------------------------------------------------------ weight gradient
// dW[n, k] += sum_m dY[m, n] X[m, k].  MMA shape M = N_out (rows of dW, <= 128 -> padded to 128), N = K_out, K = rows m.
// Both operands MN-major, no swizzle: element (mn, k) at byte (k/8)*LBO + (mn/8)*128 + (k%8)*16 + (mn%8)*2,
// LBO = MN*16: a source row m (contiguous in mn) lands as 16-byte chunks -> conflict-free 16-byte stores.
// Sub-tiles of 64 rows (4 k-steps): both operands of a sub-tile are prefetched into registers (batched loads).
constexpr int kWgRows = 64;
constexpr int kWgIt = 4;     // 16-byte chunks per thread per batch: 64 rows x 128 columns / 8 / 256 threads

// chunk c of a [64 rows x MN] operand: 8 lanes cover 8 consecutive rows of one 8-wide column chunk.  MN/8 is a power
// of two <= 32 and a thread block advances 32 chunk columns per step, so a thread's column chunk j is constant and only
// the 8-row group kg advances.
__device__ __forceinline__ void load_mnmajor(float4 (&pre)[2 * kWgIt], const float* __restrict__ src, long ld, long row0, int rows_valid, int MN,
                                             int mn_valid, int c_base, int vec_ok) {
    const int n_chunks = MN >> 3, lgc = ilog2(n_chunks);
    const int r = threadIdx.x & 7, q = (c_base >> 3) + (threadIdx.x >> 3);
    const int j = q & (n_chunks - 1), kg_step = 32 >> lgc;
    int kg = q >> lgc;
    const bool col_ok = j * 8 < mn_valid, full = vec_ok && (j * 8 + 8 <= mn_valid);
    const float* g = src + (row0 + kg * 8 + r) * ld + j * 8;
    const long gstep = (long)kg_step * 8 * ld;
#pragma unroll
    for (int i = 0; i < kWgIt; ++i) {
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
        if (kg < (kWgRows >> 3) && kg * 8 + r < rows_valid && col_ok) {
            if (full) {
                a = __ldg(reinterpret_cast<const float4*>(g));
                b = __ldg(reinterpret_cast<const float4*>(g + 4));
            } else {
                float t[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) t[e] = (j * 8 + e < mn_valid) ? __ldg(g + e) : 0.f;
                a = make_float4(t[0], t[1], t[2], t[3]);
                b = make_float4(t[4], t[5], t[6], t[7]);
            }
        }
        pre[2 * i] = a; pre[2 * i + 1] = b;
        kg += kg_step;
        g += gstep;
    }
}

template <int NSPLIT>
__device__ __forceinline__ void store_mnmajor(const float4 (&pre)[2 * kWgIt], uint8_t* hi, uint8_t* lo, int MN, int c_base, int relu, uint32_t lbo) {
    const int n_chunks = MN >> 3, lgc = ilog2(n_chunks);
    const int r = threadIdx.x & 7, q = (c_base >> 3) + (threadIdx.x >> 3);
    const int j = q & (n_chunks - 1), kg_step = 32 >> lgc;
    int kg = q >> lgc;
    uint32_t off = (uint32_t)kg * lbo + (uint32_t)j * 128u + (uint32_t)r * 16u;
#pragma unroll
    for (int i = 0; i < kWgIt; ++i) {
        if (kg < (kWgRows >> 3)) {
            float v[8] = {pre[2 * i].x, pre[2 * i].y, pre[2 * i].z, pre[2 * i].w, pre[2 * i + 1].x, pre[2 * i + 1].y, pre[2 * i + 1].z, pre[2 * i + 1].w};
            if (relu) {
#pragma unroll
                for (int e = 0; e < 8; ++e) v[e] = fmaxf(v[e], 0.f);
            }
            const uint4 h = make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
            *reinterpret_cast<uint4*>(hi + off) = h;
            if (NSPLIT == 3) {
                const uint32_t hh[4] = {h.x, h.y, h.z, h.w};
                uint32_t ll[4];
#pragma unroll
                for (int e = 0; e < 4; ++e)
                    ll[e] = pack_bf16(v[2 * e] - __uint_as_float(hh[e] << 16), v[2 * e + 1] - __uint_as_float(hh[e] & 0xFFFF0000u));
                *reinterpret_cast<uint4*>(lo + off) = make_uint4(ll[0], ll[1], ll[2], ll[3]);
            }
        }
        kg += kg_step;
        off += (uint32_t)kg_step * lbo;
    }
}

struct TcWgParams {
    const float* dY; long lddy;    // [M, N]
    const float* X; long ldx;      // [M, K]
    float* dW; long lddw;          // [N, K], accumulated with atomics
    float* db;                     // [N] bias gradient (+=) via an extra all-ones column of the X operand, or null
    long M; int N, K;
    int relu_in, dy_vec, x_vec;
    long rows_per_cta;
};

template <int NSPLIT>
__global__ void __launch_bounds__(256, 2) wgrad_tc_kernel(TcWgParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t mma_bar;
    __shared__ uint32_t tmem_slot;
    const int K = p.K;
    const int KE = p.db ? K + 16 : K;    // MMA N extent: 16 extra columns, the first one all ones -> D[:, K] = sum_m dY[m, :]
    const uint32_t a_bytes = (uint32_t)kWgRows * 128u * 2u, b_bytes = (uint32_t)kWgRows * (uint32_t)KE * 2u;
    uint8_t* a_hi = smem_raw;
    uint8_t* a_lo = a_hi + a_bytes;
    uint8_t* b_hi = smem_raw + (NSPLIT == 3 ? 2 : 1) * a_bytes;
    uint8_t* b_lo = b_hi + b_bytes;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t ncols = KE <= 32 ? 32u : (KE <= 64 ? 64u : (KE <= 128 ? 128u : 256u));
    if (warp == 0) tmem_alloc(&tmem_slot, ncols);
    if (tid == 0) mbar_init(&mma_bar, 1);
    const uint32_t a_lbo = 128u * 16u, b_lbo = (uint32_t)KE * 16u;
    if (p.db) {
        // constant part of the X operand: chunk K/8 of every row m holds [1,0,...,0], chunk K/8 + 1 holds zeros
        for (int i = tid; i < kWgRows * 2; i += 256) {
            const int m = i >> 1, jx = (K >> 3) + (i & 1);
            const uint32_t off = (uint32_t)(m >> 3) * b_lbo + (uint32_t)jx * 128u + (uint32_t)(m & 7) * 16u;
            *reinterpret_cast<uint4*>(b_hi + off) = make_uint4((i & 1) ? 0u : 0x00003F80u, 0u, 0u, 0u);   // bf16(1.0) = 0x3F80
            if (NSPLIT == 3) *reinterpret_cast<uint4*>(b_lo + off) = make_uint4(0u, 0u, 0u, 0u);
        }
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    const uint32_t idesc = make_idesc(128, KE, 1, 1);
    const int b_batches = (kWgRows * (K >> 3) + 256 * kWgIt - 1) / (256 * kWgIt);   // 1 for K <= 128, 2 for K = 256

    const long m_begin = (long)blockIdx.x * p.rows_per_cta;
    const long m_end = min(p.M, m_begin + p.rows_per_cta);
    uint32_t phase = 0, acc = 0;
    float4 pa[2 * kWgIt], pb[2 * kWgIt];
    if (m_begin < m_end) {
        const int rows = (int)min((long)kWgRows, m_end - m_begin);
        load_mnmajor(pa, p.dY, p.lddy, m_begin, rows, 128, p.N, 0, p.dy_vec);
        load_mnmajor(pb, p.X, p.ldx, m_begin, rows, K, K, 0, p.x_vec);
    }
    for (long m0 = m_begin; m0 < m_end; m0 += kWgRows) {
        const int rows = (int)min((long)kWgRows, m_end - m0);
        if (acc) {           // previous sub-tile's MMAs must have consumed the operand buffers
            mbar_wait(&mma_bar, phase);
            phase ^= 1;
        }
        store_mnmajor<NSPLIT>(pa, a_hi, a_lo, 128, 0, 0, a_lbo);
        store_mnmajor<NSPLIT>(pb, b_hi, b_lo, K, 0, p.relu_in, b_lbo);
        for (int bb = 1; bb < b_batches; ++bb) {       // K = 256: second half of the X sub-tile
            load_mnmajor(pb, p.X, p.ldx, m0, rows, K, K, bb * 256 * kWgIt, p.x_vec);
            store_mnmajor<NSPLIT>(pb, b_hi, b_lo, K, bb * 256 * kWgIt, p.relu_in, b_lbo);
        }
        const long mn = m0 + kWgRows;
        if (mn < m_end) {    // prefetch the next sub-tile: in flight under the fence / sync / MMA issue
            const int rn = (int)min((long)kWgRows, m_end - mn);
            load_mnmajor(pa, p.dY, p.lddy, mn, rn, 128, p.N, 0, p.dy_vec);
            load_mnmajor(pb, p.X, p.ldx, mn, rn, K, K, 0, p.x_vec);
        }
        fence_async_smem();
        __syncthreads();
        if (tid == 0) {
            tc_fence_after();
            const uint32_t sa_hi = smem_u32(a_hi), sa_lo = smem_u32(a_lo), sb_hi = smem_u32(b_hi), sb_lo = smem_u32(b_lo);
            for (int ks = 0; ks < kWgRows / 16; ++ks) {     // k-steps of 16 rows (2 k-groups of 8)
                const uint32_t ao = (uint32_t)ks * 2u * a_lbo, bo = (uint32_t)ks * 2u * b_lbo;
                umma_bf16(tmem, make_desc(sa_hi + ao, a_lbo, 128), make_desc(sb_hi + bo, b_lbo, 128), idesc, acc);
                acc = 1;
                if (NSPLIT == 3) {
                    umma_bf16(tmem, make_desc(sa_hi + ao, a_lbo, 128), make_desc(sb_lo + bo, b_lbo, 128), idesc, 1);
                    umma_bf16(tmem, make_desc(sa_lo + ao, a_lbo, 128), make_desc(sb_hi + bo, b_lbo, 128), idesc, 1);
                }
            }
            umma_commit(&mma_bar);
        }
        acc = 1;
    }
    if (acc) {
        mbar_wait(&mma_bar, phase);
        tc_fence_after();
        const int lane_base = 32 * (warp & 3);
        const int n = lane_base + lane;                    // row of dW
        const int split = ((K + 1) / 2 + 15) / 16 * 16;
        const int c_begin = (warp >> 2) ? split : 0, c_end = (warp >> 2) ? K : min(K, split);
        for (int c0 = c_begin; c0 < c_end; c0 += 16) {
            float v[16];
            tmem_ld16(tmem + ((uint32_t)lane_base << 16) + (uint32_t)c0, v);
            if (n < p.N) {
                float* d = p.dW + (long)n * p.lddw + c0;
                if ((p.lddw & 3) == 0 && (reinterpret_cast<uintptr_t>(p.dW) & 15) == 0) {
#pragma unroll
                    for (int j = 0; j < 16; j += 4) atomicAdd(reinterpret_cast<float4*>(d + j), make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]));
                } else {
#pragma unroll
                    for (int j = 0; j < 16; ++j) atomicAdd(d + j, v[j]);
                }
            }
        }
        if (p.db && warp < 4) {       // the ones column: D[n, K] = sum over rows of dY[:, n]
            float v[16];
            tmem_ld16(tmem + ((uint32_t)lane_base << 16) + (uint32_t)K, v);
            if (n < p.N) atomicAdd(p.db + n, v[0]);
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, ncols);
}

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }
static bool pow2(int x) { return x > 0 && (x & (x - 1)) == 0; }
static bool tc_shape_ok(int red, int out) { return pow2(red) && red >= 16 && red <= 128 && pow2(out) && out >= 16 && out <= 256; }

template <int NSPLIT>
static int launch_lin(TcLinParams& p, cudaStream_t st) {
    const size_t smem = (size_t)(NSPLIT == 3 ? 2 : 1) * (128 + p.NO) * p.KR * 2 + (size_t)(kLinThreads / 32) * 32 * kScratchLd * sizeof(float);
    static size_t reserved = 0;
    if (smem > reserved) {
        if (cudaFuncSetAttribute(linear_tc_kernel<NSPLIT, 128, 128, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024) != cudaSuccess ||
            cudaFuncSetAttribute(linear_tc_kernel<NSPLIT, 128, 128, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024) != cudaSuccess ||
            cudaFuncSetAttribute(linear_tc_kernel<NSPLIT, 128, 128, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024) != cudaSuccess ||
            cudaFuncSetAttribute(linear_tc_kernel<NSPLIT, 0, 0, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024) != cudaSuccess) {
            cudaGetLastError();
            return NPF_ENOTSUP;
        }
        reserved = 220 * 1024;
    }
    if (smem > 220 * 1024) return NPF_ENOTSUP;
    p.n_tiles = (int)cdiv(p.M, 128);
    // CTAs per SM limited by shared memory; stay persistent with one CTA per resident slot
    const int per_sm = 1;   // 512 threads + ~200 KB of shared memory: one persistent CTA per SM
    int grid = kNumSMs * per_sm;
    if (grid > p.n_tiles) grid = p.n_tiles;
    const bool hot = p.KR == 128 && p.NO == 128;
    if (hot && !p.u && !p.mask) linear_tc_kernel<NSPLIT, 128, 128, false, false><<<grid, kLinThreads, smem, st>>>(p);
    else if (hot && !p.u) linear_tc_kernel<NSPLIT, 128, 128, false, true><<<grid, kLinThreads, smem, st>>>(p);
    else if (hot && !p.mask) linear_tc_kernel<NSPLIT, 128, 128, true, false><<<grid, kLinThreads, smem, st>>>(p);
    else linear_tc_kernel<NSPLIT, 0, 0, true, true><<<grid, kLinThreads, smem, st>>>(p);
    count_launch();
    return check_launch("linear_tc_kernel");
}

int linear_fwd_tc(const float* X, int ldx, const float* W, int ldw, const float* b, float* Y, int ldy, int M, int K,
                  int N, int flags, const float* u, const float* w2, int ldw2, int precision, cudaStream_t st) {
    if (!tc_shape_ok(K, N) || (flags & NPF_ACCUM)) return NPF_ENOTSUP;
    TcLinParams p{};
    p.A = X; p.lda = ldx; p.W = W; p.ldw = ldw; p.C = Y; p.ldc = ldy; p.bias = b;
    p.u = u; p.w2 = w2; p.ldw2 = ldw2;
    p.M = M; p.KR = K; p.NO = N;
    p.relu_in = (flags & NPF_RELU_IN) ? 1 : 0; p.relu_out = (flags & NPF_RELU_OUT) ? 1 : 0;
    p.transposed_w = 0;
    p.a_vec = (ldx % 4 == 0) && aligned16(X);
    p.w_vec = (ldw % 4 == 0) && aligned16(W);
    p.c_vec = (ldy % 4 == 0) && aligned16(Y);
    return precision == NPF_PREC_BF16X3 ? launch_lin<3>(p, st) : launch_lin<1>(p, st);
}

int linear_bwd_data_tc(const float* dY, int lddy, const float* W, int ldw, float* dX, int lddx, int M, int K, int N,
                       const float* mask_src, int ldm, int flags, int precision, cudaStream_t st) {
    if (!tc_shape_ok(N, K) || (flags & NPF_ACCUM)) return NPF_ENOTSUP;
    TcLinParams p{};
    p.A = dY; p.lda = lddy; p.W = W; p.ldw = ldw; p.C = dX; p.ldc = lddx;
    p.mask = mask_src; p.ldm = ldm;
    p.M = M; p.KR = N; p.NO = K;
    p.transposed_w = 1;
    p.a_vec = (lddy % 4 == 0) && aligned16(dY);
    p.w_vec = (ldw % 4 == 0) && aligned16(W);
    p.c_vec = (lddx % 4 == 0) && aligned16(dX);
    return precision == NPF_PREC_BF16X3 ? launch_lin<3>(p, st) : launch_lin<1>(p, st);
}

template <int NSPLIT>
static int launch_wg(TcWgParams& p, cudaStream_t st) {
    const size_t smem = (size_t)(NSPLIT == 3 ? 2 : 1) * (kWgRows * 128 + kWgRows * (p.K + (p.db ? 16 : 0))) * 2;
    static bool attr = false;
    if (!attr) {
        if (cudaFuncSetAttribute(wgrad_tc_kernel<NSPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024) != cudaSuccess) {
            cudaGetLastError();
            return NPF_ENOTSUP;
        }
        attr = true;
    }
    if (smem > 220 * 1024) return NPF_ENOTSUP;
    int per_sm = (int)((220 * 1024) / (smem + 1024));
    if (per_sm < 1) per_sm = 1;
    if (per_sm > 2) per_sm = 2;
    long ctas = (long)kNumSMs * per_sm;
    long rows = cdiv(cdiv(p.M, ctas), kWgRows) * kWgRows;
    if (rows < kWgRows) rows = kWgRows;
    p.rows_per_cta = rows;
    const long grid = cdiv(p.M, rows);
    wgrad_tc_kernel<NSPLIT><<<(unsigned)grid, 256, smem, st>>>(p);
    count_launch();
    return check_launch("wgrad_tc_kernel");
}

int linear_bwd_weight_tc(const float* dY, int lddy, const float* X, int ldx, float* dW, int lddw, float* db, int* db_done, int M,
                         int K, int N, int flags, int precision, cudaStream_t st) {
    // MMA M dimension = N (rows of dW, padded to 128), MMA N dimension = K, reduction over the M rows
    if (N > 128 || N < 8 || N % 8 != 0 || !pow2(K) || K < 16 || K > 256) return NPF_ENOTSUP;
    TcWgParams p{};
    p.dY = dY; p.lddy = lddy; p.X = X; p.ldx = ldx; p.dW = dW; p.lddw = lddw;
    p.db = (db && K + 16 <= 256) ? db : nullptr;     // fused bias gradient needs MMA N = K + 16 <= 256
    *db_done = p.db != nullptr;
    p.M = M; p.N = N; p.K = K;
    p.relu_in = (flags & NPF_RELU_IN) ? 1 : 0;
    p.dy_vec = (lddy % 4 == 0) && aligned16(dY);
    p.x_vec = (ldx % 4 == 0) && aligned16(X);
    return precision == NPF_PREC_BF16X3 ? launch_wg<3>(p, st) : launch_wg<1>(p, st);
}

}  // namespace npf
