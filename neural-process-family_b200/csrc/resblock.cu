// One kernel per direction for the 1-D pre-activation residual block of the ConvCNP CNN (upstream npf/architectures/cnn.py:204-215,
// n_conv_layers = 1, Normalization = Identity: the constructor default of ConvCNP):
//
//     O[l, :] = sum_j wdw[:, j] (.) relu(X[l + j - p, :]) + bdw + X[l, :]        depthwise k taps (zero padding) + residual
//     Y[l, :] = Wpw . O[l, :] + bpw                                               pointwise 128 -> 128
//
// The unfused path runs this as npf_dwconv_fwd + npf_linear_fwd: O makes a round trip through HBM (write 50 MB, read 50 MB
// at config 2) between two launches.  Here a persistent CTA walks 128-row tiles of one task at a time:
//   * ONE thread issues a TMA bulk copy (cp.async.bulk, 1-D: the tile's rows and their +-p halo are contiguous in the
//     channel-last layout) of the raw fp32 rows into shared memory; rows outside the task are zero-filled (padding);
//   * 16 producer warps run the depthwise conv out of that raw tile (thread = 2 channels x 16 rows, the k taps of its two
//     channels in registers for the whole kernel), add bias + residual, and write O split into bf16 hi / lo straight into the
//     SWIZZLE_128B K-major A-operand image (optionally also as fp32 rows to HBM when the caller wants O saved);
//   * 1 MMA warp multiplies by the once-staged pointwise weights (tcgen05, hi.hi + hi.lo + lo.hi, accumulator double-
//     buffered in TMEM), 8 epilogue warps add the bias and store coalesced rows.
// HBM sees X once (+ 2p / 128 halo re-reads out of L2) and Y once.
#include <cstdlib>

#include "tc_common.cuh"

namespace npf {

constexpr int kRbProd = 16;
constexpr int kRbMmaWarp = kRbProd;
constexpr int kRbEpiWarp0 = kRbProd + 1;
constexpr int kRbEpi = 8;
constexpr int kRbThreads = (kRbEpiWarp0 + kRbEpi) * 32;       // 800
constexpr int kRbMaxPad = 9;                                   // k <= 19
constexpr int kRbRawRows = 128 + 2 * kRbMaxPad;                // 146
constexpr uint32_t kRbTile = 128u * 128u * 2u;                 // one bf16 128 x 128 image: 32 KB
constexpr int kRbScratchLd = 20;

struct RbFwdParams {
    const float* X;      // [B, L, 128]
    const float* wdw;    // [128, k]
    const float* bdw;    // [128] or null
    const float* wpw;    // [128, 128]
    const float* bpw;    // [128] or null
    float* O;            // [B, L, 128] or null
    float* Y;            // [B, L, 128]
    int B, L, n_lt, n_tiles;
};

__device__ __forceinline__ uint64_t rb_desc_sw128(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return make_desc(saddr, lbo_bytes, sbo_bytes) | (2ull << 61);
}
// byte offset of element (row, col) of a [128 x 128] bf16 SWIZZLE_128B K-major image (two 64-column atoms of 16 KB)
__device__ __forceinline__ uint32_t rb_img_off(uint32_t row, uint32_t col) {
    return (col >> 6) * 16384u + row * 128u + ((((col & 63u) >> 3) ^ (row & 7u)) << 4) + (col & 7u) * 2u;
}
__device__ __forceinline__ void rb_prod_sync() { asm volatile("bar.sync 1, %0;" ::"n"(kRbProd * 32) : "memory"); }

template <int KW>
__global__ void __launch_bounds__(kRbThreads, 1) resblock1d_fwd_kernel(RbFwdParams p) {
    constexpr int P = KW / 2;
    constexpr int RAW = 128 + 2 * P;                            // raw rows of a tile: row i <-> position l0 - P + i
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar_raw, bar_afull, bar_aempty, bar_tfull[2], bar_tempty[2];
    __shared__ uint32_t tmem_slot;
    __shared__ __align__(16) float s_bias[128];

    uint8_t* a_hi = smem_raw;                                   // O image (A operand)
    uint8_t* a_lo = a_hi + kRbTile;
    uint8_t* b_hi = smem_raw + 2 * kRbTile;                     // pointwise weights (B operand)
    uint8_t* b_lo = b_hi + kRbTile;
    float* raw = reinterpret_cast<float*>(smem_raw + 4 * kRbTile);          // [RAW][128] fp32
    float* scratch_all = raw + kRbRawRows * 128;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (warp == 0) tmem_alloc(&tmem_slot, 256);
    if (tid == 32) {
        mbar_init(&bar_raw, 1);
        mbar_init(&bar_afull, kRbProd * 32);
        mbar_init(&bar_aempty, 1);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&bar_tfull[i], 1);
            mbar_init(&bar_tempty[i], kRbEpi * 32);
        }
    }
    if (tid < 128) s_bias[tid] = p.bpw ? __ldg(p.bpw + tid) : 0.f;
    // contiguous, balanced tile ranges (tiles of one task stay together: halo rows come out of L2)
    const int per = p.n_tiles / (int)gridDim.x, rem = p.n_tiles - per * (int)gridDim.x;
    const int g0 = (int)blockIdx.x * per + min((int)blockIdx.x, rem), g1 = g0 + per + ((int)blockIdx.x < rem ? 1 : 0);
    pdl_trigger();

    // pointwise weights [128 x 128] fp32 row-major, staged once per CTA by the 16 producer warps (parameters: before pdl_wait)
    if (warp < kRbProd) {
        const uint32_t pchunk = (uint32_t)(lane >> 1) & 7u;
        const uint32_t psoff = (uint32_t)(lane >> 4) * 16384u + (uint32_t)(warp * 8) * 128u + (uint32_t)(lane & 1) * 8u;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float4 v = __ldg(reinterpret_cast<const float4*>(p.wpw + (long)(warp * 8 + i) * 128) + lane);
            const uint32_t off = psoff + (uint32_t)i * 128u + ((pchunk ^ (uint32_t)i) << 4);
            const uint32_t h01 = pack_bf16(v.x, v.y), h23 = pack_bf16(v.z, v.w);
            *reinterpret_cast<uint2*>(b_hi + off) = make_uint2(h01, h23);
            *reinterpret_cast<uint2*>(b_lo + off) = make_uint2(pack_bf16(v.x - __uint_as_float(h01 << 16), v.y - __uint_as_float(h01 & 0xFFFF0000u)),
                                                                 pack_bf16(v.z - __uint_as_float(h23 << 16), v.w - __uint_as_float(h23 & 0xFFFF0000u)));
        }
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    pdl_wait();

    if (warp < kRbProd) {
        // ------------------------------------------------------------------ producers: raw tile -> depthwise -> O image
        const int cp = lane + 32 * (warp & 1);                   // channel pair: channels 2 cp, 2 cp + 1
        const int rg = warp >> 1;                                // rows 16 rg .. 16 rg + 15 of the tile
        float w0[KW], w1[KW];
#pragma unroll
        for (int j = 0; j < KW; ++j) { w0[j] = __ldg(p.wdw + (2 * cp) * KW + j); w1[j] = __ldg(p.wdw + (2 * cp + 1) * KW + j); }
        const float bd0 = p.bdw ? __ldg(p.bdw + 2 * cp) : 0.f, bd1 = p.bdw ? __ldg(p.bdw + 2 * cp + 1) : 0.f;

        // raw rows of tile g: TMA bulk copy of the in-task part, zero fill of the rest (issued when the raw buffer is free)
        auto fetch = [&](int g) {
            const int b = g / p.n_lt, l0 = (g - b * p.n_lt) * 128;
            const int s0 = max(0, l0 - P), e0 = min(p.L, l0 + 128 + P);
            const int d0 = s0 - (l0 - P), d1 = d0 + (e0 - s0);                     // raw rows [d0, d1) come from HBM
            for (int i = tid; i < (d0 + (RAW - d1)) * 32; i += kRbProd * 32) {      // 32 float4 per row
                const int r = i >> 5, rr = r < d0 ? r : d1 + (r - d0);
                reinterpret_cast<float4*>(raw + rr * 128)[i & 31] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
            if (tid == 0) {
                fence_async_smem();                       // the generic reads of the previous tile are ordered before the async writes
                const uint32_t bytes = (uint32_t)(e0 - s0) * 512u;
                mbar_expect_tx(&bar_raw, bytes);
                bulk_g2s(raw + d0 * 128, p.X + ((long)b * p.L + s0) * 128, bytes, &bar_raw);
            }
        };
        if (g0 < g1) fetch(g0);
        int it = 0;
        for (int g = g0; g < g1; ++g, ++it) {
            const int b = g / p.n_lt, l0 = (g - b * p.n_lt) * 128;
            const int rows_ok = min(128, p.L - l0);
            mbar_wait(&bar_raw, (uint32_t)it & 1u);
            rb_prod_sync();                                       // the zero-filled rows of this tile are visible to every producer
            float a0[16], a1[16];
#pragma unroll
            for (int o = 0; o < 16; ++o) { a0[o] = bd0; a1[o] = bd1; }
            const float* rp = raw + (16 * rg) * 128 + 2 * cp;
#pragma unroll
            for (int i = 0; i < 16 + 2 * P; ++i) {                // raw row 16 rg + i feeds outputs o = i - j, tap j
                const float2 v = *reinterpret_cast<const float2*>(rp + i * 128);
                const float r0 = fmaxf(v.x, 0.f), r1 = fmaxf(v.y, 0.f);
#pragma unroll
                for (int j = 0; j < KW; ++j) {
                    const int o = i - j;
                    if (o >= 0 && o < 16) { a0[o] = fmaf(w0[j], r0, a0[o]); a1[o] = fmaf(w1[j], r1, a1[o]); }
                }
                if (i - P >= 0 && i - P < 16) { a0[i - P] += v.x; a1[i - P] += v.y; }          // residual: the block input itself
            }
            rb_prod_sync();                                       // every producer has finished reading the raw tile
            if (g + 1 < g1) fetch(g + 1);
            if (p.O) {                                            // O saved for a backward pass that does not recompute it
#pragma unroll
                for (int o = 0; o < 16; ++o) {
                    const int row = 16 * rg + o;
                    if (row < rows_ok) *reinterpret_cast<float2*>(p.O + ((long)b * p.L + l0 + row) * 128 + 2 * cp) = make_float2(a0[o], a1[o]);
                }
            }
            if (it > 0) mbar_wait(&bar_aempty, (uint32_t)(it - 1) & 1u);        // the MMAs of the previous tile have read the image
#pragma unroll
            for (int o = 0; o < 16; ++o) {
                const uint32_t row = (uint32_t)(16 * rg + o);
                const uint32_t off = rb_img_off(row, (uint32_t)(2 * cp));
                const uint32_t h = pack_bf16(a0[o], a1[o]);
                *reinterpret_cast<uint32_t*>(a_hi + off) = h;
                *reinterpret_cast<uint32_t*>(a_lo + off) = pack_bf16(a0[o] - __uint_as_float(h << 16), a1[o] - __uint_as_float(h & 0xFFFF0000u));
            }
            fence_async_smem();
            mbar_arrive(&bar_afull);
        }
    } else if (warp == kRbMmaWarp) {
        // ------------------------------------------------------------------ MMA issuer
        if (lane == 0) {
            const uint32_t idesc = make_idesc(128, 128, 0, 0);
            const uint32_t sa_hi = smem_u32(a_hi), sa_lo = smem_u32(a_lo), sb_hi = smem_u32(b_hi), sb_lo = smem_u32(b_lo);
            int it = 0;
            for (int g = g0; g < g1; ++g, ++it) {
                const int t = it & 1;
                mbar_wait(&bar_afull, (uint32_t)it & 1u);
                mbar_wait(&bar_tempty[t], (uint32_t)((it >> 1) & 1) ^ 1u);
                tc_fence_after();
                const uint32_t d = tmem + (uint32_t)t * 128u;
#pragma unroll
                for (int ks = 0; ks < 8; ++ks) {
                    const uint32_t ao = (uint32_t)(ks >> 2) * 16384u + (uint32_t)(ks & 3) * 32u;
                    const uint64_t a_h = rb_desc_sw128(sa_hi + ao, 16, 1024), b_h = rb_desc_sw128(sb_hi + ao, 16, 1024);
                    umma_bf16(d, a_h, b_h, idesc, ks ? 1u : 0u);
                    umma_bf16(d, a_h, rb_desc_sw128(sb_lo + ao, 16, 1024), idesc, 1);
                    umma_bf16(d, rb_desc_sw128(sa_lo + ao, 16, 1024), b_h, idesc, 1);
                }
                umma_commit(&bar_aempty);
                umma_commit(&bar_tfull[t]);
            }
        }
    } else {
        // ------------------------------------------------------------------ epilogue: TMEM -> + bias -> coalesced rows of Y
        const int e = warp - kRbEpiWarp0;
        const int lane_base = 32 * (warp & 3);
        const int col_base = (e >> 2) * 64;
        float* scratch = scratch_all + e * (32 * kRbScratchLd);
        const int r_in = lane >> 2, c4 = (lane & 3) * 4;
        int it = 0;
        for (int g = g0; g < g1; ++g, ++it) {
            const int t = it & 1;
            const int b = g / p.n_lt, l0 = (g - b * p.n_lt) * 128;
            const int rows_ok = min(128, p.L - l0) - lane_base;        // rows [0, rows_ok) of this warp's 32 exist
            float* yb = p.Y + ((long)b * p.L + l0 + lane_base) * 128;
            mbar_wait(&bar_tfull[t], (uint32_t)(it >> 1) & 1u);
            tc_fence_after();
#pragma unroll 1
            for (int ch = 0; ch < 4; ++ch) {
                const int c0 = col_base + ch * 16;
                float v[16];
                tmem_ld16(tmem + ((uint32_t)lane_base << 16) + (uint32_t)(t * 128 + c0), v);
                if (ch == 3) {
                    tc_fence_before();
                    mbar_arrive(&bar_tempty[t]);
                }
#pragma unroll
                for (int j = 0; j < 16; j += 4) *reinterpret_cast<float4*>(scratch + lane * kRbScratchLd + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                __syncwarp();
                const float4 bb = *reinterpret_cast<const float4*>(&s_bias[c0 + c4]);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int r = j * 8 + r_in;
                    if (r < rows_ok) {
                        float4 x = *reinterpret_cast<const float4*>(scratch + r * kRbScratchLd + c4);
                        x.x += bb.x; x.y += bb.y; x.z += bb.z; x.w += bb.w;
                        *reinterpret_cast<float4*>(yb + (long)r * 128 + c0 + c4) = x;
                    }
                }
                __syncwarp();
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 256);
}

template <int KW>
static int launch_rb_fwd(RbFwdParams& p, cudaStream_t st) {
    const size_t smem = (size_t)4 * kRbTile + (size_t)kRbRawRows * 512 + (size_t)kRbEpi * 32 * kRbScratchLd * sizeof(float);
    static bool attr = false;
    if (!attr) {
        if (cudaFuncSetAttribute(resblock1d_fwd_kernel<KW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
            cudaGetLastError();
            return NPF_ENOTSUP;
        }
        attr = true;
    }
    const int grid = p.n_tiles < kNumSMs ? p.n_tiles : kNumSMs;
    launch_pdl(resblock1d_fwd_kernel<KW>, dim3(grid), dim3(kRbThreads), smem, st, p);
    count_launch();
    return check_launch("resblock1d_fwd_kernel");
}

static inline bool rb_aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace npf

using namespace npf;

extern "C" int npf_resblock1d_fwd(const float* X, const float* wdw, const float* bdw, const float* wpw, const float* bpw, float* O, float* Y,
                                  int B, int L, int C, int k, int precision, npf_stream_t stream) {
    NPF_REQUIRE(X && wdw && wpw && Y, "npf_resblock1d_fwd: null pointer");
    NPF_REQUIRE(B >= 0 && L >= 1 && k >= 1 && (k & 1), "npf_resblock1d_fwd: bad shape (odd kernel size)");
    if (B == 0) return NPF_OK;
    static const bool on = [] { const char* e = getenv("NPF_RESBLOCK_FUSED"); return !(e && e[0] == '0'); }();
    if (!on || C != 128 || precision != NPF_PREC_BF16X3 || k != 11 || !rb_aligned16(X) || !rb_aligned16(Y) || !rb_aligned16(wpw) || (O && !rb_aligned16(O))) {
        set_error("npf_resblock1d_fwd: covered: 128 channels, kernel size 11, precision bf16x3, 16-byte aligned tensors; run npf_dwconv_fwd + npf_linear_fwd otherwise");
        return NPF_ENOTSUP;
    }
    RbFwdParams p{};
    p.X = X; p.wdw = wdw; p.bdw = bdw; p.wpw = wpw; p.bpw = bpw; p.O = O; p.Y = Y; p.B = B; p.L = L;
    p.n_lt = (L + 127) / 128;
    p.n_tiles = B * p.n_lt;
    return launch_rb_fwd<11>(p, as_stream(stream));
}
