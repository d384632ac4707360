// On-grid context encoding of GridConvCNP / GridConvLNP (upstream npf/neuralproc/gridconvnp.py:136-162 with the
// abs-weight depthwise conv of npf/utils/helpers.py:316-331):
//   sig = conv_|w|(Y * M), den = conv_|w|(M);  feat = [ sig / max(den, 1e-5) ; den ]        (k x k, zero pad k/2)
// 1.5 MFLOP per 32x32 image: a small streaming kernel; one thread per (pixel, y-channel).
#include "common.cuh"

namespace npf {

__global__ void __launch_bounds__(256) gridconv_in_fwd_kernel(const float* __restrict__ img, const uint8_t* __restrict__ mask, int mc,
                                                              const float* __restrict__ Wt, float* __restrict__ feat, int B, int H,
                                                              int Wd, int y, int k) {
    extern __shared__ float wabs[];  // [y][k][k]
    for (int i = threadIdx.x; i < y * k * k; i += blockDim.x) wabs[i] = fabsf(__ldg(Wt + i));
    __syncthreads();
    const long n = (long)B * H * Wd * y;
    const int p = k / 2;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        const int j = (int)(i % y);
        long r = i / y;
        const int w = (int)(r % Wd); r /= Wd;
        const int h = (int)(r % H);
        const long b = r / H;
        float sig = 0.f, den = 0.f;
        for (int di = 0; di < k; ++di) {
            const int hh = h + di - p;
            if (hh < 0 || hh >= H) continue;
            for (int dj = 0; dj < k; ++dj) {
                const int ww = w + dj - p;
                if (ww < 0 || ww >= Wd) continue;
                const long pix = (b * H + hh) * Wd + ww;
                const float m = mask[pix * mc + (mc == 1 ? 0 : j)] ? 1.f : 0.f;
                const float wv = wabs[(j * k + di) * k + dj];
                den = fmaf(wv, m, den);
                sig = fmaf(wv, m * __ldg(img + pix * y + j), sig);
            }
        }
        const long o = (i / y) * 2 * y;
        feat[o + j] = sig / fmaxf(den, 1e-5f);
        feat[o + y + j] = den;
    }
}

// dWt[j,di,dj] += sign(w) * sum_{b,h,w} [ dsig * (img*m)(h+di-p, w+dj-p) + dden * m(h+di-p, w+dj-p) ]
//   dsig = dfeat1 / cd ;  dden = dfeat2 - (den > 1e-5 ? dfeat1 * feat1 / cd : 0) ;  cd = max(den, 1e-5)
// one CTA per (image, y-channel); threads own filter taps and sweep the image.
__global__ void __launch_bounds__(128) gridconv_in_bwd_kernel(const float* __restrict__ img, const uint8_t* __restrict__ mask, int mc,
                                                              const float* __restrict__ Wt, const float* __restrict__ feat,
                                                              const float* __restrict__ dfeat, float* __restrict__ dWt, int H, int Wd,
                                                              int y, int k) {
    extern __shared__ float sm[];
    const int p = k / 2;
    const int Hp = H + 2 * p, Wp = Wd + 2 * p;
    float* xm = sm;                  // [Hp][Wp]  img*m, zero padded
    float* mm = xm + Hp * Wp;        // [Hp][Wp]  m
    float* ds = mm + Hp * Wp;        // [H][Wd]
    float* dd = ds + H * Wd;         // [H][Wd]
    const int j = blockIdx.x, b = blockIdx.y;
    for (int i = threadIdx.x; i < Hp * Wp; i += blockDim.x) {
        const int hh = i / Wp - p, ww = i % Wp - p;
        float m = 0.f, x = 0.f;
        if (hh >= 0 && hh < H && ww >= 0 && ww < Wd) {
            const long pix = ((long)b * H + hh) * Wd + ww;
            m = mask[pix * mc + (mc == 1 ? 0 : j)] ? 1.f : 0.f;
            x = m * __ldg(img + pix * y + j);
        }
        xm[i] = x; mm[i] = m;
    }
    for (int i = threadIdx.x; i < H * Wd; i += blockDim.x) {
        const long o = ((long)b * H * Wd + i) * 2 * y;
        const float f1 = __ldg(feat + o + j), den = __ldg(feat + o + y + j);
        const float g1 = __ldg(dfeat + o + j), g2 = __ldg(dfeat + o + y + j);
        const float cd = fmaxf(den, 1e-5f);
        ds[i] = g1 / cd;
        dd[i] = g2 - (den > 1e-5f ? g1 * f1 / cd : 0.f);
    }
    __syncthreads();
    for (int t = threadIdx.x; t < k * k; t += blockDim.x) {
        const int di = t / k, dj = t % k;
        float a = 0.f;
        for (int h = 0; h < H; ++h) {
            const float* xr = xm + (h + di) * Wp + dj;
            const float* mr = mm + (h + di) * Wp + dj;
            const float* dsr = ds + h * Wd;
            const float* ddr = dd + h * Wd;
            for (int w = 0; w < Wd; ++w) a = fmaf(dsr[w], xr[w], fmaf(ddr[w], mr[w], a));
        }
        const float wv = __ldg(Wt + (long)j * k * k + t);
        const float sgn = wv > 0.f ? 1.f : (wv < 0.f ? -1.f : 0.f);  // d|w|/dw, torch.abs backward (0 at 0)
        atomicAdd(dWt + (long)j * k * k + t, sgn * a);
    }
}

}  // namespace npf

using namespace npf;

extern "C" int npf_gridconv_in_fwd(const float* img, const uint8_t* mask, int mc, const float* Wt, float* feat, int B, int H,
                                   int Wd, int y, int k, npf_stream_t stream) {
    NPF_REQUIRE(img && mask && Wt && feat, "npf_gridconv_in_fwd: null pointer");
    NPF_REQUIRE(B >= 0 && H >= 1 && Wd >= 1 && y >= 1 && k >= 1 && (k & 1), "npf_gridconv_in_fwd: bad shape");
    NPF_REQUIRE(mc == 1 || mc == y, "npf_gridconv_in_fwd: mask channels must be 1 or y");
    if (B == 0) return NPF_OK;
    const long n = (long)B * H * Wd * y;
    long blocks = cdiv(n, 256);
    if (blocks > 16L * kNumSMs) blocks = 16L * kNumSMs;
    const size_t smem = sizeof(float) * (size_t)y * k * k;
    NPF_REQUIRE(smem <= 48 * 1024, "npf_gridconv_in_fwd: filter bank too large");
    gridconv_in_fwd_kernel<<<(unsigned)blocks, 256, smem, as_stream(stream)>>>(img, mask, mc, Wt, feat, B, H, Wd, y, k);
    count_launch();
    return check_launch("gridconv_in_fwd_kernel");
}

extern "C" int npf_gridconv_in_bwd(const float* img, const uint8_t* mask, int mc, const float* Wt, const float* feat,
                                   const float* dfeat, float* dWt, int B, int H, int Wd, int y, int k, npf_stream_t stream) {
    NPF_REQUIRE(img && mask && Wt && feat && dfeat && dWt, "npf_gridconv_in_bwd: null pointer");
    NPF_REQUIRE(B >= 0 && H >= 1 && Wd >= 1 && y >= 1 && k >= 1 && (k & 1), "npf_gridconv_in_bwd: bad shape");
    NPF_REQUIRE(mc == 1 || mc == y, "npf_gridconv_in_bwd: mask channels must be 1 or y");
    NPF_REQUIRE(B <= 65535, "npf_gridconv_in_bwd: batch > 65535");
    if (B == 0) return NPF_OK;
    const int p = k / 2;
    const size_t smem = sizeof(float) * (2 * (size_t)(H + 2 * p) * (Wd + 2 * p) + 2 * (size_t)H * Wd);
    if (smem > 200 * 1024) { set_error("npf_gridconv_in_bwd: image %dx%d too large for the shared-memory tile", H, Wd); return NPF_ENOTSUP; }
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(gridconv_in_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        attr_set = true;
    }
    dim3 grid((unsigned)y, (unsigned)B);
    gridconv_in_bwd_kernel<<<grid, 128, smem, as_stream(stream)>>>(img, mask, mc, Wt, feat, dfeat, dWt, H, Wd, y, k);
    count_launch();
    return check_launch("gridconv_in_bwd_kernel");
}
