// Context / target split on the device: the step BEFORE the hot path.
//
// Upstream builds every meta-batch on the host: per row an independent np.random.shuffle of arange(N) truncated to n
// (GetRandomIndcs.__call__, npf/utils/datasplit.py:108-145, indep_shuffle_ npf/utils/helpers.py:82-96), torch.gather of
// X and Y (CntxtTrgtGetter.select 246-255), a scatter into a boolean mask (RandomMasker 259-278) and, for images fed to
// the set-based models, mask.nonzero() + coordinate normalisation (GridCntxtTrgtGetter.select 423-452).  Here the data
// stays in HBM: the subset is the first n entries of a partial Fisher-Yates shuffle driven by the counter-based
// Philox-4x32-10 generator, so every (row, draw) number is addressable and the CPU checker of the test-suite reproduces
// the indices bit for bit.  All kernels are integer / byte work bounded by launch latency at these sizes
// (B x N <= 1024 x 1024): one CTA per row, no tensor cores.
#include "common.cuh"

namespace npf {

// ------------------------------------------------------------------------------------------------ Philox-4x32-10
// Salmon, Moraes, Dror, Shaw: "Parallel random numbers: as easy as 1, 2, 3" (SC'11).  10 rounds, key bumped by the
// Weyl constants between rounds.
struct U4 {
    uint32_t x, y, z, w;
};

__host__ __device__ __forceinline__ U4 philox4x32_10(U4 c, uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        if (r > 0) {
            k0 += 0x9E3779B9u;
            k1 += 0xBB67AE85u;
        }
        const uint64_t p0 = (uint64_t)0xD2511F53u * c.x;
        const uint64_t p1 = (uint64_t)0xCD9E8D57u * c.z;
        U4 n;
        n.x = (uint32_t)(p1 >> 32) ^ c.y ^ k0;
        n.y = (uint32_t)p1;
        n.z = (uint32_t)(p0 >> 32) ^ c.w ^ k1;
        n.w = (uint32_t)p0;
        c = n;
    }
    return c;
}

// j uniform in [0, m): multiply-high (Lemire) without the rejection step; the bias is < m / 2^32 (< 3e-6 for m <= 12288)
__host__ __device__ __forceinline__ uint32_t bounded(uint32_t r, uint32_t m) { return (uint32_t)(((uint64_t)r * m) >> 32); }

// The n sequential swaps of the partial Fisher-Yates shuffle of row b (perm = arange(N) on entry): draw i is component
// i & 3 of philox(counter = {i >> 2, b, 0, 0}, key = seed).  __host__ too so that the build container (no GPU) can run exactly
// this code against the CPU checker of the test-suite.
__host__ __device__ __forceinline__ void partial_shuffle(int32_t* perm, int N, int n, uint32_t b, uint32_t k0, uint32_t k1) {
    U4 r = {0, 0, 0, 0};
    for (int i = 0; i < n; ++i) {
        if ((i & 3) == 0) r = philox4x32_10(U4{(uint32_t)(i >> 2), b, 0u, 0u}, k0, k1);
        const uint32_t x = (i & 3) == 0 ? r.x : (i & 3) == 1 ? r.y : (i & 3) == 2 ? r.z : r.w;
        const int j = i + (int)bounded(x, (uint32_t)(N - i));
        const int32_t a = perm[i];
        perm[i] = perm[j];
        perm[j] = a;
    }
}

// One CTA per row.  perm = arange(N) in shared memory; thread 0 runs partial_shuffle;
// MODE 0 writes indcs[b, 0..n) = perm[0..n) (the shuffled order, as upstream's indcs[:, :n]); MODE 1 writes the byte
// mask of RandomMasker (1 at the n chosen positions).
template <int MODE>
__global__ void random_subset_kernel(int32_t* __restrict__ indcs, uint8_t* __restrict__ mask, int N, int n, uint32_t k0, uint32_t k1) {
    extern __shared__ int32_t perm[];
    const int b = blockIdx.x;
    for (int t = threadIdx.x; t < N; t += blockDim.x) perm[t] = t;
    if (MODE == 1)
        for (int t = threadIdx.x; t < N; t += blockDim.x) mask[(long)b * N + t] = 0;
    __syncthreads();
    if (threadIdx.x == 0) partial_shuffle(perm, N, n, (uint32_t)b, k0, k1);
    __syncthreads();
    for (int t = threadIdx.x; t < n; t += blockDim.x) {
        if (MODE == 0)
            indcs[(long)b * n + t] = perm[t];
        else
            mask[(long)b * N + perm[t]] = 1;
    }
}

// out[b, i, :] = src[b, indcs[b, i], :] for X (xd features) and Y (yd values) in one pass
__global__ void select_points_kernel(const float* __restrict__ X, const float* __restrict__ Y, const int32_t* __restrict__ indcs,
                                     float* __restrict__ Xo, float* __restrict__ Yo, int B, int N, int n, int xd, int yd) {
    const int d = xd + yd;
    const long total = (long)B * n * d;
    for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
        const int c = (int)(e % d);
        const long bi = e / d;
        const int b = (int)(bi / n);
        int src = __ldg(indcs + bi);
        src = src < 0 ? 0 : (src >= N ? N - 1 : src);  // memory safety only: indices are required to lie in [0, N)
        if (c < xd)
            Xo[bi * xd + c] = __ldg(X + ((long)b * N + src) * xd + c);
        else
            Yo[bi * yd + (c - xd)] = __ldg(Y + ((long)b * N + src) * yd + (c - xd));
    }
}

// GridCntxtTrgtGetter.select: per row, the masked pixels in row-major order (the order of mask.nonzero()):
//   Xo[b, k, :] = normalised grid coordinates of the k-th masked pixel, Yo[b, k, :] = its values.
// 256 threads per row, each owning a contiguous slab of pixels: count, block-wide exclusive scan, write.
constexpr int kGridSelThreads = 256;
__global__ void grid_select_kernel(const uint8_t* __restrict__ mask, const float* __restrict__ img, float* __restrict__ Xo,
                                   float* __restrict__ Yo, int32_t* __restrict__ counts, int H, int W, int n_grid_dim, int yd, int n,
                                   float sh, float sw, float upscale) {
    __shared__ int s_scan[kGridSelThreads];
    const int b = blockIdx.x, t = threadIdx.x;
    const int P = H * W;
    const int slab = (P + kGridSelThreads - 1) / kGridSelThreads;
    const int p0 = t * slab, p1 = min(P, p0 + slab);
    const uint8_t* m = mask + (long)b * P;
    int cnt = 0;
    for (int p = p0; p < p1; ++p) cnt += m[p] != 0;
    s_scan[t] = cnt;
    __syncthreads();
    for (int o = 1; o < kGridSelThreads; o <<= 1) {   // Hillis-Steele inclusive scan
        const int v = t >= o ? s_scan[t - o] : 0;
        __syncthreads();
        s_scan[t] += v;
        __syncthreads();
    }
    int k = s_scan[t] - cnt;
    if (t == kGridSelThreads - 1 && counts) counts[b] = s_scan[t];
    for (int p = p0; p < p1; ++p) {
        if (!m[p]) continue;
        if (k < n) {
            const long o = (long)b * n + k;
            // two roundings per coordinate like upstream's in-place `*= 2/(size-1)`, `-= 1`, `*= upscale` (no FMA)
            if (n_grid_dim == 2) {
                Xo[o * 2 + 0] = __fmul_rn(__fsub_rn(__fmul_rn((float)(p / W), sh), 1.f), upscale);
                Xo[o * 2 + 1] = __fmul_rn(__fsub_rn(__fmul_rn((float)(p % W), sw), 1.f), upscale);
            } else {
                Xo[o] = __fmul_rn(__fsub_rn(__fmul_rn((float)p, sw), 1.f), upscale);
            }
            for (int c = 0; c < yd; ++c) Yo[o * yd + c] = __ldg(img + ((long)b * P + p) * yd + c);
        }
        ++k;
    }
}

}  // namespace npf

using namespace npf;

static const int kSubsetMaxN = 12288;  // perm[] of one row in the default 48 KB of shared memory

template <int MODE>
static int launch_subset(int32_t* indcs, uint8_t* mask, int B, int N, int n, unsigned long long seed, npf_stream_t stream, const char* who) {
    NPF_REQUIRE(B >= 0 && N >= 0 && n >= 0 && n <= N, "%s: need 0 <= n <= N (B=%d N=%d n=%d)", who, B, N, n);
    if (N > kSubsetMaxN) {
        set_error("%s: N=%d exceeds the %d points one CTA shuffles in shared memory", who, N, kSubsetMaxN);
        return NPF_ENOTSUP;
    }
    if (B == 0 || N == 0 || (MODE == 0 && n == 0)) return NPF_OK;
    random_subset_kernel<MODE><<<B, 128, (size_t)N * sizeof(int32_t), as_stream(stream)>>>(indcs, mask, N, n, (uint32_t)seed,
                                                                                          (uint32_t)(seed >> 32));
    count_launch();
    return check_launch("random_subset_kernel");
}

extern "C" int npf_random_subset(int32_t* indcs, int B, int N, int n, unsigned long long seed, npf_stream_t stream) {
    NPF_REQUIRE(indcs || (long)B * n == 0, "npf_random_subset: null pointer");
    return launch_subset<0>(indcs, nullptr, B, N, n, seed, stream, "npf_random_subset");
}

extern "C" int npf_random_mask(uint8_t* mask, int B, int P, int n, unsigned long long seed, npf_stream_t stream) {
    NPF_REQUIRE(mask || (long)B * P == 0, "npf_random_mask: null pointer");
    return launch_subset<1>(nullptr, mask, B, P, n, seed, stream, "npf_random_mask");
}

extern "C" int npf_select_points(const float* X, const float* Y, const int32_t* indcs, float* Xo, float* Yo, int B, int N, int n, int xd,
                                 int yd, npf_stream_t stream) {
    NPF_REQUIRE(B >= 0 && N >= 0 && n >= 0 && xd >= 1 && yd >= 1, "npf_select_points: bad shape");
    const long total = (long)B * n * (xd + yd);
    if (total == 0) return NPF_OK;
    NPF_REQUIRE(N >= 1, "npf_select_points: selecting %d points out of an empty set", n);
    NPF_REQUIRE(X && Y && indcs && Xo && Yo, "npf_select_points: null pointer");
    const long blocks = cdiv(total, 256);
    select_points_kernel<<<(unsigned)(blocks > 8 * kNumSMs ? 8 * kNumSMs : blocks), 256, 0, as_stream(stream)>>>(X, Y, indcs, Xo, Yo, B, N,
                                                                                                               n, xd, yd);
    count_launch();
    return check_launch("select_points_kernel");
}

extern "C" int npf_grid_select(const uint8_t* mask, const float* img, float* Xo, float* Yo, int32_t* counts, int B, int H, int W,
                               int n_grid_dim, int yd, int n, float upscale, npf_stream_t stream) {
    NPF_REQUIRE(B >= 0 && H >= 1 && W >= 1 && yd >= 1 && n >= 0 && (long)H * W < (1l << 30), "npf_grid_select: bad shape");
    NPF_REQUIRE(n_grid_dim == 2 || (n_grid_dim == 1 && H == 1), "npf_grid_select: 1-D (H == 1) or 2-D grids only");
    NPF_REQUIRE(n <= H * W, "npf_grid_select: n=%d exceeds the %d grid points", n, H * W);
    if (B == 0) return NPF_OK;
    NPF_REQUIRE(mask && img && (n == 0 || (Xo && Yo)), "npf_grid_select: null pointer");
    // 2 / (size - 1) in double, rounded once to fp32: the python scalar upstream multiplies the fp32 tensor by
    const float sh = H > 1 ? (float)(2.0 / (H - 1)) : 0.f, sw = W > 1 ? (float)(2.0 / (W - 1)) : 0.f;
    grid_select_kernel<<<B, kGridSelThreads, 0, as_stream(stream)>>>(mask, img, Xo, Yo, counts, H, W, n_grid_dim, yd, n, sh, sw, upscale);
    count_launch();
    return check_launch("grid_select_kernel");
}
