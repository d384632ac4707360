// How much HBM bandwidth does a PERSISTENT streaming kernel reach as a function of CTAs per SM and 16-byte loads in flight per
// thread?  (copy: read 1 + write 1, buffers of 256 MB, L2 flushed by size).  Build: nvcc -arch=sm_100a -O3 stream_probe.cu -o stream_probe
#include <cstdio>
#include <cuda_runtime.h>
template <int U>
__global__ void copy_kernel(const float4* __restrict__ a, float4* __restrict__ b, long n) {
    const long stride = (long)gridDim.x * blockDim.x;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride * U) {
        float4 v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) v[u] = (i + u * stride < n) ? __ldg(a + i + u * stride) : make_float4(0, 0, 0, 0);
#pragma unroll
        for (int u = 0; u < U; ++u) if (i + u * stride < n) b[i + u * stride] = v[u];
    }
}
template <int U>
__global__ void read_kernel(const float4* __restrict__ a, float* out, long n) {
    const long stride = (long)gridDim.x * blockDim.x;
    float s = 0.f;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride * U) {
        float4 v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) v[u] = (i + u * stride < n) ? __ldg(a + i + u * stride) : make_float4(0, 0, 0, 0);
#pragma unroll
        for (int u = 0; u < U; ++u) s += v[u].x + v[u].y + v[u].z + v[u].w;
    }
    if (s == 123.456f) *out = s;
}
template <int U>
void run(const char* name, int ctas_per_sm, int threads, float4* a, float4* b, long n, bool copy) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int grid = 148 * ctas_per_sm;
    float best = 1e9f;
    for (int r = 0; r < 5; ++r) {
        cudaEventRecord(e0);
        if (copy) copy_kernel<U><<<grid, threads>>>(a, b, n); else read_kernel<U><<<grid, threads>>>(a, (float*)b, n);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    const double bytes = (copy ? 2.0 : 1.0) * n * 16;
    printf("%-5s U=%2d ctas/SM=%d threads=%4d  in-flight/SM=%6.1f KB  %7.1f GB/s\n", name, U, ctas_per_sm, threads, ctas_per_sm * threads * U * 16 / 1024.0, bytes / best / 1e6);
}
int main() {
    const long n = (256L << 20) / 16;
    float4 *a, *b; cudaMalloc(&a, n * 16); cudaMalloc(&b, n * 16); cudaMemset(a, 1, n * 16);
    for (int copy = 1; copy >= 0; --copy) {
        const char* nm = copy ? "copy" : "read";
        run<16>(nm, 1, 256, a, b, n, copy); run<16>(nm, 2, 256, a, b, n, copy); run<8>(nm, 2, 512, a, b, n, copy);
        run<4>(nm, 8, 256, a, b, n, copy); run<8>(nm, 8, 256, a, b, n, copy); run<16>(nm, 4, 256, a, b, n, copy); run<2>(nm, 8, 256, a, b, n, copy);
        run<4>(nm, 1, 1024, a, b, n, copy); run<8>(nm, 1, 1024, a, b, n, copy);
    }
    return 0;
}
