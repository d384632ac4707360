// Throughput of FFMA vs FFMA2 (packed fp32x2) on sm_100a: N independent accumulator chains per thread, W warps per SM sub-partition.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma_probe ffma_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE, int ILP>
__global__ void probe(float* out, float a, float b, int iters) {
    float2 acc[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) acc[i] = make_float2(threadIdx.x * 0.001f + i, i * 0.5f);
    const float2 a2 = make_float2(a, a * 1.0001f), b2 = make_float2(b, b * 0.9999f);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            if (MODE == 0) {          // 2 scalar FFMA
                acc[i].x = fmaf(a2.x, acc[i].x, b2.x);
                acc[i].y = fmaf(a2.y, acc[i].y, b2.y);
            } else {                  // 1 FFMA2
                acc[i] = __ffma2_rn(a2, acc[i], b2);
            }
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += acc[i].x + acc[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE, int ILP>
void run(const char* name, int threads) {
    float* out;
    cudaMalloc(&out, 148 * 4 * 1024 * sizeof(float));
    const int iters = 20000;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    probe<MODE, ILP><<<148, threads>>>(out, 1.0001f, 0.0001f, 100);
    cudaEventRecord(e0);
    probe<MODE, ILP><<<148, threads>>>(out, 1.0001f, 0.0001f, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double fma = 148.0 * threads * (double)iters * ILP * 2;
    printf("%-6s ILP=%2d threads/SM=%4d : %.2f TFMA/s (%.1f TFLOP/s), %.1f FMA/clk/SM at 1.965 GHz\n", name, ILP, threads, fma / ms / 1e9, 2 * fma / ms / 1e9,
           fma / (ms * 1e-3) / 148 / 1.965e9);
    cudaFree(out);
}

int main() {
    run<0, 8>("FFMA", 128); run<1, 8>("FFMA2", 128);
    run<0, 8>("FFMA", 256); run<1, 8>("FFMA2", 256);
    run<0, 8>("FFMA", 512); run<1, 8>("FFMA2", 512);
    run<0, 8>("FFMA", 1024); run<1, 8>("FFMA2", 1024);
    run<0, 16>("FFMA", 512); run<1, 16>("FFMA2", 512);
    return 0;
}
