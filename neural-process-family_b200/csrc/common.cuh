// Shared helpers for libnpf_b200 (sm_100a).  Error plumbing, launch accounting, small device utilities.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/npf_b200.h"

namespace npf {

void set_error(const char* fmt, ...);
void count_launch(int n = 1);

// returns NPF_OK or NPF_ECUDA after a kernel launch (no sync)
int check_launch(const char* what);

static inline cudaStream_t as_stream(npf_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

#define NPF_REQUIRE(cond, ...)            \
    do {                                  \
        if (!(cond)) {                    \
            npf::set_error(__VA_ARGS__);  \
            return NPF_EINVAL;            \
        }                                 \
    } while (0)

constexpr int kNumSMs = 148;  // B200

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// softplus with the same branch structure as ATen (threshold 20): log1p(exp(x)) for x <= 20, x otherwise
__device__ __forceinline__ float softplus_f(float x) { return x > 20.f ? x : log1pf(expf(x)); }
__device__ __forceinline__ float sigmoid_f(float x) { return 1.f / (1.f + expf(-x)); }

static inline long cdiv(long a, long b) { return (a + b - 1) / b; }

}  // namespace npf
