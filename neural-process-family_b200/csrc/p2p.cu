// One-shot all-reduce (mean) of the flat gradient bucket over NVLink peer memory: the only exchange step of the data-parallel
// path.  The bucket is 0.5-2 MB: an NCCL all-reduce of that size is pure latency (~45 us at 8 ranks,
// measured), so every rank simply READS the other ranks' buckets through peer mappings (cudaIpc handles, one process per GPU)
// and writes the mean into its own output buffer -- one kernel, no staging copies:
//     signal "my bucket is complete" to every peer  ->  wait for all peers  ->  out[i] = (sum_r in_r[i]) / world
//     ->  signal "I have finished reading"  ->  wait for all peers (a rank's bucket may be overwritten once its kernel ends).
// Signals are monotonically increasing epoch numbers written with system-scope release stores into the peers' signal blocks and
// polled with acquire loads (bounded: a lost signal traps instead of hanging); the epoch lives in device memory and is advanced by
// the kernel itself, so the launch is identical every step (CUDA-graph capturable, no host round trip).
#include "common.cuh"

namespace npf {

__device__ __forceinline__ void st_release_sys(int* p, int v) { asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ int ld_acquire_sys(const int* p) {
    int v;
    asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ float4 ld_sys_v4(const float* p) {       // never served from a stale L1 line
    float4 v;
    asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void wait_ge(const int* p, int epoch) {
    for (unsigned it = 0; it < (1u << 26); ++it) {
        if (ld_acquire_sys(p) >= epoch) return;
        __nanosleep(32);
    }
    __trap();
}

constexpr int kArMaxWorld = 16;
struct ArParams {
    float* outp[kArMaxWorld];         // every rank's output buffer (peer mappings; outp[rank] == out): two-shot variant only
    const float* in[kArMaxWorld];     // every rank's bucket (peer mappings; in[rank] is local)
    int* sig[kArMaxWorld];            // every rank's signal block: [0, world) "bucket ready" from rank r, [world, 2 world) "done reading" from rank r
    float* out;
    int* state;                       // local: [0] epoch of the last completed all-reduce, [1] CTAs finished in this launch
    int rank, world;
    long n4;                          // float4 elements
};

__global__ void __launch_bounds__(256) allreduce_mean_p2p_kernel(ArParams p) {
    __shared__ int s_epoch;
    if (threadIdx.x == 0) s_epoch = *reinterpret_cast<volatile int*>(p.state) + 1;
    __syncthreads();
    const int epoch = s_epoch;
    int* my_sig = p.sig[p.rank];
    if (blockIdx.x == 0 && threadIdx.x < p.world) st_release_sys(p.sig[threadIdx.x] + p.rank, epoch);     // my bucket is complete (stream order)
    if (threadIdx.x < p.world) wait_ge(my_sig + threadIdx.x, epoch);                                       // everybody's is
    __syncthreads();
    const float inv = 1.f / (float)p.world;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < p.n4; i += (long)gridDim.x * blockDim.x) {
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
        for (int r = 0; r < p.world; ++r) {
            const float4 v = ld_sys_v4(p.in[r] + 4 * i);
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
        reinterpret_cast<float4*>(p.out)[i] = make_float4(acc.x * inv, acc.y * inv, acc.z * inv, acc.w * inv);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const int prev = atomicAdd(p.state + 1, 1);
        if (prev == (int)gridDim.x - 1) {                 // last CTA of this rank: every read of the peers' buckets has been issued and consumed
            p.state[1] = 0;
            for (int r = 0; r < p.world; ++r) st_release_sys(p.sig[r] + p.world + p.rank, epoch);
            for (int r = 0; r < p.world; ++r) wait_ge(my_sig + p.world + r, epoch);    // nobody still reads MY bucket when this kernel ends
            p.state[0] = epoch;
            __threadfence();
        }
    }
}

// Two-shot variant (reduce-scatter + all-gather in one launch): rank r reduces only slice r of the bucket (reads (world - 1) / world
// of it from the peers) and WRITES the mean of that slice into every rank's output buffer.  Per rank 2 (world - 1) / world bucket
// sizes cross NVLink instead of (world - 1): at 8 ranks 0.96 MB instead of 3.85 MB.  Same two barriers: nobody reads a bucket
// before it is complete, and a kernel ends only when every rank has finished writing (hence reading) everything.
__global__ void __launch_bounds__(256) allreduce_mean_p2p2_kernel(ArParams p) {
    __shared__ int s_epoch;
    if (threadIdx.x == 0) s_epoch = *reinterpret_cast<volatile int*>(p.state) + 1;
    __syncthreads();
    const int epoch = s_epoch;
    int* my_sig = p.sig[p.rank];
    if (blockIdx.x == 0 && threadIdx.x < p.world) st_release_sys(p.sig[threadIdx.x] + p.rank, epoch);
    if (threadIdx.x < p.world) wait_ge(my_sig + threadIdx.x, epoch);
    __syncthreads();
    const float inv = 1.f / (float)p.world;
    const long per = (p.n4 + p.world - 1) / p.world;
    const long lo = per * p.rank, hi = min(p.n4, lo + per);
    for (long i = lo + (long)blockIdx.x * blockDim.x + threadIdx.x; i < hi; i += (long)gridDim.x * blockDim.x) {
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
        for (int r = 0; r < p.world; ++r) {
            const float4 v = ld_sys_v4(p.in[r] + 4 * i);
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
        const float4 m = make_float4(acc.x * inv, acc.y * inv, acc.z * inv, acc.w * inv);
#pragma unroll 4
        for (int r = 0; r < p.world; ++r) reinterpret_cast<float4*>(p.outp[r])[i] = m;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence_system();                           // this CTA's remote writes are visible system-wide before the signal below
        const int prev = atomicAdd(p.state + 1, 1);
        if (prev == (int)gridDim.x - 1) {
            p.state[1] = 0;
            __threadfence_system();
            for (int r = 0; r < p.world; ++r) st_release_sys(p.sig[r] + p.world + p.rank, epoch);     // my slice is in everybody's output
            for (int r = 0; r < p.world; ++r) wait_ge(my_sig + p.world + r, epoch);                    // everybody's slice is in mine
            p.state[0] = epoch;
            __threadfence();
        }
    }
}

}  // namespace npf

using namespace npf;

extern "C" int npf_p2p_alloc(void** ptr, size_t bytes) {
    NPF_REQUIRE(ptr && bytes > 0, "npf_p2p_alloc: bad arguments");
    if (cudaMalloc(ptr, bytes) != cudaSuccess || cudaMemset(*ptr, 0, bytes) != cudaSuccess) return check_launch("npf_p2p_alloc");
    return NPF_OK;
}
extern "C" int npf_p2p_free(void* ptr) { return cudaFree(ptr) == cudaSuccess ? NPF_OK : check_launch("npf_p2p_free"); }
extern "C" int npf_p2p_get_handle(void* ptr, unsigned char* handle64) {
    NPF_REQUIRE(ptr && handle64, "npf_p2p_get_handle: null pointer");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    cudaIpcMemHandle_t h;
    if (cudaIpcGetMemHandle(&h, ptr) != cudaSuccess) return check_launch("npf_p2p_get_handle");
    memcpy(handle64, &h, 64);
    return NPF_OK;
}
extern "C" int npf_p2p_open(const unsigned char* handle64, void** ptr) {
    NPF_REQUIRE(ptr && handle64, "npf_p2p_open: null pointer");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    if (cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) return check_launch("npf_p2p_open");
    return NPF_OK;
}
extern "C" int npf_p2p_close(void* ptr) { return cudaIpcCloseMemHandle(ptr) == cudaSuccess ? NPF_OK : check_launch("npf_p2p_close"); }

extern "C" int npf_allreduce_mean_p2p2(const float* const* in, int* const* sig, float* const* out, int* state, int rank, int world, long n,
                                       npf_stream_t stream) {
    NPF_REQUIRE(in && sig && out && state, "npf_allreduce_mean_p2p2: null pointer");
    NPF_REQUIRE(world >= 1 && world <= kArMaxWorld && rank >= 0 && rank < world && n >= 0 && n % 4 == 0,
                "npf_allreduce_mean_p2p2: bad world / rank / length (n must be a multiple of 4)");
    if (n == 0) return NPF_OK;
    ArParams p{};
    for (int r = 0; r < world; ++r) {
        NPF_REQUIRE(in[r] && sig[r] && out[r], "npf_allreduce_mean_p2p2: null peer pointer");
        p.in[r] = in[r]; p.sig[r] = sig[r]; p.outp[r] = out[r];
    }
    p.out = out[rank]; p.state = state; p.rank = rank; p.world = world; p.n4 = n / 4;
    const long want = cdiv(cdiv(p.n4, world), 256);
    const int grid = (int)(want < 32 ? (want < 1 ? 1 : want) : 32);
    allreduce_mean_p2p2_kernel<<<grid, 256, 0, as_stream(stream)>>>(p);
    count_launch();
    return check_launch("allreduce_mean_p2p2_kernel");
}

extern "C" int npf_allreduce_mean_p2p(const float* const* in, int* const* sig, float* out, int* state, int rank, int world, long n,
                                      npf_stream_t stream) {
    NPF_REQUIRE(in && sig && out && state, "npf_allreduce_mean_p2p: null pointer");
    NPF_REQUIRE(world >= 1 && world <= kArMaxWorld && rank >= 0 && rank < world && n >= 0 && n % 4 == 0,
                "npf_allreduce_mean_p2p: bad world / rank / length (n must be a multiple of 4)");
    if (n == 0) return NPF_OK;
    ArParams p{};
    for (int r = 0; r < world; ++r) {
        NPF_REQUIRE(in[r] && sig[r], "npf_allreduce_mean_p2p: null peer pointer");
        p.in[r] = in[r]; p.sig[r] = sig[r];
    }
    p.out = out; p.state = state; p.rank = rank; p.world = world; p.n4 = n / 4;
    const long want = cdiv(p.n4, 256);
    const int grid = (int)(want < 64 ? (want < 1 ? 1 : want) : 64);
    allreduce_mean_p2p_kernel<<<grid, 256, 0, as_stream(stream)>>>(p);
    count_launch();
    return check_launch("allreduce_mean_p2p_kernel");
}
