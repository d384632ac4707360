// tcgen05 / TMEM / mbarrier PTX wrappers and UMMA descriptor helpers shared by the tensor-core kernels
// (gemm_tc.cu, attention_tc.cu).  sm_100a only.
#pragma once
#include <cuda_bf16.h>

#include "common.cuh"

namespace npf {

// ------------------------------------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void tmem_alloc(uint32_t* slot, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");   // visible to the async proxy (tcgen05.commit)
}
// Bounded wait: a lost arrival traps (error return) instead of hanging the GPU.  The try_wait carries a suspend-time hint, so a
// waiting warp is parked by the hardware until the phase completes (or the hint expires) instead of hot-spinning: in the
// warp-specialised kernels ~20 of the 25 warps of a CTA are waiting at any time, and with a bare try_wait loop their polling
// took the issue slots the ONE MMA-issuing thread needs (measured with the timeline hook: 90-150 cycles per tcgen05.mma issued
// against 33-65 cycles of tensor time; profiles/r2/trace_*).  A short nanosleep backs the loop off further if the hint is not
// honoured.
#ifndef NPF_MBAR_HINT_NS
#define NPF_MBAR_HINT_NS 20000
#endif
#ifndef NPF_MBAR_OUTLINE
#define NPF_MBAR_OUTLINE 0
#endif
__device__ __forceinline__ uint32_t mbar_try_wait(uint32_t addr, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(addr), "r"(parity), "r"((uint32_t)NPF_MBAR_HINT_NS) : "memory");
    return ok;
}
// the retry loop, out of line: the (unrolled) poll / back-off code of ~25 call sites was most of these kernels' 60-100 KB of SASS
static __device__ __noinline__ void mbar_wait_slow(uint32_t addr, uint32_t parity) {
#pragma unroll 1
    for (uint32_t it = 0; it < (1u << 24); ++it) {
        if (mbar_try_wait(addr, parity)) return;
        if (it > 1) __nanosleep(64);
    }
    __trap();
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
#if NPF_MBAR_OUTLINE
    if (mbar_try_wait(addr, parity)) return;
    mbar_wait_slow(addr, parity);
#else
    for (uint32_t it = 0; it < (1u << 24); ++it) {
        if (mbar_try_wait(addr, parity)) return;
        if (it > 2) __nanosleep(64);
    }
    __trap();
#endif
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// arm the barrier's current phase with one arrival that expects `bytes` of asynchronous (bulk-copy) traffic
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// TMA 1-D bulk copy global -> shared (16-byte aligned addresses, size % 16 == 0); completion is counted on `bar`
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem], bf16 inputs, fp32 accumulate
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]: the A operand read from tensor memory (lane = row of A, one 32-bit column = two consecutive K elements,
// low half first; K-major only) -- no shared-memory read for A.
__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// 16 registers per thread -> 32 lanes x 16 consecutive 32-bit columns (thread = TMEM lane)
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
          "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// 32 lanes x 32 consecutive fp32 columns -> 32 registers per thread (thread = TMEM lane)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// Shared-memory matrix descriptor, SWIZZLE_NONE, descriptor version 1 (sm_100):
//   bits [0,14) start address >> 4, [16,30) leading-dim byte offset >> 4, [32,46) stride-dim byte offset >> 4,
//   [46,48) version = 1, [61,64) layout type = 0 (no swizzle).
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ull << 46);
}
// A descriptor whose start address is `bytes` further on (bytes % 16 == 0; the 14-bit address field cannot overflow: shared memory is
// < 256 KB).  The MMA loops build each operand's base descriptor ONCE and step through the k-slices with this single add: the
// issuing thread is one lane competing with 24 other warps for issue slots, and re-deriving every descriptor (mask, shift, three ORs
// in 64 bit) cost it more than the tensor core needed for the MMA.
__device__ __forceinline__ uint64_t desc_adv(uint64_t desc, uint32_t bytes) { return desc + (uint64_t)(bytes >> 4); }

// Instruction descriptor, kind::f16: c_format f32 (1 @ bit 4), a/b format bf16 (1 @ bits 7, 10), a_major @ 15,
// b_major @ 16 (1 = MN-major), N >> 3 @ bits [17,23), M >> 4 @ bits [24,29).
__host__ __device__ constexpr uint32_t make_idesc(int M, int N, int a_mn, int b_mn) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) |
           ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ float bf16_round(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }


// 32 lanes x 16 consecutive fp32 columns
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}


// 16 consecutive floats accumulated at a 4-byte aligned address: scalar atomics up to the next 16-byte boundary, float4 atomics for
// the aligned middle, scalars for the tail.  Rows of a [N, ld] gradient with ld % 4 != 0 (the 129-column SetConv resizer) cost 7 atomic
// operations per 16 values instead of 16.
template <int H>
__device__ __forceinline__ void atomic_add16_head(float* d, const float (&v)[16]) {
#pragma unroll
    for (int j = 0; j < H; ++j) atomicAdd(d + j, v[j]);
    constexpr int NV = H == 0 ? 4 : 3;
#pragma unroll
    for (int i = 0; i < NV; ++i)
        atomicAdd(reinterpret_cast<float4*>(d + H + 4 * i), make_float4(v[H + 4 * i], v[H + 4 * i + 1], v[H + 4 * i + 2], v[H + 4 * i + 3]));
#pragma unroll
    for (int j = H + 4 * NV; j < 16; ++j) atomicAdd(d + j, v[j]);
}
__device__ __forceinline__ void atomic_add16(float* d, const float (&v)[16]) {
    switch (((16u - (unsigned)(reinterpret_cast<uintptr_t>(d) & 15u)) >> 2) & 3u) {
        case 0: atomic_add16_head<0>(d, v); break;
        case 1: atomic_add16_head<1>(d, v); break;
        case 2: atomic_add16_head<2>(d, v); break;
        default: atomic_add16_head<3>(d, v); break;
    }
}

}  // namespace npf
