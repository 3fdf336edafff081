// ps_tma.cuh — bulk asynchronous copies (the TMA unit's 1-D form, cp.async.bulk) and the mbarrier that tracks them.
//
// The state records of the stepper are body-major (144 B per body, 16-byte aligned) and the bodies of a CTA's tile are
// contiguous, so a tile of the published / working / predicted state is ONE contiguous span: a single bulk copy moves it
// between HBM and shared memory in full lines, with no thread involved and no register staging (SASS: UBLKCP, SYNCS).
// Threads then read their 144-byte records from shared memory (16-byte accesses at a 144-byte stride: the eight lanes of
// a quarter warp fall on eight different 16-byte bank groups — conflict-free).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace pst_tma {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");  // visible to the async proxy before any copy names it
}
// one arrival + the number of bytes the copies issued after it will deliver
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity)
{
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1, 0x989680;\n\t"
        "@P1 bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t"
        "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// global -> shared, completion counted on the mbarrier (bytes: multiple of 16; both addresses 16-byte aligned)
__device__ __forceinline__ void bulk_g2s(void* smem, const void* gmem, unsigned bytes, uint64_t* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem)), "l"(gmem), "r"(bytes),
                 "r"(smem_u32(bar))
                 : "memory");
}
// shared -> global, tracked by the thread's bulk group
__device__ __forceinline__ void bulk_s2g(void* gmem, const void* smem, unsigned bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem), "r"(smem_u32(smem)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// the committed groups have READ their shared-memory source (the CTA may exit / reuse the tile)
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
// writes made to shared memory by threads become visible to the async proxy (call before bulk_s2g, then sync)
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

}  // namespace pst_tma
