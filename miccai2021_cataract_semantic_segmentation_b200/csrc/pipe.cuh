// Bulk-async (TMA engine) tile pipeline for the streaming passes over the logits.
//
// The logits are NCHW: the C values of a pixel are C planes apart, so a tile of TILE consecutive pixels of one image
// is C contiguous rows of TILE*4 bytes.  One elected thread per CTA issues one `cp.async.bulk` per row (plus the
// tile's label / per-pixel-state rows) into a ring of shared-memory stages and arms an mbarrier with the byte count;
// all warps wait on the barrier's phase, compute from shared memory, and a __syncthreads() hands the stage back.
// Loads for the next STAGES-1 tiles are always in flight, independent of how many registers the math needs.
#pragma once
#include "common.cuh"

__device__ __forceinline__ u32 smem_u32(const void* p) { return (u32)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(u64* bar, u32 count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(u64* bar, u32 bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// generic-proxy reads of a stage must be ordered before the async-proxy writes that refill it
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, u32 bytes, u64* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// Wait for the phase with the given parity; bounded (a pipeline bug must not hang the device).
__device__ __forceinline__ void mbar_wait(u64* bar, u32 parity) {
    const u32 addr = smem_u32(bar);
    u32 done = 0;
    for (u32 tries = 0; tries < (1u << 26); ++tries) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
        if (done) return;
    }
    __trap();
}
