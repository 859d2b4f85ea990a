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

// ---- per-thread cp.async (LDGSTS) helpers -------------------------------------------------------------------------
// Each thread copies the bytes it will later read itself, so completion is tracked per thread with commit/wait
// groups and no CTA-wide barrier is needed.
template <int BYTES>
__device__ __forceinline__ void cp_async(void* smem_dst, const void* gsrc) {
    static_assert(BYTES == 4 || BYTES == 8 || BYTES == 16, "cp.async moves 4, 8 or 16 bytes");
    if constexpr (BYTES == 16)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
    else
        asm volatile("cp.async.ca.shared.global [%0], [%1], %2;" ::"r"(smem_u32(smem_dst)), "l"(gsrc), "n"(BYTES) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// ---- warp-private cp.async tile pipeline -------------------------------------------------------------------------------
// A warp tile is WT = 32*VEC consecutive pixels of one image: C rows of WT*4 bytes.  The warp copies the rows with
// 16-byte cp.async (LDGSTS.128: one instruction moves 512 bytes) into its private ring of STAGES buffers, STAGES-1
// tiles ahead of the math.  HBM reads never wait for the math, no registers are tied up by data in flight, there is no
// CTA barrier (cp.async.wait_group + __syncwarp), and dynamic class indices address shared memory.
template <int CT, int VEC>
struct WarpTile {
    static constexpr int WT = 32 * VEC;                 // pixels per warp tile
    static constexpr int CHUNKS = WT / 4;               // 16-byte chunks per row
    static constexpr int ROWS_PER_INSTR = 32 / CHUNKS;  // rows covered by one warp-wide cp.async
    static constexpr int NINSTR = (CT + ROWS_PER_INSTR - 1) / ROWS_PER_INSTR;
    // copy the tile whose first pixel is (image n, pixel q0) into buf[CT][WT]
    static __device__ __forceinline__ void prefetch(float (*buf)[WT], const float* logits, int n, long long q0,
                                                    long long HW, int lane) {
        const int ch = lane % CHUNKS, r0 = lane / CHUNKS;
        const long long q = q0 + 4 * ch;
        if (q < HW) {
            // one running source pointer and one running shared-memory address: two adds per copy
            const float* src = logits + ((size_t)n * CT + r0) * HW + q;
            const size_t step = (size_t)ROWS_PER_INSTR * HW;
            u32 dst = smem_u32(&buf[r0][4 * ch]);
#pragma unroll
            for (int i = 0; i < NINSTR; ++i) {
                if ((i + 1) * ROWS_PER_INSTR <= CT || i * ROWS_PER_INSTR + r0 < CT)
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
                src += step;
                dst += ROWS_PER_INSTR * WT * 4;
            }
        }
    }
};

