// cp.async (LDGSTS) tile pipelines for the streaming passes over the logits.
//
// The logits are NCHW: the C values of a pixel are C planes apart, so a tile of consecutive pixels of one image is C
// contiguous rows.  (A cp.async.bulk / mbarrier variant with one request per row was measured and dropped: ~290 cycles
// per request per SM, serialised; see DESIGN.md.)
#pragma once
#include "common.cuh"

__device__ __forceinline__ u32 smem_u32(const void* p) { return (u32)__cvta_generic_to_shared(p); }

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

