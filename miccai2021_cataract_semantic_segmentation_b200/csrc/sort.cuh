// Segmented, stable LSD radix sort on (u32 key, u32 value) pairs held in fixed-capacity segment regions.
//
// Replaces the C sequential full-length torch.sort calls of losses/LovaszSoftmax.py:57 (reference): one
// segment per class (flat) or per (image, class) (per-image); only candidate elements are present.
//
// Layout in HBM: segment s owns [s*cap, s*cap + seg_count[s]) of each array.  Keys carry at most
// seg_bits[s] <= 30 significant bits, so exactly three digit passes of w = ceil(bits/3) <= 10 bits are run
// (A -> B -> A -> B; the result always lands in buffer 1).  Each pass is a single "onesweep"-style kernel:
// tiles of 4096 elements take a ticket, rank their keys stably (warp match + per-warp counters), publish
// per-bin tile counts and resolve the bins' global offsets by decoupled look-back within the segment.
#pragma once
#include "common.cuh"

#define SORT_TPB 256
#define SORT_KPT 16
#define SORT_WARPS (SORT_TPB / 32)
#define SORT_TILE (SORT_TPB * SORT_KPT)      // 4096 elements
#define SORT_MAX_BINS 1024
#define SORT_PASSES 3

struct SortArgs {
    u32* keys[2];
    u32* vals[2];
    const u32* seg_count;   // [n_seg] elements per segment
    const u32* seg_bits;    // [n_seg] significant key bits (1..30)
    int n_seg;
    long long cap;          // segment region stride (elements)
    u32* tile_start;        // [n_seg + 1] exclusive prefix of tiles per segment (written by sort_plan_kernel)
    u32* ghist;             // [n_seg][3][1024] digit histograms -> exclusive bin bases
    u32* lb[2];             // [max_tiles][1024] look-back state, ping-pong between passes
    u64* lb_chain;          // [max_tiles] spare single-chain state (zeroed here, used by the Jaccard kernel)
    u32* tickets;           // [4]
    int* status;
};

struct SortScratch {
    size_t tile_start, ghist, lb0, lb1, lb_chain, tickets, total;
    u32 max_tiles;
};

static inline SortScratch sort_scratch_layout(int n_seg, long long total_capacity) {
    SortScratch L;
    L.max_tiles = (u32)(total_capacity / SORT_TILE + n_seg + 1);
    size_t o = 0;
    L.tickets = o;    o = align_up(o + 64, 256);
    L.tile_start = o; o = align_up(o + sizeof(u32) * (size_t)(n_seg + 1), 256);
    L.ghist = o;      o = align_up(o + sizeof(u32) * (size_t)n_seg * SORT_PASSES * SORT_MAX_BINS, 256);
    L.lb0 = o;        o = align_up(o + sizeof(u32) * (size_t)L.max_tiles * SORT_MAX_BINS, 256);
    L.lb1 = o;        o = align_up(o + sizeof(u32) * (size_t)L.max_tiles * SORT_MAX_BINS, 256);
    L.lb_chain = o;   o = align_up(o + sizeof(u64) * (size_t)L.max_tiles, 256);
    L.total = o;
    return L;
}

__device__ __forceinline__ u32 sort_digit_width(u32 bits) {
    u32 w = (bits + SORT_PASSES - 1) / SORT_PASSES;
    return w < 1 ? 1 : (w > 10 ? 10 : w);
}

// largest segment whose first tile is <= t (skips empty segments)
__device__ __forceinline__ int sort_find_segment(const u32* tile_start, int n_seg, u32 t) {
    int lo = 0, hi = n_seg - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (__ldg(tile_start + mid) <= t) lo = mid; else hi = mid - 1;
    }
    return lo;
}

// ---- plan: tiles per segment -> exclusive prefix ------------------------------------------------------------
__global__ void __launch_bounds__(1024) sort_plan_kernel(SortArgs a) {
    __shared__ u32 s_warp[32];
    __shared__ u32 s_carry;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_carry = 0;
    __syncthreads();
    for (int base = 0; base < a.n_seg; base += 1024) {
        const int s = base + tid;
        const u32 nt = s < a.n_seg ? (a.seg_count[s] + SORT_TILE - 1) / SORT_TILE : 0;
        u32 v = nt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const u32 x = __shfl_up_sync(FULL_MASK, v, o); if (lane >= o) v += x; }
        if (lane == 31) s_warp[warp] = v;
        __syncthreads();
        if (warp == 0) {
            u32 wv = s_warp[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const u32 x = __shfl_up_sync(FULL_MASK, wv, o); if (lane >= o) wv += x; }
            s_warp[lane] = wv;
        }
        __syncthreads();
        const u32 incl = v + (warp ? s_warp[warp - 1] : 0) + s_carry;
        if (s < a.n_seg) a.tile_start[s] = incl - nt;
        __syncthreads();
        if (tid == 1023) s_carry = incl;
        __syncthreads();
    }
    if (tid == 0) a.tile_start[a.n_seg] = s_carry;
}

// ---- upfront digit histograms of all three passes (one read of the keys) -----------------------------------
__global__ void __launch_bounds__(SORT_TPB) sort_hist_kernel(SortArgs a) {
    __shared__ u32 s_hist[SORT_PASSES][SORT_MAX_BINS];
    const int tid = threadIdx.x, lane = tid & 31;
    const u32 total_tiles = a.tile_start[a.n_seg];
    const u32 t0 = (u32)(((u64)total_tiles * blockIdx.x) / gridDim.x);
    const u32 t1 = (u32)(((u64)total_tiles * (blockIdx.x + 1)) / gridDim.x);
    if (t0 >= t1) return;
    for (int i = tid; i < SORT_PASSES * SORT_MAX_BINS; i += SORT_TPB) (&s_hist[0][0])[i] = 0;
    __syncthreads();
    int seg = sort_find_segment(a.tile_start, a.n_seg, t0);
    const u32* kin = a.keys[0];
    for (u32 t = t0; t < t1; ++t) {
        if (t >= a.tile_start[seg + 1]) {                 // segment change: flush
            __syncthreads();
            const u32 nbins = 1u << sort_digit_width(a.seg_bits[seg]);
            for (int i = tid; i < SORT_PASSES * SORT_MAX_BINS; i += SORT_TPB) {
                const u32 v = (&s_hist[0][0])[i];
                if ((u32)(i & (SORT_MAX_BINS - 1)) < nbins && v) atomicAdd(a.ghist + (size_t)seg * SORT_PASSES * SORT_MAX_BINS + i, v);
                (&s_hist[0][0])[i] = 0;
            }
            __syncthreads();
            seg = sort_find_segment(a.tile_start, a.n_seg, t);
        }
        const u32 off = (t - a.tile_start[seg]) * SORT_TILE;
        const u32 count = a.seg_count[seg];
        const u32 n = min((u32)SORT_TILE, count - off);
        const u32 w = sort_digit_width(a.seg_bits[seg]);
        const u32 dmask = (1u << w) - 1;
        const size_t base = (size_t)seg * a.cap + off;
#pragma unroll 4
        for (int k = 0; k < SORT_KPT; ++k) {
            const u32 idx = k * SORT_TPB + tid;
            const bool valid = idx < n;
            const u32 key = valid ? kin[base + idx] : 0;
#pragma unroll
            for (int p = 0; p < SORT_PASSES; ++p) {
                const u32 d = valid ? ((key >> (p * w)) & dmask) : 0xFFFFFFFFu;
                const u32 m = __match_any_sync(FULL_MASK, d);
                if (valid && lane == __ffs(m) - 1) atomicAdd(&s_hist[p][d], (u32)__popc(m));
            }
        }
        // clear the look-back state the first pass (and the Jaccard chain) will use for this tile
        uint4* row = (uint4*)(a.lb[0] + (size_t)t * SORT_MAX_BINS);
        row[tid] = make_uint4(0, 0, 0, 0);
        if (tid == 0) a.lb_chain[t] = 0;
    }
    __syncthreads();
    const u32 nbins = 1u << sort_digit_width(a.seg_bits[seg]);
    for (int i = tid; i < SORT_PASSES * SORT_MAX_BINS; i += SORT_TPB) {
        const u32 v = (&s_hist[0][0])[i];
        if ((u32)(i & (SORT_MAX_BINS - 1)) < nbins && v) atomicAdd(a.ghist + (size_t)seg * SORT_PASSES * SORT_MAX_BINS + i, v);
    }
}

// ---- exclusive scan of each (segment, pass) histogram -> bin bases ------------------------------------------
__global__ void __launch_bounds__(SORT_MAX_BINS) sort_scan_kernel(SortArgs a) {
    __shared__ u32 s_warp[32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    u32* h = a.ghist + ((size_t)blockIdx.x * SORT_PASSES + blockIdx.y) * SORT_MAX_BINS;
    if (a.seg_count[blockIdx.x] == 0) return;
    const u32 x = h[tid];
    u32 v = x;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const u32 y = __shfl_up_sync(FULL_MASK, v, o); if (lane >= o) v += y; }
    if (lane == 31) s_warp[warp] = v;
    __syncthreads();
    if (warp == 0) {
        u32 wv = s_warp[lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const u32 y = __shfl_up_sync(FULL_MASK, wv, o); if (lane >= o) wv += y; }
        s_warp[lane] = wv;
    }
    __syncthreads();
    h[tid] = v - x + (warp ? s_warp[warp - 1] : 0);
}

// ---- one digit pass -----------------------------------------------------------------------------------------
__global__ void __launch_bounds__(SORT_TPB) sort_pass_kernel(SortArgs a, int pass) {
    __shared__ u32 s_cnt[SORT_WARPS][SORT_MAX_BINS + 1];
    __shared__ u32 s_binoff[SORT_MAX_BINS];
    __shared__ u32 s_ticket;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const u32 lt_mask = (1u << lane) - 1;
    const u32 total_tiles = a.tile_start[a.n_seg];
    // (static indexing only: a dynamically indexed kernel-parameter array would be copied to local memory)
    const bool odd = pass & 1;
    const u32* __restrict__ kin = odd ? a.keys[1] : a.keys[0];
    const u32* __restrict__ vin = odd ? a.vals[1] : a.vals[0];
    u32* __restrict__ kout = odd ? a.keys[0] : a.keys[1];
    u32* __restrict__ vout = odd ? a.vals[0] : a.vals[1];
    u32* lbr = odd ? a.lb[1] : a.lb[0];
    u32* lbn = odd ? a.lb[0] : a.lb[1];

    for (;;) {
        __syncthreads();
        if (tid == 0) s_ticket = atomicAdd(a.tickets + 1 + pass, 1u);
        __syncthreads();
        const u32 t = s_ticket;
        if (t >= total_tiles) break;
        const int seg = sort_find_segment(a.tile_start, a.n_seg, t);
        const u32 tis = t - a.tile_start[seg];                 // tile index within its segment
        const u32 off = tis * SORT_TILE;
        const u32 n = min((u32)SORT_TILE, a.seg_count[seg] - off);
        const u32 w = sort_digit_width(a.seg_bits[seg]);
        const u32 nbins = 1u << w;
        const u32 dmask = nbins - 1;
        const u32 shift = pass * w;
        const size_t base = (size_t)seg * a.cap;

        for (u32 i = tid; i < SORT_WARPS * (SORT_MAX_BINS + 1); i += SORT_TPB) {
            const u32 b = i % (SORT_MAX_BINS + 1);
            if (b <= nbins) (&s_cnt[0][0])[i] = 0;
        }
        __syncthreads();

        u32 key[SORT_KPT];
        unsigned short rnk[SORT_KPT];
        const u32 wbase = warp * (32 * SORT_KPT) + lane;
#pragma unroll
        for (int k = 0; k < SORT_KPT; ++k) {
            const u32 idx = wbase + k * 32;
            key[k] = idx < n ? kin[base + off + idx] : 0xFFFFFFFFu;
        }
#pragma unroll
        for (int k = 0; k < SORT_KPT; ++k) {
            const u32 idx = wbase + k * 32;
            const u32 d = idx < n ? ((key[k] >> shift) & dmask) : nbins;     // padding lanes share the dummy bin
            const u32 m = __match_any_sync(FULL_MASK, d);
            const int leader = __ffs(m) - 1;
            u32 old = 0;
            if (lane == leader) { old = s_cnt[warp][d]; s_cnt[warp][d] = old + __popc(m); }
            old = __shfl_sync(FULL_MASK, old, leader);
            rnk[k] = (unsigned short)(old + __popc(m & lt_mask));
            __syncwarp();
        }
        __syncthreads();

        for (u32 b = tid; b < nbins; b += SORT_TPB) {
            u32 run = 0;
#pragma unroll
            for (int w2 = 0; w2 < SORT_WARPS; ++w2) { const u32 c = s_cnt[w2][b]; s_cnt[w2][b] = run; run += c; }
            u32* mine = lbr + (size_t)t * SORT_MAX_BINS + b;
            u32 excl = 0;
            if (tis == 0) {
                st_relaxed(mine, lb_pack32(LB_INCL, run));
            } else {
                st_relaxed(mine, lb_pack32(LB_AGG, run));
                long long look = (long long)t - 1;
                const long long first = (long long)t - tis;
                while (look >= first) {
                    const u32 s = lb_wait32(lbr + (size_t)look * SORT_MAX_BINS + b, a.status);
                    excl += lb_val32(s);
                    if (lb_flag32(s) != LB_AGG) break;
                    --look;
                }
                st_relaxed(mine, lb_pack32(LB_INCL, excl + run));
            }
            s_binoff[b] = a.ghist[((size_t)seg * SORT_PASSES + pass) * SORT_MAX_BINS + b] + excl;
        }
        __syncthreads();

#pragma unroll
        for (int k = 0; k < SORT_KPT; ++k) {
            const u32 idx = wbase + k * 32;
            if (idx < n) {
                const u32 d = (key[k] >> shift) & dmask;
                const u32 pos = s_binoff[d] + s_cnt[warp][d] + rnk[k];
                kout[base + pos] = key[k];
                vout[base + pos] = vin[base + off + idx];
            }
        }
        if (pass + 1 < SORT_PASSES) {
            uint4* row = (uint4*)(lbn + (size_t)t * SORT_MAX_BINS);
            row[tid] = make_uint4(0, 0, 0, 0);
        }
    }
}

// Enqueue plan + histogram + scan + three passes.  keys/vals[0] = input, result in keys/vals[1].
static inline int sort_enqueue(const SortArgs& a, const SortScratch& L, void* scratch_base, cudaStream_t st) {
    // tickets + histograms start from zero
    CUDA_TRY(cudaMemsetAsync((char*)scratch_base + L.tickets, 0, 64, st));
    CUDA_TRY(cudaMemsetAsync((char*)scratch_base + L.ghist, 0,
                             sizeof(u32) * (size_t)a.n_seg * SORT_PASSES * SORT_MAX_BINS, st));
    sort_plan_kernel<<<1, 1024, 0, st>>>(a);
    LAUNCH_CHECK("sort_plan_kernel");
    const int sms = b200seg_sm_count();
    sort_hist_kernel<<<sms * 2, SORT_TPB, 0, st>>>(a);
    LAUNCH_CHECK("sort_hist_kernel");
    sort_scan_kernel<<<dim3(a.n_seg, SORT_PASSES), SORT_MAX_BINS, 0, st>>>(a);
    LAUNCH_CHECK("sort_scan_kernel");
    b200seg_stage(4, st);
    for (int p = 0; p < SORT_PASSES; ++p) {
        sort_pass_kernel<<<sms * 4, SORT_TPB, 0, st>>>(a, p);
        LAUNCH_CHECK("sort_pass_kernel");
    }
    b200seg_stage(5, st);
    return 0;
}
