// Segmented, stable LSD radix sort on (u32 key, u32 value) pairs held in fixed-capacity segment regions.
//
// Replaces the C sequential full-length torch.sort calls of losses/LovaszSoftmax.py:57 (reference): one
// segment per class (flat) or per (image, class) (per-image); only candidate elements are present.
//
// Layout in HBM: segment s owns [s*cap, s*cap + seg_count[s]) of each ping-pong array.  Keys carry at most
// seg_bits[s] <= 30 significant bits, so exactly three digit passes of w = ceil(bits/3) <= 10 bits are run
// (src -> B -> A -> B; the result always lands in buffer 1).  Segments have few tiles (tens), all running at the
// same time, so a chained scan would serialise; each pass is two short kernels instead:
//   count    per-tile digit histogram (shared-memory atomics) -> tilehist[tile][bin]; the CTA that finishes a segment's
//            last tile scans it: column scan over its tiles + exclusive scan over bins -> tile offsets, bin_base
//            (segments above SORT_SCAN_LOCAL_MAX tiles: scanned by the whole scatter grid instead, sort_big_scan)
//   scatter  re-read the tile (L2), stable ranks (ballot / MATCH.ANY peer masks + per-warp counters), reorder in shared
//            memory, write to the other buffer with warp-contiguous stores; launched cooperatively (grid barriers)
// The emission kernels leave every segment compact but in no particular order (tiles reserve their slices with atomics), so
// for the loss the three key passes are preceded by SORT_VAL_PASSES 8-bit passes over the value bits: equal keys then end up
// in ascending pixel order = torch.sort(stable=True).  sort_prepare_kernel plans the tiles and writes their descriptors.
// A last small kernel counts the foreground flags (value bit 0) per tile of the final order for the Jaccard scan and
// drops the dead source buffer from L2.
#pragma once
#include "common.cuh"
#include <cstdlib>

#define SORT_TPB 256
#define SORT_KPT 16
#define SORT_WARPS (SORT_TPB / 32)
#define SORT_TILE (SORT_TPB * SORT_KPT)      // 4096 elements
#define SORT_MAX_BINS 1024
#define SORT_PASSES 3                        // digit passes over the key bits
#define SORT_VAL_PASSES 4                    // optional leading 8-bit passes over the value bits (unordered input)
#define SORT_MAX_PASSES (SORT_PASSES + SORT_VAL_PASSES)
#define SORT_SCAN_LOCAL_MAX 128               // segments with more tiles are scanned by the whole grid (see sort_big_scan)
#define SORT_CHUNK 64                        // tiles per chunk of that scan
#ifndef SORT_SCATTER_MINB
#define SORT_SCATTER_MINB 3                  // resident scatter CTAs per SM the register budget is cut for
#endif

struct SortArgs {
    u32* keys[2];
    u32* vals[2];
    const u32* seg_count;   // [n_seg] elements per segment
    const u32* seg_bits;    // [n_seg] significant key bits (1..30)
    int n_seg;
    long long cap;          // segment stride of the ping-pong arrays (elements)
    // scratch
    u32* tile_start;        // [n_seg + 1] exclusive prefix of tiles per segment
    uint4* tile_desc;       // [max_tiles] {segment, first element, element count, digit width}
    u32* seg_done;          // [SORT_PASSES][n_seg] tiles counted so far (last CTA of a segment runs its scan)
    u32* tilehist;          // [max_tiles][1024]
    u32* bin_base;          // [n_seg][1024]
    u32* tile_fg;           // [max_tiles] foreground flags per tile of the final order (sort_fg_count_kernel)
    u32* big;               // [1 + n_seg] number of segments with > SORT_SCAN_LOCAL_MAX tiles, then their ids
    u32* chunksum;          // [max_chunks][1024] per-chunk digit sums -> exclusive chunk offsets (big segments only)
    u32* gbar;              // [SORT_GBAR] grid-barrier counters: 2 per pass (sort_big_scan), the rest for the fused fallback
    int* status;
    const u32* seg_sel;     // nullptr: every segment; else only segments with seg_sel[seg] != 0 are processed
    int val_passes;         // 0: stable sort by key (input order kept among equal keys); SORT_VAL_PASSES: sort by (key, value)
};
#define SORT_GBAR 40

struct SortScratch {
    size_t tile_start, tile_desc, seg_done, tilehist, bin_base, tile_fg, big, chunksum, gbar, total;
    u32 max_tiles;
};

static inline SortScratch sort_scratch_layout(int n_seg, long long total_capacity) {
    SortScratch L;
    L.max_tiles = (u32)(total_capacity / SORT_TILE + n_seg + 1);
    size_t o = 0;
    L.tile_start = o; o = align_up(o + sizeof(u32) * (size_t)(n_seg + 1), 256);
    L.tile_desc = o;  o = align_up(o + sizeof(uint4) * (size_t)L.max_tiles, 256);
    L.seg_done = o;   o = align_up(o + sizeof(u32) * (size_t)n_seg * SORT_MAX_PASSES, 256);
    L.bin_base = o;   o = align_up(o + sizeof(u32) * (size_t)n_seg * SORT_MAX_BINS, 256);
    L.tile_fg = o;    o = align_up(o + sizeof(u32) * (size_t)L.max_tiles, 256);
    L.big = o;        o = align_up(o + sizeof(u32) * (size_t)(n_seg + 1), 256);
    L.gbar = o;       o = align_up(o + sizeof(u32) * SORT_GBAR, 256);
    L.chunksum = o;   o = align_up(o + sizeof(u32) * ((size_t)L.max_tiles / SORT_CHUNK + n_seg + 2) * SORT_MAX_BINS, 256);
    L.tilehist = o;   o = align_up(o + sizeof(u32) * (size_t)L.max_tiles * SORT_MAX_BINS, 256);
    L.total = o;
    return L;
}

// ---- hybrid path (hybrid.cuh, hyb_local_kernel): shared definitions -------------------------------------------------
// local tiles: the local kernel's work unit starts every LOC_T0 elements of a segment and owns the buckets that START in
// its window; LOC_CAP bounds what it can hold in shared memory (a bucket longer than LOC_CAP - LOC_T0 may overflow it)
#define LOC_TPB 512
#define LOC_WARPS (LOC_TPB / 32)
#define LOC_KPT 16
#define LOC_CAP (LOC_TPB * LOC_KPT)
#ifndef LOC_T0
#define LOC_T0 4096
#endif
#define LOC_PER_SORT_TILE (SORT_TILE / LOC_T0)
#define HYB_MAX_W 13
#define HYB_MAX_BINS (1u << HYB_MAX_W)

struct HybArgs {
    u32* hist;          // [n_seg][HYB_MAX_BINS] bucket counts -> exclusive bucket offsets -> running cursors
    u32* fgpre;         // [n_seg][HYB_MAX_BINS] foreground flags per bucket -> foreground flags in front of the bucket
    u32* seg_done;      // [n_seg] tiles of the segment counted so far
    u32* ticket;        // [2] work ticket of the local kernel, CTAs of it that have finished
    u32* seg_ovf;       // [n_seg] 1 = a bucket of the segment did not fit: the segment goes through the LSD fallback
    u32* ovf_any;       // [1]
    double* seg_loss;   // [n_seg] loss sums of the local kernel (the fallback sums into LovaszParams::seg_loss)
};

struct HybPlan { u32 w, L; };   // partition digit = key >> L, w bits wide
__host__ __device__ __forceinline__ HybPlan hyb_plan(u32 bits, u32 n) {
#ifdef __CUDA_ARCH__
    const u32 lg = n <= 1 ? 0u : 32u - (u32)__clz((int)(n - 1));   // ceil(log2 n)
#else
    u32 lg = 0;
    while (lg < 31 && (1u << lg) < n) ++lg;
#endif
    int w = (int)lg - 6;
    if (w < 0) w = 0;
    if (w > HYB_MAX_W) w = HYB_MAX_W;
    if (w > (int)bits) w = (int)bits;
    HybPlan p;
    p.w = (u32)w;
    p.L = bits - (u32)w;
    return p;
}


__device__ __forceinline__ u32 sort_digit_width(u32 bits) {
    u32 w = (bits + SORT_PASSES - 1) / SORT_PASSES;
    return w < 1 ? 1 : (w > 10 ? 10 : w);
}

// digit of pass `pass`: the leading a.val_passes passes take 8-bit digits of the value (so that equal keys end up in
// ascending value = pixel order whatever order the input came in), the remaining SORT_PASSES passes w_key-bit digits of the key
struct PassDigit { u32 shift, w; bool by_val; };
__device__ __forceinline__ PassDigit sort_pass_digit(const SortArgs& a, int pass, u32 w_key) {
    PassDigit d;
    d.by_val = pass < a.val_passes;
    d.w = d.by_val ? 8u : w_key;
    d.shift = d.by_val ? 8u * (u32)pass : (u32)(pass - a.val_passes) * w_key;
    return d;
}

// largest segment whose first tile is <= t (skips empty segments)
__device__ __forceinline__ int sort_find_segment(const u32* tile_start, int n_seg, u32 t) {
    int lo = 0, hi = n_seg - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (tile_start[mid] <= t) lo = mid; else hi = mid - 1;
    }
    return lo;
}

// ---- prepare: everything between the emission and the first digit pass ----------------------------------------------
//   every CTA: clears the hybrid path's bucket tables of its segments (hist == nullptr: plain path, nothing to clear)
//   CTA 0:     tiles per segment -> exclusive prefix (tile_start); clears the per-pass counters; the list of big segments;
//              one descriptor per tile {segment, first element, element count, key digit width}, so the per-tile prologue
//              of every later kernel is a single 16-byte load instead of a chain of dependent ones
#define SORT_PREP_TPB 1024
__global__ void __launch_bounds__(SORT_PREP_TPB) sort_prepare_kernel(SortArgs a, HybArgs h, u32 max_tiles) {
    pdl_enter();
    __shared__ u32 s_warp[32];
    __shared__ u32 s_carry;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (h.hist) {
        for (int seg = blockIdx.x; seg < a.n_seg; seg += gridDim.x) {
            const u32 nb = 1u << hyb_plan(a.seg_bits[seg], a.seg_count[seg]).w;
            u32* gh = h.hist + (size_t)seg * HYB_MAX_BINS;
            u32* gf = h.fgpre + (size_t)seg * HYB_MAX_BINS;
            for (u32 b = tid; b < nb; b += SORT_PREP_TPB) { gh[b] = 0; gf[b] = 0; }
            if (tid == 0) h.seg_done[seg] = 0;
        }
    }
    if (blockIdx.x != 0) return;
    if (tid == 0) s_carry = 0;
    __syncthreads();
    for (int base = 0; base < a.n_seg; base += SORT_PREP_TPB) {
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
        if (tid == SORT_PREP_TPB - 1) s_carry = incl;
        __syncthreads();
    }
    const u32 total = s_carry;
    if (tid == 0) { a.tile_start[a.n_seg] = total; a.big[0] = 0; }
    if (tid < SORT_GBAR) a.gbar[tid] = 0;
    __syncthreads();
    for (int sg = tid; sg < a.n_seg; sg += SORT_PREP_TPB)
        if ((a.seg_count[sg] + SORT_TILE - 1) / SORT_TILE > SORT_SCAN_LOCAL_MAX) a.big[1 + atomicAdd(a.big, 1u)] = (u32)sg;
    for (int i = tid; i < a.n_seg * SORT_MAX_PASSES; i += SORT_PREP_TPB) a.seg_done[i] = 0;
    for (u32 i = tid; i < total && i < max_tiles; i += SORT_PREP_TPB) a.tile_fg[i] = 0;
    if (h.hist && tid < 2) h.ticket[tid] = 0;
    __syncthreads();
    for (int seg = warp; seg < a.n_seg; seg += SORT_PREP_TPB / 32) {      // descriptors: a warp per segment
        const u32 t_first = a.tile_start[seg], nt = a.tile_start[seg + 1] - t_first;
        const u32 cnt = a.seg_count[seg], w = sort_digit_width(a.seg_bits[seg]);
        for (u32 i = lane; i < nt && t_first + i < max_tiles; i += 32) {
            const u32 off = i * SORT_TILE;
            a.tile_desc[t_first + i] = make_uint4((u32)seg, off, min((u32)SORT_TILE, cnt - off), w);
        }
    }
}

// ---- per-segment scan, run by the CTA that counted the segment's last tile ---------------------------------------------
// column scan over the segment's tiles (tilehist[tile][bin] -> exclusive offset of the tile within the bin) and
// exclusive scan over the bin totals (-> bin_base).  256 threads, thread b owns bins [4b, 4b+4).
__device__ __forceinline__ void segment_scan(const SortArgs& a, u32* rows, int seg, u32 t0, u32 t1, u32 nbins, u32* s_warp) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    u32 tot[4] = {0, 0, 0, 0};
    if ((u32)(4 * tid) < nbins) {
        uint4* col = reinterpret_cast<uint4*>(rows + (size_t)t0 * SORT_MAX_BINS + 4 * tid);
        constexpr int STRIDE = SORT_MAX_BINS / 4;
        u32 t = t0;
        for (; t + 8 <= t1; t += 8, col += 8 * STRIDE) {                 // 8 independent 16-byte loads, then the sums
            uint4 x[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) x[i] = __ldcg(col + i * STRIDE);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                col[i * STRIDE] = make_uint4(tot[0], tot[1], tot[2], tot[3]);
                tot[0] += x[i].x; tot[1] += x[i].y; tot[2] += x[i].z; tot[3] += x[i].w;
            }
        }
        for (; t < t1; ++t, col += STRIDE) {
            const uint4 x = __ldcg(col);
            *col = make_uint4(tot[0], tot[1], tot[2], tot[3]);
            tot[0] += x.x; tot[1] += x.y; tot[2] += x.z; tot[3] += x.w;
        }
    }
    const u32 tsum = tot[0] + tot[1] + tot[2] + tot[3];
    u32 v = tsum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const u32 y = __shfl_up_sync(FULL_MASK, v, o); if (lane >= o) v += y; }
    if (lane == 31) s_warp[warp] = v;
    __syncthreads();
    u32 excl = v - tsum;
#pragma unroll
    for (int w2 = 0; w2 < SORT_WARPS; ++w2) if (w2 < warp) excl += s_warp[w2];
    if ((u32)(4 * tid) < nbins) {
        u32* bb = a.bin_base + (size_t)seg * SORT_MAX_BINS + 4 * tid;
        *reinterpret_cast<uint4*>(bb) = make_uint4(excl, excl + tot[0], excl + tot[0] + tot[1], excl + tot[0] + tot[1] + tot[2]);
    }
    __syncthreads();
}

// ---- count: per-tile digit histogram (+ the segment's scan once its last tile is in) ---------------------------------
__device__ __forceinline__ void sort_count_body(const SortArgs& a, int pass, u32 total_bound) {
    __shared__ __align__(16) u32 s_hist[SORT_MAX_BINS];
    __shared__ u32 s_warp[SORT_WARPS];
    __shared__ u32 s_last;
    const int tid = threadIdx.x;
    const u32 total_tiles = min(a.tile_start[a.n_seg], total_bound);
    for (u32 t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        const uint4 d4 = a.tile_desc[t];
        const int seg = (int)d4.x;
        const u32 off = d4.y, n = d4.z;
        const PassDigit pd = sort_pass_digit(a, pass, d4.w);
        const u32 nbins = 1u << pd.w, dmask = nbins - 1, shift = pd.shift;
        if (a.seg_sel && !a.seg_sel[seg]) continue;        // (CTA-uniform)
        __syncthreads();                                   // previous iteration done with s_hist
        for (u32 b = tid; b < SORT_MAX_BINS; b += SORT_TPB) s_hist[b] = 0;
        __syncthreads();
        const bool odd = pass & 1;                         // (static indexing of the kernel-parameter arrays only)
        const u32* __restrict__ src = (pd.by_val ? (odd ? a.vals[1] : a.vals[0]) : (odd ? a.keys[1] : a.keys[0])) + (size_t)seg * a.cap + off;
        u32 key[SORT_KPT];
#pragma unroll
        for (int k = 0; k < SORT_KPT; ++k) {
            const u32 idx = k * SORT_TPB + tid;
            key[k] = idx < n ? src[idx] : 0;
        }
#pragma unroll
        for (int k = 0; k < SORT_KPT; ++k)
            if (k * SORT_TPB + tid < n) atomicAdd(&s_hist[(key[k] >> shift) & dmask], 1u);
        __syncthreads();
        uint4* row = reinterpret_cast<uint4*>(a.tilehist + (size_t)t * SORT_MAX_BINS);
        if ((u32)(4 * tid) < nbins) __stcg(row + tid, reinterpret_cast<const uint4*>(s_hist)[tid]);
        // last CTA to finish a tile of this segment scans the segment
        __threadfence();
        __syncthreads();
        const u32 tseg0 = t - off / SORT_TILE;
        const u32 ntiles_seg = (a.seg_count[seg] + SORT_TILE - 1) / SORT_TILE;
        if (tid == 0) s_last = (atomicAdd(a.seg_done + (size_t)pass * a.n_seg + seg, 1u) == ntiles_seg - 1);
        __syncthreads();
        if (s_last && ntiles_seg <= SORT_SCAN_LOCAL_MAX) {   // (larger segments: sort_big_scan in the scatter kernel)
            __threadfence();
            segment_scan(a, a.tilehist, seg, tseg0, tseg0 + ntiles_seg, nbins, s_warp);
        }
    }
}
__global__ void __launch_bounds__(SORT_TPB) sort_count_kernel(SortArgs a, int pass, u32 total_bound) {
    sort_count_body(a, pass, total_bound);
}

// ---- scatter ----------------------------------------------------------------------------------------------------------------
// Elements are first placed at their tile-local sorted position in shared memory, then streamed out so that the
// lanes of a warp write consecutive addresses within each digit bin (a fully scattered 4-byte store costs the LSU
// one wavefront per lane).
// lanes holding the same digit (what match.any returns, but built from ballots: MATCH.ANY costs ~250 cycles per warp
// instruction on this part, a ballot a few).  A warp whose 32 digits agree (the usual case in the top-digit pass,
// where keys cluster) skips the loop.
template <int NBITS>
__device__ __forceinline__ u32 peer_mask(u32 d) {
    u32 peers = FULL_MASK;
#pragma unroll
    for (int b = 0; b < NBITS; ++b) {
        const u32 x = 0u - ((d >> b) & 1u);               // all ones if bit b of d is set
        const u32 m = __ballot_sync(FULL_MASK, x != 0u);
        peers &= ~(m ^ x);                                  // keep lanes whose bit b equals mine
    }
    return peers;
}

#define SORT_CNT_STRIDE (SORT_MAX_BINS + 4)                // u16 row stride, keeps 4-bin groups 8-byte aligned
struct ScatterSmem {
    unsigned short cnt[SORT_WARPS][SORT_CNT_STRIDE];     // per-warp digit counters -> exclusive offsets across warps
    unsigned short binexcl[SORT_MAX_BINS];               // exclusive prefix of the tile's bin totals
    u32 binoff[SORT_MAX_BINS];                           // global position of local sorted index i in bin d: binoff[d] + i
    u32 keys[SORT_TILE];
    u32 vals[SORT_TILE];
    u32 warp_sum[SORT_WARPS];
};

// stable rank of each of the warp's SORT_KPT rows within (warp, digit); NBITS = digit width + 1 (the extra bit is
// the dummy bin of padding lanes).  Two loops on purpose: the peer masks of all rows are independent (their votes
// pipeline), only the counter updates form a chain through shared memory.
template <int NBITS, bool USE_MATCH>
__device__ __forceinline__ void rank_rows(ScatterSmem& S, const u32 (&key)[SORT_KPT], u32 (&rnk)[SORT_KPT], u32 n,
                                          u32 wbase, u32 shift, u32 dmask, u32 nbins, int warp, int lane) {
    const u32 lt_mask = (1u << lane) - 1;
#pragma unroll
    for (int k = 0; k < SORT_KPT; ++k) {
        const u32 d = (wbase + k * 32 < n) ? ((key[k] >> shift) & dmask) : nbins;
        rnk[k] = USE_MATCH ? __match_any_sync(FULL_MASK, d) : peer_mask<NBITS>(d);
    }
#pragma unroll
    for (int k = 0; k < SORT_KPT; ++k) {
        const u32 d = (wbase + k * 32 < n) ? ((key[k] >> shift) & dmask) : nbins;
        const u32 m = rnk[k];
        const int leader = __ffs(m) - 1;
        u32 old = 0;
        if (lane == leader) { old = S.cnt[warp][d]; S.cnt[warp][d] = (unsigned short)(old + __popc(m)); }
        rnk[k] = __shfl_sync(FULL_MASK, old, leader) + __popc(m & lt_mask);
        __syncwarp();
    }
}

// ---- scatter ----------------------------------------------------------------------------------------------------------------
// Elements are first placed at their tile-local sorted position in shared memory, then streamed out so that the
// lanes of a warp write consecutive addresses within each digit bin (a fully scattered 4-byte store costs the LSU
// one wavefront per lane).
// ---- big segments: the histogram scan by the whole grid ------------------------------------------------------------------
// A segment of T tiles needs, per digit, the exclusive prefix of its T tile histograms.  The count kernel's last-CTA scan
// walks them 8 per L2 round trip: fine for tens of tiles, 200-600 us for the 500-2500 tiles of one class of confident
// logits.  For segments above SORT_SCAN_LOCAL_MAX tiles the scatter kernel does it instead, with every CTA:
//   phase 1  sum of each chunk of SORT_CHUNK tile histograms              (one CTA per chunk, round-robin)
//   phase 2  exclusive scan of a segment's chunk sums + its bin bases     (one CTA per big segment)
// and each scatter tile then adds the <= 63 raw histograms of its chunk that precede it.  Two grid barriers; all CTAs of
// the scatter grid are co-resident (the launch clamps the grid to the occupancy).
__device__ __forceinline__ void grid_barrier(u32* ctr, int* status) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(ctr, 1u);
        u32 spins = 0;
        while (ld_relaxed(ctr) < gridDim.x) {
            __nanosleep(64);
            if (++spins >= SPIN_LIMIT) {                     // never compute on incomplete data: flag it and kill the stream
                atomicOr(status, STATUS_SPIN_TIMEOUT);
                __threadfence_system();
                __trap();
            }
        }
        __threadfence();
    }
    __syncthreads();
}
__device__ __forceinline__ u32 sort_chunk_id(u32 tseg0, int seg, u32 lc) { return (tseg0 / SORT_CHUNK) + (u32)seg + lc; }

__device__ __forceinline__ void sort_big_scan(const SortArgs& a, int pass, u32* s_warp) {
    const int tid = threadIdx.x;
    const u32 n_big = a.big[0];
    u32 unit = 0;
    for (u32 j = 0; j < n_big; ++j) {                      // phase 1
        const int seg = (int)a.big[1 + j];
        if (a.seg_sel && !a.seg_sel[seg]) continue;
        const u32 t0 = a.tile_start[seg], nt = a.tile_start[seg + 1] - t0;
        const u32 nbins = 1u << sort_pass_digit(a, pass, sort_digit_width(a.seg_bits[seg])).w, nch = (nt + SORT_CHUNK - 1) / SORT_CHUNK;
        for (u32 lc = 0; lc < nch; ++lc, ++unit) {
            if (unit % gridDim.x != blockIdx.x || (u32)(4 * tid) >= nbins) continue;
            const u32 ta = t0 + lc * SORT_CHUNK, tb = min(ta + SORT_CHUNK, t0 + nt);
            const uint4* col = reinterpret_cast<const uint4*>(a.tilehist + (size_t)ta * SORT_MAX_BINS + 4 * tid);
            constexpr int STRIDE = SORT_MAX_BINS / 4;
            u32 tot[4] = {0, 0, 0, 0};
            u32 t = ta;
            for (; t + 16 <= tb; t += 16, col += 16 * STRIDE) {
                uint4 x[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) x[i] = __ldcg(col + i * STRIDE);
#pragma unroll
                for (int i = 0; i < 16; ++i) { tot[0] += x[i].x; tot[1] += x[i].y; tot[2] += x[i].z; tot[3] += x[i].w; }
            }
            for (; t < tb; ++t, col += STRIDE) { const uint4 x = __ldcg(col); tot[0] += x.x; tot[1] += x.y; tot[2] += x.z; tot[3] += x.w; }
            __stcg(reinterpret_cast<uint4*>(a.chunksum + (size_t)sort_chunk_id(t0, seg, lc) * SORT_MAX_BINS + 4 * tid),
                   make_uint4(tot[0], tot[1], tot[2], tot[3]));
        }
    }
    grid_barrier(a.gbar + 2 * pass, a.status);
    for (u32 j = blockIdx.x; j < n_big; j += gridDim.x) {  // phase 2
        const int seg = (int)a.big[1 + j];
        if (a.seg_sel && !a.seg_sel[seg]) continue;
        const u32 t0 = a.tile_start[seg], nt = a.tile_start[seg + 1] - t0;
        const u32 nbins = 1u << sort_pass_digit(a, pass, sort_digit_width(a.seg_bits[seg])).w, nch = (nt + SORT_CHUNK - 1) / SORT_CHUNK;
        const u32 c0 = sort_chunk_id(t0, seg, 0);
        __syncthreads();
        segment_scan(a, a.chunksum, seg, c0, c0 + nch, nbins, s_warp);
    }
    grid_barrier(a.gbar + 2 * pass + 1, a.status);
}

// GATHER = true: keys and values are both loaded up front (needed by the passes over the value digits).
// GATHER = false: the values are loaded only when they are placed, so key + rank are all that lives in
// registers through the ranking and the kernel fits 4 CTAs per SM (1100 tiles: 2 rounds of 592 instead of 3 of 444).
template <bool USE_MATCH, bool GATHER>
__device__ __forceinline__ void sort_scatter_body(const SortArgs& a, int pass, u32 total_bound) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    ScatterSmem& S = *reinterpret_cast<ScatterSmem*>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const u32 total_tiles = min(a.tile_start[a.n_seg], total_bound);
    // (static indexing only: a dynamically indexed kernel-parameter array would be copied to local memory)
    const bool odd = pass & 1;
    u32* __restrict__ kout = odd ? a.keys[0] : a.keys[1];
    u32* __restrict__ vout = odd ? a.vals[0] : a.vals[1];
    if (a.big[0]) sort_big_scan(a, pass, S.warp_sum);     // grid-uniform

    for (u32 t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        const uint4 d4 = a.tile_desc[t];
        const int seg = (int)d4.x;
        const u32 off = d4.y, n = d4.z;
        const PassDigit pd = sort_pass_digit(a, pass, d4.w);       // (by_val passes are launched with GATHER = true: both arrays in registers)
        const u32 w = pd.w, nbins = 1u << w, dmask = nbins - 1, shift = pd.shift;
        u32* __restrict__ ko = kout + (size_t)seg * a.cap;
        u32* __restrict__ vo = vout + (size_t)seg * a.cap;
        if (a.seg_sel && !a.seg_sel[seg]) continue;        // (CTA-uniform)
        __syncthreads();
        {
            uint2* z = reinterpret_cast<uint2*>(&S.cnt[0][0]);
            for (u32 i = tid; i < sizeof(S.cnt) / 8; i += SORT_TPB) z[i] = make_uint2(0, 0);
        }
        __syncthreads();
        const u32* __restrict__ kp = (odd ? a.keys[1] : a.keys[0]) + (size_t)seg * a.cap + off;
        const u32* __restrict__ vp = (odd ? a.vals[1] : a.vals[0]) + (size_t)seg * a.cap + off;

        u32 key[SORT_KPT], val[GATHER ? SORT_KPT : 1], rnk[SORT_KPT];
        const u32 wbase = warp * (32 * SORT_KPT) + lane;
#pragma unroll
        for (int k = 0; k < SORT_KPT; ++k) {
            const bool valid = wbase + k * 32 < n;
            key[k] = valid ? kp[wbase + k * 32] : 0xFFFFFFFFu;
            if (GATHER) val[GATHER ? k : 0] = valid ? vp[wbase + k * 32] : 0u;
        }
        if constexpr (GATHER) {
            if (pd.by_val) rank_rows<9, USE_MATCH>(S, val, rnk, n, wbase, shift, dmask, nbins, warp, lane);   // digit of the value
        }
        if (!(GATHER && pd.by_val)) {
            switch (w) {                                  // uniform per tile
                case 10: rank_rows<11, USE_MATCH>(S, key, rnk, n, wbase, shift, dmask, nbins, warp, lane); break;
                case 9: rank_rows<10, USE_MATCH>(S, key, rnk, n, wbase, shift, dmask, nbins, warp, lane); break;
                case 8: rank_rows<9, USE_MATCH>(S, key, rnk, n, wbase, shift, dmask, nbins, warp, lane); break;
                default: rank_rows<8, USE_MATCH>(S, key, rnk, n, wbase, shift, dmask, nbins, warp, lane); break;
            }
        }
        __syncthreads();
        // thread b owns bins [4b, 4b+4): four u16 counters travel as one 64-bit word (no carries: totals <= 4096)
        u64 run = 0;
        if ((u32)(4 * tid) < nbins + 1) {
#pragma unroll
            for (int w2 = 0; w2 < SORT_WARPS; ++w2) {
                u64* c4 = reinterpret_cast<u64*>(&S.cnt[w2][4 * tid]);
                const u64 c = *c4;
                *c4 = run;
                run += c;
            }
        }
        u32 tot[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) tot[j] = (4 * tid + j < (int)nbins) ? (u32)((run >> (16 * j)) & 0xFFFFu) : 0u;
        const u32 tsum = tot[0] + tot[1] + tot[2] + tot[3];
        u32 v = tsum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const u32 y = __shfl_up_sync(FULL_MASK, v, o); if (lane >= o) v += y; }
        if (lane == 31) S.warp_sum[warp] = v;
        __syncthreads();
        u32 excl = v - tsum;
#pragma unroll
        for (int w2 = 0; w2 < SORT_WARPS; ++w2) if (w2 < warp) excl += S.warp_sum[w2];
        if ((u32)(4 * tid) < nbins) {
            const uint4 bb = *reinterpret_cast<const uint4*>(a.bin_base + (size_t)seg * SORT_MAX_BINS + 4 * tid);
            uint4 th;
            const u32 tseg0 = t - off / SORT_TILE, nts = a.tile_start[seg + 1] - tseg0;
            if (nts <= SORT_SCAN_LOCAL_MAX) {             // scanned by the count kernel: already the tile's offset in the bin
                th = *reinterpret_cast<const uint4*>(a.tilehist + (size_t)t * SORT_MAX_BINS + 4 * tid);
            } else {                                      // chunk offset + the raw histograms of the chunk's earlier tiles
                const u32 lc = (t - tseg0) / SORT_CHUNK;
                th = __ldcg(reinterpret_cast<const uint4*>(a.chunksum + (size_t)sort_chunk_id(tseg0, seg, lc) * SORT_MAX_BINS + 4 * tid));
                const uint4* col = reinterpret_cast<const uint4*>(a.tilehist + (size_t)(tseg0 + lc * SORT_CHUNK) * SORT_MAX_BINS + 4 * tid);
                constexpr int STRIDE = SORT_MAX_BINS / 4;
                u32 tq = tseg0 + lc * SORT_CHUNK;
                for (; tq + 8 <= t; tq += 8, col += 8 * STRIDE) {
                    uint4 x[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) x[i] = __ldcg(col + i * STRIDE);
#pragma unroll
                    for (int i = 0; i < 8; ++i) { th.x += x[i].x; th.y += x[i].y; th.z += x[i].z; th.w += x[i].w; }
                }
                for (; tq < t; ++tq, col += STRIDE) { const uint4 x = __ldcg(col); th.x += x.x; th.y += x.y; th.z += x.z; th.w += x.w; }
            }
            const u32 e0 = excl, e1 = e0 + tot[0], e2 = e1 + tot[1], e3 = e2 + tot[2];
            *reinterpret_cast<uint2*>(&S.binexcl[4 * tid]) = make_uint2(e0 | (e1 << 16), e2 | (e3 << 16));
            *reinterpret_cast<uint4*>(&S.binoff[4 * tid]) =
                make_uint4(bb.x + th.x - e0, bb.y + th.y - e1, bb.z + th.z - e2, bb.w + th.w - e3);
        }
        __syncthreads();
        if (GATHER) {
#pragma unroll
            for (int k = 0; k < SORT_KPT; ++k) {
                if (wbase + k * 32 < n) {
                    const u32 d = ((pd.by_val ? val[GATHER ? k : 0] : key[k]) >> shift) & dmask;
                    const u32 lpos = (u32)S.binexcl[d] + S.cnt[warp][d] + rnk[k];
                    S.keys[lpos] = key[k];
                    S.vals[lpos] = val[GATHER ? k : 0];
                }
            }
        } else {                                           // values straight from the source tile (L2) to their place
#pragma unroll
            for (int k = 0; k < SORT_KPT; ++k) {
                const bool valid = wbase + k * 32 < n;
                const u32 v = valid ? vp[wbase + k * 32] : 0u;
                const u32 d = (key[k] >> shift) & dmask;
                rnk[k] = valid ? (u32)S.binexcl[d] + S.cnt[warp][d] + rnk[k] : 0xFFFFFFFFu;
                if (valid) S.keys[rnk[k]] = key[k];
                key[k] = v;                                // the key register now carries the value
            }
#pragma unroll
            for (int k = 0; k < SORT_KPT; ++k)
                if (rnk[k] != 0xFFFFFFFFu) S.vals[rnk[k]] = key[k];
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < SORT_KPT; ++k) {
            const u32 i = k * SORT_TPB + tid;
            if (i < n) {
                const u32 kk = S.keys[i], vv = S.vals[i];
                const u32 pos = S.binoff[(((GATHER && pd.by_val) ? vv : kk) >> shift) & dmask] + i;
                ko[pos] = kk;
                vo[pos] = vv;
            }
        }
    }
}
template <bool USE_MATCH, bool GATHER>
__global__ void __launch_bounds__(SORT_TPB, GATHER ? SORT_SCATTER_MINB : SORT_SCATTER_MINB + 1)
sort_scatter_kernel(SortArgs a, int pass, u32 total_bound) {
    sort_scatter_body<USE_MATCH, GATHER>(a, pass, total_bound);
}

// foreground flags (value bit 0) per tile of the final order, for the Jaccard scan
__device__ __forceinline__ void sort_fg_count_body(const SortArgs& a, u32 total_bound) {
    __shared__ u32 s_w[SORT_WARPS];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const u32 total_tiles = min(a.tile_start[a.n_seg], total_bound);
    for (u32 t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        const uint4 d4 = a.tile_desc[t];
        if (a.seg_sel && !a.seg_sel[d4.x]) continue;       // (CTA-uniform)
        const u32* v = a.vals[1] + (size_t)d4.x * a.cap + d4.y;
        {   // buffer 0 (source of the last pass) is dead: drop this tile's share of its dirty lines from L2 unwritten
            const size_t sb = (size_t)d4.x * a.cap;
            discard_dead_lines(a.keys[0] + sb + d4.y, a.keys[0] + sb + d4.y + SORT_TILE, a.keys[0] + sb, a.keys[0] + sb + a.cap);
            discard_dead_lines(a.vals[0] + sb + d4.y, a.vals[0] + sb + d4.y + SORT_TILE, a.vals[0] + sb, a.vals[0] + sb + a.cap);
        }
        u32 c = 0;
#pragma unroll
        for (int k = 0; k < SORT_KPT; ++k) {
            const u32 i = k * SORT_TPB + tid;
            c += (i < d4.z) ? (v[i] & 1u) : 0u;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(FULL_MASK, c, o);
        __syncthreads();
        if (lane == 0) s_w[warp] = c;
        __syncthreads();
        if (tid == 0) {
            u32 tot = 0;
#pragma unroll
            for (int w2 = 0; w2 < SORT_WARPS; ++w2) tot += s_w[w2];
            a.tile_fg[t] = tot;
        }
    }
}
__global__ void __launch_bounds__(SORT_TPB) sort_fg_count_kernel(SortArgs a, u32 total_bound) {
    sort_fg_count_body(a, total_bound);
}

// Enqueue plan + descriptors + three (count+scan, scatter) passes.  Result in keys/vals[1].
static inline int sort_enqueue(const SortArgs& a, const SortScratch& L, cudaStream_t st) {
    HybArgs h = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};   // (hist == nullptr: plain path)
    static bool attr_set[64] = {false};
    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64 || !attr_set[dev]) {          // opt in to > 48 KB of dynamic shared memory, once per device
        const void* fns[4] = {(const void*)sort_scatter_kernel<false, false>, (const void*)sort_scatter_kernel<false, true>,
                              (const void*)sort_scatter_kernel<true, false>, (const void*)sort_scatter_kernel<true, true>};
        for (const void* f : fns)
            CUDA_TRY(cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ScatterSmem)));
        if (dev >= 0 && dev < 64) attr_set[dev] = true;
    }
    {
        const int sms_p = b200seg_sm_count();
        const int pgrid = 1;
        sort_prepare_kernel<<<pgrid, SORT_PREP_TPB, 0, st>>>(a, h, L.max_tiles);
    }
    LAUNCH_CHECK("sort_prepare_kernel");
    b200seg_stage(4, st);
    // peer masks by MATCH.ANY cost ~ the number of distinct digits in the warp: a win only for the top digit, where
    // the keys cluster (measured: 38 vs 45 us for that pass, 65 vs 56 us for the low digits); B200SEG_SORT_MATCH = 0 / 1
    // forces ballots / match everywhere
    const int match_mode = b200seg_tuning().sort_match;
    const int sms = b200seg_sm_count();
    const u32 cgrid = L.max_tiles < (u32)sms * 8 ? L.max_tiles : (u32)sms * 8;
    // the scatter grids must be co-resident (grid barriers of sort_big_scan): clamp them to the measured occupancy
    static int occ_gather[64] = {0}, occ_compact[64] = {0};
    if (dev >= 0 && dev < 64 && occ_gather[dev] == 0) {
        int o[4] = {0, 0, 0, 0};
        CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o[0], sort_scatter_kernel<false, true>, SORT_TPB, sizeof(ScatterSmem)));
        CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o[1], sort_scatter_kernel<true, true>, SORT_TPB, sizeof(ScatterSmem)));
        CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o[2], sort_scatter_kernel<false, false>, SORT_TPB, sizeof(ScatterSmem)));
        CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o[3], sort_scatter_kernel<true, false>, SORT_TPB, sizeof(ScatterSmem)));
        occ_gather[dev] = max(1, min(min(o[0], o[1]), SORT_SCATTER_MINB));
        occ_compact[dev] = max(1, min(min(o[2], o[3]), SORT_SCATTER_MINB + 1));
    }
    const int n_passes = a.val_passes + SORT_PASSES;
    for (int p = 0; p < n_passes; ++p) {
        sort_count_kernel<<<cgrid, SORT_TPB, 0, st>>>(a, p, L.max_tiles);
        LAUNCH_CHECK("sort_count_kernel");
        const bool use_match = p >= a.val_passes && (match_mode == 1 || (match_mode == 2 && p == n_passes - 1));
        const bool gather = p < a.val_passes;              // value-digit passes keep both arrays in registers
        {   // cooperative launch: the grid is resident as a whole or not at all, so the grid barriers of sort_big_scan cannot
            // dead-lock against another partially resident grid (two loss heads on two streams)
            const u32 per_sm = (dev >= 0 && dev < 64) ? (u32)(gather ? occ_gather[dev] : occ_compact[dev]) : 1u;
            const u32 sgrid = L.max_tiles < (u32)sms * per_sm ? L.max_tiles : (u32)sms * per_sm;
            SortArgs a_copy = a;
            int pass_copy = p;
            u32 bound_copy = L.max_tiles;
            void* args[] = {(void*)&a_copy, (void*)&pass_copy, (void*)&bound_copy};
            const void* fn = gather ? (use_match ? (const void*)sort_scatter_kernel<true, true> : (const void*)sort_scatter_kernel<false, true>)
                                    : (use_match ? (const void*)sort_scatter_kernel<true, false> : (const void*)sort_scatter_kernel<false, false>);
            CUDA_TRY(cudaLaunchCooperativeKernel(fn, dim3(sgrid), dim3(SORT_TPB), args, sizeof(ScatterSmem), st));
        }
        if (p >= a.val_passes && p + 1 < n_passes) b200seg_stage(5 + p - a.val_passes, st);
    }
    sort_fg_count_kernel<<<cgrid, SORT_TPB, 0, st>>>(a, L.max_tiles);
    LAUNCH_CHECK("sort_fg_count_kernel");
    b200seg_stage(7, st);
    return 0;
}
