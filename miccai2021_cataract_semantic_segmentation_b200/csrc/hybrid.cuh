// Hybrid segmented sort, global half: ONE most-significant-digit partition pass over the candidates.
//
// Replaces (together with hyb_local_kernel in lovasz.cu) the C sequential full-length torch.sort calls plus the
// lovasz_grad scans of losses/LovaszSoftmax.py:57,83-95 (reference).  The three-pass LSD sort of sort.cuh moves every
// (key, value) pair through L2 three times and needs a separate pass for the Jaccard scan.  Here the pairs cross
// global memory once:
//   hyb_count      per-tile histogram of the top w <= 13 key bits in shared memory, added to the segment's histogram
//                  hist[seg][bin] with one global atomic per non-empty (tile, bin); the CTA that finishes a segment's
//                  last tile turns its histogram into exclusive bucket offsets (cursors)
//   hyb_partition  per tile: arrival ranks from shared-memory atomics, one atomicAdd on the bucket cursor per
//                  non-empty (tile, bin) reserves the tile's slice of the bucket, elements are grouped by bucket in
//                  shared memory and written with run-contiguous stores
// The partition is NOT stable (slices of a bucket land in the order the tiles reserve them): the local kernel sorts
// every bucket by its remaining low bits in shared memory and restores the canonical tie order (ascending pixel index
// among equal keys = torch.sort(stable=True)) from the pixel index in the value.
// Digit width per segment: w = clamp(ceil(log2 n) - 6, 0, min(13, key bits)): ~64..128 elements per bucket if the keys
// were spread evenly, a few thousand at the peaks of real distributions; a bucket that does not fit the local kernel's
// shared memory sends its whole segment through the LSD path instead (sort_fallback_kernel), so any input is handled.
#pragma once
#include "sort.cuh"

// ---- count: per-bucket element and foreground counts -------------------------------------------------------------------------
// A CTA walks a contiguous range of tiles and keeps ONE shared histogram per run of tiles of the same segment (packed:
// elements | foreground flags << 16 per tile, widened when added up), so the cost of clearing and flushing the bins is
// paid per (CTA, segment), not per tile.  The CTA that adds a segment's last tile scans the segment: hist -> exclusive
// element offsets (the partition's cursors), fgpre -> foreground flags in front of every bucket (the local kernel reads
// its foreground prefix there: no scan over the sorted order is ever needed).
__device__ __forceinline__ void hyb_count_flush(const SortArgs& a, const HybArgs& h, u32* s_hist, u32* s_warp, u32* s_last,
                                                int seg, u32 ntl) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const u32 ns = a.seg_count[seg];
    const u32 nbins = 1u << hyb_plan(a.seg_bits[seg], ns).w;
    u32* gh = h.hist + (size_t)seg * HYB_MAX_BINS;
    u32* gf = h.fgpre + (size_t)seg * HYB_MAX_BINS;
    __syncthreads();
    for (u32 b = tid; b < nbins; b += SORT_TPB) {
        const u32 c = s_hist[b], f = s_hist[HYB_MAX_BINS + b];
        if (c) atomicAdd(gh + b, c);
        if (f) atomicAdd(gf + b, f);
    }
    __threadfence();
    __syncthreads();
    const u32 ntiles_seg = (ns + SORT_TILE - 1) / SORT_TILE;
    if (tid == 0) *s_last = (atomicAdd(h.seg_done + seg, ntl) + ntl == ntiles_seg);
    __syncthreads();
    if (!*s_last) return;
    __threadfence();
    // exclusive scans over the bins: warp w owns a contiguous range, walked in rows of 32 (coalesced)
    const u32 per_warp = (nbins + SORT_WARPS - 1) / SORT_WARPS;
    const u32 wb0 = warp * per_warp, wb1 = min(wb0 + per_warp, nbins);
    u32 sc = 0, sf = 0;
    for (u32 b = wb0 + lane; b < wb1; b += 32) { sc += __ldcg(gh + b); sf += __ldcg(gf + b); }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { sc += __shfl_xor_sync(FULL_MASK, sc, o); sf += __shfl_xor_sync(FULL_MASK, sf, o); }
    if (lane == 0) { s_warp[warp] = sc; s_warp[SORT_WARPS + warp] = sf; }
    __syncthreads();
    u32 rc = 0, rf = 0;
    for (int w2 = 0; w2 < warp; ++w2) { rc += s_warp[w2]; rf += s_warp[SORT_WARPS + w2]; }
    for (u32 b0 = wb0; b0 < wb1; b0 += 32) {
        const u32 b = b0 + lane;
        const u32 c = b < wb1 ? __ldcg(gh + b) : 0u, f = b < wb1 ? __ldcg(gf + b) : 0u;
        u32 ic = c, jf = f;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const u32 x = __shfl_up_sync(FULL_MASK, ic, o), y = __shfl_up_sync(FULL_MASK, jf, o);
            if (lane >= o) { ic += x; jf += y; }
        }
        if (b < wb1) { __stcg(gh + b, rc + ic - c); __stcg(gf + b, rf + jf - f); }
        rc += __shfl_sync(FULL_MASK, ic, 31);
        rf += __shfl_sync(FULL_MASK, jf, 31);
    }
    __syncthreads();
}

__global__ void __launch_bounds__(SORT_TPB) hyb_count_kernel(SortArgs a, HybArgs h, u32 total_bound) {
    pdl_enter();
    extern __shared__ __align__(16) u32 s_hist[];          // [2][HYB_MAX_BINS]: elements, foreground flags
    __shared__ u32 s_warp[2 * SORT_WARPS];
    __shared__ u32 s_last;
    const int tid = threadIdx.x;
    const u32 total_tiles = min(a.tile_start[a.n_seg], total_bound);
    const u32 t0 = (u32)((u64)total_tiles * blockIdx.x / gridDim.x), t1 = (u32)((u64)total_tiles * (blockIdx.x + 1) / gridDim.x);
    int cur_seg = -1;
    u32 ntl = 0, L = 0;
    // software pipeline: the keys of tile t + 1 (and the descriptor of tile t + 2) are requested before the shared-memory
    // atomics of tile t, so a CTA's tiles do not each pay the descriptor -> keys round trips one after the other
    auto load_keys = [&](const uint4& d, u32 (&k)[SORT_KPT]) {
        const u32* __restrict__ kp = a.keys[0] + (size_t)d.x * a.cap + d.y;
#pragma unroll
        for (int j = 0; j < SORT_KPT; ++j) {
            const u32 idx = j * SORT_TPB + tid;
            k[j] = idx < d.z ? kp[idx] : 0u;
        }
    };
    const uint4 none = make_uint4(0, 0, 0, 0);
    uint4 d4 = t0 < t1 ? a.tile_desc[t0] : none;
    uint4 dn = t0 + 1 < t1 ? a.tile_desc[t0 + 1] : none;
    u32 key[SORT_KPT];
    if (t0 < t1) load_keys(d4, key);
    for (u32 t = t0; t < t1; ++t) {
        const uint4 dnn = t + 2 < t1 ? a.tile_desc[t + 2] : none;
        u32 nkey[SORT_KPT];
#pragma unroll
        for (int j = 0; j < SORT_KPT; ++j) nkey[j] = 0u;
        if (t + 1 < t1) load_keys(dn, nkey);
        const int seg = (int)d4.x;
        const u32 n = d4.z;
        if (seg != cur_seg) {
            if (cur_seg >= 0) hyb_count_flush(a, h, s_hist, s_warp, &s_last, cur_seg, ntl);
            const HybPlan pl = hyb_plan(a.seg_bits[seg], a.seg_count[seg]);
            L = pl.L;
            const u32 nbins = 1u << pl.w;
            __syncthreads();
            for (u32 b = tid; b < nbins; b += SORT_TPB) { s_hist[b] = 0; s_hist[HYB_MAX_BINS + b] = 0; }
            cur_seg = seg;
            ntl = 0;
        }
        __syncthreads();                                   // bins cleared
#pragma unroll
        for (int k = 0; k < SORT_KPT; ++k) {
            if (k * SORT_TPB + tid < n) {                  // (the foreground flag rides in bit 31 of the key: the values are not read)
                const u32 d = (key[k] & KEY_MASK) >> L;
                atomicAdd(&s_hist[d], 1u);
                if (key[k] & KEY_FG) atomicAdd(&s_hist[HYB_MAX_BINS + d], 1u);
            }
        }
        ++ntl;
        d4 = dn; dn = dnn;
#pragma unroll
        for (int j = 0; j < SORT_KPT; ++j) key[j] = nkey[j];
    }
    if (cur_seg >= 0) hyb_count_flush(a, h, s_hist, s_warp, &s_last, cur_seg, ntl);
}

// ---- partition ---------------------------------------------------------------------------------------------------------------------
struct PartSmem {
    u32 cnt[HYB_MAX_BINS / 2];    // u16 pairs: arrivals per bucket -> local start of the bucket in the staged tile
    u32 gd[SORT_TILE];            // indexed by the local start of a run: global slice start - local start
    u32 keys[SORT_TILE];
    u32 vals[SORT_TILE];          // (the list of non-empty buckets aliases vals until the tile is staged)
    u32 warp_sum[SORT_WARPS];
    u32 nlist;
};

__global__ void __launch_bounds__(SORT_TPB, 3) hyb_partition_kernel(SortArgs a, HybArgs h, u32 total_bound) {
    pdl_enter();
    extern __shared__ __align__(16) unsigned char smem_raw[];
    PartSmem& S = *reinterpret_cast<PartSmem*>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const u32 total_tiles = min(a.tile_start[a.n_seg], total_bound);
    u32* __restrict__ kout = a.keys[1];
    u32* __restrict__ vout = a.vals[1];
    for (u32 t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        const uint4 d4 = a.tile_desc[t];
        const int seg = (int)d4.x;
        const u32 off = d4.y, n = d4.z;
        const HybPlan pl = hyb_plan(a.seg_bits[seg], a.seg_count[seg]);
        const u32 nbins = 1u << pl.w, L = pl.L;
        const u32 nwords = (nbins + 1) / 2;
        u32* __restrict__ ko = kout + (size_t)seg * a.cap;
        u32* __restrict__ vo = vout + (size_t)seg * a.cap;
        __syncthreads();                                   // previous tile fully written out
        for (u32 b = tid; b < nwords; b += SORT_TPB) S.cnt[b] = 0;
        if (tid == 0) S.nlist = 0;
        __syncthreads();
        u32 key[SORT_KPT], val[SORT_KPT];
        unsigned short rnk[SORT_KPT];
        const u32 wbase = warp * (32 * SORT_KPT) + lane;
        {
            const u32* __restrict__ kp = a.keys[0] + (size_t)seg * a.cap + off;
            const u32* __restrict__ vp = a.vals[0] + (size_t)seg * a.cap + off;
#pragma unroll
            for (int k = 0; k < SORT_KPT; ++k) {
                const u32 idx = wbase + k * 32;
                key[k] = idx < n ? kp[idx] : 0u;
                val[k] = idx < n ? vp[idx] : 0u;
            }
        }
#pragma unroll
        for (int k = 0; k < SORT_KPT; ++k) {               // arrival rank within the bucket (u16 halves: a tile has <= 4096 elements)
            const u32 d = (key[k] & KEY_MASK) >> L, sh = (d & 1u) * 16u;
            rnk[k] = (wbase + k * 32 < n) ? (unsigned short)((atomicAdd(&S.cnt[d >> 1], 1u << sh) >> sh) & 0xFFFFu) : (unsigned short)0;
        }
        __syncthreads();                                   // all arrivals counted
        {   // local starts (exclusive scan over the buckets; warp w owns a contiguous range of words, rows of 32) and the
            // list of non-empty buckets
            const u32 per_warp = (nwords + SORT_WARPS - 1) / SORT_WARPS;
            const u32 w0 = warp * per_warp, w1 = min(w0 + per_warp, nwords);
            u32 s = 0;
            for (u32 i = w0 + lane; i < w1; i += 32) { const u32 x = S.cnt[i]; s += (x & 0xFFFFu) + (x >> 16); }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(FULL_MASK, s, o);
            if (lane == 0) S.warp_sum[warp] = s;
            __syncthreads();
            u32 run = 0;
            for (int w2 = 0; w2 < warp; ++w2) run += S.warp_sum[w2];
            for (u32 i0 = w0; i0 < w1; i0 += 32) {
                const u32 i = i0 + lane;
                const u32 x = i < w1 ? S.cnt[i] : 0u;
                const u32 c0 = x & 0xFFFFu, c1 = x >> 16, c = c0 + c1;
                u32 inc = c;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { const u32 y = __shfl_up_sync(FULL_MASK, inc, o); if (lane >= o) inc += y; }
                const u32 e0 = run + inc - c, e1 = e0 + c0;
                if (i < w1) S.cnt[i] = e0 | (e1 << 16);
                const u32 nz = (c0 ? 1u : 0u) + (c1 ? 1u : 0u);
                if (nz) {                                  // list entry: bucket | count << 16 (its local start is in S.cnt)
                    u32 at = atomicAdd(&S.nlist, nz);
                    if (c0) S.vals[at++] = (2 * i) | (c0 << 16);
                    if (c1) S.vals[at] = (2 * i + 1) | (c1 << 16);
                }
                run += __shfl_sync(FULL_MASK, inc, 31);
            }
        }
        __syncthreads();
        {   // the tile's slice of every non-empty bucket: one global atomic each, spread evenly over the threads
            u32* gh = h.hist + (size_t)seg * HYB_MAX_BINS;
            const u32 nl = S.nlist;
            for (u32 e0 = tid; e0 < nl; e0 += 4 * SORT_TPB) {
                u32 ent[4], g[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) ent[j] = e0 + j * SORT_TPB < nl ? S.vals[e0 + j * SORT_TPB] : 0u;
#pragma unroll
                for (int j = 0; j < 4; ++j) g[j] = ent[j] ? atomicAdd(gh + (ent[j] & 0xFFFFu), ent[j] >> 16) : 0u;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    if (ent[j]) {
                        const u32 d = ent[j] & 0xFFFFu;
                        const u32 ls = (S.cnt[d >> 1] >> ((d & 1u) * 16u)) & 0xFFFFu;
                        S.gd[ls] = g[j] - ls;
                    }
                }
            }
        }
        __syncthreads();                                   // (the list in S.vals is consumed)
#pragma unroll
        for (int k = 0; k < SORT_KPT; ++k) {
            if (wbase + k * 32 < n) {
                const u32 d = (key[k] & KEY_MASK) >> L;
                const u32 pos = ((S.cnt[d >> 1] >> ((d & 1u) * 16u)) & 0xFFFFu) + rnk[k];
                S.keys[pos] = key[k];
                S.vals[pos] = val[k];
            }
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < SORT_KPT; ++k) {
            const u32 i = k * SORT_TPB + tid;
            if (i < n) {
                const u32 kk = S.keys[i];
                const u32 d = (kk & KEY_MASK) >> L;
                const u32 ls = (S.cnt[d >> 1] >> ((d & 1u) * 16u)) & 0xFFFFu;
                const u32 pos = S.gd[ls] + i;              // slice start - local start + local index
                ko[pos] = kk;
                vo[pos] = S.vals[i];
            }
        }
    }
}
