// Online hard example mining cross entropy for sm_100a.  Replaces losses/OhemCrossEntropy.py:22-40 of the reference
// (softmax + per-pixel cross entropy + gather + full-length sort of the label probabilities + threshold + masked mean)
// and its autograd backward.
//
// The sort only serves to read one order statistic (the min_kept-th smallest label probability), so it becomes a
// three-level radix select over the fp32 bit patterns (probabilities are non-negative: their bits order like the values):
//   ohem_stats      one pass over the logits (4*C + label bytes read, 16 bytes written per pixel): softmax max / sum,
//                   p_label, -log p_label, and the top-10-bit histogram of p_label
//   ohem_select_reduce   one cooperative kernel: every CTA walks the 1024-bin histogram to the bin holding the wanted
//                   rank; unless that already settles the threshold, two refinements inside the bin (grid-wide
//                   histograms over the L2-resident 16 MB p_label plane, a grid barrier each); then sum / count of the
//                   losses with p_label < max(order statistic, thresh), the last CTA writes the mean
//   ohem_backward   one pass: dlogits = go / n_kept * (softmax - onehot) on kept pixels, 0 elsewhere (pixels that are
//                   not kept never read their logits)
#include "b200seg.h"
#include "common.cuh"

#define OH_TPB 256
#define OH_BINS 1024
enum { OH_NVALID = 0, OH_THR, OH_TICKET, OH_KEPT, OH_INVKEPT, OH_BAR0, OH_BAR1, OH_CTRL_WORDS = 16 };
#define OH_INVALID_BITS 0x7F800000u     // +inf marks ignored / out-of-range pixels in the p_label plane

struct OhemParams {
    const float* logits;
    const void* labels;
    int N, C;
    long long HW, P;
    int has_ignore, ignore;
    float thresh;
    long long min_kept;
    float *p_lab, *ce, *pix_m, *pix_s;
    u32* hist;          // [3][OH_BINS]
    u32* ctrl;          // OH_CTRL_WORDS
    double* ce_sum;
    float* loss_out;
    int* status;
};

struct OhemLayout { size_t p_lab, ce, pix_m, pix_s, hist, ctrl, ce_sum, total; };
static OhemLayout ohem_layout(long long P) {
    OhemLayout L;
    size_t o = 0;
    const size_t plane = align_up((size_t)P * sizeof(float), 256);
    L.p_lab = o; o += plane;
    L.ce = o; o += plane;
    L.pix_m = o; o += plane;
    L.pix_s = o; o += plane;
    L.hist = o; o += 3 * OH_BINS * sizeof(u32);
    L.ctrl = o; o += OH_CTRL_WORDS * sizeof(u32);
    L.ce_sum = o; o += 256;
    L.total = o;
    return L;
}

__device__ __forceinline__ bool ohem_valid(const OhemParams& p, int lab, u32& oob) {
    if (p.has_ignore && lab == p.ignore) return false;
    if ((unsigned)lab >= (unsigned)p.C) { oob = 1; return false; }
    return true;
}
__device__ __forceinline__ u32 ohem_bin0(u32 bits) { const u32 b = bits >> 20; return b < OH_BINS ? b : OH_BINS - 1; }

// per-pixel epilogue shared by both stats kernels: (m, s, z_label) -> p_label, loss, histogram
__device__ __forceinline__ void ohem_pixel(bool valid, float m, float s, float zl, float& pl, float& ce, u32* s_hist,
                                           u32& nvalid) {
    pl = __uint_as_float(OH_INVALID_BITS);
    ce = 0.f;
    if (valid) {
        pl = sm_prob(zl, m, s);
        ce = __fsub_rn(__fadd_rn(m, logf(s)), zl);          // -(z_label - (max + log sum)), ATen's log_softmax + nll
        atomicAdd(s_hist + ohem_bin0(__float_as_uint(pl)), 1u);
        ++nvalid;
    }
}

template <int CT, typename LT>
__global__ void __launch_bounds__(OH_TPB, 2) ohem_stats_v4(OhemParams p) {
    __shared__ u32 s_hist[OH_BINS];
    const int tid = threadIdx.x;
    for (int i = tid; i < OH_BINS; i += OH_TPB) s_hist[i] = 0;
    __syncthreads();
    constexpr int TILE_PX = OH_TPB * 4;
    const long long tpi = (p.HW + TILE_PX - 1) / TILE_PX;
    const long long ntiles = tpi * p.N;
    u32 oob = 0, nvalid = 0;
    for (long long t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const int n = (int)(t / tpi);
        const long long q0 = (t - (long long)n * tpi) * TILE_PX + tid * 4;
        if (q0 >= p.HW) continue;
        const float* lp = p.logits + (size_t)n * CT * p.HW + q0;
        float4 v[CT];
#pragma unroll
        for (int c = 0; c < CT; ++c) v[c] = ld_stream4(lp + (size_t)c * p.HW);
        int lab[4];
        const size_t px = (size_t)n * p.HW + q0;
        load_labels4<LT>(p.labels, px, lab);
        float m[4] = {v[0].x, v[0].y, v[0].z, v[0].w};
#pragma unroll
        for (int c = 1; c < CT; ++c) {
            m[0] = fmaxf(m[0], v[c].x); m[1] = fmaxf(m[1], v[c].y);
            m[2] = fmaxf(m[2], v[c].z); m[3] = fmaxf(m[3], v[c].w);
        }
        // the label's logit: one more scalar load (an L2 hit) instead of selecting among the 4 x CT registers
        bool valid[4];
        float zl[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            valid[j] = ohem_valid(p, lab[j], oob);
            zl[j] = valid[j] ? __ldg(lp + (size_t)lab[j] * p.HW + j) : 0.f;
        }
        float s[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int c = 0; c < CT; ++c) {
            s[0] = __fadd_rn(s[0], sm_exp(v[c].x, m[0])); s[1] = __fadd_rn(s[1], sm_exp(v[c].y, m[1]));
            s[2] = __fadd_rn(s[2], sm_exp(v[c].z, m[2])); s[3] = __fadd_rn(s[3], sm_exp(v[c].w, m[3]));
        }
        float pl[4], ce[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) ohem_pixel(valid[j], m[j], s[j], zl[j], pl[j], ce[j], s_hist, nvalid);
        *reinterpret_cast<float4*>(p.p_lab + px) = make_float4(pl[0], pl[1], pl[2], pl[3]);
        *reinterpret_cast<float4*>(p.ce + px) = make_float4(ce[0], ce[1], ce[2], ce[3]);
        *reinterpret_cast<float4*>(p.pix_m + px) = make_float4(m[0], m[1], m[2], m[3]);
        *reinterpret_cast<float4*>(p.pix_s + px) = make_float4(s[0], s[1], s[2], s[3]);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) nvalid += __shfl_xor_sync(FULL_MASK, nvalid, o);
    if ((tid & 31) == 0 && nvalid) atomicAdd(p.ctrl + OH_NVALID, nvalid);
    __syncthreads();
    for (int i = tid; i < OH_BINS; i += OH_TPB)
        if (s_hist[i]) atomicAdd(p.hist + i, s_hist[i]);
    if (oob) atomicOr(p.status, STATUS_LABEL_OOB);
}

template <typename LT>
__global__ void __launch_bounds__(OH_TPB) ohem_stats_generic(OhemParams p) {
    __shared__ u32 s_hist[OH_BINS];
    const int tid = threadIdx.x, C = p.C;
    for (int i = tid; i < OH_BINS; i += OH_TPB) s_hist[i] = 0;
    __syncthreads();
    u32 oob = 0, nvalid = 0;
    for (long long px = (long long)blockIdx.x * OH_TPB + tid; px < p.P; px += (long long)gridDim.x * OH_TPB) {
        const long long n = px / p.HW, q = px - n * p.HW;
        const float* lp = p.logits + (size_t)n * C * p.HW + q;
        const int lab = load_label<LT>(p.labels, (size_t)px);
        float m = __ldg(lp);
        for (int c = 1; c < C; ++c) m = fmaxf(m, __ldg(lp + (size_t)c * p.HW));
        float s = 0.f, zl = 0.f;
        for (int c = 0; c < C; ++c) {
            const float z = __ldg(lp + (size_t)c * p.HW);
            s = __fadd_rn(s, sm_exp(z, m));
            if (lab == c) zl = z;
        }
        float pl, ce;
        ohem_pixel(ohem_valid(p, lab, oob), m, s, zl, pl, ce, s_hist, nvalid);
        p.p_lab[px] = pl; p.ce[px] = ce; p.pix_m[px] = m; p.pix_s[px] = s;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) nvalid += __shfl_xor_sync(FULL_MASK, nvalid, o);
    if ((tid & 31) == 0 && nvalid) atomicAdd(p.ctrl + OH_NVALID, nvalid);
    __syncthreads();
    for (int i = tid; i < OH_BINS; i += OH_TPB)
        if (s_hist[i]) atomicAdd(p.hist + i, s_hist[i]);
    if (oob) atomicOr(p.status, STATUS_LABEL_OOB);
}

// ---- select + reduce: one cooperative kernel ---------------------------------------------------------------------
// Every CTA walks the 1024-bin histogram of a level itself (4 KB from L2), so no CTA waits for a "picker"; the two
// refinement histograms are built by the whole grid with a grid barrier each.  When the first level already shows that the
// order statistic lies below thresh (the usual case: thresh = 0.7, min_kept a small fraction of the pixels) there is no
// barrier at all and the kernel goes straight to the masked mean.
__device__ __forceinline__ void ohem_grid_barrier(u32* ctr, int* status) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(ctr, 1u);
        u32 spins = 0;
        while (ld_relaxed(ctr) < gridDim.x) {
            if (++spins > SPIN_LIMIT) { atomicOr(status, STATUS_SPIN_TIMEOUT); __threadfence_system(); __trap(); }
            __nanosleep(32);
        }
        __threadfence();
    }
    __syncthreads();
}

// bin of `hist` (OH_BINS counters) that holds rank `rank`, and the rank inside that bin; all threads get the result
__device__ __forceinline__ void ohem_pick(const u32* hist, u32 rank, u32& bin, u32& rank_in_bin, u32* s_tmp) {
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    constexpr int PER = OH_BINS / OH_TPB;
    u32 c[PER], tsum = 0;
#pragma unroll
    for (int k = 0; k < PER; ++k) { c[k] = __ldcg(hist + t * PER + k); tsum += c[k]; }
    u32 incl = tsum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const u32 up = __shfl_up_sync(FULL_MASK, incl, o);
        if (lane >= o) incl += up;
    }
    __syncthreads();                                   // s_tmp may still be read from the previous call
    if (lane == 31) s_tmp[warp] = incl;
    if (t == 0) { s_tmp[OH_TPB / 32] = OH_BINS - 1; s_tmp[OH_TPB / 32 + 1] = 0; }   // rank beyond the counts (NaNs): last bin
    __syncthreads();
    u32 base = 0;
    for (int w2 = 0; w2 < warp; ++w2) base += s_tmp[w2];
    u32 excl = base + incl - tsum;
    if (excl <= rank && rank < excl + tsum) {
#pragma unroll
        for (int k = 0; k < PER; ++k) {
            if (rank >= excl && rank < excl + c[k]) { s_tmp[OH_TPB / 32] = (u32)(t * PER + k); s_tmp[OH_TPB / 32 + 1] = rank - excl; }
            excl += c[k];
        }
    }
    __syncthreads();
    bin = s_tmp[OH_TPB / 32];
    rank_in_bin = s_tmp[OH_TPB / 32 + 1];
}

// histogram of the next 10 bits among the elements that share the prefix chosen so far
__device__ __forceinline__ void ohem_refine_hist(const OhemParams& p, int level, u32 prefix, u32* s_hist) {
    const int tid = threadIdx.x;
    for (int i = tid; i < OH_BINS; i += OH_TPB) s_hist[i] = 0;
    __syncthreads();
    const int hi = 30 - 10 * level, lo = 20 - 10 * level;
    const u32* bits = reinterpret_cast<const u32*>(p.p_lab);
    for (long long i = (long long)blockIdx.x * OH_TPB + tid; i < p.P; i += (long long)gridDim.x * OH_TPB) {
        const u32 b = __ldg(bits + i);
        if ((b >> hi) == prefix) atomicAdd(s_hist + ((b >> lo) & (OH_BINS - 1)), 1u);
    }
    __syncthreads();
    for (int i = tid; i < OH_BINS; i += OH_TPB)
        if (s_hist[i]) atomicAdd(p.hist + level * OH_BINS + i, s_hist[i]);
}

__global__ void __launch_bounds__(OH_TPB) ohem_select_reduce_kernel(OhemParams p) {
    __shared__ u32 s_hist[OH_BINS];
    __shared__ u32 s_tmp[OH_TPB / 32 + 2];
    __shared__ double s_sum[OH_TPB / 32];
    __shared__ u32 s_cnt[OH_TPB / 32];
    const int tid = threadIdx.x;
    const u32 n_valid = p.ctrl[OH_NVALID];
    float thr = p.thresh;                                           // n_valid == 0: nothing is kept, the mean is 0/0
    if (n_valid) {
        const long long last = (long long)n_valid - 1;
        u32 rank = (u32)(p.min_kept < last ? p.min_kept : last);   // min(min_kept, numel - 1), :33
        u32 prefix, bin;
        ohem_pick(p.hist, rank, prefix, rank, s_tmp);
        // every value of that bin is below thresh's bin: the order statistic is below thresh and max() picks thresh (:34)
        const bool settled = p.thresh > 0.f && p.thresh == p.thresh && prefix < ohem_bin0(__float_as_uint(p.thresh));
        if (!settled) {                                             // uniform over the grid: every CTA saw the same counts
            ohem_refine_hist(p, 1, prefix, s_hist);
            ohem_grid_barrier(p.ctrl + OH_BAR0, p.status);
            ohem_pick(p.hist + OH_BINS, rank, bin, rank, s_tmp);
            prefix = (prefix << 10) | bin;
            ohem_refine_hist(p, 2, prefix, s_hist);
            ohem_grid_barrier(p.ctrl + OH_BAR1, p.status);
            ohem_pick(p.hist + 2 * OH_BINS, rank, bin, rank, s_tmp);
            prefix = (prefix << 10) | bin;
            thr = fmaxf(__uint_as_float(prefix), p.thresh);
        }
    }
    if (blockIdx.x == 0 && tid == 0) p.ctrl[OH_THR] = __float_as_uint(thr);      // backward reads it

    double sum = 0.0;
    u32 cnt = 0;
    const long long quads = p.P / 4;                               // planes are 256-byte aligned: 128-bit loads
    for (long long i = (long long)blockIdx.x * OH_TPB + tid; i < quads; i += (long long)gridDim.x * OH_TPB) {
        const float4 pl = __ldg(reinterpret_cast<const float4*>(p.p_lab) + i);
        const float4 ce = __ldg(reinterpret_cast<const float4*>(p.ce) + i);
        float part = 0.f;                                          // :37-39 (strictly below the threshold)
        if (pl.x < thr) { part += ce.x; ++cnt; }
        if (pl.y < thr) { part += ce.y; ++cnt; }
        if (pl.z < thr) { part += ce.z; ++cnt; }
        if (pl.w < thr) { part += ce.w; ++cnt; }
        sum += (double)part;
    }
    for (long long i = quads * 4 + (long long)blockIdx.x * OH_TPB + tid; i < p.P; i += (long long)gridDim.x * OH_TPB)
        if (__ldg(p.p_lab + i) < thr) { sum += (double)__ldg(p.ce + i); ++cnt; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        sum += __shfl_xor_sync(FULL_MASK, sum, o);
        cnt += __shfl_xor_sync(FULL_MASK, cnt, o);
    }
    if ((tid & 31) == 0) { s_sum[tid >> 5] = sum; s_cnt[tid >> 5] = cnt; }
    __syncthreads();
    if (tid == 0) {
        for (int w = 1; w < OH_TPB / 32; ++w) { sum += s_sum[w]; cnt += s_cnt[w]; }
        if (cnt) { atomicAdd(p.ce_sum, sum); atomicAdd(p.ctrl + OH_KEPT, cnt); }
        __threadfence();
        if (atomicAdd(p.ctrl + OH_TICKET, 1u) == gridDim.x - 1) {      // last CTA: the mean (0/0 = NaN, like torch)
            __threadfence();
            const u32 kept = ld_relaxed(p.ctrl + OH_KEPT);
            const double total = *((volatile double*)p.ce_sum);
            *p.loss_out = (float)(total / (double)kept);
            p.ctrl[OH_INVKEPT] = __float_as_uint(kept ? 1.0f / (float)kept : 0.f);
        }
    }
}

template <int CT, typename LT>
__global__ void __launch_bounds__(OH_TPB, 3) ohem_backward_v4(OhemParams p, const float* __restrict__ go,
                                                           float* __restrict__ dlogits) {
    const int tid = threadIdx.x;
    constexpr int TILE_PX = OH_TPB * 4;
    const long long tpi = (p.HW + TILE_PX - 1) / TILE_PX;
    const long long ntiles = tpi * p.N;
    const float thr = __uint_as_float(p.ctrl[OH_THR]);
    const float scale = __ldg(go) * __uint_as_float(p.ctrl[OH_INVKEPT]);
    for (long long t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const int n = (int)(t / tpi);
        const long long q0 = (t - (long long)n * tpi) * TILE_PX + tid * 4;
        if (q0 >= p.HW) continue;
        const size_t px = (size_t)n * p.HW + q0;
        const float4 pl = __ldg(reinterpret_cast<const float4*>(p.p_lab + px));
        const bool k0 = pl.x < thr, k1 = pl.y < thr, k2 = pl.z < thr, k3 = pl.w < thr;
        float* dp = dlogits + (size_t)n * CT * p.HW + q0;
        if (!(k0 || k1 || k2 || k3)) {
#pragma unroll
            for (int c = 0; c < CT; ++c) st_stream4(dp + (size_t)c * p.HW, make_float4(0.f, 0.f, 0.f, 0.f));
            continue;
        }
        const float* lp = p.logits + (size_t)n * CT * p.HW + q0;
        const float4 m = __ldg(reinterpret_cast<const float4*>(p.pix_m + px));
        const float4 s = __ldg(reinterpret_cast<const float4*>(p.pix_s + px));
        int lab[4];
        load_labels4<LT>(p.labels, px, lab);
        // softmax = exp(z - max) * (1 / sum): one division per pixel; gradients carry the 1e-5 gate, not bit parity
        const float g0 = k0 ? scale : 0.f, g1 = k1 ? scale : 0.f, g2 = k2 ? scale : 0.f, g3 = k3 ? scale : 0.f;
        const float r0 = g0 * __frcp_rn(s.x), r1 = g1 * __frcp_rn(s.y), r2 = g2 * __frcp_rn(s.z), r3 = g3 * __frcp_rn(s.w);
        // classes are independent here (max and sum come from the forward pass): CH planes in flight per thread keep
        // the registers low enough for three CTAs per SM
        constexpr int CH = CT % 5 == 0 ? 5 : (CT % 4 == 0 ? 4 : 6);
#pragma unroll 1
        for (int c0 = 0; c0 < CT; c0 += CH) {
            float4 v[CH];
#pragma unroll
            for (int j = 0; j < CH; ++j)
                if (c0 + j < CT) v[j] = ld_stream4(lp + (size_t)(c0 + j) * p.HW);
#pragma unroll
            for (int j = 0; j < CH; ++j) {
                const int c = c0 + j;
                if (c < CT) {
                    float4 d;
                    d.x = sm_exp(v[j].x, m.x) * r0 - (lab[0] == c ? g0 : 0.f);
                    d.y = sm_exp(v[j].y, m.y) * r1 - (lab[1] == c ? g1 : 0.f);
                    d.z = sm_exp(v[j].z, m.z) * r2 - (lab[2] == c ? g2 : 0.f);
                    d.w = sm_exp(v[j].w, m.w) * r3 - (lab[3] == c ? g3 : 0.f);
                    st_stream4(dp + (size_t)c * p.HW, d);
                }
            }
        }
    }
}

template <typename LT>
__global__ void __launch_bounds__(OH_TPB) ohem_backward_generic(OhemParams p, const float* __restrict__ go,
                                                                float* __restrict__ dlogits) {
    const int C = p.C;
    const float thr = __uint_as_float(p.ctrl[OH_THR]);
    const float scale = __ldg(go) * __uint_as_float(p.ctrl[OH_INVKEPT]);
    for (long long px = (long long)blockIdx.x * OH_TPB + threadIdx.x; px < p.P; px += (long long)gridDim.x * OH_TPB) {
        const long long n = px / p.HW, q = px - n * p.HW;
        float* dp = dlogits + (size_t)n * C * p.HW + q;
        if (!(p.p_lab[px] < thr)) {
            for (int c = 0; c < C; ++c) dp[(size_t)c * p.HW] = 0.f;
            continue;
        }
        const float* lp = p.logits + (size_t)n * C * p.HW + q;
        const int lab = load_label<LT>(p.labels, (size_t)px);
        const float m = p.pix_m[px], s = p.pix_s[px];
        for (int c = 0; c < C; ++c)
            dp[(size_t)c * p.HW] = scale * (sm_prob(__ldg(lp + (size_t)c * p.HW), m, s) - (lab == c ? 1.f : 0.f));
    }
}

#define DISPATCH_LABEL(dtype, ...)                                               \
    switch (dtype) {                                                             \
        case B200SEG_LABEL_U8: { typedef uint8_t LT; __VA_ARGS__; } break;        \
        case B200SEG_LABEL_I32: { typedef int32_t LT; __VA_ARGS__; } break;       \
        case B200SEG_LABEL_I64: { typedef int64_t LT; __VA_ARGS__; } break;       \
        default: b200seg_set_error("unknown label dtype %d", dtype); return B200SEG_E_INVALID; \
    }

extern "C" int b200seg_ohem_workspace_bytes(int32_t n, int64_t hw, size_t* bytes) {
    if (!bytes || n < 0 || hw < 0 || (long double)n * hw >= (long double)(1u << 30)) {
        b200seg_set_error("b200seg_ohem_workspace_bytes: bad argument");
        return B200SEG_E_INVALID;
    }
    *bytes = ohem_layout((long long)n * hw).total;
    return 0;
}

static int ohem_fill(OhemParams& p, const float* logits, const void* labels, int32_t n, int32_t c, int64_t hw,
                     int64_t ignore_label, const void* workspace, size_t workspace_bytes) {
    if (n < 0 || hw < 0 || c < 1 || c > B200SEG_MAX_CLASSES || (long double)n * hw >= (long double)(1u << 30) ||
        (long double)n * hw * c >= (long double)(1ull << 31)) {
        b200seg_set_error("invalid shape: n_images=%d n_classes=%d plane=%lld", n, c, (long long)hw);
        return B200SEG_E_INVALID;
    }
    if (!workspace || ((!logits || !labels) && (long long)n * hw > 0)) {        // an empty batch has no data pointers
        b200seg_set_error("null pointer argument");
        return B200SEG_E_INVALID;
    }
    const OhemLayout L = ohem_layout((long long)n * hw);
    if (workspace_bytes < L.total || ((uintptr_t)workspace & 255)) {
        b200seg_set_error("ohem: workspace of %zu bytes (256-byte aligned) required, got %zu", L.total, workspace_bytes);
        return B200SEG_E_WORKSPACE;
    }
    char* ws = (char*)const_cast<void*>(workspace);
    p.logits = logits; p.labels = labels; p.N = n; p.C = c; p.HW = hw; p.P = (long long)n * hw;
    p.has_ignore = (ignore_label != B200SEG_NO_LABEL && ignore_label >= INT_MIN && ignore_label <= INT_MAX) ? 1 : 0;
    p.ignore = p.has_ignore ? (int)ignore_label : 0;
    p.p_lab = (float*)(ws + L.p_lab); p.ce = (float*)(ws + L.ce);
    p.pix_m = (float*)(ws + L.pix_m); p.pix_s = (float*)(ws + L.pix_s);
    p.hist = (u32*)(ws + L.hist); p.ctrl = (u32*)(ws + L.ctrl); p.ce_sum = (double*)(ws + L.ce_sum);
    return 0;
}

static bool ohem_v4_ok(const OhemParams& p, int label_dtype, const void* extra) {
    bool v4 = p.HW % 4 == 0 && ((uintptr_t)p.logits & 15) == 0 && ((uintptr_t)extra & 15) == 0;
    v4 = v4 && (label_dtype == B200SEG_LABEL_U8 ? ((uintptr_t)p.labels & 3) == 0 : ((uintptr_t)p.labels & 15) == 0);
    return v4 && (p.C == 8 || p.C == 17 || p.C == 25);
}

extern "C" int b200seg_ohem_ce_forward(const float* logits, const void* labels, int32_t label_dtype, int32_t n,
                                       int32_t c, int64_t hw, int64_t ignore_label, float thresh, int64_t min_kept,
                                       void* workspace, size_t workspace_bytes, float* loss_out, int32_t* status,
                                       void* stream) {
    OhemParams p;
    const int rc = ohem_fill(p, logits, labels, n, c, hw, ignore_label, workspace, workspace_bytes);
    if (rc) return rc;
    if (!loss_out || !status) { b200seg_set_error("null pointer argument"); return B200SEG_E_INVALID; }
    if (min_kept < 0) { b200seg_set_error("ohem: min_kept must be >= 0"); return B200SEG_E_INVALID; }
    p.thresh = thresh; p.min_kept = min_kept; p.loss_out = loss_out; p.status = status;
    cudaStream_t st = (cudaStream_t)stream;
    const OhemLayout L = ohem_layout(p.P);
    CUDA_TRY(cudaMemsetAsync((char*)workspace + L.hist, 0, L.total - L.hist, st));
    const int sms = b200seg_sm_count();
    if (p.P == 0) {                                             // mean over nothing
        CUDA_TRY(cudaMemsetAsync(loss_out, 0xFF, sizeof(float), st));       // NaN
        return 0;
    }
    if (ohem_v4_ok(p, label_dtype, nullptr)) {
        const long long tiles = (long long)n * ((hw + OH_TPB * 4 - 1) / (OH_TPB * 4));
        const int grid = (int)(tiles < (long long)sms * 4 ? tiles : (long long)sms * 4);
        DISPATCH_LABEL(label_dtype, {
            if (c == 8) ohem_stats_v4<8, LT><<<grid, OH_TPB, 0, st>>>(p);
            else if (c == 17) ohem_stats_v4<17, LT><<<grid, OH_TPB, 0, st>>>(p);
            else ohem_stats_v4<25, LT><<<grid, OH_TPB, 0, st>>>(p);
        });
    } else {
        const long long blocks = (p.P + OH_TPB - 1) / OH_TPB;
        const int grid = (int)(blocks < (long long)sms * 8 ? blocks : (long long)sms * 8);
        DISPATCH_LABEL(label_dtype, ohem_stats_generic<LT><<<grid, OH_TPB, 0, st>>>(p));
    }
    LAUNCH_CHECK("ohem_stats");
    {   // cooperative launch: the grid barriers of the refinement need every CTA resident
        static int occ[64] = {0};
        int dev = 0;
        CUDA_TRY(cudaGetDevice(&dev));
        int per_sm = (dev >= 0 && dev < 64) ? occ[dev] : 0;
        if (per_sm == 0) {
            CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, ohem_select_reduce_kernel, OH_TPB, 0));
            if (per_sm < 1) per_sm = 1;
            if (dev >= 0 && dev < 64) occ[dev] = per_sm;
        }
        if (per_sm > 4) per_sm = 4;
        const long long blocks = (p.P / 4 + OH_TPB - 1) / OH_TPB + 1;
        const int rgrid = (int)(blocks < (long long)sms * per_sm ? blocks : (long long)sms * per_sm);
        void* args[] = {(void*)&p};
        CUDA_TRY(cudaLaunchCooperativeKernel((const void*)ohem_select_reduce_kernel, dim3(rgrid), dim3(OH_TPB), args, 0, st));
    }
    LAUNCH_CHECK("ohem select / reduce");
    return 0;
}

extern "C" int b200seg_ohem_ce_backward(const float* logits, const void* labels, int32_t label_dtype, int32_t n,
                                        int32_t c, int64_t hw, int64_t ignore_label, const void* workspace,
                                        size_t workspace_bytes, const float* grad_out, float* dlogits, void* stream) {
    OhemParams p;
    const int rc = ohem_fill(p, logits, labels, n, c, hw, ignore_label, workspace, workspace_bytes);
    if (rc) return rc;
    if (p.P == 0) return 0;
    if (!grad_out || !dlogits) { b200seg_set_error("null pointer argument"); return B200SEG_E_INVALID; }
    cudaStream_t st = (cudaStream_t)stream;
    const int sms = b200seg_sm_count();
    if (ohem_v4_ok(p, label_dtype, dlogits)) {
        const long long tiles = (long long)n * ((hw + OH_TPB * 4 - 1) / (OH_TPB * 4));
        const int grid = (int)(tiles < (long long)sms * 4 ? tiles : (long long)sms * 4);
        DISPATCH_LABEL(label_dtype, {
            if (c == 8) ohem_backward_v4<8, LT><<<grid, OH_TPB, 0, st>>>(p, grad_out, dlogits);
            else if (c == 17) ohem_backward_v4<17, LT><<<grid, OH_TPB, 0, st>>>(p, grad_out, dlogits);
            else ohem_backward_v4<25, LT><<<grid, OH_TPB, 0, st>>>(p, grad_out, dlogits);
        });
    } else {
        const long long blocks = (p.P + OH_TPB - 1) / OH_TPB;
        const int grid = (int)(blocks < (long long)sms * 8 ? blocks : (long long)sms * 8);
        DISPATCH_LABEL(label_dtype, ohem_backward_generic<LT><<<grid, OH_TPB, 0, st>>>(p, grad_out, dlogits));
    }
    LAUNCH_CHECK("ohem_backward");
    return 0;
}
