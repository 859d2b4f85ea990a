// Lovasz-Softmax forward + backward for sm_100a.  Replaces losses/LovaszSoftmax.py:19-95 of the reference.
//
// Pipeline (all on one stream, no host sync, no allocation):
//   K1 stats      read logits once: per-pixel softmax max/sum (kept, 8 B/px), per-segment foreground count and
//                 max key (= min foreground error -> sort threshold), optional fused argmax + confusion matrix
//   K1c absent    (keep_absent only) max_i p_c(i) for considered classes without foreground
//   K1b finalize  thresholds, log-thresholds, key widths, class weights 1/n_present(/n_images)
//   K2 emit       re-read logits: every (pixel, class) with error >= threshold becomes a candidate
//                 (key = 0x3F800000 - bits(error), value = pixel<<1 | fg), written in pixel order per segment
//                 (smem bitmask ranks + chained scan across tiles) so a stable sort gives canonical tie order
//   K3..K4 sort   segmented stable LSD radix sort (sort.cuh)
//   K5 jaccard    scan of fg flags in sorted order -> Jaccard gradient, loss partials, per-candidate g
//   K5b loss      mean over present classes (sequential fp32, class order) and over images
//   K6 backward   re-read logits: dz_k = go * p_k (g_k - sum_j g_j p_j), sparse g gathered per pixel
//
// Exactness notes: candidates are a superset of every element with non-zero Jaccard gradient (SURVEY.md §7.3,
// zero-tail), so pruning changes nothing; p_c is computed by the same inlined fp32 sequence in every pass.
#include "b200seg.h"
#include "common.cuh"
#include "sort.cuh"
#include "pipe.cuh"
#include <cstdlib>

#define STATS_TPB 128
#define EMIT_TPB 256
#define BWD_TPB 128
#define JAC_TPB SORT_TPB

enum { CTRL_STATUS = 8 };

struct LovaszParams {
    const float* logits;
    const void* labels;
    int N, C;
    long long HW, P, cap;
    int per_image, has_filter, filter, keep_absent, need_grad, dbg;
    u32 class_mask;
    int groups, n_seg;
    // workspace
    u32* ctrl;
    u32 *seg_fg, *seg_maxkey, *seg_maxp, *seg_count, *grp_valid, *seg_bits;
    double* seg_loss;
    float *seg_thr, *seg_logthr, *seg_w;
    float *pix_m, *pix_s, *gown, *gbg;
    u32 *run_cnt, *run_prefix;          // [groups*n_runs][C] candidate counts per emission chunk; [n_seg][n_runs+1]
    int n_runs, tiles_per_chunk;        // emission chunks per group, emission tiles per chunk
    long long run_stride, src_cap;      // slots per chunk, holey segment stride (= n_runs * run_stride)
    u32 *keysA, *valsA, *keysB, *valsB;
    // fused confusion matrix
    unsigned long long* cm;
    int has_drop, drop;
    int* status;
    float* loss_out;
};

struct LovaszLayout {
    size_t ctrl, seg_fg, seg_maxkey, seg_maxp, seg_count, grp_valid, seg_loss, zero_end;
    size_t seg_thr, seg_logthr, seg_w, seg_bits, run_cnt, run_prefix;
    size_t pix_m, pix_s, gown, keysA, valsA, keysB, valsB, sort_scratch, total;
    SortScratch sort;
};

// Emission geometry for one pixel-vector width: tiles of EMIT_TPB*vec pixels never straddle images; a chunk is
// `tpc` consecutive tiles of one group and owns run_stride = tpc * tile_px candidate slots per class.
struct EmitGeom { long long tile_px, tpi, tpg, tpc, n_runs, run_stride, src_cap; };
#define EMIT_WARP_TILE 32            // pixels per warp tile of the pipelined emission kernel
static EmitGeom emit_geom(int N, long long HW, int per_image, long long tile_px) {
    EmitGeom G;
    G.tile_px = tile_px;
    G.tpi = (HW + G.tile_px - 1) / G.tile_px;
    G.tpg = per_image ? G.tpi : G.tpi * N;
    const long long total_tiles = G.tpi * N;
    const long long target_chunks = tile_px <= 64 ? 4096 : 1024;   // chunks over the whole batch
    G.tpc = (total_tiles + target_chunks - 1) / target_chunks;
    if (G.tpc < 1) G.tpc = 1;
    G.n_runs = (G.tpg + G.tpc - 1) / G.tpc;
    if (G.n_runs < 1) G.n_runs = 1;
    G.run_stride = G.tpc * G.tile_px;
    G.src_cap = G.n_runs * G.run_stride;
    return G;
}

static LovaszLayout lovasz_layout(int N, int C, long long HW, int per_image) {
    LovaszLayout L;
    const long long P = (long long)N * HW;
    const int groups = per_image ? N : 1;
    const size_t S = (size_t)groups * C;
    const size_t CP = (size_t)C * P;
    size_t cap_max = 0, runs_max = 0;                      // every emission variant must fit
    for (long long tile_px : {(long long)EMIT_TPB * 4, (long long)EMIT_TPB, (long long)EMIT_WARP_TILE}) {
        const EmitGeom G = emit_geom(N, HW, per_image, tile_px);
        if ((size_t)G.src_cap > cap_max) cap_max = (size_t)G.src_cap;
        if ((size_t)G.n_runs > runs_max) runs_max = (size_t)G.n_runs;
    }
    const size_t holey = S * cap_max;                      // >= CP
    const size_t runs = (size_t)groups * runs_max;
    size_t o = 0;
    L.ctrl = o;       o = align_up(o + 256, 256);
    L.seg_fg = o;     o = align_up(o + 4 * S, 256);
    L.seg_maxkey = o; o = align_up(o + 4 * S, 256);
    L.seg_maxp = o;   o = align_up(o + 4 * S, 256);
    L.seg_count = o;  o = align_up(o + 4 * S, 256);
    L.grp_valid = o;  o = align_up(o + 4 * (size_t)groups, 256);
    L.seg_loss = o;   o = align_up(o + 8 * S, 256);
    L.zero_end = o;
    L.seg_thr = o;    o = align_up(o + 4 * S, 256);
    L.seg_logthr = o; o = align_up(o + 4 * S, 256);
    L.seg_w = o;      o = align_up(o + 4 * S, 256);
    L.seg_bits = o;   o = align_up(o + 4 * S, 256);
    L.run_cnt = o;    o = align_up(o + 4 * runs * C, 256);
    L.run_prefix = o; o = align_up(o + 4 * (runs + (size_t)groups) * C, 256);
    L.pix_m = o;      o = align_up(o + 4 * (size_t)P, 256);
    L.pix_s = o;      o = align_up(o + 4 * (size_t)P, 256);
    L.gown = o;       o = align_up(o + 4 * (size_t)P, 256);
    L.keysA = o;      o = align_up(o + 4 * holey, 256);
    L.valsA = o;      o = align_up(o + 4 * holey, 256);
    L.keysB = o;      o = align_up(o + 4 * CP, 256);
    L.valsB = o;      o = align_up(o + 4 * CP, 256);
    L.sort = sort_scratch_layout((int)S, (long long)CP);
    L.sort_scratch = o; o = align_up(o + L.sort.total, 256);
    L.total = o;
    return L;
}

// --------------------------------------------------------------------------------------------------------------
// candidate predicate pieces (shared by K2 and K6 so both passes select exactly the same (pixel, class) pairs)
// --------------------------------------------------------------------------------------------------------------
// Conservative log-domain pre-test: a true candidate (p >= thr) always satisfies z >= theta + log(thr).
__device__ __forceinline__ float pre_theta(float m, float s) {
    return m + __logf(s) - (1e-3f + 1e-6f * fabsf(m));
}
// exact test; err = |fg - p|
__device__ __forceinline__ bool exact_accept(float z, float m, float s, bool fg, float thr, float& err, float& pr) {
    pr = sm_prob(z, m, s);
    if (fg) { err = __fsub_rn(1.0f, pr); return true; }
    err = pr;
    return pr >= thr;
}
#define THR_INACTIVE 2.0f
__device__ __forceinline__ bool thr_active(float thr) { return thr <= 1.5f; }

// --------------------------------------------------------------------------------------------------------------
// K1: stats (+ fused confusion matrix)
// --------------------------------------------------------------------------------------------------------------
struct StatsSmem {
    u32 fg[B200SEG_MAX_CLASSES];
    u32 key[B200SEG_MAX_CLASSES];
    u32 valid;
    u32 oob;
    u32 cm[B200SEG_MAX_CLASSES * B200SEG_MAX_CLASSES];
};

__device__ __forceinline__ void stats_flush_group(const LovaszParams& p, StatsSmem& sm, int g) {
    __syncthreads();
    const int tid = threadIdx.x;
    if (tid < p.C) {
        const size_t seg = (size_t)g * p.C + tid;
        if (sm.fg[tid]) { atomicAdd(p.seg_fg + seg, sm.fg[tid]); atomicMax(p.seg_maxkey + seg, sm.key[tid]); }
        sm.fg[tid] = 0; sm.key[tid] = 0;
    }
    if (tid == 0) { if (sm.valid) atomicAdd(p.grp_valid + g, sm.valid); sm.valid = 0; }
    __syncthreads();
}
__device__ __forceinline__ void stats_flush_cm(const LovaszParams& p, StatsSmem& sm) {
    __syncthreads();
    if (p.cm) {
        for (int i = threadIdx.x; i < p.C * p.C; i += blockDim.x)
            if (sm.cm[i]) atomicAdd(p.cm + i, (unsigned long long)sm.cm[i]);
        if (threadIdx.x == 0 && sm.oob) atomicOr(p.status, STATUS_LABEL_OOB);
    }
}
__device__ __forceinline__ void stats_pixel_tail(const LovaszParams& p, StatsSmem& sm, int lab, float e_lab, float s,
                                                 int arg, int C, u32& nvalid) {
    const bool filtered = p.has_filter && lab == p.filter;
    if (!filtered) {
        ++nvalid;
        if ((unsigned)lab < (unsigned)C) {
            const float pr = __fdiv_rn(e_lab, s);
            const u32 key = err_key(__fsub_rn(1.0f, pr));
            atomicAdd(&sm.fg[lab], 1u);
            atomicMax(&sm.key[lab], key);
        }
    }
    if (p.cm) {
        if (!(p.has_drop && lab == p.drop)) {
            if ((unsigned)lab < (unsigned)C) atomicAdd(&sm.cm[arg * C + lab], 1u);
            else sm.oob = 1;
        }
    }
}

// VEC consecutive pixels per thread, all C logits of those pixels in registers (C*VEC independent exp chains).
template <int CT, int VEC, int TPB, typename LT>
__global__ void __launch_bounds__(TPB) stats_kernel_vec(LovaszParams p) {
    __shared__ StatsSmem sm;
    const int tid = threadIdx.x;
    constexpr int TILE_PX = TPB * VEC;
    const long long tpi = (p.HW + TILE_PX - 1) / TILE_PX;
    const long long ntiles = tpi * p.N;
    const long long t0 = ntiles * blockIdx.x / gridDim.x, t1 = ntiles * (blockIdx.x + 1) / gridDim.x;
    for (int i = tid; i < (int)(sizeof(StatsSmem) / 4); i += TPB) ((u32*)&sm)[i] = 0;
    __syncthreads();
    int cur_g = -1;
    u32 nvalid = 0;
    for (long long t = t0; t < t1; ++t) {
        const int n = (int)(t / tpi);
        const long long q0 = (t - (long long)n * tpi) * TILE_PX + tid * VEC;
        const int g = p.per_image ? n : 0;
        if (g != cur_g) {
            if (cur_g >= 0) { if (nvalid) atomicAdd(&sm.valid, nvalid); nvalid = 0; stats_flush_group(p, sm, cur_g); }
            cur_g = g;
        }
        if (q0 >= p.HW) continue;
        const float* lp = p.logits + (size_t)n * CT * p.HW + q0;
        const size_t px = (size_t)n * p.HW + q0;
        int lab[VEC];
        if constexpr (VEC == 4) load_labels4<LT>(p.labels, px, lab);
        else {
#pragma unroll
            for (int j = 0; j < VEC; ++j) lab[j] = load_label<LT>(p.labels, px + j);
        }
        float z[CT][VEC];
#pragma unroll
        for (int c = 0; c < CT; ++c) {
            if constexpr (VEC == 4) {
                const float4 v = ld_stream4(lp + (size_t)c * p.HW);
                z[c][0] = v.x; z[c][1] = v.y; z[c][2] = v.z; z[c][3] = v.w;
            } else {
                const float2 v = ld_stream2(lp + (size_t)c * p.HW);
                z[c][0] = v.x; z[c][1] = v.y;
            }
        }
        // the logit of the pixel's own class: a dependent re-read (L2 hit) is cheaper than a C-way select chain
        float zl[VEC];
#pragma unroll
        for (int j = 0; j < VEC; ++j) zl[j] = (unsigned)lab[j] < (unsigned)CT ? __ldg(lp + (size_t)lab[j] * p.HW + j) : 0.f;
        float mo[VEC], so[VEC];
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
            float m = z[0][j];
#pragma unroll
            for (int c = 1; c < CT; ++c) m = fmaxf(m, z[c][j]);
            float s = 0.f;
#pragma unroll
            for (int c = 0; c < CT; ++c) s = __fadd_rn(s, sm_exp(z[c][j], m));
            int arg = 0;
            if (p.cm) {
#pragma unroll
                for (int c = CT - 1; c >= 0; --c) arg = (z[c][j] == m) ? c : arg;      // first maximum
                if (s != s || m != m) {                   // NaN / inf among the logits: torch's argmax lets NaN win
                    float best = z[0][j];
                    arg = 0;
#pragma unroll
                    for (int c = 1; c < CT; ++c) argmax_step(z[c][j], c, best, arg);
                }
            }
            mo[j] = m; so[j] = s;
            stats_pixel_tail(p, sm, lab[j], sm_exp(zl[j], m), s, arg, CT, nvalid);
        }
        if constexpr (VEC == 4) {
            *(float4*)(p.pix_m + px) = make_float4(mo[0], mo[1], mo[2], mo[3]);
            *(float4*)(p.pix_s + px) = make_float4(so[0], so[1], so[2], so[3]);
        } else {
            *(float2*)(p.pix_m + px) = make_float2(mo[0], mo[1]);
            *(float2*)(p.pix_s + px) = make_float2(so[0], so[1]);
        }
    }
    if (cur_g >= 0) { if (nvalid) atomicAdd(&sm.valid, nvalid); stats_flush_group(p, sm, cur_g); }
    stats_flush_cm(p, sm);
}

// generic: any C <= 32, any plane size / alignment; one pixel per thread, logits re-read from L1/L2
template <typename LT>
__global__ void __launch_bounds__(STATS_TPB) stats_kernel_generic(LovaszParams p) {
    __shared__ StatsSmem sm;
    const int tid = threadIdx.x;
    const int C = p.C;
    constexpr int TILE_PX = STATS_TPB;
    const long long tpi = (p.HW + TILE_PX - 1) / TILE_PX;
    const long long ntiles = tpi * p.N;
    const long long t0 = ntiles * blockIdx.x / gridDim.x, t1 = ntiles * (blockIdx.x + 1) / gridDim.x;
    for (int i = tid; i < (int)(sizeof(StatsSmem) / 4); i += STATS_TPB) ((u32*)&sm)[i] = 0;
    __syncthreads();
    int cur_g = -1;
    u32 nvalid = 0;
    for (long long t = t0; t < t1; ++t) {
        const int n = (int)(t / tpi);
        const long long q = (t - (long long)n * tpi) * TILE_PX + tid;
        const int g = p.per_image ? n : 0;
        if (g != cur_g) {
            if (cur_g >= 0) { if (nvalid) atomicAdd(&sm.valid, nvalid); nvalid = 0; stats_flush_group(p, sm, cur_g); }
            cur_g = g;
        }
        if (q >= p.HW) continue;
        const float* lp = p.logits + (size_t)n * C * p.HW + q;
        const size_t px = (size_t)n * p.HW + q;
        const int lab = load_label<LT>(p.labels, px);
        float m = __ldg(lp), best = m;
        int arg = 0;
        for (int c = 1; c < C; ++c) { const float v = __ldg(lp + (size_t)c * p.HW); m = fmaxf(m, v); argmax_step(v, c, best, arg); }
        float s = 0.f, e_lab = 0.f;
        for (int c = 0; c < C; ++c) {
            const float e = sm_exp(__ldg(lp + (size_t)c * p.HW), m);
            s = __fadd_rn(s, e);
            e_lab = (c == lab) ? e : e_lab;
        }
        p.pix_m[px] = m; p.pix_s[px] = s;
        stats_pixel_tail(p, sm, lab, e_lab, s, arg, C, nvalid);
    }
    if (cur_g >= 0) { if (nvalid) atomicAdd(&sm.valid, nvalid); stats_flush_group(p, sm, cur_g); }
    stats_flush_cm(p, sm);
}

// --------------------------------------------------------------------------------------------------------------
// K1c: max p_c over the valid pixels of a group, for considered classes without foreground (keep_absent mode)
//      reference: the loss term of an absent class degenerates to max_i p_c(i) (LovaszSoftmax.py:52-60 with fg == 0)
// --------------------------------------------------------------------------------------------------------------
template <typename LT>
__global__ void __launch_bounds__(256) absent_max_kernel(LovaszParams p) {
    __shared__ u32 s_max;
    const int seg = blockIdx.x;
    const int c = seg % p.C, g = seg / p.C;
    if (!((p.class_mask >> c) & 1u) || p.seg_fg[seg] > 0 || p.grp_valid[g] == 0) return;
    if (threadIdx.x == 0) s_max = 0;
    __syncthreads();
    u32 best = 0;
    const long long begin = (long long)g * p.cap, end = begin + p.cap;
    for (long long px = begin + (long long)blockIdx.y * blockDim.x + threadIdx.x; px < end;
         px += (long long)gridDim.y * blockDim.x) {
        const int lab = load_label<LT>(p.labels, (size_t)px);
        if (p.has_filter && lab == p.filter) continue;
        const long long n = px / p.HW, q = px - n * p.HW;
        const float z = __ldg(p.logits + ((size_t)n * p.C + c) * p.HW + q);
        const float pr = sm_prob(z, p.pix_m[px], p.pix_s[px]);
        best = max(best, __float_as_uint(pr));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) best = max(best, __shfl_xor_sync(FULL_MASK, best, o));
    if ((threadIdx.x & 31) == 0 && best) atomicMax(&s_max, best);
    __syncthreads();
    if (threadIdx.x == 0 && s_max) atomicMax(p.seg_maxp + seg, s_max);
}

// --------------------------------------------------------------------------------------------------------------
// K1b: per-group finalisation of the segment table
// --------------------------------------------------------------------------------------------------------------
__global__ void finalize_stats_kernel(LovaszParams p) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= p.groups) return;
    int nkept = 0;
    const bool any_valid = p.grp_valid[g] > 0;
    for (int c = 0; c < p.C; ++c) {
        const size_t seg = (size_t)g * p.C + c;
        const u32 fg = p.seg_fg[seg];
        const bool active = ((p.class_mask >> c) & 1u) && any_valid && (fg > 0 || p.keep_absent);
        float thr = THR_INACTIVE, logthr = __int_as_float(0x7f800000);
        u32 bits = 1;
        if (active) {
            const u32 thr_bits = fg > 0 ? (ONE_BITS - p.seg_maxkey[seg]) : p.seg_maxp[seg];
            thr = __uint_as_float(thr_bits);
            logthr = thr > 0.f ? logf(thr) : __int_as_float(0xff800000);
            const u32 maxkey = ONE_BITS - thr_bits;
            bits = maxkey ? (32 - __clz(maxkey)) : 1;
            ++nkept;
        }
        p.seg_thr[seg] = thr;
        p.seg_logthr[seg] = logthr;
        p.seg_bits[seg] = bits;
    }
    // d(mean)/d(term): the reference's mean() divides only when it averaged more than one value
    float w = 1.0f;
    if (p.groups > 1) w = w / (float)p.groups;
    if (nkept > 1) w = w / (float)nkept;
    for (int c = 0; c < p.C; ++c) {
        const size_t seg = (size_t)g * p.C + c;
        p.seg_w[seg] = thr_active(p.seg_thr[seg]) ? w : 0.f;
    }
}

// --------------------------------------------------------------------------------------------------------------
// K2: candidate emission.  Every chunk (a few consecutive tiles of one group) owns a private slot range per class,
//     so CTAs never talk to each other: within the chunk candidates are written in pixel order (shared-memory
//     bitmask ranks + running per-class offsets), chunks are ordered by construction, and the per-chunk counts are
//     turned into the sort's run prefix by run_scan_kernel.  A stable sort then yields the canonical tie order.
// --------------------------------------------------------------------------------------------------------------
template <int VEC, typename LT>
__global__ void __launch_bounds__(EMIT_TPB) emit_kernel(LovaszParams p) {
    constexpr int TILE_PX = EMIT_TPB * VEC;
    constexpr int WORDS = TILE_PX / 32;
    constexpr int NWARPS = EMIT_TPB / 32;
    __shared__ u32 s_mask[B200SEG_MAX_CLASSES][WORDS];
    __shared__ u32 s_wpre[B200SEG_MAX_CLASSES][WORDS];
    __shared__ u32 s_tot[B200SEG_MAX_CLASSES];
    __shared__ u32 s_run[B200SEG_MAX_CLASSES];
    __shared__ float s_thr[B200SEG_MAX_CLASSES], s_logthr[B200SEG_MAX_CLASSES];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int C = p.C;
    const long long tpi = (p.HW + TILE_PX - 1) / TILE_PX;
    const long long tpg = p.per_image ? tpi : tpi * p.N;
    const long long total_chunks = (long long)p.groups * p.n_runs;
    int cur_g = -1;

    for (long long chunk = blockIdx.x; chunk < total_chunks; chunk += gridDim.x) {
        const int g = (int)(chunk / p.n_runs);
        const long long r = chunk - (long long)g * p.n_runs;
        const long long gt0 = r * p.tiles_per_chunk;
        const long long gt1 = min(gt0 + (long long)p.tiles_per_chunk, tpg);
        __syncthreads();                                   // previous chunk fully written out
        if (g != cur_g) {
            if (tid < C) { s_thr[tid] = p.seg_thr[(size_t)g * C + tid]; s_logthr[tid] = p.seg_logthr[(size_t)g * C + tid]; }
            cur_g = g;
        }
        if (tid < B200SEG_MAX_CLASSES) s_run[tid] = 0;
        for (int i = tid; i < B200SEG_MAX_CLASSES * WORDS; i += EMIT_TPB) (&s_mask[0][0])[i] = 0;
        __syncthreads();

        for (long long gt = gt0; gt < gt1; ++gt) {
            const int n = p.per_image ? g : (int)(gt / tpi);
            const long long ti = p.per_image ? gt : gt - (long long)n * tpi;
            const long long q0 = ti * TILE_PX + (long long)tid * VEC;
            const bool inb = q0 < p.HW;
            const size_t px0 = (size_t)n * p.HW + q0;
            const float* lp = p.logits + (size_t)n * C * p.HW + q0;
            float m[VEC], s[VEC];
            int lab[VEC];
            u32 acc[VEC];
#pragma unroll
            for (int j = 0; j < VEC; ++j) acc[j] = 0;
            if (inb) {
                if constexpr (VEC == 4) {
                    const float4 mv = *(const float4*)(p.pix_m + px0), sv = *(const float4*)(p.pix_s + px0);
                    m[0] = mv.x; m[1] = mv.y; m[2] = mv.z; m[3] = mv.w;
                    s[0] = sv.x; s[1] = sv.y; s[2] = sv.z; s[3] = sv.w;
                    load_labels4<LT>(p.labels, px0, lab);
                } else {
                    m[0] = p.pix_m[px0]; s[0] = p.pix_s[px0]; lab[0] = load_label<LT>(p.labels, px0);
                }
                float theta[VEC];
                u32 pre[VEC];
#pragma unroll
                for (int j = 0; j < VEC; ++j) { theta[j] = pre_theta(m[j], s[j]); pre[j] = 0; }
#pragma unroll 5
                for (int c = 0; c < C; ++c) {
                    const float lt = s_logthr[c];
                    float v[VEC];
                    if constexpr (VEC == 4) {
                        const float4 x = __ldg((const float4*)(lp + (size_t)c * p.HW));
                        v[0] = x.x; v[1] = x.y; v[2] = x.z; v[3] = x.w;
                    } else {
                        v[0] = __ldg(lp + (size_t)c * p.HW);
                    }
#pragma unroll
                    for (int j = 0; j < VEC; ++j) pre[j] |= (v[j] >= theta[j] + lt) ? (1u << c) : 0u;
                }
#pragma unroll
                for (int j = 0; j < VEC; ++j) {
                    if (p.has_filter && lab[j] == p.filter) pre[j] = 0;
                    else if ((unsigned)lab[j] < (unsigned)C && thr_active(s_thr[lab[j]])) pre[j] |= 1u << lab[j];
                    u32 mm = pre[j];
                    while (mm) {
                        const int c = __ffs(mm) - 1;
                        mm &= mm - 1;
                        float err, pr;
                        if (exact_accept(__ldg(lp + (size_t)c * p.HW + j), m[j], s[j], c == lab[j], s_thr[c], err, pr))
                            acc[j] |= 1u << c;
                    }
                }
                u32 uni = 0;
#pragma unroll
                for (int j = 0; j < VEC; ++j) uni |= acc[j];
                const int bit0 = tid * VEC;
                while (uni) {
                    const int c = __ffs(uni) - 1;
                    uni &= uni - 1;
                    u32 nib = 0;
#pragma unroll
                    for (int j = 0; j < VEC; ++j) nib |= ((acc[j] >> c) & 1u) << j;
                    atomicOr(&s_mask[c][bit0 >> 5], nib << (bit0 & 31));
                }
            }
            __syncthreads();

            // exclusive popcount prefix over the words of every class
            for (int c = warp; c < C; c += NWARPS) {
                const u32 cnt = lane < WORDS ? __popc(s_mask[c][lane]) : 0;
                u32 v = cnt;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { const u32 x = __shfl_up_sync(FULL_MASK, v, o); if (lane >= o) v += x; }
                if (lane < WORDS) s_wpre[c][lane] = v - cnt;
                if (lane == 31) s_tot[c] = v;
            }
            __syncthreads();

            if (inb) {
#pragma unroll
                for (int j = 0; j < VEC; ++j) {
                    u32 mm = acc[j];
                    const int bit = tid * VEC + j;
                    while (mm) {
                        const int c = __ffs(mm) - 1;
                        mm &= mm - 1;
                        float err, pr;
                        const bool fg = c == lab[j];
                        exact_accept(__ldg(lp + (size_t)c * p.HW + j), m[j], s[j], fg, s_thr[c], err, pr);
                        const u32 rank = s_run[c] + s_wpre[c][bit >> 5] +
                                         __popc(s_mask[c][bit >> 5] & ((1u << (bit & 31)) - 1u));
                        const size_t slot = ((size_t)g * C + c) * (size_t)p.src_cap + (size_t)r * p.run_stride + rank;
                        p.keysA[slot] = err_key(err);
                        p.valsA[slot] = ((u32)(px0 + j) << 1) | (fg ? 1u : 0u);
                    }
                }
            }
            __syncthreads();
            if (tid < C) s_run[tid] += s_tot[tid];
            for (int i = tid; i < B200SEG_MAX_CLASSES * WORDS; i += EMIT_TPB) (&s_mask[0][0])[i] = 0;
            __syncthreads();
        }
        if (tid < C) p.run_cnt[(size_t)chunk * C + tid] = s_run[tid];
    }
}

// Pipelined emission (the fast path): one warp per chunk of consecutive 32-pixel warp tiles, logits prefetched with the
// warp-private cp.async ring (see WarpTile).  Ranks inside a tile come from per-class ballots, the running per-class
// offset of the chunk lives in lane c's register: no CTA barriers, no cross-warp traffic.  The tile sequence is walked with increments only (no integer divisions).
struct EmitCursor {                                       // position in the (group, chunk, tile-in-chunk) sequence
    u32 k, r, gt, ti;                                     // tile in chunk, chunk in group, tile in group, tile in image
    int g, n;                                             // group, image
};
__device__ __forceinline__ void emit_cursor_init(EmitCursor& c, u32 chunk, u32 tpc, u32 n_runs, u32 wtpi, int per_image) {
    c.g = (int)(chunk / n_runs);
    c.r = chunk - (u32)c.g * n_runs;
    c.k = 0;
    c.gt = c.r * tpc;
    c.n = per_image ? c.g : (int)(c.gt / wtpi);
    c.ti = per_image ? c.gt : c.gt - (u32)c.n * wtpi;
}
__device__ __forceinline__ void emit_cursor_next(EmitCursor& c, u32 tpc, u32 n_runs, u32 wtpi, int per_image) {
    ++c.k; ++c.gt; ++c.ti;
    if (!per_image && c.ti == wtpi) { c.ti = 0; ++c.n; }
    if (c.k == tpc) {
        c.k = 0;
        if (++c.r == n_runs) { c.r = 0; ++c.g; c.gt = 0; c.ti = 0; c.n = per_image ? c.g : 0; }
    }
}

template <int CT, int TPB, int STAGES, typename LT>
__global__ void __launch_bounds__(TPB) emit_kernel_async(LovaszParams p) {
    using W = WarpTile<CT, 1>;
    constexpr int WT = W::WT, NW = TPB / 32;
    static_assert(STAGES == 2, "the prefetch cursor runs exactly one tile ahead");
    extern __shared__ __align__(16) unsigned char pipe_smem_raw[];
    float (*Zall)[STAGES][CT][WT] = reinterpret_cast<float (*)[STAGES][CT][WT]>(pipe_smem_raw);
    __shared__ float s_thr[NW][B200SEG_MAX_CLASSES], s_logthr[NW][B200SEG_MAX_CLASSES];
    __shared__ u32 s_mask[NW][B200SEG_MAX_CLASSES], s_base[NW][B200SEG_MAX_CLASSES];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const u32 lt_mask = (1u << lane) - 1;
    float (*Z)[CT][WT] = Zall[warp];
    // all tile / chunk counts fit 32 bits (n_images * plane < 2^30)
    const u32 wtpi = (u32)((p.HW + WT - 1) / WT);
    const u32 tpg = p.per_image ? wtpi : wtpi * (u32)p.N;
    const u32 tpc = (u32)p.tiles_per_chunk, n_runs = (u32)p.n_runs;
    const u32 total_chunks = (u32)p.groups * n_runs;
    const u32 gw = blockIdx.x * NW + warp, nwarps = gridDim.x * NW;
    const u32 ch0 = (u32)((u64)total_chunks * gw / nwarps), ch1 = (u32)((u64)total_chunks * (gw + 1) / nwarps);
    const u32 ntile = (ch1 - ch0) * tpc;                  // this warp's slice of the (chunk, tile-in-chunk) sequence
    if (ntile == 0) return;

    EmitCursor cur, pre;
    emit_cursor_init(cur, ch0, tpc, n_runs, wtpi, p.per_image);
    pre = cur;
    auto prefetch = [&](const EmitCursor& c, bool live, int stage) {
        if (live && c.gt < tpg) W::prefetch(Z[stage], p.logits, c.n, (long long)c.ti * WT, p.HW, lane);
        cp_async_commit();
    };
    prefetch(pre, true, 0);

    int cur_g = -1;
    u32 run = 0;                                          // lane c: candidates of class c emitted so far in this chunk
    for (u32 it = 0; it < ntile; ++it) {
        const int stage = (int)(it & 1);
        __syncwarp();
        emit_cursor_next(pre, tpc, n_runs, wtpi, p.per_image);
        prefetch(pre, it + 1 < ntile, stage ^ 1);
        const int g = cur.g, n = cur.n;
        const u32 r = cur.r, k = cur.k;
        const bool exists = cur.gt < tpg;
        if (k == 0) run = 0;
        if (g != cur_g) {
            if (lane < CT) { s_thr[warp][lane] = p.seg_thr[(size_t)g * CT + lane]; s_logthr[warp][lane] = p.seg_logthr[(size_t)g * CT + lane]; }
            cur_g = g;
            __syncwarp();
        }
        const long long q = (long long)cur.ti * WT + lane;
        const bool inb = exists && q < p.HW;
        float m = 0.f, s = 1.f;
        int lab = -1;
        const size_t px = (size_t)n * p.HW + q;
        if (inb) { m = p.pix_m[px]; s = p.pix_s[px]; lab = load_label<LT>(p.labels, px); }
        cp_async_wait<STAGES - 1>();
        __syncwarp();
        const float (*Tz)[WT] = Z[stage];
        const float* thr = s_thr[warp];
        u32 acc = 0;                                      // accepted classes of this lane's pixel
        float e0 = 0.f, e1 = 0.f;                         // errors of the first two of them (the rest is recomputed)
        if (inb) {
            const float theta = pre_theta(m, s);
            u32 pm = 0;
#pragma unroll
            for (int c = 0; c < CT; ++c) pm |= (Tz[c][lane] >= theta + s_logthr[warp][c]) ? (1u << c) : 0u;
            if (p.has_filter && lab == p.filter) pm = 0;
            else if ((unsigned)lab < (unsigned)CT && thr_active(thr[lab & 31])) pm |= 1u << lab;
            while (pm) {
                const int c = __ffs(pm) - 1;
                pm &= pm - 1;
                float err, pr;
                if (exact_accept(Tz[c][lane], m, s, c == lab, thr[c], err, pr)) {
                    if (acc == 0) e0 = err; else if ((acc & (acc - 1)) == 0) e1 = err;
                    acc |= 1u << c;
                }
            }
        }
        if (__any_sync(FULL_MASK, acc != 0)) {
            // lane c collects the ballot of class c and reserves the slots; the table goes through shared memory
            u32 mine = 0;
#pragma unroll
            for (int c = 0; c < CT; ++c) {
                const u32 b = __ballot_sync(FULL_MASK, (acc >> c) & 1u);
                if (lane == c) mine = b;
            }
            __syncwarp();
            if (lane < CT) { s_mask[warp][lane] = mine; s_base[warp][lane] = run; run += __popc(mine); }
            __syncwarp();
            u32 mm = acc;
            int i = 0;
            while (mm) {                                   // this lane's own candidates
                const int c = __ffs(mm) - 1;
                mm &= mm - 1;
                const bool fg = c == lab;
                float err = i == 0 ? e0 : e1;
                if (i >= 2) { float pr; exact_accept(Tz[c][lane], m, s, fg, thr[c], err, pr); }
                ++i;
                const u32 rank = s_base[warp][c] + __popc(s_mask[warp][c] & lt_mask);
                const size_t slot = ((size_t)g * CT + c) * (size_t)p.src_cap + (size_t)r * p.run_stride + rank;
                p.keysA[slot] = err_key(err);
                p.valsA[slot] = ((u32)px << 1) | (fg ? 1u : 0u);
            }
        }
        if (k == tpc - 1 && lane < CT) p.run_cnt[((size_t)g * n_runs + r) * CT + lane] = run;
        emit_cursor_next(cur, tpc, n_runs, wtpi, p.per_image);
    }
    cp_async_wait<0>();
}

// per segment: exclusive prefix of the chunk counts (the sort's run prefix) and the segment's candidate count
__global__ void __launch_bounds__(256) run_scan_kernel(LovaszParams p) {
    const int lane = threadIdx.x & 31;
    const int seg = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (seg >= p.n_seg) return;
    const int g = seg / p.C, c = seg - g * p.C;
    u32* out = p.run_prefix + (size_t)seg * (p.n_runs + 1);
    u32 carry = 0;
    for (int base = 0; base < p.n_runs; base += 1024) {                 // 32 independent loads per lane, then the scans
        u32 x[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            const int r = base + i * 32 + lane;
            x[i] = r < p.n_runs ? p.run_cnt[((size_t)g * p.n_runs + r) * p.C + c] : 0;
        }
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            const int r = base + i * 32 + lane;
            u32 v = x[i];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const u32 y = __shfl_up_sync(FULL_MASK, v, o); if (lane >= o) v += y; }
            if (r < p.n_runs) out[r] = carry + v - x[i];
            carry += __shfl_sync(FULL_MASK, v, 31);
        }
    }
    if (lane == 0) { out[p.n_runs] = carry; p.seg_count[seg] = carry; }
}

// --------------------------------------------------------------------------------------------------------------
// K5: Jaccard gradient over the sorted candidates      reference: lovasz_grad, losses/LovaszSoftmax.py:83-95
// --------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(JAC_TPB) jaccard_kernel(LovaszParams p, SortArgs a) {
    __shared__ u32 s_wfg[SORT_WARPS];
    __shared__ double s_red[SORT_WARPS];
    __shared__ u32 s_excl;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const u32 le_mask = lane == 31 ? FULL_MASK : ((2u << lane) - 1u);
    const u32 total_tiles = a.tile_start[a.n_seg];
    const u32* keys = a.keys[1];
    const u32* vals = a.vals[1];
    for (u32 t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        __syncthreads();
        const uint4 d4 = a.tile_desc[t];
        const int seg = (int)d4.x;
        const u32 off = d4.y, n = d4.z;
        const u32 tis = off / SORT_TILE;
        const u32 tseg0 = t - tis;
        const size_t base = (size_t)seg * a.cap + off;
        const int c = seg % p.C;
        const float gts = (float)p.seg_fg[seg];
        const float w = p.seg_w[seg];

        u32 key[SORT_KPT], val[SORT_KPT];
        unsigned short floc[SORT_KPT];
        const u32 wbase = warp * (32 * SORT_KPT) + lane;
        u32 run = 0;
#pragma unroll
        for (int k = 0; k < SORT_KPT; ++k) {
            const u32 idx = wbase + k * 32;
            const bool valid = idx < n;
            key[k] = valid ? keys[base + idx] : 0;
            val[k] = valid ? vals[base + idx] : 0;
            const u32 b = __ballot_sync(FULL_MASK, valid && (val[k] & 1u));
            floc[k] = (unsigned short)(run + __popc(b & le_mask));
            run += __popc(b);
        }
        if (lane == 0) s_wfg[warp] = run;
        __syncthreads();
        u32 wexcl = 0, ttot = 0;
#pragma unroll
        for (int w2 = 0; w2 < SORT_WARPS; ++w2) { const u32 x = s_wfg[w2]; if (w2 < warp) wexcl += x; ttot += x; }
        if (warp == 0) {                                   // foreground flags in the tiles before this one (last sort pass)
            u32 e = 0;
            for (u32 i = lane; i < tis; i += 32) e += a.tile_fg[tseg0 + i];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) e += __shfl_xor_sync(FULL_MASK, e, o);
            if (lane == 0) s_excl = e;
        }
        __syncthreads();
        const u32 fbase = s_excl + wexcl;
        double acc = 0.0;
#pragma unroll
        for (int k = 0; k < SORT_KPT; ++k) {
            const u32 idx = wbase + k * 32;
            if (idx < n) {
                const u32 i = off + idx;                  // position in the segment's sorted order
                const u32 fgi = val[k] & 1u;
                const u32 F = fbase + floc[k];            // foreground among positions 0..i
                const u32 B = i + 1 - F;                  // background among positions 0..i
                const float J = 1.0f - __fdiv_rn(gts - (float)F, gts + (float)B);
                float grad = J;
                if (i > 0) {
                    const float Jp = 1.0f - __fdiv_rn(gts - (float)(F - fgi), gts + (float)(B - (1u - fgi)));
                    grad = __fsub_rn(J, Jp);
                }
                const float err = key_err(key[k]);
                acc += (double)err * (double)grad;
                if (p.need_grad) {
                    // d|fg - p|/dp = -sgn(fg - p): fg -> -1, bg -> +1, exactly 0 when the error is 0
                    const float gv = err > 0.f ? (fgi ? -grad : grad) * w : 0.f;
                    const u32 px = val[k] >> 1;
                    if (fgi) p.gown[px] = gv;
                    else {
                        const u32 ni = px / (u32)p.HW;
                        const u32 q = px - ni * (u32)p.HW;
                        p.gbg[((size_t)ni * p.C + c) * (size_t)p.HW + q] = gv;
                    }
                }
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(FULL_MASK, acc, o);
        if (lane == 0) s_red[warp] = acc;
        __syncthreads();
        if (tid == 0) {
            double tot = 0.0;
#pragma unroll
            for (int w2 = 0; w2 < SORT_WARPS; ++w2) tot += s_red[w2];
            atomicAdd(p.seg_loss + seg, tot);
        }
    }
}

// K5b: loss = mean over groups of (mean over kept classes)      reference: mean(), losses/LovaszSoftmax.py:102-120
__global__ void loss_finalize_kernel(LovaszParams p) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    float total = 0.f;
    for (int g = 0; g < p.groups; ++g) {
        float acc = 0.f;
        int n = 0;
        for (int c = 0; c < p.C; ++c) {
            const size_t seg = (size_t)g * p.C + c;
            if (!thr_active(p.seg_thr[seg])) continue;
            const float l = (float)p.seg_loss[seg];
            acc = n ? acc + l : l;
            ++n;
        }
        if (n > 1) acc = acc / (float)n;
        total = g ? total + acc : acc;
    }
    if (p.groups > 1) total = total / (float)p.groups;
    *p.loss_out = total;
}

// --------------------------------------------------------------------------------------------------------------
// K6: backward      reference: autograd of LovaszSoftmax.forward
// --------------------------------------------------------------------------------------------------------------
template <int CT, typename LT>
__global__ void __launch_bounds__(BWD_TPB) backward_kernel_v4(LovaszParams p, const float* __restrict__ go,
                                                              float* __restrict__ dlogits) {
    __shared__ float s_thr[B200SEG_MAX_CLASSES], s_logthr[B200SEG_MAX_CLASSES];
    const int tid = threadIdx.x;
    constexpr int TILE_PX = BWD_TPB * 4;
    const long long tpi = (p.HW + TILE_PX - 1) / TILE_PX;
    const long long ntiles = tpi * p.N;
    const long long t0 = ntiles * blockIdx.x / gridDim.x, t1 = ntiles * (blockIdx.x + 1) / gridDim.x;
    const float gsc = __ldg(go);
    int cur_g = -1;
    for (long long t = t0; t < t1; ++t) {
        const int n = (int)(t / tpi);
        const long long q0 = (t - (long long)n * tpi) * TILE_PX + tid * 4;
        const int g = p.per_image ? n : 0;
        if (g != cur_g) {
            __syncthreads();
            if (tid < CT) { s_thr[tid] = p.seg_thr[(size_t)g * CT + tid]; s_logthr[tid] = p.seg_logthr[(size_t)g * CT + tid]; }
            __syncthreads();
            cur_g = g;
        }
        if (q0 >= p.HW) continue;
        const size_t off = (size_t)n * CT * p.HW + q0;
        const float* lp = p.logits + off;
        const float* gb = p.gbg + off;
        float* dp = dlogits + off;
        const size_t px = (size_t)n * p.HW + q0;
        float z[CT][4];
#pragma unroll
        for (int c = 0; c < CT; ++c) {
            const float4 v = ld_stream4(lp + (size_t)c * p.HW);
            z[c][0] = v.x; z[c][1] = v.y; z[c][2] = v.z; z[c][3] = v.w;
        }
        int lab[4];
        load_labels4<LT>(p.labels, px, lab);
        const float4 mv = *(const float4*)(p.pix_m + px), sv = *(const float4*)(p.pix_s + px);
        const float4 gv = *(const float4*)(p.gown + px);
        const float m[4] = {mv.x, mv.y, mv.z, mv.w}, s[4] = {sv.x, sv.y, sv.z, sv.w}, go4[4] = {gv.x, gv.y, gv.z, gv.w};
        // sweep 1 (branch-free, unrolled): conservative candidate bits
        float theta[4];
        u32 pre[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) { theta[j] = pre_theta(m[j], s[j]); pre[j] = 0; }
#pragma unroll
        for (int c = 0; c < CT; ++c) {
            const float lt = s_logthr[c];
#pragma unroll
            for (int j = 0; j < 4; ++j) pre[j] |= (z[c][j] >= theta[j] + lt) ? (1u << c) : 0u;
        }
        // exact stage on the (few) flagged classes: same predicate as emit_kernel, logits re-read from L1/L2
        float dot[4], gl[4], nd[4];
        u32 fix[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const bool filt = p.has_filter && lab[j] == p.filter;
            const bool own = !filt && (unsigned)lab[j] < (unsigned)CT && thr_active(s_thr[lab[j] & 31]);
            const u32 ownbit = (unsigned)lab[j] < (unsigned)CT ? (1u << lab[j]) : 0u;
            gl[j] = own ? go4[j] : 0.f;
            float d = 0.f;
            u32 cm = 0;
            u32 mm = filt ? 0u : (pre[j] & ~ownbit);
            while (mm) {
                const int c = __ffs(mm) - 1;
                mm &= mm - 1;
                const float pr = sm_prob(__ldg(lp + (size_t)c * p.HW + j), m[j], s[j]);
                if (pr >= s_thr[c]) { cm |= 1u << c; d += gb[(size_t)c * p.HW + j] * pr; }
            }
            if (own) { d += gl[j] * sm_prob(__ldg(lp + (size_t)lab[j] * p.HW + j), m[j], s[j]); cm |= ownbit; }
            dot[j] = d; fix[j] = cm;
            nd[j] = filt ? 0.f : -gsc * d * __fdiv_rn(1.0f, s[j]);
        }
        // sweep 2 (branch-free, unrolled): every class gets -go * p_k * dot with the fast exponential ...
#pragma unroll
        for (int c = 0; c < CT; ++c) {
            float o[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) o[j] = nd[j] * __expf(z[c][j] - m[j]);
            st_stream4(dp + (size_t)c * p.HW, make_float4(o[0], o[1], o[2], o[3]));
        }
        // ... then the candidate classes are overwritten with the exact go * p_k * (g_k - dot)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            u32 mm = fix[j];
            while (mm) {
                const int c = __ffs(mm) - 1;
                mm &= mm - 1;
                const float pk = sm_prob(__ldg(lp + (size_t)c * p.HW + j), m[j], s[j]);
                const float gk = (c == lab[j]) ? gl[j] : gb[(size_t)c * p.HW + j];
                dp[(size_t)c * p.HW + j] = gsc * pk * (gk - dot[j]);
            }
        }
    }
}

// Pipelined backward (the fast path): lane l of a warp owns pixels [l*VEC, l*VEC+VEC) of the warp tile.
template <int CT, int VEC, int TPB, int STAGES, typename LT>
__global__ void __launch_bounds__(TPB, (VEC == 1 && TPB == 128) ? 8 : 1) backward_kernel_async(LovaszParams p, const float* __restrict__ go,
                                                              float* __restrict__ dlogits) {
    using W = WarpTile<CT, VEC>;
    constexpr int WT = W::WT, NW = TPB / 32;
    extern __shared__ __align__(16) unsigned char pipe_smem_raw[];
    float (*Zall)[STAGES][CT][WT] = reinterpret_cast<float (*)[STAGES][CT][WT]>(pipe_smem_raw);   // [NW][STAGES][CT][WT]
    __shared__ float s_thr[NW][B200SEG_MAX_CLASSES], s_logthr[NW][B200SEG_MAX_CLASSES];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float (*Z)[CT][WT] = Zall[warp];
    const int l0 = lane * VEC;
    const u32 wtpi = (u32)((p.HW + WT - 1) / WT);         // tile counts fit 32 bits: keep the index math off the 64-bit divider
    const u32 nwt = wtpi * (u32)p.N;
    const u32 gw = blockIdx.x * NW + warp, nwarps = gridDim.x * NW;
    const u32 t0 = (u32)((u64)nwt * gw / nwarps), t1 = (u32)((u64)nwt * (gw + 1) / nwarps);
    const float gsc = __ldg(go);

    // (image, tile-in-image) cursors advance by increments: no integer division per tile
    u32 cn = t0 / wtpi, cti = t0 - cn * wtpi;             // tile being consumed
    u32 pn = cn, pti = cti, pt = t0;                       // next tile to prefetch
    auto prefetch_next = [&](int stage) {
        if (pt < t1) W::prefetch(Z[stage], p.logits, (int)pn, (long long)pti * WT, p.HW, lane);
        cp_async_commit();
        ++pt;
        if (++pti == wtpi) { pti = 0; ++pn; }
    };
#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) prefetch_next(s);

    int cur_g = -1;
    u32 it = 0;
    for (u32 t = t0; t < t1; ++t, ++it) {
        const int stage = (int)(it % STAGES);
        __syncwarp();                                     // every lane is done reading the stage about to be refilled
        prefetch_next((int)((it + STAGES - 1) % STAGES));
        const int n = (int)cn;
        const long long q = (long long)cti * WT + l0;
        if (++cti == wtpi) { cti = 0; ++cn; }
        const int g = p.per_image ? n : 0;
        if (g != cur_g) {
            if (lane < CT) { s_thr[warp][lane] = p.seg_thr[(size_t)g * CT + lane]; s_logthr[warp][lane] = p.seg_logthr[(size_t)g * CT + lane]; }
            cur_g = g;
            __syncwarp();
        }
        const bool inb = q < p.HW;
        // per-pixel state straight from global memory (issued before the wait so it overlaps)
        float m[VEC], s[VEC], gl[VEC];
        int lab[VEC];
        const size_t px = (size_t)n * p.HW + q;
#pragma unroll
        for (int j = 0; j < VEC; ++j) { m[j] = 0.f; s[j] = 1.f; gl[j] = 0.f; lab[j] = -1; }
        if (inb) {
            if constexpr (VEC == 4) {
                const float4 mv = *(const float4*)(p.pix_m + px), sv = *(const float4*)(p.pix_s + px), gv = *(const float4*)(p.gown + px);
                m[0] = mv.x; m[1] = mv.y; m[2] = mv.z; m[3] = mv.w;
                s[0] = sv.x; s[1] = sv.y; s[2] = sv.z; s[3] = sv.w;
                gl[0] = gv.x; gl[1] = gv.y; gl[2] = gv.z; gl[3] = gv.w;
                load_labels4<LT>(p.labels, px, lab);
            } else {
#pragma unroll
                for (int j = 0; j < VEC; ++j) {
                    m[j] = p.pix_m[px + j]; s[j] = p.pix_s[px + j]; gl[j] = p.gown[px + j];
                    lab[j] = load_label<LT>(p.labels, px + j);
                }
            }
        }
        cp_async_wait<STAGES - 1>();                      // this lane's copies for tile t have landed ...
        __syncwarp();                                     // ... and so have the other lanes'
        if (!inb) continue;
        const float (*T)[WT] = Z[stage];
        const size_t off = (size_t)n * CT * p.HW + q;
        const float* gb = p.gbg + off;
        float* dp = dlogits + off;
        const float* thr = s_thr[warp];
        float theta[VEC];
        u32 pre[VEC];
#pragma unroll
        for (int j = 0; j < VEC; ++j) { theta[j] = pre_theta(m[j], s[j]); pre[j] = 0; }
        // sweep 1 (branch-free): conservative candidate bits
#pragma unroll
        for (int c = 0; c < CT; ++c) {
#pragma unroll
            for (int j = 0; j < VEC; ++j) pre[j] |= (T[c][l0 + j] >= theta[j] + s_logthr[warp][c]) ? (1u << c) : 0u;
        }
        // exact stage on the flagged classes: same predicate as the emission kernel
        float dot[VEC], nd[VEC], ownv[VEC];
        u32 fix[VEC];
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
            const bool filt = p.has_filter && lab[j] == p.filter;
            const bool own = !filt && (unsigned)lab[j] < (unsigned)CT && thr_active(thr[lab[j] & 31]);
            const u32 ownbit = (unsigned)lab[j] < (unsigned)CT ? (1u << lab[j]) : 0u;
            float d = 0.f;
            u32 cm = 0;
            u32 mm = filt ? 0u : (pre[j] & ~ownbit);
            while (mm) {
                const int c = __ffs(mm) - 1;
                mm &= mm - 1;
                const float pr = sm_prob(T[c][l0 + j], m[j], s[j]);
                if (pr >= thr[c]) { cm |= 1u << c; d += gb[(size_t)c * p.HW + j] * pr; }
            }
            float pown = 0.f;
            if (own) { pown = sm_prob(T[lab[j]][l0 + j], m[j], s[j]); d += gl[j] * pown; }
            else lab[j] = -1;                              // no class of this pixel takes the own-class path below
            dot[j] = d; fix[j] = cm;
            nd[j] = filt ? 0.f : -gsc * d * __fdiv_rn(1.0f, s[j]);
            ownv[j] = gsc * pown * (gl[j] - d);            // exact value of the own class
        }
        // sweep 2 (branch-free): -go * p_k * dot with the fast exponential, the own class takes its exact value ...
#pragma unroll
        for (int c = 0; c < CT; ++c) {
            float o[VEC];
#pragma unroll
            for (int j = 0; j < VEC; ++j) {
                const float fast = nd[j] * __expf(T[c][l0 + j] - m[j]);
                o[j] = (c == lab[j]) ? ownv[j] : fast;
            }
            if constexpr (VEC == 4) st_stream4(dp + (size_t)c * p.HW, make_float4(o[0], o[1], o[2], o[3]));
            else if constexpr (VEC == 2) st_stream2(dp + (size_t)c * p.HW, make_float2(o[0], o[1]));
            else dp[(size_t)c * p.HW] = o[0];
        }
        // ... then the (few) background-candidate classes are overwritten with the exact go * p_k * (g_k - dot)
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
            u32 mm = fix[j];
            while (mm) {
                const int c = __ffs(mm) - 1;
                mm &= mm - 1;
                const float pk = sm_prob(T[c][l0 + j], m[j], s[j]);
                dp[(size_t)c * p.HW + j] = gsc * pk * (gb[(size_t)c * p.HW + j] - dot[j]);
            }
        }
    }
    cp_async_wait<0>();
}

template <typename LT>
__global__ void __launch_bounds__(BWD_TPB) backward_kernel_generic(LovaszParams p, const float* __restrict__ go,
                                                                   float* __restrict__ dlogits) {
    __shared__ float s_thr[B200SEG_MAX_CLASSES], s_logthr[B200SEG_MAX_CLASSES];
    const int tid = threadIdx.x;
    const int C = p.C;
    constexpr int TILE_PX = BWD_TPB;
    const long long tpi = (p.HW + TILE_PX - 1) / TILE_PX;
    const long long ntiles = tpi * p.N;
    const long long t0 = ntiles * blockIdx.x / gridDim.x, t1 = ntiles * (blockIdx.x + 1) / gridDim.x;
    const float gsc = __ldg(go);
    int cur_g = -1;
    for (long long t = t0; t < t1; ++t) {
        const int n = (int)(t / tpi);
        const long long q = (t - (long long)n * tpi) * TILE_PX + tid;
        const int g = p.per_image ? n : 0;
        if (g != cur_g) {
            __syncthreads();
            if (tid < C) { s_thr[tid] = p.seg_thr[(size_t)g * C + tid]; s_logthr[tid] = p.seg_logthr[(size_t)g * C + tid]; }
            __syncthreads();
            cur_g = g;
        }
        if (q >= p.HW) continue;
        const size_t off = (size_t)n * C * p.HW + q;
        const float* lp = p.logits + off;
        const float* gb = p.gbg + off;
        float* dp = dlogits + off;
        const size_t px = (size_t)n * p.HW + q;
        const int lab = load_label<LT>(p.labels, px);
        const float m = p.pix_m[px], s = p.pix_s[px];
        const bool filt = p.has_filter && lab == p.filter;
        const bool own = !filt && (unsigned)lab < (unsigned)C && thr_active(s_thr[lab & 31]);
        const float gl = own ? p.gown[px] : 0.f;
        const float theta = pre_theta(m, s);
        float d = 0.f;
        u32 cm = 0;
        for (int c = 0; c < C; ++c) {
            const float v = __ldg(lp + (size_t)c * p.HW);
            if (c == lab) { if (own) d += gl * sm_prob(v, m, s); }
            else if (!filt && v >= theta + s_logthr[c]) {
                const float pr = sm_prob(v, m, s);
                if (pr >= s_thr[c]) { cm |= 1u << c; d += gb[(size_t)c * p.HW] * pr; }
            }
        }
        for (int c = 0; c < C; ++c) {
            const float v = __ldg(lp + (size_t)c * p.HW);
            float gk = (c == lab) ? gl : 0.f;
            if ((cm >> c) & 1u) gk = gb[(size_t)c * p.HW];
            const float pk = sm_prob(v, m, s);
            dp[(size_t)c * p.HW] = filt ? 0.f : gsc * pk * (gk - d);
        }
    }
}

// --------------------------------------------------------------------------------------------------------------
// host side
// --------------------------------------------------------------------------------------------------------------
static int check_shape(int32_t n, int32_t c, int64_t hw) {
    if (n < 0 || hw < 0 || c < 1 || c > B200SEG_MAX_CLASSES) {
        b200seg_set_error("invalid shape: n_images=%d n_classes=%d plane=%lld (need 1 <= n_classes <= %d)", n, c,
                          (long long)hw, B200SEG_MAX_CLASSES);
        return B200SEG_E_INVALID;
    }
    const long double P = (long double)n * (long double)hw;
    if (P >= (long double)(1u << 30) || P * c >= (long double)(1ull << 31)) {
        b200seg_set_error("shape too large: n_images*plane must be < 2^30 and n_images*plane*n_classes < 2^31");
        return B200SEG_E_INVALID;
    }
    return 0;
}

static bool fill_params(LovaszParams& p, const LovaszLayout& L, char* ws, const float* logits, const void* labels,
                        int32_t n, int32_t c, int64_t hw, int32_t per_image, int64_t filter_label,
                        int32_t keep_absent, uint32_t class_mask) {
    p.logits = logits; p.labels = labels;
    p.N = n; p.C = c; p.HW = hw; p.P = (long long)n * hw;
    p.per_image = per_image ? 1 : 0;
    p.groups = per_image ? n : 1;
    p.n_seg = p.groups * c;
    p.cap = per_image ? hw : p.P;
    p.has_filter = (filter_label != B200SEG_NO_LABEL && filter_label >= INT_MIN && filter_label <= INT_MAX) ? 1 : 0;
    p.filter = p.has_filter ? (int)filter_label : 0;
    p.keep_absent = keep_absent ? 1 : 0;
    p.class_mask = c == 32 ? class_mask : (class_mask & ((1u << c) - 1u));
    p.ctrl = (u32*)(ws + L.ctrl);
    p.seg_fg = (u32*)(ws + L.seg_fg); p.seg_maxkey = (u32*)(ws + L.seg_maxkey); p.seg_maxp = (u32*)(ws + L.seg_maxp);
    p.seg_count = (u32*)(ws + L.seg_count); p.grp_valid = (u32*)(ws + L.grp_valid); p.seg_bits = (u32*)(ws + L.seg_bits);
    p.seg_loss = (double*)(ws + L.seg_loss);
    p.seg_thr = (float*)(ws + L.seg_thr); p.seg_logthr = (float*)(ws + L.seg_logthr); p.seg_w = (float*)(ws + L.seg_w);
    p.pix_m = (float*)(ws + L.pix_m); p.pix_s = (float*)(ws + L.pix_s); p.gown = (float*)(ws + L.gown);
    p.run_cnt = (u32*)(ws + L.run_cnt); p.run_prefix = (u32*)(ws + L.run_prefix);
    p.n_runs = 0; p.tiles_per_chunk = 0; p.run_stride = 0; p.src_cap = 0;
    p.keysA = (u32*)(ws + L.keysA); p.valsA = (u32*)(ws + L.valsA);
    p.keysB = (u32*)(ws + L.keysB); p.valsB = (u32*)(ws + L.valsB);
    p.gbg = (float*)(ws + L.keysA);      // free again once the sort result sits in buffer B
    p.cm = nullptr; p.has_drop = 0; p.drop = 0;
    p.status = (int*)(p.ctrl + CTRL_STATUS);
    p.loss_out = nullptr; p.need_grad = 1; p.dbg = 0;
    return true;
}

static bool aligned16(const void* ptr) { return ((uintptr_t)ptr & 15) == 0; }
static bool vec4_ok(const float* logits, const void* labels, int label_dtype, int64_t hw) {
    if (hw % 4 != 0 || !aligned16(logits)) return false;
    if (label_dtype == B200SEG_LABEL_U8) return ((uintptr_t)labels & 3) == 0;
    return aligned16(labels);
}

#define DISPATCH_LABEL(dtype, ...)                                               \
    switch (dtype) {                                                             \
        case B200SEG_LABEL_U8: { typedef uint8_t LT; __VA_ARGS__; } break;        \
        case B200SEG_LABEL_I32: { typedef int32_t LT; __VA_ARGS__; } break;       \
        case B200SEG_LABEL_I64: { typedef int64_t LT; __VA_ARGS__; } break;       \
        default: b200seg_set_error("unknown label dtype %d", dtype); return B200SEG_E_INVALID; \
    }

extern "C" int b200seg_lovasz_workspace_bytes(int32_t n, int32_t c, int64_t hw, int32_t per_image, size_t* bytes) {
    if (!bytes) { b200seg_set_error("bytes is NULL"); return B200SEG_E_INVALID; }
    if (int rc = check_shape(n, c, hw)) return rc;
    *bytes = lovasz_layout(n, c, hw, per_image).total;
    return 0;
}

extern "C" int b200seg_lovasz_forward(const float* logits, const void* labels, int32_t label_dtype, int32_t n,
                                      int32_t c, int64_t hw, int32_t per_image, int64_t filter_label,
                                      int32_t keep_absent, uint32_t class_mask, int32_t need_grad, void* workspace,
                                      size_t workspace_bytes, float* loss_out, int64_t* cm, int64_t cm_drop_label,
                                      int32_t* status, void* stream) {
    if (int rc = check_shape(n, c, hw)) return rc;
    if (!loss_out) { b200seg_set_error("loss_out is NULL"); return B200SEG_E_INVALID; }
    if ((long long)n * hw == 0) {      // empty batch: nothing to read, loss 0
        CUDA_TRY(cudaMemsetAsync(loss_out, 0, sizeof(float), (cudaStream_t)stream));
        return 0;
    }
    if (!logits || !labels || !workspace || (cm && !status)) {
        b200seg_set_error("null pointer argument");
        return B200SEG_E_INVALID;
    }
    const LovaszLayout L = lovasz_layout(n, c, hw, per_image);
    if (workspace_bytes < L.total || ((uintptr_t)workspace & 255)) {
        b200seg_set_error("workspace too small or not 256-byte aligned: have %zu, need %zu", workspace_bytes, L.total);
        return B200SEG_E_WORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    char* ws = (char*)workspace;
    LovaszParams p;
    fill_params(p, L, ws, logits, labels, n, c, hw, per_image, filter_label, keep_absent, class_mask);
    p.loss_out = loss_out;
    p.need_grad = need_grad ? 1 : 0;
    p.cm = (unsigned long long*)cm;
    p.has_drop = (cm_drop_label != B200SEG_NO_LABEL && cm_drop_label >= INT_MIN && cm_drop_label <= INT_MAX) ? 1 : 0;
    p.drop = p.has_drop ? (int)cm_drop_label : 0;
    if (status) p.status = status;

    CUDA_TRY(cudaMemsetAsync(ws + L.ctrl, 0, L.zero_end - L.ctrl, st));
    const int sms = b200seg_sm_count();
    const bool v4 = vec4_ok(logits, labels, label_dtype, hw);
    b200seg_stage(0, st);

    // K1
    if (v4 && (c == 8 || c == 17 || c == 25)) {
        const int grid = sms * 3;
        DISPATCH_LABEL(label_dtype, {
            if (c == 8) stats_kernel_vec<8, 4, STATS_TPB, LT><<<grid, STATS_TPB, 0, st>>>(p);
            else if (c == 17) stats_kernel_vec<17, 4, STATS_TPB, LT><<<grid, STATS_TPB, 0, st>>>(p);
            else stats_kernel_vec<25, 4, STATS_TPB, LT><<<grid, STATS_TPB, 0, st>>>(p);
        });
    } else {
        DISPATCH_LABEL(label_dtype, stats_kernel_generic<LT><<<sms * 8, STATS_TPB, 0, st>>>(p));
    }
    LAUNCH_CHECK("stats_kernel");
    b200seg_stage(1, st);
    if (p.keep_absent) {
        DISPATCH_LABEL(label_dtype, absent_max_kernel<LT><<<dim3(p.n_seg, 32), 256, 0, st>>>(p));
        LAUNCH_CHECK("absent_max_kernel");
    }
    finalize_stats_kernel<<<(p.groups + 127) / 128, 128, 0, st>>>(p);
    LAUNCH_CHECK("finalize_stats_kernel");
    b200seg_stage(2, st);

    // K2
    {
        const bool pipe_ok = v4 && (c == 8 || c == 17 || c == 25);
        const EmitGeom G = emit_geom(n, hw, per_image, pipe_ok ? EMIT_WARP_TILE : (long long)EMIT_TPB * (v4 ? 4 : 1));
        p.n_runs = (int)G.n_runs; p.tiles_per_chunk = (int)G.tpc; p.run_stride = G.run_stride; p.src_cap = G.src_cap;
        const long long chunks = (long long)p.groups * G.n_runs;
        if (pipe_ok) {
#define LAUNCH_EMIT_ASYNC(CC)                                                                                   \
    {                                                                                                           \
        constexpr int ET = 128, ES = 2;                                                                         \
        const size_t smem = sizeof(float) * (size_t)ES * CC * ET;                                               \
        CUDA_TRY(cudaFuncSetAttribute(emit_kernel_async<CC, ET, ES, LT>,                                        \
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));                 \
        int per_sm = (int)((224 * 1024) / (smem + 2048));                                                       \
        if (per_sm * ET > 2048) per_sm = 2048 / ET;                                                             \
        long long grid = (long long)sms * per_sm;                                                               \
        if (grid * (ET / 32) > chunks) grid = (chunks + ET / 32 - 1) / (ET / 32);                               \
        emit_kernel_async<CC, ET, ES, LT><<<(int)grid, ET, smem, st>>>(p);                                      \
    }
            DISPATCH_LABEL(label_dtype, {
                if (c == 8) LAUNCH_EMIT_ASYNC(8)
                else if (c == 17) LAUNCH_EMIT_ASYNC(17)
                else LAUNCH_EMIT_ASYNC(25)
            });
#undef LAUNCH_EMIT_ASYNC
        } else {
            const int grid = (int)(chunks < (long long)sms * 8 ? chunks : (long long)sms * 8);
            if (v4) { DISPATCH_LABEL(label_dtype, emit_kernel<4, LT><<<grid, EMIT_TPB, 0, st>>>(p)); }
            else { DISPATCH_LABEL(label_dtype, emit_kernel<1, LT><<<grid, EMIT_TPB, 0, st>>>(p)); }
        }
        LAUNCH_CHECK("emit_kernel");
        run_scan_kernel<<<(p.n_seg + 7) / 8, 256, 0, st>>>(p);
        LAUNCH_CHECK("run_scan_kernel");
    }
    b200seg_stage(3, st);

    // sort: holey source in A (gathered through the run prefix), ping-pong B -> A -> B
    SortArgs a;
    char* ss = ws + L.sort_scratch;
    a.keys[0] = p.keysA; a.vals[0] = p.valsA; a.keys[1] = p.keysB; a.vals[1] = p.valsB;
    a.seg_count = p.seg_count; a.seg_bits = p.seg_bits; a.n_seg = p.n_seg; a.cap = p.cap;
    a.src_keys = p.keysA; a.src_vals = p.valsA; a.run_prefix = p.run_prefix; a.n_runs = p.n_runs;
    a.run_stride = p.run_stride; a.src_cap = p.src_cap;
    a.tile_start = (u32*)(ss + L.sort.tile_start); a.tilehist = (u32*)(ss + L.sort.tilehist);
    a.tile_desc = (uint4*)(ss + L.sort.tile_desc); a.tile_runs = (uint2*)(ss + L.sort.tile_runs);
    a.seg_done = (u32*)(ss + L.sort.seg_done);
    a.bin_base = (u32*)(ss + L.sort.bin_base); a.tile_fg = (u32*)(ss + L.sort.tile_fg);
    a.status = p.status;
    if (int rc = sort_enqueue(a, L.sort, st)) return rc;

    // K5
    jaccard_kernel<<<sms * 4, JAC_TPB, 0, st>>>(p, a);
    LAUNCH_CHECK("jaccard_kernel");
    loss_finalize_kernel<<<1, 32, 0, st>>>(p);
    LAUNCH_CHECK("loss_finalize_kernel");
    b200seg_stage(8, st);
    return 0;
}

extern "C" int b200seg_lovasz_backward(const float* logits, const void* labels, int32_t label_dtype, int32_t n,
                                       int32_t c, int64_t hw, int32_t per_image, int64_t filter_label,
                                       int32_t keep_absent, uint32_t class_mask, const void* workspace,
                                       size_t workspace_bytes, const float* grad_out, float* dlogits, void* stream) {
    if (int rc = check_shape(n, c, hw)) return rc;
    if ((long long)n * hw == 0) return 0;
    if (!logits || !labels || !workspace || !grad_out || !dlogits) {
        b200seg_set_error("null pointer argument");
        return B200SEG_E_INVALID;
    }
    const LovaszLayout L = lovasz_layout(n, c, hw, per_image);
    if (workspace_bytes < L.total || ((uintptr_t)workspace & 255)) {
        b200seg_set_error("workspace too small or not 256-byte aligned: have %zu, need %zu", workspace_bytes, L.total);
        return B200SEG_E_WORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    LovaszParams p;
    fill_params(p, L, (char*)const_cast<void*>(workspace), logits, labels, n, c, hw, per_image, filter_label,
                keep_absent, class_mask);
    if (p.P == 0) return 0;
    const int sms = b200seg_sm_count();
    const bool v4 = vec4_ok(logits, labels, label_dtype, hw) && aligned16(dlogits);
    const bool pipe_ok = v4 && (label_dtype != B200SEG_LABEL_U8 || hw % 16 == 0) && aligned16(labels) &&
                         (c == 8 || c == 17 || c == 25);
    b200seg_stage(9, st);
    if (pipe_ok) {
        const char* e = getenv("B200SEG_BWD_VARIANT");
        const int variant = e ? atoi(e) : 0;
#define LAUNCH_BWD_ASYNC(CC, VV, TT, SS)                                                                         \
    {                                                                                                            \
        const size_t smem = sizeof(float) * (size_t)SS * CC * TT * VV;                                           \
        CUDA_TRY(cudaFuncSetAttribute(backward_kernel_async<CC, VV, TT, SS, LT>,                                 \
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));                  \
        int per_sm = (int)((224 * 1024) / (smem + 2048));                                                        \
        if (per_sm < 1) per_sm = 1;                                                                              \
        if (per_sm * TT > 2048) per_sm = 2048 / TT;                                                              \
        backward_kernel_async<CC, VV, TT, SS, LT><<<sms * per_sm, TT, smem, st>>>(p, grad_out, dlogits);         \
    }
#define LAUNCH_BWD_C(VV, TT, SS)                                       \
    {                                                                  \
        if (c == 8) LAUNCH_BWD_ASYNC(8, VV, TT, SS)                    \
        else if (c == 17) LAUNCH_BWD_ASYNC(17, VV, TT, SS)             \
        else LAUNCH_BWD_ASYNC(25, VV, TT, SS)                          \
    }
        DISPATCH_LABEL(label_dtype, {
            switch (variant) {
                case 1: LAUNCH_BWD_C(1, 128, 4) break;
                case 2: LAUNCH_BWD_C(2, 128, 2) break;
                case 3: LAUNCH_BWD_C(2, 128, 3) break;
                case 4: LAUNCH_BWD_C(1, 256, 3) break;
                case 5: LAUNCH_BWD_C(1, 128, 2) break;
                case 6: LAUNCH_BWD_C(4, 128, 2) break;
                default: LAUNCH_BWD_C(1, 128, 2) break;
            }
        });
#undef LAUNCH_BWD_C
#undef LAUNCH_BWD_ASYNC
    } else if (v4 && (c == 8 || c == 17 || c == 25)) {
        const int grid = sms * 3 * 4;
        DISPATCH_LABEL(label_dtype, {
            if (c == 8) backward_kernel_v4<8, LT><<<grid, BWD_TPB, 0, st>>>(p, grad_out, dlogits);
            else if (c == 17) backward_kernel_v4<17, LT><<<grid, BWD_TPB, 0, st>>>(p, grad_out, dlogits);
            else backward_kernel_v4<25, LT><<<grid, BWD_TPB, 0, st>>>(p, grad_out, dlogits);
        });
    } else {
        DISPATCH_LABEL(label_dtype, backward_kernel_generic<LT><<<sms * 16, BWD_TPB, 0, st>>>(p, grad_out, dlogits));
    }
    LAUNCH_CHECK("backward_kernel");
    b200seg_stage(10, st);
    return 0;
}

// ---- test hook: the segmented sort on its own -------------------------------------------------------------------
extern "C" int b200seg_sort_scratch_bytes(int32_t n_segments, int64_t capacity, size_t* bytes) {
    if (!bytes || n_segments < 1 || capacity < 0 || (long double)n_segments * capacity >= (long double)(1ull << 31)) {
        b200seg_set_error("invalid sort shape");
        return B200SEG_E_INVALID;
    }
    *bytes = sort_scratch_layout(n_segments, (long long)n_segments * capacity).total;
    return 0;
}

extern "C" int b200seg_sort_segments(uint32_t* keys_in, uint32_t* vals_in, uint32_t* keys_out, uint32_t* vals_out,
                                     const uint32_t* counts, const uint32_t* key_bits, int32_t n_segments,
                                     int64_t capacity, void* scratch, size_t scratch_bytes, int32_t* status,
                                     void* stream) {
    size_t need = 0;
    if (int rc = b200seg_sort_scratch_bytes(n_segments, capacity, &need)) return rc;
    if (!keys_in || !vals_in || !keys_out || !vals_out || !counts || !key_bits || !scratch || !status) {
        b200seg_set_error("null pointer argument");
        return B200SEG_E_INVALID;
    }
    if (scratch_bytes < need || ((uintptr_t)scratch & 255)) {
        b200seg_set_error("sort scratch too small or misaligned: have %zu, need %zu", scratch_bytes, need);
        return B200SEG_E_WORKSPACE;
    }
    const SortScratch L = sort_scratch_layout(n_segments, (long long)n_segments * capacity);
    char* ss = (char*)scratch;
    SortArgs a;
    a.keys[0] = keys_in; a.vals[0] = vals_in; a.keys[1] = keys_out; a.vals[1] = vals_out;
    a.seg_count = counts; a.seg_bits = key_bits; a.n_seg = n_segments; a.cap = capacity;
    a.src_keys = nullptr; a.src_vals = nullptr; a.run_prefix = nullptr; a.n_runs = 0; a.run_stride = 0; a.src_cap = 0;
    a.tile_start = (u32*)(ss + L.tile_start); a.tilehist = (u32*)(ss + L.tilehist);
    a.tile_desc = (uint4*)(ss + L.tile_desc); a.tile_runs = (uint2*)(ss + L.tile_runs);
    a.seg_done = (u32*)(ss + L.seg_done);
    a.bin_base = (u32*)(ss + L.bin_base); a.tile_fg = (u32*)(ss + L.tile_fg);
    a.status = status;
    return sort_enqueue(a, L, (cudaStream_t)stream);
}
