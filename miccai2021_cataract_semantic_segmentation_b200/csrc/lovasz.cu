// Lovasz-Softmax forward + backward for sm_100a.  Replaces losses/LovaszSoftmax.py:19-95 of the reference.
//
// Pipeline (all on one stream, no host sync, no allocation):
//   K1 stats      read logits once: per-pixel softmax max/sum (kept, 8 B/px), compact label, per-segment foreground count
//                 and max key (= min foreground error -> sort threshold), a 20-byte candidate record per pixel (own-class key,
//                 the two most probable other classes, a bound on the rest), optional fused argmax + confusion matrix and
//                 cross-entropy sum
//   K1c absent    (keep_absent only) max_i p_c(i) for considered classes without foreground
//   K1b finalize  thresholds, key widths, class weights 1/n_present(/n_images), class order; chooses the emission path
//   K2 emit       every (pixel, class) with error >= threshold becomes a candidate (key = 0x3F800000 - bits(error) | fg << 31,
//                 value = pixel<<1 | fg); segments are written compactly, in no particular order (tiles reserve their slices
//                 with atomics); from the records (CTA kernel, no logits) or, for confident logits, from a second pass over
//                 the logits
//   K3h/K4h       hybrid sort (hybrid.cuh + hyb_local_kernel below): bucket histogram, ONE partition pass by the top key bits,
//                 then a local kernel that ranks whole buckets in shared memory in the canonical (key, pixel) order and goes
//                 straight to the Jaccard gradient, loss partials and per-candidate g; the CTA that finishes last takes the
//                 mean over present classes (sequential fp32, class order) and over images
//   K4f fallback  segments whose buckets did not fit shared memory: stable LSD radix sort (sort.cuh) + Jaccard kernel, fused
//                 into one cooperative launch; exits at once otherwise.  (sort_path = 1 runs that path for everything.)
//   K6 backward   re-read logits: dz_k = go * p_k (g_k - sum_j g_j p_j) (+ the cross-entropy gradient), sparse g per pixel
//
// Exactness notes: candidates are a superset of every element with non-zero Jaccard gradient (SURVEY.md §7.3,
// zero-tail), so pruning changes nothing; p_c is computed by the same fp32 instruction sequence in every pass.
#include "b200seg.h"
#include "common.cuh"
#include "sort.cuh"
#include "hybrid.cuh"
#include "pipe.cuh"
#include "upsample.cuh"
#include <cstdlib>
#include <cuda.h>

#define STATS_TPB 128
#define EMIT_TPB 256
#define BWD_TPB 128
#define JAC_TPB SORT_TPB
// record-driven CTA emission kernel: threads per CTA (4 pixels each), resident CTAs per SM, pixels / mask words per tile
#ifndef ECTA_TPB
#define ECTA_TPB 512
#endif
#define ECTA_MINB (1024 / ECTA_TPB)      // resident CTAs per SM at 64 registers per thread
#define ECTA_TILE (ECTA_TPB * 4)
#define ECTA_WORDS (ECTA_TILE / 32)

enum { CTRL_STATUS = 8, CTRL_FLAGS = 16, CTRL_SLOW = 24, CTRL_SAMPLES = 25, CTRL_TICKET = 26, CTRL_JTICKET = 27, CTRL_PTICKET = 28,
       CTRL_OVF_ANY = 29, CTRL_HTICKET = 30 /* 2 words */, CTRL_GEO = 32,
       CTRL_CE_SUM = 40 /* double */, CTRL_CE_CNT = 42, CTRL_CE_INV_N = 43 };
enum { EMIT_PATH_STREAM = 0, EMIT_PATH_RECORDS = 1 };

struct LovaszParams {
    const float* logits;                // full-resolution logits, or nullptr when `up` names low-resolution ones
    UpSrc up;                           // fused bilinear upsampling (lovasz_up.cuh); up.lo == nullptr: not used
    const void* labels;
    int N, C;
    long long HW, P, cap;
    u32 inv_hw;                         // floor(2^32 / HW) (HW >= 2), for the pixel -> (image, offset) split
    int per_image, has_filter, filter, keep_absent, need_grad, dbg, interleave;
    u32 class_mask;
    int groups, n_seg;
    // workspace
    u32* ctrl;
    u32 *seg_fg, *seg_maxkey, *seg_maxp, *seg_count, *grp_valid, *seg_bits;
    double* seg_loss;
    float *seg_thr, *seg_logthr, *seg_w;
    float* grp_tmin;                    // [groups] smallest threshold among the group's summed classes
    unsigned char* seg_order;           // [groups][C] classes of the group by ascending threshold
    // fused cross-entropy term (nn.CrossEntropyLoss(ignore_index), mean over the non-ignored pixels; LossWrapper.py:17-24)
    int ce_enabled, has_ce_ignore, ce_ignore;
    double* ce_sum;                     // sum over valid pixels of -log p_label
    u32* ce_cnt;                        // valid pixels
    float* ce_inv_n;                    // 1 / ce_cnt for the backward pass
    float* ce_out;                      // caller's scalar
    int have_records;                   // stats_kernel_async ran: rec16 / rec4 are valid
    int emit_force;                     // 0 = decide on the device, 1 = record path, 2 = streaming path (B200SEG_EMIT_PATH)
    float *pix_m, *pix_s, *gown, *gbg;
    unsigned char* lab8;                // per pixel: class 0..C-1, LAB8_NONE (not a class), LAB8_FILTERED
    u32* cmask;                         // per pixel: classes emitted as BACKGROUND candidates (written by K2)
    uint4* rec16;                       // per pixel candidate record (stats_kernel_async): {key_fg, p1, p2, guard p3}
    u32* rec4;                          //   label8 | class1 << 8 | class2 << 16
    u32* flags;                         // [0] emission path chosen by K1b (EMIT_PATH_*)
    EmitGeomDev geo_stream, geo_rec;    // chunk geometry of the streaming / record-driven emission kernels
    EmitGeomDev* geo;                   // the one in force (device memory; written by K1b / K1d, read by K2 and the sort)
    u32 *keysA, *valsA, *keysB, *valsB;
    // fused confusion matrix
    unsigned long long* cm;
    int has_drop, drop;
    int* status;
    float* loss_out;
};

struct LovaszLayout {
    size_t ctrl, seg_fg, seg_maxkey, seg_maxp, seg_count, grp_valid, seg_loss, seg_loss_h, seg_ovf, zero_end;
    size_t seg_thr, seg_logthr, seg_w, seg_bits, grp_tmin, seg_order;
    size_t pix_m, pix_s, gown, lab8, cmask, rec16, rec4, keysA, valsA, keysB, valsB, gbg, sort_scratch, total;
    size_t hyb_hist, hyb_fgpre, hyb_done;
    SortScratch sort;
};

// Emission work split for one tile size: tiles never straddle images; a chunk is `tpc` consecutive tiles of one group.
struct EmitGeom { long long tile_px, tpi, tpg, tpc, n_runs; };
#define EMIT_WARP_TILE 32            // pixels per warp tile of the pipelined emission kernel
static EmitGeom emit_geom(int N, long long HW, int per_image, long long tile_px) {
    EmitGeom G;
    G.tile_px = tile_px;
    G.tpi = (HW + G.tile_px - 1) / G.tile_px;
    G.tpg = per_image ? G.tpi : G.tpi * N;
    const long long total_tiles = G.tpi * N;
    // chunks over the whole batch: the pipelined kernel (32-pixel tiles) runs one chunk per resident warp
    // and the record-driven CTA kernel (ECTA_TILE-pixel tiles) one chunk per resident CTA
    const long long target_chunks = (long long)b200seg_sm_count() * (tile_px <= 64 ? 32 : (tile_px == ECTA_TILE ? ECTA_MINB : 8));
    G.tpc = (total_tiles + target_chunks - 1) / target_chunks;
    if (G.tpc < 1) G.tpc = 1;
    G.n_runs = (G.tpg + G.tpc - 1) / G.tpc;
    if (G.n_runs < 1) G.n_runs = 1;
    return G;
}

static LovaszLayout lovasz_layout(int N, int C, long long HW, int per_image) {
    LovaszLayout L;
    const long long P = (long long)N * HW;
    const int groups = per_image ? N : 1;
    const size_t S = (size_t)groups * C;
    const size_t CP = (size_t)C * P;
    size_t o = 0;
    L.ctrl = o;       o = align_up(o + 256, 256);
    L.seg_fg = o;     o = align_up(o + 4 * S, 256);
    L.seg_maxkey = o; o = align_up(o + 4 * S, 256);
    L.seg_maxp = o;   o = align_up(o + 4 * S, 256);
    L.seg_count = o;  o = align_up(o + 4 * S, 256);
    L.grp_valid = o;  o = align_up(o + 4 * (size_t)groups, 256);
    L.seg_loss = o;   o = align_up(o + 8 * S, 256);
    L.seg_loss_h = o; o = align_up(o + 8 * S, 256);
    L.seg_ovf = o;    o = align_up(o + 4 * S, 256);
    L.zero_end = o;
    L.seg_thr = o;    o = align_up(o + 4 * S, 256);
    L.seg_logthr = o; o = align_up(o + 4 * S, 256);
    L.seg_w = o;      o = align_up(o + 4 * S, 256);
    L.seg_bits = o;   o = align_up(o + 4 * S, 256);
    L.grp_tmin = o;   o = align_up(o + 4 * (size_t)groups, 256);
    L.seg_order = o;  o = align_up(o + S, 256);
    L.pix_m = o;      o = align_up(o + 4 * (size_t)P, 256);
    L.pix_s = o;      o = align_up(o + 4 * (size_t)P, 256);
    L.gown = o;       o = align_up(o + 4 * (size_t)P, 256);
    L.lab8 = o;       o = align_up(o + (size_t)P + 16, 256);
    L.cmask = o;      o = align_up(o + 4 * (size_t)P, 256);
    L.rec16 = o;      o = align_up(o + 16 * (size_t)P, 256);
    L.rec4 = o;       o = align_up(o + 4 * (size_t)P, 256);
    L.keysA = o;      o = align_up(o + 4 * CP, 256);
    L.valsA = o;      o = align_up(o + 4 * CP, 256);
    L.keysB = o;      o = align_up(o + 4 * CP, 256);
    L.valsB = o;      o = align_up(o + 4 * CP, 256);
    // background-candidate gradients, indexed like the logits.  (Not aliased onto the dead sort buffer A any more: the
    // hybrid path's fallback re-reads the emission output in A after gradients have been written.)
    L.gbg = o;        o = align_up(o + 4 * CP, 256);
    L.sort = sort_scratch_layout((int)S, (long long)CP);
    L.sort_scratch = o; o = align_up(o + L.sort.total, 256);
    L.hyb_hist = o;   o = align_up(o + 4 * S * (size_t)HYB_MAX_BINS, 256);
    L.hyb_done = o;   o = align_up(o + 4 * S, 256);
    L.hyb_fgpre = o;  o = align_up(o + 4 * S * (size_t)HYB_MAX_BINS, 256);
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
#define LAB8_NONE 255u               // label outside [0, C): the pixel is background for every class
#define LAB8_FILTERED 254u           // label == classes_to_ignore: the pixel is removed from the loss
__device__ __forceinline__ u32 lab8_encode(int lab, int C, int has_filter, int filter) {
    if (has_filter && lab == filter) return LAB8_FILTERED;
    return (unsigned)lab < (unsigned)C ? (u32)lab : LAB8_NONE;
}
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
            u32 l4 = 0;
#pragma unroll
            for (int j = 0; j < 4; ++j) l4 |= lab8_encode(lab[j], CT, p.has_filter, p.filter) << (8 * j);
            *(u32*)(p.lab8 + px) = l4;
        } else {
            *(float2*)(p.pix_m + px) = make_float2(mo[0], mo[1]);
            *(float2*)(p.pix_s + px) = make_float2(so[0], so[1]);
#pragma unroll
            for (int j = 0; j < VEC; ++j) p.lab8[px + j] = (unsigned char)lab8_encode(lab[j], CT, p.has_filter, p.filter);
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
        p.lab8[px] = (unsigned char)lab8_encode(lab, C, p.has_filter, p.filter);
        stats_pixel_tail(p, sm, lab, e_lab, s, arg, C, nvalid);
    }
    if (cur_g >= 0) { if (nvalid) atomicAdd(&sm.valid, nvalid); stats_flush_group(p, sm, cur_g); }
    stats_flush_cm(p, sm);
}

// Pipelined stats (the fast path): one warp owns a contiguous range of 32-pixel warp tiles; logits AND labels arrive
// through the warp-private cp.async ring STAGES-1 tiles ahead of the math (completion tracked by wait_group, so no
// global-load scoreboard sits on the critical path).  Lane l owns pixel l of the tile: C exps in registers, dynamic
// class indices address shared memory.  Per-class counters are warp-private in shared memory (no CTA barrier in the loop).
//
// Besides the softmax state the kernel leaves a 20-byte candidate record per pixel, so that the emission pass does not
// have to read the logits again:
//   rec16 = { key of the own-class error 1 - p_label, p of the two most probable OTHER classes, upper bound of the third }
//   rec4  = label8 | class of p1 << 8 | class of p2 << 16
// Every class outside {label, c1, c2} has p <= the guard, so a pixel whose guard is below the smallest class threshold
// has no further candidates (always true when that threshold exceeds 1/3); the rest take emit_kernel_rec's slow path.
// One pixel of K1 (shared by the pipelined kernel and the upsampling kernel): T[c][lane] holds the pixel's logits.
struct StatsAcc {
    u32 nvalid = 0, oob = 0, ce_n = 0, ce_oob = 0;
    float ce_acc = 0.f;
};
template <int CT, int WT>
__device__ __forceinline__ void stats_pixel(const LovaszParams& p, const float (*T)[WT], int lane, int lab_in, size_t px,
                                            u32* s_fg_w, u32* s_key_w, u32* s_cm, const ExpConsts& ek, StatsAcc& A) {
    int lab = lab_in;
    asm volatile("" : "+r"(lab));                          // a plain 32-bit value from here on: without this the compiler compares
                                                           // the 64-bit label it was narrowed from, two ISETPs per class
    float z[CT];
#pragma unroll
    for (int c = 0; c < CT; ++c) z[c] = T[c][lane];
    float m = z[0];
#pragma unroll
    for (int c = 1; c < CT; ++c) m = fmaxf(m, z[c]);
    // softmax denominator (ascending class order, like ATen) and the three largest exps among the other classes:
    // exps are positive, so their bit patterns order like integers; 31 - class rides in the low byte (of equal exps the
    // lowest class wins: b1 then also yields torch's first-maximum argmax, see below)
    float s = 0.f;
    int b1 = -1, b2 = -1, b3 = -1;
#pragma unroll
    for (int c = 0; c < CT; ++c) {
        const float e = sm_exp_k(z[c], m, ek);
        s = __fadd_rn(s, e);
        int v = (int)__byte_perm(__float_as_uint(e), (u32)(31 - c), 0x3214);   // low byte <- 31 - class (one PRMT)
        v = (c == lab) ? -1 : v;
        const int t1v = min(b1, v); b1 = max(b1, v);
        const int t2v = min(b2, t1v); b2 = max(b2, t1v);
        b3 = max(b3, t2v);
    }
    p.pix_m[px] = m; p.pix_s[px] = s;
    const u32 l8 = lab8_encode(lab, CT, p.has_filter, p.filter);
    p.lab8[px] = (unsigned char)l8;
    u32 kfg = 0;
    if (l8 != LAB8_FILTERED) {
        ++A.nvalid;
        if (l8 < (u32)CT) {
            const float pr = sm_prob(T[l8][lane], m, s);
            kfg = err_key(__fsub_rn(1.0f, pr));
            atomicAdd(&s_fg_w[l8], 1u);
            atomicMax(&s_key_w[l8], kfg);
        }
    }
    {
        const int c1 = 31 - (b1 & 31), c2 = 31 - (b2 & 31);   // CT >= 4: b1..b3 are real classes
        const float p1 = sm_prob(T[c1][lane], m, s), p2 = sm_prob(T[c2][lane], m, s);
        const float p3 = __fdiv_ru(__uint_as_float((u32)b3 | 255u), s);  // >= p of every class not recorded
        p.rec16[px] = make_uint4(kfg, __float_as_uint(p1), __float_as_uint(p2), __float_as_uint(p3));
        p.rec4[px] = l8 | ((u32)c1 << 8) | ((u32)c2 << 16);
    }
    if (p.cm && !(p.has_drop && lab == p.drop)) {
        if ((unsigned)lab < (unsigned)CT) {
            // torch's argmax (first maximum) without a pass over the classes: the maximum's exponential is exactly 1.0, so
            // among the other classes it is b1's class (ties go to the lowest class through the low byte) unless a SECOND
            // other class also has exponential 1.0 -- a logit within an ulp of the maximum rounds to 1.0 as well -- or NaN /
            // inf are around; those pixels (~1e-6 of them) take the plain loop.  The own class competes by its exact logit.
            const int cb = 31 - (b1 & 31);
            int arg;
            if (((u32)b2 >> 8) == (ONE_BITS >> 8) || s != s || m != m) {
                float best = T[0][lane];
                arg = 0;
#pragma unroll
                for (int c = 1; c < CT; ++c) argmax_step(T[c][lane], c, best, arg);   // NaN wins, like torch
            } else {
                const bool own_max = T[lab][lane] == m, cb_max = T[cb][lane] == m;
                arg = (own_max && (!cb_max || lab < cb)) ? lab : cb;
            }
            atomicAdd(&s_cm[arg * CT + lab], 1u);
        } else A.oob = 1;
    }
    if (p.ce_enabled && !(p.has_ce_ignore && lab == p.ce_ignore)) {
        // -log softmax(z)[label] = max + log(sum) - z_label        (log_softmax + nll_loss of nn.CrossEntropyLoss)
        if ((unsigned)lab < (unsigned)CT) { A.ce_acc += (m + logf(s)) - T[lab][lane]; ++A.ce_n; }
        else A.ce_oob = 1;                               // torch raises on such a target
    }
}

template <int CT, int TPB, int STAGES, typename LT>
__global__ void __launch_bounds__(TPB) stats_kernel_async(LovaszParams p) {
    pdl_enter();
    using W = WarpTile<CT, 1>;
    constexpr int WT = W::WT, NW = TPB / 32;
    constexpr int LAB_BYTES = WT * (int)sizeof(LT);       // label bytes of one tile
    constexpr int LCH = LAB_BYTES / 16 > 0 ? LAB_BYTES / 16 : 1;   // 16-byte chunks of the label row
    constexpr int PX_PER_CH = 16 / (int)sizeof(LT);
    constexpr int STAGE_BYTES = CT * WT * 4 + 256;
    extern __shared__ __align__(16) unsigned char pipe_smem_raw[];
    __shared__ u32 s_cm[B200SEG_MAX_CLASSES * B200SEG_MAX_CLASSES];
    __shared__ u32 s_fg[NW][B200SEG_MAX_CLASSES], s_key[NW][B200SEG_MAX_CLASSES];
    __shared__ u32 s_valid, s_oob;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    unsigned char* ring = pipe_smem_raw + (size_t)warp * STAGES * STAGE_BYTES;
    for (int i = tid; i < CT * CT; i += TPB) s_cm[i] = 0;
    s_fg[warp][lane] = 0; s_key[warp][lane] = 0;
    if (tid == 0) { s_valid = 0; s_oob = 0; }
    __syncthreads();

    const u32 wtpi = (u32)((p.HW + WT - 1) / WT);
    const u32 nwt = wtpi * (u32)p.N;
    const u32 gw = blockIdx.x * NW + warp, nwarps = gridDim.x * NW;
    // tiles of a warp: interleaved (t = gw, gw + nwarps, ...: the warps running at any moment sweep one window of
    // consecutive addresses per class plane, which DRAM likes) or a contiguous range
    // (per-image mode keeps contiguous ranges: a warp flushes its counters whenever its image changes)
    const bool il = p.interleave && !p.per_image;
    const u32 step = il ? nwarps : 1u;
    const u32 t0 = il ? gw : (u32)((u64)nwt * gw / nwarps);
    const u32 t1 = il ? nwt : (u32)((u64)nwt * (gw + 1) / nwarps);
    u32 cn = t0 / wtpi, cti = t0 - cn * wtpi;             // tile being consumed
    u32 pn = cn, pti = cti, pt = t0;                       // next tile to prefetch
    auto prefetch_next = [&](int stage) {
        if (pt < t1) {
            unsigned char* sb = ring + (size_t)stage * STAGE_BYTES;
            const long long q0 = (long long)pti * WT;
            W::prefetch(reinterpret_cast<float (*)[WT]>(sb), p.logits, (int)pn, q0, p.HW, lane);
            if (lane < LCH && q0 + (long long)lane * PX_PER_CH < p.HW)
                cp_async<16>(sb + CT * WT * 4 + lane * 16,
                             (const unsigned char*)p.labels + ((size_t)pn * p.HW + q0) * sizeof(LT) + lane * 16);
        }
        cp_async_commit();
        pt += step; pti += step;
        while (pti >= wtpi) { pti -= wtpi; ++pn; }
    };
#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) prefetch_next(s);

    auto flush_group = [&](int g, u32 nvalid) {           // warp-private counters -> global (per-image mode)
        __syncwarp();
        if (lane < CT) {
            const size_t seg = (size_t)g * CT + lane;
            const u32 f = s_fg[warp][lane];
            if (f) { atomicAdd(p.seg_fg + seg, f); atomicMax(p.seg_maxkey + seg, s_key[warp][lane]); }
            s_fg[warp][lane] = 0; s_key[warp][lane] = 0;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) nvalid += __shfl_xor_sync(FULL_MASK, nvalid, o);
        if (lane == 0 && nvalid) atomicAdd(p.grp_valid + g, nvalid);
        __syncwarp();
    };

    int cur_g = -1, stage = 0, pstage = STAGES - 1;
    StatsAcc A;
    ExpConsts ek;
    ek.load();
    for (u32 t = t0; t < t1; t += step) {
        __syncwarp();                                     // every lane is done reading the stage about to be refilled
        prefetch_next(pstage);
        if (++pstage == STAGES) pstage = 0;
        const int n = (int)cn;
        const long long q = (long long)cti * WT + lane;
        cti += step;
        while (cti >= wtpi) { cti -= wtpi; ++cn; }
        const int g = p.per_image ? n : 0;
        if (g != cur_g) {
            if (cur_g >= 0) { flush_group(cur_g, A.nvalid); A.nvalid = 0; }
            cur_g = g;
        }
        cp_async_wait<STAGES - 1>();
        __syncwarp();
        const unsigned char* sb = ring + (size_t)stage * STAGE_BYTES;
        const float (*T)[WT] = reinterpret_cast<const float (*)[WT]>(sb);
        if (++stage == STAGES) stage = 0;
        if (q >= p.HW) continue;
        const size_t px = (size_t)n * p.HW + q;
        int lab;
        if constexpr (sizeof(LT) == 8) lab = sat_i32(reinterpret_cast<const long long*>(sb + CT * WT * 4)[lane]);
        else lab = (int)reinterpret_cast<const LT*>(sb + CT * WT * 4)[lane];
        stats_pixel<CT, WT>(p, T, lane, lab, px, s_fg[warp], s_key[warp], s_cm, ek, A);
    }
    cp_async_wait<0>();
    if (p.ce_enabled) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { A.ce_acc += __shfl_xor_sync(FULL_MASK, A.ce_acc, o); A.ce_n += __shfl_xor_sync(FULL_MASK, A.ce_n, o); }
        if (lane == 0 && A.ce_n) { atomicAdd(p.ce_sum, (double)A.ce_acc); atomicAdd(p.ce_cnt, A.ce_n); }
        if (__any_sync(FULL_MASK, A.ce_oob) && lane == 0 && p.status) atomicOr(p.status, STATUS_LABEL_OOB);
    }
    if (cur_g >= 0 && p.per_image) { flush_group(cur_g, A.nvalid); A.nvalid = 0; }
    // flat mode: combine the CTA's warps first (one global atomic per class and CTA)
    if (A.oob) s_oob = 1;
    if (!p.per_image) {
        u32 nvalid = A.nvalid;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) nvalid += __shfl_xor_sync(FULL_MASK, nvalid, o);
        if (lane == 0 && nvalid) atomicAdd(&s_valid, nvalid);
    }
    __syncthreads();
    if (!p.per_image) {
        if (tid < CT) {
            u32 f = 0, k = 0;
#pragma unroll
            for (int w = 0; w < NW; ++w) { f += s_fg[w][tid]; k = max(k, s_key[w][tid]); }
            if (f) { atomicAdd(p.seg_fg + tid, f); atomicMax(p.seg_maxkey + tid, k); }
        }
        if (tid == 0 && s_valid) atomicAdd(p.grp_valid, s_valid);
    }
    if (p.cm) {
        for (int i = tid; i < CT * CT; i += TPB)
            if (s_cm[i]) atomicAdd(p.cm + i, (unsigned long long)s_cm[i]);
        if (tid == 0 && s_oob) atomicOr(p.status, STATUS_LABEL_OOB);
    }
}

// --------------------------------------------------------------------------------------------------------------
// K1 (TMA variant, stats_variant = 7): the same kernel fed by the tensor-memory accelerator instead of per-warp cp.async rings.
// The logits are described to the TMA unit as a 3-D tensor {plane pixel, class, image}; ONE cp.async.bulk.tensor.3d request
// brings the box {TMA_BOX pixels x C classes x 1 image} (12.8 KB at C = 25) into a stage, a second bulk copy the tile's labels;
// both complete on the stage's mbarrier.  One elected thread issues, all 4 warps wait on the barrier and take 32 pixels each.
// Same outputs, bit for bit, as stats_kernel_async (tests/test_gpu_paths.py); kept as the measured TMA-vs-LDGSTS comparison
// (profiles/r02_tma_experiment.txt).
// --------------------------------------------------------------------------------------------------------------
#define TMA_BOX 128
#define TMA_TPB 128
#define TMA_STAGES 2
__device__ __forceinline__ void mbar_init(u32 bar, u32 count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(u32 bar, u32 bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(u32 bar, u32 parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "MBAR_WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra MBAR_DONE_%=;\n"
        "bra MBAR_WAIT_%=;\n"
        "MBAR_DONE_%=:\n"
        "}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_3d(u32 dst, const CUtensorMap* map, u32 bar, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void bulk_load_1d(u32 dst, const void* src, u32 bytes, u32 bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

template <int CT, typename LT>
__global__ void __launch_bounds__(TMA_TPB) stats_kernel_tma(LovaszParams p, const __grid_constant__ CUtensorMap tmap) {
    constexpr int NW = TMA_TPB / 32;
    constexpr int Z_BYTES = CT * TMA_BOX * 4, LAB_BYTES = TMA_BOX * (int)sizeof(LT);
    constexpr int STAGE_BYTES = (Z_BYTES + LAB_BYTES + 127) / 128 * 128;
    extern __shared__ __align__(128) unsigned char tma_smem_raw[];
    __shared__ __align__(8) unsigned long long s_bar[TMA_STAGES];
    __shared__ u32 s_cm[B200SEG_MAX_CLASSES * B200SEG_MAX_CLASSES];
    __shared__ u32 s_fg[NW][B200SEG_MAX_CLASSES], s_key[NW][B200SEG_MAX_CLASSES];
    __shared__ u32 s_valid, s_oob;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < CT * CT; i += TMA_TPB) s_cm[i] = 0;
    s_fg[warp][lane] = 0; s_key[warp][lane] = 0;
    if (tid == 0) {
        s_valid = 0; s_oob = 0;
        for (int s2 = 0; s2 < TMA_STAGES; ++s2) mbar_init(smem_u32(&s_bar[s2]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const u32 bpi = (u32)((p.HW + TMA_BOX - 1) / TMA_BOX);             // boxes per image
    const u32 nbox = bpi * (u32)p.N;
    const u32 my_n = blockIdx.x < nbox ? (nbox - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;     // boxes of this CTA (interleaved)
    auto issue = [&](u32 it) {                             // (thread 0) request box `it` of this CTA into stage it % TMA_STAGES
        const u32 t = blockIdx.x + it * gridDim.x;
        const u32 n = t / bpi, b = t - n * bpi;
        const u32 st = it % TMA_STAGES;
        const u32 dst = smem_u32(tma_smem_raw + (size_t)st * STAGE_BYTES), bar = smem_u32(&s_bar[st]);
        const long long q0 = (long long)b * TMA_BOX;
        const u32 npx = (u32)min((long long)TMA_BOX, p.HW - q0);
        const u32 lab_bytes = npx * (u32)sizeof(LT);       // (plane % 16 == 0 and 16-byte aligned labels: a multiple of 16)
        mbar_expect_tx(bar, (u32)Z_BYTES + lab_bytes);     // the box is always written in full (out-of-range pixels: zeros)
        tma_load_3d(dst, &tmap, bar, (int)q0, 0, (int)n);
        bulk_load_1d(dst + Z_BYTES, (const unsigned char*)p.labels + ((size_t)n * p.HW + q0) * sizeof(LT), lab_bytes, bar);
    };
    if (tid == 0)
        for (u32 it = 0; it < TMA_STAGES - 1 && it < my_n; ++it) issue(it);

    auto flush_group = [&](int g, u32 nvalid) {           // warp-private counters -> global (per-image mode)
        __syncwarp();
        if (lane < CT) {
            const size_t seg = (size_t)g * CT + lane;
            const u32 f = s_fg[warp][lane];
            if (f) { atomicAdd(p.seg_fg + seg, f); atomicMax(p.seg_maxkey + seg, s_key[warp][lane]); }
            s_fg[warp][lane] = 0; s_key[warp][lane] = 0;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) nvalid += __shfl_xor_sync(FULL_MASK, nvalid, o);
        if (lane == 0 && nvalid) atomicAdd(p.grp_valid + g, nvalid);
        __syncwarp();
    };
    int cur_g = -1;
    u32 nvalid = 0, oob = 0, ce_n = 0, ce_oob = 0;
    float ce_acc = 0.f;
    ExpConsts ek;
    ek.load();
    for (u32 it = 0; it < my_n; ++it) {
        if (tid == 0 && it + TMA_STAGES - 1 < my_n) issue(it + TMA_STAGES - 1);     // (its stage was released by the barrier below)
        const u32 st = it % TMA_STAGES;
        mbar_wait(smem_u32(&s_bar[st]), (it / TMA_STAGES) & 1u);
        const u32 t = blockIdx.x + it * gridDim.x;
        const int n = (int)(t / bpi);
        const long long q = (long long)(t - (u32)n * bpi) * TMA_BOX + warp * 32 + lane;
        const int g = p.per_image ? n : 0;
        if (g != cur_g) {
            if (cur_g >= 0) { flush_group(cur_g, nvalid); nvalid = 0; }
            cur_g = g;
        }
        const unsigned char* sb = tma_smem_raw + (size_t)st * STAGE_BYTES;
        const float (*T)[TMA_BOX] = reinterpret_cast<const float (*)[TMA_BOX]>(sb);
        const int col = warp * 32 + lane;
        if (q < p.HW) {
            const size_t px = (size_t)n * p.HW + q;
            int lab;
            if constexpr (sizeof(LT) == 8) lab = sat_i32(reinterpret_cast<const long long*>(sb + Z_BYTES)[col]);
            else lab = (int)reinterpret_cast<const LT*>(sb + Z_BYTES)[col];
            float z[CT];
#pragma unroll
            for (int c = 0; c < CT; ++c) z[c] = T[c][col];
            float m = z[0];
#pragma unroll
            for (int c = 1; c < CT; ++c) m = fmaxf(m, z[c]);
            float s = 0.f;
            int b1 = -1, b2 = -1, b3 = -1;
#pragma unroll
            for (int c = 0; c < CT; ++c) {
                const float e = sm_exp_k(z[c], m, ek);
                s = __fadd_rn(s, e);
                int v = (int)__byte_perm(__float_as_uint(e), (u32)c, 0x3214);
                v = (c == lab) ? -1 : v;
                const int t1v = min(b1, v); b1 = max(b1, v);
                const int t2v = min(b2, t1v); b2 = max(b2, t1v);
                b3 = max(b3, t2v);
            }
            p.pix_m[px] = m; p.pix_s[px] = s;
            const u32 l8 = lab8_encode(lab, CT, p.has_filter, p.filter);
            p.lab8[px] = (unsigned char)l8;
            u32 kfg = 0;
            if (l8 != LAB8_FILTERED) {
                ++nvalid;
                if (l8 < (u32)CT) {
                    const float pr = sm_prob(T[l8][col], m, s);
                    kfg = err_key(__fsub_rn(1.0f, pr));
                    atomicAdd(&s_fg[warp][l8], 1u);
                    atomicMax(&s_key[warp][l8], kfg);
                }
            }
            {
                const int c1 = b1 & 31, c2 = b2 & 31;
                const float p1 = sm_prob(T[c1][col], m, s), p2 = sm_prob(T[c2][col], m, s);
                const float p3 = __fdiv_ru(__uint_as_float((u32)b3 | 255u), s);
                p.rec16[px] = make_uint4(kfg, __float_as_uint(p1), __float_as_uint(p2), __float_as_uint(p3));
                p.rec4[px] = l8 | ((u32)c1 << 8) | ((u32)c2 << 16);
            }
            if (p.cm && !(p.has_drop && lab == p.drop)) {
                if ((unsigned)lab < (unsigned)CT) {
                    int arg = 0;
#pragma unroll
                    for (int c = CT - 1; c >= 0; --c) arg = (z[c] == m) ? c : arg;
                    if (s != s || m != m) {
                        float best = z[0];
                        arg = 0;
#pragma unroll
                        for (int c = 1; c < CT; ++c) argmax_step(z[c], c, best, arg);
                    }
                    atomicAdd(&s_cm[arg * CT + lab], 1u);
                } else oob = 1;
            }
            if (p.ce_enabled && !(p.has_ce_ignore && lab == p.ce_ignore)) {
                if ((unsigned)lab < (unsigned)CT) { ce_acc += (m + logf(s)) - T[lab][col]; ++ce_n; }
                else ce_oob = 1;
            }
        }
        __syncthreads();                                   // the stage may be refilled
    }
    if (p.ce_enabled) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { ce_acc += __shfl_xor_sync(FULL_MASK, ce_acc, o); ce_n += __shfl_xor_sync(FULL_MASK, ce_n, o); }
        if (lane == 0 && ce_n) { atomicAdd(p.ce_sum, (double)ce_acc); atomicAdd(p.ce_cnt, ce_n); }
        if (__any_sync(FULL_MASK, ce_oob) && lane == 0 && p.status) atomicOr(p.status, STATUS_LABEL_OOB);
    }
    if (cur_g >= 0 && p.per_image) { flush_group(cur_g, nvalid); nvalid = 0; }
    if (oob) s_oob = 1;
    if (!p.per_image) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) nvalid += __shfl_xor_sync(FULL_MASK, nvalid, o);
        if (lane == 0 && nvalid) atomicAdd(&s_valid, nvalid);
    }
    __syncthreads();
    if (!p.per_image) {
        if (tid < CT) {
            u32 f = 0, k = 0;
#pragma unroll
            for (int w = 0; w < NW; ++w) { f += s_fg[w][tid]; k = max(k, s_key[w][tid]); }
            if (f) { atomicAdd(p.seg_fg + tid, f); atomicMax(p.seg_maxkey + tid, k); }
        }
        if (tid == 0 && s_valid) atomicAdd(p.grp_valid, s_valid);
    }
    if (p.cm) {
        for (int i = tid; i < CT * CT; i += TMA_TPB)
            if (s_cm[i]) atomicAdd(p.cm + i, (unsigned long long)s_cm[i]);
        if (tid == 0 && s_oob) atomicOr(p.status, STATUS_LABEL_OOB);
    }
}

// tensor map of the logits for stats_kernel_tma: {plane pixel, class, image}, box {TMA_BOX, C, 1}
static int make_logits_tensor_map(CUtensorMap* map, const float* logits, int n, int c, long long hw) {
    typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                 const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static EncodeFn encode = nullptr;
    if (!encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        CUDA_TRY(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
        if (!fn || qres != cudaDriverEntryPointSuccess) { b200seg_set_error("cuTensorMapEncodeTiled is not available"); return B200SEG_E_UNSUPPORTED; }
        encode = (EncodeFn)fn;
    }
    const cuuint64_t dims[3] = {(cuuint64_t)hw, (cuuint64_t)c, (cuuint64_t)n};
    const cuuint64_t strides[2] = {(cuuint64_t)hw * 4, (cuuint64_t)hw * c * 4};
    const cuuint32_t box[3] = {TMA_BOX, (cuuint32_t)c, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    const CUresult r = encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void*)logits, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                              CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { b200seg_set_error("cuTensorMapEncodeTiled failed (%d)", (int)r); return B200SEG_E_UNSUPPORTED; }
    return 0;
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
        const float z = p.up.lo ? up_logit(p.up, p.C, (int)n, c, q) : __ldg(p.logits + ((size_t)n * p.C + c) * p.HW + q);
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
// One kernel: a warp per group turns the counters of K1 into thresholds, key widths, class weights and the class order
// (lane = class), then the block samples the guards of that group's records to choose the emission path.  The record
// path is exact for any input but pays a gather per (pixel, class) whose threshold the pixel's guard reaches; the sample
// estimates how many there are (a performance heuristic only).  Flat mode (one group): every block repeats the tiny
// per-class part for itself, block 0 publishes it, all blocks share the sampling.
#define DECIDE_BLOCKS 32
#define DECIDE_TPB 256
__global__ void __launch_bounds__(DECIDE_TPB) finalize_decide_kernel(LovaszParams p) {
    pdl_enter();
    __shared__ float s_tmin;
    __shared__ float s_sorted[B200SEG_MAX_CLASSES];               // thresholds of the group, ascending
    __shared__ u32 s_slow, s_n;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool flat = p.groups == 1;
    if (blockIdx.x == 0 && tid == 0) {
        *p.geo = p.geo_stream;                                    // default; the deciding block may override it
        if (p.ce_enabled) {                                       // mean over the non-ignored pixels (0 / 0 = NaN, like torch)
            const u32 n = *p.ce_cnt;
            *p.ce_out = (float)(*p.ce_sum / (double)n);
            *p.ce_inv_n = n ? 1.0f / (float)n : 0.f;
        }
    }
    if (tid == 0) { s_slow = 0; s_n = 0; }
    u32 slow = 0, n = 0;
    for (int g = flat ? 0 : blockIdx.x; g < p.groups; g += gridDim.x) {
        __syncthreads();
        if (warp == 0) {
            const bool writer = !flat || blockIdx.x == 0;
            const int c = lane;
            const size_t seg = (size_t)g * p.C + c;
            const bool any_valid = p.grp_valid[g] > 0;
            bool active = false;
            float thr = THR_INACTIVE, logthr = __int_as_float(0x7f800000);
            u32 bits = 1;
            if (c < p.C) {
                const u32 fg = p.seg_fg[seg];
                active = ((p.class_mask >> c) & 1u) && any_valid && (fg > 0 || p.keep_absent);
                if (active) {
                    const u32 thr_bits = fg > 0 ? (ONE_BITS - p.seg_maxkey[seg]) : p.seg_maxp[seg];
                    thr = __uint_as_float(thr_bits);
                    logthr = thr > 0.f ? logf(thr) : __int_as_float(0xff800000);
                    const u32 maxkey = ONE_BITS - thr_bits;
                    bits = maxkey ? (32 - __clz(maxkey)) : 1;
                }
            }
            const int nkept = __popc(__ballot_sync(FULL_MASK, active));
            float tmin = thr;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) tmin = fminf(tmin, __shfl_xor_sync(FULL_MASK, tmin, o));
            int rank = 0;                                  // position of class c in ascending (threshold, class) order
            for (int j = 0; j < p.C; ++j) {
                const float tj = __shfl_sync(FULL_MASK, thr, j);
                rank += (tj < thr || (tj == thr && j < c)) ? 1 : 0;
            }
            // d(mean)/d(term): the reference's mean() divides only when it averaged more than one value
            float w = 1.0f;
            if (p.groups > 1) w = w / (float)p.groups;
            if (nkept > 1) w = w / (float)nkept;
            if (writer && c < p.C) {
                p.seg_thr[seg] = thr; p.seg_logthr[seg] = logthr; p.seg_bits[seg] = bits;
                p.seg_w[seg] = active ? w : 0.f;
                p.seg_order[(size_t)g * p.C + rank] = (unsigned char)c;
            }
            if (c < p.C) s_sorted[rank] = thr;
            if (lane == 0) { s_tmin = tmin; if (writer) p.grp_tmin[g] = tmin; }
        }
        __syncthreads();
        if (p.have_records) {
            const float tmin = s_tmin;
            long long target = 65536 / p.groups;
            if (target < 1024) target = 1024;
            const long long stride = p.cap > target ? p.cap / target : 1;
            const long long i0 = flat ? (long long)blockIdx.x * DECIDE_TPB + tid : tid;
            const long long di = flat ? (long long)gridDim.x * DECIDE_TPB : DECIDE_TPB;
            const u32* guard = reinterpret_cast<const u32*>(p.rec16 + (size_t)g * p.cap) + 3;   // rec16[px].w
            for (long long i = i0; i * stride < p.cap; i += 8 * di) {       // 8 independent loads in flight per thread
                u32 w[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const long long px = (i + k * di) * stride;
                    w[k] = px < p.cap ? __ldg(guard + 4 * px) : 0xFFFFFFFFu;       // NaN pattern: compares false
                }
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const float gd = __uint_as_float(w[k]);
                    if (gd >= tmin)                          // classes the record kernel would have to look at for this pixel
                        for (int j = 0; j < p.C && s_sorted[j] <= gd; ++j) ++slow;
                    n += (i + k * di) * stride < p.cap;
                }
            }
        }
    }
    if (!p.have_records) return;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { slow += __shfl_xor_sync(FULL_MASK, slow, o); n += __shfl_xor_sync(FULL_MASK, n, o); }
    if (lane == 0) { atomicAdd(&s_slow, slow); atomicAdd(&s_n, n); }
    __syncthreads();
    if (tid == 0) {
        atomicAdd(p.ctrl + CTRL_SLOW, s_slow);
        atomicAdd(p.ctrl + CTRL_SAMPLES, s_n);
        __threadfence();
        if (atomicAdd(p.ctrl + CTRL_TICKET, 1u) == gridDim.x - 1) {
            __threadfence();
            const u32 ts = ld_relaxed(p.ctrl + CTRL_SLOW), tn = ld_relaxed(p.ctrl + CTRL_SAMPLES);
            // ts / tn = expected number of logits a pixel must fetch beyond its record; the record path wins while that
            // stays well below one (benchmark distribution: 0.0005 at C=25, ~0.1 at C=17; trained-like logits: several)
            const bool rec = p.emit_force == 1 || (p.emit_force == 0 && (u64)ts * 2 <= tn);
            p.flags[0] = rec ? EMIT_PATH_RECORDS : EMIT_PATH_STREAM;
            if (rec) *p.geo = p.geo_rec;
        }
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
    __shared__ u32 s_gbase[B200SEG_MAX_CLASSES];          // the tile's slice of every segment (reserved with one atomic per class)
    __shared__ float s_thr[B200SEG_MAX_CLASSES], s_logthr[B200SEG_MAX_CLASSES];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int C = p.C;
    const long long tpi = (p.HW + TILE_PX - 1) / TILE_PX;
    const long long tpg = p.per_image ? tpi : tpi * p.N;
    const EmitGeomDev G = *p.geo;
    const long long total_chunks = (long long)p.groups * G.n_runs;
    int cur_g = -1;

    for (long long chunk = blockIdx.x; chunk < total_chunks; chunk += gridDim.x) {
        const int g = (int)(chunk / G.n_runs);
        const long long r = chunk - (long long)g * G.n_runs;
        const long long gt0 = r * G.tiles_per_chunk;
        const long long gt1 = min(gt0 + (long long)G.tiles_per_chunk, tpg);
        __syncthreads();                                   // previous chunk fully written out
        if (g != cur_g) {
            if (tid < C) { s_thr[tid] = p.seg_thr[(size_t)g * C + tid]; s_logthr[tid] = p.seg_logthr[(size_t)g * C + tid]; }
            cur_g = g;
        }
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
#pragma unroll
                for (int j = 0; j < VEC; ++j)        // background candidates of the pixel, for the backward pass
                    p.cmask[px0 + j] = (unsigned)lab[j] < (unsigned)C ? (acc[j] & ~(1u << lab[j])) : acc[j];
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
                if (lane == 31) { s_tot[c] = v; s_gbase[c] = v ? atomicAdd(p.seg_count + (size_t)g * C + c, v) : 0u; }
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
                        const u32 rank = s_gbase[c] + s_wpre[c][bit >> 5] +
                                         __popc(s_mask[c][bit >> 5] & ((1u << (bit & 31)) - 1u));
                        const size_t slot = ((size_t)g * C + c) * (size_t)p.cap + rank;
                        p.keysA[slot] = err_key(err) | (fg ? KEY_FG : 0u);
                        p.valsA[slot] = ((u32)(px0 + j) << 1) | (fg ? 1u : 0u);
                    }
                }
            }
            __syncthreads();
            for (int i = tid; i < B200SEG_MAX_CLASSES * WORDS; i += EMIT_TPB) (&s_mask[0][0])[i] = 0;
            __syncthreads();
        }
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

// One 32-pixel tile of the streaming emission (lane = pixel; Tz[c][lane] = the pixel's logits): exact candidate test per class,
// one reservation per (tile, class), lane-private stores.  Shared by emit_kernel_async and emit_kernel_up.
template <int CT, int WT>
__device__ __forceinline__ void emit_tile(const LovaszParams& p, const float (*Tz)[WT], int lane, u32 lt_mask, float m, float s,
                                          u32 l8, bool inb, size_t px, int g, const float* thr, const float* logthr_w,
                                          u32* mask_w, u32* base_w) {
    u32 acc = 0;                                      // accepted classes of this lane's pixel
    float e0 = 0.f, e1 = 0.f;                         // errors of the first two of them (the rest is recomputed)
    const int lab = l8 < (u32)CT ? (int)l8 : -1;
    if (inb) {
        const float theta = pre_theta(m, s);
        u32 pm = 0;
        const float4* lt4 = reinterpret_cast<const float4*>(logthr_w);   // broadcast 128-bit reads
#pragma unroll
        for (int c4 = 0; c4 < (CT + 3) / 4; ++c4) {
            const float4 lt = lt4[c4];
            const float l[4] = {lt.x, lt.y, lt.z, lt.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int c = 4 * c4 + j;
                if (c < CT) pm |= (Tz[c][lane] >= theta + l[j]) ? (1u << c) : 0u;
            }
        }
        if (l8 == LAB8_FILTERED) pm = 0;
        else if (lab >= 0 && thr_active(thr[lab])) pm |= 1u << lab;
        while (pm) {
            const int c = __ffs(pm) - 1;
            pm &= pm - 1;
            float err, pr;
            if (exact_accept(Tz[c][lane], m, s, c == lab, thr[c], err, pr)) {
                if (acc == 0) e0 = err; else if ((acc & (acc - 1)) == 0) e1 = err;
                acc |= 1u << c;
            }
        }
        p.cmask[px] = lab >= 0 ? (acc & ~(1u << lab)) : acc;
    }
    if (__any_sync(FULL_MASK, acc != 0)) {
        // lane c collects the ballot of class c (independent votes, unrolled: a data-dependent loop over the classes
        // present would serialise their latencies) and reserves the slots; the table goes through shared memory
        u32 mine = 0;
#pragma unroll
        for (int c = 0; c < CT; ++c) {
            const u32 b = __ballot_sync(FULL_MASK, (acc >> c) & 1u);
            if (lane == c) mine = b;
        }
        __syncwarp();
        if (lane < CT) {                               // the tile's slice of every segment: one atomic per class with candidates
            const u32 cnt = __popc(mine);
            mask_w[lane] = mine;
            base_w[lane] = cnt ? atomicAdd(p.seg_count + (size_t)g * CT + lane, cnt) : 0u;
        }
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
            const u32 rank = base_w[c] + __popc(mask_w[c] & lt_mask);
            const size_t slot = ((size_t)g * CT + c) * (size_t)p.cap + rank;
            p.keysA[slot] = err_key(err) | (fg ? KEY_FG : 0u);
            p.valsA[slot] = ((u32)px << 1) | (fg ? 1u : 0u);
        }
    }
}

template <int CT, int TPB>
__global__ void __launch_bounds__(TPB) emit_kernel_async(LovaszParams p) {
    pdl_enter();
    if (p.flags[0] != EMIT_PATH_STREAM) return;           // ctrl is zeroed per call: the default path is this one
    using W = WarpTile<CT, 1>;
    constexpr int WT = W::WT, NW = TPB / 32, STAGES = 2;  // the prefetch cursor runs exactly one tile ahead
    extern __shared__ __align__(16) unsigned char pipe_smem_raw[];
    float (*Zall)[STAGES][CT][WT] = reinterpret_cast<float (*)[STAGES][CT][WT]>(pipe_smem_raw);
    __shared__ float s_thr[NW][B200SEG_MAX_CLASSES];
    __shared__ __align__(16) float s_logthr[NW][B200SEG_MAX_CLASSES];
    __shared__ u32 s_mask[NW][B200SEG_MAX_CLASSES], s_base[NW][B200SEG_MAX_CLASSES];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const u32 lt_mask = (1u << lane) - 1;
    float (*Z)[CT][WT] = Zall[warp];
    // all tile / chunk counts fit 32 bits (n_images * plane < 2^30)
    const u32 wtpi = (u32)((p.HW + WT - 1) / WT);
    const u32 tpg = p.per_image ? wtpi : wtpi * (u32)p.N;
    const EmitGeomDev G = *p.geo;
    const u32 tpc = (u32)G.tiles_per_chunk, n_runs = (u32)G.n_runs;
    const u32 total_chunks = (u32)p.groups * n_runs;
    const u32 gw = blockIdx.x * NW + warp, nwarps = gridDim.x * NW;
    const u32 ch0 = (u32)((u64)total_chunks * gw / nwarps), ch1 = (u32)((u64)total_chunks * (gw + 1) / nwarps);
    const u32 ntile = (ch1 - ch0) * tpc;                  // this warp's slice of the (chunk, tile-in-chunk) sequence
    if (ntile == 0) return;

    EmitCursor cur, pre;
    emit_cursor_init(cur, ch0, tpc, n_runs, wtpi, p.per_image);
    pre = cur;
    // per-pixel state of the tile in flight (softmax max / denominator, compact label): loaded with the prefetch,
    // consumed one iteration later
    float nm = 0.f, ns = 1.f;
    u32 nl8 = LAB8_FILTERED;
    auto prefetch = [&](const EmitCursor& c, bool live, int stage) {
        nm = 0.f; ns = 1.f; nl8 = LAB8_FILTERED;
        if (live && c.gt < tpg) {
            W::prefetch(Z[stage], p.logits, c.n, (long long)c.ti * WT, p.HW, lane);
            const long long q = (long long)c.ti * WT + lane;
            if (q < p.HW) {
                const size_t px = (size_t)c.n * p.HW + q;
                nm = p.pix_m[px]; ns = p.pix_s[px]; nl8 = p.lab8[px];
            }
        }
        cp_async_commit();
    };
    prefetch(pre, true, 0);

    int cur_g = -1;
    for (u32 it = 0; it < ntile; ++it) {
        const int stage = (int)(it & 1);
        const float m = nm, s = ns;
        const u32 l8 = nl8;
        __syncwarp();
        emit_cursor_next(pre, tpc, n_runs, wtpi, p.per_image);
        prefetch(pre, it + 1 < ntile, stage ^ 1);
        const int g = cur.g, n = cur.n;
        const bool exists = cur.gt < tpg;
        if (g != cur_g) {
            __syncwarp();
            if (lane < CT) { s_thr[warp][lane] = p.seg_thr[(size_t)g * CT + lane]; s_logthr[warp][lane] = p.seg_logthr[(size_t)g * CT + lane]; }
            cur_g = g;
            __syncwarp();
        }
        const long long q = (long long)cur.ti * WT + lane;
        const bool inb = exists && q < p.HW;
        const size_t px = (size_t)n * p.HW + q;
        cp_async_wait<STAGES - 1>();
        __syncwarp();
        emit_tile<CT, WT>(p, Z[stage], lane, lt_mask, m, s, l8, inb, px, g, s_thr[warp], s_logthr[warp], s_mask[warp], s_base[warp]);
        emit_cursor_next(cur, tpc, n_runs, wtpi, p.per_image);
    }
    cp_async_wait<0>();
}

// Record-driven emission (the fast path after stats_kernel_async): candidates come from the 20-byte per-pixel records;
// the logits are only touched by pixels whose guard says a class beyond the two recorded ones could qualify.
// One CTA per chunk of consecutive ECTA_TILE-pixel tiles (512 threads x 4 pixels).  Per tile: every thread decodes 4 pixels
// and takes an arrival rank per candidate from a shared per-class counter (the order inside a segment is free: the sort
// restores the canonical one), the candidates are staged in shared memory grouped by class, and a warp per class reserves
// the tile's slice of the segment with one global atomic and writes it with contiguous stores.
// `emit_scan_pixel`: classes outside `skip` whose probability reaches their threshold, for one pixel whose unrecorded
// classes all have p <= guard (the rare path; deliberately not inlined so that the common path carries neither its
// instructions nor its registers).  `order` lists the classes by ascending threshold: only the leading ones with
// threshold <= guard can qualify, typically a single class.
__device__ __noinline__ u32 emit_scan_pixel(const float* lp, long long plane, int C, float m, float sden, float guard,
                                            const float* thr, const unsigned char* order, u32 skip) {
    u32 more = 0;
    for (int i = 0; i < C; ++i) {
        const int c = order[i];
        if (thr[c] > guard) break;
        if (!((skip >> c) & 1u) && sm_prob(__ldg(lp + (size_t)c * plane), m, sden) >= thr[c]) more |= 1u << c;
    }
    return more;
}

// the same scan over interpolated logits (fused upsampling: the record path is the only emission path there)
__device__ __noinline__ float up_logit_call(UpSrc u, int C, int n, int c, long long q) { return up_logit(u, C, n, c, q); }
__device__ __noinline__ u32 emit_scan_pixel_up(UpSrc u, int C, int n, long long q, float m, float sden, float guard,
                                               const float* thr, const unsigned char* order, u32 skip) {
    u32 more = 0;
    for (int i = 0; i < C; ++i) {
        const int c = order[i];
        if (thr[c] > guard) break;
        if (!((skip >> c) & 1u) && sm_prob(up_logit(u, C, n, c, q), m, sden) >= thr[c]) more |= 1u << c;
    }
    return more;
}

#define ECTA_CAP 6144                                     // staged candidates per pass over the classes of a tile (>= 3 per pixel)
struct EctaSmem {
    // counters of two tiles in turn: while a tile is staged and written out, the next tile's set is cleared
    u32 cnt2[2][B200SEG_MAX_CLASSES];                     // recorded candidates of the class in this tile (arrival counter)
    u32 cntx2[2][B200SEG_MAX_CLASSES];                    // candidates beyond the recorded ones (rare path), counted ...
    u32 curx2[2][B200SEG_MAX_CLASSES];                    // ... and placed behind the recorded ones of their class
    float thr[B200SEG_MAX_CLASSES];
    unsigned char order[B200SEG_MAX_CLASSES];             // classes by ascending threshold
    u32 wcoff[ECTA_TPB / 32][B200SEG_MAX_CLASSES];        // per-warp copy of the stage offsets of the classes
    u32 stageK[ECTA_CAP], stageV[ECTA_CAP];
};

template <int CT, bool UP>
__global__ void __launch_bounds__(ECTA_TPB, ECTA_MINB) emit_kernel_cta(LovaszParams p) {
    pdl_enter();
    if (p.flags[0] != EMIT_PATH_RECORDS) return;
    extern __shared__ __align__(16) unsigned char ecta_smem_raw[];
    EctaSmem& S = *reinterpret_cast<EctaSmem*>(ecta_smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int NW = ECTA_TPB / 32;
    const EmitGeomDev G = *p.geo;
    const long long tpi = (p.HW + ECTA_TILE - 1) / ECTA_TILE;
    const long long tpg = p.per_image ? tpi : tpi * p.N;
    const long long total_chunks = (long long)p.groups * G.n_runs;
    int cur_g = -1;
    float tmin = 0.f;
    u32 par = 0;                                          // counter set of the current tile
    if (tid < B200SEG_MAX_CLASSES) { S.cnt2[0][tid] = 0; S.cntx2[0][tid] = 0; S.curx2[0][tid] = 0; }
    __syncthreads();

    for (long long chunk = blockIdx.x; chunk < total_chunks; chunk += gridDim.x) {
        const int g = (int)(chunk / G.n_runs);
        const long long r = chunk - (long long)g * G.n_runs;
        const long long gt0 = r * G.tiles_per_chunk;
        const long long gt1 = min(gt0 + (long long)G.tiles_per_chunk, tpg);
        if (g != cur_g) {
            __syncthreads();                               // previous tile done with the thresholds
            if (tid < B200SEG_MAX_CLASSES) {
                S.thr[tid] = tid < CT ? p.seg_thr[(size_t)g * CT + tid] : THR_INACTIVE;
                S.order[tid] = tid < CT ? p.seg_order[(size_t)g * CT + tid] : 0;
            }
            tmin = p.grp_tmin[g];
            cur_g = g;
            __syncthreads();                               // thresholds visible
        }
        // (image, tile-in-image) of the chunk's first tile; advanced by increments (no 64-bit division per tile)
        int n = p.per_image ? g : (int)(gt0 / tpi);
        long long ti = p.per_image ? gt0 : gt0 - (long long)n * tpi;
        for (long long gt = gt0; gt < gt1; ++gt, ++ti) {
            if (!p.per_image && ti == tpi) { ti = 0; ++n; }
            u32* cnt = S.cnt2[par];
            u32* cntx = S.cntx2[par];
            u32* curx = S.curx2[par];
            const long long q0 = ti * ECTA_TILE + (long long)tid * 4;
            const bool inb = q0 < p.HW;                    // plane % 4 == 0: a thread's 4 pixels are in or out together
            const size_t px0 = (size_t)n * p.HW + q0;
            // ---- decode: per pixel up to three recorded candidates (own class, c1, c2), rarely more ----------------------
            u32 kfg[4], kc1[4], kc2[4], acc[4], cls[4];      // cls = label8 | c1 << 8 | c2 << 16
            unsigned short r0[4], r1[4], r2[4];              // arrival ranks of the three recorded candidates
#pragma unroll
            for (int j = 0; j < 4; ++j) { kfg[j] = 0; kc1[j] = 0; kc2[j] = 0; acc[j] = 0; cls[j] = LAB8_FILTERED; r0[j] = r1[j] = r2[j] = 0; }
            u32 extra = 0;                                 // pixels (bits 0..3) with candidates beyond the recorded ones
            if (inb) {
                const uint4 r4v = *reinterpret_cast<const uint4*>(p.rec4 + px0);
                const u32 r4[4] = {r4v.x, r4v.y, r4v.z, r4v.w};
                uint4 rec[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) rec[j] = p.rec16[px0 + j];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const u32 l8 = r4[j] & 255u, c1 = (r4[j] >> 8) & 31u, c2 = (r4[j] >> 16) & 31u;
                    cls[j] = r4[j] & 0x00FFFFFFu;
                    const float p1 = __uint_as_float(rec[j].y), p2 = __uint_as_float(rec[j].z);
                    kfg[j] = rec[j].x | KEY_FG; kc1[j] = err_key(p1); kc2[j] = err_key(p2);
                    if (l8 != LAB8_FILTERED) {
                        u32 a = 0;
                        if (l8 < (u32)CT && thr_active(S.thr[l8 & 31u])) a |= 1u << l8;
                        if (p1 >= S.thr[c1]) a |= 1u << c1;
                        if (p2 >= S.thr[c2]) a |= 1u << c2;
                        if (__uint_as_float(rec[j].w) >= tmin) {          // rare: scan the other classes (a real call)
                            const u32 skip = (l8 < (u32)CT ? 1u << l8 : 0u) | (1u << c1) | (1u << c2);
                            u32 more;
                            if constexpr (UP)
                                more = emit_scan_pixel_up(p.up, CT, n, q0 + j, p.pix_m[px0 + j], p.pix_s[px0 + j],
                                                          __uint_as_float(rec[j].w), S.thr, S.order, skip);
                            else
                                more = (p.dbg & 4) ? 0u : emit_scan_pixel(p.logits + (size_t)n * CT * p.HW + q0 + j, p.HW, CT,
                                                             p.pix_m[px0 + j], p.pix_s[px0 + j], __uint_as_float(rec[j].w),
                                                             S.thr, S.order, skip);
                            if (more) { a |= more; extra |= 1u << j; }
                        }
                        acc[j] = a;
                    }
                }
                uint4 cm4;
                {
                    u32 cmv[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const u32 l8 = cls[j] & 255u;
                        cmv[j] = l8 < (u32)CT ? (acc[j] & ~(1u << l8)) : acc[j];
                    }
                    cm4 = make_uint4(cmv[0], cmv[1], cmv[2], cmv[3]);
                }
                *reinterpret_cast<uint4*>(p.cmask + px0) = cm4;
                // ---- arrival rank of every candidate within its class ------------------------------------------------------------
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const u32 l8 = cls[j] & 255u, c1 = (cls[j] >> 8) & 31u, c2 = (cls[j] >> 16) & 31u;
                    u32 a = acc[j];
                    if (l8 < (u32)CT && ((a >> l8) & 1u)) { r0[j] = (unsigned short)atomicAdd(&cnt[l8], 1u); a &= ~(1u << l8); }
                    if ((a >> c1) & 1u) { r1[j] = (unsigned short)atomicAdd(&cnt[c1], 1u); a &= ~(1u << c1); }
                    if ((a >> c2) & 1u) { r2[j] = (unsigned short)atomicAdd(&cnt[c2], 1u); a &= ~(1u << c2); }
                    if ((extra >> j) & 1u)
                        while (a) { atomicAdd(&cntx[__ffs(a) - 1], 1u); a &= a - 1; }
                }
            }
            __syncthreads();                               // all arrivals counted (and the previous tile written out)
            if (tid < B200SEG_MAX_CLASSES) { S.cnt2[par ^ 1][tid] = 0; S.cntx2[par ^ 1][tid] = 0; S.curx2[par ^ 1][tid] = 0; }
            // ---- stage offsets: every warp scans the class totals for itself (no extra CTA barrier) ---------------------------
            u32 total_all;
            {
                const u32 tc = lane < CT ? cnt[lane] + cntx[lane] : 0u;
                u32 v = tc;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { const u32 x = __shfl_up_sync(FULL_MASK, v, o); if (lane >= o) v += x; }
                S.wcoff[warp][lane] = v - tc;
                total_all = __shfl_sync(FULL_MASK, v, 31);
                __syncwarp();
            }
            // ---- passes over class ranges whose candidates fit the stage (one pass unless > ECTA_CAP candidates) ---------
            const u32* coff = S.wcoff[warp];
            int lo = 0;
            while (lo < CT) {
                int hi = CT;
                u32 base = 0;                              // stage offset of class lo
                if (total_all > ECTA_CAP) {
                    hi = lo;
                    u32 sum = 0;
                    while (hi < CT && sum + cnt[hi] + cntx[hi] <= ECTA_CAP) { sum += cnt[hi] + cntx[hi]; ++hi; }   // a class <= 2048: hi > lo
                    base = coff[lo];
                }
                if (inb) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const u32 l8 = cls[j] & 255u, c1 = (cls[j] >> 8) & 31u, c2 = (cls[j] >> 16) & 31u;
                        const u32 v0 = (u32)(px0 + j) << 1;
                        u32 a = acc[j];
                        if (l8 < (u32)CT && ((a >> l8) & 1u)) {
                            a &= ~(1u << l8);
                            if ((int)l8 >= lo && (int)l8 < hi) { const u32 pos = coff[l8] - base + r0[j]; S.stageK[pos] = kfg[j]; S.stageV[pos] = v0 | 1u; }
                        }
                        if ((a >> c1) & 1u) {
                            a &= ~(1u << c1);
                            if ((int)c1 >= lo && (int)c1 < hi) { const u32 pos = coff[c1] - base + r1[j]; S.stageK[pos] = kc1[j]; S.stageV[pos] = v0; }
                        }
                        if ((a >> c2) & 1u) {
                            a &= ~(1u << c2);
                            if ((int)c2 >= lo && (int)c2 < hi) { const u32 pos = coff[c2] - base + r2[j]; S.stageK[pos] = kc2[j]; S.stageV[pos] = v0; }
                        }
                        if ((extra >> j) & 1u) {            // rare: classes found by the scan go behind the recorded ones of their class
                            while (a) {
                                const int c = __ffs(a) - 1;
                                a &= a - 1;
                                if (c < lo || c >= hi) continue;
                                float zc;
                                if constexpr (UP) zc = up_logit_call(p.up, CT, n, c, q0 + j);
                                else zc = __ldg(p.logits + ((size_t)n * CT + c) * p.HW + q0 + j);
                                const float pr = sm_prob(zc, p.pix_m[px0 + j], p.pix_s[px0 + j]);
                                const u32 pos = coff[c] - base + cnt[c] + atomicAdd(&curx[c], 1u);
                                S.stageK[pos] = err_key(pr); S.stageV[pos] = v0;
                            }
                        }
                    }
                }
                __syncthreads();
                // ---- write out: a warp per class reserves the tile's slice of the segment and copies it ---------------------------
                for (int c = lo + warp; c < hi; c += NW) {
                    const u32 nc = cnt[c] + cntx[c];
                    if (nc == 0) continue;                 // (warp-uniform)
                    u32 gslot = 0;
                    if (lane == 0) gslot = atomicAdd(p.seg_count + (size_t)g * CT + c, nc);
                    gslot = __shfl_sync(FULL_MASK, gslot, 0);
                    const size_t cslot = ((size_t)g * CT + c) * (size_t)p.cap + gslot;
                    const u32 co = coff[c] - base;
                    for (u32 e = lane; e < nc; e += 32) { p.keysA[cslot + e] = S.stageK[co + e]; p.valsA[cslot + e] = S.stageV[co + e]; }
                }
                lo = hi;
                if (lo < CT) __syncthreads();               // next pass restages
            }
            par ^= 1;
        }
    }
}

// K5b: loss = mean over groups of (mean over kept classes)      reference: mean(), losses/LovaszSoftmax.py:102-120
// `hyb_loss` / `hyb_ovf` (hybrid path): a segment's sum comes from the local kernel unless the segment overflowed and went
// through the LSD fallback
__device__ void loss_finalize(const LovaszParams& p, const double* hyb_loss = nullptr, const u32* hyb_ovf = nullptr) {       // one thread
    float total = 0.f;
    for (int g = 0; g < p.groups; ++g) {
        float acc = 0.f;
        int n = 0;
        for (int c = 0; c < p.C; ++c) {
            const size_t seg = (size_t)g * p.C + c;
            if (!thr_active(p.seg_thr[seg])) continue;
            const bool from_hyb = hyb_loss && !__ldcg(hyb_ovf + seg);
            const float l = (float)__ldcg((from_hyb ? hyb_loss : p.seg_loss) + seg);
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
// K5: Jaccard gradient over the sorted candidates      reference: lovasz_grad, losses/LovaszSoftmax.py:83-95
// --------------------------------------------------------------------------------------------------------------
// One sorted candidate: position i in its segment, F foreground flags among positions 0..i.  Returns err * grad (the loss
// term) and scatters the candidate's gradient g = +-grad * w to the pixel (lovasz_grad: J = 1 - I/U, first difference).
__device__ __forceinline__ double jaccard_element(const LovaszParams& p, u32 key, u32 val, u32 i, u32 F, float gts, float w, int c) {
    const u32 fgi = val & 1u;
    const u32 B = i + 1 - F;                               // background among positions 0..i
    const float J = 1.0f - __fdiv_rn(gts - (float)F, gts + (float)B);
    float grad = J;
    if (i > 0) {
        const float Jp = 1.0f - __fdiv_rn(gts - (float)(F - fgi), gts + (float)(B - (1u - fgi)));
        grad = __fsub_rn(J, Jp);
    }
    const float err = key_err(key);
    if (p.need_grad) {
        // d|fg - p|/dp = -sgn(fg - p): fg -> -1, bg -> +1, exactly 0 when the error is 0
        const float gv = err > 0.f ? (fgi ? -grad : grad) * w : 0.f;
        const u32 px = val >> 1;
        if (fgi) p.gown[px] = gv;
        else {
            u32 ni = __umulhi(px, p.inv_hw);               // floor(px / HW) or one less (inv_hw = floor(2^32 / HW))
            u32 q = px - ni * (u32)p.HW;
            if (q >= (u32)p.HW) { ++ni; q -= (u32)p.HW; }
            p.gbg[((size_t)ni * p.C + c) * (size_t)p.HW + q] = gv;
        }
    }
    return (double)err * (double)grad;
}

__device__ __forceinline__ void jaccard_body(const LovaszParams& p, const SortArgs& a) {
    __shared__ u32 s_wfg[SORT_WARPS];
    __shared__ double s_red[SORT_WARPS];
    __shared__ u32 s_excl;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const u32 le_mask = lane == 31 ? FULL_MASK : ((2u << lane) - 1u);
    const u32 total_tiles = a.tile_start[a.n_seg];
    const u32* keys = a.keys[1];
    const u32* vals = a.vals[1];
    for (u32 t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        const uint4 d4 = a.tile_desc[t];
        const int seg = (int)d4.x;
        if (a.seg_sel && !a.seg_sel[seg]) continue;        // (CTA-uniform; before any barrier of the iteration)
        __syncthreads();
        const u32 off = d4.y, n = d4.z;
        const u32 tis = off / SORT_TILE;
        const u32 tseg0 = t - tis;
        const size_t base = (size_t)seg * a.cap + off;
        const int c = seg % p.C;
        const float gts = (float)p.seg_fg[seg];
        const float w = p.seg_w[seg];

        // foreground counts of the segment's earlier tiles: requested first, summed after the tile's own loads
        u32 f0 = 0, f1 = 0;
        if (warp == 0) {
            if ((u32)lane < tis) f0 = a.tile_fg[tseg0 + lane];
            if ((u32)lane + 32 < tis) f1 = a.tile_fg[tseg0 + lane + 32];
        }
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
        {   // every load of the tile must have landed before its lines are dropped: consume the keys here
            u32 kor = 0;
#pragma unroll
            for (int k = 0; k < SORT_KPT; ++k) kor |= key[k];
            asm volatile("" ::"r"(kor) : "memory");
        }
        __syncthreads();
        {                                                  // the sorted tile is in registers: its lines in buffer B are dead
            discard_dead_lines(keys + base, keys + base + SORT_TILE, keys + (size_t)seg * a.cap, keys + (size_t)(seg + 1) * a.cap);
            discard_dead_lines(vals + base, vals + base + SORT_TILE, vals + (size_t)seg * a.cap, vals + (size_t)(seg + 1) * a.cap);
        }
        u32 wexcl = 0, ttot = 0;
#pragma unroll
        for (int w2 = 0; w2 < SORT_WARPS; ++w2) { const u32 x = s_wfg[w2]; if (w2 < warp) wexcl += x; ttot += x; }
        if (warp == 0) {                                   // foreground flags in the tiles before this one (last sort pass)
            u32 e = f0 + f1;
            for (u32 i = lane + 64; i < tis; i += 32) e += a.tile_fg[tseg0 + i];
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
                acc += jaccard_element(p, key[k], val[k], i, fbase + floc[k], gts, w, c);
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
__global__ void __launch_bounds__(JAC_TPB, 4) jaccard_kernel(LovaszParams p, SortArgs a) {
    jaccard_body(p, a);
    // the CTA that finishes last turns the per-segment sums into the loss (K5b)
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(p.ctrl + CTRL_JTICKET, 1u) == gridDim.x - 1) {
            __threadfence();
            loss_finalize(p);
        }
    }
}

// --------------------------------------------------------------------------------------------------------------
// K4h: hybrid path, local half.  After hyb_partition every segment is grouped into buckets of equal top-w key bits, in
//      bucket order.  A work unit starts every LOC_T0 elements of a segment and owns the buckets that START inside its
//      window [j*T0, (j+1)*T0): it finds them by comparing the digits of neighbouring keys (no bucket table), ranks their
//      elements in shared memory, reads the number of foreground flags in front of its first bucket from the table
//      hyb_count left (fgpre), and goes straight to the Jaccard gradient: units are independent of each other and the
//      sorted order is never materialised, not even in shared memory.
//      Ranking = ONE counting pass: the unit's keys span 2^kbits values (low L bits + the span of its bucket digits);
//      LOC_BINS bins of 2^(kbits-13) values each hold (elements | foreground flags << 16), one exclusive scan gives
//      every bin its first rank and the foreground flags before it, the elements are grouped by bin through an index
//      array, and every element finds its rank inside its bin group (canonical order: key, then pixel index =
//      torch.sort(stable=True)) by looking at the group's other members -- bins hold ~0.3 elements on average.  A unit
//      with a bin group above LOC_TIE_MAX (heavy ties, a dense cluster next to a sparse tail) is sorted by stable LSD
//      passes over value and key bits instead (loc_heavy_unit).
//      reference: torch.sort + lovasz_grad + dot, losses/LovaszSoftmax.py:57-60,83-95
// --------------------------------------------------------------------------------------------------------------
#define LOC_NONE 0xFFFFFFFFu
#define LOC_TIE_MAX 32                                     // longer bin groups take the LSD passes
#define LOC_BIN_BITS 13
#define LOC_BINS (1u << LOC_BIN_BITS)
#define LOC_CNT_STRIDE 260                                 // u16 row stride of the LSD passes: 256 bins + dummy bin, rows 8-byte aligned

struct LocSmem {
    u32 keys[LOC_CAP];                                     // as loaded (partition order); bit 31 = foreground flag
    u32 vals[LOC_CAP];                                     // copied under the counting phases: needed last
    unsigned short order[LOC_CAP];                         // element indices grouped by bin
    u32 bins[LOC_BINS];                                    // (count | fg count << 16) -> exclusive starts -> cursors (LSD passes: u16 counters)
    u32 warp_sum[LOC_WARPS];
    u32 s_first, s_end, heavy, skip;
};

// One stable counting pass over the unit's n elements by digit ((A - sub) >> shift) & 255; B is moved along.
// Element m lives in row m / 32; warp w owns rows [w * rpw, (w + 1) * rpw), so the order (warp, row, lane) is the
// element order and the per-warp counters + a scan across warps give stable positions.  (Rare path: not inlined.)
__device__ __noinline__ void loc_pass(LocSmem& S, u32* A, u32* Bv, u32 n, u32 rpw, u32 amask, u32 sub, u32 shift) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const u32 lt_mask = (1u << lane) - 1;
    unsigned short (*cnt)[LOC_CNT_STRIDE] = reinterpret_cast<unsigned short (*)[LOC_CNT_STRIDE]>(S.bins);
    unsigned short* binexcl = reinterpret_cast<unsigned short*>(S.bins) + LOC_WARPS * LOC_CNT_STRIDE;
    for (u32 i = tid; i < LOC_WARPS * LOC_CNT_STRIDE / 2; i += LOC_TPB) S.bins[i] = 0;
    u32 a[LOC_KPT], rnk[LOC_KPT];
    const u32 wb = warp * rpw * 32 + lane;
#pragma unroll
    for (int k = 0; k < LOC_KPT; ++k) a[k] = ((u32)k < rpw && wb + k * 32 < n) ? A[wb + k * 32] : 0u;
    __syncthreads();                                       // counters cleared; every element of A is in a register
#pragma unroll
    for (int k = 0; k < LOC_KPT; ++k) {
        if ((u32)k < rpw) {                                // (warp-uniform)
            const u32 d = (wb + k * 32 < n) ? ((((a[k] & amask) - sub) >> shift) & 255u) : 256u;
            const u32 m = peer_mask<9>(d);
            const int leader = __ffs(m) - 1;
            u32 old = 0;
            if (lane == leader) { old = cnt[warp][d]; cnt[warp][d] = (unsigned short)(old + __popc(m)); }
            rnk[k] = __shfl_sync(FULL_MASK, old, leader) + __popc(m & lt_mask);
            __syncwarp();
        }
    }
    __syncthreads();
    // thread b < 64 owns bins [4b, 4b+4): four u16 counters travel as one 64-bit word (no carries: totals <= LOC_CAP)
    u64 run = 0;
    if (tid < 64) {
        for (int w2 = 0; w2 < LOC_WARPS; ++w2) {
            u64* c4 = reinterpret_cast<u64*>(&cnt[w2][4 * tid]);
            const u64 c = *c4;
            *c4 = run;
            run += c;
        }
    }
    u32 tot[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) tot[j] = (u32)((run >> (16 * j)) & 0xFFFFu);
    const u32 tsum = tot[0] + tot[1] + tot[2] + tot[3];
    u32 v = tsum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const u32 y = __shfl_up_sync(FULL_MASK, v, o); if (lane >= o) v += y; }
    if (lane == 31 && warp < 2) S.warp_sum[warp] = v;
    __syncthreads();
    if (tid < 64) {
        const u32 e0 = v - tsum + (warp ? S.warp_sum[0] : 0u);
        const u32 e1 = e0 + tot[0], e2 = e1 + tot[1], e3 = e2 + tot[2];
        *reinterpret_cast<uint2*>(&binexcl[4 * tid]) = make_uint2(e0 | (e1 << 16), e2 | (e3 << 16));
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < LOC_KPT; ++k) {
        if ((u32)k < rpw && wb + k * 32 < n) {
            const u32 d = (((a[k] & amask) - sub) >> shift) & 255u;
            rnk[k] = (u32)binexcl[d] + cnt[warp][d] + rnk[k];
            A[rnk[k]] = a[k];
        }
    }
#pragma unroll
    for (int k = 0; k < LOC_KPT; ++k) a[k] = ((u32)k < rpw && wb + k * 32 < n) ? Bv[wb + k * 32] : 0u;
    __syncthreads();                                       // every element of B is in a register
#pragma unroll
    for (int k = 0; k < LOC_KPT; ++k)
        if ((u32)k < rpw && wb + k * 32 < n) Bv[rnk[k]] = a[k];
    __syncthreads();
}

// The rare unit with a long bin group: full stable sort of (K, V) in place (LSD over the value bits, then the key bits),
// then foreground prefix by ballots in sorted order and the gradient.  `fexcl` = foreground flags in front of the unit.
__device__ __noinline__ double loc_heavy_unit(const LovaszParams& p, LocSmem& S, u32* K, u32* V, u32 n, u32 sub,
                                              u32 kbits, u32 valbits, u32 pos0, u32 fexcl, float gts, float w, int c) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const u32 le_mask = lane == 31 ? FULL_MASK : ((2u << lane) - 1u);
    const u32 rpw = ((n + 31) / 32 + LOC_WARPS - 1) / LOC_WARPS;
    for (u32 sh = 0; sh < valbits; sh += 8) loc_pass(S, V, K, n, rpw, 0xFFFFFFFFu, 0u, sh);
    for (u32 sh = 0; sh < kbits; sh += 8) loc_pass(S, K, V, n, rpw, KEY_MASK, sub, sh);
    const u32 wb = warp * rpw * 32 + lane;
    u32 run = 0;
    for (u32 k = 0; k < rpw; ++k) {
        const u32 idx = wb + k * 32;
        run += __popc(__ballot_sync(FULL_MASK, idx < n && (V[idx] & 1u)));
    }
    if (lane == 0) S.warp_sum[warp] = run;
    __syncthreads();
    u32 frun = fexcl;
    for (int w2 = 0; w2 < warp; ++w2) frun += S.warp_sum[w2];
    double acc = 0.0;
    for (u32 k = 0; k < rpw; ++k) {
        const u32 idx = wb + k * 32;
        const u32 v = idx < n ? V[idx] : 0u;
        const u32 bal = __ballot_sync(FULL_MASK, idx < n && (v & 1u));
        if (idx < n) acc += jaccard_element(p, K[idx], v, pos0 + idx, frun + __popc(bal & le_mask), gts, w, c);
        frun += __popc(bal);
    }
    return acc;
}

__global__ void __launch_bounds__(LOC_TPB, 2) hyb_local_kernel(LovaszParams p, SortArgs a, HybArgs h, u32 max_tiles) {
    pdl_enter();
    extern __shared__ __align__(16) unsigned char loc_smem_raw[];
    LocSmem& S = *reinterpret_cast<LocSmem*>(loc_smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const u32 total_tiles = min(a.tile_start[a.n_seg], max_tiles);
    const u32 n_work = total_tiles * LOC_PER_SORT_TILE;
    const u32* __restrict__ gkeys = a.keys[1];
    const u32* __restrict__ gvals = a.vals[1];
    const u32 valbits = 32u - (u32)__clz((int)(2u * (u32)p.P - 1u));     // bits of the largest value (pixel << 1 | fg)
    uint4 d4_next = blockIdx.x < n_work ? a.tile_desc[blockIdx.x / LOC_PER_SORT_TILE] : make_uint4(0, 0, 0, 0);
    for (u32 lt = blockIdx.x; lt < n_work; lt += gridDim.x) {
        const uint4 d4 = d4_next;
        if (lt + gridDim.x < n_work) d4_next = a.tile_desc[(lt + gridDim.x) / LOC_PER_SORT_TILE];   // one unit ahead
        const int seg = (int)d4.x;
        const u32 p0 = d4.y + (lt % LOC_PER_SORT_TILE) * LOC_T0;
        const u32 ns = a.seg_count[seg];
        if (p0 >= ns) continue;                            // (CTA-uniform) no such unit
        const float gts = (float)p.seg_fg[seg];
        const float w = p.seg_w[seg];
        __syncthreads();                                   // previous unit done with the shared state
        // (the give-up flag of the segment can be raised by another CTA at any time: one thread reads it, all follow that reading)
        if (tid == 0) { S.s_first = LOC_NONE; S.s_end = LOC_NONE; S.heavy = 0; S.skip = ld_relaxed(h.seg_ovf + seg); }
        {   // the bins of the counting pass are cleared here: the barriers of the staging loop below cover it
            uint4* z = reinterpret_cast<uint4*>(S.bins);
            for (u32 i = tid; i < LOC_BINS / 4; i += LOC_TPB) z[i] = make_uint4(0, 0, 0, 0);
        }
        const u32 L = hyb_plan(a.seg_bits[seg], ns).L;
        const size_t sbase = (size_t)seg * a.cap + p0;
        const u32 avail = min(ns - p0, (u32)LOC_CAP);

        // ---- the unit's buckets: first digit change at or after p0, first digit change at or after p0 + T0 ----------
        // (keys and values travel global -> shared as 16-byte cp.async copies: no load -> store round trip per element)
        const bool al16 = ((sbase & 3) == 0) && ((a.cap & 3) == 0) && (((uintptr_t)gkeys & 15) == 0);
        u32 prevkey = 0;
        if (tid == 0 && p0 > 0) prevkey = gkeys[sbase - 1];
        u32 loaded = 0, checked = 0, target = min(avail, (u32)(LOC_T0 + 1024));
        for (;;) {
            if (al16) {                                    // (whole 16-byte chunks: may read up to 3 elements past `target`, inside the segment)
                for (u32 ch = loaded / 4 + tid; ch < (target + 3) / 4; ch += LOC_TPB) cp_async<16>(&S.keys[4 * ch], gkeys + sbase + 4 * ch);
            } else {
                for (u32 i = loaded + tid; i < target; i += LOC_TPB) cp_async<4>(&S.keys[i], gkeys + sbase + i);
            }
            cp_async_commit();
            cp_async_wait<0>();
            __syncthreads();
            if (S.skip) break;                             // (CTA-uniform) segment already given up
            for (u32 i = checked + tid; i < target; i += LOC_TPB) {
                bool bnd = p0 + i == 0;
                if (!bnd) {
                    const u32 prev = i > 0 ? S.keys[i - 1] : prevkey;      // (i == 0: thread 0, which holds the predecessor)
                    bnd = ((S.keys[i] & KEY_MASK) >> L) != ((prev & KEY_MASK) >> L);
                }
                if (bnd) atomicMin(i < LOC_T0 ? &S.s_first : &S.s_end, i);
            }
            __syncthreads();
            checked = target;
            loaded = al16 ? ((target + 3) & ~3u) : target;
            if (S.s_end != LOC_NONE || S.s_first == LOC_NONE || checked == avail) break;
            target = avail;
        }
        const u32 s_rel = S.s_first;
        if (S.skip || s_rel == LOC_NONE) continue;         // segment given up / the window lies inside a bucket of an earlier unit
        u32 e_rel = S.s_end;
        if (e_rel == LOC_NONE) {
            if (p0 + avail == ns) e_rel = avail;           // the segment ends here
            else {                                         // a bucket too long for shared memory: LSD fallback for the segment
                if (tid == 0) { atomicExch(h.seg_ovf + seg, 1u); atomicExch(h.ovf_any, 1u); }
                continue;
            }
        }
        const u32 n = e_rel - s_rel;
        u32* K = S.keys + s_rel;
        const u32* __restrict__ gv = gvals + sbase + s_rel;    // the unit's values, in the order of K
        u32* V = S.vals + s_rel;
        // the values are needed last (tie order, the pixel a gradient goes to): their copy runs under the counting phases
        if (al16 && (((uintptr_t)gvals & 15) == 0)) {
            for (u32 ch = s_rel / 4 + tid; ch < (e_rel + 3) / 4; ch += LOC_TPB) cp_async<16>(&S.vals[4 * ch], gvals + sbase + 4 * ch);
        } else {
            for (u32 i = tid; i < n; i += LOC_TPB) cp_async<4>(&V[i], gv + i);
        }
        cp_async_commit();
        // ---- counting pass over the span of the unit's keys ---------------------------------------------------------------------
        const u32 dmin = (K[0] & KEY_MASK) >> L, dmax = (K[n - 1] & KEY_MASK) >> L;     // buckets lie in digit order
        const u32 sub = dmin << L;
        const u32 kbits = L + (dmax > dmin ? 32u - (u32)__clz((int)(dmax - dmin)) : 0u);
        const u32 bsh = kbits > LOC_BIN_BITS ? kbits - LOC_BIN_BITS : 0u;
        const u32 fexcl = __ldcg(h.fgpre + (size_t)seg * HYB_MAX_BINS + dmin);      // foreground flags in front of the unit
        for (u32 i = tid; i < n; i += LOC_TPB) { const u32 k = K[i]; atomicAdd(&S.bins[((k & KEY_MASK) - sub) >> bsh], 1u + ((k >> 31) << 16)); }
        __syncthreads();
        {   // exclusive scan of both halves at once (totals <= LOC_CAP: no carry); thread t owns bins [16t, 16t + 16)
            uint4* b4 = reinterpret_cast<uint4*>(S.bins) + 4 * tid;
            uint4 c[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) c[j] = b4[j];
            u32 sum = 0, big = 0;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                sum += c[j].x + c[j].y + c[j].z + c[j].w;
                big |= ((c[j].x & 0xFFFFu) > LOC_TIE_MAX) | ((c[j].y & 0xFFFFu) > LOC_TIE_MAX) | ((c[j].z & 0xFFFFu) > LOC_TIE_MAX) |
                       ((c[j].w & 0xFFFFu) > LOC_TIE_MAX);
            }
            if (big) S.heavy = 1;
            u32 v = sum;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const u32 y = __shfl_up_sync(FULL_MASK, v, o); if (lane >= o) v += y; }
            if (lane == 31) S.warp_sum[warp] = v;
            __syncthreads();
            u32 run = v - sum;
#pragma unroll
            for (int w2 = 0; w2 < LOC_WARPS; ++w2) if (w2 < warp) run += S.warp_sum[w2];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                uint4 o4;
                o4.x = run; run += c[j].x;
                o4.y = run; run += c[j].y;
                o4.z = run; run += c[j].z;
                o4.w = run; run += c[j].w;
                b4[j] = o4;
            }
        }
        __syncthreads();
        const bool heavy = S.heavy != 0;
        if (!heavy)                                        // group the elements by bin (arrival order inside a bin)
            for (u32 i = tid; i < n; i += LOC_TPB)
                S.order[atomicAdd(&S.bins[((K[i] & KEY_MASK) - sub) >> bsh], 1u) & 0xFFFFu] = (unsigned short)i;
        cp_async_wait<0>();                                // the values have landed
        __syncthreads();
        const int c = seg % p.C;
        const u32 pos0 = p0 + s_rel;                       // position of the unit's first element in its segment
        double acc = 0.0;
        if (heavy) acc = loc_heavy_unit(p, S, K, V, n, sub, kbits, valbits, pos0, fexcl, gts, w, c);     // (CTA-uniform)
        else {
            // ---- rank inside the bin group, Jaccard gradient ------------------------------------------------------------------
            for (u32 i = tid; i < n; i += LOC_TPB) {
                const u32 v = V[i];
                const u32 kf = K[i], k = kf & KEY_MASK;
                const u32 b = (k - sub) >> bsh;
                const u32 lo = b ? S.bins[b - 1] : 0u;     // bins[b - 1]: (end of bin b-1 = start of b) | fg flags before bin b-1 ...
                const u32 hi = S.bins[b];                  // ... and bins[b]: end of b | fg flags before bin b
                const u32 start = lo & 0xFFFFu, stop = hi & 0xFFFFu;
                u32 rank = start, F = fexcl + (hi >> 16) + (kf >> 31);
                for (u32 j = start; j < stop; ++j) {       // the other members of the bin group: (key, value) order
                    const u32 o = S.order[j];
                    const u32 kf2 = K[o], k2 = kf2 & KEY_MASK;
                    const bool less = k2 < k || (k2 == k && V[o] < v);
                    rank += less;
                    F += less ? (kf2 >> 31) : 0u;
                }
                acc += jaccard_element(p, k, v, pos0 + rank, F, gts, w, c);
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(FULL_MASK, acc, o);
        if (lane == 0 && acc != 0.0) atomicAdd(h.seg_loss + seg, acc);
    }
    __syncthreads();
    // the CTA that finishes last turns the per-segment sums into the loss, unless a segment needs the fallback
    if (tid == 0) {
        __threadfence();
        if (atomicAdd(h.ticket + 1, 1u) == gridDim.x - 1) {
            __threadfence();
            if (!ld_relaxed(h.ovf_any)) loss_finalize(p, h.seg_loss, h.seg_ovf);
        }
    }
}

// Fallback of the hybrid path: the segments the local kernel gave up on go through the three-pass LSD sort and the
// Jaccard kernel of the plain path, fused into one cooperative launch (phases separated by grid barriers) that returns
// at once when no segment overflowed -- the usual case.
__global__ void __launch_bounds__(SORT_TPB, 2) sort_fallback_kernel(LovaszParams p, SortArgs a, HybArgs h, u32 max_tiles) {
    if (!ld_relaxed(h.ovf_any)) return;                    // (grid-uniform: written by an earlier launch)
    const int n_passes = a.val_passes + SORT_PASSES;
    u32* bar = a.gbar + 2 * SORT_MAX_PASSES;               // (the first 2 * SORT_MAX_PASSES counters belong to sort_big_scan)
    for (int pass = 0; pass < n_passes; ++pass) {
        sort_count_body(a, pass, max_tiles);
        grid_barrier(bar + 2 * pass, a.status);
        if (pass < a.val_passes) sort_scatter_body<false, true>(a, pass, max_tiles);
        else sort_scatter_body<false, false>(a, pass, max_tiles);
        grid_barrier(bar + 2 * pass + 1, a.status);
    }
    sort_fg_count_body(a, max_tiles);
    grid_barrier(bar + 2 * SORT_MAX_PASSES, a.status);
    jaccard_body(p, a);
    grid_barrier(bar + 2 * SORT_MAX_PASSES + 1, a.status);
    if (blockIdx.x == 0 && threadIdx.x == 0) loss_finalize(p, h.seg_loss, h.seg_ovf);
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

// Pipelined backward (the fast path): lane l of a warp owns pixel l of a 32-pixel warp tile.  Everything a tile needs
// arrives through the warp-private cp.async ring one tile ahead of the math -- the C logit rows, the per-pixel state
// (softmax max / denominator, own-class gradient, compact label), the first two background-candidate gradients
// (addressed through the candidate mask, which therefore travels two tiles ahead) -- and completion is tracked with
// cp.async groups, so no global-load scoreboard sits on the critical path.
template <int CT>
struct BwdStage {
    float z[CT][32];
    float m[32], s[32], gl[32];
    u32 maskn[32];                                        // candidate mask of the tile AFTER this one
    float g1[32], g2[32];
    unsigned char lab8[32];
};
template <int CT, int TPB>
__global__ void __launch_bounds__(TPB) backward_kernel_async(LovaszParams p, const float* __restrict__ go,
                                                             const float* __restrict__ go_ce,
                                                             float* __restrict__ dlogits) {
    using W = WarpTile<CT, 1>;
    using Stage = BwdStage<CT>;
    constexpr int WT = W::WT, NW = TPB / 32, STAGES = 2;
    static_assert(sizeof(Stage) % 16 == 0, "stage alignment");
    extern __shared__ __align__(16) unsigned char pipe_smem_raw[];
    __shared__ float s_thr[NW][B200SEG_MAX_CLASSES];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    Stage* S = reinterpret_cast<Stage*>(pipe_smem_raw) + (size_t)warp * STAGES;
    const u32 wtpi = (u32)((p.HW + WT - 1) / WT);         // tile counts fit 32 bits: keep the index math off the 64-bit divider
    const u32 nwt = wtpi * (u32)p.N;
    const u32 gw = blockIdx.x * NW + warp, nwarps = gridDim.x * NW;
    const u32 step = p.interleave ? nwarps : 1u;         // see stats_kernel_async
    const u32 t0 = p.interleave ? gw : (u32)((u64)nwt * gw / nwarps);
    const u32 t1 = p.interleave ? nwt : (u32)((u64)nwt * (gw + 1) / nwarps);
    if (t0 >= t1) return;
    const float gsc = __ldg(go);
    // fused cross-entropy term: d/dz_k = go_ce * (p_k - [k == label]) / n_valid on the non-ignored pixels
    const float gce = (p.ce_enabled && go_ce) ? __ldg(go_ce) * *p.ce_inv_n : 0.f;
    const size_t plane = (size_t)p.HW;

    // (image, tile-in-image) cursors advance by increments: no integer division per tile
    u32 cn = t0 / wtpi, cti = t0 - cn * wtpi;             // tile being consumed
    u32 n1 = cn, ti1 = cti, pt1 = t0;                      // next tile whose state / logits are requested
    u32 n2 = cn, ti2 = cti + step, pt2 = t0 + step;        // next tile whose candidate mask is requested
    while (ti2 >= wtpi) { ti2 -= wtpi; ++n2; }
    const int sub = lane & 7, strm = lane >> 3;            // 16-byte chunk / stream of the combined state copy
    // request tile pt1 (its mask is `mask1`) into stage `st`, and the mask of tile pt2 = pt1 + 1 along with it
    auto request = [&](Stage& st, u32 mask1) {
        if (pt1 < t1) {
            const long long q0 = (long long)ti1 * WT;
            W::prefetch(st.z, p.logits, (int)n1, q0, p.HW, lane);
            const size_t px0 = (size_t)n1 * plane + q0;
            if (strm < 3) {
                if (q0 + 4 * sub < p.HW) {
                    const float* src = strm == 0 ? p.pix_m : (strm == 1 ? p.pix_s : p.gown);
                    float* dst = strm == 0 ? st.m : (strm == 1 ? st.s : st.gl);
                    cp_async<16>(dst + 4 * sub, src + px0 + 4 * sub);
                }
            } else if (pt2 < t1) {
                const long long q2 = (long long)ti2 * WT + 4 * sub;
                if (q2 < p.HW) cp_async<16>(st.maskn + 4 * sub, p.cmask + (size_t)n2 * plane + q2);
            }
            if (lane < 2 && q0 + 16 * lane < p.HW) cp_async<16>(st.lab8 + 16 * lane, p.lab8 + px0 + 16 * lane);
            if (mask1 && q0 + lane < p.HW) {
                const float* gb = p.gbg + (size_t)n1 * CT * plane + q0 + lane;
                cp_async<4>(st.g1 + lane, gb + (size_t)(__ffs(mask1) - 1) * plane);
                const u32 rest = mask1 & (mask1 - 1);
                if (rest) cp_async<4>(st.g2 + lane, gb + (size_t)(__ffs(rest) - 1) * plane);
            }
        }
        cp_async_commit();
        pt1 += step; pt2 += step; ti1 += step; ti2 += step;
        while (ti1 >= wtpi) { ti1 -= wtpi; ++n1; }
        while (ti2 >= wtpi) { ti2 -= wtpi; ++n2; }
    };
    u32 mask = 0;                                          // candidate mask of the tile being consumed
    {
        const long long q = (long long)cti * WT + lane;
        if (q < p.HW) mask = p.cmask[(size_t)cn * plane + q];
    }
    request(S[0], mask);

    int cur_g = -1;
    u32 it = 0;
    for (u32 t = t0; t < t1; t += step, ++it) {
        Stage& C = S[it & 1];
        cp_async_wait<0>();                               // this lane's copies for tile t have landed ...
        __syncwarp();                                     // ... so have the other lanes', and nobody still reads the other stage
        const u32 nmask = (t + step < t1) ? C.maskn[lane] : 0u;
        request(S[(it & 1) ^ 1], nmask);
        const int n = (int)cn;
        const long long q = (long long)cti * WT + lane;
        cti += step;
        while (cti >= wtpi) { cti -= wtpi; ++cn; }
        const int g = p.per_image ? n : 0;
        if (g != cur_g) {
            __syncwarp();
            if (lane < CT) s_thr[warp][lane] = p.seg_thr[(size_t)g * CT + lane];
            cur_g = g;
            __syncwarp();
        }
        const u32 cmask_cur = mask;
        mask = nmask;
        if (q >= p.HW) continue;
        const float (*T)[WT] = C.z;
        const float m = C.m[lane], s = C.s[lane], gl = C.gl[lane];
        const u32 l8 = C.lab8[lane];
        const size_t off = (size_t)n * CT * plane + q;
        const float* gb = p.gbg + off;
        float* dp = dlogits + off;
        const bool filt = l8 == LAB8_FILTERED;
        int lab = l8 < (u32)CT ? (int)l8 : -1;
        const float gcp = (lab >= 0 && !(p.has_ce_ignore && lab == p.ce_ignore)) ? gce : 0.f;   // CE weight of this pixel
        const int ce_lab = lab;
        if (lab >= 0 && !thr_active(s_thr[warp][lab])) lab = -1;   // class not summed: no own-class term
        // dot = sum_j g_j p_j over the pixel's candidates (background candidates ascending, then the own class)
        float d = 0.f, pk1 = 0.f, pk2 = 0.f, g1 = 0.f, g2 = 0.f;
        // candidates beyond the two the ring prefetched (trained-like logits: ~7 per pixel): their gradients are requested
        // together, so the pixel pays one round trip instead of one per candidate
        constexpr int NX = 6;
        float gx[NX];
#pragma unroll
        for (int j = 0; j < NX; ++j) gx[j] = 0.f;
        if (cmask_cur) {
            g1 = C.g1[lane]; g2 = C.g2[lane];
            u32 rest = cmask_cur & (cmask_cur - 1);
            rest &= rest - 1;                              // without the first two
#pragma unroll
            for (int j = 0; j < NX; ++j) {
                if (rest) { gx[j] = gb[(size_t)(__ffs(rest) - 1) * plane]; rest &= rest - 1; }
            }
            u32 mm = cmask_cur;
            int i = 0;
            while (mm) {
                const int c = __ffs(mm) - 1;
                mm &= mm - 1;
                const float pr = sm_prob(T[c][lane], m, s);
                float gk;
                if (i == 0) { gk = g1; pk1 = pr; } else if (i == 1) { gk = g2; pk2 = pr; }
                else if (i < 2 + NX) {
                    gk = gx[0];
#pragma unroll
                    for (int j = 1; j < NX; ++j) gk = (i == 2 + j) ? gx[j] : gk;
                } else gk = gb[(size_t)c * plane];
                d += gk * pr;
                ++i;
            }
        }
        float ownv = 0.f;
        if (lab >= 0) {
            const float pown = sm_prob(T[lab][lane], m, s);
            d += gl * pown;
            ownv = gsc * pown * (gl - d) + gcp * (pown - 1.0f);     // exact value of the own class
        }
        const float inv_s = __fdiv_rn(1.0f, s);
        const float nd = (filt ? 0.f : -gsc * d * inv_s) + gcp * inv_s;
        // every class gets -go * p_k * dot with the fast exponential, the own class takes its exact value ...
#pragma unroll
        for (int c = 0; c < CT; ++c) {
            const float fast = nd * __expf(T[c][lane] - m);
            dp[(size_t)c * plane] = (c == lab) ? ownv : fast;
        }
        // ... then the (few) background-candidate classes are overwritten with the exact go * p_k * (g_k - dot)
        if (cmask_cur) {
            u32 mm = cmask_cur;
            int i = 0;
            while (mm) {
                const int c = __ffs(mm) - 1;
                mm &= mm - 1;
                float gk, pk;
                if (i == 0) { gk = g1; pk = pk1; } else if (i == 1) { gk = g2; pk = pk2; }
                else {
                    pk = sm_prob(T[c][lane], m, s);
                    if (i < 2 + NX) {
                        gk = gx[0];
#pragma unroll
                        for (int j = 1; j < NX; ++j) gk = (i == 2 + j) ? gx[j] : gk;
                    } else gk = gb[(size_t)c * plane];
                }
                dp[(size_t)c * plane] = gsc * pk * (gk - d) + gcp * pk;
                ++i;
            }
        }
        // cross-entropy only: the label's class is not summed by the Lovasz term, its "- 1" was not applied above
        if (gcp != 0.f && lab < 0) dp[(size_t)ce_lab * plane] = nd * __expf(T[ce_lab][lane] - m) - gcp;
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

// Enqueue the hybrid path: prepare, bucket histogram, partition, local sort + Jaccard, fallback (a no-op unless a
// segment overflowed).  All decisions are taken on the device.
#include "lovasz_up.cuh"

static int hybrid_enqueue(const LovaszParams& p, const SortArgs& a, const HybArgs& h, const SortScratch& L, cudaStream_t st) {
    static bool attr_set[64] = {false};
    static int fb_occ[64] = {0};
    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    const bool cached = dev >= 0 && dev < 64;
    if (!cached || !attr_set[dev]) {
        CUDA_TRY(cudaFuncSetAttribute(hyb_count_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(2 * HYB_MAX_BINS * sizeof(u32))));
        CUDA_TRY(cudaFuncSetAttribute(hyb_partition_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(PartSmem)));
        CUDA_TRY(cudaFuncSetAttribute(hyb_local_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(LocSmem)));
        CUDA_TRY(cudaFuncSetAttribute(sort_fallback_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ScatterSmem)));
        int occ = 0;
        CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, sort_fallback_kernel, SORT_TPB, sizeof(ScatterSmem)));
        if (cached) { fb_occ[dev] = occ < 1 ? 1 : occ; attr_set[dev] = true; }
    }
    const int sms = b200seg_sm_count();
    {
        const int pgrid = a.n_seg < sms ? a.n_seg : sms;     // CTA 0 plans the tiles, all of them clear the bucket tables
        CUDA_TRY(launch_chained(sort_prepare_kernel, dim3(pgrid), dim3(SORT_PREP_TPB), 0, st, a, h, L.max_tiles));
        LAUNCH_CHECK("sort_prepare_kernel");
    }
    b200seg_stage(4, st);
    const u32 cgrid = L.max_tiles < (u32)sms * 3 ? L.max_tiles : (u32)sms * 3;   // 3 CTAs per SM by shared memory; 1x / 2x measured slower
    CUDA_TRY(launch_chained(hyb_count_kernel, dim3(cgrid), dim3(SORT_TPB), 2 * HYB_MAX_BINS * sizeof(u32), st, a, h, L.max_tiles));
    LAUNCH_CHECK("hyb_count_kernel");
    b200seg_stage(5, st);
    const u32 pgrid2 = L.max_tiles < (u32)sms * 3 ? L.max_tiles : (u32)sms * 3;
    CUDA_TRY(launch_chained(hyb_partition_kernel, dim3(pgrid2), dim3(SORT_TPB), sizeof(PartSmem), st, a, h, L.max_tiles));
    LAUNCH_CHECK("hyb_partition_kernel");
    b200seg_stage(6, st);
    if (p.dbg & 128) return 0;                             // (debugging: stop after the partition)
    const u32 units = L.max_tiles * LOC_PER_SORT_TILE;
    const u32 lgrid = units < (u32)sms * 2 ? units : (u32)sms * 2;
    CUDA_TRY(launch_chained(hyb_local_kernel, dim3(lgrid), dim3(LOC_TPB), sizeof(LocSmem), st, p, a, h, L.max_tiles));
    LAUNCH_CHECK("hyb_local_kernel");
    b200seg_stage(7, st);
    {   // cooperative: the grid barriers between its phases need the whole grid resident
        SortArgs a_fb = a;
        a_fb.seg_sel = h.seg_ovf;
        LovaszParams p_copy = p;
        HybArgs h_copy = h;
        u32 bound = L.max_tiles;
        const u32 per_sm = cached ? (u32)fb_occ[dev] : 1u;
        const u32 fgrid = L.max_tiles < (u32)sms * per_sm ? L.max_tiles : (u32)sms * per_sm;
        void* args[] = {(void*)&p_copy, (void*)&a_fb, (void*)&h_copy, (void*)&bound};
        CUDA_TRY(cudaLaunchCooperativeKernel((const void*)sort_fallback_kernel, dim3(fgrid), dim3(SORT_TPB), args,
                                             sizeof(ScatterSmem), st));
    }
    return 0;
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
    p.up = UpSrc{nullptr, 0, 0, 0, 0, 0.f, 0.f, 1};
    p.N = n; p.C = c; p.HW = hw; p.P = (long long)n * hw;
    p.inv_hw = hw >= 2 ? (u32)((1ull << 32) / (unsigned long long)hw) : 0xFFFFFFFFu;
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
    p.grp_tmin = (float*)(ws + L.grp_tmin); p.seg_order = (unsigned char*)(ws + L.seg_order); p.have_records = 0; p.emit_force = 0;
    p.pix_m = (float*)(ws + L.pix_m); p.pix_s = (float*)(ws + L.pix_s); p.gown = (float*)(ws + L.gown);
    p.lab8 = (unsigned char*)(ws + L.lab8); p.cmask = (u32*)(ws + L.cmask);
    p.rec16 = (uint4*)(ws + L.rec16); p.rec4 = (u32*)(ws + L.rec4);
    p.flags = p.ctrl + CTRL_FLAGS;
    p.geo = reinterpret_cast<EmitGeomDev*>(p.ctrl + CTRL_GEO);
    p.ce_enabled = 0; p.has_ce_ignore = 0; p.ce_ignore = 0; p.ce_out = nullptr;
    p.ce_sum = reinterpret_cast<double*>(p.ctrl + CTRL_CE_SUM); p.ce_cnt = p.ctrl + CTRL_CE_CNT;
    p.ce_inv_n = reinterpret_cast<float*>(p.ctrl + CTRL_CE_INV_N);
    p.geo_stream = EmitGeomDev{0, 0}; p.geo_rec = EmitGeomDev{0, 0};
    p.keysA = (u32*)(ws + L.keysA); p.valsA = (u32*)(ws + L.valsA);
    p.keysB = (u32*)(ws + L.keysB); p.valsB = (u32*)(ws + L.valsB);
    p.gbg = (float*)(ws + L.gbg);
    p.cm = nullptr; p.has_drop = 0; p.drop = 0;
    p.status = (int*)(p.ctrl + CTRL_STATUS);
    p.loss_out = nullptr; p.need_grad = 1; p.dbg = 0;
    p.interleave = b200seg_tuning().interleave;
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

// the pipelined kernels (the only ones that carry the fused cross-entropy term) cover this call
static bool pipelined_ok(const float* logits, const void* labels, int label_dtype, int32_t c, int64_t hw) {
    return vec4_ok(logits, labels, label_dtype, hw) && (c == 8 || c == 17 || c == 25) && hw % 16 == 0 && aligned16(labels) &&
           b200seg_tuning().stats_variant != 1;
}

extern "C" int b200seg_lovasz_ce_supported(const float* logits, const void* labels, int32_t label_dtype, int32_t n,
                                           int32_t c, int64_t hw, const float* dlogits) {
    if (check_shape(n, c, hw) != 0 || (long long)n * hw == 0) return 0;
    return pipelined_ok(logits, labels, label_dtype, c, hw) && (!dlogits || aligned16(dlogits)) ? 1 : 0;
}

static int lovasz_forward_impl(const float* logits, const void* labels, int32_t label_dtype, int32_t n,
                               int32_t c, int64_t hw, int32_t per_image, int64_t filter_label,
                               int32_t keep_absent, uint32_t class_mask, int32_t need_grad, void* workspace,
                               size_t workspace_bytes, float* loss_out, int64_t* cm, int64_t cm_drop_label,
                               int32_t* status, void* stream, bool ce_enabled, int64_t ce_ignore, float* ce_out,
                               const UpSrc* up = nullptr) {
    if (int rc = check_shape(n, c, hw)) return rc;
    if (!loss_out) { b200seg_set_error("loss_out is NULL"); return B200SEG_E_INVALID; }
    if ((long long)n * hw == 0) {      // empty batch: nothing to read, loss 0 (cross entropy of nothing: NaN, like torch)
        CUDA_TRY(cudaMemsetAsync(loss_out, 0, sizeof(float), (cudaStream_t)stream));
        if (ce_enabled) CUDA_TRY(cudaMemsetAsync(ce_out, 0xFF, sizeof(float), (cudaStream_t)stream));
        return 0;
    }
    if (ce_enabled) {
        if (!ce_out || !status) { b200seg_set_error("ce_out / status is NULL"); return B200SEG_E_INVALID; }
        if (!up && !pipelined_ok(logits, labels, label_dtype, c, hw)) {
            b200seg_set_error("the fused cross-entropy term needs the pipelined kernels (C in {8,17,25}, plane %% 16 == 0, "
                              "16-byte aligned tensors): ask b200seg_lovasz_ce_supported first");
            return B200SEG_E_UNSUPPORTED;
        }
        // with classes_to_ignore the compact label loses the class of the removed pixels: only the usual case is fused
        if (filter_label != B200SEG_NO_LABEL && filter_label >= 0 && filter_label < c && filter_label != ce_ignore) {
            b200seg_set_error("fused cross entropy: classes_to_ignore names a real class that the cross entropy keeps");
            return B200SEG_E_UNSUPPORTED;
        }
    }
    if ((!logits && !up) || !labels || !workspace || (cm && !status)) {
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
    if (up) p.up = *up;
    p.loss_out = loss_out;
    p.need_grad = need_grad ? 1 : 0;
    p.dbg = b200seg_tuning().dbg;
    p.ce_enabled = ce_enabled ? 1 : 0;
    p.has_ce_ignore = (ce_enabled && ce_ignore >= INT_MIN && ce_ignore <= INT_MAX) ? 1 : 0;
    p.ce_ignore = p.has_ce_ignore ? (int)ce_ignore : 0;
    p.ce_out = ce_out;
    p.cm = (unsigned long long*)cm;
    p.has_drop = (cm_drop_label != B200SEG_NO_LABEL && cm_drop_label >= INT_MIN && cm_drop_label <= INT_MAX) ? 1 : 0;
    p.drop = p.has_drop ? (int)cm_drop_label : 0;
    if (status) p.status = status;

    CUDA_TRY(cudaMemsetAsync(ws + L.ctrl, 0, L.zero_end - L.ctrl, st));
    const int sms = b200seg_sm_count();
    const bool v4 = up ? true : vec4_ok(logits, labels, label_dtype, hw);   // (fused upsampling: W % 32 == 0, labels read one by one)
    b200seg_stage(0, st);

    // K1
    const bool known_c = (c == 8 || c == 17 || c == 25);
    const int stats_variant = b200seg_tuning().stats_variant, emit_force = b200seg_tuning().emit_path;
    const bool async_ok = v4 && known_c && hw % 16 == 0 && aligned16(labels);
    if (up) {
        p.have_records = 1;
        p.emit_force = emit_force;
#define LAUNCH_STATS_UP(CC)                                                                                       \
    {                                                                                                             \
        constexpr int TT = 128;                                                                                   \
        const size_t smem = (size_t)(TT / 32) * 3 * CC * 32 * 4;                                                  \
        CUDA_TRY(cudaFuncSetAttribute(stats_kernel_up<CC, TT, LT>, cudaFuncAttributeMaxDynamicSharedMemorySize,   \
                                      (int)smem));                                                                \
        int per_sm = (int)((224 * 1024) / (smem + 6 * 1024));                                                     \
        if (per_sm * TT > 2048) per_sm = 2048 / TT;                                                               \
        stats_kernel_up<CC, TT, LT><<<sms * per_sm, TT, smem, st>>>(p);                                           \
    }
        DISPATCH_LABEL(label_dtype, {
            if (c == 8) LAUNCH_STATS_UP(8)
            else if (c == 17) LAUNCH_STATS_UP(17)
            else LAUNCH_STATS_UP(25)
        });
#undef LAUNCH_STATS_UP
    } else if (async_ok && stats_variant != 1) {
        p.have_records = 1;
        p.emit_force = emit_force;
#define LAUNCH_STATS_ASYNC(CC, TT, SS)                                                                          \
    {                                                                                                           \
        const size_t smem = (size_t)(TT / 32) * SS * (CC * 32 * 4 + 256);                                       \
        CUDA_TRY(cudaFuncSetAttribute(stats_kernel_async<CC, TT, SS, LT>,                                       \
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));                 \
        int per_sm = (int)((224 * 1024) / (smem + 6 * 1024));                                                   \
        if (per_sm < 1) per_sm = 1;                                                                             \
        if (per_sm * TT > 2048) per_sm = 2048 / TT;                                                             \
        stats_kernel_async<CC, TT, SS, LT><<<sms * per_sm, TT, smem, st>>>(p);                                  \
    }
#define LAUNCH_STATS_C(TT, SS)                                         \
    {                                                                  \
        if (c == 8) LAUNCH_STATS_ASYNC(8, TT, SS)                      \
        else if (c == 17) LAUNCH_STATS_ASYNC(17, TT, SS)               \
        else LAUNCH_STATS_ASYNC(25, TT, SS)                            \
    }
        if (stats_variant == 7) {                          // the TMA-fed variant (see stats_kernel_tma)
            CUtensorMap tmap;
            if (int rc = make_logits_tensor_map(&tmap, logits, n, c, hw)) return rc;
#define LAUNCH_STATS_TMA(CC)                                                                                     \
    {                                                                                                           \
        const size_t stage = ((size_t)CC * TMA_BOX * 4 + TMA_BOX * sizeof(LT) + 127) / 128 * 128;               \
        const size_t smem = stage * TMA_STAGES + 128;                                                           \
        CUDA_TRY(cudaFuncSetAttribute(stats_kernel_tma<CC, LT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        int per_sm = (int)((224 * 1024) / (smem + 8 * 1024));                                                   \
        if (per_sm < 1) per_sm = 1;                                                                             \
        stats_kernel_tma<CC, LT><<<sms * per_sm, TMA_TPB, smem, st>>>(p, tmap);                                 \
    }
            DISPATCH_LABEL(label_dtype, {
                if (c == 8) LAUNCH_STATS_TMA(8)
                else if (c == 17) LAUNCH_STATS_TMA(17)
                else LAUNCH_STATS_TMA(25)
            });
#undef LAUNCH_STATS_TMA
        } else
        DISPATCH_LABEL(label_dtype, {
            switch (stats_variant) {
                // measured (C=25, 8x540x960, in the stream): <384,2> 149 us, <256,2> 152, <512,2> 153, <128,2> 161, <128,3> 169
                case 2: LAUNCH_STATS_C(128, 2) break;
                case 3: LAUNCH_STATS_C(128, 4) break;
                case 4: LAUNCH_STATS_C(256, 2) break;
                case 5: LAUNCH_STATS_C(128, 3) break;
                case 6: LAUNCH_STATS_C(512, 2) break;
                default: LAUNCH_STATS_C(384, 2) break;
            }
        });
#undef LAUNCH_STATS_C
#undef LAUNCH_STATS_ASYNC
    } else if (v4 && known_c) {
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
    if (cm) b200seg_cm_ready(st);                          // the fused confusion matrix and the status word are complete
    b200seg_stage(1, st);
    if (p.keep_absent) {
        DISPATCH_LABEL(label_dtype, absent_max_kernel<LT><<<dim3(p.n_seg, 32), 256, 0, st>>>(p));
        LAUNCH_CHECK("absent_max_kernel");
    }
    // emission geometries (the kernels read the one in force from device memory: the path is chosen on the device)
    const bool pipe_ok = v4 && known_c;
    const EmitGeom Gs = emit_geom(n, hw, per_image, pipe_ok ? EMIT_WARP_TILE : (long long)EMIT_TPB * (v4 ? 4 : 1));
    const EmitGeom Gr = emit_geom(n, hw, per_image, ECTA_TILE);
    p.geo_stream = EmitGeomDev{(int)Gs.n_runs, (int)Gs.tpc};
    p.geo_rec = EmitGeomDev{(int)Gr.n_runs, (int)Gr.tpc};
    CUDA_TRY(launch_chained(finalize_decide_kernel, dim3(DECIDE_BLOCKS), dim3(DECIDE_TPB), 0, st, p));
    LAUNCH_CHECK("finalize_decide_kernel");
    b200seg_stage(2, st);

    // K2
    {
        const long long chunks = (long long)p.groups * Gs.n_runs;
        if (p.have_records) {                              // record path (no-op unless K1d selected it)
            const long long rchunks = (long long)p.groups * Gr.n_runs;
            const int grid = (int)(rchunks < (long long)sms * ECTA_MINB ? rchunks : (long long)sms * ECTA_MINB);
            const size_t smem = sizeof(EctaSmem);
#define LAUNCH_EMIT_CTA(CC)                                                                                          \
    {                                                                                                                \
        if (up) {                                                                                                    \
            CUDA_TRY(cudaFuncSetAttribute(emit_kernel_cta<CC, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
            emit_kernel_cta<CC, true><<<grid, ECTA_TPB, smem, st>>>(p);                                              \
        } else {                                                                                                     \
            CUDA_TRY(cudaFuncSetAttribute(emit_kernel_cta<CC, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
            CUDA_TRY(launch_chained(emit_kernel_cta<CC, false>, dim3(grid), dim3(ECTA_TPB), smem, st, p));           \
        }                                                                                                            \
    }
            if (c == 8) LAUNCH_EMIT_CTA(8)
            else if (c == 17) LAUNCH_EMIT_CTA(17)
            else LAUNCH_EMIT_CTA(25)
#undef LAUNCH_EMIT_CTA
            LAUNCH_CHECK("emit_kernel_cta");
        }
        if (up) {
#define LAUNCH_EMIT_UP(CC)                                                                                        \
    {                                                                                                             \
        constexpr int ET = 128;                                                                                   \
        const size_t smem = (size_t)(ET / 32) * 3 * CC * 32 * 4;                                                  \
        CUDA_TRY(cudaFuncSetAttribute(emit_kernel_up<CC, ET>, cudaFuncAttributeMaxDynamicSharedMemorySize,        \
                                      (int)smem));                                                                \
        int per_sm = (int)((224 * 1024) / (smem + 2048));                                                         \
        if (per_sm * ET > 2048) per_sm = 2048 / ET;                                                               \
        emit_kernel_up<CC, ET><<<sms * per_sm, ET, smem, st>>>(p);                                                \
    }
            if (c == 8) LAUNCH_EMIT_UP(8)
            else if (c == 17) LAUNCH_EMIT_UP(17)
            else LAUNCH_EMIT_UP(25)
#undef LAUNCH_EMIT_UP
        } else if (pipe_ok) {
#define LAUNCH_EMIT_ASYNC(CC)                                                                                   \
    {                                                                                                           \
        constexpr int ET = 128, ES = 2;                                                                         \
        const size_t smem = sizeof(float) * (size_t)ES * CC * ET;                                               \
        CUDA_TRY(cudaFuncSetAttribute(emit_kernel_async<CC, ET>,                                                \
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));                 \
        int per_sm = (int)((224 * 1024) / (smem + 2048));                                                       \
        if (per_sm * ET > 2048) per_sm = 2048 / ET;                                                             \
        long long grid = (long long)sms * per_sm;                                                               \
        if (grid * (ET / 32) > chunks) grid = (chunks + ET / 32 - 1) / (ET / 32);                               \
        CUDA_TRY(launch_chained(emit_kernel_async<CC, ET>, dim3((unsigned)grid), dim3(ET), smem, st, p));       \
    }
            if (c == 8) LAUNCH_EMIT_ASYNC(8)
            else if (c == 17) LAUNCH_EMIT_ASYNC(17)
            else LAUNCH_EMIT_ASYNC(25)
#undef LAUNCH_EMIT_ASYNC
        } else {
            const int grid = (int)(chunks < (long long)sms * 8 ? chunks : (long long)sms * 8);
            if (v4) { DISPATCH_LABEL(label_dtype, emit_kernel<4, LT><<<grid, EMIT_TPB, 0, st>>>(p)); }
            else { DISPATCH_LABEL(label_dtype, emit_kernel<1, LT><<<grid, EMIT_TPB, 0, st>>>(p)); }
        }
        LAUNCH_CHECK("emit_kernel");
    }
    b200seg_stage(3, st);

    // sort: compact, unordered emission output in A; hybrid path A -> B (buckets); LSD path ping-pong, result in B
    SortArgs a;
    char* ss = ws + L.sort_scratch;
    a.keys[0] = p.keysA; a.vals[0] = p.valsA; a.keys[1] = p.keysB; a.vals[1] = p.valsB;
    a.seg_count = p.seg_count; a.seg_bits = p.seg_bits; a.n_seg = p.n_seg; a.cap = p.cap;
    a.tile_start = (u32*)(ss + L.sort.tile_start); a.tilehist = (u32*)(ss + L.sort.tilehist);
    a.tile_desc = (uint4*)(ss + L.sort.tile_desc);
    a.seg_done = (u32*)(ss + L.sort.seg_done);
    a.bin_base = (u32*)(ss + L.sort.bin_base); a.tile_fg = (u32*)(ss + L.sort.tile_fg);
    a.big = (u32*)(ss + L.sort.big); a.chunksum = (u32*)(ss + L.sort.chunksum); a.gbar = (u32*)(ss + L.sort.gbar);
    a.status = p.status;
    a.seg_sel = nullptr;
    a.val_passes = SORT_VAL_PASSES;                        // the emission order is arbitrary: canonical ties come from the values
    if (b200seg_tuning().sort_path == 1) {                 // plain path: three LSD passes, then the Jaccard kernel
        if (int rc = sort_enqueue(a, L.sort, st)) return rc;
        jaccard_kernel<<<sms * 4, JAC_TPB, 0, st>>>(p, a);
        LAUNCH_CHECK("jaccard_kernel");
    } else {                                               // hybrid path: one MSD partition pass + the fused local kernel
        HybArgs h;
        h.hist = (u32*)(ws + L.hyb_hist); h.fgpre = (u32*)(ws + L.hyb_fgpre); h.seg_done = (u32*)(ws + L.hyb_done);
        h.ticket = p.ctrl + CTRL_HTICKET; h.seg_ovf = (u32*)(ws + L.seg_ovf); h.ovf_any = p.ctrl + CTRL_OVF_ANY;
        h.seg_loss = (double*)(ws + L.seg_loss_h);
        if (int rc = hybrid_enqueue(p, a, h, L.sort, st)) return rc;
    }
    b200seg_stage(8, st);
    return 0;
}

extern "C" int b200seg_lovasz_forward(const float* logits, const void* labels, int32_t label_dtype, int32_t n,
                                      int32_t c, int64_t hw, int32_t per_image, int64_t filter_label,
                                      int32_t keep_absent, uint32_t class_mask, int32_t need_grad, void* workspace,
                                      size_t workspace_bytes, float* loss_out, int64_t* cm, int64_t cm_drop_label,
                                      int32_t* status, void* stream) {
    return lovasz_forward_impl(logits, labels, label_dtype, n, c, hw, per_image, filter_label, keep_absent, class_mask,
                               need_grad, workspace, workspace_bytes, loss_out, cm, cm_drop_label, status, stream,
                               false, B200SEG_NO_LABEL, nullptr);
}

extern "C" int b200seg_lovasz_ce_forward(const float* logits, const void* labels, int32_t label_dtype, int32_t n,
                                         int32_t c, int64_t hw, int32_t per_image, int64_t filter_label,
                                         int32_t keep_absent, uint32_t class_mask, int32_t need_grad, void* workspace,
                                         size_t workspace_bytes, float* loss_out, int64_t ce_ignore_index, float* ce_out,
                                         int64_t* cm, int64_t cm_drop_label, int32_t* status, void* stream) {
    return lovasz_forward_impl(logits, labels, label_dtype, n, c, hw, per_image, filter_label, keep_absent, class_mask,
                               need_grad, workspace, workspace_bytes, loss_out, cm, cm_drop_label, status, stream,
                               true, ce_ignore_index, ce_out);
}

static int lovasz_backward_impl(const float* logits, const void* labels, int32_t label_dtype, int32_t n,
                                int32_t c, int64_t hw, int32_t per_image, int64_t filter_label,
                                int32_t keep_absent, uint32_t class_mask, const void* workspace,
                                size_t workspace_bytes, const float* grad_out, float* dlogits, void* stream,
                                bool ce_enabled, int64_t ce_ignore, const float* grad_ce, const UpSrc* up = nullptr) {
    if (int rc = check_shape(n, c, hw)) return rc;
    if ((long long)n * hw == 0) return 0;
    if ((!logits && !up) || !labels || !workspace || !grad_out || !dlogits) {
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
    p.ce_enabled = ce_enabled ? 1 : 0;
    p.has_ce_ignore = (ce_enabled && ce_ignore >= INT_MIN && ce_ignore <= INT_MAX) ? 1 : 0;
    p.ce_ignore = p.has_ce_ignore ? (int)ce_ignore : 0;
    const int sms = b200seg_sm_count();
    if (up) {                                              // fused upsampling: the gradient goes to the low-resolution logits
        p.up = *up;
        p.dbg = b200seg_tuning().dbg;
        b200seg_stage(9, st);
        CUDA_TRY(cudaMemsetAsync(dlogits, 0, sizeof(float) * (size_t)n * c * up->h * up->w, st));
#define LAUNCH_BWD_UP(CC)                                                                                         \
    {                                                                                                             \
        constexpr int TT = 128;                                                                                   \
        const size_t smem = (size_t)(TT / 32) * sizeof(UpBwdSmem<CC>);                                            \
        CUDA_TRY(cudaFuncSetAttribute(backward_kernel_up<CC, TT>, cudaFuncAttributeMaxDynamicSharedMemorySize,    \
                                      (int)smem));                                                                \
        int per_sm = (int)((224 * 1024) / (smem + 2048));                                                         \
        if (per_sm > 4) per_sm = 4;                                                                               \
        backward_kernel_up<CC, TT><<<sms * per_sm, TT, smem, st>>>(p, grad_out, grad_ce, dlogits);                \
    }
        if (c == 8) LAUNCH_BWD_UP(8)
        else if (c == 17) LAUNCH_BWD_UP(17)
        else LAUNCH_BWD_UP(25)
#undef LAUNCH_BWD_UP
        LAUNCH_CHECK("backward_kernel_up");
        b200seg_stage(10, st);
        return 0;
    }
    const bool v4 = vec4_ok(logits, labels, label_dtype, hw) && aligned16(dlogits);
    const bool pipe_ok = v4 && hw % 16 == 0 && (c == 8 || c == 17 || c == 25);
    if (ce_enabled && !pipe_ok) {
        b200seg_set_error("the fused cross-entropy term needs the pipelined backward kernel (see b200seg_lovasz_ce_supported)");
        return B200SEG_E_UNSUPPORTED;
    }
    b200seg_stage(9, st);
    if (pipe_ok) {
#define LAUNCH_BWD_ASYNC(CC, TT)                                                                                 \
    {                                                                                                            \
        const size_t smem = (size_t)2 * (TT / 32) * sizeof(BwdStage<CC>);                                        \
        CUDA_TRY(cudaFuncSetAttribute(backward_kernel_async<CC, TT>,                                             \
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));                  \
        int per_sm = (int)((224 * 1024) / (smem + 2048));                                                        \
        if (per_sm < 1) per_sm = 1;                                                                              \
        if (per_sm * TT > 2048) per_sm = 2048 / TT;                                                              \
        backward_kernel_async<CC, TT><<<sms * per_sm, TT, smem, st>>>(p, grad_out, grad_ce, dlogits);                     \
    }
        if (c == 8) LAUNCH_BWD_ASYNC(8, 128)
        else if (c == 17) LAUNCH_BWD_ASYNC(17, 128)
        else LAUNCH_BWD_ASYNC(25, 128)
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

extern "C" int b200seg_lovasz_backward(const float* logits, const void* labels, int32_t label_dtype, int32_t n,
                                       int32_t c, int64_t hw, int32_t per_image, int64_t filter_label,
                                       int32_t keep_absent, uint32_t class_mask, const void* workspace,
                                       size_t workspace_bytes, const float* grad_out, float* dlogits, void* stream) {
    return lovasz_backward_impl(logits, labels, label_dtype, n, c, hw, per_image, filter_label, keep_absent, class_mask,
                                workspace, workspace_bytes, grad_out, dlogits, stream, false, B200SEG_NO_LABEL, nullptr);
}

extern "C" int b200seg_lovasz_ce_backward(const float* logits, const void* labels, int32_t label_dtype, int32_t n,
                                          int32_t c, int64_t hw, int32_t per_image, int64_t filter_label,
                                          int32_t keep_absent, uint32_t class_mask, const void* workspace,
                                          size_t workspace_bytes, const float* grad_lovasz, int64_t ce_ignore_index,
                                          const float* grad_ce, float* dlogits, void* stream) {
    if (!grad_ce) { b200seg_set_error("grad_ce is NULL"); return B200SEG_E_INVALID; }
    return lovasz_backward_impl(logits, labels, label_dtype, n, c, hw, per_image, filter_label, keep_absent, class_mask,
                                workspace, workspace_bytes, grad_lovasz, dlogits, stream, true, ce_ignore_index, grad_ce);
}

// ---- fused bilinear upsampling (align_corners = True) + loss: the logits stay at the model's resolution -------------------------
static int up_check(const float* lowres, int32_t n, int32_t c, int32_t h, int32_t w, int32_t H, int32_t W, UpSrc* u) {
    if (h < 1 || w < 1 || H < 1 || W < 1) { b200seg_set_error("invalid upsampling shape %dx%d -> %dx%d", h, w, H, W); return B200SEG_E_INVALID; }
    if (int rc = check_shape(n, c, (int64_t)H * W)) return rc;
    if (!(c == 8 || c == 17 || c == 25) || W % 32 != 0) {
        b200seg_set_error("fused upsampling covers C in {8, 17, 25} and output widths that are multiples of 32 (got C=%d, W=%d)", c, W);
        return B200SEG_E_UNSUPPORTED;
    }
    *u = make_up_src(lowres, h, w, H, W);
    // source columns one 32-pixel strip can touch (backward flush table)
    if ((int)ceilf(31.0f * u->rx) + 2 > UP_MAX_COLS) {
        b200seg_set_error("fused upsampling needs a horizontal scale factor of at least ~3.2 (got %d -> %d)", w, W);
        return B200SEG_E_UNSUPPORTED;
    }
    if ((long double)n * c * h * w >= (long double)(1ull << 31)) { b200seg_set_error("low-resolution logits too large"); return B200SEG_E_INVALID; }
    return 0;
}

extern "C" int b200seg_lovasz_up_supported(int32_t n, int32_t c, int32_t h, int32_t w, int32_t H, int32_t W) {
    UpSrc u;
    return up_check(nullptr, n, c, h, w, H, W, &u) == 0 ? 1 : 0;
}

extern "C" int b200seg_lovasz_up_forward(const float* lowres, int32_t h, int32_t w, const void* labels, int32_t label_dtype,
                                         int32_t n, int32_t c, int32_t H, int32_t W, int32_t per_image, int64_t filter_label,
                                         int32_t keep_absent, uint32_t class_mask, int32_t need_grad, void* workspace,
                                         size_t workspace_bytes, float* loss_out, int32_t ce_enabled, int64_t ce_ignore_index,
                                         float* ce_out, int64_t* cm, int64_t cm_drop_label, int32_t* status, void* stream) {
    UpSrc u;
    if (int rc = up_check(lowres, n, c, h, w, H, W, &u)) return rc;
    if (!lowres && (long long)n * H * W != 0) { b200seg_set_error("null pointer argument"); return B200SEG_E_INVALID; }
    return lovasz_forward_impl(nullptr, labels, label_dtype, n, c, (int64_t)H * W, per_image, filter_label, keep_absent,
                               class_mask, need_grad, workspace, workspace_bytes, loss_out, cm, cm_drop_label, status, stream,
                               ce_enabled != 0, ce_ignore_index, ce_out, &u);
}

extern "C" int b200seg_lovasz_up_backward(const float* lowres, int32_t h, int32_t w, const void* labels, int32_t label_dtype,
                                          int32_t n, int32_t c, int32_t H, int32_t W, int32_t per_image, int64_t filter_label,
                                          int32_t keep_absent, uint32_t class_mask, const void* workspace,
                                          size_t workspace_bytes, const float* grad_lovasz, int32_t ce_enabled,
                                          int64_t ce_ignore_index, const float* grad_ce, float* dlowres, void* stream) {
    UpSrc u;
    if (int rc = up_check(lowres, n, c, h, w, H, W, &u)) return rc;
    if (!lowres && (long long)n * H * W != 0) { b200seg_set_error("null pointer argument"); return B200SEG_E_INVALID; }
    if (ce_enabled && !grad_ce) { b200seg_set_error("grad_ce is NULL"); return B200SEG_E_INVALID; }
    return lovasz_backward_impl(nullptr, labels, label_dtype, n, c, (int64_t)H * W, per_image, filter_label, keep_absent,
                                class_mask, workspace, workspace_bytes, grad_lovasz, dlowres, stream, ce_enabled != 0,
                                ce_ignore_index, grad_ce, &u);
}

// ---- test hook: byte offsets of a few workspace regions (tests inspect the per-pixel records) -----------------------
extern "C" int b200seg_debug_layout(int32_t n, int32_t c, int64_t hw, int32_t per_image, size_t* offsets, int32_t n_offsets) {
    if (!offsets || n_offsets < 8) { b200seg_set_error("need room for 8 offsets"); return B200SEG_E_INVALID; }
    if (int rc = check_shape(n, c, hw)) return rc;
    const LovaszLayout L = lovasz_layout(n, c, hw, per_image);
    offsets[0] = L.pix_m; offsets[1] = L.pix_s; offsets[2] = L.lab8; offsets[3] = L.cmask;
    offsets[4] = L.rec16; offsets[5] = L.rec4; offsets[6] = L.seg_thr; offsets[7] = L.grp_tmin;
    if (n_offsets >= 14) {
        offsets[8] = L.keysB; offsets[9] = L.valsB; offsets[10] = L.seg_count; offsets[11] = L.seg_bits;
        offsets[12] = L.hyb_hist; offsets[13] = L.hyb_fgpre;
    }
    return 0;
}

// ---- test hook: the register-constant exponential against expf(), bit for bit -------------------------------------------
__global__ void exp_check_kernel(const float* __restrict__ x, int n, int* __restrict__ mismatches) {
    ExpConsts k;
    k.load();
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float a = sm_exp_k(x[i], 0.f, k), b = expf(__fsub_rn(x[i], 0.f));
        if (__float_as_uint(a) != __float_as_uint(b) && !(a != a && b != b)) atomicAdd(mismatches, 1);
    }
}
extern "C" int b200seg_debug_exp_mismatches(const float* x, int32_t n, int32_t* mismatches, void* stream) {
    if (!x || !mismatches || n < 0) { b200seg_set_error("invalid argument"); return B200SEG_E_INVALID; }
    exp_check_kernel<<<256, 256, 0, (cudaStream_t)stream>>>(x, n, mismatches);
    LAUNCH_CHECK("exp_check_kernel");
    return 0;
}

// ---- test hook: bilinear upsampling (align_corners = True) as the fused kernels compute it ----------------------------------
// pattern < 0: the product's arithmetic; 0..35: the candidate contraction patterns of tools/upsample_pattern.py
extern "C" int b200seg_debug_upsample(const float* lowres, int32_t planes, int32_t h, int32_t w, int32_t H, int32_t W,
                                      float* out, int32_t pattern, void* stream) {
    if (!lowres || !out || planes < 1 || h < 1 || w < 1 || H < 1 || W < 1) { b200seg_set_error("invalid argument"); return B200SEG_E_INVALID; }
    const UpSrc u = make_up_src(lowres, h, w, H, W);
    upsample_debug_kernel<<<b200seg_sm_count() * 8, 256, 0, (cudaStream_t)stream>>>(u, planes, out, pattern);
    LAUNCH_CHECK("upsample_debug_kernel");
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
    a.tile_start = (u32*)(ss + L.tile_start); a.tilehist = (u32*)(ss + L.tilehist);
    a.tile_desc = (uint4*)(ss + L.tile_desc);
    a.seg_done = (u32*)(ss + L.seg_done);
    a.bin_base = (u32*)(ss + L.bin_base); a.tile_fg = (u32*)(ss + L.tile_fg);
    a.big = (u32*)(ss + L.big); a.chunksum = (u32*)(ss + L.chunksum); a.gbar = (u32*)(ss + L.gbar);
    a.status = status; a.seg_sel = nullptr; a.val_passes = 0;    // the hook sorts by key only, stably
    return sort_enqueue(a, L, (cudaStream_t)stream);
}
