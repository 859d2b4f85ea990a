// Confusion-matrix accumulation with fused argmax, and the C x C IoU / accuracy summary, for sm_100a.
// Replaces utils/torch_utils.py:221-241 (t_get_confusion_matrix: transpose copy + argmax + two int64 one-hots +
// fp32 GEMM) and the arithmetic of :259-332 (t_get_pixel_accuracy, t_get_miou) of the reference.
//
// One streaming pass over the logits (4*C + label bytes per pixel, the roofline of this path): every thread takes
// four consecutive pixels with 128-bit loads per class plane, keeps the running first-maximum, and adds
// (pred*C + label) to a shared-memory histogram private to the CTA (same-bin lanes serialise inside the atomic unit;
// cheaper than any warp pre-aggregation, see cm_add).  CTAs are persistent (a few per SM) and flush non-zero bins to
// the int64 matrix once at the end.
#include "b200seg.h"
#include "common.cuh"
#include "upsample.cuh"

#define CM_TPB 256

struct ConfmatParams {
    const float* pred;
    const void* labels;
    int N, C;
    long long HW;
    int has_drop, drop;
    unsigned long long* cm;
    int* status;
};

// Plain shared-memory atomics: lanes hitting one bin serialise at ~1 lane/cycle (32 cycles for a fully uniform warp),
// far cheaper than aggregating with match.any first (MATCH.ANY measured at ~250 cycles per warp instruction here).
__device__ __forceinline__ void cm_add(u32* s_cm, u32 bin) {
    if (bin != 0xFFFFFFFFu) atomicAdd(s_cm + bin, 1u);
}
__device__ __forceinline__ u32 cm_bin(const ConfmatParams& p, int lab, int arg, int C, u32& oob) {
    if (p.has_drop && lab == p.drop) return 0xFFFFFFFFu;
    if ((unsigned)lab >= (unsigned)C) { oob = 1; return 0xFFFFFFFFu; }
    return (u32)(arg * C + lab);
}

template <int CT, typename LT>
__global__ void __launch_bounds__(CM_TPB) confmat_kernel_v4(ConfmatParams p) {
    __shared__ u32 s_cm[B200SEG_MAX_CLASSES * B200SEG_MAX_CLASSES];
    const int tid = threadIdx.x;
    constexpr int TILE_PX = CM_TPB * 4;
    const long long tpi = (p.HW + TILE_PX - 1) / TILE_PX;
    const long long ntiles = tpi * p.N;
    const long long t0 = ntiles * blockIdx.x / gridDim.x, t1 = ntiles * (blockIdx.x + 1) / gridDim.x;
    for (int i = tid; i < CT * CT; i += CM_TPB) s_cm[i] = 0;
    __syncthreads();
    u32 oob = 0;
    for (long long t = t0; t < t1; ++t) {
        const int n = (int)(t / tpi);
        const long long q0 = (t - (long long)n * tpi) * TILE_PX + tid * 4;
        u32 bin[4] = {0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu};
        if (q0 < p.HW) {
            const float* lp = p.pred + (size_t)n * CT * p.HW + q0;
            float4 v[CT];
#pragma unroll
            for (int c = 0; c < CT; ++c) v[c] = ld_stream4(lp + (size_t)c * p.HW);
            int lab[4];
            load_labels4<LT>(p.labels, (size_t)n * p.HW + q0, lab);
            float best[4] = {v[0].x, v[0].y, v[0].z, v[0].w};
            int arg[4] = {0, 0, 0, 0};
#pragma unroll
            for (int c = 1; c < CT; ++c) {
                argmax_step(v[c].x, c, best[0], arg[0]);
                argmax_step(v[c].y, c, best[1], arg[1]);
                argmax_step(v[c].z, c, best[2], arg[2]);
                argmax_step(v[c].w, c, best[3], arg[3]);
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) bin[j] = cm_bin(p, lab[j], arg[j], CT, oob);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) cm_add(s_cm, bin[j]);
    }
    __syncthreads();
    for (int i = tid; i < CT * CT; i += CM_TPB)
        if (s_cm[i]) atomicAdd(p.cm + i, (unsigned long long)s_cm[i]);
    if (oob) atomicOr(p.status, STATUS_LABEL_OOB);
}

template <typename LT>
__global__ void __launch_bounds__(CM_TPB) confmat_kernel_generic(ConfmatParams p) {
    __shared__ u32 s_cm[B200SEG_MAX_CLASSES * B200SEG_MAX_CLASSES];
    const int tid = threadIdx.x;
    const int C = p.C;
    constexpr int TILE_PX = CM_TPB;
    const long long tpi = (p.HW + TILE_PX - 1) / TILE_PX;
    const long long ntiles = tpi * p.N;
    const long long t0 = ntiles * blockIdx.x / gridDim.x, t1 = ntiles * (blockIdx.x + 1) / gridDim.x;
    for (int i = tid; i < C * C; i += CM_TPB) s_cm[i] = 0;
    __syncthreads();
    u32 oob = 0;
    for (long long t = t0; t < t1; ++t) {
        const int n = (int)(t / tpi);
        const long long q = (t - (long long)n * tpi) * TILE_PX + tid;
        u32 bin = 0xFFFFFFFFu;
        if (q < p.HW) {
            const float* lp = p.pred + (size_t)n * C * p.HW + q;
            float best = __ldg(lp);
            int arg = 0;
            for (int c = 1; c < C; ++c) argmax_step(__ldg(lp + (size_t)c * p.HW), c, best, arg);
            bin = cm_bin(p, load_label<LT>(p.labels, (size_t)n * p.HW + q), arg, C, oob);
        }
        cm_add(s_cm, bin);
    }
    __syncthreads();
    for (int i = tid; i < C * C; i += CM_TPB)
        if (s_cm[i]) atomicAdd(p.cm + i, (unsigned long long)s_cm[i]);
    if (oob) atomicOr(p.status, STATUS_LABEL_OOB);
}

#define DISPATCH_LABEL(dtype, ...)                                               \
    switch (dtype) {                                                             \
        case B200SEG_LABEL_U8: { typedef uint8_t LT; __VA_ARGS__; } break;        \
        case B200SEG_LABEL_I32: { typedef int32_t LT; __VA_ARGS__; } break;       \
        case B200SEG_LABEL_I64: { typedef int64_t LT; __VA_ARGS__; } break;       \
        default: b200seg_set_error("unknown label dtype %d", dtype); return B200SEG_E_INVALID; \
    }

extern "C" int b200seg_confmat_accumulate(const float* prediction, const void* labels, int32_t label_dtype,
                                          int32_t n, int32_t c, int64_t hw, int64_t drop_label, int64_t* cm,
                                          int32_t* status, void* stream) {
    if (n < 0 || hw < 0 || c < 1 || c > B200SEG_MAX_CLASSES || (long double)n * hw >= (long double)(1u << 30)) {
        b200seg_set_error("invalid shape: n_images=%d n_classes=%d plane=%lld", n, c, (long long)hw);
        return B200SEG_E_INVALID;
    }
    if (!prediction || !labels || !cm || !status) { b200seg_set_error("null pointer argument"); return B200SEG_E_INVALID; }
    if ((long long)n * hw == 0) return 0;
    ConfmatParams p;
    p.pred = prediction; p.labels = labels; p.N = n; p.C = c; p.HW = hw;
    p.has_drop = (drop_label != B200SEG_NO_LABEL && drop_label >= INT_MIN && drop_label <= INT_MAX) ? 1 : 0;
    p.drop = p.has_drop ? (int)drop_label : 0;
    p.cm = (unsigned long long*)cm; p.status = status;
    cudaStream_t st = (cudaStream_t)stream;
    const int sms = b200seg_sm_count();
    bool v4 = hw % 4 == 0 && ((uintptr_t)prediction & 15) == 0;
    v4 = v4 && (label_dtype == B200SEG_LABEL_U8 ? ((uintptr_t)labels & 3) == 0 : ((uintptr_t)labels & 15) == 0);
    if (v4 && (c == 8 || c == 17 || c == 25)) {
        const long long tiles = (long long)n * ((hw + CM_TPB * 4 - 1) / (CM_TPB * 4));
        const int grid = (int)(tiles < (long long)sms * 4 ? tiles : (long long)sms * 4);
        DISPATCH_LABEL(label_dtype, {
            if (c == 8) confmat_kernel_v4<8, LT><<<grid, CM_TPB, 0, st>>>(p);
            else if (c == 17) confmat_kernel_v4<17, LT><<<grid, CM_TPB, 0, st>>>(p);
            else confmat_kernel_v4<25, LT><<<grid, CM_TPB, 0, st>>>(p);
        });
    } else {
        const long long tiles = (long long)n * ((hw + CM_TPB - 1) / CM_TPB);
        const int grid = (int)(tiles < (long long)sms * 8 ? tiles : (long long)sms * 8);
        DISPATCH_LABEL(label_dtype, confmat_kernel_generic<LT><<<grid, CM_TPB, 0, st>>>(p));
    }
    LAUNCH_CHECK("confmat_kernel");
    return 0;
}

// ---- confusion matrix straight from low-resolution logits (SURVEY 8 F2 applied to the metric path) --------------------------------
// The validation loop's matrix is t_get_confusion_matrix(model(img), lbl) (managers/OCRNet_Manager.py:161,
// managers/BaseManager.py:640-688) on logits the model has just upsampled with F.interpolate(align_corners=True)
// (models/OCR.py:126-131).  This kernel interpolates in shared memory with ATen's arithmetic (upsample.cuh), so the argmax is
// the one torch would take on the upsampled tensor, and the 4*C bytes per pixel of that tensor are never written or read.
// Work decomposition as in lovasz_up.cuh: a warp owns (image, strip of 32 columns, source-row interval), interpolates the two
// source rows horizontally once, then every output row is a vertical mix with a running first-maximum.
struct ConfmatUpParams {
    UpSrc up;
    const void* labels;
    int N, C;
    int has_drop, drop;
    unsigned long long* cm;
    int* status;
};
#define CMU_TPB 128
template <int CT, typename LT>                             // CT = 0: class count known at run time only
__global__ void __launch_bounds__(CMU_TPB) confmat_up_kernel(ConfmatUpParams p) {
    extern __shared__ __align__(16) float cmu_tiles[];
    __shared__ u32 s_cm[B200SEG_MAX_CLASSES * B200SEG_MAX_CLASSES];
    const int C = CT ? CT : p.C;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    float* H0 = cmu_tiles + (size_t)warp * 2 * C * 32;
    float* H1 = H0 + (size_t)C * 32;
    for (int i = tid; i < C * C; i += CMU_TPB) s_cm[i] = 0;
    __syncthreads();
    const UpSrc u = p.up;
    const u32 items = up_item_count(u, p.N);
    const u32 gw = blockIdx.x * (CMU_TPB / 32) + warp, nwarps = gridDim.x * (CMU_TPB / 32);
    const size_t pl = (size_t)u.h * u.w, HW = (size_t)u.H * u.W;
    u32 oob = 0;
    for (u32 item = gw; item < items; item += nwarps) {
        UpItem wi;
        if (!up_item(u, item, wi)) continue;
        const int X = wi.sx * 32 + lane;
        const UpAxis ax = up_axis(u.rx, X, u.w);
        const float* img = u.lo + (size_t)wi.n * C * pl;
        const int k1 = wi.k + ((wi.k < u.h - 1) ? 1 : 0);
        if (CT) {
            up_fill_row<CT ? CT : 1>(reinterpret_cast<float (*)[32]>(H0), img, wi.k, u, ax, lane);
            if (k1 != wi.k) up_fill_row<CT ? CT : 1>(reinterpret_cast<float (*)[32]>(H1), img, k1, u, ax, lane);
        } else {                                           // run-time class count: batches of four classes
            const float* r0 = img + (size_t)wi.k * u.w;
            const float* r1 = img + (size_t)k1 * u.w;
            for (int c0 = 0; c0 < C; c0 += 4) {
                float a0[4], b0[4], a1[4], b1[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const size_t o = (size_t)min(c0 + j, C - 1) * pl;
                    a0[j] = __ldg(r0 + o + ax.i0); b0[j] = __ldg(r0 + o + ax.i1);
                    a1[j] = __ldg(r1 + o + ax.i0); b1[j] = __ldg(r1 + o + ax.i1);
                }
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (c0 + j < C) {
                        H0[(c0 + j) * 32 + lane] = up_row(ax.l0, a0[j], ax.l1, b0[j]);
                        H1[(c0 + j) * 32 + lane] = up_row(ax.l0, a1[j], ax.l1, b1[j]);
                    }
            }
        }
        const float* Hb = (CT && k1 == wi.k) ? H0 : H1;     // (templated path: the bottom row is the top row at the last source row)
        size_t px = (size_t)wi.n * HW + (size_t)wi.ya * u.W + X;
        int lab_next = load_label<LT>(p.labels, px);
        for (int Y = wi.ya; Y < wi.yb; ++Y, px += u.W) {
            const int lab = lab_next;
            if (Y + 1 < wi.yb) lab_next = load_label<LT>(p.labels, px + u.W);
            const UpAxis ay = up_axis(u.ry, Y, u.h);
            float best = up_col(ay.l0, H0[lane], ay.l1, Hb[lane]);
            int arg = 0;
            if (CT) {
#pragma unroll
                for (int c = 1; c < CT; ++c) argmax_step(up_col(ay.l0, H0[c * 32 + lane], ay.l1, Hb[c * 32 + lane]), c, best, arg);
            } else {
                for (int c = 1; c < C; ++c) argmax_step(up_col(ay.l0, H0[c * 32 + lane], ay.l1, Hb[c * 32 + lane]), c, best, arg);
            }
            if (!(p.has_drop && lab == p.drop)) {
                if ((unsigned)lab < (unsigned)C) atomicAdd(s_cm + arg * C + lab, 1u);
                else oob = 1;
            }
        }
    }
    __syncthreads();
    for (int i = tid; i < C * C; i += CMU_TPB)
        if (s_cm[i]) atomicAdd(p.cm + i, (unsigned long long)s_cm[i]);
    if (oob) atomicOr(p.status, STATUS_LABEL_OOB);
}

extern "C" int b200seg_confmat_up_supported(int32_t n, int32_t c, int32_t h, int32_t w, int32_t H, int32_t W) {
    if (n < 0 || c < 1 || c > B200SEG_MAX_CLASSES || h < 1 || w < 1 || H < 1 || W < 1) return 0;
    if ((long double)n * H * W >= (long double)(1u << 30) || (long double)n * c * h * w >= (long double)(1ull << 31)) return 0;
    return W % 32 == 0 ? 1 : 0;
}

extern "C" int b200seg_confmat_up_accumulate(const float* lowres, int32_t h, int32_t w, const void* labels, int32_t label_dtype,
                                             int32_t n, int32_t c, int32_t H, int32_t W, int64_t drop_label, int64_t* cm,
                                             int32_t* status, void* stream) {
    if (n < 0 || c < 1 || c > B200SEG_MAX_CLASSES || h < 1 || w < 1 || H < 1 || W < 1 ||
        (long double)n * H * W >= (long double)(1u << 30) || (long double)n * c * h * w >= (long double)(1ull << 31)) {
        b200seg_set_error("invalid shape: n_images=%d n_classes=%d %dx%d -> %dx%d", n, c, h, w, H, W);
        return B200SEG_E_INVALID;
    }
    if (W % 32 != 0) {
        b200seg_set_error("confusion matrix from low-resolution logits needs an output width that is a multiple of 32 (got %d)", W);
        return B200SEG_E_UNSUPPORTED;
    }
    if (!lowres || !labels || !cm || !status) { b200seg_set_error("null pointer argument"); return B200SEG_E_INVALID; }
    if (n == 0) return 0;
    ConfmatUpParams p;
    p.up = make_up_src(lowres, h, w, H, W);
    p.labels = labels; p.N = n; p.C = c;
    p.has_drop = (drop_label != B200SEG_NO_LABEL && drop_label >= INT_MIN && drop_label <= INT_MAX) ? 1 : 0;
    p.drop = p.has_drop ? (int)drop_label : 0;
    p.cm = (unsigned long long*)cm; p.status = status;
    cudaStream_t st = (cudaStream_t)stream;
    const int sms = b200seg_sm_count();
    const size_t smem = (size_t)(CMU_TPB / 32) * 2 * c * 32 * sizeof(float);
    int per_sm = (int)((224 * 1024) / (smem + 5 * 1024));
    if (per_sm * CMU_TPB > 2048) per_sm = 2048 / CMU_TPB;
    const long long items = (long long)(W / 32) * h * p.up.jmax * n;
    long long grid = (long long)sms * per_sm;
    if (grid * (CMU_TPB / 32) > items) grid = (items + CMU_TPB / 32 - 1) / (CMU_TPB / 32);
    if (grid < 1) grid = 1;
#define LAUNCH_CMU(CC)                                                                                            \
    {                                                                                                             \
        CUDA_TRY(cudaFuncSetAttribute(confmat_up_kernel<CC, LT>, cudaFuncAttributeMaxDynamicSharedMemorySize,     \
                                      (int)smem));                                                                \
        confmat_up_kernel<CC, LT><<<(int)grid, CMU_TPB, smem, st>>>(p);                                           \
    }
    DISPATCH_LABEL(label_dtype, {
        if (c == 8) LAUNCH_CMU(8)
        else if (c == 17) LAUNCH_CMU(17)
        else if (c == 25) LAUNCH_CMU(25)
        else LAUNCH_CMU(0)
    });
#undef LAUNCH_CMU
    LAUNCH_CHECK("confmat_up_kernel");
    return 0;
}

// ---- IoU / accuracy summary -----------------------------------------------------------------------------------
#define MAX_SETS 8
struct MetricSets { u32 mask[MAX_SETS]; int n; };

__global__ void metrics_kernel(const long long* __restrict__ cm, int C, u32 miou_mask, MetricSets sets,
                               float* __restrict__ iou_out, float* __restrict__ summary) {
    __shared__ float s_iou[B200SEG_MAX_CLASSES], s_pac[B200SEG_MAX_CLASSES];
    __shared__ long long s_diag[B200SEG_MAX_CLASSES], s_row[B200SEG_MAX_CLASSES];
    __shared__ long long s_cm[B200SEG_MAX_CLASSES * (B200SEG_MAX_CLASSES + 1)];     // rows padded: column reads conflict-free
    for (int i = threadIdx.x; i < C * C; i += blockDim.x) s_cm[(i / C) * (C + 1) + i % C] = cm[i];   // one coalesced sweep
    __syncthreads();
    const int c = threadIdx.x;
    if (c < C) {
        long long row = 0, col = 0;                 // row: prediction totals (sum over dim 1); col: ground-truth totals
        for (int k = 0; k < C; ++k) { row += s_cm[c * (C + 1) + k]; col += s_cm[k * (C + 1) + c]; }
        const long long d = s_cm[c * (C + 1) + c];
        // utils/torch_utils.py:322-327: diag / (sum(dim=0) + sum(dim=1) - diag), NaN -> 0
        const float den = __fsub_rn(__fadd_rn((float)col, (float)row), (float)d);
        float iou = __fdiv_rn((float)d, den);
        if (iou != iou) iou = 0.f;
        s_iou[c] = iou;
        iou_out[c] = iou;
        // utils/torch_utils.py:266-270: correct / max(prediction row sum, 1)
        s_pac[c] = __fdiv_rn((float)d, row == 0 ? 1.0f : (float)row);
        s_diag[c] = d; s_row[c] = row;
    }
    __syncthreads();
    if (c == 0) {
        float acc = 0.f; int cnt = 0;
        for (int k = 0; k < C; ++k) if ((miou_mask >> k) & 1u) { acc += s_iou[k]; ++cnt; }
        summary[0] = cnt ? acc / (float)cnt : 0.f;
        long long dsum = 0, all = 0;
        float pac = 0.f;
        for (int k = 0; k < C; ++k) { dsum += s_diag[k]; all += s_row[k]; pac += s_pac[k]; }
        summary[1] = __fdiv_rn((float)dsum, (float)all);
        summary[2] = pac / (float)C;
        for (int s = 0; s < sets.n; ++s) {
            float a2 = 0.f; int n2 = 0;
            for (int k = 0; k < C; ++k) if ((sets.mask[s] >> k) & 1u) { a2 += s_iou[k]; ++n2; }
            summary[3 + s] = n2 ? a2 / (float)n2 : 0.f;
        }
    }
}

extern "C" int b200seg_metrics_from_confmat(const int64_t* cm, int32_t c, uint32_t miou_mask,
                                            const uint32_t* category_masks, int32_t n_sets, float* iou_out,
                                            float* summary_out, void* stream) {
    if (!cm || !iou_out || !summary_out || c < 1 || c > B200SEG_MAX_CLASSES || n_sets < 0 || n_sets > MAX_SETS ||
        (n_sets > 0 && !category_masks)) {
        b200seg_set_error("invalid argument to b200seg_metrics_from_confmat");
        return B200SEG_E_INVALID;
    }
    MetricSets sets;
    sets.n = n_sets;
    for (int i = 0; i < MAX_SETS; ++i) sets.mask[i] = i < n_sets ? category_masks[i] : 0;
    metrics_kernel<<<1, 256, 0, (cudaStream_t)stream>>>((const long long*)cm, c, miou_mask, sets, iou_out, summary_out);
    LAUNCH_CHECK("metrics_kernel");
    return 0;
}
