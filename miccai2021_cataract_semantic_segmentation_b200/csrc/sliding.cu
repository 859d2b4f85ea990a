// Windowed mean IoU map for sm_100a.  Replaces utils/torch_utils.py:189-218 (sliding_miou: argmax + two fp32 one-hots
// + two unfolds of [N, C*k*k, windows] + int casts + and/or reductions) of the reference.
//
// Two kernels.  class_map_kernel is the only pass over the logits (4*C + label bytes read and 2 bytes written per pixel,
// the roofline of this path): first-maximum argmax like the confusion matrix, prediction and label packed into one u16
// per pixel.  window_iou_kernel takes one window per thread: it counts, per class, predicted / labelled / agreeing
// pixels of the window in 16-bit shared-memory counters private to the thread (the 8 MB class map is L2-resident, every
// pixel is revisited ~(k/stride)^2 times), then averages I/U over the classes (U == 0 counts as 1); only the classes that
// occur in the window are visited.
#include "b200seg.h"
#include "common.cuh"

#define SM_TPB 256      // class map: threads per CTA, 4 pixels each
#define SW_TPB 128      // windows per CTA

struct ClassMapParams {
    const float* pred;
    const void* labels;
    int N, C;
    long long HW;
    unsigned short* map;
    int* status;
};

__device__ __forceinline__ unsigned short pack_class(int arg, int lab, int C, u32& oob) {
    if ((unsigned)lab >= (unsigned)C) { oob = 1; lab = 255; }       // never equal to a class id (C <= 32)
    return (unsigned short)(arg | (lab << 8));
}

template <int CT, typename LT>
__global__ void __launch_bounds__(SM_TPB) class_map_kernel_v4(ClassMapParams p) {
    constexpr int TILE_PX = SM_TPB * 4;
    const long long tpi = (p.HW + TILE_PX - 1) / TILE_PX;
    const long long ntiles = tpi * p.N;
    u32 oob = 0;
    for (long long t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const int n = (int)(t / tpi);
        const long long q0 = (t - (long long)n * tpi) * TILE_PX + threadIdx.x * 4;
        if (q0 >= p.HW) continue;
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
        ushort4 o;
        o.x = pack_class(arg[0], lab[0], CT, oob); o.y = pack_class(arg[1], lab[1], CT, oob);
        o.z = pack_class(arg[2], lab[2], CT, oob); o.w = pack_class(arg[3], lab[3], CT, oob);
        *reinterpret_cast<ushort4*>(p.map + (size_t)n * p.HW + q0) = o;
    }
    if (oob) atomicOr(p.status, STATUS_LABEL_OOB);
}

template <typename LT>
__global__ void __launch_bounds__(SM_TPB) class_map_kernel_generic(ClassMapParams p) {
    const int C = p.C;
    const long long P = (long long)p.N * p.HW;
    u32 oob = 0;
    for (long long px = (long long)blockIdx.x * SM_TPB + threadIdx.x; px < P; px += (long long)gridDim.x * SM_TPB) {
        const long long n = px / p.HW, q = px - n * p.HW;
        const float* lp = p.pred + (size_t)n * C * p.HW + q;
        float best = __ldg(lp);
        int arg = 0;
        for (int c = 1; c < C; ++c) argmax_step(__ldg(lp + (size_t)c * p.HW), c, best, arg);
        p.map[px] = pack_class(arg, load_label<LT>(p.labels, (size_t)px), C, oob);
    }
    if (oob) atomicOr(p.status, STATUS_LABEL_OOB);
}

// counters: s_cnt[(kind * C + class) * SW_TPB + thread], kind 0 = predicted only, 1 = labelled only, 2 = both
__global__ void __launch_bounds__(SW_TPB) window_iou_kernel(const unsigned short* __restrict__ map, int N, int C, int H,
                                                            int W, int K, int S, int VW, int HWIN,
                                                            float* __restrict__ out) {
    extern __shared__ unsigned short s_cnt[];
    const int tid = threadIdx.x;
    const long long total = (long long)N * VW * HWIN;
    const float inv_c = 1.0f / (float)C;
    for (int r = 0; r < 3 * C; ++r) s_cnt[r * SW_TPB + tid] = 0;     // thread-private counters: no barrier needed
    for (long long base = (long long)blockIdx.x * SW_TPB; base < total; base += (long long)gridDim.x * SW_TPB) {
        const long long id = base + tid;
        if (id >= total) continue;
        const int wx = (int)(id % HWIN);
        const long long rest = id / HWIN;
        const int wy = (int)(rest % VW), n = (int)(rest / VW);
        const unsigned short* src = map + ((size_t)n * H + (size_t)wy * S) * W + (size_t)wx * S;
        u32 seen = 0;                                                // classes predicted or labelled in this window
        for (int dy = 0; dy < K; ++dy) {
            const unsigned short* row = src + (size_t)dy * W;
            for (int dx = 0; dx < K; ++dx) {
                const u32 v = __ldg(row + dx);
                const u32 pc = v & 255u, tc = v >> 8;
                seen |= 1u << pc;
                if (tc == pc) {                                      // agreeing pixel: one counter instead of three
                    s_cnt[(2 * C + pc) * SW_TPB + tid] += 1;
                } else {
                    s_cnt[pc * SW_TPB + tid] += 1;
                    if (tc < (u32)C) {
                        s_cnt[(C + tc) * SW_TPB + tid] += 1;
                        seen |= 1u << tc;
                    }
                }
            }
        }
        // classes absent from the window score 1 each (0/0 -> 1); the others add I/U in class order and hand their
        // counters back zeroed
        float sum = (float)(C - __popc(seen));
        while (seen) {
            const int c = __ffs(seen) - 1;
            seen &= seen - 1;
            const u32 np = s_cnt[c * SW_TPB + tid], nt = s_cnt[(C + c) * SW_TPB + tid];
            const u32 ni = s_cnt[(2 * C + c) * SW_TPB + tid];
            s_cnt[c * SW_TPB + tid] = 0; s_cnt[(C + c) * SW_TPB + tid] = 0; s_cnt[(2 * C + c) * SW_TPB + tid] = 0;
            sum += __fdiv_rn((float)ni, (float)(np + nt + ni));
        }
        out[id] = sum * inv_c;
    }
}

#define DISPATCH_LABEL(dtype, ...)                                               \
    switch (dtype) {                                                             \
        case B200SEG_LABEL_U8: { typedef uint8_t LT; __VA_ARGS__; } break;        \
        case B200SEG_LABEL_I32: { typedef int32_t LT; __VA_ARGS__; } break;       \
        case B200SEG_LABEL_I64: { typedef int64_t LT; __VA_ARGS__; } break;       \
        default: b200seg_set_error("unknown label dtype %d", dtype); return B200SEG_E_INVALID; \
    }

static int window_count(int extent, int k, int s) { return extent < k ? 0 : (extent - k) / s + 1; }

extern "C" int b200seg_sliding_miou_scratch_bytes(int32_t n, int64_t h, int64_t w, size_t* bytes) {
    if (!bytes || n < 0 || h < 0 || w < 0) { b200seg_set_error("b200seg_sliding_miou_scratch_bytes: bad argument"); return B200SEG_E_INVALID; }
    *bytes = align_up((size_t)n * (size_t)h * (size_t)w * sizeof(unsigned short), 256);
    return 0;
}

extern "C" int b200seg_sliding_miou(const float* prediction, const void* labels, int32_t label_dtype, int32_t n,
                                    int32_t c, int32_t h, int32_t w, int32_t kernel_size, int32_t stride,
                                    void* scratch, size_t scratch_bytes, float* out, int32_t* status, void* stream) {
    if (n < 0 || h < 0 || w < 0 || c < 1 || c > B200SEG_MAX_CLASSES ||
        (long double)n * h * w >= (long double)(1u << 30)) {
        b200seg_set_error("invalid shape: n_images=%d n_classes=%d %dx%d", n, c, h, w);
        return B200SEG_E_INVALID;
    }
    if (kernel_size < 1 || kernel_size > 255 || kernel_size % 2 == 0 || stride < 1) {
        b200seg_set_error("sliding_miou: kernel size must be odd and in [1, 255], stride >= 1 (got %d, %d)", kernel_size, stride);
        return B200SEG_E_INVALID;
    }
    const int vw = window_count(h, kernel_size, stride), hwin = window_count(w, kernel_size, stride);
    const long long windows = (long long)n * vw * hwin;
    if (windows == 0) return 0;
    size_t need = 0;
    b200seg_sliding_miou_scratch_bytes(n, h, w, &need);
    if (!prediction || !labels || !scratch || !out || !status) { b200seg_set_error("null pointer argument"); return B200SEG_E_INVALID; }
    if (scratch_bytes < need) {
        b200seg_set_error("sliding_miou: scratch of %zu bytes, need %zu", scratch_bytes, need);
        return B200SEG_E_WORKSPACE;
    }
    ClassMapParams p;
    p.pred = prediction; p.labels = labels; p.N = n; p.C = c; p.HW = (long long)h * w;
    p.map = (unsigned short*)scratch; p.status = status;
    cudaStream_t st = (cudaStream_t)stream;
    const int sms = b200seg_sm_count();
    bool v4 = p.HW % 4 == 0 && ((uintptr_t)prediction & 15) == 0 && ((uintptr_t)scratch & 7) == 0;
    v4 = v4 && (label_dtype == B200SEG_LABEL_U8 ? ((uintptr_t)labels & 3) == 0 : ((uintptr_t)labels & 15) == 0);
    if (v4 && (c == 8 || c == 17 || c == 25)) {
        const long long tiles = (long long)n * ((p.HW + SM_TPB * 4 - 1) / (SM_TPB * 4));
        const int grid = (int)(tiles < (long long)sms * 8 ? tiles : (long long)sms * 8);
        DISPATCH_LABEL(label_dtype, {
            if (c == 8) class_map_kernel_v4<8, LT><<<grid, SM_TPB, 0, st>>>(p);
            else if (c == 17) class_map_kernel_v4<17, LT><<<grid, SM_TPB, 0, st>>>(p);
            else class_map_kernel_v4<25, LT><<<grid, SM_TPB, 0, st>>>(p);
        });
    } else {
        const long long blocks = ((long long)n * p.HW + SM_TPB - 1) / SM_TPB;
        const int grid = (int)(blocks < (long long)sms * 8 ? blocks : (long long)sms * 8);
        DISPATCH_LABEL(label_dtype, class_map_kernel_generic<LT><<<grid, SM_TPB, 0, st>>>(p));
    }
    LAUNCH_CHECK("class_map_kernel");
    const size_t smem = (size_t)3 * c * SW_TPB * sizeof(unsigned short);
    const long long blocks = (windows + SW_TPB - 1) / SW_TPB;
    const int grid = (int)(blocks < (long long)sms * 16 ? blocks : (long long)sms * 16);
    window_iou_kernel<<<grid, SW_TPB, smem, st>>>(p.map, n, c, h, w, kernel_size, stride, vw, hwin, out);
    LAUNCH_CHECK("window_iou_kernel");
    return 0;
}
