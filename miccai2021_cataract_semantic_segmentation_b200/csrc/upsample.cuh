// Bilinear upsampling (align_corners = True) of low-resolution logits, fused into the loss kernels.
// The reference's models produce logits at stride 8 (OCRNet, models/OCR.py:126-131) or 4 (DeepLabv3+,
// models/DeepLabv3Plus.py:65-68) and call F.interpolate(..., mode='bilinear', align_corners=True) right before the loss.
// The arithmetic here reproduces ATen's CUDA kernel (upsample_bilinear2d_out_frame) bit for bit, so argmax, probabilities,
// thresholds and tie order are those of the reference run on the same device:
//     src = scale * dst,  i = (int)src,  l1 = src - i,  l0 = 1 - l1        (scale = (in - 1) / (out - 1) in fp32)
//     val = l0y * (l0x * a + l1x * b) + l1y * (l0x * c + l1x * d)
// with the multiply-add contractions nvcc applies to that source (UP_* pattern constants below; pinned by
// tests/test_gpu_upsample.py against torch on the device, found with tools/upsample_pattern.py).
#pragma once
#include "common.cuh"

struct UpSrc {
    const float* lo;     // [N, C, h, w] low-resolution logits; nullptr = the kernels read full-resolution logits
    int h, w, H, W;
    float ry, rx;        // (h - 1) / (H - 1), (w - 1) / (W - 1), fp32 division of the converted integers (0 when H or W is 1)
    int jmax;            // work items per (strip, source-row interval): an interval's rows are walked in chunks of UP_ROWS_MAX
};
#define UP_ROWS_MAX 32   // output rows per work item at most (bounds the sequential fp32 accumulation of the backward pass and
                         // keeps the warps busy when the source has very few rows)

struct UpAxis {
    int i0, i1;          // the two source indices (i1 = i0 at the far edge)
    float l0, l1;        // their weights
};

// pattern bits: 0 = lambda via fma(scale, dst, -i), 1..2 = inner sum (0: fma(l0,a,l1*b), 1: fma(l1,b,l0*a), 2: unfused),
// 3..4 = outer sum likewise
#define UP_PATTERN_DEFAULT 0
template <int LFMA>
__device__ __forceinline__ UpAxis up_axis_t(float r, int dst, int size_in) {
    const float fd = (float)dst;
    const float src = __fmul_rn(r, fd);
    const int i = (int)src;
    UpAxis a;
    a.i0 = i;
    a.i1 = i + ((i < size_in - 1) ? 1 : 0);
    a.l1 = LFMA ? __fmaf_rn(r, fd, -(float)i) : __fsub_rn(src, (float)i);
    a.l0 = __fsub_rn(1.0f, a.l1);
    return a;
}
template <int MODE>
__device__ __forceinline__ float up_mix_t(float w0, float a, float w1, float b) {
    if (MODE == 0) return __fmaf_rn(w0, a, __fmul_rn(w1, b));
    if (MODE == 1) return __fmaf_rn(w1, b, __fmul_rn(w0, a));
    return __fadd_rn(__fmul_rn(w0, a), __fmul_rn(w1, b));
}

// ---- the product's pattern (set from the experiment; see the header comment) -------------------------------------------------
#ifndef UP_LFMA
#define UP_LFMA 0
#endif
#ifndef UP_INNER
#define UP_INNER 0
#endif
#ifndef UP_OUTER
#define UP_OUTER 0
#endif
__device__ __forceinline__ UpAxis up_axis(float r, int dst, int size_in) { return up_axis_t<UP_LFMA>(r, dst, size_in); }
__device__ __forceinline__ float up_row(float l0x, float a, float l1x, float b) { return up_mix_t<UP_INNER>(l0x, a, l1x, b); }
__device__ __forceinline__ float up_col(float l0y, float top, float l1y, float bot) { return up_mix_t<UP_OUTER>(l0y, top, l1y, bot); }

// one interpolated logit (the rare paths: guard-tripped pixels of the emission, absent classes)
__device__ __forceinline__ float up_logit(const UpSrc& u, int C, int n, int c, long long q) {
    const int Y = (int)(q / u.W), X = (int)(q - (long long)Y * u.W);
    const UpAxis ay = up_axis(u.ry, Y, u.h), ax = up_axis(u.rx, X, u.w);
    const float* pl = u.lo + ((size_t)n * C + c) * ((size_t)u.h * u.w);
    const float* r0 = pl + (size_t)ay.i0 * u.w;
    const float* r1 = pl + (size_t)ay.i1 * u.w;
    const float top = up_row(ax.l0, __ldg(r0 + ax.i0), ax.l1, __ldg(r0 + ax.i1));
    const float bot = up_row(ax.l0, __ldg(r1 + ax.i0), ax.l1, __ldg(r1 + ax.i1));
    return up_col(ay.l0, top, ay.l1, bot);
}

// ---- work decomposition shared by the fused kernels (lovasz_up.cuh, confmat.cu) -------------------------------------------
__device__ __forceinline__ int up_src_row(const UpSrc& u, int Y) { return (int)__fmul_rn(u.ry, (float)Y); }
// smallest output row in [0, H] whose upper source row is >= k (the source index is monotone in the output row)
__device__ __forceinline__ int up_first_row(const UpSrc& u, int k) {
    if (k <= 0) return 0;
    if (!(u.ry > 0.f)) return u.H;
    float est = ceilf((float)k / u.ry);
    int y = est >= (float)u.H ? u.H : (int)est;
    while (y > 0 && up_src_row(u, y - 1) >= k) --y;
    while (y < u.H && up_src_row(u, y) < k) ++y;
    return y;
}

// Work item -> (image, strip, source-row interval, output rows [ya, yb)); false when the item is empty.
struct UpItem { int n, sx, k, ya, yb; };
__device__ __forceinline__ u32 up_item_count(const UpSrc& u, int N) { return (u32)(u.W / 32) * (u32)u.h * (u32)u.jmax * (u32)N; }
__device__ __forceinline__ bool up_item(const UpSrc& u, u32 item, UpItem& it) {
    const u32 nsx = (u32)(u.W / 32);
    it.sx = (int)(item % nsx);
    u32 t = item / nsx;
    const int j = (int)(t % (u32)u.jmax);
    t /= (u32)u.jmax;
    it.k = (int)(t % (u32)u.h);
    it.n = (int)(t / (u32)u.h);
    const int y0 = up_first_row(u, it.k), y1 = up_first_row(u, it.k + 1);
    it.ya = y0 + j * UP_ROWS_MAX;
    it.yb = min(y1, it.ya + UP_ROWS_MAX);
    return it.ya < it.yb;
}

// horizontal interpolation of source row `ys` of image n for the strip's 32 output columns: Hd[c][lane].
// All 2*CT loads are issued before the first use.  Left to itself the compiler sinks every load next to its multiply-add (32
// registers, ~16 loads in flight): issue is in order, so the warp stalls at the first use and the row costs several dependent
// L2 round trips -- 48 % of the samples of the confusion-matrix kernel.  The loads are volatile (kept in program order) and
// both weights are made to depend on the last pair, so no multiply-add can be scheduled before the last load is out.
__device__ __forceinline__ float ldg_ordered(const float* p) {
    float v;
    asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}
template <int CT>
__device__ __forceinline__ void up_fill_row(float (*Hd)[32], const float* __restrict__ img, int ys, const UpSrc& u,
                                            const UpAxis& ax, int lane) {
    const size_t pl = (size_t)u.h * u.w;
    const float* p0 = img + (size_t)ys * u.w + ax.i0;
    const float* p1 = img + (size_t)ys * u.w + ax.i1;
    float a[CT], b[CT];
#pragma unroll
    for (int c = 0; c < CT; ++c) { a[c] = ldg_ordered(p0); b[c] = ldg_ordered(p1); p0 += pl; p1 += pl; }
    // zero, but not to the assembler: a data dependency of both weights on the last pair of loads
    u32 z;                                                 // (popc(a) + popc(b)) >> 7 == 0; opaque to both compiler stages
    asm volatile("{\n\t.reg .u32 t0, t1;\n\tpopc.b32 t0, %1;\n\tpopc.b32 t1, %2;\n\tadd.u32 t0, t0, t1;\n\tshr.u32 %0, t0, 7;\n\t}"
                 : "=r"(z) : "r"(__float_as_uint(a[CT - 1])), "r"(__float_as_uint(b[CT - 1])));
    const float l0 = __uint_as_float(__float_as_uint(ax.l0) | z), l1 = __uint_as_float(__float_as_uint(ax.l1) | z);
#pragma unroll
    for (int c = 0; c < CT; ++c) Hd[c][lane] = up_row(l0, a[c], l1, b[c]);
}

// debug / test kernel: materialise the upsampled logits with a chosen contraction pattern
template <int LFMA, int INNER, int OUTER>
__device__ __forceinline__ float up_value_t(const UpSrc& u, const float* pl, int Y, int X) {
    const UpAxis ay = up_axis_t<LFMA>(u.ry, Y, u.h), ax = up_axis_t<LFMA>(u.rx, X, u.w);
    const float* r0 = pl + (size_t)ay.i0 * u.w;
    const float* r1 = pl + (size_t)ay.i1 * u.w;
    const float top = up_mix_t<INNER>(ax.l0, r0[ax.i0], ax.l1, r0[ax.i1]);
    const float bot = up_mix_t<INNER>(ax.l0, r1[ax.i0], ax.l1, r1[ax.i1]);
    return up_mix_t<OUTER>(ay.l0, top, ay.l1, bot);
}
static __global__ void upsample_debug_kernel(UpSrc u, long long planes, float* __restrict__ out, int pattern) {
    const long long HW = (long long)u.H * u.W, total = planes * HW;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long pc = i / HW, q = i - pc * HW;
        const int Y = (int)(q / u.W), X = (int)(q - (long long)Y * u.W);
        const float* pl = u.lo + (size_t)pc * ((size_t)u.h * u.w);
        float v;
        if (pattern < 0) {                                  // the product's own pattern
            const UpAxis ay = up_axis(u.ry, Y, u.h), ax = up_axis(u.rx, X, u.w);
            const float* r0 = pl + (size_t)ay.i0 * u.w;
            const float* r1 = pl + (size_t)ay.i1 * u.w;
            v = up_col(ay.l0, up_row(ax.l0, r0[ax.i0], ax.l1, r0[ax.i1]), ay.l1, up_row(ax.l0, r1[ax.i0], ax.l1, r1[ax.i1]));
        } else {
            const int lf = pattern & 1, in = (pattern >> 1) & 3, ou = (pattern >> 3) & 3;
#define UPV(L, I, O) if (lf == L && in == I && ou == O) v = up_value_t<L, I, O>(u, pl, Y, X);
            v = 0.f;
            UPV(0, 0, 0) UPV(0, 0, 1) UPV(0, 0, 2) UPV(0, 1, 0) UPV(0, 1, 1) UPV(0, 1, 2) UPV(0, 2, 0) UPV(0, 2, 1) UPV(0, 2, 2)
            UPV(1, 0, 0) UPV(1, 0, 1) UPV(1, 0, 2) UPV(1, 1, 0) UPV(1, 1, 1) UPV(1, 1, 2) UPV(1, 2, 0) UPV(1, 2, 1) UPV(1, 2, 2)
#undef UPV
        }
        out[i] = v;
    }
}

static inline UpSrc make_up_src(const float* lo, int h, int w, int H, int W) {
    UpSrc u;
    u.lo = lo; u.h = h; u.w = w; u.H = H; u.W = W;
    // ATen: area_pixel_compute_scale<float>(in, out, align_corners = true) = out > 1 ? (float)(in - 1) / (out - 1) : 0
    u.ry = H > 1 ? (float)(h - 1) / (float)(H - 1) : 0.f;
    u.rx = W > 1 ? (float)(w - 1) / (float)(W - 1) : 0.f;
    // output rows that share one upper source row: at most floor(1 / ry) + 1 (+1 for the fp32 rounding of ry * Y)
    const long long per_interval = u.ry > 0.f ? (long long)(1.0f / u.ry) + 2 : (long long)H;
    u.jmax = (int)((per_interval + UP_ROWS_MAX - 1) / UP_ROWS_MAX);
    if (u.jmax < 1) u.jmax = 1;
    return u;
}
