// K1 / K6 with the bilinear upsampling of the reference's models fused in (models/OCR.py:126-131,
// models/DeepLabv3Plus.py:65-68).  Included by lovasz.cu after the full-resolution kernels (shares stats_pixel, sm_prob, ...).
//
// Work item of a warp = (image n, strip of 32 output columns, source-row interval k): the output rows whose upper source
// row is k (about `scale` of them).  The warp interpolates source rows k and k+1 horizontally ONCE (H0, H1: C x 32 floats
// each, lane = output column), then every output row of the item is one vertical mix of the two — 2 shared-memory reads
// and 2 flops per logit instead of 4 loads and 6 flops — written to the warp's tile T[c][lane], on which the unchanged
// per-pixel code of the full-resolution kernels runs.  The low-resolution logits (6.5 MB at 8 x 25 x 68 x 120) stay in L2;
// the 415 MB upsampled tensor never exists, neither do its gradient nor ATen's two interpolation kernels.
//
// Backward: the adjoint of the interpolation is separable.  Each lane accumulates its column's gradient over the item's
// rows into two register rows (weights l0y / l1y: source rows k and k+1), then the warp multiplies by the horizontal
// weights through a small table in shared memory (columns of the strip x 32 lanes) and adds the C x ~6 sums per source row
// to the low-resolution gradient with float atomics (rows are shared by two intervals, columns by two strips).  Like
// ATen's upsample_bilinear2d_backward the summation order is not fixed, so the low-resolution gradient is reproducible to
// rounding, not bit for bit.
#pragma once

#define UP_MAX_COLS 12               // source columns a 32-pixel strip may touch (scale factors >= ~3.2 along x)

// ---- K1 -----------------------------------------------------------------------------------------------------------------------
template <int CT, int TPB, typename LT>
__global__ void __launch_bounds__(TPB) stats_kernel_up(LovaszParams p) {
    constexpr int NW = TPB / 32;
    extern __shared__ __align__(16) unsigned char pipe_smem_raw[];
    __shared__ u32 s_cm[B200SEG_MAX_CLASSES * B200SEG_MAX_CLASSES];
    __shared__ u32 s_fg[NW][B200SEG_MAX_CLASSES], s_key[NW][B200SEG_MAX_CLASSES];
    __shared__ u32 s_valid, s_oob;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    typedef float Tile[CT][32];
    Tile* tiles = reinterpret_cast<Tile*>(pipe_smem_raw) + (size_t)warp * 3;     // H0, H1, T
    for (int i = tid; i < CT * CT; i += TPB) s_cm[i] = 0;
    s_fg[warp][lane] = 0; s_key[warp][lane] = 0;
    if (tid == 0) { s_valid = 0; s_oob = 0; }
    __syncthreads();

    const UpSrc u = p.up;
    const u32 items = up_item_count(u, p.N);
    const u32 gw = blockIdx.x * NW + warp, nwarps = gridDim.x * NW;
    const size_t img_lo = (size_t)CT * u.h * u.w;

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
    StatsAcc A;
    ExpConsts ek;
    ek.load();
    for (u32 item = gw; item < items; item += nwarps) {
        UpItem wi;
        if (!up_item(u, item, wi)) continue;               // (warp-uniform: no output row of this chunk has this upper source row)
        const int sx = wi.sx, k = wi.k, n = wi.n, y0 = wi.ya, y1 = wi.yb;
        const int g = p.per_image ? n : 0;
        if (g != cur_g) {
            if (cur_g >= 0) { flush_group(cur_g, A.nvalid); A.nvalid = 0; }
            cur_g = g;
        }
        const int X = sx * 32 + lane;
        const UpAxis ax = up_axis(u.rx, X, u.w);
        const float* img = u.lo + (size_t)n * img_lo;
        const int k1 = k + ((k < u.h - 1) ? 1 : 0);
        up_fill_row<CT>(tiles[0], img, k, u, ax, lane);
        if (k1 != k) up_fill_row<CT>(tiles[1], img, k1, u, ax, lane);
        const float (*H0)[32] = tiles[0];
        const float (*H1)[32] = (k1 != k) ? tiles[1] : tiles[0];
        float (*T)[32] = tiles[2];
        size_t px = (size_t)n * p.HW + (size_t)y0 * u.W + X;
        int lab_next = load_label<LT>(p.labels, px);
        for (int Y = y0; Y < y1; ++Y, px += u.W) {
            const int lab = lab_next;
            if (Y + 1 < y1) lab_next = load_label<LT>(p.labels, px + u.W);
            const UpAxis ay = up_axis(u.ry, Y, u.h);       // ay.i0 == k, ay.i1 == k1
#pragma unroll
            for (int c = 0; c < CT; ++c) T[c][lane] = up_col(ay.l0, H0[c][lane], ay.l1, H1[c][lane]);
            stats_pixel<CT, 32>(p, T, lane, lab, px, s_fg[warp], s_key[warp], s_cm, ek, A);
        }
    }
    if (p.ce_enabled) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { A.ce_acc += __shfl_xor_sync(FULL_MASK, A.ce_acc, o); A.ce_n += __shfl_xor_sync(FULL_MASK, A.ce_n, o); }
        if (lane == 0 && A.ce_n) { atomicAdd(p.ce_sum, (double)A.ce_acc); atomicAdd(p.ce_cnt, A.ce_n); }
        if (__any_sync(FULL_MASK, A.ce_oob) && lane == 0 && p.status) atomicOr(p.status, STATUS_LABEL_OOB);
    }
    if (cur_g >= 0 && p.per_image) { flush_group(cur_g, A.nvalid); A.nvalid = 0; }
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
            u32 f = 0, kk = 0;
#pragma unroll
            for (int w = 0; w < NW; ++w) { f += s_fg[w][tid]; kk = max(kk, s_key[w][tid]); }
            if (f) { atomicAdd(p.seg_fg + tid, f); atomicMax(p.seg_maxkey + tid, kk); }
        }
        if (tid == 0 && s_valid) atomicAdd(p.grp_valid, s_valid);
    }
    if (p.cm) {
        for (int i = tid; i < CT * CT; i += TPB)
            if (s_cm[i]) atomicAdd(p.cm + i, (unsigned long long)s_cm[i]);
        if (tid == 0 && s_oob) atomicOr(p.status, STATUS_LABEL_OOB);
    }
}

// ---- K2' (streaming emission over interpolated logits; runs when K1b chose the streaming path) ---------------------------------
template <int CT, int TPB>
__global__ void __launch_bounds__(TPB) emit_kernel_up(LovaszParams p) {
    if (p.flags[0] != EMIT_PATH_STREAM) return;
    constexpr int NW = TPB / 32;
    extern __shared__ __align__(16) unsigned char pipe_smem_raw[];
    __shared__ float s_thr[NW][B200SEG_MAX_CLASSES];
    __shared__ __align__(16) float s_logthr[NW][B200SEG_MAX_CLASSES];
    __shared__ u32 s_mask[NW][B200SEG_MAX_CLASSES], s_base[NW][B200SEG_MAX_CLASSES];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const u32 lt_mask = (1u << lane) - 1;
    typedef float Tile[CT][32];
    Tile* tiles = reinterpret_cast<Tile*>(pipe_smem_raw) + (size_t)warp * 3;     // H0, H1, T
    const UpSrc u = p.up;
    const u32 items = up_item_count(u, p.N);
    const u32 gw = blockIdx.x * NW + warp, nwarps = gridDim.x * NW;
    const size_t img_lo = (size_t)CT * u.h * u.w;
    int cur_g = -1;
    for (u32 item = gw; item < items; item += nwarps) {
        UpItem wi;
        if (!up_item(u, item, wi)) continue;               // (warp-uniform: no output row of this chunk has this upper source row)
        const int sx = wi.sx, k = wi.k, n = wi.n, y0 = wi.ya, y1 = wi.yb;
        const int g = p.per_image ? n : 0;
        if (g != cur_g) {
            __syncwarp();
            if (lane < CT) { s_thr[warp][lane] = p.seg_thr[(size_t)g * CT + lane]; s_logthr[warp][lane] = p.seg_logthr[(size_t)g * CT + lane]; }
            cur_g = g;
            __syncwarp();
        }
        const int X = sx * 32 + lane;
        const UpAxis ax = up_axis(u.rx, X, u.w);
        const float* img = u.lo + (size_t)n * img_lo;
        const int k1 = k + ((k < u.h - 1) ? 1 : 0);
        up_fill_row<CT>(tiles[0], img, k, u, ax, lane);
        if (k1 != k) up_fill_row<CT>(tiles[1], img, k1, u, ax, lane);
        const float (*H0)[32] = tiles[0];
        const float (*H1)[32] = (k1 != k) ? tiles[1] : tiles[0];
        float (*T)[32] = tiles[2];
        size_t px = (size_t)n * p.HW + (size_t)y0 * u.W + X;
        float m_n = p.pix_m[px], s_n = p.pix_s[px];
        u32 l8_n = p.lab8[px];
        for (int Y = y0; Y < y1; ++Y, px += u.W) {
            const float m = m_n, s = s_n;
            const u32 l8 = l8_n;
            if (Y + 1 < y1) { m_n = p.pix_m[px + u.W]; s_n = p.pix_s[px + u.W]; l8_n = p.lab8[px + u.W]; }
            const UpAxis ay = up_axis(u.ry, Y, u.h);
#pragma unroll
            for (int c = 0; c < CT; ++c) T[c][lane] = up_col(ay.l0, H0[c][lane], ay.l1, H1[c][lane]);
            emit_tile<CT, 32>(p, T, lane, lt_mask, m, s, l8, true, px, g, s_thr[warp], s_logthr[warp], s_mask[warp], s_base[warp]);
        }
    }
}

// ---- K6 -----------------------------------------------------------------------------------------------------------------------
template <int CT>
struct UpBwdSmem {
    float H0[CT][32], H1[CT][32];                          // horizontally interpolated source rows; reused as the staging of a flush
    float T[CT][32];                                       // logits of the row; exact gradients replace the logits of their classes
    float wt[UP_MAX_COLS][32];                             // horizontal weight of lane j's column for source column xs + i
    float thr[B200SEG_MAX_CLASSES];
    unsigned char jlo[UP_MAX_COLS], jhi[UP_MAX_COLS];     // lanes [jlo, jhi) touch the column
    unsigned char pad_[8];
};

template <int CT, int TPB>
__global__ void __launch_bounds__(TPB, 4) backward_kernel_up(LovaszParams p, const float* __restrict__ go,
                                                              const float* __restrict__ go_ce, float* __restrict__ dlow) {
    constexpr int NW = TPB / 32;
    constexpr int PS = 33;                                 // row stride of the flush staging (bank-conflict-free columns)
    static_assert(2 * CT * 32 >= CT * PS, "staging fits H0 + H1");
    extern __shared__ __align__(16) unsigned char pipe_smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    UpBwdSmem<CT>& S = reinterpret_cast<UpBwdSmem<CT>*>(pipe_smem_raw)[warp];
    const UpSrc u = p.up;
    const u32 items = up_item_count(u, p.N);
    const u32 gw = blockIdx.x * NW + warp, nwarps = gridDim.x * NW;
    const size_t img_lo = (size_t)CT * u.h * u.w, pl_lo = (size_t)u.h * u.w;
    const float gsc = __ldg(go);
    const float gce = (p.ce_enabled && go_ce) ? __ldg(go_ce) * *p.ce_inv_n : 0.f;
    const size_t plane = (size_t)p.HW;
    int cur_g = -1;

    for (u32 item = gw; item < items; item += nwarps) {
        UpItem wi;
        if (!up_item(u, item, wi)) continue;               // (warp-uniform: no output row of this chunk has this upper source row)
        const int sx = wi.sx, k = wi.k, n = wi.n, y0 = wi.ya, y1 = wi.yb;
        const int g = p.per_image ? n : 0;
        if (g != cur_g) {
            __syncwarp();
            if (lane < CT) S.thr[lane] = p.seg_thr[(size_t)g * CT + lane];
            cur_g = g;
            __syncwarp();
        }
        const int X = sx * 32 + lane;
        const UpAxis ax = up_axis(u.rx, X, u.w);
        const float* img = u.lo + (size_t)n * img_lo;
        const int k1 = k + ((k < u.h - 1) ? 1 : 0);
        up_fill_row<CT>(S.H0, img, k, u, ax, lane);
        if (k1 != k) up_fill_row<CT>(S.H1, img, k1, u, ax, lane);
        const float (*H0)[32] = S.H0;
        const float (*H1)[32] = (k1 != k) ? S.H1 : S.H0;
        float (*T)[32] = S.T;
        float acc0[CT], acc1[CT];
#pragma unroll
        for (int c = 0; c < CT; ++c) { acc0[c] = 0.f; acc1[c] = 0.f; }

        size_t px = (size_t)n * plane + (size_t)y0 * u.W + X;
        // per-pixel state one row ahead of the math
        float m_n = p.pix_m[px], s_n = p.pix_s[px], gl_n = p.gown[px];
        u32 l8_n = p.lab8[px], mask_n = p.cmask[px];
        for (int Y = y0; Y < y1; ++Y, px += u.W) {
            const float m = m_n, s = s_n, gl = gl_n;
            const u32 l8 = l8_n, cmask_cur = mask_n;
            const long long q = (long long)Y * u.W + X;
            const float* gb = p.gbg + (size_t)n * CT * plane + q;
            // gradients of the first candidates: requested now, used after the tile is formed
            constexpr int NX = 6;
            float g1 = 0.f, g2 = 0.f, gx[NX];
#pragma unroll
            for (int j = 0; j < NX; ++j) gx[j] = 0.f;
            if (cmask_cur) {
                u32 rest = cmask_cur;
                g1 = gb[(size_t)(__ffs(rest) - 1) * plane]; rest &= rest - 1;
                if (rest) { g2 = gb[(size_t)(__ffs(rest) - 1) * plane]; rest &= rest - 1; }
#pragma unroll
                for (int j = 0; j < NX; ++j)
                    if (rest) { gx[j] = gb[(size_t)(__ffs(rest) - 1) * plane]; rest &= rest - 1; }
            }
            if (Y + 1 < y1) {
                const size_t pn = px + u.W;
                m_n = p.pix_m[pn]; s_n = p.pix_s[pn]; gl_n = p.gown[pn]; l8_n = p.lab8[pn]; mask_n = p.cmask[pn];
            }
            const UpAxis ay = up_axis(u.ry, Y, u.h);
#pragma unroll
            for (int c = 0; c < CT; ++c) T[c][lane] = up_col(ay.l0, H0[c][lane], ay.l1, H1[c][lane]);

            const bool filt = l8 == LAB8_FILTERED;
            int lab = l8 < (u32)CT ? (int)l8 : -1;
            const float gcp = (lab >= 0 && !(p.has_ce_ignore && lab == p.ce_ignore)) ? gce : 0.f;
            const int ce_lab = lab;
            if (lab >= 0 && !thr_active(S.thr[lab])) lab = -1;
            float d = 0.f, pk1 = 0.f, pk2 = 0.f;
            if (cmask_cur) {
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
                ownv = gsc * pown * (gl - d) + gcp * (pown - 1.0f);
            }
            const float inv_s = __fdiv_rn(1.0f, s);
            const float nd = (filt ? 0.f : -gsc * d * inv_s) + gcp * inv_s;
            // every class gets -go * p_k * dot with the fast exponential; the classes in `exact` take an exact value instead,
            // which replaces their logit in the tile (each needs only its own logit, read before it is overwritten)
            u32 exact = cmask_cur;
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
                    T[c][lane] = gsc * pk * (gk - d) + gcp * pk;
                    ++i;
                }
            }
            if (lab >= 0) { T[lab][lane] = ownv; exact |= 1u << lab; }
            // cross-entropy only: the label's class is not summed by the Lovasz term, its "- 1" is applied here
            if (gcp != 0.f && lab < 0) { T[ce_lab][lane] = nd * __expf(T[ce_lab][lane] - m) - gcp; exact |= 1u << ce_lab; }
            if (p.dbg & 256) {                             // debugging aid: the full-resolution gradient, into the dead sort buffer B
                float* dbg = reinterpret_cast<float*>(p.keysB) + (size_t)n * CT * plane + q;
#pragma unroll
                for (int c = 0; c < CT; ++c) dbg[(size_t)c * plane] = ((exact >> c) & 1u) ? T[c][lane] : nd * __expf(T[c][lane] - m);
            }
            // vertical part of the adjoint: this row's gradient goes to source rows k (l0) and k1 (l1)
#pragma unroll
            for (int c = 0; c < CT; ++c) {
                const float t = T[c][lane];
                const float dv = ((exact >> c) & 1u) ? t : nd * __expf(t - m);
                acc0[c] = fmaf(ay.l0, dv, acc0[c]);
                acc1[c] = fmaf(ay.l1, dv, acc1[c]);
            }
        }

        // ---- flush: horizontal part of the adjoint, then atomics into the low-resolution gradient --------------------------------
        const int xs = __shfl_sync(FULL_MASK, ax.i0, 0);
        const int ncols = __shfl_sync(FULL_MASK, ax.i1, 31) - xs + 1;            // <= UP_MAX_COLS (checked by the host)
        __syncwarp();
        for (int i = 0; i < ncols; ++i) {
            const int col = xs + i;
            const bool hit0 = ax.i0 == col, hit1 = ax.i1 == col;
            S.wt[i][lane] = (hit0 ? ax.l0 : 0.f) + (hit1 ? ax.l1 : 0.f);
            // lanes that touch the column form one contiguous range (the source index is monotone in the output column)
            const u32 touch = __ballot_sync(FULL_MASK, hit0 || hit1);
            if (lane == 0) { S.jlo[i] = (unsigned char)(__ffs(touch) - 1); S.jhi[i] = (unsigned char)(32 - __clz(touch)); }
        }
        float* stg = &S.H0[0][0];                                              // [CT][PS], spans H0 and H1 (dead once the rows are done)
        float* out_n = dlow + (size_t)n * img_lo;
        const int nout = CT * ncols;
#pragma unroll 1
        for (int half = 0; half < 2; ++half) {             // source row k, then k1 (the same row at the bottom edge: both are added)
            __syncwarp();
#pragma unroll
            for (int c = 0; c < CT; ++c) stg[c * PS + lane] = half == 0 ? acc0[c] : acc1[c];
            __syncwarp();
            const int ys = half == 0 ? k : k1;
            for (int o = lane; o < nout; o += 32) {
                const int c = o / ncols, i = o - c * ncols;
                const float* wrow = S.wt[i];
                const float* srow = stg + c * PS;
                float sum = 0.f;
                const int j1 = S.jhi[i];
                for (int j = S.jlo[i]; j < j1; ++j) sum = fmaf(wrow[j], srow[j], sum);
                if (sum != 0.f) atomicAdd(out_n + (size_t)c * pl_lo + (size_t)ys * u.w + xs + i, sum);
            }
        }
        __syncwarp();
    }
}
