/*
 * Plain-C scalar restatement of the reference hot path.  TEST INFRASTRUCTURE ONLY -- never linked
 * into or called by the shipped package; built by oracle/Makefile into oracle/_build/.
 *
 * Follows (paths relative to /root/reference):
 *   losses/LovaszSoftmax.py:19-32   forward (softmax over C, flat or per-image)
 *   losses/LovaszSoftmax.py:34-61   lovasz_softmax_flat (per-class |fg-p|, descending sort, dot with Jaccard grad)
 *   losses/LovaszSoftmax.py:63-80   flatten_probabilities (pixel order n,h,w; optional label filter)
 *   losses/LovaszSoftmax.py:83-95   lovasz_grad (J = 1 - I/U in fp32, first difference)
 *   losses/LovaszSoftmax.py:102-120 mean (sequential sum in class / image order, divide by count)
 *   utils/torch_utils.py:221-241    t_get_confusion_matrix (argmax first-max, cm[pred][gt], ignore column dropped)
 * The backward is NOT autograd here but the closed form the CUDA kernels use:
 *   dL/dp_c(pixel at sorted position i) = -sgn(fg - p) * grad_i * w,   dz_k = p_k (G_k - sum_j G_j p_j)
 * so agreement between this file, the torch port (autograd) and the golden vectors checks that algebra.
 * Ties sort canonically: descending error, ascending flattened pixel index (torch.sort(stable=True)).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct { float err; int64_t idx; } item_t;

static int cmp_desc_err_asc_idx(const void* a, const void* b) {
    const item_t* x = (const item_t*)a;
    const item_t* y = (const item_t*)b;
    if (x->err > y->err) return -1;
    if (x->err < y->err) return 1;
    return (x->idx > y->idx) - (x->idx < y->idx);
}

/* softmax of one pixel: exp(z - max) / sum, classes summed in ascending order, all fp32 */
static void softmax_px(const float* logits, int64_t plane, int C, float* p) {
    float m = logits[0];
    for (int k = 1; k < C; ++k) { float v = logits[(int64_t)k * plane]; if (v > m) m = v; }
    float s = 0.f;
    for (int k = 0; k < C; ++k) { p[k] = expf(logits[(int64_t)k * plane] - m); s += p[k]; }
    for (int k = 0; k < C; ++k) p[k] = p[k] / s;
}

/*
 * class_mode: 0 = skip absent classes ('present'), 1 = keep absent classes ('all' / explicit list)
 * class_mask: C bytes, 1 = class is summed (all ones for 'present' / 'all')
 * filter_label: pixels with this label are removed (classes_to_ignore); pass INT64_MIN for none
 * grad (may be NULL): d loss / d logits, same layout as logits [N,C,HW]
 * returns 0, or -1 on allocation failure
 */
int oracle_lovasz(const float* logits, const int64_t* labels, int N, int C, int64_t HW,
                  int per_image, int64_t filter_label, int class_mode, const uint8_t* class_mask,
                  float* loss_out, float* grad) {
    const int64_t P = (int64_t)N * HW;
    float* prob = (float*)malloc(sizeof(float) * (size_t)P * C);       /* [P][C] */
    double* G = (double*)calloc((size_t)P * C, sizeof(double));        /* dL/dp, [P][C] */
    item_t* items = (item_t*)malloc(sizeof(item_t) * (size_t)P);
    float* jac = (float*)malloc(sizeof(float) * (size_t)P);
    if (!prob || !G || !items || !jac) { free(prob); free(G); free(items); free(jac); return -1; }

    for (int n = 0; n < N; ++n)
        for (int64_t q = 0; q < HW; ++q)
            softmax_px(logits + (int64_t)n * C * HW + q, HW, C, prob + ((int64_t)n * HW + q) * C);

    const int groups = per_image ? N : 1;
    const int64_t gsize = per_image ? HW : P;
    float group_acc = 0.f;       /* sequential fp32 sum of group losses */
    float* class_loss = (float*)malloc(sizeof(float) * (size_t)C);
    int* kept_class = (int*)malloc(sizeof(int) * (size_t)C);

    for (int g = 0; g < groups; ++g) {
        const int64_t base = (int64_t)g * gsize;
        int nkept = 0;
        int64_t nvalid = 0;
        for (int64_t i = 0; i < gsize; ++i) nvalid += (labels[base + i] != filter_label);
        for (int c = 0; c < C && nvalid > 0; ++c) {
            if (!class_mask[c]) continue;
            int64_t m = 0, gts = 0;
            for (int64_t i = 0; i < gsize; ++i) {
                const int64_t px = base + i;
                if (labels[px] == filter_label) continue;
                const int fg = labels[px] == c;
                const float fgf = fg ? 1.f : 0.f;
                items[m].err = fabsf(fgf - prob[px * C + c]);
                items[m].idx = px;
                gts += fg;
                ++m;
            }
            if (class_mode == 0 && gts == 0) continue;
            qsort(items, (size_t)m, sizeof(item_t), cmp_desc_err_asc_idx);
            /* Jaccard gradient: integer counts converted to fp32 (the reference's fp32 cumsums hold them exactly) */
            int64_t cfg = 0, cbg = 0;
            double dot = 0.0;
            float prev = 0.f;
            for (int64_t i = 0; i < m; ++i) {
                const int fg = labels[items[i].idx] == c;
                cfg += fg; cbg += !fg;
                const float inter = (float)gts - (float)cfg;
                const float uni = (float)gts + (float)cbg;
                const float j = 1.0f - inter / uni;
                jac[i] = (i == 0) ? j : (j - prev);
                prev = j;
                dot += (double)items[i].err * (double)jac[i];
            }
            class_loss[nkept] = (float)dot;
            kept_class[nkept] = c;
            ++nkept;
            /* stash unscaled dL/dp; scaled by 1/nkept (and 1/groups) below */
            for (int64_t i = 0; i < m; ++i) {
                const int64_t px = items[i].idx;
                const int fg = labels[px] == c;
                const float p = prob[px * C + c];
                const float d = (fg ? 1.f : 0.f) - p;                 /* fg - p */
                const double sgn = (d > 0.f) ? 1.0 : ((d < 0.f) ? -1.0 : 0.0);
                G[px * C + c] = -sgn * (double)jac[i];                /* d|fg-p|/dp * grad_i */
            }
        }
        float gl = 0.f;
        if (nkept > 0) {
            gl = class_loss[0];
            for (int k = 1; k < nkept; ++k) gl += class_loss[k];
            if (nkept > 1) gl = gl / (float)nkept;
        }
        const double w = (nkept > 0 ? 1.0 / (double)nkept : 0.0) * (groups > 1 ? 1.0 / (double)groups : 1.0);
        for (int64_t i = 0; i < gsize; ++i)
            for (int c = 0; c < C; ++c) G[(base + i) * C + c] *= w;
        if (g == 0) group_acc = gl; else group_acc += gl;
    }
    if (groups > 1) group_acc = group_acc / (float)groups;
    *loss_out = group_acc;

    if (grad) {
        for (int n = 0; n < N; ++n)
            for (int64_t q = 0; q < HW; ++q) {
                const int64_t px = (int64_t)n * HW + q;
                double s = 0.0;
                for (int k = 0; k < C; ++k) s += G[px * C + k] * (double)prob[px * C + k];
                for (int k = 0; k < C; ++k)
                    grad[((int64_t)n * C + k) * HW + q] = (float)((double)prob[px * C + k] * (G[px * C + k] - s));
            }
    }
    free(prob); free(G); free(items); free(jac); free(class_loss); free(kept_class);
    return 0;
}

/*
 * cm[pred * C + gt] += 1 for every pixel with 0 <= label < C.  drop_label >= 0: pixels with that label are
 * skipped silently (the ignore column the reference slices off, utils/torch_utils.py:232-234); any other label
 * outside [0, C) sets *oob (the reference's one_hot raises there).  argmax: first maximum, NaN counts as maximum.
 */
int oracle_confmat(const float* pred, const int64_t* labels, int N, int C, int64_t HW,
                   int64_t drop_label, int64_t* cm, int* oob) {
    *oob = 0;
    for (int n = 0; n < N; ++n)
        for (int64_t q = 0; q < HW; ++q) {
            const float* z = pred + (int64_t)n * C * HW + q;
            int best = 0; float bv = z[0];
            for (int k = 1; k < C; ++k) {
                const float v = z[(int64_t)k * HW];
                if (bv != bv) break;                       /* NaN already won */
                if (v > bv || v != v) { bv = v; best = k; }
            }
            const int64_t t = labels[(int64_t)n * HW + q];
            if (t == drop_label) continue;
            if (t < 0 || t >= C) { *oob = 1; continue; }
            cm[(int64_t)best * C + t] += 1;
        }
    return 0;
}
