/*
 * b200seg -- C ABI of the B200-native (sm_100a) Lovasz-Softmax + confusion-matrix mIoU hot path.
 *
 * This header is the drop-in boundary.  Each entry point names the reference interface it replaces
 * (paths relative to RViMLab/MICCAI2021_Cataract_semantic_segmentation).  The reference has no FFI of
 * its own (it is pure PyTorch); INTEGRATION.md shows the ctypes stub a maintainer would add.
 *
 * Conventions
 *   - every pointer except `bytes`/`info` outputs is a DEVICE pointer owned by the caller
 *     (on the PyTorch side: tensors from the caching allocator, workspace included);
 *   - every call is asynchronous on `stream` (pass torch.cuda.current_stream().cuda_stream);
 *     no internal allocation, no host synchronisation, no device switch;
 *   - return value: 0 = ok, <0 = invalid argument (B200SEG_E_*), >0 = cudaError_t of a failed launch;
 *     b200seg_last_error() returns a thread-local message for the last non-zero return;
 *   - logits / prediction: fp32, contiguous NCHW, `plane` = H*W elements per (n, c) plane;
 *   - labels: contiguous [N, H, W] of dtype `label_dtype`; values are compared after saturation to int32;
 *   - limits: 1 <= n_classes <= 32, n_images*plane < 2^30, n_images*plane*n_classes < 2^31.
 */
#ifndef B200SEG_H_
#define B200SEG_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200SEG_VERSION 1

/* label dtypes */
#define B200SEG_LABEL_U8 0
#define B200SEG_LABEL_I32 1
#define B200SEG_LABEL_I64 2

/* "no such label" for filter_label / drop_label */
#define B200SEG_NO_LABEL INT64_MIN

/* negative return codes */
#define B200SEG_E_INVALID (-1)   /* bad argument (null pointer, size out of range, unknown dtype) */
#define B200SEG_E_WORKSPACE (-2) /* workspace too small or misaligned */
#define B200SEG_E_UNSUPPORTED (-3) /* valid request that this entry point does not cover (see its comment) */

/* bits of the device-side status word (`status`, int32, OR-ed into by kernels; caller zeroes it) */
#define B200SEG_STATUS_LABEL_OOB 1 /* a label outside [0, C) other than drop_label reached the confusion matrix
                                       (the reference's one_hot raises RuntimeError there) */
#define B200SEG_STATUS_SPIN_TIMEOUT 2 /* internal grid-barrier watchdog fired (results invalid; a bug) */

int b200seg_version(void);
const char* b200seg_last_error(void);

/* ------------------------------------------------------------------------------------------------
 * Lovasz-Softmax  --  replaces LovaszSoftmax.forward + its autograd backward
 *   reference: losses/LovaszSoftmax.py:19-32 (forward), :34-61 (lovasz_softmax_flat),
 *              :63-80 (flatten_probabilities), :83-95 (lovasz_grad), :102-120 (mean)
 *
 *   per_image     LovaszSoftmax.py:14,27-29   0 = one loss over the whole batch, 1 = mean of per-image losses
 *   filter_label  LovaszSoftmax.py:15,74-80   `classes_to_ignore`: pixels with this label are removed;
 *                                             B200SEG_NO_LABEL = keep every pixel
 *   keep_absent   LovaszSoftmax.py:53         0 = 'present' (skip classes without foreground),
 *                                             1 = 'all' / explicit list (absent classes contribute max p_c)
 *   class_mask    LovaszSoftmax.py:46         bit c set = class c is summed ('all'/'present': all C bits)
 * ------------------------------------------------------------------------------------------------ */

/* Bytes of caller-provided scratch for one forward(+backward) call; sized for the worst case in which every
 * (pixel, class) pair is a sort candidate.  The same workspace must be handed, untouched, to backward. */
int b200seg_lovasz_workspace_bytes(int32_t n_images, int32_t n_classes, int64_t plane, int32_t per_image,
                                   size_t* bytes);

/* loss_out[0] (device, fp32) <- Lovasz-Softmax loss.
 * If cm != NULL the confusion matrix of (argmax_c logits, labels) is accumulated into cm[C*C] (int64,
 * cm[pred*C+gt]) in the same pass over the logits (fused metrics; see b200seg_confmat_accumulate for
 * drop_label / status semantics).  `status` may be NULL only if cm is NULL.
 * need_grad = 0 skips the work only backward needs (validation under torch.no_grad()). */
int b200seg_lovasz_forward(const float* logits, const void* labels, int32_t label_dtype,
                           int32_t n_images, int32_t n_classes, int64_t plane,
                           int32_t per_image, int64_t filter_label, int32_t keep_absent, uint32_t class_mask,
                           int32_t need_grad, void* workspace, size_t workspace_bytes,
                           float* loss_out, int64_t* cm, int64_t cm_drop_label, int32_t* status, void* stream);

/* dlogits[N,C,H,W] <- grad_out[0] * d loss / d logits, using the state forward left in `workspace`.
 * Same logits / labels / shape / mode arguments as the forward call.  grad_out is a DEVICE scalar. */
int b200seg_lovasz_backward(const float* logits, const void* labels, int32_t label_dtype,
                            int32_t n_images, int32_t n_classes, int64_t plane,
                            int32_t per_image, int64_t filter_label, int32_t keep_absent, uint32_t class_mask,
                            const void* workspace, size_t workspace_bytes,
                            const float* grad_out, float* dlogits, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Lovasz-Softmax + cross entropy in one pass  --  replaces the pair LossWrapper evaluates on the same logits
 *   reference: losses/LossWrapper.py:17-24 (nn.CrossEntropyLoss(ignore_index = 17 | 25 | -100)), :43-73
 *              (loss = sum_k w_k * loss_k over {'CrossEntropyLoss', 'LovaszSoftmax'})
 *
 * Same arguments as b200seg_lovasz_forward / _backward plus:
 *   ce_ignore_index   pixels with this label do not enter the cross entropy (B200SEG_NO_LABEL = none);
 *                     any other label outside [0, C) sets B200SEG_STATUS_LABEL_OOB (torch raises there)
 *   ce_out[0]         (device, fp32) <- mean over the non-ignored pixels of -log softmax(logits)[label]
 *                     (NaN when every pixel is ignored, like torch)
 *   grad_lovasz, grad_ce   DEVICE scalars: dlogits <- grad_lovasz * dLovasz/dlogits + grad_ce * dCE/dlogits
 * The cross entropy shares the softmax max / denominator of the first kernel and its gradient rides in the backward
 * kernel's single pass over the logits: no extra HBM traffic.  Only the pipelined kernels carry it
 * (b200seg_lovasz_ce_supported: C in {8, 17, 25}, plane % 16 == 0, 16-byte aligned tensors, `status` required);
 * otherwise B200SEG_E_UNSUPPORTED and the caller evaluates the cross entropy separately.  Also unsupported:
 * filter_label naming a real class that the cross entropy keeps.
 * ------------------------------------------------------------------------------------------------ */
int b200seg_lovasz_ce_supported(const float* logits, const void* labels, int32_t label_dtype,
                                int32_t n_images, int32_t n_classes, int64_t plane, const float* dlogits);
int b200seg_lovasz_ce_forward(const float* logits, const void* labels, int32_t label_dtype,
                              int32_t n_images, int32_t n_classes, int64_t plane,
                              int32_t per_image, int64_t filter_label, int32_t keep_absent, uint32_t class_mask,
                              int32_t need_grad, void* workspace, size_t workspace_bytes,
                              float* loss_out, int64_t ce_ignore_index, float* ce_out,
                              int64_t* cm, int64_t cm_drop_label, int32_t* status, void* stream);
int b200seg_lovasz_ce_backward(const float* logits, const void* labels, int32_t label_dtype,
                               int32_t n_images, int32_t n_classes, int64_t plane,
                               int32_t per_image, int64_t filter_label, int32_t keep_absent, uint32_t class_mask,
                               const void* workspace, size_t workspace_bytes,
                               const float* grad_lovasz, int64_t ce_ignore_index, const float* grad_ce,
                               float* dlogits, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Confusion matrix  --  replaces t_get_confusion_matrix
 *   reference: utils/torch_utils.py:221-241 (torch), utils/metrics.py:5-25 (numpy twin)
 *
 * cm[pred*C + gt] += #{pixels: argmax_c prediction == pred, label == gt}, int64, accumulated in place
 * (the reference's `existing_matrix` argument).  argmax takes the first maximum; NaN counts as maximum.
 * drop_label: pixels with this label are skipped (the ignore column the reference slices off for
 * C in {17, 25}); any other label outside [0, C) sets B200SEG_STATUS_LABEL_OOB in *status and is skipped.
 * ------------------------------------------------------------------------------------------------ */
int b200seg_confmat_accumulate(const float* prediction, const void* labels, int32_t label_dtype,
                               int32_t n_images, int32_t n_classes, int64_t plane,
                               int64_t drop_label, int64_t* cm, int32_t* status, void* stream);

/* ------------------------------------------------------------------------------------------------
 * IoU / accuracy summary on the device  --  replaces the C x C arithmetic of
 *   t_get_mean_iou / t_get_miou (utils/torch_utils.py:274-332) and t_get_pixel_accuracy (:259-271)
 *
 * iou_out[C] <- diag / (colsum + rowsum - diag) in fp32 with 0/0 -> 0 (:321-327);
 * summary_out[0] = mean IoU over the classes whose bit is set in miou_mask, [1] = PA, [2] = PAC,
 * [3 + k] = mean IoU over set k of `category_masks[n_sets]` (anatomies / instruments / rare ...).
 * Sums are formed in int64 and converted once, so they equal the reference's fp32 sums whenever those are
 * exact (every partial sum < 2^24).
 * ------------------------------------------------------------------------------------------------ */
int b200seg_metrics_from_confmat(const int64_t* cm, int32_t n_classes, uint32_t miou_mask,
                                 const uint32_t* category_masks /* host pointer */, int32_t n_sets,
                                 float* iou_out, float* summary_out, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Windowed mean IoU map  --  replaces sliding_miou (utils/torch_utils.py:189-218, called per image by
 *   calculate_performance, :24-35, with kernel 7 / stride 4 from utils/defaults.py:335-336)
 *
 * out[n][i][j] (fp32, [N, (H-k)/s+1, (W-k)/s+1]) <- mean over the C classes of I_c / U_c inside the k x k window
 * whose top-left corner is (i*s, j*s); I_c / U_c = pixels where argmax prediction AND / OR label equal c; a class
 * absent from both counts as 1 (:203-204).  Counts are exact integers; the class mean is summed in class order.
 * kernel_size must be odd (:191) and <= 255.  Labels outside [0, C) (the reference's scatter_ rejects them) set
 * B200SEG_STATUS_LABEL_OOB in *status and match no class.  `scratch` holds the per-pixel class map.
 * The `original_size` expansion (:208-214: repeat + pad) is index arithmetic left to the host side.
 * ------------------------------------------------------------------------------------------------ */
int b200seg_sliding_miou_scratch_bytes(int32_t n_images, int64_t height, int64_t width, size_t* bytes);
int b200seg_sliding_miou(const float* prediction, const void* labels, int32_t label_dtype,
                         int32_t n_images, int32_t n_classes, int32_t height, int32_t width,
                         int32_t kernel_size, int32_t stride, void* scratch, size_t scratch_bytes,
                         float* out, int32_t* status, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Online hard example mining cross entropy  --  replaces OhemCrossEntropy.forward + its autograd backward
 *   reference: losses/OhemCrossEntropy.py:22-40 (forward), :9-20 (thresh / min_kept / ignore_label)
 *
 * loss_out[0] (device, fp32) <- mean of -log p_label over the pixels with p_label < max(v, thresh), where v is the
 * element of rank min(min_kept, n_valid - 1) (0-based) of the ascending label probabilities of the n_valid pixels
 * whose label != ignore_label (:32-39).  The full-length sort of the reference only serves to read v: here it is a
 * three-level radix select over the fp32 bit patterns.  No pixel below the threshold (or n_valid == 0, where the
 * reference raises) gives 0/0 = NaN like torch's mean of an empty tensor.  Labels outside [0, C) other than
 * ignore_label set B200SEG_STATUS_LABEL_OOB and are treated as ignored.  `min_kept` is the already clamped value
 * (max(1, config) at :13).  The bilinear resize of :23-26 (score and target of different size) is left to the caller.
 * The workspace must reach backward untouched.  dlogits <- grad_out[0] / n_kept * (softmax - onehot) on the kept
 * pixels, 0 elsewhere.
 * ------------------------------------------------------------------------------------------------ */
int b200seg_ohem_workspace_bytes(int32_t n_images, int64_t plane, size_t* bytes);
int b200seg_ohem_ce_forward(const float* logits, const void* labels, int32_t label_dtype,
                            int32_t n_images, int32_t n_classes, int64_t plane, int64_t ignore_label,
                            float thresh, int64_t min_kept, void* workspace, size_t workspace_bytes,
                            float* loss_out, int32_t* status, void* stream);
int b200seg_ohem_ce_backward(const float* logits, const void* labels, int32_t label_dtype,
                             int32_t n_images, int32_t n_classes, int64_t plane, int64_t ignore_label,
                             const void* workspace, size_t workspace_bytes, const float* grad_out,
                             float* dlogits, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Measurement hook (no reference counterpart): process-wide, hand the library up to B200SEG_N_STAGES
 * cudaEvent_t handles; while set, b200seg_lovasz_forward / _backward record events[i] on their stream at stage
 * boundary i, so a caller can time each kernel group with cudaEventElapsedTime without a profiler.
 *   0 forward entry        1 stats kernel done          2 threshold finalisation done
 *   3 candidate emission (+ run scan) done                4 sort plan + tile descriptors done
 *   5 / 6 / 7 hybrid path: bucket histogram / partition / local rank + Jaccard + loss done
 *             (LSD path, B200SEG_SORT_PATH=1: key passes 0 / 1 / 2 done, 7 includes the per-tile foreground count)
 *   8 hybrid path: overflow fallback done (LSD path: Jaccard + loss done)  9 backward entry            10 backward kernel done
 * Pass NULL / 0 to clear.  Entries that are NULL are skipped.
 * ------------------------------------------------------------------------------------------------ */
#define B200SEG_N_STAGES 11
int b200seg_set_stage_events(void* const* events, int32_t n_events);

/* ------------------------------------------------------------------------------------------------
 * Data-parallel hook: an event (cudaEvent_t, owned by the caller) that b200seg_lovasz_forward /
 * b200seg_lovasz_ce_forward record on their stream right after the first kernel, i.e. as soon as the
 * fused confusion matrix and the label-range flag of the call are complete.  The host side makes its
 * all-reduce of the matrix wait for THIS event instead of for the whole forward pass, so the collective
 * runs under the emission / sort kernels (which leave room on the SMs) rather than queueing behind the
 * backward kernel (which fills them).  Replaces nothing in the reference (single-GPU there,
 * managers/BaseManager.py:83-86).  NULL clears.  Process-wide, like the stage events.
 * ------------------------------------------------------------------------------------------------ */
int b200seg_set_confmat_event(void* event);

/* ------------------------------------------------------------------------------------------------
 * Test hook: the segmented stable radix sort used inside b200seg_lovasz_forward, exposed so tests can
 * compare it bit for bit with a stable CPU sort.  Segment s occupies [s*capacity, s*capacity + counts[s])
 * of keys_in/vals_in (uint32); keys are sorted ascending on their low `key_bits[s]` bits (1..30), ties keep
 * input order.  Output lands in keys_out/vals_out with the same segment layout.  The four arrays must not
 * overlap; keys_in/vals_in are clobbered.  `scratch` from b200seg_sort_scratch_bytes.
 * ------------------------------------------------------------------------------------------------ */
int b200seg_sort_scratch_bytes(int32_t n_segments, int64_t capacity, size_t* bytes);
int b200seg_sort_segments(uint32_t* keys_in, uint32_t* vals_in, uint32_t* keys_out, uint32_t* vals_out,
                          const uint32_t* counts, const uint32_t* key_bits, int32_t n_segments, int64_t capacity,
                          void* scratch, size_t scratch_bytes, int32_t* status, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Kernel-selection knobs (no reference counterpart; process-wide).  Every setting computes the same results;
 * they exist so tests can exercise every code path and profiles can compare variants in one process.
 *   "interleave"    1 (default) warps of the streaming kernels take interleaved tiles, 0 contiguous ranges
 *   "stats_variant" 0 (default) pipelined stats kernel, 384-thread CTAs, 2 ring stages; 2..6: other CTA sizes /
 *                   stage counts; 1: register-tile kernel without per-pixel records (forces the streaming emission);
 *                   7: the same kernel fed by TMA (3-D tensor map, cp.async.bulk.tensor + mbarrier)
 *   "emit_path"     0 (default) chosen on the device from the records; 1 record-driven; 2 streaming
 *   "sort_path"     0 (default) hybrid: one most-significant-digit partition pass + a shared-memory ranking kernel fused
 *                   with the Jaccard gradient (segments whose buckets overflow shared memory fall back to 1);
 *                   1: stable LSD radix sort by (key, value) in seven passes + a separate Jaccard kernel
 *   "sort_match"    (LSD path) 2 (default) MATCH.ANY peer masks in the top digit pass only; 0 ballots; 1 MATCH.ANY
 *   "pdl"           1 (default) the kernels of the forward chain are launched with programmatic dependent launch
 *                   (cudaLaunchAttributeProgrammaticStreamSerialization): a kernel's CTAs are scheduled while the previous
 *                   kernel drains and wait (griddepcontrol.wait) for its completion before touching memory; 0 plain launches
 *   "dbg"           timing experiments only (skips work: results become wrong)
 * Initial values come from the environment variables B200SEG_INTERLEAVE, B200SEG_STATS_VARIANT,
 * B200SEG_EMIT_PATH, B200SEG_SORT_PATH, B200SEG_SORT_MATCH, B200SEG_PDL, B200SEG_DBG.
 * ------------------------------------------------------------------------------------------------ */
int b200seg_set_tuning(const char* key, int32_t value);

/* Test hook: byte offsets inside the Lovasz workspace of {pix_m, pix_s, label8, candidate mask, rec16, rec4,
 * seg_thr, grp_tmin} (see DESIGN.md "Data layout"), so tests can check the per-pixel candidate records that
 * b200seg_lovasz_forward leaves behind against a straightforward softmax / top-k. */
int b200seg_debug_layout(int32_t n_images, int32_t n_classes, int64_t plane, int32_t per_image,
                         size_t* offsets, int32_t n_offsets);

/* Test hook: *mismatches (device int32, caller-zeroed) += number of x[i] for which the exponential the stats kernel
 * uses (expf()'s instruction sequence with two constants held in registers) differs from expf(x[i]) in any bit. */
int b200seg_debug_exp_mismatches(const float* x, int32_t n, int32_t* mismatches, void* stream);

/* ---- Lovasz-Softmax (+ cross entropy, + confusion matrix) from LOW-RESOLUTION logits ----------------------
 * The reference's models upsample their stride-8 / stride-4 logits with F.interpolate(size=input_resolution,
 * mode='bilinear', align_corners=True) right before the loss (models/OCR.py:126-131, models/DeepLabv3Plus.py:65-68).
 * These entry points take lowres[n_images, n_classes, h, w] (fp32, contiguous) and labels[n_images, H, W] and compute
 * exactly what b200seg_lovasz_forward / _ce_forward compute on the upsampled tensor -- the interpolation is ATen's CUDA
 * arithmetic bit for bit, so loss, confusion matrix and tie order are unchanged -- without forming that tensor; the
 * backward pass returns the gradient with respect to `lowres` (dlowres[n_images, n_classes, h, w]; float atomics, like
 * ATen's upsample_bilinear2d_backward: reproducible to rounding).  Workspace: b200seg_lovasz_workspace_bytes(n_images,
 * n_classes, H * W, per_image).  ce_enabled != 0 adds the cross-entropy term (ce_out, status required) as
 * b200seg_lovasz_ce_forward does.  Covered: n_classes in {8, 17, 25}, W % 32 == 0, horizontal scale >= ~3.2 (a 32-pixel
 * strip touches at most 12 source columns); anything else returns B200SEG_E_UNSUPPORTED (ask b200seg_lovasz_up_supported,
 * which returns 1 / 0). */
int b200seg_lovasz_up_supported(int32_t n_images, int32_t n_classes, int32_t h, int32_t w, int32_t H, int32_t W);
int b200seg_lovasz_up_forward(const float* lowres, int32_t h, int32_t w, const void* labels, int32_t label_dtype,
                              int32_t n_images, int32_t n_classes, int32_t H, int32_t W, int32_t per_image,
                              int64_t filter_label, int32_t keep_absent, uint32_t class_mask, int32_t need_grad,
                              void* workspace, size_t workspace_bytes, float* loss_out, int32_t ce_enabled,
                              int64_t ce_ignore_index, float* ce_out, int64_t* cm, int64_t cm_drop_label,
                              int32_t* status, void* stream);
int b200seg_lovasz_up_backward(const float* lowres, int32_t h, int32_t w, const void* labels, int32_t label_dtype,
                               int32_t n_images, int32_t n_classes, int32_t H, int32_t W, int32_t per_image,
                               int64_t filter_label, int32_t keep_absent, uint32_t class_mask, const void* workspace,
                               size_t workspace_bytes, const float* grad_lovasz, int32_t ce_enabled,
                               int64_t ce_ignore_index, const float* grad_ce, float* dlowres, void* stream);

/* Confusion matrix of (argmax_c upsample(lowres), labels) without the upsampled tensor: the metric-side twin of
 * b200seg_lovasz_up_forward for the validation loop (managers/OCRNet_Manager.py:161, managers/BaseManager.py:640-688 after
 * models/OCR.py:126-131).  Same cm / drop_label / status semantics as b200seg_confmat_accumulate; the argmax is the one torch
 * takes on F.interpolate(lowres, (H, W), mode='bilinear', align_corners=True) (NCHW-contiguous), bit for bit.
 * Any n_classes <= 32 and any scale; W % 32 == 0 is required (B200SEG_E_UNSUPPORTED otherwise; _supported returns 1 / 0). */
int b200seg_confmat_up_supported(int32_t n_images, int32_t n_classes, int32_t h, int32_t w, int32_t H, int32_t W);
int b200seg_confmat_up_accumulate(const float* lowres, int32_t h, int32_t w, const void* labels, int32_t label_dtype,
                                  int32_t n_images, int32_t n_classes, int32_t H, int32_t W, int64_t drop_label,
                                  int64_t* cm, int32_t* status, void* stream);

/* Test hook: out[planes, H, W] = bilinear upsampling (align_corners = True) of lowres[planes, h, w] exactly as the fused
 * b200seg_lovasz_up_* kernels compute it (pattern < 0), i.e. bit for bit what F.interpolate gives on the same device
 * (models/OCR.py:126, models/DeepLabv3Plus.py:65); pattern 0..35 selects one of the candidate multiply-add contractions
 * (tools/upsample_pattern.py). */
int b200seg_debug_upsample(const float* lowres, int32_t planes, int32_t h, int32_t w, int32_t H, int32_t W,
                           float* out, int32_t pattern, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* B200SEG_H_ */
