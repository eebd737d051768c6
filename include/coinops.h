/*
 * coinops.h -- C ABI of libcoinops.so: the sm_100a RoI / box-decode / IoU-match / NMS path of COIN.
 *
 * The reference (Flashkong/COIN) is pure Python; on this path it binds, through detectron2 0.5 and
 * torchvision 0.10.1, the torch dispatcher operators
 *     torchvision::roi_align, torchvision::_roi_align_backward, torchvision::nms
 * and a few dozen ATen elementwise launches per call of pairwise_iou / Matcher / Box2BoxTransform /
 * Boxes.clip, plus its own Python loops (coin/layers/nms.py, coin/engine/trainer.py:338-485).
 * Each entry point below names the reference interface it replaces (file:line under the
 * reference root). INTEGRATION.md shows the ctypes binding a maintainer of the reference adds.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer on the current CUDA device unless its name ends in _host;
 *   - the caller owns every buffer, including workspaces (query sizes with the *_workspace_bytes
 *     functions); the library never allocates, frees or retains device memory;
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as void*); no call synchronises;
 *   - variable-length results are written into caller-sized worst-case buffers plus a device-side
 *     int32 count;
 *   - return value 0 = ok, otherwise a COIN_ERR_* code; coin_last_error() gives the message of the
 *     last failure on the calling thread. No exception crosses this boundary;
 *   - boxes are fp32 [n,4] xyxy; indices are int64; Matcher labels are int8 -- as in the reference.
 */
#ifndef COINOPS_H_
#define COINOPS_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default) /* the library is built with -fvisibility=hidden */
#endif

#define COIN_OK 0
#define COIN_ERR_INVALID 1     /* bad shape / null pointer / bad enum            */
#define COIN_ERR_UNSUPPORTED 2 /* legal in the reference, not supported here     */
#define COIN_ERR_CUDA 3        /* launch or runtime failure (message has detail) */
#define COIN_ERR_CAPACITY 4    /* caller buffer / workspace too small            */

#define COIN_F32 0
#define COIN_F16 1

#define COIN_LAYOUT_NCHW 0
#define COIN_LAYOUT_NHWC 1

#define COIN_MAX_LEVELS 8

typedef void* coin_stream_t; /* cudaStream_t */

const char* coin_last_error(void);
int coin_version(void);
/* Number of kernels this library has launched in the process so far (bench.py reports the delta
 * over its timed region as `gpu_launches`; the radix sort of N > 4096 boxes counts as one). */
long long coin_launch_count(void);
/* Mode / tuning switches, by name (e.g. "COIN_ROI_EXACT" = 1: the bit-exact ROIAlign forward kernel, un-fused tap
 * arithmetic in torchvision's order; "COIN_ROI_REG" = 0: separable kernels for every shape). A name is read from the
 * environment once, on first use; coin_set_option overrides it for the process, coin_unset_option returns to the
 * built-in default. No reference counterpart: the reference selects kernels through the torch dispatcher. */
int coin_set_option(const char* name_host, int value);
int coin_unset_option(const char* name_host);
int coin_get_option(const char* name_host, int dflt);

/* ---------------------------------------------------------------------------------------------
 * ROIAlign / ROIPooler
 *   replaces: detectron2.layers.ROIAlign.forward -> torchvision::roi_align and its autograd
 *   backward torchvision::_roi_align_backward, reached from
 *   coin/modeling/roi_heads/clip_roi_heads.py:51-63,142-147,172-176 (ROIPooler(...)(features, boxes)).
 * ------------------------------------------------------------------------------------------- */

/* One pyramid level: a feature map in fp32 NHWC ([N,H,W,C]) and its spatial scale. */
typedef struct {
    const float* feat_nhwc;
    int H, W;
    float spatial_scale;
} coin_level_t;

/* [N,C,H,W] (fp32 or fp16) -> fp32 [N,H,W,C]. The gather kernels read channel-last fp32. */
int coin_nchw_to_nhwc_f32(const void* in, int in_dtype, float* out, int N, int C, int H, int W,
                          coin_stream_t stream);
/* fp32 [N,H,W,C] -> [N,C,H,W] in out_dtype (used by the backward pass). */
int coin_nhwc_f32_to_nchw(const float* in, void* out, int out_dtype, int N, int C, int H, int W,
                          coin_stream_t stream);

/* Forward over `nlevels` maps. rois: [K,5] fp32 (batch, x1, y1, x2, y2) in input-image pixels.
 * roi_level: int32 [K] level of each RoI, or NULL when nlevels == 1. out: [K,C,PH,PW] out_dtype.
 * sampling_ratio <= 0 selects the adaptive grid ceil(roi/pooled); aligned = ROIAlignV2. */
int coin_roi_align_fwd(const coin_level_t* levels_host, int nlevels, const float* rois,
                       const int32_t* roi_level, void* out, int out_dtype, int C, int K, int PH,
                       int PW, int sampling_ratio, int aligned, coin_stream_t stream);

/* Backward of the above into per-level fp32 NHWC gradient maps (levels_host[i].feat_nhwc is the
 * WRITABLE gradient buffer of level i here and must be zeroed by the caller beforehand).
 * grad_out: [K,C,PH,PW] in grad_dtype. Accumulation uses fp32 atomics. */
int coin_roi_align_bwd(const coin_level_t* grad_levels_host, int nlevels, const float* rois,
                       const int32_t* roi_level, const void* grad_out, int grad_dtype, int C, int K,
                       int PH, int PW, int sampling_ratio, int aligned, coin_stream_t stream);

/* detectron2 ROIPooler level rule: floor(canonical_level + log2(sqrt(area)/canonical_size + 1e-8))
 * clamped to [min_level, max_level], minus min_level. boxes: [n,4]; out: int32 [n]. */
int coin_roi_pooler_levels(const float* boxes, int64_t n, int min_level, int max_level,
                           int canonical_box_size, int canonical_level, int32_t* out_levels,
                           coin_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Box codec
 *   replaces: detectron2 Box2BoxTransform.apply_deltas / get_deltas, Boxes.clip, Boxes.scale
 *   reached from coin/modeling/roi_heads/fast_rcnn.py:145-147,297,619-622,691,729,
 *   coin/engine/base.py:80-126 (scale + flip), d2 RPN decode (<- proposal_generator/rpn.py:113).
 * ------------------------------------------------------------------------------------------- */

/* deltas: [R, 4*kreg]; boxes: [R,4]; out: [R, 4*kreg]. If clip != 0 the result is also clamped to
 * x in [0, clip_w], y in [0, clip_h] (Boxes.clip fused). */
int coin_apply_deltas(const float* deltas, const float* boxes, float* out, int64_t R, int kreg,
                      float wx, float wy, float ww, float wh, float scale_clamp, int clip,
                      float clip_h, float clip_w, coin_stream_t stream);

/* src, tgt: [F,4] -> out [F,4]. *invalid_flag (device int32, caller-zeroed, may be NULL) is set to 1
 * if any source width is <= 0 (the reference asserts on it). */
int coin_get_deltas(const float* src, const float* tgt, float* out, int64_t F, float wx, float wy,
                    float ww, float wh, int32_t* invalid_flag, coin_stream_t stream);

/* In-place Boxes.clip((h, w)). */
int coin_boxes_clip(float* boxes, int64_t n, float h, float w, coin_stream_t stream);

/* out = flip(scale(in)); flip: 0 none, 1 horizontal (x1' = net_w - x2, x2' = net_w - x1),
 * 2 vertical. in may equal out. */
int coin_boxes_scale_flip(const float* in, float* out, int64_t n, float sx, float sy, int flip,
                          float net_w, float net_h, coin_stream_t stream);

/* GDINO.resize_boxes (coin/modeling/meta_arch/gdino.py:144-160): cxcywh in [0,1] -> xyxy in pixels of an
 * img_h x img_w image; clip != 0 also applies the Boxes.clip that follows it (gdino.py:135-136). in may equal out. */
int coin_boxes_cxcywh_to_xyxy(const float* in, float* out, int64_t n, float img_h, float img_w, int clip,
                              coin_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * IoU and Matcher
 *   replaces: detectron2.structures.pairwise_iou and detectron2.modeling.matcher.Matcher, reached
 *   from coin/engine/trainer.py:364-366,373, coin/utils/util.py:468,
 *   coin/modeling/roi_heads/clip_roi_heads.py:301-304,311-314,353-356,
 *   coin/modeling/proposal_generator/rpn.py:159-160,169-170,212-213.
 * ------------------------------------------------------------------------------------------- */

/* out[i*M + j] = IoU(b1[i], b2[j]) (0 where the intersection is empty). */
int coin_pairwise_iou(const float* b1, int64_t N, const float* b2, int64_t M, float* out,
                      coin_stream_t stream);

/* Matcher.__call__ on a materialised [N,M] quality matrix. thresholds_host: the nthr user
 * thresholds (without the +-inf the Matcher adds); labels_host: nthr+1 labels in {-1,0,1}.
 * matches: int64 [M]; match_labels: int8 [M]; matched_vals: fp32 [M] or NULL.
 * row_max_ws: fp32 [N] workspace, required iff allow_low_quality. */
int coin_matcher(const float* quality, int64_t N, int64_t M, const float* thresholds_host, int nthr,
                 const int8_t* labels_host, int allow_low_quality, int64_t* matches,
                 int8_t* match_labels, float* matched_vals, float* row_max_ws, coin_stream_t stream);

/* pairwise_iou + Matcher fused: the [N,M] matrix is never written. gt: [N,4], boxes: [M,4].
 * row_max_ws (required iff allow_low_quality): fp32 [coin_iou_match_workspace_floats(N, M)] - the row maxima plus
 * each CTA's own row maxima, so that the low-quality pass re-evaluates a row only where its maximum was reached. */
size_t coin_iou_match_workspace_floats(int64_t N, int64_t M);
int coin_iou_match(const float* gt, int64_t N, const float* boxes, int64_t M,
                   const float* thresholds_host, int nthr, const int8_t* labels_host,
                   int allow_low_quality, int64_t* matches, int8_t* match_labels,
                   float* matched_vals, float* row_max_ws, coin_stream_t stream);

/* Relabelling epilogue of the RoI-head labelling (clip_roi_heads.py:358-362): label -1 where the
 * match is foreground and fell on a private (C) pseudo box, i.e. index in [c_begin, c_end). */
int coin_relabel_roi(const int64_t* matches, int8_t* match_labels, int64_t M, int64_t c_begin,
                     int64_t c_end, coin_stream_t stream);

/* RPN variant (rpn.py:214-228): labels/matches updated in place; distillation outputs written. */
int coin_relabel_rpn(int64_t* matches, int8_t* labels, int64_t M, int64_t len_a, int64_t len_c,
                     int64_t* distill_idx, int8_t* distill_labels, coin_stream_t stream);

/* All (i, j) with IoU(b1[i], b2[j]) >= thr in row-major order (== nonzero() of the thresholded
 * matrix, trainer.py:364-366). pairs: int64 [capacity, 2]; count: device int32 (total found, may
 * exceed capacity: only the first `capacity` pairs are written). ws: coin_iou_pairs_workspace_bytes. */
size_t coin_iou_pairs_workspace_bytes(int64_t N, int64_t M);
int coin_iou_pairs_ge(const float* b1, int64_t N, const float* b2, int64_t M, float thr,
                      int64_t* pairs, int32_t* count, int64_t capacity, void* ws, size_t ws_bytes,
                      coin_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * NMS
 *   replaces: torchvision::nms, torchvision.ops.boxes.batched_nms and the detectron2 wrapper
 *   detectron2.layers.batched_nms, reached from coin/modeling/roi_heads/fast_rcnn.py:164,
 *   coin/layers/nms.py:207, coin/modeling/meta_arch/clip_rcnn.py:161, d2 find_top_rpn_proposals.
 * ------------------------------------------------------------------------------------------- */

#define COIN_NMS_PLAIN 0   /* class-agnostic                                               */
#define COIN_NMS_TRICK 1   /* torchvision coordinate trick: boxes + idx * (max + 1)        */
#define COIN_NMS_VANILLA 2 /* per-class NMS on the original coordinates                    */
#define COIN_NMS_AUTO 3    /* the strategy the reference's CPU dependency picks for this n */

size_t coin_nms_workspace_bytes(int64_t n);

/* keep: int64 [n] (kept ORIGINAL indices in descending-score order, ties by lower index);
 * nkeep: device int32. max_keep < 0 keeps all; otherwise the sweep stops after max_keep boxes
 * (== keep[:max_keep]). idxs may be NULL for COIN_NMS_PLAIN. iou_threshold is compared as the
 * torchvision CPU kernel does (float IoU > double threshold). */
int coin_batched_nms(const float* boxes, const float* scores, const int64_t* idxs, int64_t n,
                     double iou_threshold, int strategy, int64_t max_keep, int64_t* keep,
                     int32_t* nkeep, void* ws, size_t ws_bytes, coin_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Probabilistic-fusion NMS (MyNMS)
 *   replaces: coin/layers/nms.py:84-238 (mynms.nms with CLOUD.NMS_METHOD in ps/pa/pm/as/aa/am/ms/ma),
 *   called from coin/modeling/meta_arch/gdino_processor.py:164-182.
 * ------------------------------------------------------------------------------------------- */
#define COIN_SCORE_PROBEN 0
#define COIN_SCORE_AVG 1
#define COIN_SCORE_MAX 2
#define COIN_BOX_SAVG 0
#define COIN_BOX_AVG 1
#define COIN_BOX_MAX 2

size_t coin_fusion_nms_workspace_bytes(int64_t n, int k1);

/* boxes [n,4], probs [n,k1], labels int64 [n]. per_class_offset != 0 applies the label offset
 * labels * (max + 1) before the legacy "+1" IoU (nms.py:196-203); 0 restricts clusters to equal
 * labels instead (the >= 40000 branch, nms.py:222-238). Outputs are sorted by fused score
 * (descending, ties by sweep order); capacity n rows each. status: device int32, set non-zero when
 * an assertion of the reference would fire (mixed classes in a cluster, argmax != label for probEn). */
int coin_fusion_nms(const float* boxes, const float* probs, const int64_t* labels, int64_t n, int k1,
                    float iou_threshold, int score_method, int box_method, int per_class_offset,
                    int64_t* keep, float* out_boxes, float* out_scores, float* out_probs,
                    int64_t* out_classes, int32_t* nkeep, int32_t* status, void* ws, size_t ws_bytes,
                    coin_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Detection post-processing (filter -> NMS -> top-k)
 *   replaces: fast_rcnn_inference_single_image, coin/modeling/roi_heads/fast_rcnn.py:116-175.
 * ------------------------------------------------------------------------------------------- */
size_t coin_det_postprocess_workspace_bytes(int64_t R, int k1);

/* boxes: [R, 4*kreg] decoded boxes (clipped here); scores: [R,k1] with background last.
 * Rows holding a non-finite value are dropped first. Candidates = (roi, class) with
 * score > score_thresh in row-major order; batched NMS (COIN_NMS_AUTO); first topk kept
 * (topk < 0: all). Outputs have capacity `out_capacity` rows. */
int coin_det_postprocess(const float* boxes, const float* scores, int64_t R, int k1, int kreg,
                         float img_h, float img_w, float score_thresh, double nms_thresh,
                         int64_t topk, int64_t out_capacity, float* out_boxes, float* out_scores,
                         float* out_probs, int64_t* out_classes, int64_t* out_roi_index,
                         int32_t* out_count, void* ws, size_t ws_bytes, coin_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Knowledge separation (consistent A / inconsistent B / private C)
 *   replaces: CoinTrainer.match_dual_teacher, coin/engine/trainer.py:338-461, with
 *   delete_duplicate_boxes / filter_result / online_boxes_merging (coin/utils/util.py:434-507)
 *   and merge_boxes (trainer.py:480-485, coin/layers/nms.py:24-31).
 * ------------------------------------------------------------------------------------------- */
#define COIN_TAG_RCNN 0
#define COIN_TAG_RPN 1
#define COIN_ABC_MAX 1024 /* per-side detection limit of the single-CTA kernel */

/* Row r of the A/B outputs is the pair (on_index[r], off_index[r]) of cloud ("online") and
 * CLIP-detector ("offline") detections with merged box out_boxes[r]; -1 marks "no such side"
 * (the empty-side branches, trainer.py:343-361, where both sides are the same detection).
 * C rows reference exactly one side. All index outputs have capacity cap_pairs = nc*nd + nc + nd
 * (A, B) and nc + nd (C). counts: device int32 [8] = {nA, nB, nC, status, nC_off, 0, 0, 0}: C rows
 * [0, nC_off) are CLIP-detector rows (c_off valid), rows [nC_off, nC) cloud rows (c_on valid);
 * status bits: 4 = pair capacity exceeded; 8 = a duplicate group holds several boxes of the matched class (the
 * reference fails at trainer.py:402 on that input); 16 = a cloud self-cluster with a single class (the reference
 * asserts, util.py:488); 32 = a self-cluster too large for the on-chip replay of CPython's set order (then, and only
 * then, clusters are taken lowest-member-first). Rows come in the REFERENCE's order: the Python-set iterations of
 * trainer.py:369,391 and util.py:481 are replayed exactly (csrc/pyset.cuh); random.randint picks take the first
 * element, i.e. the reference with randint pinned to its lower bound ("determinism policy"). */
size_t coin_match_abc_workspace_bytes(int64_t nc, int64_t nd);
int coin_match_abc(const float* on_boxes, const int64_t* on_classes, const float* on_scores, int64_t nc,
                   const float* off_boxes, const int64_t* off_classes, const float* off_scores,
                   int64_t nd, int tag, float iou_thr, float weight_for_box_a,
                   int64_t cap_pairs, int32_t* a_on, int32_t* a_off, float* a_boxes, int32_t* b_on,
                   int32_t* b_off, float* b_boxes, int32_t* c_on, int32_t* c_off, int32_t* counts,
                   void* ws, size_t ws_bytes, coin_stream_t stream);

/* Launch order of the ROIAlign grids (scheduling only, results are unchanged): perm = 0..K-1 with the `small_pct` %
 * smallest RoIs (by box area) moved to the end and the `big_pct` % largest to the front (every part in input order), so
 * that the last wave of CTAs holds cheap RoIs and a map-sized RoI starts with the grid instead of wherever it happens to
 * stand (d2's ROIPooler hands the kernels the proposals in sampling order, clip_roi_heads.py:172-176).
 * rois: [K_cap,5]; k_dev: optional device live count; perm: int32 [K_cap] (entries beyond the live count = identity).
 * K_cap <= 8192 (larger grids have no tail worth ordering). One single-CTA launch. */
int coin_roi_launch_order(const float* rois, int K_cap, const int32_t* k_dev, int small_pct, int big_pct, int32_t* perm,
                          coin_stream_t stream);
/* coin_roi_launch_order and the size split below in ONE launch: perm = [largest big_pct % | rest | smallest small_pct % | diverted],
 * where "diverted" are the RoIs with area > area_thr or a side > side_thr (pixels) - at most divert_cap of them, otherwise
 * none is diverted; perm_divert = the diverted RoIs alone; counts = int32 [2] = {K_live - diverted, diverted}. A caller
 * launches its main kernel over perm with the live count counts[0] and the separable kernel over perm_divert with counts[1]
 * (ops.roi_align_forward_planned); perm as a whole stays a permutation of the live RoIs (for the backward). */
int coin_roi_launch_plan(const float* rois, int K_cap, const int32_t* k_dev, int small_pct, int big_pct, float area_thr,
                         float side_thr, int divert_cap, int32_t* perm, int32_t* perm_divert, int32_t* counts,
                         coin_stream_t stream);
/* Two RoI index lists by box size (pixels): perm_big = RoIs with area > area_thr or a side > side_thr, perm_small = the
 * rest, both in input order - or in the order of `order` (a coin_roi_launch_order result, may be NULL);
 * perm_big holds at most big_cap RoIs (the capacity its launch is sized for; further big RoIs stay in perm_small);
 * counts = int32 [2] device lengths (small, big). With coin_roi_align_fwd_ord (perm = a list, k_dev = its length, same `out`) a caller
 * pools the two subsets with different kernels: the step sends the rare map-sized private box to the separable kernel. */
int coin_roi_split_by_area(const float* rois, int K_cap, const int32_t* k_dev, float area_thr, float side_thr, int big_cap,
                           const int32_t* order, int32_t* perm_small, int32_t* perm_big, int32_t* counts,
                           coin_stream_t stream);
/* coin_roi_align_fwd_dev / coin_roi_align_bwd with a launch order / RoI subset (perm may be NULL): CTA group i works on RoI
 * perm[i], i < *k_dev. */
int coin_roi_align_fwd_ord(const coin_level_t* levels_host, int nlevels, const float* rois,
                           const int32_t* roi_level, void* out, int out_dtype, int C, int K_cap, int PH,
                           int PW, int sampling_ratio, int aligned, const int32_t* k_dev, const int32_t* perm,
                           coin_stream_t stream);
int coin_roi_align_bwd_ord(const coin_level_t* grad_levels_host, int nlevels, const float* rois,
                           const int32_t* roi_level, const void* grad_out, int grad_dtype, int C, int K,
                           int PH, int PW, int sampling_ratio, int aligned, const int32_t* k_dev, const int32_t* perm,
                           coin_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Sync-free ("device-count") variants
 *   In the reference every stage boundary of CoinTrainer.run_step (coin/engine/trainer.py:160-218)
 *   is a host round trip: nonzero()/tolist()/len() on device tensors decide the shapes of the next
 *   stage (trainer.py:364-391,469; clip_roi_heads.py:345-362; rpn.py:209-228; fast_rcnn.py:151-166).
 *   These entry points take the live lengths as DEVICE int32 values next to host-known capacities:
 *   every launch is sized for the capacity and threads beyond the live length exit, so the whole
 *   step is a fixed launch sequence (capturable in a CUDA graph) with one count read-back at the end.
 * ------------------------------------------------------------------------------------------- */

/* coin_roi_align_fwd with a device-side live RoI count *k_dev <= K_cap (NULL: K_cap). Rows of `out`
 * beyond the live count are left untouched. */
int coin_roi_align_fwd_dev(const coin_level_t* levels_host, int nlevels, const float* rois,
                           const int32_t* roi_level, void* out, int out_dtype, int C, int K_cap, int PH,
                           int PW, int sampling_ratio, int aligned, const int32_t* k_dev,
                           coin_stream_t stream);

/* coin_iou_match with device-side live counts *n_dev <= N_cap (gt rows) and *m_dev <= M_cap (columns);
 * either may be NULL. A live N of 0 applies Matcher's empty-matrix rule on the device. Entries of the
 * outputs beyond the live M are left untouched. row_max_ws: fp32 [coin_iou_match_workspace_floats(N_cap, M_cap)]
 * iff allow_low_quality. */
int coin_iou_match_dev(const float* gt, int64_t N_cap, const int32_t* n_dev, const float* boxes,
                       int64_t M_cap, const int32_t* m_dev, const float* thresholds_host, int nthr,
                       const int8_t* labels_host, int allow_low_quality, int64_t* matches,
                       int8_t* match_labels, float* matched_vals, float* row_max_ws, coin_stream_t stream);

/* coin_relabel_roi / coin_relabel_rpn with the pseudo-GT set lengths read from device memory
 * (clip_roi_heads.py:358-362: C rows are [len_a+len_b, len_a+len_b+len_c); rpn.py:214-228). */
int coin_relabel_roi_dev(const int64_t* matches, int8_t* match_labels, int64_t M_cap, const int32_t* m_dev,
                         const int32_t* len_a, const int32_t* len_b, const int32_t* len_c,
                         coin_stream_t stream);
int coin_relabel_rpn_dev(int64_t* matches, int8_t* labels, int64_t M, const int32_t* len_a,
                         const int32_t* len_c, int64_t* distill_idx, int8_t* distill_labels,
                         coin_stream_t stream);

/* coin_match_abc with the CLIP-detector detection count read from device memory (*nd_dev <= nd_cap):
 * chains behind coin_det_postprocess without reading its count back. */
int coin_match_abc_dev(const float* on_boxes, const int64_t* on_classes, const float* on_scores, int64_t nc,
                       const float* off_boxes, const int64_t* off_classes, const float* off_scores,
                       int64_t nd_cap, const int32_t* nd_dev, int tag, float iou_thr, float weight_for_box_a,
                       int64_t cap_pairs, int32_t* a_on, int32_t* a_off, float* a_boxes, int32_t* b_on,
                       int32_t* b_off, float* b_boxes, int32_t* c_on, int32_t* c_off, int32_t* counts,
                       void* ws, size_t ws_bytes, coin_stream_t stream);

/* Both tags of one image in ONE launch: 'RCNN' and 'RPN' share everything up to the A/B split (trainer.py:401), so
 * match_boxes (trainer.py:463-478) need not run the matching twice. rcnn_* / rpn_* as a_* / b_* above (tag 'RPN' has
 * no B set), c_on / c_off are common to both tags, counts_rcnn / counts_rpn are two device int32 [8] vectors. */
int coin_match_abc_both_dev(const float* on_boxes, const int64_t* on_classes, const float* on_scores, int64_t nc,
                            const float* off_boxes, const int64_t* off_classes, const float* off_scores,
                            int64_t nd_cap, const int32_t* nd_dev, float iou_thr, float weight_for_box_a,
                            int64_t cap_pairs, int32_t* rcnn_a_on, int32_t* rcnn_a_off, float* rcnn_a_boxes,
                            int32_t* rcnn_b_on, int32_t* rcnn_b_off, float* rcnn_b_boxes, int32_t* rpn_a_on,
                            int32_t* rpn_a_off, float* rpn_a_boxes, int32_t* c_on, int32_t* c_off,
                            int32_t* counts_rcnn, int32_t* counts_rpn, void* ws, size_t ws_bytes,
                            coin_stream_t stream);

/* One segment of a row concatenation: `count` rows at `ptr` (or *count_dev <= count rows when
 * count_dev is non-NULL); `prefix` is the value of the extra leading column, if any. */
typedef struct {
    const float* ptr;
    const int32_t* count_dev;
    int64_t count;
    float prefix;
} coin_seg_t;

/* out = cat(segments) row-wise; width_out == width_in, or width_in + 1 to prepend the per-segment
 * `prefix` column (the batch index of convert_boxes_to_pooler_format). *out_count (device int32, may be
 * NULL) receives the total number of rows written (clamped to out_cap). Replaces Boxes.cat /
 * add_ground_truth_to_proposals (clip_roi_heads.py:345-353, rpn.py:209-212) for device-length sets. */
int coin_concat_rows(const coin_seg_t* segs_host, int nseg, int width_in, int width_out, float* out,
                     int64_t out_cap, int32_t* out_count, coin_stream_t stream);

/* One detection set (device pointers): boxes [n,4], classes int64 [n], scores [n], probs [n,k1]. */
typedef struct {
    const float* boxes;
    const int64_t* classes;
    const float* scores;
    const float* probs;
} coin_dets_t;

/* One pseudo-label set of match_dual_teacher, capacity rows each (trainer.py:393-455). A and B: `classes`
 * = the CLIP-detector ("offline") class, classes_online (B only) = the cloud class, scores/probs of both
 * sides; their boxes are the a_boxes / b_boxes of coin_match_abc. C: boxes, classes, scores_online = the
 * single score, probs_online = the single prob row (the *_offline members are ignored). */
typedef struct {
    float* boxes;
    int64_t* classes;
    int64_t* classes_online;
    float* scores_online;
    float* scores_offline;
    float* probs_online;
    float* probs_offline;
} coin_pseudo_t;

/* Gathers the A / B / C fields from the index lists and counts of coin_match_abc[_dev]. */
int coin_abc_pack(const coin_dets_t* online_host, int64_t nc, const coin_dets_t* offline_host,
                  int64_t nd_cap, const int32_t* nd_dev, int k1, int tag, const int32_t* a_on,
                  const int32_t* a_off, const int32_t* b_on, const int32_t* b_off, const int32_t* c_on,
                  const int32_t* c_off, const int32_t* counts, const coin_pseudo_t* a_out_host,
                  const coin_pseudo_t* b_out_host, const coin_pseudo_t* c_out_host, int64_t cap_pairs,
                  coin_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * RPN proposal selection for one image and one feature level.
 *   replaces: detectron2 0.5 RPN.predict_proposals = _decode_proposals (Box2BoxTransform.apply_deltas on ALL
 *   anchors) + find_top_rpn_proposals (sort the objectness logits, keep pre_nms_topk, drop non-finite rows,
 *   Boxes.clip, Boxes.nonempty(min_box_size), batched_nms, keep[:post_nms_topk]) as reached from
 *   coin/modeling/proposal_generator/rpn.py:64,113 (DualTeacherRPN.forward).
 * Only the selected anchors are decoded; no host round trip (the reference syncs in nonempty() and nms()).
 *   anchors, deltas: float32 [A,4] (16-byte aligned); logits: float32 [A]; weights wx..wh = (1,1,1,1) for the RPN.
 *   out_boxes: float32 [min(A, pre, post), 4]; out_logits: float32 [same]; out_count: device int32 (live rows);
 *   status: device int32, bit 0 set when a selected row was non-finite (detectron2 raises FloatingPointError in
 *   training; the Python mirror does the same). Ties of equal logits: lower anchor index first.
 * ---------------------------------------------------------------------------------------------- */
size_t coin_rpn_proposals_workspace_bytes(int64_t A, int64_t pre_nms_topk);
int coin_rpn_proposals(const float* anchors, const float* deltas, const float* logits, int64_t A,
                       int64_t pre_nms_topk, int64_t post_nms_topk, double nms_thresh, float min_box_size,
                       float img_h, float img_w, float wx, float wy, float ww, float wh, float scale_clamp,
                       float* out_boxes, float* out_logits, int32_t* out_count, int32_t* status, void* ws,
                       size_t ws_bytes, coin_stream_t stream);

/* coin_rpn_proposals with the anchors of detectron2's DefaultAnchorGenerator (<- rpn.py:64) generated on the fly:
 * anchor(i) = cell_anchors[i % ncell] + (x, y, x, y) * stride (+ offset * stride), i = (y * Wf + x) * ncell + c, the order
 * and the single fp32 addition of grid_anchors - bit-identical to passing the materialised [Hf*Wf*ncell, 4] array, which
 * is never read. cell_anchors_host: HOST pointer to [ncell, 4] floats (ncell <= 32). Workspace: as coin_rpn_proposals. */
int coin_rpn_proposals_grid(const float* cell_anchors_host, int ncell, int Hf, int Wf, float stride, float offset,
                            const float* deltas, const float* logits, int64_t pre_nms_topk, int64_t post_nms_topk,
                            double nms_thresh, float min_box_size, float img_h, float img_w, float wx, float wy,
                            float ww, float wh, float scale_clamp, float* out_boxes, float* out_logits,
                            int32_t* out_count, int32_t* status, void* ws, size_t ws_bytes, coin_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Sampling and loss-side reductions
 * ------------------------------------------------------------------------------------------- */

/* detectron2 ROIHeads._sample_proposals, first half (<- coin/modeling/roi_heads/clip_roi_heads.py:317,363):
 * out[i] = num_classes where matched_labels[i] == 0, -1 where it is -1, else gt_classes[matched_idxs[i]]; num_classes for
 * every row when there is no ground truth. Rows beyond the live count *m_dev get -1 (the sampler ignores them). */
int coin_proposal_classes(const int64_t* matched_idxs, const int8_t* matched_labels, const int64_t* gt_classes,
                          int64_t n_gt_cap, const int32_t* n_gt_dev, int64_t m_cap, const int32_t* m_dev,
                          int64_t num_classes, int64_t* out_classes, coin_stream_t stream);

/* detectron2 modeling/sampling.py::subsample_labels (<- clip_roi_heads.py:363 via _sample_proposals, rpn.py:231 via
 * _subsample_labels). labels: int64 (label_is_int8 = 0) or int8 [m_cap]; positives = labels != -1 && != bg_label,
 * negatives = labels == bg_label. num_pos_target = int(num_samples * positive_fraction), computed by the caller.
 * The random draw: perm_pos / perm_neg non-NULL -> pos_idx = positive[perm_pos[:num_pos]], neg_idx = negative[perm_neg[:num_neg]]
 * (replay of the reference's two torch.randperm draws; they must be permutations of [0, P) and [0, N), which a first call
 * with count_only = 1 reports in counts[2], counts[3]); NULL -> device generator Philox4x32-10 keyed by `seed`, counter
 * (element, set, offset): the num smallest (key, element) pairs of each set, in key order.
 * counts: int32 [5] = {num_pos, num_neg, P, N, status (1: a permutation entry out of range)}. num_samples <= 4096. */
size_t coin_subsample_labels_workspace_bytes(int64_t m_cap);
int coin_subsample_labels(const void* labels, int label_is_int8, int64_t m_cap, const int32_t* m_dev, int num_samples,
                          int num_pos_target, int64_t bg_label, const int64_t* perm_pos, const int64_t* perm_neg,
                          uint64_t seed, uint64_t offset, int count_only, int64_t* pos_idx, int64_t* neg_idx,
                          int32_t* counts, void* ws, size_t ws_bytes, coin_stream_t stream);

/* coin/modeling/proposal_generator/rpn.py:95-98: out[i] = sum(gt_probs[matched[i], :-1]) (0 when there are no C boxes). */
int coin_rpn_teacher_probs(const float* gt_probs, int64_t nc_cap, const int32_t* nc_dev, int k1, const int64_t* matched,
                           int64_t n, float* out, coin_stream_t stream);

/* coin/modeling/roi_heads/fast_rcnn.py:541-545: loss = mean over [n, k1] of q * (log q - log(softmax(scores) + 1e-7))
 * (torch.nn.KLDivLoss, reduction 'mean'); n may be a device count. ws: coin_kl_workspace_bytes() bytes.
 * _bwd: grad_scores = d loss / d scores * *grad_loss (rows beyond the live count are zero-filled). */
size_t coin_kl_workspace_bytes(void);
int coin_kl_distill_roi_fwd(const float* scores, const float* gt_probs, int64_t n_cap, const int32_t* n_dev, int k1,
                            float* loss, void* ws, coin_stream_t stream);
int coin_kl_distill_roi_bwd(const float* scores, const float* gt_probs, int64_t n_cap, const int32_t* n_dev, int k1,
                            const float* grad_loss, float* grad_scores, coin_stream_t stream);
/* coin/modeling/proposal_generator/rpn.py:326-340: the two-column KL between (sigmoid(logit), 1 - sigmoid(logit)) and
 * (teacher, 1 - teacher) over the anchors whose distillation label is > 0, reduction 'mean' (sum / (2 * n_valid));
 * n_valid (device int32) is written by _fwd and read by _bwd; loss = 0 when no anchor is valid. */
int coin_kl_distill_rpn_fwd(const float* logits, const int8_t* distillation_labels, const float* teacher_probs, int64_t n,
                            float* loss, int32_t* n_valid, void* ws, coin_stream_t stream);
int coin_kl_distill_rpn_bwd(const float* logits, const int8_t* distillation_labels, const float* teacher_probs, int64_t n,
                            const int32_t* n_valid, const float* grad_loss, float* grad_logits, coin_stream_t stream);

/* Packs the live prefixes of many result buffers into one contiguous staging buffer, so that a step's ~100 variable-length
 * results (detections, A/B/C fields, labels, keep lists: trainer.py:393-459 hands them on one by one) leave the device in ONE
 * copy. items_dev: DEVICE array of n_items records {const void* ptr; int64 row_bytes; int64 cap_rows; int32 count_index;
 * int32 pad} (32 bytes each; count_index = index into counts_dev of the live row count, -1 = all rows). Segment i starts at
 * offsets_dev[i] (16-byte aligned); offsets_dev[n_items] = total bytes. packed_cap >= the sum of the padded capacities. */
int coin_pack_rows(const void* items_dev, int n_items, const int32_t* counts_dev, void* packed, int64_t packed_cap,
                   int64_t* offsets_dev, coin_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Evaluation
 * ------------------------------------------------------------------------------------------- */

/* order = stable descending argsort of fp32 scores (ties: lower index first; NaN first): what np.argsort(-confidence)
 * (cloud_pascal_voc_evaluation.py:252) and torch's descending sorts on this path compute. n <= 262144. */
size_t coin_argsort_desc_workspace_bytes(int64_t n);
int coin_argsort_desc(const float* scores, int64_t n, int64_t* order, void* ws, size_t ws_bytes, coin_stream_t stream);

/* The TP / FP marking loop of voc_eval for one class (coin/evaluation/cloud_pascal_voc_evaluation.py:259-308).
 * det_image: int32 [nd] image index of every detection; det_boxes: float64 [nd,4]; order: int64 [nd] detections by
 * descending confidence; ground truth of the class as CSR: gt_boxes float64 [ng,4], gt_offsets int32 [n_images+1],
 * gt_difficult uint8 [ng]. tp / fp: float64 [nd] in `order` (0 or 1), ready for the cumulative sums. */
size_t coin_voc_match_workspace_bytes(int64_t nd, int64_t ng);
int coin_voc_match(const int32_t* det_image, const double* det_boxes, const int64_t* order, int64_t nd,
                   const double* gt_boxes, const int32_t* gt_offsets, const uint8_t* gt_difficult, int64_t ng,
                   double ovthresh, double* tp, double* fp, void* ws, size_t ws_bytes, coin_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Gradient all-reduce over NVLink peer memory (BASELINE.json configs[3]; coin/engine/trainer.py:66-72 wraps the model in DDP)
 * ------------------------------------------------------------------------------------------- */

/* In-place sum over the n ranks of [offset, offset + nelem) fp32 elements of peer-mapped buffers (one process per GPU, n <= 8;
 * the buffers and the 24-word flag arrays of all ranks are mapped into every process through CUDA IPC by the caller:
 * coin_b200/p2p.py). data_ptrs / flag_ptrs: HOST arrays of n device pointers as seen from this process. offset, nelem: multiples
 * of 4 * n. epoch: grows with every call on the group, the same on every rank. err: device int32, non-zero if a peer did not
 * arrive (the kernels give up instead of hanging). max_ctas: grid size of the two data kernels (0: 4 per SM). Five launches of
 * 224-thread / one-warp CTAs: sized to run beside the ROIAlign grids, which NCCL's kernels cannot (tools/ar_overlap.py). */
int coin_p2p_all_reduce(void* const* data_ptrs, void* const* flag_ptrs, int rank, int n, int64_t offset, int64_t nelem,
                        uint32_t epoch, int32_t* err, int max_ctas, coin_stream_t stream);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif /* COINOPS_H_ */
