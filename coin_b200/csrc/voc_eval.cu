// voc_eval.cu -- the matching loop of the reference's Pascal-VOC evaluator for one class, without the serial walk
// (SURVEY.md 8(f) rank 4).
//
// Replaces coin/evaluation/cloud_pascal_voc_evaluation.py:259-308 (voc_eval: "go down dets and mark TPs and FPs"), run by
// Cloud_PascalVOCDetectionEvaluator.evaluate every EVAL_PERIOD iterations over every test image and class. The reference
// walks the detections in descending confidence and, per detection, evaluates numpy IoUs against the image's ground-truth
// boxes and updates a per-box "already detected" flag. The only sequential dependence is that flag; it resolves without
// a walk: the best ground-truth box of a detection (first maximum of the legacy "+1" IoU, float64 like numpy) does not
// depend on other detections, and a box is claimed by the matching detection of LOWEST rank - an atomicMin.
//   voc_match_kernel     one thread per detection (in confidence order): argmax IoU over its image's boxes, claim
//   voc_mark_kernel      tp = matched, not difficult, owns the claim; fp = unmatched, or matched a claimed box;
//                        a match on a 'difficult' box is neither (cloud_pascal_voc_evaluation.py:299-306)
// The cumulative sums and the AP integral stay with the caller (coin_b200/evaluation.py).
#include "common.cuh"

namespace coin {

__global__ void voc_match_kernel(const int32_t* __restrict__ det_image, const double* __restrict__ det_boxes,
                                 const int64_t* __restrict__ order, int64_t nd, const double* __restrict__ gt_boxes,
                                 const int32_t* __restrict__ gt_offsets, const uint8_t* __restrict__ gt_difficult,
                                 double thr, int32_t* __restrict__ claim, int32_t* __restrict__ jbest,
                                 uint8_t* __restrict__ over) {
    const int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (r >= nd) return;
    const int64_t d = order[r];
    const int img = det_image[d];
    const double bx0 = det_boxes[4 * d], by0 = det_boxes[4 * d + 1], bx1 = det_boxes[4 * d + 2], by1 = det_boxes[4 * d + 3];
    const double area = (bx1 - bx0 + 1.0) * (by1 - by0 + 1.0);
    double ovmax = -INFINITY;
    int jmax = -1;
    for (int j = gt_offsets[img]; j < gt_offsets[img + 1]; ++j) {
        const double gx0 = gt_boxes[4 * j], gy0 = gt_boxes[4 * j + 1], gx1 = gt_boxes[4 * j + 2], gy1 = gt_boxes[4 * j + 3];
        const double iw = fmax(fmin(gx1, bx1) - fmax(gx0, bx0) + 1.0, 0.0);
        const double ih = fmax(fmin(gy1, by1) - fmax(gy0, by0) + 1.0, 0.0);
        const double inters = iw * ih;
        const double uni = area + (gx1 - gx0 + 1.0) * (gy1 - gy0 + 1.0) - inters;
        const double ov = inters / uni;
        if (ov > ovmax || (jmax < 0 && !(ov <= ovmax))) { ovmax = ov; jmax = j; }     // np.argmax: first maximum (NaN wins)
    }
    const bool hit = ovmax > thr;
    jbest[r] = jmax;
    over[r] = hit;
    if (hit && !gt_difficult[jmax]) atomicMin(claim + jmax, (int32_t)r);
}

__global__ void voc_mark_kernel(int64_t nd, const int32_t* __restrict__ jbest, const uint8_t* __restrict__ over,
                                const uint8_t* __restrict__ gt_difficult, const int32_t* __restrict__ claim,
                                double* __restrict__ tp, double* __restrict__ fp) {
    const int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (r >= nd) return;
    double t = 0.0, f = 0.0;
    if (over[r]) {
        const int j = jbest[r];
        if (!gt_difficult[j]) { if (claim[j] == (int32_t)r) t = 1.0; else f = 1.0; }
    } else {
        f = 1.0;
    }
    tp[r] = t;
    fp[r] = f;
}

}  // namespace coin
using namespace coin;

extern "C" size_t coin_voc_match_workspace_bytes(int64_t nd, int64_t ng) {
    Carver c(nullptr);
    c.take<int32_t>((size_t)std::max<int64_t>(ng, 1));
    c.take<int32_t>((size_t)std::max<int64_t>(nd, 1));
    c.take<uint8_t>((size_t)std::max<int64_t>(nd, 1));
    return c.used() + 256;
}

extern "C" int coin_voc_match(const int32_t* det_image, const double* det_boxes, const int64_t* order, int64_t nd,
                              const double* gt_boxes, const int32_t* gt_offsets, const uint8_t* gt_difficult, int64_t ng,
                              double ovthresh, double* tp, double* fp, void* ws, size_t ws_bytes, coin_stream_t stream) {
    COIN_REQUIRE(nd >= 0 && ng >= 0 && nd < (1ll << 31) && ng < (1ll << 31), "voc_match: bad sizes");
    if (nd == 0) return COIN_OK;
    COIN_REQUIRE(det_image && det_boxes && order && gt_offsets && tp && fp && ws, "voc_match: null pointer");
    COIN_REQUIRE(ng == 0 || (gt_boxes && gt_difficult), "voc_match: null ground truth");
    if (ws_bytes < coin_voc_match_workspace_bytes(nd, ng)) return fail(COIN_ERR_CAPACITY, "voc_match: workspace too small");
    Carver c(ws);
    int32_t* claim = c.take<int32_t>((size_t)std::max<int64_t>(ng, 1));
    int32_t* jbest = c.take<int32_t>((size_t)nd);
    uint8_t* over = c.take<uint8_t>((size_t)nd);
    cudaStream_t s = as_stream(stream);
    fill_bytes(claim, 0x7f, (size_t)std::max<int64_t>(ng, 1) * sizeof(int32_t), s);     // 0x7f7f7f7f: larger than any rank
    const unsigned blocks = (unsigned)ceil_div(nd, 256);
    voc_match_kernel<<<blocks, 256, 0, s>>>(det_image, det_boxes, order, nd, gt_boxes, gt_offsets, gt_difficult, ovthresh, claim,
                                            jbest, over);
    if (int rc = check_launch("voc_match_kernel")) return rc;
    voc_mark_kernel<<<blocks, 256, 0, s>>>(nd, jbest, over, gt_difficult, claim, tp, fp);
    return check_launch("voc_mark_kernel");
}
