// det_postprocess.cu -- score filter -> per-class NMS -> top-k for one image, with no host sync.
//
// Replaces fast_rcnn_inference_single_image, coin/modeling/roi_heads/fast_rcnn.py:116-175: drop
// non-finite rows (:136-139), strip the background column, Boxes.clip (:145-147), candidates
// (roi, class) with score > score_thresh in nonzero() (row-major) order (:151-161), batched_nms
// (:164), keep[:topk] (:165-166) and the gathers (:167-174). The reference issues a nonzero() sync,
// an NMS sync and a dozen indexing launches; here the candidate count stays on the device and feeds
// the NMS pipeline of nms.cu directly.
#include "common.cuh"

namespace coin {

size_t nms_pipeline_workspace_bytes(int64_t n_cap);
int nms_sorted_pipeline(const float* boxes, const float* scores, const int64_t* idxs, int64_t n_cap,
                        const int32_t* n_dev, double thr, int strategy, int64_t max_keep, int64_t* keep,
                        int32_t* nkeep, void* ws, size_t ws_bytes, cudaStream_t s);

__global__ void dp_rows_kernel(const float* __restrict__ boxes, const float* __restrict__ scores, int R, int k1,
                               int kreg, float score_thresh, int32_t* __restrict__ row_valid,
                               int32_t* __restrict__ row_cnt) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= R) return;
    bool ok = true;
    for (int i = 0; i < 4 * kreg; ++i) ok &= isfinite(boxes[(size_t)r * 4 * kreg + i]);
    int cnt = 0;
    for (int k = 0; k < k1; ++k) {
        const float s = scores[(size_t)r * k1 + k];
        ok &= isfinite(s);
        cnt += (k < k1 - 1 && s > score_thresh);
    }
    row_valid[r] = ok;
    row_cnt[r] = ok ? cnt : 0;
}

// single block: exclusive scans of row_valid and row_cnt
__global__ void dp_scan_kernel(const int32_t* __restrict__ row_valid, const int32_t* __restrict__ row_cnt, int R,
                               int32_t* __restrict__ row_rank, int32_t* __restrict__ row_off,
                               int32_t* __restrict__ n_cand) {
    __shared__ int32_t wsum[2][32];
    __shared__ int32_t carry[2];
    if (threadIdx.x < 2) carry[threadIdx.x] = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int base = 0; base < R; base += blockDim.x) {
        const int i = base + threadIdx.x;
        const int32_t v0 = i < R ? row_valid[i] : 0, v1 = i < R ? row_cnt[i] : 0;
        int32_t x0 = v0, x1 = v1;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int32_t y0 = __shfl_up_sync(0xffffffffu, x0, o), y1 = __shfl_up_sync(0xffffffffu, x1, o);
            if (lane >= o) { x0 += y0; x1 += y1; }
        }
        if (lane == 31) { wsum[0][warp] = x0; wsum[1][warp] = x1; }
        __syncthreads();
        if (warp < 2) {
            int32_t sv = lane < nw ? wsum[warp][lane] : 0;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int32_t y = __shfl_up_sync(0xffffffffu, sv, o);
                if (lane >= o) sv += y;
            }
            wsum[warp][lane] = sv;
        }
        __syncthreads();
        const int32_t b0 = carry[0] + (warp ? wsum[0][warp - 1] : 0) + x0 - v0;
        const int32_t b1 = carry[1] + (warp ? wsum[1][warp - 1] : 0) + x1 - v1;
        if (i < R) { row_rank[i] = b0; row_off[i] = b1; }
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1) { carry[0] = b0 + v0; carry[1] = b1 + v1; }
        __syncthreads();
    }
    if (threadIdx.x == 0) *n_cand = carry[1];
}

__global__ void dp_compact_kernel(const float* __restrict__ boxes, const float* __restrict__ scores, int R, int k1,
                                  int kreg, float img_h, float img_w, float score_thresh,
                                  const int32_t* __restrict__ row_valid, const int32_t* __restrict__ row_rank,
                                  const int32_t* __restrict__ row_off, float4* __restrict__ cand_box,
                                  float* __restrict__ cand_score, int64_t* __restrict__ cand_cls,
                                  int32_t* __restrict__ cand_row, int32_t* __restrict__ cand_src) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= R || !row_valid[r]) return;
    int pos = row_off[r];
    for (int k = 0; k < k1 - 1; ++k) {
        const float s = scores[(size_t)r * k1 + k];
        if (!(s > score_thresh)) continue;
        const float* b = boxes + ((size_t)r * kreg + (kreg == 1 ? 0 : k)) * 4;
        cand_box[pos] = make_float4(fminf(fmaxf(b[0], 0.0f), img_w), fminf(fmaxf(b[1], 0.0f), img_h),
                                    fminf(fmaxf(b[2], 0.0f), img_w), fminf(fmaxf(b[3], 0.0f), img_h));
        cand_score[pos] = s;
        cand_cls[pos] = k;
        cand_row[pos] = row_rank[r];
        cand_src[pos] = r;
        ++pos;
    }
}

__global__ void dp_gather_kernel(const int64_t* __restrict__ keep, const int32_t* __restrict__ nkeep, int64_t cap,
                                 const float4* __restrict__ cand_box, const float* __restrict__ cand_score,
                                 const int64_t* __restrict__ cand_cls, const int32_t* __restrict__ cand_row,
                                 const int32_t* __restrict__ cand_src, const float* __restrict__ scores, int k1,
                                 float4* __restrict__ out_boxes, float* __restrict__ out_scores,
                                 float* __restrict__ out_probs, int64_t* __restrict__ out_classes,
                                 int64_t* __restrict__ out_roi, int32_t* __restrict__ out_count) {
    const int64_t total = min((int64_t)*nkeep, cap);
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) *out_count = (int32_t)total;
    if (i >= total) return;
    const int64_t c = keep[i];
    out_boxes[i] = cand_box[c];
    out_scores[i] = cand_score[c];
    out_classes[i] = cand_cls[c];
    out_roi[i] = cand_row[c];
    const float* srow = scores + (size_t)cand_src[c] * k1;
    for (int k = 0; k < k1; ++k) out_probs[(size_t)i * k1 + k] = srow[k];
}

struct DpWs {
    int32_t *row_valid, *row_cnt, *row_rank, *row_off, *n_cand, *nkeep, *cand_row, *cand_src;
    float4* cand_box;
    float* cand_score;
    int64_t *cand_cls, *keep;
    void* nms_ws;
    size_t nms_bytes, total;
};

static DpWs carve_dp(void* ws, int64_t R, int k1) {
    DpWs w;
    Carver c(ws);
    const int64_t cap = R * (k1 - 1);
    w.row_valid = c.take<int32_t>((size_t)R); w.row_cnt = c.take<int32_t>((size_t)R);
    w.row_rank = c.take<int32_t>((size_t)R);  w.row_off = c.take<int32_t>((size_t)R);
    w.n_cand = c.take<int32_t>(64);           w.nkeep = c.take<int32_t>(64);
    w.cand_box = c.take<float4>((size_t)cap); w.cand_score = c.take<float>((size_t)cap);
    w.cand_cls = c.take<int64_t>((size_t)cap); w.cand_row = c.take<int32_t>((size_t)cap);
    w.cand_src = c.take<int32_t>((size_t)cap); w.keep = c.take<int64_t>((size_t)cap);
    w.nms_bytes = nms_pipeline_workspace_bytes(cap);
    w.nms_ws = c.take<char>(w.nms_bytes);
    w.total = c.used();
    return w;
}

}  // namespace coin
using namespace coin;

extern "C" size_t coin_det_postprocess_workspace_bytes(int64_t R, int k1) {
    if (R <= 0 || k1 < 2) return 256;
    return carve_dp(nullptr, R, k1).total + 256;
}

extern "C" int coin_det_postprocess(const float* boxes, const float* scores, int64_t R, int k1, int kreg,
                                    float img_h, float img_w, float score_thresh, double nms_thresh,
                                    int64_t topk, int64_t out_capacity, float* out_boxes, float* out_scores,
                                    float* out_probs, int64_t* out_classes, int64_t* out_roi_index,
                                    int32_t* out_count, void* ws, size_t ws_bytes, coin_stream_t stream) {
    COIN_REQUIRE(R >= 0 && k1 >= 2 && out_count, "det_postprocess: bad arguments");
    COIN_REQUIRE(kreg == 1 || kreg == k1 - 1, "det_postprocess: kreg must be 1 or the number of foreground classes");
    cudaStream_t s = as_stream(stream);
    if (R == 0 || out_capacity == 0 || topk == 0) {
        fill_bytes(out_count, 0, sizeof(int32_t), s);
        return COIN_OK;
    }
    COIN_REQUIRE(boxes && scores && out_boxes && out_scores && out_probs && out_classes && out_roi_index && ws,
                 "det_postprocess: null pointer");
    COIN_REQUIRE((reinterpret_cast<uintptr_t>(out_boxes) & 15) == 0, "det_postprocess: out_boxes must be 16-byte aligned");
    DpWs w = carve_dp(ws, R, k1);
    if (ws_bytes < w.total) return fail(COIN_ERR_CAPACITY, "det_postprocess: workspace too small (%zu < %zu)", ws_bytes, w.total);
    const int64_t cap = R * (k1 - 1);
    const unsigned rb = (unsigned)ceil_div(R, 128);
    dp_rows_kernel<<<rb, 128, 0, s>>>(boxes, scores, (int)R, k1, kreg, score_thresh, w.row_valid, w.row_cnt);
    if (int rc = check_launch("dp_rows_kernel")) return rc;
    dp_scan_kernel<<<1, 256, 0, s>>>(w.row_valid, w.row_cnt, (int)R, w.row_rank, w.row_off, w.n_cand);
    if (int rc = check_launch("dp_scan_kernel")) return rc;
    dp_compact_kernel<<<rb, 128, 0, s>>>(boxes, scores, (int)R, k1, kreg, img_h, img_w, score_thresh, w.row_valid,
                                         w.row_rank, w.row_off, w.cand_box, w.cand_score, w.cand_cls, w.cand_row, w.cand_src);
    if (int rc = check_launch("dp_compact_kernel")) return rc;
    const int64_t max_keep = topk >= 0 ? std::min(topk, out_capacity) : out_capacity;
    if (int rc = nms_sorted_pipeline(reinterpret_cast<const float*>(w.cand_box), w.cand_score, w.cand_cls, cap, w.n_cand,
                                     nms_thresh, COIN_NMS_AUTO, max_keep, w.keep, w.nkeep, w.nms_ws, w.nms_bytes, s))
        return rc;
    dp_gather_kernel<<<(unsigned)ceil_div(max_keep, 128), 128, 0, s>>>(
        w.keep, w.nkeep, out_capacity, w.cand_box, w.cand_score, w.cand_cls, w.cand_row, w.cand_src, scores, k1,
        reinterpret_cast<float4*>(out_boxes), out_scores, out_probs, out_classes, out_roi_index, out_count);
    return check_launch("dp_gather_kernel");
}
