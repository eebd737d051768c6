// rpn_proposals.cu -- RPN proposal selection for one image and one feature level, with no host sync.
//
// Replaces detectron2 0.5 RPN._decode_proposals + find_top_rpn_proposals as reached from
// coin/modeling/proposal_generator/rpn.py:64,113 (DualTeacherRPN.forward -> predict_proposals), SURVEY 8(f) rank 1:
//   sort the objectness logits (descending), keep the first pre_nms_topk, Box2BoxTransform.apply_deltas on the
//   selected anchors, drop non-finite rows, Boxes.clip, Boxes.nonempty(min_box_size), batched_nms (one level =
//   plain NMS), keep[:post_nms_topk].
// The reference decodes ALL H*W*A anchors (~20 ATen launches over 41 625 boxes), sorts, indexes, calls nonempty()
// (a host sync on keep.sum().item()) and torchvision NMS (another sync). Here only the selected anchors are
// decoded, the live count stays on the device and feeds the NMS pipeline of nms.cu directly.
#include <cub/device/device_radix_sort.cuh>

#include "common.cuh"

namespace coin {

size_t nms_pipeline_workspace_bytes(int64_t n_cap);
int nms_sorted_pipeline(const float* boxes, const float* scores, const int64_t* idxs, int64_t n_cap,
                        const int32_t* n_dev, double thr, int strategy, int64_t max_keep, int64_t* keep,
                        int32_t* nkeep, void* ws, size_t ws_bytes, cudaStream_t s);

// descending logit, ties by ascending anchor index (stable radix sort over index-ordered input)
__global__ void rpn_keys_kernel(const float* __restrict__ logits, int A, uint32_t* __restrict__ keys,
                                uint32_t* __restrict__ vals) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= A) return;
    float s = __ldg(logits + i) + 0.0f;
    uint32_t u = __float_as_uint(s);
    if (s != s) u = 0x7fc00000u;                       // NaN first, as torch's descending sort
    u ^= (u >> 31) ? 0xffffffffu : 0x80000000u;
    keys[i] = ~u;
    vals[i] = (uint32_t)i;
}

// decode + clip + validity of the k best anchors (same operation order as apply_deltas_kernel / the oracle)
__global__ void rpn_decode_kernel(const uint32_t* __restrict__ sorted_idx, const float4* __restrict__ anchors,
                                  const float4* __restrict__ deltas, const float* __restrict__ logits, int k,
                                  float wx, float wy, float ww, float wh, float scale_clamp, float img_h,
                                  float img_w, float min_size, float4* __restrict__ boxes, float* __restrict__ scores,
                                  int32_t* __restrict__ flags, int32_t* __restrict__ status) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= k) return;
    const uint32_t src = sorted_idx[i];
    const float4 b = __ldg(anchors + src);
    const float4 d = __ldg(deltas + src);
    const float sc = __ldg(logits + src);
    const float w = b.z - b.x, h = b.w - b.y;
    const float cx = b.x + 0.5f * w, cy = b.y + 0.5f * h;
    const float dx = d.x / wx, dy = d.y / wy;
    const float dw = fminf(d.z / ww, scale_clamp), dh = fminf(d.w / wh, scale_clamp);
    const float pcx = dx * w + cx, pcy = dy * h + cy;
    const float pw = expf(dw) * w, ph = expf(dh) * h;
    float4 o = make_float4(pcx - 0.5f * pw, pcy - 0.5f * ph, pcx + 0.5f * pw, pcy + 0.5f * ph);
    const bool finite = isfinite(o.x) && isfinite(o.y) && isfinite(o.z) && isfinite(o.w) && isfinite(sc);
    if (!finite) atomicOr(status, 1);
    o.x = fminf(fmaxf(o.x, 0.0f), img_w);
    o.y = fminf(fmaxf(o.y, 0.0f), img_h);
    o.z = fminf(fmaxf(o.z, 0.0f), img_w);
    o.w = fminf(fmaxf(o.w, 0.0f), img_h);
    const bool nonempty = (o.z - o.x) > min_size && (o.w - o.y) > min_size;
    boxes[i] = o;
    scores[i] = sc;
    flags[i] = finite && nonempty;
}

// order-preserving compaction of the flagged rows (single CTA, warp ballots; k <= a few 10^4)
__global__ void __launch_bounds__(256)
rpn_compact_kernel(const float4* __restrict__ boxes, const float* __restrict__ scores, const int32_t* __restrict__ flags,
                   int k, float4* __restrict__ cboxes, float* __restrict__ cscores, int32_t* __restrict__ n_live) {
    __shared__ int wcount[8];
    __shared__ int base;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) base = 0;
    __syncthreads();
    for (int i0 = 0; i0 < k; i0 += blockDim.x) {
        const int i = i0 + threadIdx.x;
        const bool f = i < k && flags[i] != 0;
        const unsigned m = __ballot_sync(0xffffffffu, f);
        if (lane == 0) wcount[warp] = __popc(m);
        __syncthreads();
        int before = base;
        for (int q = 0; q < warp; ++q) before += wcount[q];
        if (f) {
            const int at = before + __popc(m & ((1u << lane) - 1u));
            cboxes[at] = boxes[i];
            cscores[at] = scores[i];
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            int tot = 0;
            for (int q = 0; q < (int)(blockDim.x >> 5); ++q) tot += wcount[q];
            base += tot;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) *n_live = base;
}

__global__ void rpn_gather_kernel(const int64_t* __restrict__ keep, const int32_t* __restrict__ nkeep,
                                  const float4* __restrict__ cboxes, const float* __restrict__ cscores, int cap,
                                  float4* __restrict__ out_boxes, float* __restrict__ out_logits,
                                  int32_t* __restrict__ out_count) {
    const int n = min(*nkeep, cap);
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) *out_count = n;
    if (i >= n) return;
    const int64_t src = keep[i];
    out_boxes[i] = cboxes[src];
    out_logits[i] = cscores[src];
}

struct RpnWs {
    uint32_t *keys, *keys_alt, *vals, *vals_alt;
    void* cub_tmp;
    size_t cub_bytes;
    float4 *boxes, *cboxes;
    float *scores, *cscores;
    int32_t *flags, *n_live, *nkeep;
    int64_t* keep;
    void* nms_ws;
    size_t nms_bytes, total;
};

static RpnWs carve_rpn(void* ws, int64_t A, int64_t k) {
    RpnWs w;
    Carver c(ws);
    w.keys = c.take<uint32_t>((size_t)A); w.keys_alt = c.take<uint32_t>((size_t)A);
    w.vals = c.take<uint32_t>((size_t)A); w.vals_alt = c.take<uint32_t>((size_t)A);
    w.cub_bytes = 0;
    cub::DoubleBuffer<uint32_t> dk(nullptr, nullptr), dv(nullptr, nullptr);
    if (cub::DeviceRadixSort::SortPairs(nullptr, w.cub_bytes, dk, dv, (int)A) != cudaSuccess) {
        cudaGetLastError();
        w.cub_bytes = (size_t)A * 16 + (1 << 20);
    }
    w.cub_tmp = c.take<char>(w.cub_bytes);
    w.boxes = c.take<float4>((size_t)k); w.cboxes = c.take<float4>((size_t)k);
    w.scores = c.take<float>((size_t)k); w.cscores = c.take<float>((size_t)k);
    w.flags = c.take<int32_t>((size_t)k);
    w.n_live = c.take<int32_t>(16);
    w.nkeep = w.n_live + 4;
    w.keep = c.take<int64_t>((size_t)k);
    w.nms_bytes = nms_pipeline_workspace_bytes(k);
    w.nms_ws = c.take<char>(w.nms_bytes);
    w.total = c.used();
    return w;
}

}  // namespace coin
using namespace coin;

extern "C" size_t coin_rpn_proposals_workspace_bytes(int64_t A, int64_t pre_nms_topk) {
    if (A <= 0) return 256;
    return carve_rpn(nullptr, A, std::min<int64_t>(A, std::max<int64_t>(pre_nms_topk, 1))).total + 256;
}

extern "C" int coin_rpn_proposals(const float* anchors, const float* deltas, const float* logits, int64_t A,
                                  int64_t pre_nms_topk, int64_t post_nms_topk, double nms_thresh, float min_box_size,
                                  float img_h, float img_w, float wx, float wy, float ww, float wh,
                                  float scale_clamp, float* out_boxes, float* out_logits, int32_t* out_count,
                                  int32_t* status, void* ws, size_t ws_bytes, coin_stream_t stream) {
    COIN_REQUIRE(A >= 0 && pre_nms_topk >= 0 && post_nms_topk >= 0 && out_count && status, "rpn_proposals: bad arguments");
    COIN_REQUIRE(A < (1ll << 31), "rpn_proposals: too many anchors");
    cudaStream_t s = as_stream(stream);
    fill_bytes(out_count, 0, sizeof(int32_t), s);
    fill_bytes(status, 0, sizeof(int32_t), s);
    const int64_t k = std::min(A, pre_nms_topk);
    if (k == 0 || post_nms_topk == 0) return COIN_OK;
    COIN_REQUIRE(anchors && deltas && logits && out_boxes && out_logits && ws, "rpn_proposals: null pointer");
    COIN_REQUIRE(((reinterpret_cast<uintptr_t>(anchors) | reinterpret_cast<uintptr_t>(deltas) |
                   reinterpret_cast<uintptr_t>(out_boxes)) & 15) == 0, "rpn_proposals: boxes must be 16-byte aligned");
    RpnWs w = carve_rpn(ws, A, k);
    if (ws_bytes < w.total) return fail(COIN_ERR_CAPACITY, "rpn_proposals: workspace too small (%zu < %zu)", ws_bytes, w.total);
    rpn_keys_kernel<<<(unsigned)ceil_div(A, 256), 256, 0, s>>>(logits, (int)A, w.keys, w.vals);
    if (int rc = check_launch("rpn_keys_kernel")) return rc;
    cub::DoubleBuffer<uint32_t> dk(w.keys, w.keys_alt), dv(w.vals, w.vals_alt);
    size_t bytes = w.cub_bytes;
    cudaError_t e = cub::DeviceRadixSort::SortPairs(w.cub_tmp, bytes, dk, dv, (int)A, 0, 32, s);
    if (e != cudaSuccess) return fail(COIN_ERR_CUDA, "rpn_proposals: radix sort failed: %s", cudaGetErrorString(e));
    count_launch();
    rpn_decode_kernel<<<(unsigned)ceil_div(k, 256), 256, 0, s>>>(
        dv.Current(), reinterpret_cast<const float4*>(anchors), reinterpret_cast<const float4*>(deltas), logits, (int)k, wx,
        wy, ww, wh, scale_clamp, img_h, img_w, min_box_size, w.boxes, w.scores, w.flags, status);
    if (int rc = check_launch("rpn_decode_kernel")) return rc;
    rpn_compact_kernel<<<1, 256, 0, s>>>(w.boxes, w.scores, w.flags, (int)k, w.cboxes, w.cscores, w.n_live);
    if (int rc = check_launch("rpn_compact_kernel")) return rc;
    if (int rc = nms_sorted_pipeline(reinterpret_cast<const float*>(w.cboxes), w.cscores, nullptr, k, w.n_live, nms_thresh,
                                     COIN_NMS_PLAIN, post_nms_topk, w.keep, w.nkeep, w.nms_ws, w.nms_bytes, s))
        return rc;
    const int64_t cap = std::min(k, post_nms_topk);
    rpn_gather_kernel<<<(unsigned)ceil_div(cap, 256), 256, 0, s>>>(w.keep, w.nkeep, w.cboxes, w.cscores, (int)cap,
                                                                   reinterpret_cast<float4*>(out_boxes), out_logits, out_count);
    return check_launch("rpn_gather_kernel");
}
