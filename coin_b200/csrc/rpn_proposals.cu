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
//
// Selection (A <= 262 144 anchors): every 4096-logit chunk is sorted by its own CTA, then ONE kernel ranks each key
// against the other chunks by binary search and - for the keys whose rank is below pre_nms_topk - decodes the anchor
// straight into its sorted slot: two launches for top-k + decode (the radix sort this replaces was ~15).
// Anchors: either the caller's [A,4] array, or GENERATED ON THE FLY from the cell anchors and the grid of detectron2's
// DefaultAnchorGenerator (<- rpn.py:64): anchor(i) = cell[i % ncell] + (x * stride, y * stride, x * stride, y * stride),
// the same single fp32 addition grid_anchors performs, so the 41 625 x 16 bytes are never read.
#include <cub/device/device_radix_sort.cuh>

#include "sort_common.cuh"

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

constexpr int kMaxCell = 32;
constexpr int kMaxChunksRpn = 64;     // A <= 262144 anchors take the chunked selection

// detectron2 DefaultAnchorGenerator for one level: cell anchors + a stride grid (ncell == 0: read the anchors array)
struct AnchorGrid {
    float4 cell[kMaxCell];
    int ncell, wf;
    float stride, shift0;     // shift of grid position x: shift0 + x * stride  (shift0 = offset * stride)
};

struct DecodeArgs {
    const float4* anchors;
    const float4* deltas;
    const float* logits;
    float wx, wy, ww, wh, scale_clamp, img_h, img_w, min_size;
    float4* boxes;
    float* scores;
    int32_t* flags;
    int32_t* status;
};

// decode + clip + validity of anchor `src` into sorted slot i (same operation order as apply_deltas_kernel / the oracle)
__device__ __forceinline__ void rpn_decode_one(const DecodeArgs& a, const AnchorGrid& grid, const uint32_t src, const int i) {
    float4 b;
    if (grid.ncell > 0) {
        const uint32_t loc = src / (uint32_t)grid.ncell, c = src - loc * (uint32_t)grid.ncell;
        const uint32_t y = loc / (uint32_t)grid.wf, x = loc - y * (uint32_t)grid.wf;
        const float sx = grid.shift0 + (float)x * grid.stride, sy = grid.shift0 + (float)y * grid.stride;
        const float4 cb = grid.cell[c];
        b = make_float4(sx + cb.x, sy + cb.y, sx + cb.z, sy + cb.w);
    } else {
        b = __ldg(a.anchors + src);
    }
    const float4 d = __ldg(a.deltas + src);
    const float sc = __ldg(a.logits + src);
    const float wx = a.wx, wy = a.wy, ww = a.ww, wh = a.wh, scale_clamp = a.scale_clamp;
    const float img_w = a.img_w, img_h = a.img_h, min_size = a.min_size;
    float4* boxes = a.boxes;
    float* scores = a.scores;
    int32_t* flags = a.flags;
    int32_t* status = a.status;
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

// large-A path: the radix-sorted order is given
__global__ void rpn_decode_kernel(const uint32_t* __restrict__ sorted_idx, const DecodeArgs a, const AnchorGrid grid, int k) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= k) return;
    rpn_decode_one(a, grid, sorted_idx[i], i);
}

// chunked selection, step 1: keys (descending logit, ascending anchor index) of one 4096-logit chunk, sorted
__global__ void COIN_SORT_BOUNDS rpn_chunk_sort_kernel(const float* __restrict__ logits, int A, uint64_t* __restrict__ ckeys) {
    extern __shared__ uint64_t skeys[];
    const int base = blockIdx.x * kChunk;
    for (int i = threadIdx.x; i < kChunk; i += blockDim.x) {
        const int g = base + i;
        skeys[i] = g < A ? sort_key(__ldg(logits + g), (uint32_t)g) : ~0ull;
    }
    __syncthreads();
    int npow = 2;
    while (npow < min(A - base, kChunk)) npow <<= 1;
    bitonic_sort_smem<kSortThreads>(skeys, npow);
    for (int i = threadIdx.x; i < kChunk; i += blockDim.x) ckeys[base + i] = skeys[i];
}

// step 2: rank of every key = position in its chunk + number of smaller keys in the other chunks; the keys ranked below k
// (= the k largest logits, ties by anchor index) are decoded into slot `rank`. A key whose in-chunk position is already
// >= k cannot make it and skips the searches.
__global__ void rpn_rank_decode_kernel(const uint64_t* __restrict__ ckeys, int nchunks, const DecodeArgs a,
                                       const AnchorGrid grid, int k) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= nchunks * kChunk) return;
    const uint64_t key = ckeys[e];
    if (key == ~0ull) return;
    const int g = e / kChunk;
    int rank = e - g * kChunk;
    for (int h = 0; h < nchunks && rank < k; ++h) {
        if (h == g) continue;
        const uint64_t* c = ckeys + (size_t)h * kChunk;
        int lo = 0, hi = kChunk;
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (__ldg(c + mid) < key) lo = mid + 1; else hi = mid;
        }
        rank += lo;
    }
    if (rank < k) rpn_decode_one(a, grid, (uint32_t)(key & 0xffffffffu), rank);
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
    uint64_t* ckeys;
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
    w.ckeys = nullptr;
    w.keys = w.keys_alt = w.vals = w.vals_alt = nullptr;
    w.cub_tmp = nullptr;
    w.cub_bytes = 0;
    if (A <= (int64_t)kChunk * kMaxChunksRpn) {
        w.ckeys = c.take<uint64_t>((size_t)ceil_div(A, kChunk) * kChunk);
    } else {
        w.keys = c.take<uint32_t>((size_t)A); w.keys_alt = c.take<uint32_t>((size_t)A);
        w.vals = c.take<uint32_t>((size_t)A); w.vals_alt = c.take<uint32_t>((size_t)A);
        cub::DoubleBuffer<uint32_t> dk(nullptr, nullptr), dv(nullptr, nullptr);
        if (cub::DeviceRadixSort::SortPairs(nullptr, w.cub_bytes, dk, dv, (int)A) != cudaSuccess) {
            cudaGetLastError();
            w.cub_bytes = (size_t)A * 16 + (1 << 20);
        }
        w.cub_tmp = c.take<char>(w.cub_bytes);
    }
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

static int rpn_proposals_impl(const float* anchors, const AnchorGrid& grid, const float* deltas, const float* logits, int64_t A,
                              int64_t pre_nms_topk, int64_t post_nms_topk, double nms_thresh, float min_box_size,
                              float img_h, float img_w, float wx, float wy, float ww, float wh, float scale_clamp,
                              float* out_boxes, float* out_logits, int32_t* out_count, int32_t* status, void* ws,
                              size_t ws_bytes, coin_stream_t stream) {
    COIN_REQUIRE(A >= 0 && pre_nms_topk >= 0 && post_nms_topk >= 0 && out_count && status, "rpn_proposals: bad arguments");
    COIN_REQUIRE(A < (1ll << 31), "rpn_proposals: too many anchors");
    cudaStream_t s = as_stream(stream);
    fill_bytes(out_count, 0, sizeof(int32_t), s);
    fill_bytes(status, 0, sizeof(int32_t), s);
    const int64_t k = std::min(A, pre_nms_topk);
    if (k == 0 || post_nms_topk == 0) return COIN_OK;
    COIN_REQUIRE((anchors || grid.ncell > 0) && deltas && logits && out_boxes && out_logits && ws, "rpn_proposals: null pointer");
    COIN_REQUIRE(((reinterpret_cast<uintptr_t>(anchors) | reinterpret_cast<uintptr_t>(deltas) |
                   reinterpret_cast<uintptr_t>(out_boxes)) & 15) == 0, "rpn_proposals: boxes must be 16-byte aligned");
    RpnWs w = carve_rpn(ws, A, k);
    if (ws_bytes < w.total) return fail(COIN_ERR_CAPACITY, "rpn_proposals: workspace too small (%zu < %zu)", ws_bytes, w.total);
    DecodeArgs da;
    da.anchors = reinterpret_cast<const float4*>(anchors); da.deltas = reinterpret_cast<const float4*>(deltas); da.logits = logits;
    da.wx = wx; da.wy = wy; da.ww = ww; da.wh = wh; da.scale_clamp = scale_clamp; da.img_h = img_h; da.img_w = img_w;
    da.min_size = min_box_size; da.boxes = w.boxes; da.scores = w.scores; da.flags = w.flags; da.status = status;
    if (w.ckeys) {
        const int nchunks = (int)ceil_div(A, kChunk);
        rpn_chunk_sort_kernel<<<nchunks, kSortThreads, kChunk * sizeof(uint64_t), s>>>(logits, (int)A, w.ckeys);
        if (int rc = check_launch("rpn_chunk_sort_kernel")) return rc;
        rpn_rank_decode_kernel<<<(unsigned)ceil_div((int64_t)nchunks * kChunk, 256), 256, 0, s>>>(w.ckeys, nchunks, da, grid, (int)k);
        if (int rc = check_launch("rpn_rank_decode_kernel")) return rc;
    } else {
        rpn_keys_kernel<<<(unsigned)ceil_div(A, 256), 256, 0, s>>>(logits, (int)A, w.keys, w.vals);
        if (int rc = check_launch("rpn_keys_kernel")) return rc;
        cub::DoubleBuffer<uint32_t> dk(w.keys, w.keys_alt), dv(w.vals, w.vals_alt);
        size_t bytes = w.cub_bytes;
        cudaError_t e = cub::DeviceRadixSort::SortPairs(w.cub_tmp, bytes, dk, dv, (int)A, 0, 32, s);
        if (e != cudaSuccess) return fail(COIN_ERR_CUDA, "rpn_proposals: radix sort failed: %s", cudaGetErrorString(e));
        count_launch();
        rpn_decode_kernel<<<(unsigned)ceil_div(k, 256), 256, 0, s>>>(dv.Current(), da, grid, (int)k);
        if (int rc = check_launch("rpn_decode_kernel")) return rc;
    }
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

extern "C" int coin_rpn_proposals(const float* anchors, const float* deltas, const float* logits, int64_t A,
                                  int64_t pre_nms_topk, int64_t post_nms_topk, double nms_thresh, float min_box_size,
                                  float img_h, float img_w, float wx, float wy, float ww, float wh,
                                  float scale_clamp, float* out_boxes, float* out_logits, int32_t* out_count,
                                  int32_t* status, void* ws, size_t ws_bytes, coin_stream_t stream) {
    AnchorGrid grid;
    grid.ncell = 0; grid.wf = 1; grid.stride = 0.0f; grid.shift0 = 0.0f;
    return rpn_proposals_impl(anchors, grid, deltas, logits, A, pre_nms_topk, post_nms_topk, nms_thresh, min_box_size, img_h,
                              img_w, wx, wy, ww, wh, scale_clamp, out_boxes, out_logits, out_count, status, ws, ws_bytes, stream);
}

extern "C" int coin_rpn_proposals_grid(const float* cell_anchors_host, int ncell, int Hf, int Wf, float stride, float offset,
                                       const float* deltas, const float* logits, int64_t pre_nms_topk, int64_t post_nms_topk,
                                       double nms_thresh, float min_box_size, float img_h, float img_w, float wx, float wy,
                                       float ww, float wh, float scale_clamp, float* out_boxes, float* out_logits,
                                       int32_t* out_count, int32_t* status, void* ws, size_t ws_bytes, coin_stream_t stream) {
    COIN_REQUIRE(cell_anchors_host && ncell >= 1 && ncell <= kMaxCell && Hf >= 0 && Wf >= 0 && stride > 0.0f,
                 "rpn_proposals_grid: bad anchor grid (1 <= ncell <= %d)", kMaxCell);
    AnchorGrid grid;
    for (int c = 0; c < ncell; ++c)
        grid.cell[c] = make_float4(cell_anchors_host[4 * c], cell_anchors_host[4 * c + 1], cell_anchors_host[4 * c + 2],
                                   cell_anchors_host[4 * c + 3]);
    grid.ncell = ncell; grid.wf = std::max(Wf, 1); grid.stride = stride; grid.shift0 = offset * stride;
    return rpn_proposals_impl(nullptr, grid, deltas, logits, (int64_t)Hf * Wf * ncell, pre_nms_topk, post_nms_topk, nms_thresh,
                              min_box_size, img_h, img_w, wx, wy, ww, wh, scale_clamp, out_boxes, out_logits, out_count, status,
                              ws, ws_bytes, stream);
}
