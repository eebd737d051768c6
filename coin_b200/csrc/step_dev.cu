// step_dev.cu -- sync-free glue of the RoI-path step: device-count row concatenation and the field
// gathers of the knowledge-separation result.
//
// In the reference every stage boundary of CoinTrainer.run_step is a host round trip: nonzero() /
// tolist() / len() on device tensors (coin/engine/trainer.py:364-391,469; clip_roi_heads.py:345-362;
// rpn.py:209-228) decide the SHAPES of the next stage's inputs. Here the variable-length sets stay in
// worst-case buffers with their lengths in device memory, so the whole step is a fixed launch sequence
// that can be captured in a CUDA graph:
//   coin_concat_rows   torch.cat of pseudo-GT / proposal box sets whose lengths are device counts
//                      (clip_roi_heads.py:345-353 add_ground_truth_to_proposals + Boxes.cat, rpn.py:209-212),
//                      optionally prefixing the batch-index column of convert_boxes_to_pooler_format;
//   coin_abc_pack      the A / B / C Instances fields of match_dual_teacher (trainer.py:393-455) gathered
//                      from the index lists coin_match_abc emits.
#include "common.cuh"

namespace coin {

constexpr int kMaxSegs = 8;

struct ConcatArgs {
    const float* ptr[kMaxSegs];
    const int32_t* count_dev[kMaxSegs];
    int count[kMaxSegs];      // host count, or the capacity when count_dev is set
    float prefix[kMaxSegs];
    int nseg, width_in, width_out;
    float* out;
    int32_t* out_count;
    int out_cap;
};

__global__ void concat_rows_kernel(const ConcatArgs a) {
    __shared__ int s_start[kMaxSegs + 1];
    if (threadIdx.x == 0) {
        int acc = 0;
        for (int s = 0; s < a.nseg; ++s) {
            s_start[s] = acc;
            const int n = a.count_dev[s] ? min(max(__ldg(a.count_dev[s]), 0), a.count[s]) : a.count[s];
            acc += n;
        }
        s_start[a.nseg] = acc;
        if (blockIdx.x == 0 && a.out_count) *a.out_count = min(acc, a.out_cap);
    }
    __syncthreads();
    const int total = min(s_start[a.nseg], a.out_cap);
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < total; r += gridDim.x * blockDim.x) {
        int s = 0;
        while (s + 1 < a.nseg && r >= s_start[s + 1]) ++s;
        const float* src = a.ptr[s] + (size_t)(r - s_start[s]) * a.width_in;
        float* dst = a.out + (size_t)r * a.width_out;
        int o = 0;
        if (a.width_out > a.width_in) dst[o++] = a.prefix[s];
        for (int i = 0; i < a.width_in; ++i) dst[o + i] = src[i];
    }
}

struct PackArgs {
    // detection sets (online = cloud, offline = CLIP detector)
    const float4 *on_boxes, *off_boxes;
    const int64_t *on_cls, *off_cls;
    const float *on_scores, *off_scores, *on_probs, *off_probs;
    int nc, nd_cap, k1, tag;
    const int32_t* nd_dev;
    // index lists and counts from coin_match_abc
    const int32_t *a_on, *a_off, *b_on, *b_off, *c_on, *c_off, *counts;
    // outputs (capacity rows each)
    int64_t *a_cls, *b_cls_off, *b_cls_on, *c_cls;
    float *a_s_on, *a_s_off, *a_p_on, *a_p_off;
    float *b_s_on, *b_s_off, *b_p_on, *b_p_off;
    float4* c_boxes;
    float *c_scores, *c_probs;
};

// One launch per (image, tag): blockIdx.y selects the set (0 = A, 1 = B, 2 = C); threads stride over
// (row, field column) items. In the empty-side branches (trainer.py:343-361) both members of a pair
// index the non-empty detection set.
__global__ void abc_pack_kernel(const PackArgs a) {
    const int nd = a.nd_dev ? min(max(__ldg(a.nd_dev), 0), a.nd_cap) : a.nd_cap;
    const bool on_empty = a.nc == 0, off_empty = nd == 0;
    const int64_t* ONC = on_empty ? a.off_cls : a.on_cls;
    const float* ONS = on_empty ? a.off_scores : a.on_scores;
    const float* ONP = on_empty ? a.off_probs : a.on_probs;
    const int64_t* OFC = off_empty ? a.on_cls : a.off_cls;
    const float* OFS = off_empty ? a.on_scores : a.off_scores;
    const float* OFP = off_empty ? a.on_probs : a.off_probs;
    const int k1 = a.k1;
    const int set = blockIdx.y;
    const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
    if (set < 2) {
        if (set == 1 && a.tag != COIN_TAG_RCNN) return;
        const int n = a.counts[set];
        const int32_t* on = set == 0 ? a.a_on : a.b_on;
        const int32_t* off = set == 0 ? a.a_off : a.b_off;
        int64_t* cls_off = set == 0 ? a.a_cls : a.b_cls_off;
        float* s_on = set == 0 ? a.a_s_on : a.b_s_on;
        float* s_off = set == 0 ? a.a_s_off : a.b_s_off;
        float* p_on = set == 0 ? a.a_p_on : a.b_p_on;
        float* p_off = set == 0 ? a.a_p_off : a.b_p_off;
        for (int it = tid; it < n * (k1 + 1); it += nth) {
            const int r = it / (k1 + 1), c = it - r * (k1 + 1);
            const int io = on[r], jf = off[r];
            if (c == k1) {
                cls_off[r] = OFC[jf];
                if (set == 1) a.b_cls_on[r] = ONC[io];
                s_on[r] = ONS[io];
                s_off[r] = OFS[jf];
            } else {
                p_on[(size_t)r * k1 + c] = ONP[(size_t)io * k1 + c];
                p_off[(size_t)r * k1 + c] = OFP[(size_t)jf * k1 + c];
            }
        }
    } else {
        // C rows [0, nC_off) reference the CLIP-detector set, rows [nC_off, nC) the cloud set
        const int n = a.counts[2], n_off = a.counts[4];
        for (int it = tid; it < n * (k1 + 1); it += nth) {
            const int r = it / (k1 + 1), c = it - r * (k1 + 1);
            const bool from_off = r < n_off;
            const int src = from_off ? a.c_off[r] : a.c_on[r];
            if (c == k1) {
                a.c_boxes[r] = from_off ? a.off_boxes[src] : a.on_boxes[src];
                a.c_cls[r] = from_off ? a.off_cls[src] : a.on_cls[src];
                a.c_scores[r] = from_off ? a.off_scores[src] : a.on_scores[src];
            } else {
                a.c_probs[(size_t)r * k1 + c] = from_off ? a.off_probs[(size_t)src * k1 + c] : a.on_probs[(size_t)src * k1 + c];
            }
        }
    }
}

// ---- live prefixes of many result buffers -> one contiguous staging buffer (one D2H copy per step) --------------------
struct PackItem {               // 32 bytes; the table lives in device memory (built once by the caller)
    const uint8_t* ptr;
    int64_t row_bytes;
    int64_t cap_rows;
    int32_t count_index;        // index into the step's counts vector, -1: all cap_rows rows are live
    int32_t pad;
};

__global__ void __launch_bounds__(256)
pack_rows_kernel(const PackItem* __restrict__ items, int n, const int32_t* __restrict__ counts, uint8_t* __restrict__ packed,
                 long long cap, long long* __restrict__ offsets) {
    __shared__ long long s_part[256];
    auto live_bytes = [&](int j) {
        const PackItem it = items[j];
        long long rows = it.cap_rows;
        if (it.count_index >= 0) rows = min((long long)max(counts[it.count_index], 0), (long long)it.cap_rows);
        return (rows * it.row_bytes + 15) / 16 * 16;      // every segment starts 16-byte aligned
    };
    const int b = blockIdx.x;
    long long part = 0;
    for (int j = threadIdx.x; j < b; j += 256) part += live_bytes(j);
    s_part[threadIdx.x] = part;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) s_part[threadIdx.x] += s_part[threadIdx.x + o];
        __syncthreads();
    }
    const long long off = s_part[0];
    const long long mine = live_bytes(b);
    if (threadIdx.x == 0) {
        offsets[b] = off;
        if (b == n - 1) offsets[n] = off + mine;
    }
    if (off + mine > cap) return;                         // the caller sized `packed` for the capacities: cannot happen
    const PackItem it = items[b];
    long long rows = it.cap_rows;
    if (it.count_index >= 0) rows = min((long long)max(counts[it.count_index], 0), (long long)it.cap_rows);
    const long long nbytes = rows * it.row_bytes;
    uint8_t* dst = packed + off;
    if ((reinterpret_cast<uintptr_t>(it.ptr) & 15) == 0) {
        const long long n16 = nbytes / 16;
        const uint4* s4 = reinterpret_cast<const uint4*>(it.ptr);
        uint4* d4 = reinterpret_cast<uint4*>(dst);
        for (long long i = threadIdx.x; i < n16; i += 256) d4[i] = s4[i];
        for (long long i = n16 * 16 + threadIdx.x; i < nbytes; i += 256) dst[i] = it.ptr[i];
    } else {
        for (long long i = threadIdx.x; i < nbytes; i += 256) dst[i] = it.ptr[i];
    }
}

}  // namespace coin
using namespace coin;

extern "C" int coin_pack_rows(const void* items_dev, int n_items, const int32_t* counts_dev, void* packed, int64_t packed_cap,
                              int64_t* offsets_dev, coin_stream_t stream) {
    COIN_REQUIRE(n_items >= 0 && packed_cap >= 0, "pack_rows: bad arguments");
    if (n_items == 0) return COIN_OK;
    COIN_REQUIRE(items_dev && packed && offsets_dev, "pack_rows: null pointer");
    COIN_REQUIRE((reinterpret_cast<uintptr_t>(packed) & 15) == 0, "pack_rows: packed must be 16-byte aligned");
    pack_rows_kernel<<<(unsigned)n_items, 256, 0, as_stream(stream)>>>(static_cast<const PackItem*>(items_dev), n_items, counts_dev,
                                                                     static_cast<uint8_t*>(packed), (long long)packed_cap,
                                                                     reinterpret_cast<long long*>(offsets_dev));
    return check_launch("pack_rows_kernel");
}

extern "C" int coin_concat_rows(const coin_seg_t* segs_host, int nseg, int width_in, int width_out, float* out,
                                int64_t out_cap, int32_t* out_count, coin_stream_t stream) {
    COIN_REQUIRE(segs_host && nseg >= 1 && nseg <= kMaxSegs, "concat_rows: nseg=%d out of [1,%d]", nseg, kMaxSegs);
    COIN_REQUIRE(width_in >= 1 && (width_out == width_in || width_out == width_in + 1), "concat_rows: bad widths");
    COIN_REQUIRE(out_cap >= 0 && out_cap < (1ll << 31), "concat_rows: bad capacity");
    ConcatArgs a;
    int64_t worst = 0;
    for (int s = 0; s < nseg; ++s) {
        COIN_REQUIRE(segs_host[s].count >= 0 && segs_host[s].count < (1ll << 31), "concat_rows: bad segment count");
        COIN_REQUIRE(segs_host[s].count == 0 || segs_host[s].ptr, "concat_rows: null segment");
        a.ptr[s] = segs_host[s].ptr;
        a.count_dev[s] = segs_host[s].count_dev;
        a.count[s] = (int)segs_host[s].count;
        a.prefix[s] = segs_host[s].prefix;
        worst += segs_host[s].count;
    }
    COIN_REQUIRE(worst == 0 || out, "concat_rows: out is null");
    a.nseg = nseg; a.width_in = width_in; a.width_out = width_out;
    a.out = out; a.out_count = out_count; a.out_cap = (int)out_cap;
    const int64_t rows = std::min<int64_t>(worst, out_cap);
    const unsigned blocks = (unsigned)std::max<int64_t>(1, std::min<int64_t>(ceil_div(rows, 256), 4 * kNumSMs));
    concat_rows_kernel<<<blocks, 256, 0, as_stream(stream)>>>(a);
    return check_launch("concat_rows_kernel");
}

extern "C" int coin_abc_pack(const coin_dets_t* online_host, int64_t nc, const coin_dets_t* offline_host,
                             int64_t nd_cap, const int32_t* nd_dev, int k1, int tag, const int32_t* a_on,
                             const int32_t* a_off, const int32_t* b_on, const int32_t* b_off, const int32_t* c_on,
                             const int32_t* c_off, const int32_t* counts, const coin_pseudo_t* a_out_host,
                             const coin_pseudo_t* b_out_host, const coin_pseudo_t* c_out_host, int64_t cap_pairs,
                             coin_stream_t stream) {
    COIN_REQUIRE(online_host && offline_host && a_out_host && c_out_host && counts, "abc_pack: null argument");
    COIN_REQUIRE(nc >= 0 && nd_cap >= 0 && k1 >= 1, "abc_pack: bad sizes");
    COIN_REQUIRE(tag == COIN_TAG_RCNN || tag == COIN_TAG_RPN, "abc_pack: bad tag %d", tag);
    COIN_REQUIRE(tag != COIN_TAG_RCNN || b_out_host, "abc_pack: B outputs are required for tag RCNN");
    if (nc == 0 && nd_cap == 0) return COIN_OK;
    PackArgs a;
    a.on_boxes = reinterpret_cast<const float4*>(online_host->boxes);
    a.off_boxes = reinterpret_cast<const float4*>(offline_host->boxes);
    a.on_cls = online_host->classes; a.off_cls = offline_host->classes;
    a.on_scores = online_host->scores; a.off_scores = offline_host->scores;
    a.on_probs = online_host->probs; a.off_probs = offline_host->probs;
    a.nc = (int)nc; a.nd_cap = (int)nd_cap; a.k1 = k1; a.tag = tag; a.nd_dev = nd_dev;
    a.a_on = a_on; a.a_off = a_off; a.b_on = b_on; a.b_off = b_off; a.c_on = c_on; a.c_off = c_off; a.counts = counts;
    a.a_cls = a_out_host->classes; a.a_s_on = a_out_host->scores_online; a.a_s_off = a_out_host->scores_offline;
    a.a_p_on = a_out_host->probs_online; a.a_p_off = a_out_host->probs_offline;
    if (tag == COIN_TAG_RCNN) {
        a.b_cls_off = b_out_host->classes; a.b_cls_on = b_out_host->classes_online;
        a.b_s_on = b_out_host->scores_online; a.b_s_off = b_out_host->scores_offline;
        a.b_p_on = b_out_host->probs_online; a.b_p_off = b_out_host->probs_offline;
    } else {
        a.b_cls_off = a.b_cls_on = nullptr;
        a.b_s_on = a.b_s_off = a.b_p_on = a.b_p_off = nullptr;
    }
    a.c_boxes = reinterpret_cast<float4*>(c_out_host->boxes); a.c_cls = c_out_host->classes;
    a.c_scores = c_out_host->scores_online; a.c_probs = c_out_host->probs_online;
    (void)cap_pairs;
    dim3 grid(8, 3);
    abc_pack_kernel<<<grid, 256, 0, as_stream(stream)>>>(a);
    return check_launch("abc_pack_kernel");
}
