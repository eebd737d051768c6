// fusion_nms.cu -- COIN's probabilistic-fusion NMS (MyNMS) as ONE single-CTA kernel.
//
// Replaces the Python `while` loop of coin/layers/nms.py:84-194 (nms_bayesian) with its wrappers
// batch_nms_bayesian (:196-203) and Probabilistic_Fusion (:213-238), called per image from
// coin/modeling/meta_arch/gdino_processor.py:164-182 on the cloud detector's ~10-100 boxes. The
// reference spends ~25 tensor ops and several host syncs per kept box; here the whole call is one
// launch: per-class offset, legacy "+1" IoU, greedy clustering in descending-score order, fused
// score / probability vector / box per cluster, final re-sort by fused score.
//
// Order of the members of a cluster is the reference's: matched boxes in sweep order, the pivot
// last (nms.py:130-133); sums run sequentially in that order.
//
// Up to 4096 boxes (every call COIN makes: <= 900 GDINO queries per pass) the working set lives in shared memory. Above
// that - the reference takes up to 39 999 boxes in one batched call and unbounded per-class subsets beyond (nms.py:213-238)
// - the same kernel runs with its arrays in the caller's workspace (template parameter): still one launch and one CTA,
// tens of milliseconds at 40 000 boxes, where the reference's Python loop needs seconds.
#include "common.cuh"

namespace coin {

constexpr int kFusionMax = 4096;

struct FusionArgs {
    const float4* boxes;
    const float* probs;
    const int64_t* labels;
    int n, k1;
    float thr;
    int score_method, box_method, per_class_offset;
    int64_t* keep;
    float4* out_boxes;
    float* out_scores;
    float* out_probs;
    int64_t* out_classes;
    int32_t* nkeep;
    int32_t* status;
    // working set of the large-n variant (the small one keeps these in shared memory)
    uint64_t* g_keys;    // [npow]
    float4* g_nbox;      // [n]
    float* g_area;       // [n]
    float* g_sscore;     // [n]
    int32_t* g_src;      // [n]
    int32_t* g_alive;    // [n]
    // global scratch
    int32_t* cid;        // [n] sorted position of the pivot owning each sorted position
    int32_t* pivots;     // [n] sorted positions of the pivots, sweep order
    float4* f_box;       // [n] fused box per pivot (sweep order)
    float* f_score;      // [n]
    float* f_prob;       // [n, k1]
    int64_t* f_cls;      // [n]
};

__device__ __forceinline__ uint64_t desc_key(float s, uint32_t idx) {
    s = s + 0.0f;
    uint32_t u = __float_as_uint(s);
    if (s != s) u = 0x7fc00000u;
    u ^= (u >> 31) ? 0xffffffffu : 0x80000000u;
    return ((uint64_t)(~u) << 32) | idx;
}

__device__ void bitonic_sort_smem(uint64_t* keys, int npow) {
    for (int k = 2; k <= npow; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < npow; i += blockDim.x) {
                const int p = i ^ j;
                if (p > i) {
                    const uint64_t a = keys[i], b = keys[p];
                    if ((a > b) == ((i & k) == 0)) { keys[i] = b; keys[p] = a; }
                }
            }
            __syncthreads();
        }
    }
}

template <bool GLOBAL>
__global__ void __launch_bounds__(1024) fusion_nms_kernel(const FusionArgs a) {
    extern __shared__ unsigned char smem_raw[];
    const int n = a.n, k1 = a.k1;
    int npow = 2;
    while (npow < n) npow <<= 1;
    uint64_t* keys = GLOBAL ? a.g_keys : reinterpret_cast<uint64_t*>(smem_raw);       // [npow]
    float4* nbox = GLOBAL ? a.g_nbox : reinterpret_cast<float4*>(keys + npow);         // [n] offset boxes, sorted order
    float* area = GLOBAL ? a.g_area : reinterpret_cast<float*>(nbox + n);              // [n]
    float* sscore = GLOBAL ? a.g_sscore : area + n;                                    // [n] score, sorted order
    int32_t* src = GLOBAL ? a.g_src : reinterpret_cast<int32_t*>(sscore + n);          // [n] original index
    volatile int32_t* alive = GLOBAL ? a.g_alive : src + n;                            // [n]
    __shared__ float s_red[32];
    __shared__ int s_next, s_npiv;

    // 1. score = probs[i, label[i]], max coordinate, sort keys
    float m = -INFINITY;
    for (int i = threadIdx.x; i < npow; i += blockDim.x) {
        if (i < n) {
            const float sc = a.probs[(size_t)i * k1 + (int)a.labels[i]];
            keys[i] = desc_key(sc, (uint32_t)i);
            const float4 b = a.boxes[i];
            m = fmaxf(fmaxf(m, fmaxf(b.x, b.y)), fmaxf(b.z, b.w));
        } else {
            keys[i] = ~0ull;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = m;
    __syncthreads();
    float mx = s_red[0];
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) mx = fmaxf(mx, s_red[w]);
    bitonic_sort_smem(keys, npow);

    // 2. gather in sorted order; legacy "+1" areas on the offset boxes (nms.py:86-91,199-201)
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const int o = (int)(keys[i] & 0xffffffffu);
        float4 b = a.boxes[o];
        if (a.per_class_offset) {
            const float off = (float)a.labels[o] * (mx + 1.0f);
            b.x += off; b.y += off; b.z += off; b.w += off;
        }
        nbox[i] = b;
        area[i] = (b.z - b.x + 1.0f) * (b.w - b.y + 1.0f);
        sscore[i] = a.probs[(size_t)o * k1 + (int)a.labels[o]];
        src[i] = o;
        alive[i] = 1;
        a.cid[i] = -1;
    }
    if (threadIdx.x == 0) { s_next = 0; s_npiv = 0; }
    __syncthreads();

    // 3. greedy sweep: the first alive box is the pivot; alive boxes with ovr > thr join its cluster
    int p = 0;
    while (true) {
        while (p < n && !alive[p]) ++p;   // uniform: every thread reads the same shared flags
        if (p >= n) break;
        const float4 bp = nbox[p];
        const float ap = area[p];
        const int64_t lp = a.labels[src[p]];
        __syncthreads();                  // all threads have found p before flags change
        for (int q = p + 1 + threadIdx.x; q < n; q += blockDim.x) {
            if (!alive[q]) continue;
            if (!a.per_class_offset && a.labels[src[q]] != lp) continue;
            const float4 bq = nbox[q];
            const float w = fmaxf(0.0f, fminf(bp.z, bq.z) - fmaxf(bp.x, bq.x) + 1.0f);
            const float h = fmaxf(0.0f, fminf(bp.w, bq.w) - fmaxf(bp.y, bq.y) + 1.0f);
            const float inter = w * h;
            const float ovr = inter / (ap + area[q] - inter);
            if (ovr > a.thr) { alive[q] = 0; a.cid[q] = p; }
        }
        if (threadIdx.x == 0) {
            alive[p] = 0;
            a.cid[p] = p;
            a.pivots[s_npiv++] = p;
        }
        __syncthreads();
        ++p;
    }
    __syncthreads();
    const int npiv = s_npiv;

    // 4. fuse every cluster: one warp per pivot. The lanes scan the sorted positions behind the pivot 32 at a time
    //    (ballot of cid == pivot), lane 0 visits the members in sweep order - sequential sums, the reference's order
    const int lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    for (int c = threadIdx.x >> 5; c < npiv; c += nwarps) {
        const int pv = a.pivots[c];
        float* fp = a.f_prob + (size_t)c * k1;
        int members = 0;
        for (int q0 = pv + 1; q0 < n; q0 += 32) {
            const int q = q0 + lane;
            members += __popc(__ballot_sync(0xffffffffu, q < n && a.cid[q] == pv));
        }
        const int64_t cls = a.labels[src[pv]];
        if (members == 0) {
            if (lane == 0) {
                a.f_box[c] = a.boxes[src[pv]];
                a.f_score[c] = sscore[pv];
                a.f_cls[c] = cls;
            }
            for (int k = lane; k < k1; k += 32) fp[k] = a.probs[(size_t)src[pv] * k1 + k];
            continue;
        }
        const float count = (float)(members + 1);
        // member iteration order: matched (ascending sorted position), then the pivot
        float ssum = 0.0f, best = -INFINITY;
        int best_q = pv;
        bool mixed = false, bad_argmax = false;
        // every lane runs the (identical) sequential fusion below on the same data; only lane 0 stores
        float* acc = fp;
        if (lane == 0) for (int k = 0; k < k1; ++k) acc[k] = 0.0f;
        __syncwarp();
        auto for_members = [&](auto&& fn) {     // ascending sorted position, then the pivot
            for (int q0 = pv + 1; q0 < n; q0 += 32) {
                const int q = q0 + lane;
                uint32_t bits = __ballot_sync(0xffffffffu, q < n && a.cid[q] == pv);
                while (bits) {
                    fn(q0 + __ffs((int)bits) - 1);
                    bits &= bits - 1;
                }
            }
            fn(pv);
        };
        auto visit = [&](int q) {
            const int o = src[q];
            const float sc = sscore[q];
            ssum += sc;
            if (sc > best) { best = sc; best_q = q; }
            mixed |= (a.labels[o] != cls);
            if (a.score_method == COIN_SCORE_PROBEN) {
                int am = 0;
                float av = a.probs[(size_t)o * k1];
                for (int k = 0; k < k1; ++k) {
                    const float pr = a.probs[(size_t)o * k1 + k];
                    if (pr > av) { av = pr; am = k; }
                    if (lane == 0) acc[k] += logf(pr);
                }
                bad_argmax |= (am != (int)a.labels[o]);
            } else if (a.score_method == COIN_SCORE_AVG) {
                if (lane == 0) for (int k = 0; k < k1; ++k) acc[k] += a.probs[(size_t)o * k1 + k];
            }
        };
        for_members(visit);
        if (lane == 0 && mixed) atomicOr(a.status, 1);
        if (lane == 0 && bad_argmax) atomicOr(a.status, 2);
        __syncwarp();

        float fscore = 0.0f;
        if (lane == 0) {
            if (a.score_method == COIN_SCORE_PROBEN) {
                float esum = 0.0f;
                for (int k = 0; k < k1; ++k) { fp[k] = expf(fp[k]); esum += fp[k]; }
                for (int k = 0; k < k1; ++k) fp[k] = fp[k] / esum;
                fscore = fp[(int)cls];
            } else if (a.score_method == COIN_SCORE_AVG) {
                for (int k = 0; k < k1; ++k) fp[k] = fp[k] / count;
                fscore = ssum / count;
            } else {
                const int o = src[best_q];
                for (int k = 0; k < k1; ++k) fp[k] = a.probs[(size_t)o * k1 + k];
                fscore = best;
            }
        }

        float4 fb = make_float4(0.f, 0.f, 0.f, 0.f);
        if (a.box_method == COIN_BOX_MAX) {
            fb = a.boxes[src[best_q]];
        } else {
            auto add = [&](int q) {
                const float4 b = a.boxes[src[q]];
                if (a.box_method == COIN_BOX_SAVG) {
                    const float wgt = sscore[q] / ssum;
                    fb.x += b.x * wgt; fb.y += b.y * wgt; fb.z += b.z * wgt; fb.w += b.w * wgt;
                } else {
                    fb.x += b.x; fb.y += b.y; fb.z += b.z; fb.w += b.w;
                }
            };
            for_members(add);
            if (a.box_method == COIN_BOX_AVG) { fb.x /= count; fb.y /= count; fb.z /= count; fb.w /= count; }
        }
        if (lane == 0) {
            a.f_box[c] = fb;
            a.f_score[c] = fscore;
            a.f_cls[c] = cls;
        }
    }
    __syncthreads();

    // 5. re-sort the clusters by fused score (descending; ties keep sweep order) and emit
    int ppow = 1;
    while (ppow < npiv) ppow <<= 1;
    for (int i = threadIdx.x; i < ppow; i += blockDim.x)
        keys[i] = i < npiv ? desc_key(a.f_score[i], (uint32_t)i) : ~0ull;
    __syncthreads();
    bitonic_sort_smem(keys, ppow);
    for (int i = threadIdx.x; i < npiv; i += blockDim.x) {
        const int c = (int)(keys[i] & 0xffffffffu);
        a.keep[i] = src[a.pivots[c]];
        a.out_boxes[i] = a.f_box[c];
        a.out_scores[i] = a.f_score[c];
        a.out_classes[i] = a.f_cls[c];
        for (int k = 0; k < k1; ++k) a.out_probs[(size_t)i * k1 + k] = a.f_prob[(size_t)c * k1 + k];
    }
    if (threadIdx.x == 0) *a.nkeep = npiv;
}

static size_t fusion_smem_bytes(int n) {
    int npow = 2;
    while (npow < n) npow <<= 1;
    return (size_t)npow * 8 + (size_t)n * (16 + 4 + 4 + 4 + 4);
}

}  // namespace coin
using namespace coin;

static void carve_fusion(FusionArgs& a, void* ws, int64_t n, int k1, size_t* used) {
    Carver c(ws);
    a.cid = c.take<int32_t>((size_t)n);
    a.pivots = c.take<int32_t>((size_t)n);
    a.f_box = c.take<float4>((size_t)n);
    a.f_score = c.take<float>((size_t)n);
    a.f_prob = c.take<float>((size_t)n * k1);
    a.f_cls = c.take<int64_t>((size_t)n);
    a.g_keys = nullptr; a.g_nbox = nullptr; a.g_area = a.g_sscore = nullptr; a.g_src = a.g_alive = nullptr;
    if (n > kFusionMax || option("COIN_FUSION_FORCE_GLOBAL", 0)) {      // large-n variant: the working set moves to the workspace
        size_t npow = 2;
        while ((int64_t)npow < n) npow <<= 1;
        a.g_keys = c.take<uint64_t>(npow);
        a.g_nbox = c.take<float4>((size_t)n);
        a.g_area = c.take<float>((size_t)n);
        a.g_sscore = c.take<float>((size_t)n);
        a.g_src = c.take<int32_t>((size_t)n);
        a.g_alive = c.take<int32_t>((size_t)n);
    }
    *used = c.used();
}

extern "C" size_t coin_fusion_nms_workspace_bytes(int64_t n, int k1) {
    FusionArgs a;
    size_t used = 0;
    carve_fusion(a, nullptr, n, k1, &used);
    return used + 256;
}

extern "C" int coin_fusion_nms(const float* boxes, const float* probs, const int64_t* labels, int64_t n, int k1,
                               float iou_threshold, int score_method, int box_method, int per_class_offset,
                               int64_t* keep, float* out_boxes, float* out_scores, float* out_probs,
                               int64_t* out_classes, int32_t* nkeep, int32_t* status, void* ws, size_t ws_bytes,
                               coin_stream_t stream) {
    COIN_REQUIRE(n >= 0 && k1 >= 1 && nkeep && status, "fusion_nms: bad arguments");
    COIN_REQUIRE(score_method >= 0 && score_method <= 2 && box_method >= 0 && box_method <= 2, "fusion_nms: bad method");
    cudaStream_t s = as_stream(stream);
    fill_bytes(status, 0, sizeof(int32_t), s);
    if (n == 0) {
        fill_bytes(nkeep, 0, sizeof(int32_t), s);
        return COIN_OK;
    }
    COIN_REQUIRE(n < (1ll << 24), "fusion_nms: n=%lld exceeds the supported 16M boxes", (long long)n);
    COIN_REQUIRE(boxes && probs && labels && keep && out_boxes && out_scores && out_probs && out_classes && ws,
                 "fusion_nms: null pointer");
    COIN_REQUIRE((reinterpret_cast<uintptr_t>(boxes) & 15) == 0 && (reinterpret_cast<uintptr_t>(out_boxes) & 15) == 0,
                 "fusion_nms: boxes must be 16-byte aligned");
    if (ws_bytes < coin_fusion_nms_workspace_bytes(n, k1))
        return fail(COIN_ERR_CAPACITY, "fusion_nms: workspace too small");
    FusionArgs a;
    size_t used = 0;
    carve_fusion(a, ws, n, k1, &used);
    a.boxes = reinterpret_cast<const float4*>(boxes);
    a.probs = probs; a.labels = labels; a.n = (int)n; a.k1 = k1; a.thr = iou_threshold;
    a.score_method = score_method; a.box_method = box_method; a.per_class_offset = per_class_offset;
    a.keep = keep; a.out_boxes = reinterpret_cast<float4*>(out_boxes); a.out_scores = out_scores;
    a.out_probs = out_probs; a.out_classes = out_classes; a.nkeep = nkeep; a.status = status;
    if (a.g_keys) {
        fusion_nms_kernel<true><<<1, 1024, 0, s>>>(a);
        return check_launch("fusion_nms_kernel");
    }
    const size_t smem = fusion_smem_bytes((int)n);
    if (smem > 48 * 1024) cudaFuncSetAttribute(fusion_nms_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const int threads = n <= 128 ? 128 : (n <= 512 ? 256 : 1024);
    fusion_nms_kernel<false><<<1, threads, smem, s>>>(a);
    return check_launch("fusion_nms_kernel");
}
