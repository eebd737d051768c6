// sampling_loss.cu -- the sampling and loss-side reductions that sit between the matcher and ROIAlign / behind the heads
// (SURVEY.md 8(f) rank 2):
//
//   coin_proposal_classes     detectron2 ROIHeads._sample_proposals, first half (<- clip_roi_heads.py:317,363): the class of
//                             every proposal from (matched_idxs, matched_labels, gt_classes): background where the label is
//                             0, ignore (-1) where it is -1.
//   coin_subsample_labels     detectron2 modeling/sampling.py::subsample_labels (<- ROIHeads._sample_proposals,
//                             RPN._subsample_labels <- clip_roi_heads.py:363, rpn.py:231): a random subset of at most
//                             int(num_samples * positive_fraction) positives and the rest negatives. The reference draws two
//                             torch.randperm()s on the host generator; here the draw is either REPLAYED from caller-supplied
//                             permutations (bit-exact against the reference's seeded outputs) or made on the device from a
//                             counter-based generator (Philox4x32-10, key = seed, counter = (element, set, offset lo, offset hi)): every
//                             element gets a 32-bit key and the `num` smallest (key, index) pairs of a set are its sample, in
//                             key order - a uniform random subset in uniform random order, a function of (seed, offset) only.
//                             One CTA: ordered compaction of the two sets by a block scan, radix-select of the key of rank
//                             `num`, ordered collection, bitonic sort of the <= 4096 selected pairs.
//   coin_rpn_teacher_probs    rpn.py:95-98: teacher_probs = gt_probs[:, :-1].sum(1)[all_matched_idxs] (zeros without C boxes).
//   coin_kl_distill_roi_*     fast_rcnn.py:541-545: KLDiv(log(softmax(scores_c) + 1e-7), gt_probs), reduction 'mean', forward
//   coin_kl_distill_rpn_*     and gradient; rpn.py:326-340: the two-column (p, 1-p) KL over the anchors with a positive
//                             distillation label. Single pass each: per-block fp32 partial sums accumulated in double, the
//                             last block to finish writes the mean.
#include "common.cuh"

namespace coin {

constexpr int kSampleThreads = 1024;
constexpr int kSampleMax = 4096;       // num_samples per call (shared-memory sort of the selected pairs)

__device__ __forceinline__ uint32_t philox_word(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    return c0;
}

template <typename L>
__device__ __forceinline__ int64_t load_label(const void* p, int i) { return (int64_t) reinterpret_cast<const L*>(p)[i]; }

__global__ void proposal_classes_kernel(const int64_t* __restrict__ matched_idxs, const int8_t* __restrict__ matched_labels,
                                        const int64_t* __restrict__ gt_classes, int64_t n_gt_cap, const int32_t* __restrict__ n_gt_dev,
                                        int64_t m_cap, const int32_t* __restrict__ m_dev, int64_t num_classes,
                                        int64_t* __restrict__ out) {
    const int64_t m = m_dev ? min((int64_t)max(*m_dev, 0), m_cap) : m_cap;
    const int64_t n_gt = n_gt_dev ? min((int64_t)max(*n_gt_dev, 0), n_gt_cap) : n_gt_cap;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < m_cap; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t c = -1;                           // rows beyond the live count: ignored by the sampler
        if (i < m) {
            if (n_gt > 0) {
                const int8_t l = matched_labels[i];
                c = l == 0 ? num_classes : (l == -1 ? -1 : gt_classes[matched_idxs[i]]);
            } else {
                c = num_classes;
            }
        }
        out[i] = c;
    }
}

struct SubsampleArgs {
    const void* labels;
    int label_i8;
    int m_cap;
    const int32_t* m_dev;
    int num_samples, num_pos_target;
    int64_t bg_label;
    const int64_t* perm_pos;
    const int64_t* perm_neg;
    uint32_t seed_lo, seed_hi, off_lo, off_hi;
    int count_only;
    int64_t* pos_idx;
    int64_t* neg_idx;
    int32_t* counts;     // [n_pos, n_neg, P, N, status]
    int32_t* pos_list;   // [m_cap] workspace
    int32_t* neg_list;   // [m_cap] workspace
};

// the `num` smallest (key, element) pairs of list[0, len) in key order -> out[0, num)      (whole CTA; num <= kSampleMax)
__device__ void philox_select(const int32_t* __restrict__ list, int len, int num, uint32_t set, const SubsampleArgs& a,
                              int64_t* __restrict__ out, uint64_t* skeys, uint32_t* hist, uint32_t* s_sel, int* s_warp) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (num <= 0) return;
    auto key_of = [&](int e) { return philox_word((uint32_t)e, set, a.off_lo, a.off_hi, a.seed_lo, a.seed_hi); };
    // radix select: the key of rank num-1 among the list's keys
    if (tid == 0) { s_sel[0] = 0; s_sel[1] = (uint32_t)(num - 1); }
    for (int pass = 3; pass >= 0; --pass) {
        const int shift = 8 * pass;
        if (tid < 256) hist[tid] = 0;
        __syncthreads();
        const uint32_t prefix = s_sel[0];
        for (int i = tid; i < len; i += kSampleThreads) {
            const uint32_t k = key_of(list[i]);
            if (pass == 3 || (k >> (shift + 8)) == (prefix >> (shift + 8))) atomicAdd(&hist[(k >> shift) & 255u], 1u);
        }
        __syncthreads();
        if (tid == 0) {
            uint32_t rank = s_sel[1], b = 0;
            while (b < 255 && hist[b] <= rank) { rank -= hist[b]; ++b; }
            s_sel[1] = rank;
            s_sel[0] = prefix | (b << shift);
        }
        __syncthreads();
    }
    const uint32_t thr = s_sel[0];
    const int ties_wanted = (int)s_sel[1] + 1;          // elements with key == thr to take, in list order
    // ordered collection: thread t owns the contiguous slice [t * per, (t + 1) * per) of the list
    const int per = (len + kSampleThreads - 1) / kSampleThreads;
    const int i0 = min(tid * per, len), i1 = min(i0 + per, len);
    int n_less = 0, n_tie = 0;
    for (int i = i0; i < i1; ++i) {
        const uint32_t k = key_of(list[i]);
        n_less += k < thr;
        n_tie += k == thr;
    }
    int p_less = n_less, p_tie = n_tie;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int vl = __shfl_up_sync(0xffffffffu, p_less, o), vt = __shfl_up_sync(0xffffffffu, p_tie, o);
        if (lane >= o) { p_less += vl; p_tie += vt; }
    }
    if (lane == 31) { s_warp[warp] = p_less; s_warp[32 + warp] = p_tie; }
    __syncthreads();
    int o_less = p_less - n_less, o_tie = p_tie - n_tie, total_less = 0;
    for (int w = 0; w < kSampleThreads / 32; ++w) {
        if (w < warp) { o_less += s_warp[w]; o_tie += s_warp[32 + w]; }
        total_less += s_warp[w];
    }
    for (int i = i0; i < i1; ++i) {
        const int e = list[i];
        const uint32_t k = key_of(e);
        if (k < thr) skeys[o_less++] = ((uint64_t)k << 32) | (uint32_t)e;
        else if (k == thr) { if (o_tie < ties_wanted) skeys[total_less + o_tie] = ((uint64_t)k << 32) | (uint32_t)e; ++o_tie; }
    }
    int npow = 2;
    while (npow < num) npow <<= 1;
    __syncthreads();
    for (int i = num + tid; i < npow; i += kSampleThreads) skeys[i] = ~0ull;
    __syncthreads();
    for (int kk = 2; kk <= npow; kk <<= 1)
        for (int j = kk >> 1; j > 0; j >>= 1) {
            for (int t = tid; t < (npow >> 1); t += kSampleThreads) {
                const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1)), p = i | j;
                const uint64_t x = skeys[i], y = skeys[p];
                if ((x > y) == ((i & kk) == 0)) { skeys[i] = y; skeys[p] = x; }
            }
            __syncthreads();
        }
    for (int i = tid; i < num; i += kSampleThreads) out[i] = (int64_t)(skeys[i] & 0xffffffffu);
    __syncthreads();
}

__global__ void __launch_bounds__(kSampleThreads) subsample_labels_kernel(const SubsampleArgs a) {
    __shared__ uint64_t skeys[kSampleMax];
    __shared__ uint32_t hist[256];
    __shared__ uint32_t s_sel[2];
    __shared__ int s_warp[64];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int m = a.m_dev ? min(max(*a.m_dev, 0), a.m_cap) : a.m_cap;
    // 1. ordered lists of the positive / negative elements (torch.nonzero order)
    const int per = (m + kSampleThreads - 1) / kSampleThreads;
    const int i0 = min(tid * per, m), i1 = min(i0 + per, m);
    auto label = [&](int i) { return a.label_i8 ? load_label<int8_t>(a.labels, i) : load_label<int64_t>(a.labels, i); };
    int np = 0, nn = 0;
    for (int i = i0; i < i1; ++i) {
        const int64_t l = label(i);
        np += (l != -1 && l != a.bg_label);
        nn += (l == a.bg_label);
    }
    int pp = np, pn = nn;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int vp = __shfl_up_sync(0xffffffffu, pp, o), vn = __shfl_up_sync(0xffffffffu, pn, o);
        if (lane >= o) { pp += vp; pn += vn; }
    }
    if (lane == 31) { s_warp[warp] = pp; s_warp[32 + warp] = pn; }
    __syncthreads();
    int op = pp - np, on = pn - nn, P = 0, N = 0;
    for (int w = 0; w < kSampleThreads / 32; ++w) {
        if (w < warp) { op += s_warp[w]; on += s_warp[32 + w]; }
        P += s_warp[w];
        N += s_warp[32 + w];
    }
    for (int i = i0; i < i1; ++i) {
        const int64_t l = label(i);
        if (l != -1 && l != a.bg_label) a.pos_list[op++] = i;
        else if (l == a.bg_label) a.neg_list[on++] = i;
    }
    const int num_pos = min(P, a.num_pos_target);
    const int num_neg = min(N, a.num_samples - num_pos);
    if (tid == 0) {
        a.counts[0] = a.count_only ? 0 : num_pos;
        a.counts[1] = a.count_only ? 0 : num_neg;
        a.counts[2] = P;
        a.counts[3] = N;
        a.counts[4] = 0;
    }
    __syncthreads();                 // (also makes the lists visible to the whole CTA)
    if (a.count_only) return;
    if (a.perm_pos || a.perm_neg) {  // replay of the caller's permutations: positive[perm1[:num_pos]], negative[perm2[:num_neg]]
        bool bad = false;
        for (int j = tid; j < num_pos; j += kSampleThreads) {
            const int64_t q = a.perm_pos[j];
            if (q < 0 || q >= P) { bad = true; a.pos_idx[j] = -1; } else a.pos_idx[j] = a.pos_list[q];
        }
        for (int j = tid; j < num_neg; j += kSampleThreads) {
            const int64_t q = a.perm_neg[j];
            if (q < 0 || q >= N) { bad = true; a.neg_idx[j] = -1; } else a.neg_idx[j] = a.neg_list[q];
        }
        if (bad) a.counts[4] = 1;
        return;
    }
    philox_select(a.pos_list, P, num_pos, 0u, a, a.pos_idx, skeys, hist, s_sel, s_warp);
    philox_select(a.neg_list, N, num_neg, 1u, a, a.neg_idx, skeys, hist, s_sel, s_warp);
}

__global__ void rpn_teacher_probs_kernel(const float* __restrict__ gt_probs, int64_t nc_cap, const int32_t* __restrict__ nc_dev,
                                         int k1, const int64_t* __restrict__ matched, int64_t n, float* __restrict__ out) {
    const int64_t nc = nc_dev ? min((int64_t)max(*nc_dev, 0), nc_cap) : nc_cap;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        float s = 0.0f;
        if (nc > 0) {
            const float* row = gt_probs + matched[i] * k1;
            for (int k = 0; k + 1 < k1; ++k) s += row[k];
        }
        out[i] = s;
    }
}

// ---- reductions: block partial (fp32) -> double accumulator -> the last block writes the mean -------------------
struct ReduceWs { double sum; unsigned long long count; unsigned int done; unsigned int pad; };

__device__ __forceinline__ float block_sum(float v, float* s_red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
    __syncthreads();
    float t = 0.0f;
    if (threadIdx.x < 32) {
        t = threadIdx.x < (blockDim.x >> 5) ? s_red[threadIdx.x] : 0.0f;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    }
    __syncthreads();
    return t;   // valid in thread 0
}

__device__ __forceinline__ float xlogy(float x, float y) { return x == 0.0f ? 0.0f : x * logf(y); }

// one thread per row (k1 is ~9): softmax in registers
__global__ void __launch_bounds__(256)
kl_roi_fwd_kernel(const float* __restrict__ scores, const float* __restrict__ q, int64_t n_cap, const int32_t* __restrict__ n_dev,
                  int k1, ReduceWs* __restrict__ ws, float* __restrict__ loss) {
    __shared__ float s_red[32];
    const int64_t n = n_dev ? min((int64_t)max(*n_dev, 0), n_cap) : n_cap;
    float acc = 0.0f;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float* s = scores + i * k1;
        const float* t = q + i * k1;
        float mx = -INFINITY;
        for (int k = 0; k < k1; ++k) mx = fmaxf(mx, s[k]);
        float den = 0.0f;
        for (int k = 0; k < k1; ++k) den += expf(s[k] - mx);
        for (int k = 0; k < k1; ++k) {
            const float p = expf(s[k] - mx) / den;
            acc += xlogy(t[k], t[k]) - t[k] * logf(p + 1e-7f);
        }
    }
    const float part = block_sum(acc, s_red);
    if (threadIdx.x == 0) {
        atomicAdd(&ws->sum, (double)part);
        __threadfence();
        if (atomicAdd(&ws->done, 1u) == gridDim.x - 1) {
            __threadfence();
            const double total = *reinterpret_cast<volatile double*>(&ws->sum);
            *loss = n > 0 ? (float)(total / (double)(n * k1)) : 0.0f;
        }
    }
}

__global__ void __launch_bounds__(256)
kl_roi_bwd_kernel(const float* __restrict__ scores, const float* __restrict__ q, int64_t n_cap, const int32_t* __restrict__ n_dev,
                  int k1, const float* __restrict__ grad_loss, float* __restrict__ grad_scores) {
    const int64_t n = n_dev ? min((int64_t)max(*n_dev, 0), n_cap) : n_cap;
    const float go = *grad_loss;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n_cap; i += (int64_t)gridDim.x * blockDim.x) {
        float* g = grad_scores + i * k1;
        if (i >= n) { for (int k = 0; k < k1; ++k) g[k] = 0.0f; continue; }
        const float* s = scores + i * k1;
        const float* t = q + i * k1;
        float mx = -INFINITY;
        for (int k = 0; k < k1; ++k) mx = fmaxf(mx, s[k]);
        float den = 0.0f;
        for (int k = 0; k < k1; ++k) den += expf(s[k] - mx);
        const float scale = go / (float)(n * k1);
        float dot = 0.0f;                       // sum_j (dL/dp_j) p_j
        for (int k = 0; k < k1; ++k) {
            const float p = expf(s[k] - mx) / den;
            dot += -t[k] / (p + 1e-7f) * p;
        }
        for (int k = 0; k < k1; ++k) {
            const float p = expf(s[k] - mx) / den;
            g[k] = scale * p * (-t[k] / (p + 1e-7f) - dot);
        }
    }
}

__global__ void __launch_bounds__(256)
kl_rpn_fwd_kernel(const float* __restrict__ logits, const int8_t* __restrict__ labels, const float* __restrict__ teacher,
                  int64_t n, ReduceWs* __restrict__ ws, float* __restrict__ loss, int32_t* __restrict__ n_valid) {
    __shared__ float s_red[32];
    float acc = 0.0f, cnt = 0.0f;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        if (labels[i] <= 0) continue;
        const float p = 1.0f / (1.0f + expf(-logits[i]));
        const float t = teacher[i];
        const float p1 = 1.0f - p, t1 = 1.0f - t;
        acc += xlogy(t, t) - t * logf(p + 1e-7f) + xlogy(t1, t1) - t1 * logf(p1 + 1e-7f);
        cnt += 1.0f;
    }
    const float part = block_sum(acc, s_red);
    const float pc = block_sum(cnt, s_red);
    if (threadIdx.x == 0) {
        atomicAdd(&ws->sum, (double)part);
        atomicAdd(&ws->count, (unsigned long long)(pc + 0.5f));
        __threadfence();
        if (atomicAdd(&ws->done, 1u) == gridDim.x - 1) {
            __threadfence();
            const double total = *reinterpret_cast<volatile double*>(&ws->sum);
            const unsigned long long nv = *reinterpret_cast<volatile unsigned long long*>(&ws->count);
            *loss = nv ? (float)(total / (2.0 * (double)nv)) : 0.0f;
            *n_valid = (int32_t)nv;
        }
    }
}

__global__ void __launch_bounds__(256)
kl_rpn_bwd_kernel(const float* __restrict__ logits, const int8_t* __restrict__ labels, const float* __restrict__ teacher,
                  int64_t n, const int32_t* __restrict__ n_valid, const float* __restrict__ grad_loss,
                  float* __restrict__ grad_logits) {
    const float nv = (float)max(*n_valid, 1);
    const float scale = *grad_loss / (2.0f * nv);
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        float g = 0.0f;
        if (labels[i] > 0) {
            const float p = 1.0f / (1.0f + expf(-logits[i]));
            const float t = teacher[i];
            const float p1 = 1.0f - p, t1 = 1.0f - t;
            g = scale * (-t / (p + 1e-7f) + t1 / (p1 + 1e-7f)) * p * p1;
        }
        grad_logits[i] = g;
    }
}

}  // namespace coin
using namespace coin;

extern "C" int coin_proposal_classes(const int64_t* matched_idxs, const int8_t* matched_labels, const int64_t* gt_classes,
                                     int64_t n_gt_cap, const int32_t* n_gt_dev, int64_t m_cap, const int32_t* m_dev,
                                     int64_t num_classes, int64_t* out_classes, coin_stream_t stream) {
    COIN_REQUIRE(m_cap >= 0 && n_gt_cap >= 0, "proposal_classes: bad sizes");
    if (m_cap == 0) return COIN_OK;
    COIN_REQUIRE(matched_idxs && matched_labels && out_classes && (n_gt_cap == 0 || gt_classes), "proposal_classes: null pointer");
    const unsigned blocks = (unsigned)std::min<int64_t>(ceil_div(m_cap, 256), 4 * kNumSMs);
    proposal_classes_kernel<<<blocks, 256, 0, as_stream(stream)>>>(matched_idxs, matched_labels, gt_classes, n_gt_cap, n_gt_dev,
                                                                 m_cap, m_dev, num_classes, out_classes);
    return check_launch("proposal_classes_kernel");
}

extern "C" size_t coin_subsample_labels_workspace_bytes(int64_t m_cap) { return (size_t)std::max<int64_t>(m_cap, 1) * 8 + 512; }

extern "C" int coin_subsample_labels(const void* labels, int label_is_int8, int64_t m_cap, const int32_t* m_dev,
                                     int num_samples, int num_pos_target, int64_t bg_label, const int64_t* perm_pos,
                                     const int64_t* perm_neg, uint64_t seed, uint64_t offset, int count_only,
                                     int64_t* pos_idx, int64_t* neg_idx, int32_t* counts, void* ws, size_t ws_bytes,
                                     coin_stream_t stream) {
    COIN_REQUIRE(m_cap >= 0 && m_cap < (1ll << 31) && counts, "subsample_labels: bad arguments");
    COIN_REQUIRE(num_samples >= 0 && num_samples <= kSampleMax && num_pos_target >= 0 && num_pos_target <= num_samples,
                 "subsample_labels: num_samples=%d (pos %d) out of [0,%d]", num_samples, num_pos_target, kSampleMax);
    COIN_REQUIRE((perm_pos == nullptr) == (perm_neg == nullptr), "subsample_labels: pass both permutations or neither");
    COIN_REQUIRE(m_cap == 0 || (labels && ws), "subsample_labels: null pointer");
    COIN_REQUIRE(count_only || num_samples == 0 || (pos_idx && neg_idx), "subsample_labels: null output");
    if (ws_bytes < coin_subsample_labels_workspace_bytes(m_cap)) return fail(COIN_ERR_CAPACITY, "subsample_labels: workspace too small");
    SubsampleArgs a;
    a.labels = labels; a.label_i8 = label_is_int8; a.m_cap = (int)m_cap; a.m_dev = m_dev;
    a.num_samples = num_samples; a.num_pos_target = num_pos_target; a.bg_label = bg_label;
    a.perm_pos = perm_pos; a.perm_neg = perm_neg;
    a.seed_lo = (uint32_t)seed; a.seed_hi = (uint32_t)(seed >> 32); a.off_lo = (uint32_t)offset; a.off_hi = (uint32_t)(offset >> 32);
    a.count_only = count_only; a.pos_idx = pos_idx; a.neg_idx = neg_idx; a.counts = counts;
    Carver c(ws);
    a.pos_list = c.take<int32_t>((size_t)std::max<int64_t>(m_cap, 1));
    a.neg_list = c.take<int32_t>((size_t)std::max<int64_t>(m_cap, 1));
    subsample_labels_kernel<<<1, kSampleThreads, 0, as_stream(stream)>>>(a);
    return check_launch("subsample_labels_kernel");
}

extern "C" int coin_rpn_teacher_probs(const float* gt_probs, int64_t nc_cap, const int32_t* nc_dev, int k1,
                                      const int64_t* matched, int64_t n, float* out, coin_stream_t stream) {
    COIN_REQUIRE(n >= 0 && nc_cap >= 0 && k1 >= 1, "rpn_teacher_probs: bad arguments");
    if (n == 0) return COIN_OK;
    COIN_REQUIRE(out && matched && (nc_cap == 0 || gt_probs), "rpn_teacher_probs: null pointer");
    const unsigned blocks = (unsigned)std::min<int64_t>(ceil_div(n, 256), 4 * kNumSMs);
    rpn_teacher_probs_kernel<<<blocks, 256, 0, as_stream(stream)>>>(gt_probs, nc_cap, nc_dev, k1, matched, n, out);
    return check_launch("rpn_teacher_probs_kernel");
}

extern "C" size_t coin_kl_workspace_bytes(void) { return 256; }

extern "C" int coin_kl_distill_roi_fwd(const float* scores, const float* gt_probs, int64_t n_cap, const int32_t* n_dev, int k1,
                                       float* loss, void* ws, coin_stream_t stream) {
    COIN_REQUIRE(n_cap >= 0 && k1 >= 1 && k1 <= 1024 && loss && ws, "kl_distill_roi: bad arguments");
    COIN_REQUIRE(n_cap == 0 || (scores && gt_probs), "kl_distill_roi: null pointer");
    cudaStream_t s = as_stream(stream);
    fill_bytes(ws, 0, sizeof(ReduceWs), s);
    const unsigned blocks = (unsigned)std::max<int64_t>(1, std::min<int64_t>(ceil_div(n_cap, 256), kNumSMs));
    kl_roi_fwd_kernel<<<blocks, 256, 0, s>>>(scores, gt_probs, n_cap, n_dev, k1, static_cast<ReduceWs*>(ws), loss);
    return check_launch("kl_roi_fwd_kernel");
}

extern "C" int coin_kl_distill_roi_bwd(const float* scores, const float* gt_probs, int64_t n_cap, const int32_t* n_dev, int k1,
                                       const float* grad_loss, float* grad_scores, coin_stream_t stream) {
    COIN_REQUIRE(n_cap >= 0 && k1 >= 1 && grad_loss, "kl_distill_roi_bwd: bad arguments");
    if (n_cap == 0) return COIN_OK;
    COIN_REQUIRE(scores && gt_probs && grad_scores, "kl_distill_roi_bwd: null pointer");
    const unsigned blocks = (unsigned)std::min<int64_t>(ceil_div(n_cap, 256), 4 * kNumSMs);
    kl_roi_bwd_kernel<<<blocks, 256, 0, as_stream(stream)>>>(scores, gt_probs, n_cap, n_dev, k1, grad_loss, grad_scores);
    return check_launch("kl_roi_bwd_kernel");
}

extern "C" int coin_kl_distill_rpn_fwd(const float* logits, const int8_t* distillation_labels, const float* teacher_probs,
                                       int64_t n, float* loss, int32_t* n_valid, void* ws, coin_stream_t stream) {
    COIN_REQUIRE(n >= 0 && loss && n_valid && ws, "kl_distill_rpn: bad arguments");
    COIN_REQUIRE(n == 0 || (logits && distillation_labels && teacher_probs), "kl_distill_rpn: null pointer");
    cudaStream_t s = as_stream(stream);
    fill_bytes(ws, 0, sizeof(ReduceWs), s);
    const unsigned blocks = (unsigned)std::max<int64_t>(1, std::min<int64_t>(ceil_div(n, 1024), 2 * kNumSMs));
    kl_rpn_fwd_kernel<<<blocks, 256, 0, s>>>(logits, distillation_labels, teacher_probs, n, static_cast<ReduceWs*>(ws), loss, n_valid);
    return check_launch("kl_rpn_fwd_kernel");
}

extern "C" int coin_kl_distill_rpn_bwd(const float* logits, const int8_t* distillation_labels, const float* teacher_probs,
                                       int64_t n, const int32_t* n_valid, const float* grad_loss, float* grad_logits,
                                       coin_stream_t stream) {
    COIN_REQUIRE(n >= 0 && n_valid && grad_loss, "kl_distill_rpn_bwd: bad arguments");
    if (n == 0) return COIN_OK;
    COIN_REQUIRE(logits && distillation_labels && teacher_probs && grad_logits, "kl_distill_rpn_bwd: null pointer");
    const unsigned blocks = (unsigned)std::min<int64_t>(ceil_div(n, 1024), 4 * kNumSMs);
    kl_rpn_bwd_kernel<<<blocks, 256, 0, as_stream(stream)>>>(logits, distillation_labels, teacher_probs, n, n_valid, grad_loss,
                                                           grad_logits);
    return check_launch("kl_rpn_bwd_kernel");
}
