// iou_match.cu -- pairwise IoU, Matcher, the fused IoU+Matcher, relabelling epilogues and the
// thresholded pair list used by the knowledge-separation step.
//
// Replaces detectron2 pairwise_iou (~10 ATen launches materialising [N,M,2] temporaries) and
// Matcher (max(dim=0) + masked fills + max(dim=1)/==/nonzero for low-quality matches) as reached from
//   coin/modeling/roi_heads/clip_roi_heads.py:301-304,311-314,353-356   (GT x ~2000 proposals)
//   coin/modeling/proposal_generator/rpn.py:159-160,169-170,212-213     (GT x 41 625 anchors)
//   coin/engine/trainer.py:364-366,373, coin/utils/util.py:468          (cloud x CLIP-detector dets)
// Columns (proposals / anchors) are the coalesced dimension; the GT rows are staged in shared
// memory in tiles. Row maxima for the low-quality rule use warp-shuffle reductions + one
// atomicMax per warp and row. All float compares follow the oracle bit for bit (-fmad=false,
// IEEE division, thresholds rounded to fp32 exactly as torch does for tensor-vs-scalar compares).
#include "common.cuh"

namespace coin {

constexpr int kRowTile = 256;  // GT rows staged per shared-memory tile
constexpr int kMaxThr = 8;

struct MatcherCfg {
    float thr[kMaxThr + 2];     // [-inf, t0, ..., +inf]
    int8_t labels[kMaxThr + 1];
    int nbuckets;               // nthr + 1
};

__device__ __forceinline__ int8_t bucket_label(const MatcherCfg& m, float v) {
    int8_t lab = 1;  // detectron2 initialises labels to 1 and overwrites per bucket
#pragma unroll 1
    for (int b = 0; b < m.nbuckets; ++b)
        if (v >= m.thr[b] && v < m.thr[b + 1]) lab = m.labels[b];
    return lab;
}

// ------------------------------------------------------------------------------------------------
__global__ void pairwise_iou_kernel(const float4* __restrict__ b1, int64_t N, const float4* __restrict__ b2,
                                    int64_t M, float* __restrict__ out, int rows_per_block) {
    extern __shared__ float4 srow[];  // rows_per_block boxes
    const int64_t r0 = (int64_t)blockIdx.y * rows_per_block;
    const int nr = (int)min((int64_t)rows_per_block, N - r0);
    for (int t = threadIdx.x; t < nr; t += blockDim.x) srow[t] = __ldg(b1 + r0 + t);
    __syncthreads();
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= M) return;
    const float4 b = __ldg(b2 + j);
    const float area_b = box_area(b);
    for (int t = 0; t < nr; ++t) {
        const float4 a = srow[t];
        out[(r0 + t) * M + j] = iou_d2(a, box_area(a), b, area_b);
    }
}

// Fused IoU + column arg-max (+ optional row maxima). A CTA of 256 threads owns CW = 256 / G consecutive columns;
// G thread groups split the GT rows (row t of a tile goes to group t % G) so that a few thousand columns still fill
// the machine and a thread's serial chain is N / G IoUs. n_dev / m_dev: optional device-side live counts (<= the host
// capacities N / M) for sync-free chaining.
// Row maxima (low-quality rule): one REDUX per warp and row on the IoU bit pattern (IoU >= 0, so the unsigned order is the
// float order), one shared-memory atomicMax per warp and row, one global atomicMax per CTA and row - and the CTA's own
// maximum of every row is kept in cta_rmax[cta][row]: the second kernel then re-evaluates a row only in the CTAs that
// reached the global maximum (about one CTA per row) instead of re-evaluating every pair.
template <bool kRowMax, int G>
__global__ void __launch_bounds__(256)
iou_match_kernel(const float4* __restrict__ gt, int64_t N, const float4* __restrict__ boxes, int64_t M, MatcherCfg cfg,
                 int64_t* __restrict__ matches, int8_t* __restrict__ labels, float* __restrict__ vals,
                 unsigned* __restrict__ row_max, unsigned* __restrict__ cta_rmax, const int32_t* __restrict__ n_dev,
                 const int32_t* __restrict__ m_dev) {
    constexpr int CW = 256 / G;
    __shared__ float4 srow[kRowTile];
    __shared__ float sarea[kRowTile];
    __shared__ unsigned s_rmax[kRowMax ? kRowTile : 1];
    __shared__ float s_best[G > 1 ? 256 : 1];
    __shared__ int s_besti[G > 1 ? 256 : 1];
    const int64_t n_cap = N;
    if (n_dev) N = min((int64_t)max(__ldg(n_dev), 0), N);
    if (m_dev) M = min((int64_t)max(__ldg(m_dev), 0), M);
    if ((int64_t)blockIdx.x * CW >= M) return;   // capacity launch: whole CTA beyond the live columns
    const int c = threadIdx.x % CW, g = threadIdx.x / CW;
    const int64_t j = (int64_t)blockIdx.x * CW + c;
    const bool live = j < M;
    const float4 b = live ? __ldg(boxes + j) : make_float4(0.f, 0.f, 0.f, 0.f);
    const float area_b = box_area(b);
    float best = -1.0f;  // IoU >= 0, so the first row always wins the first comparison (torch: first max)
    int best_i = 0;
    for (int64_t r0 = 0; r0 < N; r0 += kRowTile) {
        const int nr = (int)min((int64_t)kRowTile, N - r0);
        __syncthreads();
        for (int t = threadIdx.x; t < nr; t += blockDim.x) {
            const float4 a = __ldg(gt + r0 + t);
            srow[t] = a;
            sarea[t] = box_area(a);
            if (kRowMax) s_rmax[t] = 0u;
        }
        __syncthreads();
        for (int t = g; t < nr; t += G) {
            const float v = live ? iou_d2(srow[t], sarea[t], b, area_b) : 0.0f;
            if (v > best) { best = v; best_i = (int)r0 + t; }
            if (kRowMax) {
                const unsigned m = __reduce_max_sync(0xffffffffu, __float_as_uint(v));
                if ((threadIdx.x & 31) == 0 && m != 0u) atomicMax(&s_rmax[t], m);
            }
        }
        if (kRowMax) {
            __syncthreads();
            for (int t = threadIdx.x; t < nr; t += blockDim.x) {
                const unsigned m = s_rmax[t];
                cta_rmax[(size_t)blockIdx.x * n_cap + r0 + t] = m;
                if (m != 0u) atomicMax(row_max + r0 + t, m);
            }
        }
    }
    if (G > 1) {   // combine the row groups: larger IoU wins, equal IoU -> lower row (torch.max: first maximum)
        __syncthreads();
        s_best[threadIdx.x] = best;
        s_besti[threadIdx.x] = best_i;
        __syncthreads();
        if (g != 0) return;
        for (int q = 1; q < G; ++q) {
            const float v = s_best[q * CW + c];
            const int i = s_besti[q * CW + c];
            if (v > best || (v == best && i < best_i)) { best = v; best_i = i; }
        }
    }
    if (!live) return;
    if (N == 0) {   // Matcher's empty-matrix rule (decided on the device when N is a device count)
        matches[j] = 0;
        labels[j] = cfg.labels[0];
        if (vals) vals[j] = 0.0f;
        return;
    }
    matches[j] = best_i;
    labels[j] = bucket_label(cfg, best);
    if (vals) vals[j] = best;
}

// Low-quality rule: every column whose IoU with some GT row equals that row's maximum gets label 1. Same column
// partition as iou_match_kernel; a CTA re-evaluates only the rows whose maximum it reached itself (cta_rmax == row_max,
// which includes the rows that overlap nothing: their maximum 0 is reached by every column, as in detectron2).
template <int G>
__global__ void __launch_bounds__(256)
iou_low_quality_kernel(const float4* __restrict__ gt, int64_t N, const float4* __restrict__ boxes, int64_t M,
                       const unsigned* __restrict__ row_max, const unsigned* __restrict__ cta_rmax,
                       int8_t* __restrict__ labels, const int32_t* __restrict__ n_dev, const int32_t* __restrict__ m_dev) {
    constexpr int CW = 256 / G;
    __shared__ int s_rows[kRowTile];
    __shared__ int s_n;
    __shared__ int s_hit[CW];
    const int64_t n_cap = N;
    if (n_dev) N = min((int64_t)max(__ldg(n_dev), 0), N);
    if (m_dev) M = min((int64_t)max(__ldg(m_dev), 0), M);
    if ((int64_t)blockIdx.x * CW >= M) return;
    const int c = threadIdx.x % CW, g = threadIdx.x / CW;
    const int64_t j = (int64_t)blockIdx.x * CW + c;
    const bool live = j < M;
    const float4 b = live ? __ldg(boxes + j) : make_float4(0.f, 0.f, 0.f, 0.f);
    const float area_b = box_area(b);
    if (threadIdx.x < CW) s_hit[threadIdx.x] = 0;
    bool hit = false;
    for (int64_t r0 = 0; r0 < N; r0 += kRowTile) {
        const int nr = (int)min((int64_t)kRowTile, N - r0);
        __syncthreads();
        if (threadIdx.x == 0) s_n = 0;
        __syncthreads();
        for (int t = threadIdx.x; t < nr; t += blockDim.x)
            if (cta_rmax[(size_t)blockIdx.x * n_cap + r0 + t] == __ldg(row_max + r0 + t)) s_rows[atomicAdd(&s_n, 1)] = (int)r0 + t;
        __syncthreads();
        const int n = s_n;
        for (int q = g; q < n; q += G) {
            const int r = s_rows[q];
            const float4 a = __ldg(gt + r);
            hit |= (__float_as_uint(iou_d2(a, box_area(a), b, area_b)) == __ldg(row_max + r));
        }
    }
    if (G > 1) {
        if (hit) s_hit[c] = 1;
        __syncthreads();
        hit = s_hit[c] != 0;
        if (g != 0) return;
    }
    if (live && hit) labels[j] = 1;
}

// Matcher on a materialised matrix: column arg-max, coalesced along M.
__global__ void matcher_kernel(const float* __restrict__ q, int64_t N, int64_t M, MatcherCfg cfg,
                               int64_t* __restrict__ matches, int8_t* __restrict__ labels, float* __restrict__ vals) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= M) return;
    float best = __ldg(q + j);
    int64_t best_i = 0;
    for (int64_t i = 1; i < N; ++i) {
        const float v = __ldg(q + i * M + j);
        if (v > best) { best = v; best_i = i; }
    }
    matches[j] = best_i;
    labels[j] = bucket_label(cfg, best);
    if (vals) vals[j] = best;
}

__global__ void row_max_kernel(const float* __restrict__ q, int64_t M, float* __restrict__ row_max) {
    __shared__ float part[32];
    const float* row = q + (int64_t)blockIdx.x * M;
    float m = -INFINITY;
    for (int64_t j = threadIdx.x; j < M; j += blockDim.x) m = fmaxf(m, __ldg(row + j));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x < 32) {
        m = threadIdx.x < (blockDim.x >> 5) ? part[threadIdx.x] : -INFINITY;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
        if (threadIdx.x == 0) row_max[blockIdx.x] = m;
    }
}

__global__ void matrix_low_quality_kernel(const float* __restrict__ q, int64_t N, int64_t M,
                                          const float* __restrict__ row_max, int8_t* __restrict__ labels) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= M) return;
    bool hit = false;
    for (int64_t i = 0; i < N; ++i) hit |= (__ldg(q + i * M + j) == __ldg(row_max + i));
    if (hit) labels[j] = 1;
}

__global__ void empty_matcher_kernel(int64_t M, int8_t label0, int64_t* __restrict__ matches,
                                     int8_t* __restrict__ labels, float* __restrict__ vals) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= M) return;
    matches[j] = 0;
    labels[j] = label0;
    if (vals) vals[j] = 0.0f;
}

// ------------------------------------------------------------------------------------------------
// lens: optional device int32 lengths {len_a, len_b, len_c} of the pseudo-GT sets (c_begin = a+b, c_end = a+b+c)
__global__ void relabel_roi_kernel(const int64_t* __restrict__ matches, int8_t* __restrict__ labels, int64_t M,
                                   int64_t c_begin, int64_t c_end, const int32_t* __restrict__ len_a,
                                   const int32_t* __restrict__ len_b, const int32_t* __restrict__ len_c,
                                   const int32_t* __restrict__ m_dev) {
    if (len_a) { c_begin = (int64_t)__ldg(len_a) + __ldg(len_b); c_end = c_begin + __ldg(len_c); }
    if (m_dev) M = min((int64_t)max(__ldg(m_dev), 0), M);
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= M) return;
    const int64_t m = matches[j];
    if (m >= c_begin && m < c_end && labels[j] != 0) labels[j] = -1;
}

__global__ void relabel_rpn_kernel(int64_t* __restrict__ matches, int8_t* __restrict__ labels, int64_t M,
                                   int64_t len_a, int64_t len_c, int64_t* __restrict__ distill_idx,
                                   int8_t* __restrict__ distill_labels, const int32_t* __restrict__ len_a_dev,
                                   const int32_t* __restrict__ len_c_dev) {
    if (len_a_dev) { len_a = __ldg(len_a_dev); len_c = __ldg(len_c_dev); }
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= M) return;
    const int64_t m = matches[j];
    const bool in_c = m >= len_a && m < len_a + len_c;
    const bool fg_c = in_c && labels[j] != 0;
    if (fg_c) labels[j] = -1;
    if (in_c) matches[j] = 0;
    distill_idx[j] = fg_c ? m - len_a : 0;
    distill_labels[j] = fg_c ? 1 : 0;
}

// ------------------------------------------------------------------------------------------------
// IoU >= thr pair list in row-major order: ballot words + per-row counts, scan, ordered write.
// ------------------------------------------------------------------------------------------------
__global__ void pairs_ballot_kernel(const float4* __restrict__ b1, int64_t N, const float4* __restrict__ b2,
                                    int64_t M, float thr, uint32_t* __restrict__ words, int words_per_row,
                                    int32_t* __restrict__ row_count) {
    const int64_t i = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (i >= N) return;
    const int lane = threadIdx.x & 31;
    const float4 a = __ldg(b1 + i);
    const float area_a = box_area(a);
    int cnt = 0;
    for (int w = 0; w < words_per_row; ++w) {
        const int64_t j = (int64_t)w * 32 + lane;
        bool p = false;
        if (j < M) {
            const float4 b = __ldg(b2 + j);
            p = iou_d2(a, area_a, b, box_area(b)) >= thr;
        }
        const uint32_t word = __ballot_sync(0xffffffffu, p);
        if (lane == 0) words[i * words_per_row + w] = word;
        cnt += __popc(word);
    }
    if (lane == 0) row_count[i] = cnt;
}

__global__ void pairs_scan_kernel(const int32_t* __restrict__ row_count, int64_t N, int32_t* __restrict__ row_off,
                                  int32_t* __restrict__ total) {
    // single block; N is small (<= a few thousand rows on this path): chunked warp-shuffle scan
    __shared__ int32_t warp_sum[32];
    __shared__ int32_t carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int64_t base = 0; base < N; base += blockDim.x) {
        const int64_t i = base + threadIdx.x;
        const int32_t v = i < N ? row_count[i] : 0;
        int32_t x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int32_t y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += y;
        }
        if (lane == 31) warp_sum[warp] = x;
        __syncthreads();
        if (warp == 0) {
            int32_t s = lane < (blockDim.x >> 5) ? warp_sum[lane] : 0;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int32_t y = __shfl_up_sync(0xffffffffu, s, o);
                if (lane >= o) s += y;
            }
            warp_sum[lane] = s;
        }
        __syncthreads();
        const int32_t before = carry + (warp ? warp_sum[warp - 1] : 0) + x - v;
        if (i < N) row_off[i] = before;
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1) carry = before + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = carry;
}

__global__ void pairs_write_kernel(const uint32_t* __restrict__ words, int words_per_row, int64_t N,
                                   const int32_t* __restrict__ row_off, int64_t* __restrict__ pairs, int64_t capacity) {
    const int64_t i = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (i >= N) return;
    const int lane = threadIdx.x & 31;
    int64_t pos = row_off[i];
    for (int w = 0; w < words_per_row; ++w) {
        const uint32_t word = words[i * words_per_row + w];
        if (word >> lane & 1u) {
            const int64_t at = pos + __popc(word & ((1u << lane) - 1u));
            if (at < capacity) {
                pairs[2 * at] = i;
                pairs[2 * at + 1] = (int64_t)w * 32 + lane;
            }
        }
        pos += __popc(word);
    }
}

static int make_cfg(MatcherCfg& cfg, const float* thr, int nthr, const int8_t* labels) {
    COIN_REQUIRE(thr && labels && nthr >= 1 && nthr <= kMaxThr, "matcher: nthr=%d out of [1,%d]", nthr, kMaxThr);
    COIN_REQUIRE(thr[0] > 0.0f, "matcher: thresholds[0] must be > 0");
    cfg.thr[0] = -INFINITY;
    for (int i = 0; i < nthr; ++i) {
        COIN_REQUIRE(i == 0 || thr[i - 1] <= thr[i], "matcher: thresholds must be ascending");
        cfg.thr[i + 1] = thr[i];
    }
    cfg.thr[nthr + 1] = INFINITY;
    for (int i = 0; i <= nthr; ++i) {
        COIN_REQUIRE(labels[i] >= -1 && labels[i] <= 1, "matcher: labels must be in {-1,0,1}");
        cfg.labels[i] = labels[i];
    }
    cfg.nbuckets = nthr + 1;
    return COIN_OK;
}

}  // namespace coin
using namespace coin;

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

extern "C" int coin_pairwise_iou(const float* b1, int64_t N, const float* b2, int64_t M, float* out,
                                 coin_stream_t stream) {
    COIN_REQUIRE(N >= 0 && M >= 0, "pairwise_iou: bad sizes");
    if (N == 0 || M == 0) return COIN_OK;
    COIN_REQUIRE(b1 && b2 && out && aligned16(b1) && aligned16(b2), "pairwise_iou: null or misaligned pointer");
    // rows per block chosen so that the grid covers the 148 SMs a few times without re-reading b2 much
    int rows = 16;
    while (rows < 256 && ceil_div(M, 256) * ceil_div(N, rows) > 8 * kNumSMs) rows *= 2;
    dim3 grid((unsigned)ceil_div(M, 256), (unsigned)ceil_div(N, rows));
    pairwise_iou_kernel<<<grid, 256, rows * sizeof(float4), as_stream(stream)>>>(
        reinterpret_cast<const float4*>(b1), N, reinterpret_cast<const float4*>(b2), M, out, rows);
    return check_launch("pairwise_iou_kernel");
}

// column partition of the fused kernels: G row groups x (256 / G) columns per CTA
// (M is a capacity: the live column count may be far smaller, so only very wide problems take one row group)
static int match_groups(int64_t M) { return M >= 65536 ? 1 : 8; }

extern "C" size_t coin_iou_match_workspace_floats(int64_t N, int64_t M) {
    if (N <= 0 || M <= 0) return 1;
    const int64_t cw = 256 / match_groups(M);
    return (size_t)N * (size_t)(1 + ceil_div(M, cw));
}

static int iou_match_impl(const float* gt, int64_t N, const float* boxes, int64_t M,
                          const float* thresholds_host, int nthr, const int8_t* labels_host,
                          int allow_low_quality, int64_t* matches, int8_t* match_labels,
                          float* matched_vals, float* row_max_ws, const int32_t* n_dev, const int32_t* m_dev,
                          coin_stream_t stream) {
    MatcherCfg cfg;
    if (int rc = make_cfg(cfg, thresholds_host, nthr, labels_host)) return rc;
    COIN_REQUIRE(N >= 0 && M >= 0, "iou_match: bad sizes");
    if (M == 0) return COIN_OK;
    COIN_REQUIRE(matches && match_labels, "iou_match: null output");
    cudaStream_t s = as_stream(stream);
    if (N == 0) {  // Matcher's empty-matrix rule
        empty_matcher_kernel<<<(unsigned)ceil_div(M, 256), 256, 0, s>>>(M, cfg.labels[0], matches, match_labels, matched_vals);
        return check_launch("empty_matcher_kernel");
    }
    COIN_REQUIRE(gt && boxes && aligned16(gt) && aligned16(boxes), "iou_match: null or misaligned boxes");
    const int G = match_groups(M);
    const unsigned blocks = (unsigned)ceil_div(M, 256 / G);
    const float4* g4 = reinterpret_cast<const float4*>(gt);
    const float4* b4 = reinterpret_cast<const float4*>(boxes);
    if (allow_low_quality) {
        COIN_REQUIRE(row_max_ws, "iou_match: row_max_ws is required with allow_low_quality");
        unsigned* row_max = reinterpret_cast<unsigned*>(row_max_ws);
        unsigned* cta_rmax = row_max + N;
        fill_bytes(row_max, 0, N * sizeof(float), s);
        if (G == 1) {
            iou_match_kernel<true, 1><<<blocks, 256, 0, s>>>(g4, N, b4, M, cfg, matches, match_labels, matched_vals, row_max,
                                                             cta_rmax, n_dev, m_dev);
            if (int rc = check_launch("iou_match_kernel")) return rc;
            iou_low_quality_kernel<1><<<blocks, 256, 0, s>>>(g4, N, b4, M, row_max, cta_rmax, match_labels, n_dev, m_dev);
        } else {
            iou_match_kernel<true, 8><<<blocks, 256, 0, s>>>(g4, N, b4, M, cfg, matches, match_labels, matched_vals, row_max,
                                                             cta_rmax, n_dev, m_dev);
            if (int rc = check_launch("iou_match_kernel")) return rc;
            iou_low_quality_kernel<8><<<blocks, 256, 0, s>>>(g4, N, b4, M, row_max, cta_rmax, match_labels, n_dev, m_dev);
        }
        return check_launch("iou_low_quality_kernel");
    }
    if (G == 1)
        iou_match_kernel<false, 1><<<blocks, 256, 0, s>>>(g4, N, b4, M, cfg, matches, match_labels, matched_vals, nullptr,
                                                          nullptr, n_dev, m_dev);
    else
        iou_match_kernel<false, 8><<<blocks, 256, 0, s>>>(g4, N, b4, M, cfg, matches, match_labels, matched_vals, nullptr,
                                                          nullptr, n_dev, m_dev);
    return check_launch("iou_match_kernel");
}

extern "C" int coin_iou_match(const float* gt, int64_t N, const float* boxes, int64_t M,
                              const float* thresholds_host, int nthr, const int8_t* labels_host,
                              int allow_low_quality, int64_t* matches, int8_t* match_labels,
                              float* matched_vals, float* row_max_ws, coin_stream_t stream) {
    return iou_match_impl(gt, N, boxes, M, thresholds_host, nthr, labels_host, allow_low_quality, matches, match_labels,
                          matched_vals, row_max_ws, nullptr, nullptr, stream);
}

extern "C" int coin_iou_match_dev(const float* gt, int64_t N_cap, const int32_t* n_dev, const float* boxes,
                                  int64_t M_cap, const int32_t* m_dev, const float* thresholds_host, int nthr,
                                  const int8_t* labels_host, int allow_low_quality, int64_t* matches,
                                  int8_t* match_labels, float* matched_vals, float* row_max_ws,
                                  coin_stream_t stream) {
    COIN_REQUIRE(N_cap >= 1, "iou_match_dev: N_cap must be >= 1 (an empty set is a device count of 0)");
    return iou_match_impl(gt, N_cap, boxes, M_cap, thresholds_host, nthr, labels_host, allow_low_quality, matches,
                          match_labels, matched_vals, row_max_ws, n_dev, m_dev, stream);
}

extern "C" int coin_matcher(const float* quality, int64_t N, int64_t M, const float* thresholds_host, int nthr,
                            const int8_t* labels_host, int allow_low_quality, int64_t* matches,
                            int8_t* match_labels, float* matched_vals, float* row_max_ws, coin_stream_t stream) {
    MatcherCfg cfg;
    if (int rc = make_cfg(cfg, thresholds_host, nthr, labels_host)) return rc;
    COIN_REQUIRE(N >= 0 && M >= 0, "matcher: bad sizes");
    if (M == 0) return COIN_OK;
    COIN_REQUIRE(matches && match_labels, "matcher: null output");
    cudaStream_t s = as_stream(stream);
    const unsigned blocks = (unsigned)ceil_div(M, 128);
    if (N == 0) {
        empty_matcher_kernel<<<(unsigned)ceil_div(M, 256), 256, 0, s>>>(M, cfg.labels[0], matches, match_labels, matched_vals);
        return check_launch("empty_matcher_kernel");
    }
    COIN_REQUIRE(quality, "matcher: quality is null");
    matcher_kernel<<<blocks, 128, 0, s>>>(quality, N, M, cfg, matches, match_labels, matched_vals);
    if (int rc = check_launch("matcher_kernel")) return rc;
    if (allow_low_quality) {
        COIN_REQUIRE(row_max_ws, "matcher: row_max_ws is required with allow_low_quality");
        row_max_kernel<<<(unsigned)N, 256, 0, s>>>(quality, M, row_max_ws);
        if (int rc = check_launch("row_max_kernel")) return rc;
        matrix_low_quality_kernel<<<blocks, 128, 0, s>>>(quality, N, M, row_max_ws, match_labels);
        return check_launch("matrix_low_quality_kernel");
    }
    return COIN_OK;
}

extern "C" int coin_relabel_roi(const int64_t* matches, int8_t* match_labels, int64_t M, int64_t c_begin,
                                int64_t c_end, coin_stream_t stream) {
    COIN_REQUIRE(M >= 0, "relabel_roi: bad size");
    if (M == 0) return COIN_OK;
    COIN_REQUIRE(matches && match_labels, "relabel_roi: null pointer");
    relabel_roi_kernel<<<(unsigned)ceil_div(M, 256), 256, 0, as_stream(stream)>>>(matches, match_labels, M, c_begin, c_end,
                                                                                nullptr, nullptr, nullptr, nullptr);
    return check_launch("relabel_roi_kernel");
}

extern "C" int coin_relabel_roi_dev(const int64_t* matches, int8_t* match_labels, int64_t M_cap, const int32_t* m_dev,
                                    const int32_t* len_a, const int32_t* len_b, const int32_t* len_c,
                                    coin_stream_t stream) {
    COIN_REQUIRE(M_cap >= 0, "relabel_roi_dev: bad size");
    if (M_cap == 0) return COIN_OK;
    COIN_REQUIRE(matches && match_labels && len_a && len_b && len_c, "relabel_roi_dev: null pointer");
    relabel_roi_kernel<<<(unsigned)ceil_div(M_cap, 256), 256, 0, as_stream(stream)>>>(matches, match_labels, M_cap, 0, 0,
                                                                                    len_a, len_b, len_c, m_dev);
    return check_launch("relabel_roi_kernel");
}

extern "C" int coin_relabel_rpn(int64_t* matches, int8_t* labels, int64_t M, int64_t len_a, int64_t len_c,
                                int64_t* distill_idx, int8_t* distill_labels, coin_stream_t stream) {
    COIN_REQUIRE(M >= 0, "relabel_rpn: bad size");
    if (M == 0) return COIN_OK;
    COIN_REQUIRE(matches && labels && distill_idx && distill_labels, "relabel_rpn: null pointer");
    relabel_rpn_kernel<<<(unsigned)ceil_div(M, 256), 256, 0, as_stream(stream)>>>(matches, labels, M, len_a, len_c,
                                                                                distill_idx, distill_labels, nullptr, nullptr);
    return check_launch("relabel_rpn_kernel");
}

extern "C" int coin_relabel_rpn_dev(int64_t* matches, int8_t* labels, int64_t M, const int32_t* len_a,
                                    const int32_t* len_c, int64_t* distill_idx, int8_t* distill_labels,
                                    coin_stream_t stream) {
    COIN_REQUIRE(M >= 0, "relabel_rpn_dev: bad size");
    if (M == 0) return COIN_OK;
    COIN_REQUIRE(matches && labels && distill_idx && distill_labels && len_a && len_c, "relabel_rpn_dev: null pointer");
    relabel_rpn_kernel<<<(unsigned)ceil_div(M, 256), 256, 0, as_stream(stream)>>>(matches, labels, M, 0, 0, distill_idx,
                                                                                distill_labels, len_a, len_c);
    return check_launch("relabel_rpn_kernel");
}

extern "C" size_t coin_iou_pairs_workspace_bytes(int64_t N, int64_t M) {
    Carver c(nullptr);
    c.take<uint32_t>((size_t)N * (size_t)ceil_div(M, 32));
    c.take<int32_t>((size_t)N);
    c.take<int32_t>((size_t)N);
    return c.used() + 256;
}

extern "C" int coin_iou_pairs_ge(const float* b1, int64_t N, const float* b2, int64_t M, float thr,
                                 int64_t* pairs, int32_t* count, int64_t capacity, void* ws, size_t ws_bytes,
                                 coin_stream_t stream) {
    COIN_REQUIRE(N >= 0 && M >= 0 && capacity >= 0 && count, "iou_pairs_ge: bad arguments");
    cudaStream_t s = as_stream(stream);
    if (N == 0 || M == 0) {
        fill_bytes(count, 0, sizeof(int32_t), s);
        return COIN_OK;
    }
    COIN_REQUIRE(b1 && b2 && aligned16(b1) && aligned16(b2) && (pairs || capacity == 0), "iou_pairs_ge: null or misaligned pointer");
    if (!ws || ws_bytes < coin_iou_pairs_workspace_bytes(N, M))
        return fail(COIN_ERR_CAPACITY, "iou_pairs_ge: workspace too small (%zu < %zu)", ws_bytes, coin_iou_pairs_workspace_bytes(N, M));
    Carver c(ws);
    const int wpr = (int)ceil_div(M, 32);
    uint32_t* words = c.take<uint32_t>((size_t)N * wpr);
    int32_t* row_count = c.take<int32_t>((size_t)N);
    int32_t* row_off = c.take<int32_t>((size_t)N);
    const unsigned blocks = (unsigned)ceil_div(N, 4);
    pairs_ballot_kernel<<<blocks, 128, 0, s>>>(reinterpret_cast<const float4*>(b1), N, reinterpret_cast<const float4*>(b2),
                                               M, thr, words, wpr, row_count);
    if (int rc = check_launch("pairs_ballot_kernel")) return rc;
    pairs_scan_kernel<<<1, 1024, 0, s>>>(row_count, N, row_off, count);
    if (int rc = check_launch("pairs_scan_kernel")) return rc;
    pairs_write_kernel<<<blocks, 128, 0, s>>>(words, wpr, N, row_off, pairs, capacity);
    return check_launch("pairs_write_kernel");
}
