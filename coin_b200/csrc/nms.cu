// nms.cu -- greedy NMS / batched NMS as a sort + 64x64 bit-mask + single-CTA sweep, with early exit on max_keep.
//
// Pipeline (all launches fixed by the host-known capacity; live counts and the early-exit flag stay on the device):
//   sort      n <= 4096: one CTA, bitonic in shared memory. n <= 65536: every 4096-box chunk is sorted by its own CTA,
//             then ONE kernel ranks each key across the other chunks by binary search (keys are unique: score, index) and
//             gathers the boxes straight into sorted order - 2 launches instead of CUB's ~14. Above that: CUB radix sort.
//   part 1    mask + sweep over the first R1 = max(1024, 1.5*max_keep rounded up to 256) sorted boxes only (COIN_NMS_R1_PCT). Callers that
//             ask for keep[:k] (detections: k = 100; RPN proposals: k = 2000; d2's find_top_rpn_proposals) are usually
//             satisfied here: the mask shrinks from n^2/2 to R1^2/2 pairs and the sweep from n/64 to R1/64 tiles.
//   part 2    only if part 1 neither filled max_keep nor covered all boxes (device flag): the kept rows of part 1 are
//             OR-ed into the removed vector of the remaining columns without storing their mask words, the remaining rows
//             get their upper-triangular mask, and the sweep continues where part 1 stopped. Its kernels return at once
//             when the flag says done.
//
// Replaces torchvision::nms / torchvision.ops.boxes.batched_nms behind detectron2.layers.batched_nms,
// reached from coin/modeling/roi_heads/fast_rcnn.py:164 (final detections, per class),
// coin/layers/nms.py:207 ('nms'/'mm' pseudo-label NMS), coin/modeling/meta_arch/clip_rcnn.py:161 and
// the d2 RPN proposal selection (<- coin/modeling/proposal_generator/rpn.py:113).
//
// Semantics follow the CPU oracle (the torchvision CPU kernel): stable descending sort (ties: lower
// original index first), areas (x2-x1)*(y2-y1), suppress when inter/(a_i+a_j-inter) > thr with the
// float IoU compared against the DOUBLE threshold (done here by rounding the threshold down to the
// nearest float, which is equivalent for a strict '>'), keep list in descending-score order.
// batched variants: TRICK adds idx*(max+1) to the coordinates in fp32 exactly as torchvision does
// (so near-threshold roundings match), VANILLA restricts suppression to equal classes.
#include <climits>

#include <cub/device/device_radix_sort.cuh>

#include "sort_common.cuh"

namespace coin {

// meta[0] = live box count n (<= n_cap), meta[1] = resolved strategy. Every kernel of the pipeline is
// launched for the host-known capacity n_cap and reads the live n from `meta`, so a caller can chain
// NMS behind a device-side compaction without a host synchronisation.
__device__ __forceinline__ int resolve_n(int n_cap, const int32_t* n_dev) {
    return n_dev ? min(max(*n_dev, 0), n_cap) : n_cap;
}
__device__ __forceinline__ int resolve_strategy(int strategy, int n) {
    // torchvision CPU rule (boxes.numel() > 4000 -> per-class) behind the d2 wrapper
    if (strategy == COIN_NMS_AUTO) return (n * 4 > 4000) ? COIN_NMS_VANILLA : COIN_NMS_TRICK;
    return strategy;
}

// Large-n path: 32-bit keys (descending-score order) + 32-bit values (original index). The radix sort is
// stable and the input is in index order, so exact ties keep the lower index first -- the same order as the
// 64-bit (key, index) composite of the single-CTA path, in 4 radix passes instead of 8.
__global__ void make_keys_kernel(const float* __restrict__ scores, int n_cap, const int32_t* __restrict__ n_dev,
                                 int strategy, uint32_t* __restrict__ keys, uint32_t* __restrict__ vals,
                                 int32_t* __restrict__ meta) {
    const int n = resolve_n(n_cap, n_dev);
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) { meta[0] = n; meta[1] = resolve_strategy(strategy, n); }
    if (i < n_cap) {
        keys[i] = i < n ? (uint32_t)(sort_key(__ldg(scores + i), 0u) >> 32) : 0xffffffffu;
        vals[i] = (uint32_t)i;
    }
}

// max over all 4n coordinates (for the coordinate trick); result as orderable int via atomicMax
__global__ void max_coord_kernel(const float4* __restrict__ boxes, const int32_t* __restrict__ meta,
                                 float* __restrict__ out_max) {
    const int n = meta[0];
    if (meta[1] != COIN_NMS_TRICK) return;
    float m = -INFINITY;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float4 b = __ldg(boxes + i);
        m = fmaxf(fmaxf(m, fmaxf(b.x, b.y)), fmaxf(b.z, b.w));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) {
        // float max through integer atomics: order-preserving map for any sign
        int bits = __float_as_int(m);
        bits = bits >= 0 ? bits : bits ^ 0x7fffffff;
        atomicMax(reinterpret_cast<int*>(out_max), bits);
    }
}

__device__ __forceinline__ float decode_max(const float* p) {
    int bits = *reinterpret_cast<const int*>(p);
    bits = bits >= 0 ? bits : bits ^ 0x7fffffff;
    return __int_as_float(bits);
}

// sorted position -> box (with the class offset when strategy == TRICK), class id, original index
__global__ void gather_sorted_kernel(const uint32_t* __restrict__ sorted_idx, const float4* __restrict__ boxes,
                                     const int64_t* __restrict__ idxs, const int32_t* __restrict__ meta,
                                     const float* __restrict__ max_coord, float4* __restrict__ sboxes,
                                     int32_t* __restrict__ scls, int32_t* __restrict__ order) {
    const int n = meta[0], strategy = meta[1];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t src = sorted_idx[i];
    float4 b = __ldg(boxes + src);
    const int64_t cls = idxs ? __ldg(idxs + src) : 0;
    if (strategy == COIN_NMS_TRICK) {
        const float off = (float)cls * (decode_max(max_coord) + 1.0f);
        b.x += off; b.y += off; b.z += off; b.w += off;
    }
    sboxes[i] = b;
    scls[i] = (int32_t)cls;
    order[i] = (int32_t)src;
}

// Single-CTA path for n <= kSmallSort: key generation, bitonic sort, max-reduce and gather fused.
__global__ void COIN_SORT_BOUNDS
small_sort_gather_kernel(const float* __restrict__ scores, const float4* __restrict__ boxes,
                         const int64_t* __restrict__ idxs, int n_cap, const int32_t* __restrict__ n_dev,
                         int strategy_in, float4* __restrict__ sboxes, int32_t* __restrict__ scls,
                         int32_t* __restrict__ order, int32_t* __restrict__ meta) {
    extern __shared__ uint64_t skeys[];
    __shared__ float smax[32];
    const int n = resolve_n(n_cap, n_dev);
    const int strategy = resolve_strategy(strategy_in, n);
    if (threadIdx.x == 0) { meta[0] = n; meta[1] = strategy; }
    int npow = 1;
    while (npow < n) npow <<= 1;
    float m = -INFINITY;
    for (int i = threadIdx.x; i < npow; i += blockDim.x) {
        skeys[i] = i < n ? sort_key(__ldg(scores + i), (uint32_t)i) : ~0ull;
        if (strategy == COIN_NMS_TRICK && i < n) {
            const float4 b = __ldg(boxes + i);
            m = fmaxf(fmaxf(m, fmaxf(b.x, b.y)), fmaxf(b.z, b.w));
        }
    }
    if (strategy == COIN_NMS_TRICK) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
        if ((threadIdx.x & 31) == 0) smax[threadIdx.x >> 5] = m;
    }
    __syncthreads();
    bitonic_sort_smem<kSortThreads>(skeys, npow);
    float mx = 0.0f;
    if (strategy == COIN_NMS_TRICK) {
        mx = smax[0];
        for (int w = 1; w < (int)(blockDim.x >> 5); ++w) mx = fmaxf(mx, smax[w]);
    }
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const uint32_t src = (uint32_t)(skeys[i] & 0xffffffffu);
        float4 b = __ldg(boxes + src);
        const int64_t cls = idxs ? __ldg(idxs + src) : 0;
        if (strategy == COIN_NMS_TRICK) {
            const float off = (float)cls * (mx + 1.0f);
            b.x += off; b.y += off; b.z += off; b.w += off;
        }
        sboxes[i] = b;
        scls[i] = (int32_t)cls;
        order[i] = (int32_t)src;
    }
}

// ---- mid-size sort: 4096-box chunks sorted independently, then ranked against each other ---------------------
constexpr int kMaxChunks = 16;   // n <= 65536 takes this path

// ---- class-segmented pipeline (per-class NMS of many boxes: the 20-class operator sweep) --------------
// Sorting by (class, score desc, index) puts every class in one contiguous segment, so (1) the suppression mask is only
// evaluated for tile pairs whose class ranges meet (K classes: ~K times fewer IoUs and mask words), (2) every class is
// swept by its own CTA, and (3) a last sort of the kept boxes by (score desc, index) restores the order batched_nms
// returns. The 64-bit key packs class:10 | ~orderable(score):32 | index:22; class ids outside [0, 1024) or the
// coordinate-trick strategy fold everything into one segment (correct, just not faster than the dense pipeline).
constexpr int kSegBuckets = 1024;
constexpr int kMaxChunksSeg = 64;             // n <= 262144
constexpr int kSegMinAlloc = 1024;            // smallest n whose workspace holds the segmented pipeline's buffers
constexpr uint64_t kSegIdxMask = (1ull << 22) - 1;
constexpr uint64_t kSegOrderMask = (1ull << 54) - 1;

__device__ __forceinline__ uint64_t seg_key(int bucket, float s, uint32_t idx) {
    return ((uint64_t)bucket << 54) | ((sort_key(s, 0u) >> 32) << 22) | idx;
}

// meta[2] = 1 when a class id does not fit a bucket (then everything is one segment)
__global__ void seg_class_range_kernel(const int64_t* __restrict__ idxs, int n_cap, const int32_t* __restrict__ n_dev,
                                       int32_t* __restrict__ meta) {
    const int n = resolve_n(n_cap, n_dev);
    bool bad = false;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int64_t c = __ldg(idxs + i);
        bad |= c < 0 || c >= kSegBuckets;
    }
    if (__any_sync(0xffffffffu, bad) && (threadIdx.x & 31) == 0) atomicOr(meta + 2, 1);
}

__global__ void COIN_SORT_BOUNDS
chunk_sort_kernel(const float* __restrict__ scores, const float4* __restrict__ boxes, int n_cap,
                  const int32_t* __restrict__ n_dev, int strategy_in, uint64_t* __restrict__ ckeys,
                  float* __restrict__ max_coord, int32_t* __restrict__ meta, const int64_t* __restrict__ seg_idxs) {
    extern __shared__ uint64_t skeys[];
    const int n = resolve_n(n_cap, n_dev);
    const int strategy = resolve_strategy(strategy_in, n);
    if (blockIdx.x == 0 && threadIdx.x == 0) { meta[0] = n; meta[1] = strategy; }
    // class-major keys (segmented pipeline) unless a class id is out of range or the strategy is the coordinate trick
    const bool seg = seg_idxs && strategy == COIN_NMS_VANILLA && meta[2] == 0;
    const int base = blockIdx.x * kChunk;
    float m = -INFINITY;
    for (int i = threadIdx.x; i < kChunk; i += blockDim.x) {
        const int g = base + i;
        if (seg_idxs) skeys[i] = g < n ? seg_key(seg ? (int)__ldg(seg_idxs + g) : 0, __ldg(scores + g), (uint32_t)g) : ~0ull;
        else skeys[i] = g < n ? sort_key(__ldg(scores + g), (uint32_t)g) : ~0ull;
        if (strategy == COIN_NMS_TRICK && g < n) {
            const float4 b = __ldg(boxes + g);
            m = fmaxf(fmaxf(m, fmaxf(b.x, b.y)), fmaxf(b.z, b.w));
        }
    }
    if (strategy == COIN_NMS_TRICK) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
        if ((threadIdx.x & 31) == 0) {
            int bits = __float_as_int(m);
            bits = bits >= 0 ? bits : bits ^ 0x7fffffff;
            atomicMax(reinterpret_cast<int*>(max_coord), bits);
        }
    }
    __syncthreads();
    if (base < n) {     // chunks beyond the live boxes hold only padding; a partly filled chunk sorts its live prefix
        int npow = 2;
        while (npow < min(n - base, kChunk)) npow <<= 1;
        bitonic_sort_smem<kSortThreads>(skeys, npow);
    }
    for (int i = threadIdx.x; i < kChunk; i += blockDim.x) ckeys[base + i] = skeys[i];
}

// rank of a key = its position in its own chunk + the number of smaller keys in every other chunk (binary search;
// keys are unique because they carry the box index). The gather of gather_sorted_kernel is fused in.
__global__ void merge_rank_gather_kernel(const uint64_t* __restrict__ ckeys, int nchunks, const float4* __restrict__ boxes,
                                         const int64_t* __restrict__ idxs, const int32_t* __restrict__ meta,
                                         const float* __restrict__ max_coord, float4* __restrict__ sboxes,
                                         int32_t* __restrict__ scls, int32_t* __restrict__ order,
                                         uint64_t* __restrict__ okeys) {
    const int n = meta[0], strategy = meta[1];
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= nchunks * kChunk) return;
    const uint64_t key = ckeys[e];
    if (key == ~0ull) return;                       // padding
    const int g = e / kChunk;
    int rank = e - g * kChunk;
    const int live_chunks = (n + kChunk - 1) / kChunk;
    for (int h = 0; h < live_chunks; ++h) {
        if (h == g) continue;
        const uint64_t* c = ckeys + (size_t)h * kChunk;
        int lo = 0, hi = kChunk;                    // first position with c[pos] >= key
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (__ldg(c + mid) < key) lo = mid + 1; else hi = mid;
        }
        rank += lo;
    }
    const uint32_t src = okeys ? (uint32_t)(key & kSegIdxMask) : (uint32_t)(key & 0xffffffffu);
    float4 b = __ldg(boxes + src);
    const int64_t cls = idxs ? __ldg(idxs + src) : 0;
    if (strategy == COIN_NMS_TRICK) {
        const float off = (float)cls * (decode_max(max_coord) + 1.0f);
        b.x += off; b.y += off; b.z += off; b.w += off;
    }
    sboxes[rank] = b;
    scls[rank] = (int32_t)cls;
    order[rank] = (int32_t)src;
    if (okeys) okeys[rank] = key;
}

// Mask of the segmented pipeline: CTA (j, rt) owns row tile rt and walks the groups of 4 column tiles rt/4 + j,
// rt/4 + j + gridDim.x, ... until the group's first class lies beyond the row tile's last class (the sequence is sorted
// by class, so every later group does too). Same words as nms_mask_kernel: mask[i][ct] for ct >= rt (bits j > i only),
// lower[i] = transposed diagonal word, rowflags[i] = bitmap of the non-zero words. Words of skipped tile pairs are never
// written and never read (the sweep only follows rowflags).
template <bool FAST>
__global__ void __launch_bounds__(256)
nms_seg_mask_kernel(const float4* __restrict__ sboxes, const int32_t* __restrict__ scls, const uint64_t* __restrict__ okeys,
                    const int32_t* __restrict__ meta, int stride, float thr, uint64_t* __restrict__ mask,
                    uint64_t* __restrict__ lower, uint64_t* __restrict__ rowflags, int fw) {
    const int n = meta[0];
    const bool same_class_only = meta[1] == COIN_NMS_VANILLA;
    const int rt = blockIdx.y;
    if (rt * 64 >= n) return;
    const int colblocks = (n + 63) >> 6;
    const int r = threadIdx.x;
    const int i = rt * 64 + r;
    const int row_last = (int)(__ldg(okeys + min(rt * 64 + 63, n - 1)) >> 54);
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
    int32_t cls_a = -1;
    if (i < n) { a = __ldg(sboxes + i); cls_a = __ldg(scls + i); }
    const float area_a = box_area(a);
    __shared__ float4 cb[4][64];
    __shared__ float ca[4][64];
    __shared__ int32_t cc[4][64];
    for (int cg = (rt >> 2) + blockIdx.x; cg * 4 < colblocks; cg += gridDim.x) {
        const int ctf = max(cg * 4, rt);                                      // first column tile of the group that matters
        if ((int)(__ldg(okeys + ctf * 64) >> 54) > row_last) break;           // CTA-uniform: no common class from here on
        const int ct = cg * 4 + threadIdx.y;
        __syncthreads();
        {
            const int j = ct * 64 + r;
            float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
            int32_t c = -1;
            if (j < n) { b = __ldg(sboxes + j); c = __ldg(scls + j); }
            cb[threadIdx.y][r] = b;
            ca[threadIdx.y][r] = box_area(b);
            cc[threadIdx.y][r] = c;
        }
        __syncthreads();
        if (i >= n || ct >= colblocks || ct < rt) continue;
        const int jn = min(64, n - ct * 64);
        const bool diag = ct == rt;
        uint64_t word = 0, low = 0;
        for (int j = 0; j < jn; ++j) {
            if (diag && j == r) continue;
            const float4 b = cb[threadIdx.y][j];
            bool hit;
            if (FAST) {
                const float w = fminf(a.z, b.z) - fmaxf(a.x, b.x);
                const float h = fminf(a.w, b.w) - fmaxf(a.y, b.y);
                if (!(w > 0.0f && h > 0.0f)) continue;     // empty intersection (or NaN): IoU is 0 or NaN
                const float inter = w * h;
                hit = inter / (area_a + ca[threadIdx.y][j] - inter) > thr;
            } else {
                hit = iou_tv(a, area_a, b, ca[threadIdx.y][j]) > thr;
            }
            if (hit && (!same_class_only || cc[threadIdx.y][j] == cls_a)) {
                if (!diag || j > r) word |= 1ull << j; else low |= 1ull << j;
            }
        }
        mask[(size_t)i * stride + ct] = word;
        if (word) atomicOr(reinterpret_cast<unsigned long long*>(rowflags + (size_t)i * fw + (ct >> 6)), 1ull << (ct & 63));
        if (diag) lower[i] = low;
    }
}

// One CTA per class bucket: finds its segment [lo, hi) of the class-sorted sequence by binary search and sweeps the
// segment's 64-row tiles exactly like nms_sweep_kernel (fixed-point iteration on the transposed diagonal words, kept rows
// OR their flagged mask words into the shared `removed` vector), restricted to the rows of the segment. Tiles shared with
// a neighbouring segment are visited by both CTAs, each with its own rows. Output: keptbits[tile] (atomicOr).
__global__ void __launch_bounds__(256)
nms_seg_sweep_kernel(const uint64_t* __restrict__ mask, const uint64_t* __restrict__ lower,
                     const uint64_t* __restrict__ rowflags, int fw, const uint64_t* __restrict__ okeys,
                     const int32_t* __restrict__ meta, int stride, uint64_t* __restrict__ keptbits) {
    extern __shared__ uint64_t removed[];       // one word per tile of the segment
    __shared__ int s_lo, s_hi;
    __shared__ uint64_t s_kept;
    const int n = meta[0];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid < 2) {                              // first position whose bucket is >= blockIdx.x + tid
        const uint64_t want = (uint64_t)(blockIdx.x + tid);
        int lo = 0, hi = n;
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if ((__ldg(okeys + mid) >> 54) < want) lo = mid + 1; else hi = mid;
        }
        if (tid == 0) s_lo = lo; else s_hi = lo;
    }
    __syncthreads();
    const int lo = s_lo, hi = s_hi;
    if (lo >= hi) return;
    const int t0 = lo >> 6, t1 = (hi - 1) >> 6;
    for (int c = tid; c <= t1 - t0; c += blockDim.x) removed[c] = 0;
    __syncthreads();
    uint64_t low0 = 0, low1 = 0;
    if (warp == 0) {
        if (t0 * 64 + lane < n) low0 = lower[t0 * 64 + lane];
        if (t0 * 64 + lane + 32 < n) low1 = lower[t0 * 64 + lane + 32];
    }
    for (int t = t0; t <= t1; ++t) {
        const int base = t * 64;
        if (warp == 0) {
            const int rlo = max(lo - base, 0), rhi = min(hi - base, 64);      // the segment's rows of this tile
            uint64_t valid = rhi >= 64 ? ~0ull : ((1ull << rhi) - 1ull);
            valid &= ~((1ull << rlo) - 1ull);
            const uint64_t cur = removed[t - t0] | ~valid;
            const bool a0 = !((cur >> lane) & 1ull), a1 = !((cur >> (lane + 32)) & 1ull);
            uint64_t kept = (uint64_t)__ballot_sync(0xffffffffu, a1) << 32 | __ballot_sync(0xffffffffu, a0);
            while (true) {
                const bool k0 = a0 && !(low0 & kept), k1 = a1 && !(low1 & kept);
                const uint64_t nxt = (uint64_t)__ballot_sync(0xffffffffu, k1) << 32 | __ballot_sync(0xffffffffu, k0);
                if (nxt == kept) break;
                kept = nxt;
            }
            if (lane == 0) {
                s_kept = kept;
                if (kept) atomicOr(reinterpret_cast<unsigned long long*>(keptbits + t), (unsigned long long)kept);
            }
            low0 = low1 = 0;                                                   // next tile's diagonal words
            if (t < t1) {
                if (base + 64 + lane < n) low0 = lower[base + 64 + lane];
                if (base + 96 + lane < n) low1 = lower[base + 96 + lane];
            }
        }
        __syncthreads();
        if (t < t1) {
            const uint64_t kept = s_kept;
            const int f0 = (t + 1) >> 6, f1 = t1 >> 6, nf = f1 - f0 + 1;
            for (int w = tid; w < 64 * nf; w += blockDim.x) {
                const int rr = w / nf, f = f0 + (w - rr * nf);
                if (!((kept >> rr) & 1ull)) continue;
                const size_t row = (size_t)base + rr;
                uint64_t bits = rowflags[row * fw + f];
                const int shift = t - f * 64 + 1;                              // clear the columns <= t
                if (shift >= 64) bits = 0; else if (shift > 0) bits &= ~0ull << shift;
                while (bits) {
                    const int c = f * 64 + __ffsll((long long)bits) - 1;
                    bits &= bits - 1;
                    if (c <= t1)
                        atomicOr(reinterpret_cast<unsigned long long*>(&removed[c - t0]), (unsigned long long)mask[row * stride + c]);
                }
            }
        }
        __syncthreads();
    }
}

// kept positions -> keys ordered by (score desc, index) for the final sort; everything else becomes padding
__global__ void seg_final_keys_kernel(const uint64_t* __restrict__ okeys, const uint64_t* __restrict__ keptbits,
                                      const int32_t* __restrict__ meta, int total, uint64_t* __restrict__ ckeys,
                                      int32_t* __restrict__ nkeep) {
    const int n = meta[0];
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    bool kept = false;
    if (p < n) kept = (keptbits[p >> 6] >> (p & 63)) & 1ull;
    if (p < total) ckeys[p] = kept ? (okeys[p] & kSegOrderMask) : ~0ull;
    const int cnt = __popc(__ballot_sync(0xffffffffu, kept));
    if ((threadIdx.x & 31) == 0 && cnt) atomicAdd(nkeep, cnt);
}

__global__ void COIN_SORT_BOUNDS
chunk_sort_keys_kernel(uint64_t* __restrict__ ckeys, const int32_t* __restrict__ meta) {
    extern __shared__ uint64_t skeys[];
    const int base = blockIdx.x * kChunk;
    if (base >= meta[0]) return;                 // only padding
    for (int i = threadIdx.x; i < kChunk; i += blockDim.x) skeys[i] = ckeys[base + i];
    __syncthreads();
    bitonic_sort_smem<kSortThreads>(skeys, kChunk);
    for (int i = threadIdx.x; i < kChunk; i += blockDim.x) ckeys[base + i] = skeys[i];
}

__global__ void merge_rank_keep_kernel(const uint64_t* __restrict__ ckeys, int nchunks, const int32_t* __restrict__ meta,
                                       int64_t* __restrict__ keep) {
    const int n = meta[0];
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= nchunks * kChunk) return;
    const uint64_t key = ckeys[e];
    if (key == ~0ull) return;                       // not kept / padding
    const int g = e / kChunk;
    int rank = e - g * kChunk;
    const int live_chunks = (n + kChunk - 1) / kChunk;
    for (int h = 0; h < live_chunks; ++h) {
        if (h == g) continue;
        const uint64_t* c = ckeys + (size_t)h * kChunk;
        int lo = 0, hi = kChunk;
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (__ldg(c + mid) < key) lo = mid + 1; else hi = mid;
        }
        rank += lo;
    }
    keep[rank] = (int64_t)(key & kSegIdxMask);
}

// 64 x (4*64) tile of the upper-triangular suppression mask per CTA; row-major [n][colblocks] so
// that the sweep reads whole rows coalesced. Bits are set only for j > i. For the diagonal tiles the
// transposed word is emitted as well: lower[i] = { j < i in the same 64-row tile : IoU(j, i) > thr }
// (IoU is bitwise symmetric), which lets the sweep resolve a tile by fixed-point iteration.
// Pairs with an empty intersection are rejected before the division (their IoU is 0 or NaN, never > thr
// for thr >= 0); FAST = false keeps the plain evaluation for negative thresholds.
// Two-part operation (early exit on max_keep): part 1 is launched with n_limit = R1 and t1 = 0; part 2 with ct0 = t1 =
// R1 / 64 and n_limit = INT_MAX: it returns at once when state[1] (done) is set; rows of part 1 (row tile < t1) that the
// first sweep KEPT (keptbits) OR their words into removed_g[column tile] instead of storing them.
template <bool FAST>
__global__ void __launch_bounds__(256)
nms_mask_kernel(const float4* __restrict__ sboxes, const int32_t* __restrict__ scls,
                const int32_t* __restrict__ meta, int colblocks, float thr, uint64_t* __restrict__ mask,
                uint64_t* __restrict__ lower, uint64_t* __restrict__ rowflags, int fw, int ct0, int n_limit,
                const int32_t* __restrict__ state, const uint64_t* __restrict__ keptbits,
                uint64_t* __restrict__ removed_g, int t1) {
    if (t1 > 0 && state[1]) return;            // part 1 already delivered max_keep boxes (or covered every box)
    const int n = min(meta[0], n_limit);
    const bool same_class_only = meta[1] == COIN_NMS_VANILLA;
    const int rt = blockIdx.y;                 // row tile
    const int ct = ct0 + blockIdx.x * 4 + threadIdx.y;  // col tile of this thread row
    if (ct0 + blockIdx.x * 4 + 3 < rt) return;       // whole CTA below the diagonal
    if (rt * 64 >= n || (ct0 + blockIdx.x * 4) * 64 >= n) return;  // beyond the live boxes (capacity launch)
    __shared__ float4 cb[4][64];
    __shared__ float ca[4][64];
    __shared__ int32_t cc[4][64];
    const int r = threadIdx.x;
    {
        const int j = ct * 64 + r;
        float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
        int32_t c = -1;
        if (ct < colblocks && j < n) { b = __ldg(sboxes + j); c = __ldg(scls + j); }
        cb[threadIdx.y][r] = b;
        ca[threadIdx.y][r] = box_area(b);
        cc[threadIdx.y][r] = c;
    }
    __syncthreads();
    const int i = rt * 64 + r;
    if (i >= n || ct >= colblocks || ct < rt) return;
    const bool part1_row = rt < t1;
    if (part1_row && !((keptbits[rt] >> r) & 1ull)) return;    // suppressed in part 1: suppresses nobody
    const float4 a = __ldg(sboxes + i);
    const float area_a = box_area(a);
    const int32_t cls_a = __ldg(scls + i);
    const int jn = min(64, n - ct * 64);
    const bool diag = ct == rt;
    uint64_t word = 0, low = 0;
    for (int j = diag ? 0 : 0; j < jn; ++j) {
        if (diag && j == r) continue;
        const float4 b = cb[threadIdx.y][j];
        bool hit;
        if (FAST) {
            const float w = fminf(a.z, b.z) - fmaxf(a.x, b.x);
            const float h = fminf(a.w, b.w) - fmaxf(a.y, b.y);
            if (!(w > 0.0f && h > 0.0f)) continue;     // empty intersection (or NaN): IoU is 0 or NaN
            const float inter = w * h;
            hit = inter / (area_a + ca[threadIdx.y][j] - inter) > thr;
        } else {
            hit = iou_tv(a, area_a, b, ca[threadIdx.y][j]) > thr;
        }
        if (hit && (!same_class_only || cc[threadIdx.y][j] == cls_a)) {
            if (!diag || j > r) word |= 1ull << j; else low |= 1ull << j;
        }
    }
    if (part1_row) {
        if (word) atomicOr(reinterpret_cast<unsigned long long*>(removed_g + ct), (unsigned long long)word);
        return;
    }
    mask[(size_t)i * colblocks + ct] = word;
    // rowflags[i]: bitmap of the column blocks with a non-zero word (zero-filled by the host side). Overlap is
    // sparse, so the sweep reads this bitmap and then only the few non-zero words of a kept row.
    if (word) atomicOr(reinterpret_cast<unsigned long long*>(rowflags + (size_t)i * fw + (ct >> 6)), 1ull << (ct & 63));
    if (diag) lower[i] = low;
}

// Single-CTA sweep over 64-row tiles. Warp 0 resolves a tile by FIXED-POINT iteration on the transposed
// diagonal words: kept = alive & ~any(lower & kept), repeated until stable (the unique solution of the
// greedy recurrence; a few rounds instead of a 64-step serial chain). The other warps OR the mask rows of
// the PREVIOUS tile's kept boxes into the `removed` bit-vector one tile behind: one thread per kept row reads the
// row's bitmap of non-zero column blocks and fetches only those words (shared-memory atomicOr); the one
// column warp 0 needs immediately (tile t-1 -> column t) is prefetched by warp 0 itself before the kept set is
// known. 256 threads / ~2 KB of static shared memory: the CTA fits
// next to resident HBM-bound kernels of other streams.
__global__ void __launch_bounds__(256)
nms_sweep_kernel(const uint64_t* __restrict__ mask, const uint64_t* __restrict__ lower,
                 const uint64_t* __restrict__ rowflags, int fw, const int32_t* __restrict__ order,
                 const int32_t* __restrict__ meta, int stride, int64_t max_keep, int64_t* __restrict__ keep,
                 int32_t* __restrict__ nkeep, int t_begin, int n_limit, int32_t* __restrict__ state,
                 uint64_t* __restrict__ keptbits, const uint64_t* __restrict__ removed_g) {
    // two-part operation: part 1 = (t_begin 0, n_limit R1, state set): sweeps the first R1 boxes, records the kept rows
    // of every tile and whether the job is done; part 2 = (t_begin R1/64, n_limit INT_MAX): continues unless done.
    extern __shared__ uint64_t removed[];  // colblocks words
    if (t_begin > 0 && state[1]) return;
    const int n_all = meta[0];
    const int n = min(n_all, n_limit);
    const int colblocks = (n + 63) >> 6;   // live column blocks; `stride` is the row pitch of `mask`
    __shared__ uint64_t s_kept[2];
    __shared__ int s_krow[2][64];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int c = tid; c < colblocks; c += blockDim.x) removed[c] = (t_begin > 0 && c >= t_begin) ? removed_g[c] : 0;
    if (tid < 2) s_kept[tid] = 0;
    __syncthreads();
    int64_t nk = t_begin > 0 ? (int64_t)state[0] : 0;
    const int64_t limit = max_keep >= 0 ? max_keep : (int64_t)n_all;
    const bool fast_flags = 64 * fw <= (int)blockDim.x - 32;   // one helper thread per (row, flag word) of a tile
    uint64_t myflag = 0;                    // helpers: the flag word loaded one iteration earlier
    uint64_t pref0 = 0, pref1 = 0;          // warp 0: mask[(t-1)*64 + lane (+32)][t]
    uint64_t low0 = 0, low1 = 0;            // warp 0: lower words of tile t
    if (warp == 0) {
        if (t_begin * 64 + lane < n) low0 = lower[t_begin * 64 + lane];
        if (t_begin * 64 + lane + 32 < n) low1 = lower[t_begin * 64 + lane + 32];
    }
    for (int t = t_begin; t < colblocks && nk < limit; ++t) {
        const int buf = t & 1;
        if (warp == 0) {
            const uint64_t prev = t > t_begin ? s_kept[buf ^ 1] : 0ull;
            uint64_t v = ((prev >> lane) & 1ull ? pref0 : 0ull) | ((prev >> (lane + 32)) & 1ull ? pref1 : 0ull);
            const uint32_t vlo = __reduce_or_sync(0xffffffffu, (uint32_t)v);
            const uint32_t vhi = __reduce_or_sync(0xffffffffu, (uint32_t)(v >> 32));
            uint64_t cur = removed[t] | ((uint64_t)vhi << 32 | vlo);
            const int nr = min(64, n - t * 64);
            if (nr < 64) cur |= ~0ull << nr;
            // prefetch for the next tile (independent of this tile's result)
            uint64_t npref0 = 0, npref1 = 0, nlow0 = 0, nlow1 = 0;
            if (t + 1 < colblocks) {
                const int r0 = t * 64 + lane, r1 = r0 + 32;
                if (r0 < n) npref0 = mask[(size_t)r0 * stride + t + 1];
                if (r1 < n) npref1 = mask[(size_t)r1 * stride + t + 1];
                const int q0 = (t + 1) * 64 + lane, q1 = q0 + 32;
                if (q0 < n) nlow0 = lower[q0];
                if (q1 < n) nlow1 = lower[q1];
            }
            const bool a0 = !((cur >> lane) & 1ull), a1 = !((cur >> (lane + 32)) & 1ull);
            uint64_t kept = (uint64_t)__ballot_sync(0xffffffffu, a1) << 32 | __ballot_sync(0xffffffffu, a0);
            while (true) {
                const bool k0 = a0 && !(low0 & kept), k1 = a1 && !(low1 & kept);
                const uint64_t nxt = (uint64_t)__ballot_sync(0xffffffffu, k1) << 32 | __ballot_sync(0xffffffffu, k0);
                if (nxt == kept) break;
                kept = nxt;
            }
            // honour max_keep: drop the kept rows beyond the limit (keep[:max_keep])
            const int64_t room = limit - nk;
            while ((int64_t)__popcll(kept) > room) kept &= ~(1ull << (63 - __clzll(kept)));
            if (lane == 0) s_kept[buf] = kept;
            if ((kept >> lane) & 1ull) {
                const int pos = __popcll(kept & ((1ull << lane) - 1ull));
                s_krow[buf][pos] = lane;
                keep[nk + pos] = order[t * 64 + lane];
            }
            if ((kept >> (lane + 32)) & 1ull) {
                const int pos = __popcll(kept & ((1ull << (lane + 32)) - 1ull));
                s_krow[buf][pos] = lane + 32;
                keep[nk + pos] = order[t * 64 + lane + 32];
            }
            pref0 = npref0; pref1 = npref1; low0 = nlow0; low1 = nlow1;
        } else if (fast_flags) {
            // helpers, n <= 12288 (<= 3 flag words per row): thread h owns flag word (row h / fw, word h % fw) of
            // the current tile. The flag word of tile t is loaded NOW (independent of the kept set) and used
            // in the next iteration, when kept(t) is known: one dependent L2 round trip per tile.
            const int h = tid - 32;
            uint64_t nflag = 0;
            if (h < 64 * fw) {
                const int r = t * 64 + h / fw;
                if (r < n) nflag = rowflags[(size_t)r * fw + h % fw];
            }
            if (t > t_begin && h < 64 * fw) {
                const uint64_t prev = s_kept[buf ^ 1];
                const int rr = h / fw, f = h - rr * fw;
                if ((prev >> rr) & 1ull) {
                    uint64_t bits = myflag;
                    const int shift = t - f * 64 + 1;          // clear the columns <= t
                    if (shift >= 64) bits = 0; else if (shift > 0) bits &= ~0ull << shift;
                    const size_t rowbase = ((size_t)(t - 1) * 64 + rr) * stride + f * 64;
                    while (bits) {                              // up to 4 loads in flight
                        int c[4];
                        uint64_t w[4];
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            c[u] = bits ? __ffsll((long long)bits) - 1 : -1;
                            if (bits) bits &= bits - 1;
                            w[u] = c[u] >= 0 ? mask[rowbase + c[u]] : 0ull;
                        }
#pragma unroll
                        for (int u = 0; u < 4; ++u)
                            if (c[u] >= 0 && w[u])
                                atomicOr(reinterpret_cast<unsigned long long*>(&removed[f * 64 + c[u]]), (unsigned long long)w[u]);
                    }
                }
            }
            myflag = nflag;
        } else if (t > t_begin) {
            // helpers, generic: tile t-1's kept rows -> removed[c] for c >= t+1 (column t was warp 0's prefetch)
            const uint64_t prev = s_kept[buf ^ 1];
            const int np = __popcll(prev);
            const int* rows = s_krow[buf ^ 1];
            const size_t base = (size_t)(t - 1) * 64;
            for (int q = tid - 32; q < np; q += blockDim.x - 32) {
                const size_t row = base + rows[q];
                for (int f = t >> 6; f < fw; ++f) {           // flag words that can hold a column > t
                    uint64_t bits = rowflags[row * fw + f];
                    const int shift = t - f * 64 + 1;          // clear the columns <= t
                    if (shift >= 64) bits = 0; else if (shift > 0) bits &= ~0ull << shift;
                    while (bits) {
                        const int c = f * 64 + __ffsll((long long)bits) - 1;
                        bits &= bits - 1;
                        atomicOr(reinterpret_cast<unsigned long long*>(&removed[c]),
                                 (unsigned long long)mask[row * stride + c]);
                    }
                }
            }
        }
        __syncthreads();
        nk += __popcll(s_kept[buf]);
        if (tid == 0 && state && t_begin == 0) keptbits[t] = s_kept[buf];
    }
    if (tid == 0) {
        const bool done = !state || t_begin > 0 || nk >= limit || n_all <= n_limit;
        if (state) { state[0] = (int32_t)nk; state[1] = done ? 1 : 0; }
        if (done) *nkeep = (int32_t)nk;
    }
}

struct NmsWs {
    uint64_t *ckeys, *okeys, *keptbits, *removed_g;
    int32_t* state;
    uint32_t *keys, *keys_alt, *vals, *vals_alt;
    void* cub_tmp;
    size_t cub_bytes;
    int32_t *order, *scls, *meta;
    float4* sboxes;
    float* max_coord;
    uint64_t* mask;
    uint64_t* lower;
    uint64_t* rowflags;
    size_t total;
};

static size_t cub_sort_bytes(int64_t n) {
    size_t bytes = 0;
    cub::DoubleBuffer<uint32_t> dk(nullptr, nullptr), dv(nullptr, nullptr);
    if (cub::DeviceRadixSort::SortPairs(nullptr, bytes, dk, dv, (int)n) != cudaSuccess) {
        cudaGetLastError();
        bytes = (size_t)n * 16 + (1 << 20);  // conservative when no device can be queried
    }
    return bytes;
}

static NmsWs carve_nms(void* ws, int64_t n) {
    NmsWs w;
    Carver c(ws);
    const int64_t colblocks = ceil_div(n, 64);
    w.meta = c.take<int32_t>(64);
    w.max_coord = c.take<float>(64);
    w.order = c.take<int32_t>((size_t)n);
    w.scls = c.take<int32_t>((size_t)n);
    w.sboxes = c.take<float4>((size_t)n);
    w.keys = w.keys_alt = w.vals = w.vals_alt = nullptr;
    w.cub_tmp = nullptr;
    w.cub_bytes = 0;
    w.ckeys = w.okeys = nullptr;
    w.state = c.take<int32_t>(64);
    w.keptbits = c.take<uint64_t>((size_t)colblocks);
    w.removed_g = c.take<uint64_t>((size_t)colblocks);
    if (n >= kSegMinAlloc && n <= (int64_t)kChunk * kMaxChunksSeg) {
        w.ckeys = c.take<uint64_t>((size_t)ceil_div(n, kChunk) * kChunk);
        w.okeys = c.take<uint64_t>((size_t)n);
    }
    if (n > (int64_t)kChunk * kMaxChunks) {
        w.keys = c.take<uint32_t>((size_t)n);
        w.keys_alt = c.take<uint32_t>((size_t)n);
        w.vals = c.take<uint32_t>((size_t)n);
        w.vals_alt = c.take<uint32_t>((size_t)n);
        w.cub_bytes = cub_sort_bytes(n);
        w.cub_tmp = c.take<char>(w.cub_bytes);
    }
    w.mask = c.take<uint64_t>((size_t)n * colblocks);
    w.lower = c.take<uint64_t>((size_t)n);
    w.rowflags = c.take<uint64_t>((size_t)n * ceil_div(colblocks, 64));
    w.total = c.used();
    return w;
}

size_t nms_pipeline_workspace_bytes(int64_t n_cap) { return n_cap <= 0 ? 256 : carve_nms(nullptr, n_cap).total + 256; }

// largest float <= t  (so that  iou > result  <=>  (double)iou > t)
static float round_down_to_float(double t) {
    float f = (float)t;
    if ((double)f > t) f = nextafterf(f, -INFINITY);
    return f;
}

// n_cap: host-known capacity; n_dev: optional device int32 with the live count (<= n_cap).
int nms_sorted_pipeline(const float* boxes, const float* scores, const int64_t* idxs, int64_t n_cap,
                        const int32_t* n_dev, double thr, int strategy, int64_t max_keep, int64_t* keep,
                        int32_t* nkeep, void* ws, size_t ws_bytes, cudaStream_t s) {
    NmsWs w = carve_nms(ws, n_cap);
    if (ws_bytes < w.total) return fail(COIN_ERR_CAPACITY, "nms: workspace too small (%zu < %zu)", ws_bytes, w.total);
    const int n = (int)n_cap;
    const int colblocks = (int)ceil_div(n_cap, 64);
    const float4* b4 = reinterpret_cast<const float4*>(boxes);
    const float thr_f = round_down_to_float(thr);
    const int fw = (int)ceil_div(colblocks, 64);
    // class-segmented pipeline: per-class NMS of many boxes with no keep[:k] cut (COIN_NMS_SEG_MIN: smallest n that takes it)
    const bool segmented = idxs && strategy != COIN_NMS_PLAIN && max_keep < 0 && w.okeys &&
                           n >= option("COIN_NMS_SEG_MIN", 3500);
    if (segmented) {
        const int nchunks = (int)ceil_div(n, kChunk);
        fill_bytes(w.meta, 0, 64 * sizeof(int32_t), s);
        fill_bytes(w.max_coord, 0x80, sizeof(int), s);
        fill_bytes(w.keptbits, 0, (size_t)colblocks * sizeof(uint64_t), s);
        fill_bytes(w.rowflags, 0, (size_t)n * fw * sizeof(uint64_t), s);
        fill_bytes(nkeep, 0, sizeof(int32_t), s);
        seg_class_range_kernel<<<kNumSMs, 256, 0, s>>>(idxs, n, n_dev, w.meta);
        if (int rc = check_launch("seg_class_range_kernel")) return rc;
        chunk_sort_kernel<<<nchunks, kSortThreads, kChunk * sizeof(uint64_t), s>>>(scores, b4, n, n_dev, strategy, w.ckeys,
                                                                           w.max_coord, w.meta, idxs);
        if (int rc = check_launch("chunk_sort_kernel")) return rc;
        merge_rank_gather_kernel<<<(unsigned)ceil_div((int64_t)nchunks * kChunk, 256), 256, 0, s>>>(
            w.ckeys, nchunks, b4, idxs, w.meta, w.max_coord, w.sboxes, w.scls, w.order, w.okeys);
        if (int rc = check_launch("merge_rank_gather_kernel")) return rc;
        const dim3 grid((unsigned)std::min(8, (int)ceil_div(colblocks, 4)), (unsigned)colblocks), block(64, 4);
        if (thr_f >= 0.0f)
            nms_seg_mask_kernel<true><<<grid, block, 0, s>>>(w.sboxes, w.scls, w.okeys, w.meta, colblocks, thr_f, w.mask, w.lower,
                                                             w.rowflags, fw);
        else
            nms_seg_mask_kernel<false><<<grid, block, 0, s>>>(w.sboxes, w.scls, w.okeys, w.meta, colblocks, thr_f, w.mask, w.lower,
                                                              w.rowflags, fw);
        if (int rc = check_launch("nms_seg_mask_kernel")) return rc;
        const size_t smem = (size_t)colblocks * sizeof(uint64_t);
        if (smem > 48 * 1024) cudaFuncSetAttribute(nms_seg_sweep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        nms_seg_sweep_kernel<<<kSegBuckets, 256, smem, s>>>(w.mask, w.lower, w.rowflags, fw, w.okeys, w.meta, colblocks, w.keptbits);
        if (int rc = check_launch("nms_seg_sweep_kernel")) return rc;
        const int total = nchunks * kChunk;
        seg_final_keys_kernel<<<(unsigned)ceil_div(total, 256), 256, 0, s>>>(w.okeys, w.keptbits, w.meta, total, w.ckeys, nkeep);
        if (int rc = check_launch("seg_final_keys_kernel")) return rc;
        chunk_sort_keys_kernel<<<nchunks, kSortThreads, kChunk * sizeof(uint64_t), s>>>(w.ckeys, w.meta);
        if (int rc = check_launch("chunk_sort_keys_kernel")) return rc;
        merge_rank_keep_kernel<<<(unsigned)ceil_div(total, 256), 256, 0, s>>>(w.ckeys, nchunks, w.meta, keep);
        return check_launch("merge_rank_keep_kernel");
    }
    if (n <= kSmallSort) {
        int npow = 1;
        while (npow < n) npow <<= 1;
        small_sort_gather_kernel<<<1, kSortThreads, npow * sizeof(uint64_t), s>>>(scores, b4, idxs, n, n_dev, strategy, w.sboxes,
                                                                          w.scls, w.order, w.meta);
        if (int rc = check_launch("small_sort_gather_kernel")) return rc;
    } else if (n <= kChunk * kMaxChunks) {
        const int nchunks = (int)ceil_div(n, kChunk);
        fill_bytes(w.max_coord, 0x80, sizeof(int), s);  // 0x80808080 decodes to ~ -3.4e38
        chunk_sort_kernel<<<nchunks, kSortThreads, kChunk * sizeof(uint64_t), s>>>(scores, b4, n, n_dev, strategy, w.ckeys,
                                                                           w.max_coord, w.meta, nullptr);
        if (int rc = check_launch("chunk_sort_kernel")) return rc;
        merge_rank_gather_kernel<<<(unsigned)ceil_div((int64_t)nchunks * kChunk, 256), 256, 0, s>>>(
            w.ckeys, nchunks, b4, idxs, w.meta, w.max_coord, w.sboxes, w.scls, w.order, nullptr);
        if (int rc = check_launch("merge_rank_gather_kernel")) return rc;
    } else {
        make_keys_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, s>>>(scores, n, n_dev, strategy, w.keys, w.vals, w.meta);
        if (int rc = check_launch("make_keys_kernel")) return rc;
        cub::DoubleBuffer<uint32_t> dk(w.keys, w.keys_alt), db(w.vals, w.vals_alt);
        size_t bytes = w.cub_bytes;
        cudaError_t e = cub::DeviceRadixSort::SortPairs(w.cub_tmp, bytes, dk, db, n, 0, 32, s);
        if (e != cudaSuccess) return fail(COIN_ERR_CUDA, "nms: radix sort failed: %s", cudaGetErrorString(e));
        count_launch();
        if (strategy == COIN_NMS_TRICK || strategy == COIN_NMS_AUTO) {
            fill_bytes(w.max_coord, 0x80, sizeof(int), s);  // 0x80808080 decodes to ~ -3.4e38
            max_coord_kernel<<<kNumSMs, 256, 0, s>>>(b4, w.meta, w.max_coord);
            if (int rc = check_launch("max_coord_kernel")) return rc;
        }
        gather_sorted_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, s>>>(db.Current(), b4, idxs, w.meta, w.max_coord,
                                                                        w.sboxes, w.scls, w.order);
        if (int rc = check_launch("gather_sorted_kernel")) return rc;
    }
    fill_bytes(w.rowflags, 0, (size_t)n * fw * sizeof(uint64_t), s);
    const size_t smem = (size_t)colblocks * sizeof(uint64_t);
    if (smem > 48 * 1024) cudaFuncSetAttribute(nms_sweep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    dim3 block(64, 4);
    auto launch_mask = [&](dim3 grid, int ct0, int n_limit, int t1) {
        if (thr_f >= 0.0f)
            nms_mask_kernel<true><<<grid, block, 0, s>>>(w.sboxes, w.scls, w.meta, colblocks, thr_f, w.mask, w.lower, w.rowflags,
                                                         fw, ct0, n_limit, w.state, w.keptbits, w.removed_g, t1);
        else
            nms_mask_kernel<false><<<grid, block, 0, s>>>(w.sboxes, w.scls, w.meta, colblocks, thr_f, w.mask, w.lower, w.rowflags,
                                                          fw, ct0, n_limit, w.state, w.keptbits, w.removed_g, t1);
        return check_launch("nms_mask_kernel");
    };
    // part 1 covers the first R1 sorted boxes; with no max_keep (or few boxes) it is the whole job
    int64_t r1 = n;
    // COIN_NMS_R1_PCT: size of part 1 in percent of max_keep (default 150; a smaller part 1 does less mask work when the
    // input overlaps little, and hands over to part 2 earlier when it overlaps a lot - the result is the same)
    const int64_t pct = std::max(100, option("COIN_NMS_R1_PCT", 150));
    if (max_keep >= 0) r1 = std::min<int64_t>(n, std::max<int64_t>(1024, (max_keep * pct / 100 + 255) / 256 * 256));
    if (option("COIN_NMS_TWO_PART", 1) == 0) r1 = n;
    const int t1 = (int)(r1 / 64);
    if (r1 >= n) {
        if (int rc = launch_mask(dim3((unsigned)ceil_div(colblocks, 4), (unsigned)colblocks), 0, INT_MAX, 0)) return rc;
        nms_sweep_kernel<<<1, 256, smem, s>>>(w.mask, w.lower, w.rowflags, fw, w.order, w.meta, colblocks, max_keep, keep, nkeep,
                                              0, INT_MAX, nullptr, nullptr, nullptr);
        return check_launch("nms_sweep_kernel");
    }
    fill_bytes(w.removed_g, 0, (size_t)colblocks * sizeof(uint64_t), s);
    if (int rc = launch_mask(dim3((unsigned)ceil_div(t1, 4), (unsigned)t1), 0, (int)r1, 0)) return rc;
    nms_sweep_kernel<<<1, 256, smem, s>>>(w.mask, w.lower, w.rowflags, fw, w.order, w.meta, colblocks, max_keep, keep, nkeep, 0,
                                          (int)r1, w.state, w.keptbits, w.removed_g);
    if (int rc = check_launch("nms_sweep_kernel")) return rc;
    if (int rc = launch_mask(dim3((unsigned)ceil_div(colblocks - t1, 4), (unsigned)colblocks), t1, INT_MAX, t1)) return rc;
    nms_sweep_kernel<<<1, 256, smem, s>>>(w.mask, w.lower, w.rowflags, fw, w.order, w.meta, colblocks, max_keep, keep, nkeep, t1,
                                          INT_MAX, w.state, w.keptbits, w.removed_g);
    return check_launch("nms_sweep_kernel");
}

}  // namespace coin
using namespace coin;

extern "C" size_t coin_nms_workspace_bytes(int64_t n) { return nms_pipeline_workspace_bytes(n); }

// ---- stable descending argsort (the sort stage of the pipeline on its own) ------------------------------------------
namespace coin {
__global__ void merge_rank_order_kernel(const uint64_t* __restrict__ ckeys, int nchunks, int n, int64_t* __restrict__ order) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= nchunks * kChunk) return;
    const uint64_t key = ckeys[e];
    if (key == ~0ull) return;
    const int g = e / kChunk;
    int rank = e - g * kChunk;
    for (int h = 0; h < nchunks; ++h) {
        if (h == g) continue;
        const uint64_t* c = ckeys + (size_t)h * kChunk;
        int lo = 0, hi = kChunk;
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (__ldg(c + mid) < key) lo = mid + 1; else hi = mid;
        }
        rank += lo;
    }
    order[rank] = (int64_t)(key & 0xffffffffu);
}
}  // namespace coin

extern "C" size_t coin_argsort_desc_workspace_bytes(int64_t n) {
    return (size_t)ceil_div(std::max<int64_t>(n, 1), kChunk) * kChunk * sizeof(uint64_t) + 1024;
}

extern "C" int coin_argsort_desc(const float* scores, int64_t n, int64_t* order, void* ws, size_t ws_bytes, coin_stream_t stream) {
    COIN_REQUIRE(n >= 0, "argsort_desc: bad n");
    if (n == 0) return COIN_OK;
    COIN_REQUIRE(scores && order && ws, "argsort_desc: null pointer");
    if (n > (int64_t)kChunk * kMaxChunksSeg)
        return fail(COIN_ERR_UNSUPPORTED, "argsort_desc: n=%lld exceeds %d keys", (long long)n, kChunk * kMaxChunksSeg);
    if (ws_bytes < coin_argsort_desc_workspace_bytes(n)) return fail(COIN_ERR_CAPACITY, "argsort_desc: workspace too small");
    Carver c(ws);
    int32_t* meta = c.take<int32_t>(64);
    const int nchunks = (int)ceil_div(n, kChunk);
    uint64_t* ckeys = c.take<uint64_t>((size_t)nchunks * kChunk);
    cudaStream_t s = as_stream(stream);
    chunk_sort_kernel<<<nchunks, kSortThreads, kChunk * sizeof(uint64_t), s>>>(scores, nullptr, (int)n, nullptr, COIN_NMS_PLAIN, ckeys,
                                                                       nullptr, meta, nullptr);
    if (int rc = check_launch("chunk_sort_kernel")) return rc;
    merge_rank_order_kernel<<<(unsigned)ceil_div((int64_t)nchunks * kChunk, 256), 256, 0, s>>>(ckeys, nchunks, (int)n, order);
    return check_launch("merge_rank_order_kernel");
}

extern "C" int coin_batched_nms(const float* boxes, const float* scores, const int64_t* idxs, int64_t n,
                                double iou_threshold, int strategy, int64_t max_keep, int64_t* keep,
                                int32_t* nkeep, void* ws, size_t ws_bytes, coin_stream_t stream) {
    COIN_REQUIRE(n >= 0 && nkeep, "batched_nms: bad arguments");
    COIN_REQUIRE(strategy >= COIN_NMS_PLAIN && strategy <= COIN_NMS_AUTO, "batched_nms: bad strategy %d", strategy);
    COIN_REQUIRE(n < (1ll << 22), "batched_nms: n=%lld exceeds the supported 4M boxes", (long long)n);
    cudaStream_t s = as_stream(stream);
    if (n == 0 || max_keep == 0) {
        fill_bytes(nkeep, 0, sizeof(int32_t), s);
        return COIN_OK;
    }
    COIN_REQUIRE(boxes && scores && keep && ws, "batched_nms: null pointer");
    COIN_REQUIRE((reinterpret_cast<uintptr_t>(boxes) & 15) == 0, "batched_nms: boxes must be 16-byte aligned");
    if (strategy != COIN_NMS_PLAIN) COIN_REQUIRE(idxs, "batched_nms: idxs is required for a batched strategy");
    if (strategy == COIN_NMS_PLAIN) idxs = nullptr;
    return nms_sorted_pipeline(boxes, scores, idxs, n, nullptr, iou_threshold, strategy, max_keep, keep, nkeep, ws,
                               ws_bytes, s);
}
