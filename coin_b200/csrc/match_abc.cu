// match_abc.cu -- COIN's knowledge separation (consistent A / inconsistent B / private C) on device.
//
// Replaces CoinTrainer.match_dual_teacher, coin/engine/trainer.py:338-461, together with its helpers
// delete_duplicate_boxes (coin/utils/util.py:434-457), filter_result/find_same (:459-482),
// online_boxes_merging (:484-507) and merge_boxes (trainer.py:480-485 -> coin/layers/nms.py:24-31).
// The reference runs this on CPU tensors with Python sets, .tolist() and per-group loops, after a
// D2H copy of the teacher's detections (trainer.py:469) and before an H2D copy of the result
// (:457-459). Here one single-CTA launch per (image, tag) emits INDEX lists into the two detection
// sets plus the merged boxes; the field gathers are plain device index_selects in the wrapper.
//
// The problem is tiny (<= ~100 detections per side) and mostly sequential bookkeeping, so the goal
// is one launch and no host round trip, not bandwidth. Pairwise parts (IoU, equality, grouping) run
// across the CTA; list compaction uses warp ballots.
//
// Determinism policy: random.randint picks -> first element; Python-set iteration
// order -> ascending index. oracle/coin_ref.py implements the same policy (and the literal one).
#include "common.cuh"

namespace coin {

struct AbcArgs {
    const float4 *onb, *offb;
    const int64_t *oncls, *offcls;
    const float *ons, *offs;
    int nc, nd, tag, use_smem;
    const int32_t* nd_dev;   // optional device-side count of CLIP-detector detections (<= nd)
    float thr, w_a;
    int cap;
    int32_t *a_on, *a_off, *b_on, *b_off, *c_on, *c_off, *counts;
    float4 *a_box, *b_box;
    // scratch (global)
    float* key;        // [L]
    int32_t *first, *cnt, *isgrp, *glist, *single;  // [L]
    int32_t *uniq, *offgl;                           // [nd]
    int32_t *on_used, *off_matched, *label;          // [nc], [nd], [nc]
    int32_t *rowcnt, *rowoff;                        // [nc]
    int32_t *pon[2], *poff[2];                       // [cap]
    int32_t *flag_a, *flag_b;                        // [cap]
    int32_t *raw;                                    // [cap] row list of the set being packed
    float4* mbox;                                    // [cap]
    int32_t *g_i0, *g_m;                             // [nd]
    int32_t *outlist;                                // [cap]
};

__device__ __forceinline__ bool box_eq(const float4& a, const float4& b) {
    return a.x == b.x && a.y == b.y && a.z == b.z && a.w == b.w;
}

// Ordered compaction of {r in [0,L) : pred(r)} appended to out[*count...]; executed by warp 0 of the
// CTA, every thread must call it (it ends with a barrier). Returns the new count.
template <class Pred, class Map>
__device__ int compact_append(int L, Pred pred, Map map, int32_t* out, int count, int cap, int* s_tmp) {
    if (threadIdx.x < 32) {
        const int lane = threadIdx.x;
        int c = count;
        for (int base = 0; base < L; base += 32) {
            const int r = base + lane;
            const bool p = r < L && pred(r);
            const unsigned m = __ballot_sync(0xffffffffu, p);
            if (p) {
                const int at = c + __popc(m & ((1u << lane) - 1u));
                if (at < cap) out[at] = map(r);
            }
            c += __popc(m);
        }
        if (lane == 0) *s_tmp = c;
    }
    __syncthreads();
    const int res = *s_tmp;
    __syncthreads();
    return res;
}

// util.py:434-457 grouping. After the call: single[r] = 1 if row r is not part of a true duplicate
// group; glist[0..ng) = leaders (lowest row) of the true groups in ascending-key order;
// first[r] = leader candidate of r's key group. Returns ng. All threads must call.
__device__ int dup_groups(const float4* box, int L, const AbcArgs& a, int* s_tmp) {
    for (int r = threadIdx.x; r < L; r += blockDim.x) {
        const float4 b = box[r];
        a.key[r] = b.x + b.y + b.z + b.w;  // sequential fp32 sum, as tensor.sum(1) on 4 columns
    }
    __syncthreads();
    for (int r = threadIdx.x; r < L; r += blockDim.x) {
        const float k = a.key[r];
        int c = 0, f = -1;
        for (int r2 = 0; r2 < L; ++r2)
            if (a.key[r2] == k) { ++c; if (f < 0) f = r2; }
        a.cnt[r] = c;
        a.first[r] = f;
    }
    __syncthreads();
    for (int r = threadIdx.x; r < L; r += blockDim.x) {
        int g = 0;
        if (a.first[r] == r && a.cnt[r] > 1) {
            const float4 b0 = box[r];
            const float k = a.key[r];
            float s = 0.0f;
            for (int r2 = r; r2 < L; ++r2)
                if (a.key[r2] == k) {
                    const float4 b = box[r2];
                    s += b.x - b0.x; s += b.y - b0.y; s += b.z - b0.z; s += b.w - b0.w;
                }
            g = (s == 0.0f);
        }
        a.isgrp[r] = g;
    }
    __syncthreads();
    for (int r = threadIdx.x; r < L; r += blockDim.x) {
        const int f = a.first[r];
        a.single[r] = !(f >= 0 && a.cnt[r] > 1 && a.isgrp[f]);
        if (a.isgrp[r]) {
            int rank = 0;
            for (int r2 = 0; r2 < L; ++r2) rank += (a.isgrp[r2] && a.key[r2] < a.key[r]);
            a.glist[rank] = r;
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) *s_tmp = 0;
    __syncthreads();
    int mine = 0;
    for (int r = threadIdx.x; r < L; r += blockDim.x) mine += a.isgrp[r];
    if (mine) atomicAdd(s_tmp, mine);
    __syncthreads();
    const int ng = *s_tmp;
    __syncthreads();
    return ng;
}

// delete_duplicate_boxes(return_split=False) with the "first" policy: rows of `box` (length L)
// -> order list: singles in order, then one (the first) row per true group in ascending-key order.
__device__ int dedup_order(const float4* box, int L, const AbcArgs& a, int32_t* out, int* s_tmp) {
    const int ng = dup_groups(box, L, a, s_tmp);
    int n = compact_append(L, [&](int r) { return a.single[r] != 0; }, [&](int r) { return r; }, out, 0, a.cap, s_tmp);
    for (int g = threadIdx.x; g < ng; g += blockDim.x)
        if (n + g < a.cap) out[n + g] = a.glist[g];
    __syncthreads();
    return n + ng;
}

// Scratch placement: the kernel is a long chain of small dependent phases, so the latency of its scratch
// arrays is what it costs. With <= kAbcN detections per side the per-detection and per-row arrays live in
// shared memory, and so do the pair lists when the number of common pairs fits kAbcPairs; otherwise the
// global workspace is used (same code, pointers swapped).
constexpr int kAbcN = 128;
constexpr int kAbcPairs = 320;
constexpr size_t kAbcInputBytes = (size_t)kAbcN * (16 + 16 + 8 + 8 + 4 + 4);   // staged copies of both detection sets
constexpr size_t kAbcSmemBytes = (size_t)kAbcPairs * (6 * 4 + 8 * 4 + 16) + (size_t)kAbcN * 9 * 4 + kAbcInputBytes + 64;

__device__ __forceinline__ void abc_use_shared_small(AbcArgs& b, unsigned char* sm) {
    int32_t* w = reinterpret_cast<int32_t*>(sm + (size_t)kAbcPairs * 16);   // after the float4 block
    b.key = reinterpret_cast<float*>(w); w += kAbcPairs;
    b.first = w; w += kAbcPairs; b.cnt = w; w += kAbcPairs; b.isgrp = w; w += kAbcPairs;
    b.glist = w; w += kAbcPairs; b.single = w; w += kAbcPairs;
    b.uniq = w; w += kAbcN; b.offgl = w; w += kAbcN; b.on_used = w; w += kAbcN; b.off_matched = w; w += kAbcN;
    b.label = w; w += kAbcN; b.rowcnt = w; w += kAbcN; b.rowoff = w; w += kAbcN; b.g_i0 = w; w += kAbcN;
    b.g_m = w; w += kAbcN;
}
__device__ __forceinline__ void abc_use_shared_pairs(AbcArgs& b, unsigned char* sm) {
    b.mbox = reinterpret_cast<float4*>(sm);
    int32_t* w = reinterpret_cast<int32_t*>(sm + (size_t)kAbcPairs * 16) + 6 * kAbcPairs + 9 * kAbcN;
    b.pon[0] = w; w += kAbcPairs; b.pon[1] = w; w += kAbcPairs; b.poff[0] = w; w += kAbcPairs; b.poff[1] = w; w += kAbcPairs;
    b.flag_a = w; w += kAbcPairs; b.flag_b = w; w += kAbcPairs; b.raw = w; w += kAbcPairs; b.outlist = w; w += kAbcPairs;
}

__global__ void __launch_bounds__(256) match_abc_kernel(const AbcArgs a_in) {
    extern __shared__ __align__(16) unsigned char abc_smem[];
    __shared__ int s_tmp, s_flag[4];
    AbcArgs a = a_in;   // mutable copy: scratch pointers may be redirected to shared memory
    const int nc = a.nc, nd = a.nd_dev ? min(max(*a.nd_dev, 0), a.nd) : a.nd;
    const bool small = a.use_smem && nc <= kAbcN && nd <= kAbcN;
    if (small) {
        abc_use_shared_small(a, abc_smem);
        abc_use_shared_pairs(a, abc_smem);   // tentative: undone below if the pair count does not fit
        // stage both detection sets: every later phase re-reads them many times
        unsigned char* in = abc_smem + (size_t)kAbcPairs * (6 * 4 + 8 * 4 + 16) + (size_t)kAbcN * 9 * 4;
        float4* s_onb = reinterpret_cast<float4*>(in);
        float4* s_offb = s_onb + kAbcN;
        int64_t* s_oncls = reinterpret_cast<int64_t*>(s_offb + kAbcN);
        int64_t* s_offcls = s_oncls + kAbcN;
        float* s_ons = reinterpret_cast<float*>(s_offcls + kAbcN);
        float* s_offs = s_ons + kAbcN;
        for (int i = threadIdx.x; i < nc; i += blockDim.x) { s_onb[i] = a.onb[i]; s_oncls[i] = a.oncls[i]; s_ons[i] = a.ons[i]; }
        for (int j = threadIdx.x; j < nd; j += blockDim.x) { s_offb[j] = a.offb[j]; s_offcls[j] = a.offcls[j]; s_offs[j] = a.offs[j]; }
        a.onb = s_onb; a.offb = s_offb; a.oncls = s_oncls; a.offcls = s_offcls; a.ons = s_ons; a.offs = s_offs;
        __syncthreads();
    }
    int status = 0;
    // In the empty-side branches both members of a pair come from the same detection set.
    const bool on_empty = (nc == 0), off_empty = (nd == 0);
    const float4* ONB = on_empty ? a.offb : a.onb;
    const int64_t* ONC = on_empty ? a.offcls : a.oncls;
    const float* ONS = on_empty ? a.offs : a.ons;
    const float4* OFB = off_empty ? a.onb : a.offb;
    const int64_t* OFC = off_empty ? a.oncls : a.offcls;
    const float* OFS = off_empty ? a.ons : a.offs;

    int P = 0;        // common pairs
    int cur = 0;      // which pon/poff buffer is live
    int nC = 0;       // private rows written so far
    int nC_off = 0;   // ... of which CLIP-detector rows (they come first)

    if (on_empty && off_empty) {
        // nothing
    } else if (on_empty) {
        // trainer.py:349-355: confident CLIP-detector boxes are "common", the rest private
        P = compact_append(nd, [&](int j) { return a.offs[j] > 0.8f; }, [&](int j) { return j; }, a.pon[0], 0, a.cap, &s_tmp);
        for (int p = threadIdx.x; p < min(P, a.cap); p += blockDim.x) a.poff[0][p] = a.pon[0][p];
        nC = compact_append(nd, [&](int j) { return !(a.offs[j] > 0.8f); }, [&](int j) { return j; }, a.c_off, 0, nc + nd, &s_tmp);
        for (int r = threadIdx.x; r < nC; r += blockDim.x) a.c_on[r] = -1;
        nC_off = nC;
        __syncthreads();
    } else if (off_empty) {
        // trainer.py:356-361: every cloud box is "common" with itself
        for (int i = threadIdx.x; i < nc; i += blockDim.x) { a.pon[0][i] = i; a.poff[0][i] = i; }
        P = nc;
        __syncthreads();
    } else {
        // ---- A. split CLIP-detector detections into unique boxes and exact-duplicate groups (trainer.py:363)
        const int ng = dup_groups(a.offb, nd, a, &s_tmp);
        for (int j = threadIdx.x; j < nd; j += blockDim.x) {
            a.offgl[j] = a.single[j] ? -1 : a.first[j];
            a.off_matched[j] = 0;
        }
        for (int i = threadIdx.x; i < nc; i += blockDim.x) a.on_used[i] = 0;
        __syncthreads();
        const int nu = compact_append(nd, [&](int j) { return a.single[j] != 0; }, [&](int j) { return j; }, a.uniq, 0, nd, &s_tmp);
        // group leaders are kept in glist; copy them out because later dedups reuse the scratch
        for (int g = threadIdx.x; g < ng; g += blockDim.x) a.g_m[g] = a.glist[g];
        __syncthreads();

        // ---- B. all (cloud i, unique j) with IoU >= thr, row-major (trainer.py:364-368)
        for (int i = threadIdx.x; i < nc; i += blockDim.x) {
            const float4 bi = a.onb[i];
            const float ai = box_area(bi);
            int c = 0;
            for (int u = 0; u < nu; ++u) {
                const float4 bj = a.offb[a.uniq[u]];
                c += iou_d2(bi, ai, bj, box_area(bj)) >= a.thr;
            }
            a.rowcnt[i] = c;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            int acc = 0;
            for (int i = 0; i < nc; ++i) { a.rowoff[i] = acc; acc += a.rowcnt[i]; }
            s_tmp = acc;
        }
        __syncthreads();
        P = s_tmp;
        __syncthreads();
        if (small && P + nd > kAbcPairs) {   // too many pairs for shared memory: back to the global workspace
            a.pon[0] = a_in.pon[0]; a.pon[1] = a_in.pon[1]; a.poff[0] = a_in.poff[0]; a.poff[1] = a_in.poff[1];
            a.flag_a = a_in.flag_a; a.flag_b = a_in.flag_b; a.raw = a_in.raw; a.outlist = a_in.outlist; a.mbox = a_in.mbox;
            a.key = a_in.key; a.first = a_in.first; a.cnt = a_in.cnt; a.isgrp = a_in.isgrp; a.glist = a_in.glist;
            a.single = a_in.single;
        }
        if (P > a.cap) { status |= 4; }
        for (int i = threadIdx.x; i < nc; i += blockDim.x) {
            const float4 bi = a.onb[i];
            const float ai = box_area(bi);
            int at = a.rowoff[i];
            for (int u = 0; u < nu; ++u) {
                const int j = a.uniq[u];
                const float4 bj = a.offb[j];
                if (iou_d2(bi, ai, bj, box_area(bj)) >= a.thr) {
                    if (at < a.cap) { a.pon[0][at] = i; a.poff[0][at] = j; }
                    ++at;
                    a.on_used[i] = 1;
                    a.off_matched[j] = 1;
                }
            }
        }
        __syncthreads();
        P = min(P, a.cap);
        // unmatched unique CLIP-detector boxes -> private (trainer.py:369)
        nC = compact_append(nu, [&](int u) { return !a.off_matched[a.uniq[u]]; }, [&](int u) { return a.uniq[u]; },
                            a.c_off, 0, nc + nd, &s_tmp);

        // ---- C. duplicate groups (trainer.py:372-387): first matching cloud box, prefer the member
        //         with the same class, else the first member
        for (int g = threadIdx.x; g < ng; g += blockDim.x) {
            const int leader = a.g_m[g];
            int i0 = -1;
            for (int i = 0; i < nc && i0 < 0; ++i) {
                const float4 bi = a.onb[i];
                const float ai = box_area(bi);
                for (int j = leader; j < nd; ++j)
                    if (a.offgl[j] == leader) {
                        const float4 bj = a.offb[j];
                        if (iou_d2(bi, ai, bj, box_area(bj)) >= a.thr) { i0 = i; break; }
                    }
            }
            int pick = leader;
            if (i0 >= 0) {
                int same = 0, first_same = -1;
                for (int j = leader; j < nd; ++j)
                    if (a.offgl[j] == leader && a.offcls[j] == a.oncls[i0]) { if (first_same < 0) first_same = j; ++same; }
                if (same >= 1) pick = first_same;
                if (same > 1) atomicOr(&a.counts[3], 8);  // the reference would append several rows here
            }
            a.g_i0[g] = i0;
            a.flag_a[g] = pick;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            int p = P, c = nC;
            for (int g = 0; g < ng; ++g) {
                if (a.g_i0[g] >= 0) {
                    if (p < a.cap) { a.pon[0][p] = a.g_i0[g]; a.poff[0][p] = a.flag_a[g]; ++p; }
                    a.on_used[a.g_i0[g]] = 1;
                } else {
                    a.c_off[c++] = a.flag_a[g];
                }
            }
            s_flag[0] = p;
            s_flag[1] = c;
        }
        __syncthreads();
        P = s_flag[0];
        nC = s_flag[1];
        __syncthreads();
        for (int r = threadIdx.x; r < nC; r += blockDim.x) a.c_on[r] = -1;

        // ---- D. online_boxes_merging (util.py:484-507): cloud self-clusters at IoU >= 0.95
        for (int i = threadIdx.x; i < nc; i += blockDim.x) a.label[i] = i;
        __syncthreads();
        while (true) {  // min-label propagation over the >= 0.95 graph (transitive closure, util.py:459-482)
            int changed = 0;
            for (int i = threadIdx.x; i < nc; i += blockDim.x) {
                const float4 bi = a.onb[i];
                const float ai = box_area(bi);
                int best = a.label[i];
                for (int j = 0; j < nc; ++j) {
                    const float4 bj = a.onb[j];
                    if (iou_d2(bi, ai, bj, box_area(bj)) >= 0.95f) best = min(best, a.label[j]);
                }
                if (best < a.label[i]) { a.label[i] = best; changed = 1; }
            }
            if (!__syncthreads_or(changed)) break;
        }
        // size and "mixed classes" flag of every cluster, once, in parallel (rowcnt / rowoff are free by now)
        for (int i = threadIdx.x; i < nc; i += blockDim.x) { a.rowcnt[i] = 0; a.rowoff[i] = 0; }
        __syncthreads();
        for (int i = threadIdx.x; i < nc; i += blockDim.x) {
            const int r = a.label[i];
            atomicAdd(&a.rowcnt[r], 1);
            if (a.oncls[i] != a.oncls[r]) atomicOr(&a.rowoff[r], 1);
        }
        __syncthreads();
        for (int root = 0; root < nc; ++root) {  // clusters in ascending order of their lowest member
            if (a.label[root] != root) continue;
            const int size = a.rowcnt[root], mixed = a.rowoff[root];
            if (size < 2) continue;                     // uniform: every thread evaluates the same data
            if (!mixed) status |= 16;                   // the reference asserts here (util.py:488)
            const float4 broot = a.onb[root];
            // flags over the current common list
            int any_first = 0;
            for (int p = threadIdx.x; p < P; p += blockDim.x) {
                const float4 bp = a.onb[a.pon[cur][p]];
                int touched = 0;
                for (int i = root; i < nc; ++i)
                    if (a.label[i] == root && box_eq(a.onb[i], bp)) { touched = 1; break; }
                a.flag_a[p] = touched;
                const int wf = box_eq(broot, bp);
                a.flag_b[p] = wf;
                any_first |= wf;
            }
            any_first = __syncthreads_or(any_first);
            // classes the CLIP detector gave to the pairs of the first cluster box: unanimous?
            if (threadIdx.x == 0) { s_flag[0] = -1; s_flag[1] = 0; s_flag[2] = 0x7fffffff; }
            __syncthreads();
            if (any_first) {
                int mine = 0x7fffffff;   // lowest flagged pair of this thread
                for (int p = threadIdx.x; p < P; p += blockDim.x)
                    if (a.flag_b[p]) { mine = p; break; }
                if (mine != 0x7fffffff) atomicMin(&s_flag[2], mine);
                __syncthreads();
                const int cls = (int)a.offcls[a.poff[cur][s_flag[2]]];
                int multi = 0;
                for (int p = threadIdx.x; p < P; p += blockDim.x)
                    multi |= (a.flag_b[p] && (int)a.offcls[a.poff[cur][p]] != cls);
                multi = __syncthreads_or(multi);
                if (threadIdx.x == 0) { s_flag[0] = cls; s_flag[1] = multi; }
                __syncthreads();
            }
            const int ucls = s_flag[0];
            const bool unanimous = any_first && !s_flag[1];
            int agree_any = 0;
            if (unanimous) {
                for (int p = threadIdx.x; p < P; p += blockDim.x)
                    agree_any |= (a.flag_a[p] && (int)a.oncls[a.pon[cur][p]] == ucls);
            }
            agree_any = __syncthreads_or(agree_any);
            // keep rule for touched entries
            for (int p = threadIdx.x; p < P; p += blockDim.x) {
                if (!a.flag_a[p]) { a.flag_b[p] = 0; continue; }
                const int oc = (int)a.oncls[a.pon[cur][p]], fc = (int)a.offcls[a.poff[cur][p]];
                int keep;
                if (unanimous) keep = agree_any ? (oc == ucls) : 1;
                else keep = (oc != fc);
                a.flag_b[p] = keep;
            }
            __syncthreads();
            const int nxt = cur ^ 1;
            int q = compact_append(P, [&](int p) { return a.flag_a[p] == 0; }, [&](int p) { return p; }, a.outlist, 0, a.cap, &s_tmp);
            q = compact_append(P, [&](int p) { return a.flag_b[p] != 0; }, [&](int p) { return p; }, a.outlist, q, a.cap, &s_tmp);
            for (int r = threadIdx.x; r < q; r += blockDim.x) {
                a.pon[nxt][r] = a.pon[cur][a.outlist[r]];
                a.poff[nxt][r] = a.poff[cur][a.outlist[r]];
            }
            __syncthreads();
            cur = nxt;
            P = q;
        }

        // ---- E. cloud boxes never used in a pair -> private (trainer.py:391)
        const int before = nC;
        nC_off = nC;
        nC = compact_append(nc, [&](int i) { return !a.on_used[i]; }, [&](int i) { return i; }, a.c_on, nC, nc + nd, &s_tmp);
        for (int r = before + threadIdx.x; r < nC; r += blockDim.x) a.c_off[r] = -1;
        __syncthreads();
    }

    // ---- F. split the common pairs into A / B, merge boxes, de-duplicate (trainer.py:401-455)
    int nA = 0, nB = 0;
    for (int pass = 0; pass < 2; ++pass) {
        // pass 0 -> A (same class, or everything for 'RPN'); pass 1 -> B ('RCNN' only)
        if (pass == 1 && a.tag != COIN_TAG_RCNN) break;
        const int nraw = compact_append(
            P,
            [&](int p) {
                if (a.tag != COIN_TAG_RCNN) return true;
                const bool same = OFC[a.poff[cur][p]] == ONC[a.pon[cur][p]];
                return pass == 0 ? same : !same;
            },
            [&](int p) { return p; }, a.raw, 0, a.cap, &s_tmp);
        for (int r = threadIdx.x; r < nraw; r += blockDim.x) {
            const int p = a.raw[r];
            const float4 bo = ONB[a.pon[cur][p]];
            float4 m = bo;
            if (a.w_a != 1.0f) {  // weighted_box_fusion_split, nms.py:24-31
                const float4 bf = OFB[a.poff[cur][p]];
                const float so = ONS[a.pon[cur][p]], sf = OFS[a.poff[cur][p]];
                const float tot = so + sf;
                const float wo = so / tot, wf = sf / tot;
                m = make_float4(bo.x * wo + bf.x * wf, bo.y * wo + bf.y * wf, bo.z * wo + bf.z * wf, bo.w * wo + bf.w * wf);
            }
            a.mbox[r] = m;
        }
        __syncthreads();
        const int nout = dedup_order(a.mbox, nraw, a, a.outlist, &s_tmp);
        if (pass == 0) {
            for (int r = threadIdx.x; r < nout; r += blockDim.x) {
                const int p = a.raw[a.outlist[r]];
                a.a_on[r] = a.pon[cur][p];
                a.a_off[r] = a.poff[cur][p];
                a.a_box[r] = a.mbox[a.outlist[r]];
            }
            nA = nout;
            __syncthreads();
        } else {
            // drop B rows whose box equals an A box exactly (trainer.py:434-439)
            for (int r = threadIdx.x; r < nout; r += blockDim.x) {
                const float4 bb = a.mbox[a.outlist[r]];
                int clash = 0;
                for (int q = 0; q < nA; ++q) clash |= box_eq(bb, a.a_box[q]);
                a.flag_a[r] = !clash;
            }
            __syncthreads();
            nB = compact_append(nout, [&](int r) { return a.flag_a[r] != 0; }, [&](int r) { return a.outlist[r]; },
                                a.flag_b, 0, a.cap, &s_tmp);
            for (int r = threadIdx.x; r < nB; r += blockDim.x) {
                const int row = a.flag_b[r];
                const int p = a.raw[row];
                a.b_on[r] = a.pon[cur][p];
                a.b_off[r] = a.poff[cur][p];
                a.b_box[r] = a.mbox[row];
            }
            __syncthreads();
        }
    }
    if (threadIdx.x == 0) {
        a.counts[0] = nA;
        a.counts[1] = nB;
        a.counts[2] = nC;
        a.counts[4] = nC_off;
        if (status) atomicOr(&a.counts[3], status);
    }
}

static void carve_abc(AbcArgs& a, void* ws, int64_t nc, int64_t nd, int64_t cap, size_t* total) {
    Carver c(ws);
    const size_t L = (size_t)std::max<int64_t>(std::max(nd, cap), 1);
    const size_t snc = (size_t)std::max<int64_t>(nc, 1), snd = (size_t)std::max<int64_t>(nd, 1), scap = (size_t)std::max<int64_t>(cap, 1);
    a.key = c.take<float>(L);
    a.first = c.take<int32_t>(L); a.cnt = c.take<int32_t>(L); a.isgrp = c.take<int32_t>(L);
    a.glist = c.take<int32_t>(L); a.single = c.take<int32_t>(L);
    a.uniq = c.take<int32_t>(snd); a.offgl = c.take<int32_t>(snd);
    a.on_used = c.take<int32_t>(snc); a.off_matched = c.take<int32_t>(snd); a.label = c.take<int32_t>(snc);
    a.rowcnt = c.take<int32_t>(snc); a.rowoff = c.take<int32_t>(snc);
    for (int b = 0; b < 2; ++b) { a.pon[b] = c.take<int32_t>(scap); a.poff[b] = c.take<int32_t>(scap); }
    a.flag_a = c.take<int32_t>(std::max(scap, snd)); a.flag_b = c.take<int32_t>(scap);
    a.raw = c.take<int32_t>(scap);
    a.mbox = c.take<float4>(scap);
    a.g_i0 = c.take<int32_t>(snd); a.g_m = c.take<int32_t>(snd);
    a.outlist = c.take<int32_t>(scap);
    *total = c.used();
}

}  // namespace coin
using namespace coin;

extern "C" size_t coin_match_abc_workspace_bytes(int64_t nc, int64_t nd) {
    AbcArgs a;
    size_t total = 0;
    carve_abc(a, nullptr, nc, nd, nc * nd + nc + nd, &total);
    return total + 256;
}

static int match_abc_impl(const float* on_boxes, const int64_t* on_classes, const float* on_scores, int64_t nc,
                          const float* off_boxes, const int64_t* off_classes, const float* off_scores,
                          int64_t nd, const int32_t* nd_dev, int tag, float iou_thr, float weight_for_box_a,
                          int64_t cap_pairs, int32_t* a_on, int32_t* a_off, float* a_boxes, int32_t* b_on,
                          int32_t* b_off, float* b_boxes, int32_t* c_on, int32_t* c_off, int32_t* counts, void* ws,
                          size_t ws_bytes, coin_stream_t stream) {
    COIN_REQUIRE(nc >= 0 && nd >= 0 && counts, "match_abc: bad arguments");
    COIN_REQUIRE(nc <= COIN_ABC_MAX && nd <= COIN_ABC_MAX, "match_abc: at most %d detections per side", COIN_ABC_MAX);
    COIN_REQUIRE(tag == COIN_TAG_RCNN || tag == COIN_TAG_RPN, "match_abc: bad tag %d", tag);
    cudaStream_t s = as_stream(stream);
    cudaMemsetAsync(counts, 0, 8 * sizeof(int32_t), s);
    if (nc == 0 && nd == 0) return COIN_OK;
    COIN_REQUIRE(cap_pairs >= nc * nd + nc + nd, "match_abc: cap_pairs must be >= nc*nd + nc + nd");
    COIN_REQUIRE(a_on && a_off && a_boxes && c_on && c_off && ws, "match_abc: null pointer");
    COIN_REQUIRE(tag != COIN_TAG_RCNN || (b_on && b_off && b_boxes), "match_abc: B outputs are required for tag RCNN");
    COIN_REQUIRE(nc == 0 || (on_boxes && on_classes && on_scores), "match_abc: null cloud inputs");
    COIN_REQUIRE(nd == 0 || (off_boxes && off_classes && off_scores), "match_abc: null CLIP-detector inputs");
    AbcArgs a;
    size_t total = 0;
    carve_abc(a, ws, nc, nd, cap_pairs, &total);
    if (ws_bytes < total) return fail(COIN_ERR_CAPACITY, "match_abc: workspace too small (%zu < %zu)", ws_bytes, total);
    a.onb = reinterpret_cast<const float4*>(on_boxes); a.offb = reinterpret_cast<const float4*>(off_boxes);
    a.oncls = on_classes; a.offcls = off_classes; a.ons = on_scores; a.offs = off_scores;
    a.nc = (int)nc; a.nd = (int)nd; a.nd_dev = nd_dev; a.tag = tag; a.thr = iou_thr; a.w_a = weight_for_box_a; a.cap = (int)cap_pairs;
    a.a_on = a_on; a.a_off = a_off; a.b_on = b_on; a.b_off = b_off; a.c_on = c_on; a.c_off = c_off; a.counts = counts;
    a.a_box = reinterpret_cast<float4*>(a_boxes); a.b_box = reinterpret_cast<float4*>(b_boxes);
    a.use_smem = (nc <= kAbcN && nd <= kAbcN) ? 1 : 0;
    match_abc_kernel<<<1, 256, a.use_smem ? kAbcSmemBytes : 0, s>>>(a);
    return check_launch("match_abc_kernel");
}

extern "C" int coin_match_abc(const float* on_boxes, const int64_t* on_classes, const float* on_scores, int64_t nc,
                              const float* off_boxes, const int64_t* off_classes, const float* off_scores,
                              int64_t nd, int tag, float iou_thr, float weight_for_box_a, int64_t cap_pairs,
                              int32_t* a_on, int32_t* a_off, float* a_boxes, int32_t* b_on, int32_t* b_off,
                              float* b_boxes, int32_t* c_on, int32_t* c_off, int32_t* counts, void* ws,
                              size_t ws_bytes, coin_stream_t stream) {
    return match_abc_impl(on_boxes, on_classes, on_scores, nc, off_boxes, off_classes, off_scores, nd, nullptr, tag,
                          iou_thr, weight_for_box_a, cap_pairs, a_on, a_off, a_boxes, b_on, b_off, b_boxes, c_on, c_off,
                          counts, ws, ws_bytes, stream);
}

extern "C" int coin_match_abc_dev(const float* on_boxes, const int64_t* on_classes, const float* on_scores, int64_t nc,
                                  const float* off_boxes, const int64_t* off_classes, const float* off_scores,
                                  int64_t nd_cap, const int32_t* nd_dev, int tag, float iou_thr,
                                  float weight_for_box_a, int64_t cap_pairs, int32_t* a_on, int32_t* a_off,
                                  float* a_boxes, int32_t* b_on, int32_t* b_off, float* b_boxes, int32_t* c_on,
                                  int32_t* c_off, int32_t* counts, void* ws, size_t ws_bytes, coin_stream_t stream) {
    COIN_REQUIRE(nd_cap >= 1 && nd_dev, "match_abc_dev: nd_cap must be >= 1 and nd_dev non-null");
    return match_abc_impl(on_boxes, on_classes, on_scores, nc, off_boxes, off_classes, off_scores, nd_cap, nd_dev, tag,
                          iou_thr, weight_for_box_a, cap_pairs, a_on, a_off, a_boxes, b_on, b_off, b_boxes, c_on, c_off,
                          counts, ws, ws_bytes, stream);
}
