// match_abc.cu -- COIN's knowledge separation (consistent A / inconsistent B / private C) on device.
//
// Replaces CoinTrainer.match_dual_teacher, coin/engine/trainer.py:338-461, together with its helpers
// delete_duplicate_boxes (coin/utils/util.py:434-457), filter_result/find_same (:459-482),
// online_boxes_merging (:484-507) and merge_boxes (trainer.py:480-485 -> coin/layers/nms.py:24-31).
// The reference runs this on CPU tensors with Python sets, .tolist() and per-group loops, after a
// D2H copy of the teacher's detections (trainer.py:469) and before an H2D copy of the result
// (:457-459). Here ONE single-CTA launch per image serves both tags ('RCNN' and 'RPN' share everything
// up to the A/B split) and emits INDEX lists into the two detection sets plus the merged boxes; the
// field gathers are coin_abc_pack (step_dev.cu).
//
// The problem is tiny (<= ~100 detections per side) and is a chain of small dependent phases, so what
// it costs is latency, not bandwidth:
//   * every pairwise IoU is evaluated ONCE, by all warps (lane = column, one ballot per 32 pairs), into
//     two bit matrices in shared memory: cloud x CLIP-detector at the match threshold and cloud x cloud
//     at 0.95. Pair lists, "first matching cloud box of a duplicate group", matched / used flags and
//     the self-cluster graph are then popcounts and bit scans;
//   * both detection sets and every scratch array live in shared memory (<= kAbcN boxes per side and
//     <= kAbcPairs common pairs; beyond that the same code runs on the global workspace);
//   * the serial pieces (CPython set replays, appending the duplicate groups) run on single threads of
//     DIFFERENT warps between the same two barriers.
//
// Order. The reference turns Python sets into lists at trainer.py:369,391 and util.py:481, so the row
// order of C and the "first box" of a self-cluster (which decides A/B membership at util.py:497) are
// those of CPython's set table. pyset.cuh replays that table exactly (fuzzed against the interpreter in
// tests/test_pyset_cpu.py); the kernel therefore returns the reference's rows in the reference's order.
// If a replay exceeds its scratch (a self-cluster of dozens of boxes) the kernel falls back to ascending
// order / lowest member and raises status bit 32. random.randint picks (trainer.py:385,387; util.py:450)
// take the first element, i.e. the reference with randint pinned to its lower bound.
#include "common.cuh"
#include "pyset.cuh"

namespace coin {

struct AbcOut {            // one tag's outputs; a_on == nullptr: tag not requested
    int32_t *a_on, *a_off, *b_on, *b_off;
    float4 *a_box, *b_box;
    int32_t* counts;       // [8]: nA, nB, nC, status, nC_off, 0, 0, 0
};

struct AbcArgs {
    const float4 *onb, *offb;
    const int64_t *oncls64, *offcls64;
    const float *ons, *offs;
    int nc, nd, use_smem;
    const int32_t* nd_dev;   // optional device-side count of CLIP-detector detections (<= nd)
    float thr, w_a;
    int cap;
    AbcOut out[2];           // [0] = 'RCNN', [1] = 'RPN'
    int32_t *c_on, *c_off;
    // scratch (global workspace; redirected to shared memory when it fits)
    int32_t *oncls, *offcls;                          // [nc], [nd] classes as int32
    uint32_t *mbits, *abits;                          // [nc][Wd], [nc][Wc]
    uint32_t *uniqmask, *matched, *gmask;             // [Wd], [Wd], [nd][Wd] (row g = members of group g)
    float* key;                                       // [L]
    int32_t *first, *isgrp, *glist, *single;          // [L]
    int32_t *uniq, *urank, *ginv;                     // [nd]
    int32_t *on_used, *label;                         // [nc]
    int32_t *rowcnt, *rowoff;                         // [nc]
    int32_t *pon[2], *poff[2];                        // [cap]
    int32_t *flag_a, *flag_b;                         // [cap]
    int32_t *raw;                                     // [cap]
    float4* mbox;                                     // [cap]
    int32_t *g_i0, *g_pick;                           // [nd]
    int32_t *outlist;                                 // [cap]
    int32_t *ord_off, *ord_on;                        // [nd], [nc] set-difference orders
    int16_t *pool_cl, *pool_d0, *pool_d1;             // pyset pools
    pyset::Handle *sets, *clusters;                   // [nc], [kAbcClusters]
    pyset::Frame* frames;                             // [pyset::kMaxDepth]
};

constexpr int kAbcN = 128;          // boxes per side held in shared memory
constexpr int kAbcPairs = 320;      // common pairs held in shared memory
constexpr int kAbcClusters = 64;    // self-clusters replayed in set order
constexpr int kPoolCl = 4096;       // int16 slots for the self-cluster set replay
constexpr int kPoolDiff = 8 + 32 + 128 + 512;

__device__ __forceinline__ bool box_eq(const float4& a, const float4& b) {
    return a.x == b.x && a.y == b.y && a.z == b.z && a.w == b.w;
}

// Ordered compaction of {r in [0,L) : pred(r)} appended to out[count...]; executed by warp 0 of the
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

// list(set(range(n)) - other), warp-cooperative; keep(i) = "i is not in other" (so len(other) = n - #kept). The common
// case (the survivors sit in their own slots of the CPython table: ascending) is a ballot compaction; the rare remainder
// (a few large survivors in a small table) is replayed by lane 0. Every lane must call; returns the count.
template <class Keep>
__device__ int difference_order_warp(int n, Keep keep, int32_t* out, int16_t* pool_base, int* pool_ctr) {
    const int lane = threadIdx.x & 31;
    int m = 0, mx = -1;
    for (int base = 0; base < n; base += 32) {
        const int i = base + lane;
        const unsigned b = __ballot_sync(0xffffffffu, i < n && keep(i));
        m += __popc(b);
        if (b) mx = base + 31 - __clz(b);
    }
    const int other_size = n - m;
    if (((n >> 2) > other_size) || mx < pyset::table_size_after_adds(m)) {
        int c = 0;
        for (int base = 0; base < n; base += 32) {
            const int i = base + lane;
            const bool p = i < n && keep(i);
            const unsigned b = __ballot_sync(0xffffffffu, p);
            if (p) out[c + __popc(b & ((1u << lane) - 1u))] = i;
            c += __popc(b);
        }
        return c;
    }
    int c = 0;
    if (lane == 0) {
        pyset::Pool pool{pool_base, kPoolDiff, pool_ctr, pool_ctr + 1};
        c = pyset::difference_order(n, keep, other_size, out, pool);
    }
    return __shfl_sync(0xffffffffu, c, 0);
}

// util.py:434-457 grouping, one pass per row: rows sharing the fp32 sum of their coordinates form a candidate group;
// it is a true duplicate group when the summed difference to its first row is exactly zero (the reference's test).
// After the call: single[r] = 1 if row r is not part of a true group; isgrp[r] = 1 on the leader (lowest row) of a true
// group; first[r] = leader candidate of r's key group; glist[0..ng) = leaders in ascending-key order. Returns ng.
__device__ int dup_groups(const float4* box, int L, const AbcArgs& a, int* s_tmp) {
    for (int r = threadIdx.x; r < L; r += blockDim.x) {
        const float4 b = box[r];
        a.key[r] = b.x + b.y + b.z + b.w;  // sequential fp32 sum, as tensor.sum(1) on 4 columns
    }
    if (threadIdx.x == 0) *s_tmp = 0;
    __syncthreads();
    for (int r = threadIdx.x; r < L; r += blockDim.x) {
        const float k = a.key[r];
        int c = 0, f = -1;
        float4 b0 = box[r];
        float s = 0.0f;
        for (int r2 = 0; r2 < L; ++r2)
            if (a.key[r2] == k) {
                const float4 b = box[r2];
                if (f < 0) { f = r2; b0 = b; }
                ++c;
                s += b.x - b0.x; s += b.y - b0.y; s += b.z - b0.z; s += b.w - b0.w;
            }
        const int grp = (c > 1 && s == 0.0f);
        a.first[r] = f;
        a.single[r] = !grp;
        a.isgrp[r] = grp && f == r;
    }
    __syncthreads();
    int mine = 0;
    for (int r = threadIdx.x; r < L; r += blockDim.x)
        if (a.isgrp[r]) {
            int rank = 0;
            for (int r2 = 0; r2 < L; ++r2) rank += (a.isgrp[r2] && a.key[r2] < a.key[r]);
            a.glist[rank] = r;
            ++mine;
        }
    if (mine) atomicAdd(s_tmp, mine);
    __syncthreads();
    const int ng = *s_tmp;
    __syncthreads();
    return ng;
}

// delete_duplicate_boxes(return_split=False), first-member policy: singles in order, then the first row of every true
// group in ascending-key order.
__device__ int dedup_order(const float4* box, int L, const AbcArgs& a, int32_t* out, int* s_tmp) {
    const int ng = dup_groups(box, L, a, s_tmp);
    int n = compact_append(L, [&](int r) { return a.single[r] != 0; }, [&](int r) { return r; }, out, 0, a.cap, s_tmp);
    for (int g = threadIdx.x; g < ng; g += blockDim.x)
        if (n + g < a.cap) out[n + g] = a.glist[g];
    __syncthreads();
    return n + ng;
}

// shared-memory layout (small mode)
constexpr int kWN = kAbcN / 32;
constexpr size_t kSmA = (size_t)kAbcPairs * 16;                                  // mbox
constexpr size_t kSmB = kSmA + (size_t)2 * kAbcN * 16;                           // staged boxes
constexpr size_t kSmC = kSmB + (size_t)kAbcN * 4 * 4;                            // cls (int32) + scores, both sides
constexpr size_t kSmD = kSmC + (size_t)2 * kAbcN * kWN * 4;                      // mbits, abits
constexpr size_t kSmE = kSmD + (size_t)(2 * kWN + kAbcN * kWN) * 4;              // uniqmask, matched, gmask
constexpr size_t kSmF = kSmE + (size_t)kAbcPairs * 5 * 4;                        // key first isgrp glist single
constexpr size_t kSmG = kSmF + (size_t)kAbcN * 11 * 4;                           // per-box int arrays
constexpr size_t kSmH = kSmG + (size_t)kAbcPairs * 8 * 4;                        // pair arrays
constexpr size_t kSmI = kSmH + (size_t)(kPoolCl + 2 * kPoolDiff) * 2;            // pyset pools
constexpr size_t kSmJ = kSmI + (size_t)(kAbcN + kAbcClusters) * sizeof(pyset::Handle) + (size_t)pyset::kMaxDepth * sizeof(pyset::Frame);
constexpr size_t kAbcSmemBytes = kSmJ + 64;

__device__ __forceinline__ void abc_use_shared(AbcArgs& b, unsigned char* sm, bool pairs_too) {
    int32_t* w;
    w = reinterpret_cast<int32_t*>(sm + kSmB);
    b.oncls = w; w += kAbcN; b.offcls = w; w += kAbcN;       // (scores follow, set by the caller)
    uint32_t* u = reinterpret_cast<uint32_t*>(sm + kSmC);
    b.mbits = u; u += kAbcN * kWN; b.abits = u;
    u = reinterpret_cast<uint32_t*>(sm + kSmD);
    b.uniqmask = u; u += kWN; b.matched = u; u += kWN; b.gmask = u;
    w = reinterpret_cast<int32_t*>(sm + kSmF);
    b.uniq = w; w += kAbcN; b.urank = w; w += kAbcN; b.ginv = w; w += kAbcN; b.on_used = w; w += kAbcN;
    b.label = w; w += kAbcN; b.rowcnt = w; w += kAbcN; b.rowoff = w; w += kAbcN; b.g_i0 = w; w += kAbcN;
    b.g_pick = w; w += kAbcN; b.ord_off = w; w += kAbcN; b.ord_on = w;
    int16_t* h = reinterpret_cast<int16_t*>(sm + kSmH);
    b.pool_cl = h; h += kPoolCl; b.pool_d0 = h; h += kPoolDiff; b.pool_d1 = h;
    pyset::Handle* hd = reinterpret_cast<pyset::Handle*>(sm + kSmI);
    b.sets = hd; hd += kAbcN; b.clusters = hd; hd += kAbcClusters;
    b.frames = reinterpret_cast<pyset::Frame*>(hd);
    if (pairs_too) {
        b.mbox = reinterpret_cast<float4*>(sm);
        w = reinterpret_cast<int32_t*>(sm + kSmE);
        b.key = reinterpret_cast<float*>(w); w += kAbcPairs;
        b.first = w; w += kAbcPairs; b.isgrp = w; w += kAbcPairs; b.glist = w; w += kAbcPairs; b.single = w;
        w = reinterpret_cast<int32_t*>(sm + kSmG);
        b.pon[0] = w; w += kAbcPairs; b.pon[1] = w; w += kAbcPairs; b.poff[0] = w; w += kAbcPairs; b.poff[1] = w; w += kAbcPairs;
        b.flag_a = w; w += kAbcPairs; b.flag_b = w; w += kAbcPairs; b.raw = w; w += kAbcPairs; b.outlist = w;
    }
}

// <= 56 registers x 256 threads = 14336: the CTA fits into the register slot ONE retiring ROIAlign CTA frees (224 threads
// x 72 registers = 16128), so inside the step it becomes resident at once instead of waiting for two slots on one SM.
__global__ void __maxnreg__(56) match_abc_kernel(const AbcArgs a_in) {
    extern __shared__ __align__(16) unsigned char abc_smem[];
    __shared__ int s_tmp, s_flag[6], s_status, s_pc[6];   // s_pc: (used, overflow) of the three pyset pools
    AbcArgs a = a_in;   // mutable copy: scratch pointers may be redirected to shared memory
    const int nc = a.nc, nd = a.nd_dev ? min(max(*a.nd_dev, 0), a.nd) : a.nd;
    const bool small = a.use_smem && nc <= kAbcN && nd <= kAbcN;
    const int Wd = (nd + 31) >> 5, Wc = (nc + 31) >> 5;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    if (small) {
        abc_use_shared(a, abc_smem, true);   // pair arrays tentatively too: undone below if the pair count does not fit
        float4* s_onb = reinterpret_cast<float4*>(abc_smem + kSmA);
        float4* s_offb = s_onb + kAbcN;
        float* s_ons = reinterpret_cast<float*>(abc_smem + kSmB) + 2 * kAbcN;
        float* s_offs = s_ons + kAbcN;
        for (int i = threadIdx.x; i < nc; i += blockDim.x) { s_onb[i] = a.onb[i]; s_ons[i] = a.ons[i]; }
        for (int j = threadIdx.x; j < nd; j += blockDim.x) { s_offb[j] = a.offb[j]; s_offs[j] = a.offs[j]; }
        a.onb = s_onb; a.offb = s_offb; a.ons = s_ons; a.offs = s_offs;
    }
    for (int i = threadIdx.x; i < nc; i += blockDim.x) a.oncls[i] = (int32_t)a.oncls64[i];
    for (int j = threadIdx.x; j < nd; j += blockDim.x) a.offcls[j] = (int32_t)a.offcls64[j];
    if (threadIdx.x == 0) s_status = 0;
    if (threadIdx.x < 6) s_pc[threadIdx.x] = 0;
    __syncthreads();
    int status = 0;
    // In the empty-side branches both members of a pair come from the same detection set.
    const bool on_empty = (nc == 0), off_empty = (nd == 0);
    const float4* ONB = on_empty ? a.offb : a.onb;
    const int32_t* ONC = on_empty ? a.offcls : a.oncls;
    const float* ONS = on_empty ? a.offs : a.ons;
    const float4* OFB = off_empty ? a.onb : a.offb;
    const int32_t* OFC = off_empty ? a.oncls : a.offcls;
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

        // ---- B. every pairwise IoU once: bit (i, j) of mbits = IoU(cloud i, CLIP-detector j) >= thr (trainer.py:364-365),
        //         of abits = IoU(cloud i, cloud j) >= 0.95 (util.py:468). lane = column, one ballot per 32 pairs.
        for (int w = 0; w < Wd; ++w) {
            const int j = w * 32 + lane;
            const float4 bj = j < nd ? a.offb[j] : make_float4(0.f, 0.f, 0.f, 0.f);
            const float aj = box_area(bj);
            for (int i = warp; i < nc; i += nwarps) {
                const float4 bi = a.onb[i];
                const unsigned m = __ballot_sync(0xffffffffu, j < nd && iou_d2(bi, box_area(bi), bj, aj) >= a.thr);
                if (lane == 0) a.mbits[i * Wd + w] = m;
            }
        }
        for (int w = 0; w < Wc; ++w) {
            const int j = w * 32 + lane;
            const float4 bj = j < nc ? a.onb[j] : make_float4(0.f, 0.f, 0.f, 0.f);
            const float aj = box_area(bj);
            for (int i = warp; i < nc; i += nwarps) {
                const float4 bi = a.onb[i];
                const unsigned m = __ballot_sync(0xffffffffu, j < nc && iou_d2(bi, box_area(bi), bj, aj) >= 0.95f);
                if (lane == 0) a.abits[i * Wc + w] = m;
            }
        }
        // masks of the unique CLIP-detector boxes, members of every duplicate group, rank of a unique box among them
        for (int w = warp; w < Wd; w += nwarps) {
            const int j = w * 32 + lane;
            const unsigned m = __ballot_sync(0xffffffffu, j < nd && a.single[j]);
            if (lane == 0) { a.uniqmask[w] = m; a.matched[w] = 0u; }
        }
        for (int t = threadIdx.x; t < ng * Wd; t += blockDim.x) a.gmask[t] = 0u;
        // the leaders move to g_pick: glist belongs to the de-duplication scratch, which is reused (and may move) below
        for (int g = threadIdx.x; g < ng; g += blockDim.x) { a.ginv[a.glist[g]] = g; a.g_pick[g] = a.glist[g]; }
        __syncthreads();
        for (int j = threadIdx.x; j < nd; j += blockDim.x) {
            if (a.single[j]) {
                int r = __popc(a.uniqmask[j >> 5] & ((1u << (j & 31)) - 1u));
                for (int w = 0; w < (j >> 5); ++w) r += __popc(a.uniqmask[w]);
                a.urank[j] = r;
                a.uniq[r] = j;
            } else {
                atomicOr(&a.gmask[a.ginv[a.first[j]] * Wd + (j >> 5)], 1u << (j & 31));
            }
        }
        // ---- C. all (cloud i, unique j) pairs, row-major (trainer.py:366-368)
        for (int i = threadIdx.x; i < nc; i += blockDim.x) {
            int c = 0;
            for (int w = 0; w < Wd; ++w) {
                const unsigned m = a.mbits[i * Wd + w] & a.uniqmask[w];
                c += __popc(m);
                if (m) atomicOr(&a.matched[w], m);
            }
            a.rowcnt[i] = c;
            a.on_used[i] = c > 0;
        }
        __syncthreads();
        if (warp == 0) {      // exclusive scan of rowcnt
            int carry = 0;
            for (int base = 0; base < nc; base += 32) {
                const int i = base + lane;
                const int v = i < nc ? a.rowcnt[i] : 0;
                int x = v;
                for (int d = 1; d < 32; d <<= 1) { const int y = __shfl_up_sync(0xffffffffu, x, d); if (lane >= d) x += y; }
                if (i < nc) a.rowoff[i] = carry + x - v;
                carry += __shfl_sync(0xffffffffu, x, 31);
            }
            if (lane == 0) s_tmp = carry;
        }
        int nu = 0;
        for (int w = 0; w < Wd; ++w) nu += __popc(a.uniqmask[w]);
        __syncthreads();
        P = s_tmp;
        __syncthreads();
        if (small && P + nd > kAbcPairs) {   // too many pairs for shared memory: pair arrays back to the global workspace
            a.pon[0] = a_in.pon[0]; a.pon[1] = a_in.pon[1]; a.poff[0] = a_in.poff[0]; a.poff[1] = a_in.poff[1];
            a.flag_a = a_in.flag_a; a.flag_b = a_in.flag_b; a.raw = a_in.raw; a.outlist = a_in.outlist; a.mbox = a_in.mbox;
            a.key = a_in.key; a.first = a_in.first; a.isgrp = a_in.isgrp; a.glist = a_in.glist; a.single = a_in.single;
        }
        if (P > a.cap) { status |= 4; }
        for (int i = threadIdx.x; i < nc; i += blockDim.x) {
            int at = a.rowoff[i];
            for (int w = 0; w < Wd; ++w) {
                unsigned m = a.mbits[i * Wd + w] & a.uniqmask[w];
                while (m) {
                    const int b = __ffs(m) - 1;
                    m &= m - 1;
                    if (at < a.cap) { a.pon[0][at] = i; a.poff[0][at] = w * 32 + b; }
                    ++at;
                }
            }
        }
        // ---- D. duplicate groups (trainer.py:372-387): first cloud box matching any member; prefer the member with the
        //         same class, else the first member. (the group pairs are appended, in order, by one thread below)
        for (int g = threadIdx.x; g < ng; g += blockDim.x) {
            const int leader = a.g_pick[g];
            int i0 = -1;
            for (int i = 0; i < nc && i0 < 0; ++i)
                for (int w = 0; w < Wd; ++w)
                    if (a.mbits[i * Wd + w] & a.gmask[g * Wd + w]) { i0 = i; break; }
            int pick = leader;
            if (i0 >= 0) {
                int same = 0, first_same = -1;
                for (int w = 0; w < Wd; ++w) {
                    unsigned m = a.gmask[g * Wd + w];
                    while (m) {
                        const int j = w * 32 + __ffs(m) - 1;
                        m &= m - 1;
                        if (a.offcls[j] == a.oncls[i0]) { if (first_same < 0) first_same = j; ++same; }
                    }
                }
                if (same >= 1) pick = first_same;
                if (same > 1) status |= 8;  // the reference appends several rows here and fails (trainer.py:402)
            }
            a.g_i0[g] = i0;
            a.g_pick[g] = pick;
        }
        __syncthreads();
        P = min(P, a.cap);
        // ---- E. order-carrying pieces (trainer.py:369,391; util.py:459-482), spread over the warps:
        //   threads 0..nc-1 classify the nodes of the 0.95 graph and build the CPython sets that matter (pyset.cuh);
        //   one thread of an otherwise idle warp appends the group pairs / unmatched groups, in group order
        auto adjw = [&](int i, int w) { return a.abits[i * Wc + w]; };
        pyset::Pool cpool{a.pool_cl, kPoolCl, &s_pc[0], &s_pc[1]};
        for (int i = threadIdx.x; i < nc; i += blockDim.x) {
            const int kind = pyset::classify_node(i, Wc, adjw);
            a.label[i] = kind;
            if (kind == pyset::kPairLeader || kind == pyset::kActive) a.sets[i] = pyset::build_row_set(i, Wc, adjw, cpool);
        }
        if (threadIdx.x == blockDim.x - 32) {
            int p = P, extra = 0;
            for (int g = 0; g < ng; ++g) {
                if (a.g_i0[g] >= 0) {
                    if (p < a.cap) { a.pon[0][p] = a.g_i0[g]; a.poff[0][p] = a.g_pick[g]; ++p; }
                    a.on_used[a.g_i0[g]] = 1;
                } else {
                    a.g_i0[g] = -2 - extra;     // position among the unmatched groups
                    ++extra;
                }
            }
            s_flag[0] = p;
            s_flag[1] = extra;
        }
        __syncthreads();
        //   warp 0: list(set(range(nc)) - used cloud boxes); warp 1: list(set(range(nu)) - matched unique boxes);
        //   warp 2: the nodes that are neither isolated nor a plain pair are replayed serially (rare: chains, triples)
        if (warp == 0) {
            const int c = difference_order_warp(nc, [&](int i) { return a.on_used[i] == 0; }, a.ord_on, a.pool_d0, &s_pc[2]);
            if (lane == 0) s_flag[3] = c;
        } else if (warp == 1) {
            const int c = difference_order_warp(
                nu, [&](int u) { const int j = a.uniq[u]; return ((a.matched[j >> 5] >> (j & 31)) & 1u) == 0u; }, a.ord_off,
                a.pool_d1, &s_pc[4]);
            if (lane == 0) s_flag[2] = c;
        } else if (warp == 2) {
            int16_t* act = reinterpret_cast<int16_t*>(a.rowoff);      // rowoff is free by now
            int na = 0;
            for (int base = 0; base < nc; base += 32) {
                const int i = base + lane;
                const bool p = i < nc && a.label[i] == pyset::kActive;
                const unsigned b = __ballot_sync(0xffffffffu, p);
                if (p) act[na + __popc(b & ((1u << lane) - 1u))] = (int16_t)i;
                na += __popc(b);
            }
            __syncwarp();
            int rc = 0;
            if (lane == 0 && na > 0 && !s_pc[1]) rc = pyset::filter_clusters_built(na, act, a.sets, a.frames, cpool);
            if (lane == 0) s_flag[4] = (rc < 0 || s_pc[1]) ? -1 : 0;
        }
        __syncthreads();
        // clusters in ascending order of their owner node: pair leaders and the replayed nodes that kept a set of len > 1
        int ncl = s_flag[4];
        if (ncl == 0) {
            ncl = compact_append(nc,
                                 [&](int i) {
                                     const int kind = a.label[i];
                                     return kind == pyset::kPairLeader ||
                                            (kind == pyset::kActive && a.sets[i].off >= 0 && a.sets[i].used != 1);
                                 },
                                 [&](int i) { return i; }, a.rowcnt, 0, nc, &s_tmp);
            if (ncl > kAbcClusters) ncl = -1;
            for (int k = threadIdx.x; k < ncl; k += blockDim.x) a.clusters[k] = a.sets[a.rowcnt[k]];
            __syncthreads();
        }
        P = s_flag[0];
        const int n_unmatched_groups = s_flag[1], n_off_only = s_flag[2], n_on_only = s_flag[3];
        // private rows: unmatched unique CLIP-detector boxes (set order), one box per unmatched group (group order), ...
        for (int r = threadIdx.x; r < n_off_only; r += blockDim.x) { a.c_off[r] = a.uniq[a.ord_off[r]]; a.c_on[r] = -1; }
        for (int g = threadIdx.x; g < ng; g += blockDim.x)
            if (a.g_i0[g] <= -2) { const int r = n_off_only + (-2 - a.g_i0[g]); a.c_off[r] = a.g_pick[g]; a.c_on[r] = -1; }
        nC_off = n_off_only + n_unmatched_groups;
        // ... then the cloud boxes never used in a pair (trainer.py:391)
        for (int r = threadIdx.x; r < n_on_only; r += blockDim.x) { a.c_on[nC_off + r] = a.ord_on[r]; a.c_off[nC_off + r] = -1; }
        nC = nC_off + n_on_only;

        // ---- F. online_boxes_merging (util.py:484-507) over the self-clusters
        if (ncl < 0) {
            // the set replay ran out of scratch: components by min-label propagation, lowest member first
            status |= 32;
            for (int i = threadIdx.x; i < nc; i += blockDim.x) a.label[i] = i;
            __syncthreads();
            while (true) {
                int changed = 0;
                for (int i = threadIdx.x; i < nc; i += blockDim.x) {
                    int best = a.label[i];
                    for (int w = 0; w < Wc; ++w) {
                        unsigned m = a.abits[i * Wc + w];
                        while (m) { const int j = w * 32 + __ffs(m) - 1; m &= m - 1; best = min(best, a.label[j]); }
                    }
                    if (best < a.label[i]) { a.label[i] = best; changed = 1; }
                }
                if (!__syncthreads_or(changed)) break;
            }
        }
        const int nloop = ncl < 0 ? nc : ncl;
        for (int k = 0; k < nloop; ++k) {
            int root = -1;              // first member of the cluster
            pyset::Handle h{};
            if (ncl >= 0) {
                h = a.clusters[k];
                for (int sl = 0; sl <= h.mask && root < 0; ++sl) { const int v = a.pool_cl[h.off + sl]; if (v != pyset::kEmpty) root = v; }
            } else {
                if (a.label[k] != k) continue;
                int size = 0;
                for (int i = k; i < nc; ++i) size += a.label[i] == k;
                if (size < 2) continue;
                root = k;
            }
            auto member = [&](int idx, int& out_i) -> bool {    // idx-th candidate slot / index -> member?
                if (ncl >= 0) { const int v = a.pool_cl[h.off + idx]; out_i = v; return v != pyset::kEmpty; }
                out_i = idx;
                return a.label[idx] == k;
            };
            const int nslots = ncl >= 0 ? h.mask + 1 : nc;
            // the reference asserts that a cluster holds more than one class (util.py:488)
            {
                int mixed = 0, mi;
                for (int s = 0; s < nslots; ++s)
                    if (member(s, mi) && a.oncls[mi] != a.oncls[root]) { mixed = 1; break; }
                if (!mixed) status |= 16;
            }
            const float4 broot = a.onb[root];
            // flags over the current common list: pair touched by the cluster / by its first box
            int any_first = 0;
            for (int p = threadIdx.x; p < P; p += blockDim.x) {
                const float4 bp = a.onb[a.pon[cur][p]];
                int touched = 0, mi;
                for (int s = 0; s < nslots; ++s)
                    if (member(s, mi) && box_eq(a.onb[mi], bp)) { touched = 1; break; }
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
                const int cls = a.offcls[a.poff[cur][s_flag[2]]];
                int multi = 0;
                for (int p = threadIdx.x; p < P; p += blockDim.x)
                    multi |= (a.flag_b[p] && a.offcls[a.poff[cur][p]] != cls);
                multi = __syncthreads_or(multi);
                if (threadIdx.x == 0) { s_flag[0] = cls; s_flag[1] = multi; }
                __syncthreads();
            }
            const int ucls = s_flag[0];
            const bool unanimous = any_first && !s_flag[1];
            int agree_any = 0;
            if (unanimous) {
                for (int p = threadIdx.x; p < P; p += blockDim.x)
                    agree_any |= (a.flag_a[p] && a.oncls[a.pon[cur][p]] == ucls);
            }
            agree_any = __syncthreads_or(agree_any);
            // keep rule for touched entries
            for (int p = threadIdx.x; p < P; p += blockDim.x) {
                if (!a.flag_a[p]) { a.flag_b[p] = 0; continue; }
                const int oc = a.oncls[a.pon[cur][p]], fc = a.offcls[a.poff[cur][p]];
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
        __syncthreads();
    }

    // ---- G. split the common pairs into A / B, merge boxes, de-duplicate (trainer.py:401-455), per requested tag
    for (int tag = 0; tag < 2; ++tag) {
        const AbcOut& o = a.out[tag];
        if (!o.a_on) continue;
        int nA = 0, nB = 0;
        for (int pass = 0; pass < 2; ++pass) {
            // pass 0 -> A (same class, or everything for 'RPN'); pass 1 -> B ('RCNN' only)
            if (pass == 1 && tag != COIN_TAG_RCNN) break;
            const int nraw = compact_append(
                P,
                [&](int p) {
                    if (tag != COIN_TAG_RCNN) return true;
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
                    o.a_on[r] = a.pon[cur][p];
                    o.a_off[r] = a.poff[cur][p];
                    o.a_box[r] = a.mbox[a.outlist[r]];
                }
                nA = nout;
                __syncthreads();
            } else {
                // drop B rows whose box equals an A box exactly (trainer.py:434-439)
                for (int r = threadIdx.x; r < nout; r += blockDim.x) {
                    const float4 bb = a.mbox[a.outlist[r]];
                    int clash = 0;
                    for (int q = 0; q < nA; ++q) clash |= box_eq(bb, o.a_box[q]);
                    a.flag_a[r] = !clash;
                }
                __syncthreads();
                nB = compact_append(nout, [&](int r) { return a.flag_a[r] != 0; }, [&](int r) { return a.outlist[r]; },
                                    a.flag_b, 0, a.cap, &s_tmp);
                for (int r = threadIdx.x; r < nB; r += blockDim.x) {
                    const int row = a.flag_b[r];
                    const int p = a.raw[row];
                    o.b_on[r] = a.pon[cur][p];
                    o.b_off[r] = a.poff[cur][p];
                    o.b_box[r] = a.mbox[row];
                }
                __syncthreads();
            }
        }
        if (status) atomicOr(&s_status, status);
        __syncthreads();
        if (threadIdx.x == 0) {
            o.counts[0] = nA;
            o.counts[1] = nB;
            o.counts[2] = nC;
            o.counts[3] = s_status;
            o.counts[4] = nC_off;
            o.counts[5] = 0; o.counts[6] = 0; o.counts[7] = 0;
        }
    }
}

static void carve_abc(AbcArgs& a, void* ws, int64_t nc, int64_t nd, int64_t cap, size_t* total) {
    Carver c(ws);
    const size_t L = (size_t)std::max<int64_t>(std::max(nd, cap), 1);
    const size_t snc = (size_t)std::max<int64_t>(nc, 1), snd = (size_t)std::max<int64_t>(nd, 1), scap = (size_t)std::max<int64_t>(cap, 1);
    const size_t wd = (snd + 31) / 32, wc = (snc + 31) / 32;
    a.oncls = c.take<int32_t>(snc); a.offcls = c.take<int32_t>(snd);
    a.mbits = c.take<uint32_t>(snc * wd); a.abits = c.take<uint32_t>(snc * wc);
    a.uniqmask = c.take<uint32_t>(wd); a.matched = c.take<uint32_t>(wd); a.gmask = c.take<uint32_t>(snd * wd);
    a.key = c.take<float>(L);
    a.first = c.take<int32_t>(L); a.isgrp = c.take<int32_t>(L);
    a.glist = c.take<int32_t>(L); a.single = c.take<int32_t>(L);
    a.uniq = c.take<int32_t>(snd); a.urank = c.take<int32_t>(snd); a.ginv = c.take<int32_t>(snd);
    a.on_used = c.take<int32_t>(snc); a.label = c.take<int32_t>(snc);
    a.rowcnt = c.take<int32_t>(snc); a.rowoff = c.take<int32_t>(snc);
    for (int b = 0; b < 2; ++b) { a.pon[b] = c.take<int32_t>(scap); a.poff[b] = c.take<int32_t>(scap); }
    a.flag_a = c.take<int32_t>(std::max(scap, snd)); a.flag_b = c.take<int32_t>(scap);
    a.raw = c.take<int32_t>(scap);
    a.mbox = c.take<float4>(scap);
    a.g_i0 = c.take<int32_t>(snd); a.g_pick = c.take<int32_t>(snd);
    a.outlist = c.take<int32_t>(scap);
    a.ord_off = c.take<int32_t>(snd); a.ord_on = c.take<int32_t>(snc);
    a.pool_cl = c.take<int16_t>(kPoolCl); a.pool_d0 = c.take<int16_t>(kPoolDiff); a.pool_d1 = c.take<int16_t>(kPoolDiff);
    a.sets = c.take<pyset::Handle>(snc); a.clusters = c.take<pyset::Handle>(kAbcClusters);
    a.frames = c.take<pyset::Frame>(pyset::kMaxDepth);
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

struct AbcHostOut {
    int32_t *a_on, *a_off;
    float* a_boxes;
    int32_t *b_on, *b_off;
    float* b_boxes;
    int32_t* counts;
};

static int match_abc_impl(const float* on_boxes, const int64_t* on_classes, const float* on_scores, int64_t nc,
                          const float* off_boxes, const int64_t* off_classes, const float* off_scores,
                          int64_t nd, const int32_t* nd_dev, float iou_thr, float weight_for_box_a, int64_t cap_pairs,
                          const AbcHostOut* rcnn, const AbcHostOut* rpn, int32_t* c_on, int32_t* c_off, void* ws,
                          size_t ws_bytes, coin_stream_t stream) {
    COIN_REQUIRE(nc >= 0 && nd >= 0 && (rcnn || rpn), "match_abc: bad arguments");
    COIN_REQUIRE(nc <= COIN_ABC_MAX && nd <= COIN_ABC_MAX, "match_abc: at most %d detections per side", COIN_ABC_MAX);
    cudaStream_t s = as_stream(stream);
    for (const AbcHostOut* o : {rcnn, rpn})
        if (o) {
            COIN_REQUIRE(o->counts, "match_abc: counts is null");
            if (nc == 0 && nd == 0) fill_bytes(o->counts, 0, 8 * sizeof(int32_t), s);   // otherwise the kernel writes all 8
        }
    if (nc == 0 && nd == 0) return COIN_OK;
    COIN_REQUIRE(cap_pairs >= nc * nd + nc + nd, "match_abc: cap_pairs must be >= nc*nd + nc + nd");
    COIN_REQUIRE(c_on && c_off && ws, "match_abc: null pointer");
    COIN_REQUIRE(!rcnn || (rcnn->a_on && rcnn->a_off && rcnn->a_boxes && rcnn->b_on && rcnn->b_off && rcnn->b_boxes),
                 "match_abc: A and B outputs are required for tag RCNN");
    COIN_REQUIRE(!rpn || (rpn->a_on && rpn->a_off && rpn->a_boxes), "match_abc: A outputs are required for tag RPN");
    COIN_REQUIRE(nc == 0 || (on_boxes && on_classes && on_scores), "match_abc: null cloud inputs");
    COIN_REQUIRE(nd == 0 || (off_boxes && off_classes && off_scores), "match_abc: null CLIP-detector inputs");
    AbcArgs a;
    size_t total = 0;
    carve_abc(a, ws, nc, nd, cap_pairs, &total);
    if (ws_bytes < total) return fail(COIN_ERR_CAPACITY, "match_abc: workspace too small (%zu < %zu)", ws_bytes, total);
    a.onb = reinterpret_cast<const float4*>(on_boxes); a.offb = reinterpret_cast<const float4*>(off_boxes);
    a.oncls64 = on_classes; a.offcls64 = off_classes; a.ons = on_scores; a.offs = off_scores;
    a.nc = (int)nc; a.nd = (int)nd; a.nd_dev = nd_dev; a.thr = iou_thr; a.w_a = weight_for_box_a; a.cap = (int)cap_pairs;
    const AbcHostOut* outs[2] = {rcnn, rpn};
    for (int t = 0; t < 2; ++t) {
        AbcOut& o = a.out[t];
        o = AbcOut{};
        if (!outs[t]) continue;
        o.a_on = outs[t]->a_on; o.a_off = outs[t]->a_off; o.b_on = outs[t]->b_on; o.b_off = outs[t]->b_off;
        o.a_box = reinterpret_cast<float4*>(outs[t]->a_boxes); o.b_box = reinterpret_cast<float4*>(outs[t]->b_boxes);
        o.counts = outs[t]->counts;
    }
    a.c_on = c_on; a.c_off = c_off;
    a.use_smem = (nc <= kAbcN && nd <= kAbcN) ? 1 : 0;
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(match_abc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kAbcSmemBytes);
        attr_set = true;
    }
    match_abc_kernel<<<1, 256, a.use_smem ? kAbcSmemBytes : 0, s>>>(a);
    return check_launch("match_abc_kernel");
}

static int match_abc_one(const float* on_boxes, const int64_t* on_classes, const float* on_scores, int64_t nc,
                         const float* off_boxes, const int64_t* off_classes, const float* off_scores, int64_t nd,
                         const int32_t* nd_dev, int tag, float iou_thr, float weight_for_box_a, int64_t cap_pairs,
                         int32_t* a_on, int32_t* a_off, float* a_boxes, int32_t* b_on, int32_t* b_off, float* b_boxes,
                         int32_t* c_on, int32_t* c_off, int32_t* counts, void* ws, size_t ws_bytes, coin_stream_t stream) {
    COIN_REQUIRE(tag == COIN_TAG_RCNN || tag == COIN_TAG_RPN, "match_abc: bad tag %d", tag);
    AbcHostOut o{a_on, a_off, a_boxes, b_on, b_off, b_boxes, counts};
    return match_abc_impl(on_boxes, on_classes, on_scores, nc, off_boxes, off_classes, off_scores, nd, nd_dev, iou_thr,
                          weight_for_box_a, cap_pairs, tag == COIN_TAG_RCNN ? &o : nullptr, tag == COIN_TAG_RPN ? &o : nullptr,
                          c_on, c_off, ws, ws_bytes, stream);
}

extern "C" int coin_match_abc(const float* on_boxes, const int64_t* on_classes, const float* on_scores, int64_t nc,
                              const float* off_boxes, const int64_t* off_classes, const float* off_scores,
                              int64_t nd, int tag, float iou_thr, float weight_for_box_a, int64_t cap_pairs,
                              int32_t* a_on, int32_t* a_off, float* a_boxes, int32_t* b_on, int32_t* b_off,
                              float* b_boxes, int32_t* c_on, int32_t* c_off, int32_t* counts, void* ws,
                              size_t ws_bytes, coin_stream_t stream) {
    return match_abc_one(on_boxes, on_classes, on_scores, nc, off_boxes, off_classes, off_scores, nd, nullptr, tag,
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
    return match_abc_one(on_boxes, on_classes, on_scores, nc, off_boxes, off_classes, off_scores, nd_cap, nd_dev, tag,
                         iou_thr, weight_for_box_a, cap_pairs, a_on, a_off, a_boxes, b_on, b_off, b_boxes, c_on, c_off,
                         counts, ws, ws_bytes, stream);
}

extern "C" int coin_match_abc_both_dev(const float* on_boxes, const int64_t* on_classes, const float* on_scores, int64_t nc,
                                       const float* off_boxes, const int64_t* off_classes, const float* off_scores,
                                       int64_t nd_cap, const int32_t* nd_dev, float iou_thr, float weight_for_box_a,
                                       int64_t cap_pairs, int32_t* rcnn_a_on, int32_t* rcnn_a_off, float* rcnn_a_boxes,
                                       int32_t* rcnn_b_on, int32_t* rcnn_b_off, float* rcnn_b_boxes, int32_t* rpn_a_on,
                                       int32_t* rpn_a_off, float* rpn_a_boxes, int32_t* c_on, int32_t* c_off,
                                       int32_t* counts_rcnn, int32_t* counts_rpn, void* ws, size_t ws_bytes,
                                       coin_stream_t stream) {
    AbcHostOut r{rcnn_a_on, rcnn_a_off, rcnn_a_boxes, rcnn_b_on, rcnn_b_off, rcnn_b_boxes, counts_rcnn};
    AbcHostOut p{rpn_a_on, rpn_a_off, rpn_a_boxes, nullptr, nullptr, nullptr, counts_rpn};
    return match_abc_impl(on_boxes, on_classes, on_scores, nc, off_boxes, off_classes, off_scores, nd_cap, nd_dev, iou_thr,
                          weight_for_box_a, cap_pairs, &r, &p, c_on, c_off, ws, ws_bytes, stream);
}
