// pyset.cuh -- iteration order of CPython `set` objects holding small non-negative ints, reproduced exactly.
//
// Why this exists: three places of the reference's knowledge separation turn a Python set into a list,
//     list(set(range(n)) - set(matched))                 coin/engine/trainer.py:369,391   (row order of the private set C)
//     result[list(i)] for i in sets                      coin/utils/util.py:481           (member order of a self-cluster;
//                                                        util.py:497 keys a decision on the FIRST member)
// and the order of list(set) is the slot order of CPython's open-addressing table (Objects/setobject.c: hash(int) = int,
// slot = hash & mask, 1 + LINEAR_PROBES linear probes, then i = i*5 + 1 + (perturb >>= 5); growth when fill*5 >= mask*3
// to the first power of two above 4*used; set_merge's pre-sizing rule; set_difference's copy-and-discard shortcut).
// To return what the reference returns, row for row, the device replays those table operations. The functions below
// are plain integer code shared by the CUDA kernel (match_abc.cu) and a host harness (tests/csrc/pyset_host.cpp) that
// fuzzes them against the running interpreter's real sets.
//
// Keys are < 32768 and stored as int16 (kEmpty = unused slot). Sets are immutable once built (the reference rebinds
// `sets[i] = sets[i] | ...`, it never mutates in place), so a set is a (table offset, mask, used) handle into a bump pool.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define COIN_PYSET_HD __host__ __device__ __forceinline__
#else
#define COIN_PYSET_HD inline
#endif

namespace coin {
namespace pyset {

constexpr int kLinearProbes = 9;
constexpr int kPerturbShift = 5;
constexpr int16_t kEmpty = -1;

struct Pool {            // bump allocator; `used` / `overflow` live outside so that several threads can share one pool
    int16_t* base;
    int cap;
    int* used;
    int* overflow;
};

struct Set {        // tab == nullptr <=> no storage (only legal for used == 0)
    int16_t* tab;
    int mask, fill, used;
};

COIN_PYSET_HD int16_t* pool_alloc(Pool& p, int slots) {
#if defined(__CUDA_ARCH__)
    const int off = atomicAdd(p.used, slots);
#else
    const int off = *p.used;
    *p.used += slots;
#endif
    if (off + slots > p.cap) { *p.overflow = 1; return nullptr; }
    int16_t* r = p.base + off;
    for (int i = 0; i < slots; ++i) r[i] = kEmpty;
    return r;
}

COIN_PYSET_HD Set make_empty(Pool& p) {
    Set s;
    s.tab = pool_alloc(p, 8);   // the object's smalltable
    s.mask = 7; s.fill = 0; s.used = 0;
    return s;                   // tab == nullptr on overflow: every caller checks *p.overflow before touching it
}

// set_insert_clean: the key is known to be absent and the table has no dummies
COIN_PYSET_HD void insert_clean(int16_t* tab, int mask, int key) {
    uint32_t perturb = (uint32_t)key, i = (uint32_t)key & (uint32_t)mask;
    while (true) {
        if (tab[i] == kEmpty) { tab[i] = (int16_t)key; return; }
        if (i + kLinearProbes <= (uint32_t)mask) {
            for (int j = 1; j <= kLinearProbes; ++j)
                if (tab[i + j] == kEmpty) { tab[i + j] = (int16_t)key; return; }
        }
        perturb >>= kPerturbShift;
        i = (i * 5 + 1 + perturb) & (uint32_t)mask;
    }
}

// set_table_resize(so, minused)
COIN_PYSET_HD void resize(Set& s, int minused, Pool& p) {
    int newsize = 8;
    while (newsize <= minused) newsize <<= 1;
    if (newsize == 8 && s.mask == 7 && s.fill == s.used) return;   // smalltable, no dummies: nothing to do
    int16_t* nt = pool_alloc(p, newsize);
    if (!nt) return;
    for (int i = 0; i <= s.mask; ++i)
        if (s.tab[i] != kEmpty) insert_clean(nt, newsize - 1, s.tab[i]);
    s.tab = nt; s.mask = newsize - 1; s.fill = s.used;
}

// set_add_entry (no dummies can be present in the sets this file builds)
COIN_PYSET_HD void add(Set& s, int key, Pool& p) {
    if ((*p.overflow)) return;
    uint32_t perturb = (uint32_t)key, i = (uint32_t)key & (uint32_t)s.mask;
    uint32_t e;
    while (true) {
        int probes = (i + kLinearProbes <= (uint32_t)s.mask) ? kLinearProbes : 0;
        e = i;
        bool found = false;
        do {
            if (s.tab[e] == kEmpty) { found = true; break; }
            if (s.tab[e] == (int16_t)key) return;      // already a member
            ++e;
        } while (probes--);
        if (found) break;
        perturb >>= kPerturbShift;
        i = (i * 5 + 1 + perturb) & (uint32_t)s.mask;
    }
    s.tab[e] = (int16_t)key;
    ++s.fill; ++s.used;
    if (s.fill * 5 < s.mask * 3) return;
    resize(s, s.used * 4, p);      // used <= 50000 always
}

// set_merge(so, other): so |= other
COIN_PYSET_HD void merge(Set& so, const Set& other, Pool& p) {
    if ((*p.overflow) || other.used == 0) return;
    if ((so.fill + other.used) * 5 >= so.mask * 3) resize(so, (so.used + other.used) * 2, p);
    if ((*p.overflow)) return;
    if (so.fill == 0 && so.mask == other.mask && other.fill == other.used) {   // same size, empty target: slot copy
        for (int i = 0; i <= other.mask; ++i) so.tab[i] = other.tab[i];
        so.fill = other.fill; so.used = other.used;
        return;
    }
    if (so.fill == 0) {
        for (int i = 0; i <= other.mask; ++i)
            if (other.tab[i] != kEmpty) insert_clean(so.tab, so.mask, other.tab[i]);
        so.fill = other.used; so.used = other.used;
        return;
    }
    for (int i = 0; i <= other.mask; ++i)
        if (other.tab[i] != kEmpty) add(so, other.tab[i], p);
}

// a | b  (set_or: set_copy(a) then set_update_internal(result, b))
COIN_PYSET_HD Set set_union(const Set& a, const Set& b, Pool& p) {
    Set r = make_empty(p);
    if ((*p.overflow)) return r;
    merge(r, a, p);
    merge(r, b, p);
    return r;
}

COIN_PYSET_HD bool contains(const Set& s, int key) {
    if (s.used == 0) return false;
    for (int i = 0; i <= s.mask; ++i)
        if (s.tab[i] == (int16_t)key) return true;
    return false;
}

// (a - b) == set()
COIN_PYSET_HD bool is_subset(const Set& a, const Set& b) {
    if (a.used == 0) return true;
    for (int i = 0; i <= a.mask; ++i)
        if (a.tab[i] != kEmpty && !contains(b, a.tab[i])) return false;
    return true;
}

// Table size a set reaches when `m` distinct keys are ADDED one by one to an empty set (set_add_entry growth rule).
COIN_PYSET_HD int table_size_after_adds(int m) {
    int mask = 7, fill = 0;
    for (int k = 0; k < m; ++k) {
        ++fill;
        if (fill * 5 >= mask * 3) {
            int newsize = 8;
            while (newsize <= fill * 4) newsize <<= 1;
            mask = newsize - 1;
        }
    }
    return mask + 1;
}

// list(set(range(n)) - other): `keep(i)` tells whether i (0 <= i < n) is NOT in `other`; other_size = len(other).
// Writes the order into out[] and returns its length. `pool` needs 8 + 32 + 128 + 512 slots at most for n <= 1024.
template <class Keep>
COIN_PYSET_HD int difference_order(int n, Keep keep, int other_size, int32_t* out, Pool& pool) {
    int m = 0, mx = -1;
    for (int i = 0; i < n; ++i)
        if (keep(i)) { ++m; mx = i; }
    // set_difference: a much larger left operand is copied and the right one's members discarded -> the layout of
    // set(range(n)) (every key in its own slot, ascending) survives. Otherwise the survivors are ADDED, ascending, to a
    // new set; when all of them are smaller than the final table they also sit in their own slots.
    const bool ascending = ((n >> 2) > other_size) || mx < table_size_after_adds(m);
    if (ascending) {
        int c = 0;
        for (int i = 0; i < n; ++i)
            if (keep(i)) out[c++] = i;
        return c;
    }
    *pool.used = 0; *pool.overflow = 0;
    Set r = make_empty(pool);
    for (int i = 0; i < n && !(*pool.overflow); ++i)
        if (keep(i)) add(r, i, pool);
    if ((*pool.overflow)) {            // cannot happen within the documented sizes; stay deterministic
        int c = 0;
        for (int i = 0; i < n; ++i)
            if (keep(i)) out[c++] = i;
        return c;
    }
    int c = 0;
    for (int i = 0; i <= r.mask; ++i)
        if (r.tab[i] != kEmpty) out[c++] = r.tab[i];
    return c;
}

// ----------------------------------------------------------------------------------------------------------------
// filter_result / find_same (coin/utils/util.py:459-482), replayed literally.
//   adj_word(i, w) : bits [32w, 32w+32) of row i of `IoU(box i, box j) >= thresh` (the reference's iou_matrix, diagonal
//                    included), W words per row - the rows are walked by set bit, never by an O(n^2) scan
//   Output     : clusters[k] = handle of the k-th surviving set with len != 1 (len 0 sets are dropped first);
//                their slot order is the order of `result[list(i)]`.
// Node states: a handle per node (nodes whose set is exactly {i} never allocate: they take no part in anything).
// ----------------------------------------------------------------------------------------------------------------
struct Handle { int32_t off; int16_t mask, used; };     // off < 0: kNone (empty set) / kSelf (the singleton {i})
constexpr int32_t kNone = -1, kSelf = -2;
constexpr int kMaxDepth = 128;   // recursion depth of find_same <= size of a connected component

struct Frame { int16_t node; Handle h; int16_t pos; };

COIN_PYSET_HD Set view(const Handle& h, Pool& p) {
    Set s;
    s.tab = h.off >= 0 ? p.base + h.off : nullptr;
    s.mask = h.mask; s.fill = h.used; s.used = h.used;
    return s;
}
COIN_PYSET_HD Handle handle_of(const Set& s, Pool& p) {
    Handle h;
    h.off = s.used == 0 ? kNone : (int32_t)(s.tab - p.base);
    h.mask = (int16_t)s.mask; h.used = (int16_t)s.used;
    return h;
}

// Returns the number of clusters (<= max_clusters) or -1 on overflow of the pool / recursion stack / cluster list
// (the caller then falls back to its order-free policy and flags the result).
COIN_PYSET_HD int popc32(uint32_t x) {
#if defined(__CUDA_ARCH__)
    return __popc(x);
#else
    return __builtin_popcount(x);
#endif
}
COIN_PYSET_HD int ctz32(uint32_t x) {
#if defined(__CUDA_ARCH__)
    return __ffs((int)x) - 1;
#else
    return __builtin_ctz(x);
#endif
}

COIN_PYSET_HD int filter_clusters_built(int n, const int16_t* active, Handle* sets, Frame* stack, Pool& pool);

// set(iou_matrix[i].nonzero()[:,0].tolist()) for one node: its neighbours added in ascending order
template <class AdjWord>
COIN_PYSET_HD Handle build_row_set(int i, int W, AdjWord adj_word, Pool& pool) {
    Handle h;
    int deg = 0, only = -1;
    for (int w = 0; w < W; ++w) {
        const uint32_t m = adj_word(i, w);
        if (m) { deg += popc32(m); only = w * 32 + ctz32(m); }
    }
    if (deg == 0) { h.off = kNone; h.mask = 7; h.used = 0; return h; }
    if (deg == 1 && only == i) { h.off = kSelf; h.mask = 7; h.used = 1; return h; }
    Set s = make_empty(pool);
    for (int w = 0; w < W && !*pool.overflow; ++w) {
        uint32_t m = adj_word(i, w);
        while (m && !*pool.overflow) { add(s, w * 32 + ctz32(m), pool); m &= m - 1; }
    }
    if (*pool.overflow) { h.off = kNone; h.mask = 7; h.used = 0; return h; }
    return handle_of(s, pool);
}

// Serial reference form: build every node's set, replay, list the clusters.
template <class AdjWord>
COIN_PYSET_HD int filter_clusters(int n, int W, AdjWord adj_word, Handle* sets /*[n]*/, Frame* stack /*[kMaxDepth]*/,
                                  Pool& pool, Handle* clusters, int max_clusters) {
    // sets[i] = set(iou_matrix[i].nonzero()[:,0].tolist())   (the caller resets the pool)
    for (int i = 0; i < n; ++i) {
        sets[i] = build_row_set(i, W, adj_word, pool);
        if (*pool.overflow) return -1;
    }
    if (filter_clusters_built(n, nullptr, sets, stack, pool) < 0) return -1;
    int nc = 0;
    for (int i = 0; i < n; ++i) {
        if (sets[i].off < 0 || sets[i].used == 1) continue;    // len 0 dropped, len 1 skipped (util.py:480-481)
        if (nc == max_clusters) return -1;
        clusters[nc++] = sets[i];
    }
    return nc;
}

// Node classes for the parallel form. Near-duplicate detections almost always come in PAIRS: i ~ j and nothing else.
// For such a component the replay is a fixed point: sets[i] | find_same(...) copies the 8-slot table of sets[i] slot by
// slot and adds members it already has, so the cluster is the initial set of the lower node, as built. (From three
// members on, set_merge pre-sizes the copy and re-lays it out, so larger components are replayed for real.)
constexpr int kIdle = 0, kPairLeader = 1, kPairFollower = 2, kActive = 3;

template <class AdjWord>
COIN_PYSET_HD int classify_node(int i, int W, AdjWord adj_word) {
    int deg = 0, lo = -1, hi = -1;
    for (int w = 0; w < W; ++w) {
        uint32_t m = adj_word(i, w);
        deg += popc32(m);
        while (m) { const int b = w * 32 + ctz32(m); if (lo < 0) lo = b; hi = b; m &= m - 1; }
    }
    if (deg == 0 || (deg == 1 && lo == i)) return kIdle;
    if (deg == 2 && (lo == i || hi == i)) {
        const int other = lo == i ? hi : lo;
        bool same = true;
        for (int w = 0; w < W; ++w) same = same && (adj_word(other, w) == adj_word(i, w));
        if (same) return i < other ? kPairLeader : kPairFollower;
    }
    return kActive;
}

// Parallel-friendly form, written serially here (the kernel runs steps 1-2 with one thread per node and step 4 as a
// ballot compaction; step 3 is the only serial part and visits the kActive nodes only). Same result as filter_clusters.
template <class AdjWord>
COIN_PYSET_HD int filter_clusters_fast(int n, int W, AdjWord adj_word, int* kind /*[n]*/, Handle* sets /*[n]*/,
                                       int16_t* active /*[n]*/, Frame* stack, Pool& pool, Handle* clusters, int max_clusters) {
    int n_active = 0;
    for (int i = 0; i < n; ++i) {                                   // 1, 2: classify, build the sets that matter
        kind[i] = classify_node(i, W, adj_word);
        if (kind[i] == kPairLeader || kind[i] == kActive) sets[i] = build_row_set(i, W, adj_word, pool);
        if (*pool.overflow) return -1;
        if (kind[i] == kActive) active[n_active++] = (int16_t)i;
    }
    if (n_active && filter_clusters_built(n_active, active, sets, stack, pool) < 0) return -1;   // 3
    int nc = 0;
    for (int i = 0; i < n; ++i) {                                   // 4: clusters in ascending order of their owner
        const bool is_cluster = kind[i] == kPairLeader || (kind[i] == kActive && sets[i].off >= 0 && sets[i].used != 1);
        if (!is_cluster) continue;
        if (nc == max_clusters) return -1;
        clusters[nc++] = sets[i];
    }
    return nc;
}

// The replay proper (util.py:471-478 with find_same inlined as an explicit stack) over sets[] already built. `active`:
// optional ascending list of the nodes to visit (nullptr: all of 0..n-1); results stay in sets[]. Returns 0, or -1 on
// overflow of the pool / recursion stack.
COIN_PYSET_HD int filter_clusters_built(int n, const int16_t* active, Handle* sets, Frame* stack, Pool& pool) {
    for (int q = 0; q < n; ++q) {
        const int i = active ? active[q] : q;
        if (sets[i].off < 0) continue;       // {} and {i}: both loops of util.py:471-478 do nothing
        // for j in sets[i]  (the object bound NOW; later rebinding of sets[i] does not affect this iteration)
        const Handle it = sets[i];
        for (int sl = 0; sl <= it.mask; ++sl) {
            const int j = pool.base[it.off + sl];
            if (j == kEmpty || j == i) continue;
            // ---- r = find_same(sets, [i], j), with an explicit stack
            Handle ret;
            int depth = 0;
            stack[0].node = (int16_t)j; stack[0].h = sets[j]; stack[0].pos = 0;
            while (true) {
                Frame& f = stack[depth];
                const int fi = f.node;
                int nxt = -1;
                if (f.h.off >= 0) {
                    while (f.pos <= f.h.mask) {
                        const int cand = pool.base[f.h.off + f.pos];
                        ++f.pos;
                        if (cand == kEmpty || cand == fi || cand == i) continue;      // j != i and j not in ups ...
                        bool in_ups = false;
                        for (int d = 0; d < depth; ++d) in_ups |= (stack[d].node == cand);
                        if (in_ups) continue;
                        // if sets[j] - sets[i] == set(): pass
                        if (sets[cand].off == kNone) continue;
                        Set sj, si;
                        int16_t self_j = (int16_t)cand, self_i = (int16_t)fi;
                        if (sets[cand].off == kSelf) { sj.tab = &self_j; sj.mask = 0; sj.fill = sj.used = 1; }
                        else sj = view(sets[cand], pool);
                        if (sets[fi].off == kSelf) { si.tab = &self_i; si.mask = 0; si.fill = si.used = 1; }
                        else si = view(sets[fi], pool);
                        if (is_subset(sj, si)) continue;
                        nxt = cand;
                        break;
                    }
                }
                if (nxt >= 0) {                      // sets[i] = sets[i] | find_same(sets, ups + [i], j)
                    if (depth + 1 >= kMaxDepth) return -1;
                    ++depth;
                    stack[depth].node = (int16_t)nxt; stack[depth].h = sets[nxt]; stack[depth].pos = 0;
                    continue;
                }
                ret = sets[fi];                      // return sets[i]
                if (depth == 0) break;
                --depth;
                const int pi = stack[depth].node;
                if (sets[pi].off == kSelf || ret.off == kSelf) return -1;   // unreachable for a symmetric adj; be safe
                Set u = set_union(view(sets[pi], pool), view(ret, pool), pool);
                if ((*pool.overflow)) return -1;
                sets[pi] = handle_of(u, pool);
            }
            if (ret.off == kSelf) return -1;
            Set u = set_union(view(sets[i], pool), view(ret, pool), pool);
            if ((*pool.overflow)) return -1;
            sets[i] = handle_of(u, pool);
        }
        // for j in sets[i]: if j != i: sets[j] = set()
        const Handle fin = sets[i];
        for (int sl = 0; sl <= fin.mask; ++sl) {
            const int j = pool.base[fin.off + sl];
            if (j != kEmpty && j != i) { sets[j].off = kNone; sets[j].mask = 7; sets[j].used = 0; }
        }
    }
    return 0;
}

}  // namespace pyset
}  // namespace coin
