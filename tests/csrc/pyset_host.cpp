// Host harness for coin_b200/csrc/pyset.cuh (the CPython-set replay the match_abc kernel runs on the device):
// built by tests/test_pyset_cpu.py with g++ and fuzzed against the running interpreter's real `set` objects.
#include <cstdint>
#include <vector>

#include "../../coin_b200/csrc/pyset.cuh"

using namespace coin::pyset;

extern "C" int pyset_difference_order(int n, const uint8_t* keep, int other_size, int32_t* out) {
    std::vector<int16_t> store(8 + 32 + 128 + 512 + 2048 + 8192);
    int used = 0, overflow = 0;
    Pool pool{store.data(), (int)store.size(), &used, &overflow};
    return difference_order(n, [&](int i) { return keep[i] != 0; }, other_size, out, pool);
}

// adj: n*n bytes. members: concatenated cluster members in list(set) order; offsets[k]..offsets[k+1]. Returns the
// number of clusters or -1 on overflow.
extern "C" int pyset_filter_clusters(int n, const uint8_t* adj, int pool_slots, int32_t* members, int32_t* offsets, int max_clusters) {
    std::vector<int16_t> store(pool_slots);
    int used = 0, overflow = 0;
    Pool pool{store.data(), pool_slots, &used, &overflow};
    std::vector<Handle> sets(n), clusters(max_clusters);
    std::vector<Frame> stack(kMaxDepth);
    const int W = (n + 31) / 32;
    std::vector<uint32_t> words((size_t)n * W, 0u);
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j)
            if (adj[(size_t)i * n + j]) words[(size_t)i * W + j / 32] |= 1u << (j % 32);
    const int nc = filter_clusters(n, W, [&](int i, int w) { return words[(size_t)i * W + w]; }, sets.data(), stack.data(), pool,
                                   clusters.data(), max_clusters);
    if (nc < 0) return nc;
    int at = 0;
    for (int k = 0; k < nc; ++k) {
        offsets[k] = at;
        for (int s = 0; s <= clusters[k].mask; ++s) {
            const int v = pool.base[clusters[k].off + s];
            if (v != kEmpty) members[at++] = v;
        }
    }
    offsets[nc] = at;
    return nc;
}

// the classify / pair shortcut / replay-the-rest form the kernel runs (same arguments and result as above)
extern "C" int pyset_filter_clusters_fast(int n, const uint8_t* adj, int pool_slots, int32_t* members, int32_t* offsets,
                                          int max_clusters) {
    std::vector<int16_t> store(pool_slots);
    int used = 0, overflow = 0;
    Pool pool{store.data(), pool_slots, &used, &overflow};
    std::vector<Handle> sets(n), clusters(max_clusters);
    std::vector<Frame> stack(kMaxDepth);
    std::vector<int> kind(n);
    std::vector<int16_t> active(n);
    const int W = (n + 31) / 32;
    std::vector<uint32_t> words((size_t)n * W, 0u);
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j)
            if (adj[(size_t)i * n + j]) words[(size_t)i * W + j / 32] |= 1u << (j % 32);
    const int nc = filter_clusters_fast(n, W, [&](int i, int w) { return words[(size_t)i * W + w]; }, kind.data(), sets.data(),
                                        active.data(), stack.data(), pool, clusters.data(), max_clusters);
    if (nc < 0) return nc;
    int at = 0;
    for (int k = 0; k < nc; ++k) {
        offsets[k] = at;
        for (int s = 0; s <= clusters[k].mask; ++s) {
            const int v = pool.base[clusters[k].off + s];
            if (v != kEmpty) members[at++] = v;
        }
    }
    offsets[nc] = at;
    return nc;
}

extern "C" int pyset_list_of_set_from_list(int n, const int32_t* keys, int32_t* out) {   // list(set([k0, k1, ...]))
    std::vector<int16_t> store(1 << 16);
    int used = 0, overflow = 0;
    Pool pool{store.data(), (int)store.size(), &used, &overflow};
    Set s = make_empty(pool);
    for (int i = 0; i < n; ++i) add(s, keys[i], pool);
    int c = 0;
    for (int i = 0; i <= s.mask; ++i)
        if (s.tab[i] != kEmpty) out[c++] = s.tab[i];
    return c;
}

extern "C" int pyset_list_of_union(int na, const int32_t* a, int nb, const int32_t* b, int32_t* out) {   // list(set(a) | set(b))
    std::vector<int16_t> store(1 << 16);
    int used = 0, overflow = 0;
    Pool pool{store.data(), (int)store.size(), &used, &overflow};
    Set sa = make_empty(pool), sb = make_empty(pool);
    for (int i = 0; i < na; ++i) add(sa, a[i], pool);
    for (int i = 0; i < nb; ++i) add(sb, b[i], pool);
    Set u = set_union(sa, sb, pool);
    int c = 0;
    for (int i = 0; i <= u.mask; ++i)
        if (u.tab[i] != kEmpty) out[c++] = u.tab[i];
    return c;
}
