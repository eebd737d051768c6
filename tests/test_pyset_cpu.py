"""coin_b200/csrc/pyset.cuh replays CPython's set table (slot order = order of list(set)) so that the device returns
the rows the reference returns, in the reference's order (trainer.py:369,391; util.py:459-482). The header is plain
integer code: here it is compiled for the host and fuzzed against the running interpreter's real sets."""
import ctypes
import os
import random
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
SRC = os.path.join(ROOT, "tests", "csrc", "pyset_host.cpp")
OUT = os.path.join(ROOT, "tests", "csrc", "_build", "libpyset_host.so")


@pytest.fixture(scope="module")
def lib():
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    hdr = os.path.join(ROOT, "coin_b200", "csrc", "pyset.cuh")
    if not os.path.exists(OUT) or os.path.getmtime(OUT) < max(os.path.getmtime(SRC), os.path.getmtime(hdr)):
        env = dict(os.environ)
        env.pop("CC", None), env.pop("CXX", None)
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-x", "c++", SRC, "-o", OUT], env=env)
    return ctypes.CDLL(OUT)


def _i32(n):
    return (ctypes.c_int32 * max(n, 1))()


def test_list_of_set_and_union_order(lib):
    rnd = random.Random(1)
    for trial in range(3000):
        hi = rnd.choice([8, 12, 40, 130, 600, 1024])
        a = [rnd.randrange(hi) for _ in range(rnd.randrange(0, rnd.choice([4, 9, 30, 100, 400])))]
        out = _i32(len(a))
        n = lib.pyset_list_of_set_from_list(len(a), (ctypes.c_int32 * max(len(a), 1))(*a), out)
        assert list(out[:n]) == list(set(a)), (trial, a)
        b = [rnd.randrange(hi) for _ in range(rnd.randrange(0, rnd.choice([4, 9, 30, 100])))]
        out = _i32(len(a) + len(b))
        n = lib.pyset_list_of_union(len(a), (ctypes.c_int32 * max(len(a), 1))(*a), len(b), (ctypes.c_int32 * max(len(b), 1))(*b), out)
        assert list(out[:n]) == list(set(a) | set(b)), (trial, a, b)


def test_difference_order_matches_python(lib):
    """list(set(range(n)) - set(matched)), trainer.py:369,391."""
    rnd = random.Random(2)
    nontrivial = 0
    for trial in range(4000):
        n = rnd.choice([0, 1, 5, 9, 20, 33, 60, 100, 128, 300, 1000])
        frac = rnd.choice([0.0, 0.1, 0.3, 0.7, 0.9, 0.97, 1.0])
        matched = [i for i in range(n) if rnd.random() < frac]
        rnd.shuffle(matched)
        matched = matched + matched[: len(matched) // 3]          # duplicates, as match_inds[:, 1].tolist() has
        want = list(set([i for i in range(n)]) - set(matched))
        keep = np.ones(max(n, 1), dtype=np.uint8)
        keep[list(set(matched))] = 0
        out = _i32(n)
        m = lib.pyset_difference_order(n, keep.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8)), len(set(matched)), out)
        assert list(out[:m]) == want, (trial, n, sorted(set(matched)))
        nontrivial += want != sorted(want)
    assert nontrivial > 200        # the fuzz does reach the non-ascending layouts


def _literal_clusters(adj):
    """coin/utils/util.py:459-482 (filter_result + find_same) on an adjacency matrix, real Python sets."""
    n = len(adj)
    sets = [set(j for j in range(n) if adj[i][j]) for i in range(n)]

    def find_same(sets, ups, i):
        for j in sets[i]:
            if j != i and j not in ups:
                if sets[j] - sets[i] == set():
                    pass
                else:
                    sets[i] = sets[i] | find_same(sets, ups + [i], j)
        return sets[i]

    for i in range(len(sets)):
        for j in sets[i]:
            ups = []
            if j != i:
                sets[i] = sets[i] | find_same(sets, ups + [i], j)
        for j in sets[i]:
            if j != i:
                sets[j] = set()
    sets = [s for s in sets if len(s) != 0]
    return [list(s) for s in sets if len(s) != 1]


def _random_graph(rnd, n, kind):
    adj = [[i == j for j in range(n)] for i in range(n)]
    def link(a, b):
        adj[a][b] = adj[b][a] = True
    if kind == "pairs":
        for _ in range(rnd.randrange(1, max(n // 3, 2))):
            a, b = rnd.randrange(n), rnd.randrange(n)
            link(a, b)
    elif kind == "chains":
        for _ in range(rnd.randrange(1, 4)):
            nodes = rnd.sample(range(n), min(n, rnd.randrange(2, 9)))
            for a, b in zip(nodes, nodes[1:]):
                link(a, b)
    elif kind == "cliques":
        for _ in range(rnd.randrange(1, 4)):
            nodes = rnd.sample(range(n), min(n, rnd.randrange(2, 7)))
            for a in nodes:
                for b in nodes:
                    link(a, b)
    else:   # dense random
        p = rnd.choice([0.02, 0.05, 0.15])
        for a in range(n):
            for b in range(a):
                if rnd.random() < p:
                    link(a, b)
        for a in rnd.sample(range(n), n // 10):   # zero-area boxes: IoU with themselves is 0
            for b in range(n):
                adj[a][b] = adj[b][a] = False
    return adj


def test_filter_clusters_matches_literal_python(lib):
    rnd = random.Random(3)
    seen_reordered = 0
    for trial in range(1500):
        n = rnd.choice([2, 5, 9, 17, 40, 100, 128])
        adj = _random_graph(rnd, n, rnd.choice(["pairs", "chains", "cliques", "random"]))
        want = _literal_clusters([row[:] for row in adj])
        flat = np.array(adj, dtype=np.uint8).reshape(-1)
        members, offsets = _i32(n * n + n), _i32(n + 2)
        kind = "?"
        nc = lib.pyset_filter_clusters(n, flat.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8)), 1 << 22, members, offsets, n + 1)
        assert nc >= 0, trial
        got = [list(members[offsets[k]: offsets[k + 1]]) for k in range(nc)]
        assert got == want, (trial, n, got, want)
        seen_reordered += any(c != sorted(c) for c in want)
        # the form the kernel runs (pairs by closed form, the rest replayed) gives the same clusters in the same order
        nc3 = lib.pyset_filter_clusters_fast(n, flat.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8)), 1 << 22, members, offsets, n + 1)
        assert nc3 >= 0 and [list(members[offsets[k]: offsets[k + 1]]) for k in range(nc3)] == want, (trial, n)
        # with the DEVICE's pool (4096 slots in shared memory) the answer is the same or an explicit overflow
        nc2 = lib.pyset_filter_clusters(n, flat.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8)), 4096, members, offsets, n + 1)
        assert nc2 == -1 or [list(members[offsets[k]: offsets[k + 1]]) for k in range(nc2)] == want
    assert seen_reordered > 50


def test_device_pool_suffices_for_detection_like_graphs(lib):
    """Near-duplicate cloud boxes form pairs, short chains and small cliques: 4096 slots never overflow there."""
    rnd = random.Random(4)
    for trial in range(600):
        n = rnd.choice([30, 100, 128])
        adj = _random_graph(rnd, n, rnd.choice(["pairs", "chains", "cliques"]))
        flat = np.array(adj, dtype=np.uint8).reshape(-1)
        members, offsets = _i32(n * n + n), _i32(n + 2)
        assert lib.pyset_filter_clusters(n, flat.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8)), 4096, members, offsets, n + 1) >= 0
        assert lib.pyset_filter_clusters_fast(n, flat.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8)), 4096, members, offsets, n + 1) >= 0


def test_filter_clusters_reports_overflow_instead_of_guessing(lib):
    n = 64
    adj = np.ones((n, n), dtype=np.uint8)       # one 64-clique: needs far more than 256 slots
    members, offsets = _i32(n * n + n), _i32(n + 2)
    nc = lib.pyset_filter_clusters(n, adj.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8)), 256, members, offsets, n + 1)
    assert nc == -1
