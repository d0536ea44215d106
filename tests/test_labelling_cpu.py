"""Host-side check of the algorithm behind K9 (`k9_label_relax` / `k9_label_jump`, mvskit_b200/csrc/pmk_filter.cuh): the reference's
filterSmallGroups (filter.cpp:432-578) labels breadth-first in m_ppatches order over a DIRECTED neighbour relation, skipping patches that
already carry a label.  The device computes label(v) = the smallest id among the patches that can reach v (v included) as the fixed point of
L[v] = min(L[v], L[u]) over edges u -> v, with the shortcut L[v] = L[L[v]].  Both are restated here in numpy / plain Python and compared on
random directed graphs: same partition into groups, hence same sizes and same removals.  No GPU, no library."""
import numpy as np
import pytest


def bfs_in_order(n, offs, adj):
    """the reference's loop: for every patch in order, if unlabelled start a group and flood along out-edges through unlabelled patches"""
    label = np.full(n, -1, np.int64)
    groups = 0
    for s in range(n):
        if label[s] >= 0:
            continue
        label[s] = groups
        queue = [s]
        while queue:
            u = queue.pop(0)
            for v in adj[offs[u]:offs[u + 1]]:
                if label[v] < 0:
                    label[v] = groups
                    queue.append(v)
        groups += 1
    return label


def min_ancestor_fixed_point(n, offs, adj, rounds_out=None):
    """the device's rounds: relax every edge once (atomicMin order does not matter), then pointer-jump every label to its own label's label"""
    L = np.arange(n, dtype=np.int64)
    src = np.repeat(np.arange(n), np.diff(offs))
    rounds = 0
    while True:
        rounds += 1
        before = L.copy()
        np.minimum.at(L, adj, L[src])                       # k9_label_relax (reads may see this round's writes on the device: still monotone)
        while True:                                          # k9_label_jump, to its own fixed point as the kernel's while loop does
            LL = L[L]
            if np.array_equal(LL, L):
                break
            L = LL
        if np.array_equal(L, before):
            break
    if rounds_out is not None:
        rounds_out.append(rounds)
    return L


def random_digraph(rng, n, mean_degree, clusters):
    """edges mostly inside `clusters` index ranges (islands, as thinned patch stores have), a few across, asymmetric on purpose"""
    size = max(1, n // clusters)
    lists = []
    for u in range(n):
        k = rng.poisson(mean_degree)
        lo = (u // size) * size
        local = rng.randint(lo, min(n, lo + size), k)
        far = rng.randint(0, n, rng.binomial(1, 0.02))
        lists.append(np.unique(np.concatenate([local, far])))
    offs = np.zeros(n + 1, np.int64)
    offs[1:] = np.cumsum([len(x) for x in lists])
    adj = np.concatenate(lists).astype(np.int64) if offs[-1] else np.zeros(0, np.int64)
    return offs, adj


@pytest.mark.parametrize("seed,n,deg,clusters", [(0, 1, 0.0, 1), (1, 50, 0.3, 5), (2, 400, 0.8, 40), (3, 400, 2.5, 8), (4, 3000, 1.2, 150),
                                                 (5, 3000, 0.5, 30), (6, 2000, 3.0, 400)])
def test_min_ancestor_labels_are_the_reference_groups(seed, n, deg, clusters):
    rng = np.random.RandomState(seed)
    offs, adj = random_digraph(rng, n, deg, clusters)
    ref = bfs_in_order(n, offs, adj)
    rounds = []
    L = min_ancestor_fixed_point(n, offs, adj, rounds)
    # same partition: the group of v is named after its first member on both sides
    first = np.full(ref.max() + 1, -1, np.int64)
    for v in range(n):
        if first[ref[v]] < 0:
            first[ref[v]] = v
    assert np.array_equal(first[ref], L)
    # hence the same sizes and the same removals at any threshold
    size_ref = np.bincount(ref)[ref]
    size_dev = np.bincount(L, minlength=n)[L]
    assert np.array_equal(size_ref, size_dev)
    assert rounds[0] <= 64                                   # pointer jumping keeps the rounds to a handful even on chains


def test_a_chain_against_the_index_order_needs_few_rounds():
    """worst case for plain relaxation: edges v+1 -> v, the smallest id sits at the far end of a path of n hops"""
    n = 4096
    # node u >= 1 has one out-edge, to u - 1; node 0 has none
    offs = np.concatenate([[0, 0], np.arange(1, n)]).astype(np.int64)
    adj = np.arange(0, n - 1, dtype=np.int64)
    ref = bfs_in_order(n, offs, adj)
    rounds = []
    L = min_ancestor_fixed_point(n, offs, adj, rounds)
    # patch 0 reaches nobody, patch 1 only reaches the already labelled patch 0, ...: every patch is a group of its own on both sides
    assert np.array_equal(ref, np.arange(n)) and np.array_equal(L, np.arange(n))
    # and the opposite orientation (u -> u + 1): one group, found in O(log n) rounds thanks to the jumps
    offs = np.concatenate([np.arange(0, n), [n - 1]]).astype(np.int64)
    adj = np.arange(1, n, dtype=np.int64)
    ref = bfs_in_order(n, offs, adj)
    rounds = []
    L = min_ancestor_fixed_point(n, offs, adj, rounds)
    assert (ref == 0).all() and (L == 0).all() and rounds[0] <= 16, rounds
