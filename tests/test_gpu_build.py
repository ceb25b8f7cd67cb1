"""GPU tests of the LBVH builder pieces: onesweep sort, tree invariants."""
import numpy as np
import pytest

from owlraytracing_b200 import datasets

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n", [1, 2, 31, 32, 33, 4095, 4096, 4097, 100_003, 1_500_000])
def test_onesweep_sort_matches_stable_argsort(knn, n):
    rng = np.random.default_rng(n)
    keys = rng.integers(0, 2**63, n, dtype=np.uint64)
    vals = np.arange(n, dtype=np.uint32)
    order = np.argsort(keys, kind="stable")
    k2, v2 = knn.sort_pairs(keys.copy(), vals.copy())
    assert (k2 == keys[order]).all() and (v2 == vals[order]).all()


@pytest.mark.parametrize("distinct", [1, 2, 3, 256, 70000])
def test_onesweep_sort_is_stable(knn, distinct):
    n = 300_000
    rng = np.random.default_rng(distinct)
    pool = rng.integers(0, 2**64 - 1, distinct, dtype=np.uint64)
    keys = pool[rng.integers(0, distinct, n)]
    vals = np.arange(n, dtype=np.uint32)
    order = np.argsort(keys, kind="stable")
    k2, v2 = knn.sort_pairs(keys.copy(), vals.copy())
    assert (k2 == keys[order]).all() and (v2 == order.astype(np.uint32)).all()


def test_onesweep_sort_extreme_keys(knn):
    keys = np.array([2**64 - 1, 0, 2**63, 1, 2**64 - 1, 0, 255, 256, 65535, 65536], dtype=np.uint64)
    keys = np.tile(keys, 1000)
    vals = np.arange(keys.size, dtype=np.uint32)
    order = np.argsort(keys, kind="stable")
    k2, v2 = knn.sort_pairs(keys.copy(), vals.copy())
    assert (k2 == keys[order]).all() and (v2 == order.astype(np.uint32)).all()


@pytest.mark.parametrize("n", [1, 33, 4096, 4097, 250_001])
def test_onesweep_keys_only(knn, n):
    """The keys-only kernel variant (the builder's packed (code, index) keys) over all 64 key bits."""
    rng = np.random.default_rng(n + 7)
    keys = rng.integers(0, 2**64 - 1, n, dtype=np.uint64)
    keys[: n // 3] = keys[0]  # a long run of equal keys
    k2, _ = knn.sort_pairs(keys.copy())
    assert (k2 == np.sort(keys)).all()


@pytest.mark.parametrize("cloud", ["uniform", "lidar", "dups"])
def test_packed_sort_equals_pair_sort(knn, cloud):
    """TKNN_OPT_SORT_MODE: packed (code << index bits | index) keys sorted keys-only put the points in the same order
    as the stable (code, index) pair sort with the same number of code bits."""
    n = 70_001
    x = {"uniform": lambda: datasets.uniform(n, seed=3), "lidar": lambda: datasets.lidar_like(n, seed=3),
         "dups": lambda: np.tile(datasets.uniform(701, seed=4), (100, 1))}[cloud]()
    got = []
    for mode in (0, 1):
        knn.set_option("sort_mode", mode)
        knn.set_option("morton_bits", 12)
        knn.build(x)
        nodes, pts, leaf_start = knn.get_bvh()
        _walk(nodes, pts, leaf_start)
        got.append((pts.copy(), leaf_start.copy()))
    assert (got[0][0].view(np.int32) == got[1][0].view(np.int32)).all()
    if cloud != "dups":  # points that share a code: the radix tree splits them by index bits (packed) or by position (pairs)
        assert got[0][1].shape == got[1][1].shape and (got[0][1] == got[1][1]).all()


def _walk(nodes, pts, leaf_start):
    """Returns per-node (lo, hi, count) computed from the leaves; asserts stored child boxes are exact."""
    refs = nodes.view(np.int32)
    n_nodes = nodes.shape[0]
    seen_pts = np.zeros(pts.shape[0], np.int32)
    visited = np.zeros(n_nodes, bool)
    starts = set(leaf_start[:-1].tolist())

    def box_of_child(node, slot):
        # interleaved layout (common.cuh: Node): every pair of words is (child 0, child 1)
        lo = nodes[node, [0 + slot, 2 + slot, 4 + slot]]
        hi = nodes[node, [6 + slot, 8 + slot, 10 + slot]]
        ref = int(refs[node, 12 + slot])
        cnt = int(refs[node, 14 + slot])
        if cnt > 0:
            assert ref in starts
            p = pts[ref: ref + cnt, :3]
            seen_pts[ref: ref + cnt] += 1
            tlo, thi = p.min(0), p.max(0)
        else:
            assert 0 < ref < n_nodes and not visited[ref]
            tlo, thi = visit(ref)
        assert (lo == tlo).all() and (hi == thi).all(), (node, slot)
        return tlo, thi

    def visit(node):
        visited[node] = True
        l0, h0 = box_of_child(node, 0)
        l1, h1 = box_of_child(node, 1)
        return np.minimum(l0, l1), np.maximum(h0, h1)

    import sys
    sys.setrecursionlimit(10000)
    visit(0)
    assert visited.all()
    assert (seen_pts == 1).all()


@pytest.mark.parametrize("cloud,leaf,policy", [("uniform", 32, 0), ("uniform", 8, 0), ("lidar", 32, 0), ("lidar", 16, 1),
                                              ("dups", 32, 0)])
def test_bvh_invariants(knn, cloud, leaf, policy):
    n = 20_000
    if cloud == "uniform":
        x = datasets.uniform(n, seed=1)
    elif cloud == "lidar":
        x = datasets.lidar_like(n, seed=1)
    else:
        x = np.tile(datasets.uniform(50, seed=2), (400, 1))
    knn.set_option("leaf_size", leaf)
    knn.set_option("leaf_policy", policy)
    knn.build(x)
    nodes, pts, leaf_start = knn.get_bvh()
    st = knn.stats()
    assert nodes.shape[0] == st["n_nodes"] == st["n_leaves"] - 1
    # the sorted points are a permutation of the input carrying their original index
    ids = pts.view(np.int32)[:, 3]
    assert (np.sort(ids) == np.arange(x.shape[0])).all()
    assert (pts[:, :3] == x[ids]).all()
    # leaves: contiguous, non-empty, at most leaf_size points
    counts = np.diff(leaf_start.astype(np.int64))
    assert leaf_start[0] == 0 and leaf_start[-1] == x.shape[0]
    assert (counts >= 1).all() and (counts <= leaf).all()
    if policy == 0 and cloud == "uniform":
        assert counts.mean() > leaf * 0.5  # maximal subtrees are at least half full on average
    _walk(nodes, pts, leaf_start)
