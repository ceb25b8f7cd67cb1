"""BASELINE.json's full-size configurations on the GPU, checked against the oracle (full compare where
the CPU finishes in seconds, sampled compare + size-independent properties otherwise)."""
import numpy as np
import pytest

from helpers import assert_knn_equal, rows_sorted
from owlraytracing_b200 import datasets

pytestmark = pytest.mark.gpu


def _properties(x, idx, dist, k):
    n = x.shape[0]
    assert idx.min() >= 0 and idx.max() < n                       # every slot filled with a valid index
    assert (idx != np.arange(n)[:, None]).all()                    # self excluded by index
    assert rows_sorted(dist)                                       # ascending
    # reported distance == sqrt(fma-chain d2) recomputed in float64 on a sample (<= 1e-6 relative)
    rows = np.arange(0, n, max(1, n // 5000))
    d = np.linalg.norm(x[rows, None, :].astype(np.float64) - x[idx[rows]].astype(np.float64), axis=-1)
    assert np.allclose(dist[rows], d, rtol=2e-6, atol=1e-12)
    # no repeated neighbour inside a row
    s = np.sort(idx[rows], axis=1)
    assert (np.diff(s, axis=1) != 0).all()


def test_cfg2_10m_uniform_k10_every_query(knn, oracle):
    """configs[1]: 10 M uniform, k = 10 — every query against the host kd-tree oracle."""
    x = datasets.uniform(10_000_000, seed=42)
    idx, dist = knn.build(x).search(10)
    _properties(x, idx, dist, 10)
    ref_idx, ref_dist = oracle.knn_kdtree(x, 10)
    assert_knn_equal(idx, dist, ref_idx, ref_dist, "cfg2")
    st = knn.stats()
    assert st["rounds"] <= 4


def test_cfg3_10m_lidar_k64_sampled(knn, oracle):
    """configs[2]: 10 M clustered LiDAR-like, k = 64 — properties on all queries, 300 K sampled queries vs the oracle."""
    x = datasets.lidar_like(10_000_000, seed=7)
    idx, dist = knn.build(x).search(64)
    _properties(x, idx, dist, 64)
    tree = oracle.KdTree(x)
    ids = np.sort(np.random.default_rng(1).choice(x.shape[0], 300_000, replace=False)).astype(np.int32)
    ri, rd = tree.query(x[ids], 64, self_ids=ids)
    tree.close()
    assert_knn_equal(idx[ids], dist[ids], ri, rd, "cfg3 sample")


def test_cfg4_100m_shard_sampled(knn, oracle):
    """configs[3] shape: 100 M uniform points, BVH replicated, one of 8 query shards on this GPU; sampled
    queries against the GPU brute-force kernel and a CPU brute-force subsample."""
    import torch

    n, k = 100_000_000, 10
    xd = torch.empty((n, 3), dtype=torch.float32, device="cuda")
    knn.generate_uniform(42, 0, n, out=xd)
    knn.build(xd)
    qid, idx, dist = knn.search_shard(k, 3, 8)
    m = int(qid.shape[0])
    assert abs(m - n // 8) <= 32
    pick = torch.from_numpy(np.random.default_rng(3).choice(m, 2048, replace=False)).cuda()
    ids = qid[pick].contiguous()
    bi, bd = knn.brute_force(ids, k)
    assert torch.equal(bi, idx[pick]) and torch.equal(bd, dist[pick])
    # CPU brute force on a handful of those queries against all 100 M points
    sub = ids[:48].cpu().numpy()
    x = xd.cpu().numpy()
    ci, cd = oracle.knn_brute_queries(x, x[sub], k, self_ids=sub)
    assert (bi[:48].cpu().numpy() == ci).all() and (bd[:48].cpu().numpy() == cd).all()
    assert (x[:1000] == datasets.uniform(1000, seed=42)).all()      # device generator == host generator
