"""GPU tests of the multi-GPU drivers INSIDE the library (include/trueknn.h "multi-GPU", csrc/dist.cu), run on
ONE GPU: a device list that names device 0 several times runs that many ranks on it over the in-process transport
(NCCL refuses two ranks on one GPU), so the whole pipelines — slice upload + all-gather + replicated build, Morton-
range redistribution, local LBVH with global ids, summaries, boundary-query exchange, remote capped search, merges,
peer scatter into file-order arrays — are exercised by the single-GPU suite.  The NCCL transport itself is covered by
tools/multigpu_check.py on a multi-GPU box (profiles/r2_multigpu_check_*.log)."""
import os
import subprocess
import threading

import numpy as np
import pytest

from helpers import assert_knn_equal
from owlraytracing_b200 import MultiTrueKNN, TrueKNNError, datasets

pytestmark = pytest.mark.gpu


def _cloud(kind, n):
    if kind == "uniform":
        return datasets.uniform(n, seed=21)
    if kind == "lidar":
        return datasets.lidar_like(n, seed=9)
    if kind == "dups":      # every point twice + one cluster of 300 coincident points: ties across rank borders
        base = datasets.uniform(n // 2 - 150, seed=5)
        return np.ascontiguousarray(np.concatenate([base, base, np.tile(base[:1], (300, 1))]), np.float32)
    if kind == "lattice":
        return datasets.lattice(28)[:n]
    if kind == "planar":    # 2-D data: z = 0 everywhere
        x = datasets.uniform(n, seed=6)
        x[:, 2] = 0
        return x
    raise ValueError(kind)


@pytest.mark.parametrize("ranks", [1, 2, 3, 5])
@pytest.mark.parametrize("kind", ["uniform", "lidar"])
def test_query_sharded_in_library(oracle, ranks, kind):
    x = _cloud(kind, 40_000)
    k = 10
    with MultiTrueKNN([0] * ranks, "shard") as m:
        idx, dist = m.build(x).search(k)
        sizes = [m.rank(r).stats()["n_queries"] for r in range(ranks)]
    ref_i, ref_d = oracle.knn_kdtree(x, k)
    assert_knn_equal(idx, dist, ref_i, ref_d, f"shard x{ranks} {kind}")
    assert sum(sizes) == x.shape[0] and max(sizes) - min(sizes) <= 32   # contiguous Morton slices of 32-query groups


@pytest.mark.parametrize("ranks", [1, 2, 4, 7])
@pytest.mark.parametrize("kind,k", [("uniform", 10), ("lidar", 16), ("dups", 4), ("lattice", 6), ("planar", 8)])
def test_point_partitioned_in_library(oracle, ranks, kind, k):
    x = _cloud(kind, 21_000)
    with MultiTrueKNN([0] * ranks, "partition") as m:
        idx, dist = m.build(x).search(k)
        st = [m.rank(r).dist_stats() for r in range(ranks)]
    ref_i, ref_d = oracle.knn_kdtree(x, k)
    assert_knn_equal(idx, dist, ref_i, ref_d, f"partition x{ranks} {kind}")
    assert sum(s["n_owned"] for s in st) == x.shape[0]          # every point has exactly one owner
    assert all(s["n_global"] == x.shape[0] for s in st)
    if ranks > 1:
        sent = sum(s["boundary_sent"] for s in st)
        assert sent == sum(s["boundary_received"] for s in st)
        assert sent > 0                                           # the exchange really happened
        if kind == "uniform":
            assert sent < x.shape[0] // 2                         # ... for a minority of the queries


@pytest.mark.parametrize("r0", [1e-3, 0.05, float("inf")])
def test_partitioned_start_radius_is_result_invariant(oracle, r0):
    x = _cloud("uniform", 15_000)
    with MultiTrueKNN([0, 0, 0], "partition") as m:
        idx, dist = m.build(x).search(7, r0)
    ref_i, ref_d = oracle.knn_kdtree(x, 7)
    assert_knn_equal(idx, dist, ref_i, ref_d, f"partition r0={r0}")


def _run_ranks(m, fn):
    """Drive the per-rank C ABI (tknn_partition_*) from one Python thread per rank: ctypes releases the GIL."""
    out, err = [None] * m.n_ranks, [None] * m.n_ranks

    def body(r):
        try:
            out[r] = fn(r, m.rank(r))
        except Exception as e:  # noqa: BLE001
            err[r] = e

    th = [threading.Thread(target=body, args=(r,)) for r in range(m.n_ranks)]
    for t in th:
        t.start()
    for t in th:
        t.join(timeout=300)
    assert all(not t.is_alive() for t in th), "a rank hangs"
    for e in err:
        if e is not None:
            raise e
    return out


@pytest.mark.parametrize("kind", ["uniform", "lidar"])
def test_per_rank_partition_api_and_distributed_verification(oracle, kind):
    """The one-rank-per-process entry points (what bench.py's cfg5 uses under torchrun), driven per rank: arbitrary
    index slices in, sharded results out, and tknn_partition_verify — every rank brute-forces all sampled queries
    against its own points, partial lists are all-gathered and merged — agrees, and notices a planted error."""
    import torch

    n, k, ranks = 60_000, 10, 4
    x = _cloud(kind, n)
    cuts = [0, 9_000, 31_000, 31_500, n]     # deliberately uneven starting slices
    ref_i, ref_d = oracle.knn_kdtree(x, k)
    with MultiTrueKNN([0] * ranks, "partition") as m:
        def work(r, eng):
            lo, hi = cuts[r], cuts[r + 1]
            xd = torch.from_numpy(x[lo:hi]).cuda()
            eng.partition_build(xd, lo)
            gid, idx, dist = eng.partition_search(k)
            ok = eng.partition_verify(k, 256, gid, idx, dist)
            # plant one wrong neighbour in a sampled row of rank 2 (sample t -> row rows * t / samples; t = 1)
            if r == 2:
                row = gid.shape[0] * 1 // 256
                idx[row, k - 1] = idx[row, k - 1] ^ 1
            bad = eng.partition_verify(k, 256, gid, idx, dist)
            if r == 2:
                idx[row, k - 1] = idx[row, k - 1] ^ 1
            host = eng.partition_verify(k, 64, gid.cpu().numpy(), idx.cpu().numpy(), dist.cpu().numpy())  # host arrays too
            return gid.cpu().numpy(), idx.cpu().numpy(), dist.cpu().numpy(), ok, bad, host, eng.dist_stats()

        res = _run_ranks(m, work)
    owned = np.concatenate([r[0] for r in res])
    assert np.array_equal(np.sort(owned), np.arange(n))
    for gid, idx, dist, ok, bad, host, st in res:
        assert_knn_equal(idx, dist, ref_i[gid], ref_d[gid], f"per-rank partition {kind}")
        assert ok == (ranks * 256, 0)
        assert bad == (ranks * 256, 1)
        assert host == (ranks * 64, 0)
        assert st["n_ranks"] == ranks and st["search_total_ms"] > 0


def test_replicated_build_from_slices_per_rank(oracle):
    """tknn_build_replicated + tknn_search_shard per rank (bench.py's cfg4 path under torchrun)."""
    n, k, ranks = 50_001, 10, 3
    x = _cloud("uniform", n)
    per = (n + ranks - 1) // ranks
    ref_i, ref_d = oracle.knn_kdtree(x, k)
    with MultiTrueKNN([0] * ranks, "shard") as m:
        def work(r, eng):
            lo = min(n, per * r)
            eng.build_replicated(x[lo: min(n, lo + per)], lo, n)
            qid, idx, dist = eng.search_shard(k, r, ranks)
            return np.asarray(qid), np.asarray(idx), np.asarray(dist), eng.dist_stats()

        res = _run_ranks(m, work)
        # a slice that does not tile the cloud is refused by EVERY rank (nobody is left waiting in a collective)
        def wrong(r, eng):
            lo = min(n, per * r)
            try:
                eng.build_replicated(x[lo: min(n, lo + per) - (1 if r == 1 else 0)], lo, n)
            except TrueKNNError as e:
                return e.code
            return 0

        codes = _run_ranks(m, wrong)
    assert all(c == 1 for c in codes)  # TKNN_EINVAL everywhere
    seen = np.zeros(n, bool)
    for qid, idx, dist, st in res:
        assert not seen[qid].any()
        seen[qid] = True
        assert_knn_equal(idx, dist, ref_i[qid], ref_d[qid], "replicated build")
        assert st["allgather_ms"] >= 0 and st["n_global"] == n
    assert seen.all()


def test_partition_rejects_what_it_cannot_answer():
    x = _cloud("uniform", 64)
    with MultiTrueKNN([0, 0, 0, 0], "partition") as m:
        m.build(x)
        with pytest.raises(TrueKNNError):      # k >= points owned by a rank: refused by all ranks together
            m.search(40)
        idx, dist = m.search(3)                # and the handle stays usable
        assert (idx >= 0).all()
    with MultiTrueKNN([0, 0], "partition") as m:
        with pytest.raises(TrueKNNError):      # all points in one cell: a rank would own nothing
            m.build(np.zeros((100, 3), np.float32))
        bad = x.copy()
        bad[5, 1] = np.nan
        with pytest.raises(TrueKNNError):
            m.build(bad)
    with pytest.raises(TrueKNNError):
        MultiTrueKNN([0, 99], "shard")         # no such device


@pytest.mark.parametrize("mode", ["shard", "partition"])
def test_cpp_cli_multi_gpu_modes(oracle, tmp_path, mode):
    """tools/trueknn --devices 0,0 --mode shard|partition: the C++ host drives all ranks through tknn_create_multi."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "tools", "trueknn")
    assert os.path.exists(exe), "build it with __graft_entry__.build()"
    x = datasets.uniform(20_000, seed=33)
    pts = tmp_path / "pts.f32"
    x.tofile(str(pts))
    out, nn = tmp_path / "time.txt", tmp_path / "nn.csv"
    r = subprocess.run([exe, str(pts), "20000", "3", "0", "10", str(out), "--neighbours", str(nn), "--json", "--devices", "0,0",
                        "--mode", mode], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    assert "Build time:" in r.stdout and "True KNN time:" in r.stdout and '"gpus": 2' in r.stdout
    got = np.loadtxt(str(nn), delimiter=",")
    ref_i, ref_d = oracle.knn_kdtree(x, 10)
    assert (got[:, 1].astype(np.int32).reshape(20000, 10) == ref_i).all()
    assert np.allclose(got[:, 2].astype(np.float32).reshape(20000, 10), ref_d, rtol=1e-6, atol=0)
