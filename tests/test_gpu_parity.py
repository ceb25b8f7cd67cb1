"""GPU parity: the CUDA path through the C ABI vs the CPU oracle, bit-exact on indices."""
import numpy as np
import pytest

from helpers import assert_knn_equal, rows_sorted
from owlraytracing_b200 import TrueKNNError, datasets

pytestmark = pytest.mark.gpu


def test_cfg1_100k_uniform_k5(knn, oracle):
    """BASELINE.json configs[0]: 100K uniform, k=5, every query vs the host oracle."""
    x = datasets.uniform(100_000, seed=42)
    idx, dist = knn.build(x).search(5)
    ref_idx, ref_dist = oracle.knn_kdtree(x, 5)
    assert_knn_equal(idx, dist, ref_idx, ref_dist, "cfg1")
    st = knn.stats()
    assert st["rounds"] >= 1 and st["n_queries"] == 100_000
    # brute force on a subset pins the kd-tree oracle itself at this size
    sub = np.arange(0, 100_000, 97, dtype=np.int32)
    bi, bd = oracle.knn_brute_queries(x, x[sub], 5, self_ids=sub)
    assert (bi == ref_idx[sub]).all() and (bd == ref_dist[sub]).all()


@pytest.mark.parametrize("k", [1, 2, 7, 10, 16, 33, 64])
def test_uniform_various_k(knn, oracle, k):
    x = datasets.uniform(30_000, seed=3)
    idx, dist = knn.build(x).search(k)
    ref = oracle.knn_kdtree(x, k)
    assert_knn_equal(idx, dist, *ref, f"k={k}")
    assert rows_sorted(dist)


@pytest.mark.parametrize("m,k", [(6, 6), (7, 18), (8, 26), (5, 7)])
def test_lattice_known_answers(knn, oracle, m, k):
    """Unit lattice: shells at distance 1 (6), sqrt2 (12), sqrt3 (8): every shell is an index tie set."""
    x = datasets.lattice(m)
    idx, dist = knn.build(x).search(k)
    ref = oracle.knn_brute(x, k)
    assert_knn_equal(idx, dist, *ref, f"lattice {m}^3 k={k}")
    centre = (m // 2) * (m * m + m + 1)
    if m >= 5 and k >= 6:
        assert (dist[centre][:6] == 1.0).all()
        # lowest-index tie-break: the six unit neighbours appear in ascending index order
        assert (np.diff(idx[centre][:6]) > 0).all()


def test_duplicates_and_coincident_points(knn, oracle):
    rng = np.random.default_rng(5)
    base = rng.random((3000, 3), dtype=np.float32)
    x = np.concatenate([base, base[:1500], base[:700], np.tile(base[:1], (200, 1))]).astype(np.float32)
    for k in (3, 10, 40):
        idx, dist = knn.build(x).search(k)
        assert_knn_equal(idx, dist, *oracle.knn_brute(x, k), f"duplicates k={k}")
    # self is excluded by INDEX, so a duplicate is a legal neighbour at distance 0 (deviceCode.cu:103)
    assert dist[0][0] == 0.0 and idx[0][0] != 0


def test_all_points_identical(knn, oracle):
    x = np.full((500, 3), 0.25, np.float32)
    idx, dist = knn.build(x).search(9)
    assert_knn_equal(idx, dist, *oracle.knn_brute(x, 9), "identical")
    assert (dist == 0).all()


@pytest.mark.parametrize("shape", ["line", "plane", "two_clusters"])
def test_degenerate_geometry(knn, oracle, shape):
    rng = np.random.default_rng(11)
    n = 20_000
    if shape == "line":
        t = rng.random(n, dtype=np.float32)
        x = np.stack([t, 2 * t, np.zeros_like(t)], 1)
    elif shape == "plane":
        x = np.concatenate([rng.random((n, 2), dtype=np.float32), np.full((n, 1), 3.0, np.float32)], 1)
    else:
        a = rng.normal(0, 1e-3, (n // 2, 3)).astype(np.float32)
        b = (rng.normal(0, 1e-3, (n // 2, 3)) + 1000.0).astype(np.float32)
        x = np.concatenate([a, b])
    x = np.ascontiguousarray(x, np.float32)
    idx, dist = knn.build(x).search(8)
    assert_knn_equal(idx, dist, *oracle.knn_kdtree(x, 8), shape)


def test_two_dimensional_input(knn, oracle):
    """dim == 2 => z = 0 (hostCode.cpp:114-118)."""
    rng = np.random.default_rng(2)
    xy = rng.random((10_000, 2), dtype=np.float32)
    idx, dist = knn.build(xy, dim=2).search(6)
    x3 = np.concatenate([xy, np.zeros((xy.shape[0], 1), np.float32)], 1)
    assert_knn_equal(idx, dist, *oracle.knn_kdtree(x3, 6), "2-D")
    # a strided 3-column array read as 2-D ignores the third column
    junk = np.concatenate([xy, rng.random((xy.shape[0], 1), dtype=np.float32)], 1)
    idx2, dist2 = knn.build(junk, dim=2).search(6)
    assert (idx2 == idx).all() and (dist2 == dist).all()


@pytest.mark.parametrize("n", [2, 3, 5, 31, 32, 33, 64, 65, 1000])
def test_small_and_k_equals_n_minus_1(knn, oracle, n):
    rng = np.random.default_rng(n)
    x = rng.random((n, 3), dtype=np.float32)
    for k in sorted({1, min(5, n - 1), n - 1}):
        if k > 512:
            continue
        idx, dist = knn.build(x).search(k)
        assert_knn_equal(idx, dist, *oracle.knn_brute(x, k), f"n={n} k={k}")


@pytest.mark.parametrize("r0", [1e-4, 0.003, 0.02, 0.5, float("inf"), 0.0, -1.0])
def test_start_radius_does_not_change_results(knn, oracle, r0):
    """Rounds (hostCode.cpp:285-340) are an execution strategy: any start radius gives the exact answer."""
    x = datasets.uniform(50_000, seed=9)
    ref = oracle.knn_kdtree(x, 10)
    idx, dist = knn.build(x).search(10, start_radius=r0)
    assert_knn_equal(idx, dist, *ref, f"r0={r0}")
    st = knn.stats()
    if r0 == 1e-4:
        assert st["rounds"] > 4  # doubling from a tiny radius takes several rounds
        assert st["round_queries"][0] == 50_000 and st["round_queries"][1] <= 50_000
    if r0 == float("inf"):
        assert st["rounds"] == 1


def test_clustered_lidar_like_k64(knn, oracle):
    """cfg3 shape at 300K points: skewed density, duplicates, k = 64."""
    x = datasets.lidar_like(300_000, seed=7)
    idx, dist = knn.build(x).search(64)
    assert_knn_equal(idx, dist, *oracle.knn_kdtree(x, 64), "lidar k=64")


@pytest.mark.parametrize("leaf,policy", [(4, 0), (8, 0), (16, 0), (32, 0), (32, 1), (8, 1)])
def test_leaf_options_are_result_invariant(knn, oracle, leaf, policy):
    x = datasets.lidar_like(40_000, seed=3)
    ref = oracle.knn_kdtree(x, 12)
    knn.set_option("leaf_size", leaf)
    knn.set_option("leaf_policy", policy)
    idx, dist = knn.build(x).search(12)
    assert_knn_equal(idx, dist, *ref, f"leaf={leaf} policy={policy}")


def test_counters_and_stats(knn, oracle):
    x = datasets.uniform(60_000, seed=1)
    knn.set_option("counters", 1)
    idx, dist = knn.build(x).search(10)
    assert_knn_equal(idx, dist, *oracle.knn_kdtree(x, 10), "counters build")
    st = knn.stats()
    assert st["points_tested"] >= 10 * 60_000 and st["nodes_visited"] > 0 and st["heap_inserts"] >= 10 * 60_000
    assert st["warp_point_loads"] * 32 >= st["points_tested"]
    assert st["n_leaves"] == st["n_nodes"] + 1 and st["build_ms"] > 0 and st["search_ms"] > 0


def test_query_sharded_union_equals_full(knn, oracle):
    x = datasets.uniform(70_001, seed=4)
    ref = oracle.knn_kdtree(x, 10)
    knn.build(x)
    for shards in (1, 2, 3, 8):
        seen = np.zeros(x.shape[0], bool)
        for s in range(shards):
            qid, idx, dist = knn.search_shard(10, s, shards)
            assert not seen[qid].any()
            seen[qid] = True
            assert_knn_equal(idx, dist, ref[0][qid], ref[1][qid], f"shard {s}/{shards}")
        assert seen.all()


def test_separate_query_set(knn, oracle):
    rng = np.random.default_rng(8)
    x = datasets.uniform(40_000, seed=6)
    q = (rng.random((5_003, 3), dtype=np.float32) * 1.2 - 0.1).astype(np.float32)  # some outside the scene box
    knn.build(x)
    idx, dist = knn.query(q, 9)
    assert_knn_equal(idx, dist, *oracle.knn_brute_queries(x, q, 9), "query set")
    # data points as queries with their own index excluded == all-points search
    sub = rng.choice(x.shape[0], 3000, replace=False).astype(np.int32)
    idx, dist = knn.query(x[sub], 9, self_ids=sub)
    full = oracle.knn_kdtree(x, 9)
    assert_knn_equal(idx, dist, full[0][sub], full[1][sub], "query with self ids")
    # per-query radius caps (boundary queries of the point-partitioned driver)
    rad2 = (rng.random(q.shape[0], dtype=np.float32) * 0.05) ** 2
    idx, dist = knn.query(q, 9, init_radius2=rad2)
    for i in range(0, q.shape[0], 50):
        ri, rd = oracle.knn_brute_queries(x, q[i:i + 1], 9, radius2=rad2[i])
        assert (idx[i] == ri[0]).all() and (dist[i] == rd[0]).all()


@pytest.mark.parametrize("cloud", ["uniform", "duplicates", "lidar"])
def test_range_count(knn, oracle, cloud):
    """Fixed-radius neighbour counts (DBSCAN core test): self excluded by index, coincident duplicates counted."""
    if cloud == "uniform":
        x = datasets.uniform(8_000, seed=12)
    elif cloud == "duplicates":
        base = datasets.uniform(6_000, seed=13)
        x = np.ascontiguousarray(np.concatenate([base, base[:1500], base[:300]], 0))
    else:
        x = datasets.lidar_like(9_000, seed=4)
    knn.build(x)
    scale = 1.0 if cloud != "lidar" else 40.0
    for r in (0.0, 0.01 * scale, 0.05 * scale, 0.2 * scale):
        got = knn.range_count(r)
        assert (np.asarray(got).astype(np.uint32) == oracle.range_count(x, r)).all(), r


def test_gpu_brute_force_second_oracle(knn, oracle):
    x = datasets.lidar_like(50_000, seed=2)
    knn.build(x)
    ids = np.random.default_rng(0).choice(x.shape[0], 777, replace=False).astype(np.int32)
    idx, dist = knn.brute_force(ids, 20)
    ref = oracle.knn_kdtree(x, 20)
    assert_knn_equal(idx, dist, ref[0][ids], ref[1][ids], "gpu brute force")


def test_estimated_radius_is_plausible(knn):
    x = datasets.uniform(200_000, seed=42)
    knn.build(x)
    r = knn.estimate_start_radius(10)
    expect = (10 / (4.0 / 3.0 * np.pi * 200_000)) ** (1 / 3)
    assert 0.5 * expect < r < 3 * expect


def test_device_generator_matches_numpy(knn):
    a = knn.generate_uniform(42, 1000, 5000)
    assert (a == datasets.uniform(5000, seed=42, first=1000)).all()


def test_errors(knn):
    x = datasets.uniform(100, seed=1)
    with pytest.raises(TrueKNNError) as e:
        knn.search(3)
    assert e.value.code == 5  # TKNN_ESTATE: search before build
    knn.build(x)
    for bad_k in (0, 100, 513):
        with pytest.raises(TrueKNNError) as e:
            knn.search(bad_k)  # the reference never terminates for k > n-1 (hostCode.cpp:285,321-323)
        assert e.value.code == 1
    y = x.copy()
    y[7, 1] = np.nan
    with pytest.raises(TrueKNNError) as e:
        knn.build(y)
    assert e.value.code == 1
    with pytest.raises(TrueKNNError):
        knn.build(x[:1])
    with pytest.raises(TrueKNNError):
        knn.build(x, dim=4)


def test_rebuild_many_times(knn, oracle):
    """Shape of the reference's tests/t02-group-rebuilds: rebuild repeatedly, stay correct, no leak."""
    rng = np.random.default_rng(0)
    for it in range(30):
        n = int(rng.integers(50, 5000))
        x = rng.random((n, 3), dtype=np.float32)
        idx, dist = knn.build(x).search(4)
        if it % 10 == 0:
            assert_knn_equal(idx, dist, *oracle.knn_brute(x, 4), f"rebuild {it}")


def test_torch_device_tensors(knn, oracle):
    import torch

    x = datasets.uniform(20_000, seed=5)
    xd = torch.from_numpy(x).cuda()
    knn.set_stream(torch.cuda.current_stream().cuda_stream)
    idx, dist = knn.build(xd).search(10)
    assert idx.is_cuda and dist.is_cuda
    assert_knn_equal(idx.cpu().numpy(), dist.cpu().numpy(), *oracle.knn_kdtree(x, 10), "torch tensors")


def test_merge_topk_and_morton_codes_match_the_cpu_stand_in(knn):
    """The two C-ABI helpers of the point-partitioned driver against the numpy stand-in the gloo tests use."""
    import torch

    from cpu_engine import CpuEngine

    rng = np.random.default_rng(4)
    x = datasets.lidar_like(20_000, seed=9)
    box = np.concatenate([x.min(0), x.max(0)]).astype(np.float32)
    assert (knn.morton_codes(x, box).astype(np.int64) == CpuEngine().morton_codes(x, box).numpy()).all()
    codes_dev = knn.morton_codes(torch.from_numpy(x).cuda(), box)
    assert (codes_dev.cpu().numpy() == CpuEngine().morton_codes(x, box).numpy()).all()
    # partial lists with ties, duplicates across parts and short lists
    parts, nq, k = 3, 500, 7
    ip = rng.integers(0, 40, (parts, nq, k)).astype(np.int32)
    dp = (rng.integers(0, 6, (parts, nq, k)) * 0.25).astype(np.float32)
    # a point has one distance: make d2 a function of the index so that duplicates agree
    dp = ((ip % 6) * 0.25).astype(np.float32)
    order = np.lexsort((ip, dp), axis=-1)  # ascending (d2, idx) inside each list
    ip, dp = np.take_along_axis(ip, order, -1), np.take_along_axis(dp, order, -1)
    for s in range(parts):  # remove in-list duplicates, pad with -1
        for q in range(nq):
            _, first = np.unique(ip[s, q], return_index=True)
            keep = np.sort(first)
            m = keep.size
            ip[s, q, :m], dp[s, q, :m] = ip[s, q, keep], dp[s, q, keep]
            ip[s, q, m:], dp[s, q, m:] = -1, np.finfo(np.float32).max
    ref_i, ref_d = CpuEngine().merge_topk(ip, dp)
    gi, gd = knn.merge_topk(torch.from_numpy(ip).cuda(), torch.from_numpy(dp).cuda())
    assert (gi.cpu().numpy() == ref_i.numpy()).all()
    assert np.allclose(gd.cpu().numpy(), ref_d.numpy(), rtol=1e-6, atol=0)


def test_cpp_cli_end_to_end(oracle, tmp_path):
    """tools/trueknn: the sample's six positional arguments (hostCode.cpp:66-73) through the C++ host."""
    import os
    import subprocess

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "tools", "trueknn")
    assert os.path.exists(exe), "build it with __graft_entry__.build()"
    x = datasets.lidar_like(6000, seed=4)
    csv = tmp_path / "pts.csv"
    with open(csv, "w") as f:
        for p in x:
            f.write(",".join(repr(float(v)) for v in p) + "\n")
    out, nn = tmp_path / "time.txt", tmp_path / "nn.csv"
    for r0 in ("0.05", "0"):
        r = subprocess.run([exe, str(csv), "5000", "3", r0, "7", str(out), "--neighbours", str(nn), "--json"],
                           capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        assert "Build time:" in r.stdout and "True KNN time:" in r.stdout and "Total time:" in r.stdout
        got = np.loadtxt(str(nn), delimiter=",")
        ref_i, ref_d = oracle.knn_kdtree(x[:5000], 7)   # "always selects first n points" (README.md:10)
        assert (got[:, 1].astype(np.int32).reshape(5000, 7) == ref_i).all()
        assert np.allclose(got[:, 2].astype(np.float32).reshape(5000, 7), ref_d, rtol=1e-6, atol=0)
    assert len(open(out).read().split()) == 2  # one total per run appended (hostCode.cpp:350-356)
    # k > n - 1: the reference loops forever; the CLI reports and exits non-zero
    r = subprocess.run([exe, str(csv), "5", "3", "0.1", "5", str(out)], capture_output=True, text=True)
    assert r.returncode != 0 and "k = 5" in r.stderr


@pytest.mark.parametrize("cloud", ["uniform", "offset_1e4", "offset_1e6_tiny_spread", "lattice", "lidar", "two_scales", "huge"])
def test_prefilter_never_rejects_a_true_candidate(knn, oracle, cloud):
    """The 3-FMA pre-filter is conservative: the counting build audits every (query, point)
    pair against the exact test; results with the filter on and off are identical and equal the oracle."""
    rng = np.random.default_rng(17)
    if cloud == "uniform":
        x = datasets.uniform(60_000, seed=2)
    elif cloud == "offset_1e4":
        x = datasets.uniform(40_000, seed=3) + np.float32(1e4)
    elif cloud == "offset_1e6_tiny_spread":
        x = (rng.random((30_000, 3)) * 8).astype(np.float32) + np.float32(1e6)   # spacing ~ a few ulps: massive ties
    elif cloud == "lattice":
        x = datasets.lattice(30) * np.float32(0.1) + np.float32(3.0)
    elif cloud == "lidar":
        x = datasets.lidar_like(80_000, seed=11)
    elif cloud == "two_scales":
        x = np.concatenate([datasets.uniform(20_000, seed=5) * np.float32(1e-3), datasets.uniform(20_000, seed=6) * np.float32(1e3)])
    else:
        x = (datasets.uniform(20_000, seed=7) - np.float32(0.5)) * np.float32(1e18)  # squares overflow fp32
    x = np.ascontiguousarray(x, np.float32)
    k = 12
    ref = oracle.knn_kdtree(x, k)
    knn.set_option("counters", 1)
    knn.set_option("approx_filter", 1)
    idx, dist = knn.build(x).search(k)
    st = knn.stats()
    assert st["filter_violations"] == 0
    assert_knn_equal(idx, dist, *ref, f"{cloud} (pre-filter on)")
    knn.set_option("approx_filter", 0)
    idx0, dist0 = knn.search(k)
    assert (idx0 == idx).all() and (dist0 == dist).all()
    for r0 in (float("inf"), 1e-3):
        knn.set_option("approx_filter", 1)
        i2, d2 = knn.search(k, start_radius=r0)
        assert knn.stats()["filter_violations"] == 0
        assert (i2 == idx).all() and (d2 == dist).all()


@pytest.mark.parametrize("k", [24, 25, 100, 256, 512])
def test_large_k_heap_path(knn, oracle, k):
    """k > 24 switches the per-lane k-list from the ascending list to the binary heap; k = 512 is TKNN_MAX_K."""
    x = datasets.lidar_like(6000, seed=13)
    ref = oracle.knn_kdtree(x, k)
    for r0 in (0.0, 0.05):
        idx, dist = knn.build(x).search(k, start_radius=r0)
        assert_knn_equal(idx, dist, *ref, f"k={k} r0={r0}")


def test_pipelined_host_output_of_a_shard(knn, oracle):
    """tknn_search_shard with HOST outputs searches the shard in Morton slices and overlaps each slice's
    device->host copy with the next slice's search; results equal the single-pass device-output path."""
    import torch

    x = datasets.uniform(700_001, seed=21)
    ref = oracle.knn_kdtree(x, 10)
    knn.build(x)                                   # numpy => host outputs => pipelined (n >= 2^18)
    qid, idx, dist = knn.search_shard(10, 0, 1)
    assert np.array_equal(np.sort(qid), np.arange(x.shape[0]))
    assert_knn_equal(idx, dist, ref[0][qid], ref[1][qid], "pipelined host output")
    for chunks in (1, 3, 64):
        knn.set_option("output_chunks", chunks)
        q2, i2, d2 = knn.search_shard(10, 1, 2)
        assert_knn_equal(i2, d2, ref[0][q2], ref[1][q2], f"chunks={chunks}")
    knn.build(torch.from_numpy(x).cuda())          # device outputs: one pass
    q3, i3, d3 = knn.search_shard(10, 0, 1)
    assert np.array_equal(q3.cpu().numpy(), qid) and np.array_equal(i3.cpu().numpy(), idx)


def test_pipelined_host_output_in_file_order(knn, oracle):
    """tknn_search with HOST outputs slices the queries by original index (contiguous, final output rows per
    slice) and overlaps each slice's copy with the next slice's search; any slice count gives the exact answer."""
    x = datasets.lidar_like(400_000, seed=23)
    ref = oracle.knn_kdtree(x, 8)
    knn.build(x)
    for chunks in (4, 1, 3, 64):
        knn.set_option("file_order_chunks", chunks)
        idx, dist = knn.search(8)
        assert_knn_equal(idx, dist, *ref, f"file-order chunks={chunks}")
    idx, dist = knn.search(8, start_radius=0.02)
    assert_knn_equal(idx, dist, *ref, "file-order chunks, fixed radius")


@pytest.mark.parametrize("k", [4, 10, 40])
def test_big_cluster_of_coincident_points(knn, oracle, k):
    """Sensor-style degenerate data: thousands of returns at one location.  Exactness needs the k LOWEST indices
    of the cluster; the tie-pruning kernel variant (index-aware) keeps it fast: every query in or
    next to the cluster ties exactly at its bound with every cluster leaf."""
    rng = np.random.default_rng(k)
    x = datasets.uniform(60_000, seed=31)
    dup = rng.choice(60_000, 9_000, replace=False)          # scattered file positions, one location
    x[dup] = x[dup[0]]
    x[rng.choice(60_000, 50, replace=False)] = x[dup[0]] + np.float32(1e-4)   # close neighbours of the cluster
    ref = oracle.knn_kdtree(x, k)
    idx, dist = knn.build(x).search(k)
    assert_knn_equal(idx, dist, *ref, f"coincident cluster k={k}")
    st = knn.stats()
    assert st["search_ms"] < 200.0                            # 9 000 duplicates took seconds without tie pruning
    q = x[dup[:100]] + np.float32(3e-5)
    qi, qd = knn.query(q, k)
    assert_knn_equal(qi, qd, *oracle.knn_brute_queries(x, q, k), "queries next to the cluster")


def test_reach_mask_kernel_is_conservative_and_tight(knn):
    """tknn_reach_mask (boundary-query routing of the point-partitioned driver): never misses a rank whose
    summary box the ball really touches, and flags nothing the all-cells reference would not."""
    import torch

    rng = np.random.default_rng(12)
    w, bits = 5, 3
    x = datasets.lidar_like(40_000, seed=21)
    box = np.concatenate([x.min(0), x.max(0)]).astype(np.float32)
    codes = knn.morton_codes(x, box).astype(np.int64)
    owner = rng.integers(0, w, x.shape[0])
    cell = codes >> (63 - 3 * bits)
    summ = np.empty((w, 1 << (3 * bits), 6), np.float32)
    summ[:, :, :3], summ[:, :, 3:] = np.inf, -np.inf
    for s in range(w):
        for c in np.unique(cell[owner == s]):
            p = x[(owner == s) & (cell == c)]
            summ[s, c, :3], summ[s, c, 3:] = p.min(0), p.max(0)
    q = x[rng.choice(x.shape[0], 5000, replace=False)]
    r2 = (rng.random(5000).astype(np.float32) * 3.0) ** 2
    r2[:50] = np.inf
    r2[50:100] = 0.0
    got = knn.reach_mask(torch.from_numpy(q).cuda(), torch.from_numpy(r2).cuda(), box, torch.from_numpy(summ).cuda(), 2, bits)
    got = got.cpu().numpy()
    p64, lo, hi = q.astype(np.float64)[:, None, None, :], summ[None, :, :, :3].astype(np.float64), summ[None, :, :, 3:].astype(np.float64)
    with np.errstate(invalid="ignore"):
        d2 = (np.maximum(np.maximum(lo - p64, p64 - hi), 0.0) ** 2).sum(-1)           # [nq, w, cells]
    d2 = np.where(np.isnan(d2), np.inf, d2)
    exact = (d2 <= r2.astype(np.float64)[:, None, None]).any(-1)
    loose = (d2 <= (r2.astype(np.float64) * (1 + 1e-4) + 1e-20)[:, None, None]).any(-1)
    exact[:, 2] = loose[:, 2] = False                                                 # never the asking rank itself
    bits_got = ((got[:, None] >> np.arange(w)[None]) & 1).astype(bool)
    assert (bits_got | ~exact).all(), "a reachable rank was missed"
    assert (~bits_got | loose).all(), "a rank out of reach was flagged"


def _kernel_choice_clouds():
    rng = np.random.default_rng(21)
    base = rng.random((4000, 3), dtype=np.float32)
    return {
        "uniform": datasets.uniform(12_000, seed=9),
        "lidar": datasets.lidar_like(15_000, seed=5),
        "duplicates": np.ascontiguousarray(np.concatenate([base, base[:2000], np.tile(base[:1], (300, 1))]), np.float32),
        "lattice": datasets.lattice(12),
    }


@pytest.mark.parametrize("cloud", ["uniform", "lidar", "duplicates", "lattice"])
@pytest.mark.parametrize("k", [1, 10, 25, 33, 64, 100])
def test_warp_per_query_kernel_every_round(oracle, cloud, k):
    """TKNN_OPT_WARP_ROUND_MAX above the cloud size: every round (and the start-radius sample) runs the
    warp-per-query kernel; a small start radius forces several rounds.  Same exact results."""
    from owlraytracing_b200 import TrueKNN

    x = _kernel_choice_clouds()[cloud]
    ref = oracle.knn_kdtree(x, k)
    with TrueKNN(0, warp_round_max=1 << 30) as t:
        idx, dist = t.build(x).search(k)
        assert_knn_equal(idx, dist, *ref, f"warp kernel {cloud} k={k} auto radius")
        r_small = 0.5 * float(np.median(ref[1][:, 0])) + 1e-6    # most queries miss even their first neighbour
        idx, dist = t.search(k, r_small)
        assert_knn_equal(idx, dist, *ref, f"warp kernel {cloud} k={k} small radius")
        assert t.stats()["rounds"] >= 2


@pytest.mark.parametrize("spec", [0, 1 << 20])
@pytest.mark.parametrize("wmax", [0, 49152, 1 << 30])
def test_round_kernels_agree(oracle, wmax, spec):
    """The three round kernels (cooperative, thread-per-query, warp-per-query) are interchangeable: the
    result does not depend on which of them a round ran on — nor on whether round 2 was launched speculatively as the
    final round (TKNN_OPT_SPECULATIVE_MAX; with the warp kernel allowed for every round size it then covers ALL
    leftovers, so the search ends after two rounds) or the host waited for the count and kept doubling."""
    from owlraytracing_b200 import TrueKNN

    x = datasets.lidar_like(120_000, seed=8)
    ref = oracle.knn_kdtree(x, 12)
    with TrueKNN(0, warp_round_max=wmax, speculative_max=spec) as t:
        idx, dist = t.build(x).search(12, 0.02)                   # many rounds: 0.02 m against a 200 m scene
        assert_knn_equal(idx, dist, *ref, f"warp_round_max={wmax} speculative_max={spec}")
        assert t.stats()["rounds"] >= (2 if (spec and wmax == 1 << 30) else 4)
        idx, dist = t.search(12)                                  # estimated radius: read from the device by round 1
        assert_knn_equal(idx, dist, *ref, f"auto radius, warp_round_max={wmax} speculative_max={spec}")
        assert t.stats()["start_radius"] > 0 and t.stats()["start_radius"] == pytest.approx(t.estimate_start_radius(12))
        if wmax == 0:
            t.set_option("counters", 1)
            t.search(12)
            assert t.stats()["heap_inserts"] >= 12 * x.shape[0]


@pytest.mark.parametrize("case", ["uniform", "lidar", "lattice", "dups"])
def test_committed_golden_vectors(knn, case):
    """The CUDA path against the committed fixtures (tests/golden/knn_golden.npz, written by the brute-force
    oracle through tests/golden/make_golden.py): indices bit-exact, distances bit-exact (sqrtf of the same d2)."""
    import os

    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "knn_golden.npz"))
    x, k = np.ascontiguousarray(g[f"{case}_x"], np.float32), int(g[f"{case}_k"])
    idx, dist = knn.build(x).search(k)
    assert (idx == g[f"{case}_idx"]).all()
    assert (dist == g[f"{case}_dist"]).all()
    # and from a tiny start radius (many rounds), same answer
    idx2, dist2 = knn.search(k, 1e-4)
    assert (idx2 == g[f"{case}_idx"]).all() and (dist2 == g[f"{case}_dist"]).all()


def test_non_finite_queries_are_rejected(knn):
    """tknn_query with a NaN / inf coordinate returns TKNN_EINVAL instead of a silent row of -1."""
    from owlraytracing_b200 import TrueKNNError

    x = datasets.uniform(5000, seed=4)
    knn.build(x)
    q = datasets.uniform(200, seed=5)
    idx, dist = knn.query(q, 5)
    assert (idx >= 0).all()
    for bad in (np.nan, np.inf):
        qb = q.copy()
        qb[17, 1] = bad
        with pytest.raises(TrueKNNError) as e:
            knn.query(qb, 5)
        assert e.value.code == 1
    idx2, _ = knn.query(q, 5)          # the context stays usable
    assert (idx2 == idx).all()


def test_shared_memory_bandwidth_probe(knn):
    """bench.py's roofline denominator: conflict-free LDS.128 delivers the 128 B/clk/SM crossbar (about 37 TB/s on a B200),
    warp-broadcast LDS.128 about twice that in bytes delivered to lanes."""
    m = knn.measure_smem_bandwidth()
    assert 20_000 < m["smem_conflict_free_gbs"] < 45_000
    assert m["smem_broadcast_gbs"] > m["smem_conflict_free_gbs"]


@pytest.mark.parametrize("curve", [0, 1])
def test_space_filling_curve_is_result_invariant(oracle, curve):
    """TKNN_OPT_CURVE: Hilbert (default) or Morton keys order the points and the tree's children differently; the k
    nearest neighbours do not depend on it, and the Hilbert order needs fewer node visits and point tests per query."""
    from owlraytracing_b200 import TrueKNN

    x = datasets.lidar_like(150_000, seed=12)
    ref = oracle.knn_kdtree(x, 10)
    with TrueKNN(0, curve=curve, counters=1) as t:
        idx, dist = t.build(x).search(10)
        assert_knn_equal(idx, dist, *ref, f"curve={curve}")
        qi, qd = t.query(x[:5000], 10, self_ids=np.arange(5000, dtype=np.int32))   # queries are coded on the same curve
        assert_knn_equal(qi, qd, ref[0][:5000], ref[1][:5000], f"query, curve={curve}")


@pytest.mark.parametrize("team", [0, 4, 8, 16])
@pytest.mark.parametrize("cloud,k", [("uniform", 10), ("lidar", 33), ("duplicates", 5)])
def test_sparse_round_kernels_agree(oracle, team, cloud, k):
    """Sparse rounds (more leftovers than the warp-per-query threshold, fewer than n / 8): one thread per query
    (TKNN_OPT_SPARSE_TEAM = 0) or teams of 4 / 8 / 16 lanes per query give the same exact answer.  warp_round_max = 64
    and a small start radius push most later rounds through the sparse kernel."""
    from owlraytracing_b200 import TrueKNN

    x = _kernel_choice_clouds()[cloud]
    ref = oracle.knn_kdtree(x, k)
    r_small = float(np.quantile(ref[1][:, -1], 0.96))                 # ~4 % of the queries miss in round 1
    with TrueKNN(0, warp_round_max=64, speculative_max=0, sparse_team=team, counters=1) as t:
        idx, dist = t.build(x).search(k, r_small)
        st = t.stats()
        assert_knn_equal(idx, dist, *ref, f"sparse_team={team} {cloud} k={k}")
        assert st["rounds"] >= 2 and 64 < st["round_queries"][1] * 1 and st["round_queries"][1] * 8 <= x.shape[0]


@pytest.mark.parametrize("where", ["host", "device"])
def test_indices_only_output(knn, oracle, where):
    """dist_out = NULL (tknn_search): the same indices, no distances copied; the sliced host-output path included."""
    import torch

    n, k = 300_000, 7
    x = datasets.uniform(n, seed=11)
    ref_idx, _ = oracle.knn_kdtree(x, k)
    if where == "host":
        knn.build(x)
        idx, dist = knn.search(k, indices_only=True)
        assert dist is None
        assert knn.stats()["d2h_bytes"] == n * k * 4
        assert (idx == ref_idx).all()
    else:
        xd = torch.from_numpy(x).cuda()
        knn.build(xd)
        idx, dist = knn.search(k, indices_only=True)
        assert dist is None and idx.is_cuda
        assert (idx.cpu().numpy() == ref_idx).all()
