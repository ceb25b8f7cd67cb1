"""Shared assertions for the parity tests."""
import numpy as np

REL_TOL = 1e-6  # north star: fp32 distances within 1e-6 relative; indices bit-exact


def assert_knn_equal(idx, dist, ref_idx, ref_dist, what=""):
    idx, dist = np.asarray(idx), np.asarray(dist)
    assert idx.shape == ref_idx.shape, (what, idx.shape, ref_idx.shape)
    bad = np.nonzero((idx != ref_idx).any(axis=1))[0]
    assert bad.size == 0, (f"{what}: {bad.size} / {idx.shape[0]} queries differ in neighbour indices; first row {bad[0]}: "
                           f"got {idx[bad[0]]} want {ref_idx[bad[0]]} (dist {dist[bad[0]]} vs {ref_dist[bad[0]]})")
    filled = ref_idx >= 0
    err = np.abs(dist[filled] - ref_dist[filled])
    assert (err <= REL_TOL * np.abs(ref_dist[filled])).all(), f"{what}: distance error {err.max()}"
    assert (dist[~filled] == ref_dist[~filled]).all(), f"{what}: sentinel distances differ"


def rows_sorted(dist):
    d = np.asarray(dist)
    return bool((np.diff(d, axis=1) >= 0).all())
