"""ctypes binding of oracle/liboracle.so — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import
this module.  Nothing under owlraytracing_b200/ does.  See knn_oracle.c for what each function
restates (reference file:line) and for the "parity unpinned" statement.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")
_lib = None


def build(force: bool = False) -> str:
    """Compile liboracle.so with the committed Makefile (gcc + OpenMP)."""
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(
        os.path.join(_HERE, "knn_oracle.c")
    ):
        subprocess.run(["make", "-C", _HERE, "-s"], check=True)
    return _LIB_PATH


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        L = C.CDLL(_LIB_PATH)
        f32p, i32p, u32p = C.POINTER(C.c_float), C.POINTER(C.c_int32), C.POINTER(C.c_uint32)
        L.tko_num_threads.restype = C.c_int
        L.tko_dist2.restype = C.c_float
        L.tko_dist2.argtypes = [f32p, f32p]
        L.tko_knn_brute.argtypes = [f32p, C.c_int64, C.c_int, i32p, f32p]
        L.tko_knn_brute_queries.argtypes = [f32p, C.c_int64, f32p, C.c_int64, i32p, C.c_int, C.c_float, f32p, i32p, f32p]
        L.tko_set_squared_output.argtypes = [C.c_int]
        L.tko_knn_kdtree.argtypes = [f32p, C.c_int64, C.c_int, i32p, f32p, C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.tko_kdtree_build.restype = C.c_void_p
        L.tko_kdtree_build.argtypes = [f32p, C.c_int64, C.c_int]
        L.tko_kdtree_free.argtypes = [C.c_void_p]
        L.tko_kdtree_knn.argtypes = [C.c_void_p, f32p, C.c_int64, i32p, C.c_int, C.c_float, i32p, f32p]
        L.tko_kdtree_knn_positions.argtypes = [C.c_void_p, C.POINTER(C.c_int64), C.c_int64, C.c_int, C.c_float, i32p, f32p]
        L.tko_kdtree_ids.argtypes = [C.c_void_p, C.POINTER(C.c_int64), C.c_int64, i32p]
        L.tko_set_num_threads.argtypes = [C.c_int]
        L.tko_range_count.argtypes = [f32p, C.c_int64, C.c_float, u32p]
        L.tko_reference_trueknn.argtypes = [f32p, C.c_int64, C.c_int, C.c_float, C.c_int, i32p, f32p,
                                            C.POINTER(C.c_int), C.POINTER(C.c_float)]
        L.tko_parse_points.restype = C.c_int64
        L.tko_parse_points.argtypes = [C.c_char_p, C.c_int64, C.c_int64, C.c_int, f32p, C.c_int64]
        _lib = L
    return _lib


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t)) if a is not None else None


def num_threads() -> int:
    return int(lib().tko_num_threads())


def set_num_threads(n: int):
    """Override OMP_NUM_THREADS (torchrun exports 1 to its workers)."""
    lib().tko_set_num_threads(int(n))


def dist2(q, p) -> np.float32:
    q, p = _f32(q), _f32(p)
    return np.float32(lib().tko_dist2(_p(q, C.c_float), _p(p, C.c_float)))


def knn_brute(xyz, k):
    """Ground truth all-points kNN: returns (idx [n,k] int32, dist [n,k] float32)."""
    xyz = _f32(xyz).reshape(-1, 3)
    n = xyz.shape[0]
    idx = np.empty((n, k), np.int32)
    dist = np.empty((n, k), np.float32)
    rc = lib().tko_knn_brute(_p(xyz, C.c_float), n, k, _p(idx, C.c_int32), _p(dist, C.c_float))
    assert rc == 0
    return idx, dist


def set_squared_output(on: bool):
    """Report d2 instead of sqrtf(d2) (used by the stand-in engine of the multi-GPU host tests)."""
    lib().tko_set_squared_output(1 if on else 0)


def knn_brute_queries(xyz, queries, k, self_ids=None, radius2=np.inf, caps2=None):
    xyz = _f32(xyz).reshape(-1, 3)
    q = _f32(queries).reshape(-1, 3)
    nq = q.shape[0]
    sid = None
    if self_ids is not None:
        sid = np.ascontiguousarray(self_ids, dtype=np.int32)
    else:
        sid = np.full(nq, -1, np.int32)
    idx = np.empty((nq, k), np.int32)
    dist = np.empty((nq, k), np.float32)
    caps = _f32(caps2) if caps2 is not None else None
    rc = lib().tko_knn_brute_queries(_p(xyz, C.c_float), xyz.shape[0], _p(q, C.c_float), nq, _p(sid, C.c_int32), k,
                                     C.c_float(radius2), _p(caps, C.c_float), _p(idx, C.c_int32), _p(dist, C.c_float))
    assert rc == 0
    return idx, dist


def knn_kdtree(xyz, k, return_times=False):
    """Exact kd-tree all-points kNN (OpenMP on all cores); bit-identical to knn_brute."""
    xyz = _f32(xyz).reshape(-1, 3)
    n = xyz.shape[0]
    idx = np.empty((n, k), np.int32)
    dist = np.empty((n, k), np.float32)
    b, q = C.c_double(0), C.c_double(0)
    rc = lib().tko_knn_kdtree(_p(xyz, C.c_float), n, k, _p(idx, C.c_int32), _p(dist, C.c_float), C.byref(b), C.byref(q))
    assert rc == 0
    if return_times:
        return idx, dist, b.value, q.value
    return idx, dist


class KdTree:
    """Reusable exact kd-tree over a point set (for sampled-query checks and the CPU baseline)."""

    def __init__(self, xyz, leaf=8):
        self.xyz = _f32(xyz).reshape(-1, 3)
        self._h = lib().tko_kdtree_build(_p(self.xyz, C.c_float), self.xyz.shape[0], leaf)
        assert self._h

    def query(self, queries, k, self_ids=None, radius2=np.inf):
        q = _f32(queries).reshape(-1, 3)
        nq = q.shape[0]
        sid = np.ascontiguousarray(self_ids, dtype=np.int32) if self_ids is not None else np.full(nq, -1, np.int32)
        idx = np.empty((nq, k), np.int32)
        dist = np.empty((nq, k), np.float32)
        rc = lib().tko_kdtree_knn(self._h, _p(q, C.c_float), nq, _p(sid, C.c_int32), k, C.c_float(radius2),
                                  _p(idx, C.c_int32), _p(dist, C.c_float))
        assert rc == 0
        return idx, dist

    def query_positions(self, pos, k, radius2=np.inf):
        """Queries = the points at tree positions `pos` (runs of consecutive positions are spatially coherent).
        Returns (ids [m] original indices of the queries, idx [m,k], dist [m,k])."""
        pos = np.ascontiguousarray(pos, dtype=np.int64)
        m = pos.shape[0]
        ids = np.empty(m, np.int32)
        idx = np.empty((m, k), np.int32)
        dist = np.empty((m, k), np.float32)
        pp = pos.ctypes.data_as(C.POINTER(C.c_int64))
        assert lib().tko_kdtree_ids(self._h, pp, m, _p(ids, C.c_int32)) == 0
        assert lib().tko_kdtree_knn_positions(self._h, pp, m, k, C.c_float(radius2), _p(idx, C.c_int32), _p(dist, C.c_float)) == 0
        return ids, idx, dist

    def close(self):
        if self._h:
            lib().tko_kdtree_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def range_count(xyz, radius):
    xyz = _f32(xyz).reshape(-1, 3)
    out = np.empty(xyz.shape[0], np.uint32)
    rc = lib().tko_range_count(_p(xyz, C.c_float), xyz.shape[0], C.c_float(radius), _p(out, C.c_uint32))
    assert rc == 0
    return out


def reference_trueknn(xyz, k, radius, max_rounds=64):
    """The reference's own (inexact) round algorithm; returns (idx, dist, rounds, final_radius, rc)."""
    xyz = _f32(xyz).reshape(-1, 3)
    n = xyz.shape[0]
    idx = np.empty((n, k), np.int32)
    dist = np.empty((n, k), np.float32)
    rounds, fr = C.c_int(0), C.c_float(0)
    rc = lib().tko_reference_trueknn(_p(xyz, C.c_float), n, k, C.c_float(radius), max_rounds, _p(idx, C.c_int32),
                                     _p(dist, C.c_float), C.byref(rounds), C.byref(fr))
    return idx, dist, rounds.value, fr.value, rc


def parse_points(text: bytes, n: int, dim: int):
    """Reference point-file grammar (hostCode.cpp:83-124). Returns [m,3] float32 or raises ValueError."""
    cap = max(1, text.count(b"\n") + 2) * 64
    out = np.empty((cap, 3), np.float32)
    m = lib().tko_parse_points(text, len(text), n, dim, _p(out, C.c_float), cap)
    if m < 0:
        raise ValueError(f"tko_parse_points failed: {m}")
    return out[:m].copy()
