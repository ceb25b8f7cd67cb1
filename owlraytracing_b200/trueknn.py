"""Host-side mirror of the reference's TrueKNN sample interface over the libtrueknn C ABI.

The reference has no library API for kNN — its interface is the command line of
samples/s01-trueknn (hostCode.cpp:66-73: file, n, dim, start radius, k, out file) and the in-memory
`Neigh` buffer (GeomTypes.h:22-28, read at hostCode.cpp:294).  `run_sample()` below keeps those six
positional arguments; `TrueKNN` is the object behind it.  Every array argument may be a numpy array
(host memory, staged inside the C call) or a CUDA `torch.Tensor` (device memory, zero-copy).
"""
from __future__ import annotations

import ctypes as C
import re
import time

import numpy as np

from . import _lib
from ._lib import Stats


class TrueKNNError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"{_lib.ERROR_NAMES.get(code, code)}: {message}")
        self.code = code


def _is_tensor(a) -> bool:
    return hasattr(a, "data_ptr") and hasattr(a, "is_cuda")


def _ptr(a):
    """Raw address of a numpy array or torch tensor (None -> NULL)."""
    if a is None:
        return None
    if _is_tensor(a):
        if not a.is_contiguous():
            raise ValueError("tensor must be contiguous")
        return C.c_void_p(a.data_ptr())
    if not a.flags["C_CONTIGUOUS"]:
        raise ValueError("array must be C-contiguous")
    return C.c_void_p(a.ctypes.data)


def _check_dtype(a, np_dtype, name):
    if _is_tensor(a):
        import torch

        want = {np.float32: torch.float32, np.int32: torch.int32, np.uint32: torch.int32, np.uint64: torch.int64}[np_dtype]
        if a.dtype != want and not (np_dtype is np.uint32 and a.dtype == torch.uint32):
            raise TypeError(f"{name}: expected {want}, got {a.dtype}")
    elif a.dtype != np_dtype:
        raise TypeError(f"{name}: expected {np.dtype(np_dtype)}, got {a.dtype}")


class TrueKNN:
    """One LBVH + traversal context on one CUDA device (`tknn_ctx`)."""

    def __init__(self, device: int = 0, **options):
        self._L = _lib.load()
        h = C.c_void_p()
        rc = self._L.tknn_create(int(device), C.byref(h))
        if rc != _lib.OK:
            raise TrueKNNError(rc, f"tknn_create(device={device}) failed — a CUDA sm_100 device is required, "
                                   "there is no CPU fallback")
        self._h = h
        self.device = int(device)
        self._own_stream = True  # until set_stream() hands over a caller stream
        self.n = 0
        self._like = None
        for k, v in options.items():
            self.set_option(k, v)

    # ---- plumbing ----
    def _check(self, rc: int):
        if rc != _lib.OK:
            raise TrueKNNError(rc, self._L.tknn_last_error(self._h).decode())

    def close(self):
        if getattr(self, "_h", None):
            self._L.tknn_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    _OPTS = {
        "leaf_size": _lib.OPT_LEAF_SIZE, "counters": _lib.OPT_COUNTERS, "leaf_policy": _lib.OPT_LEAF_POLICY,
        "sample_groups": _lib.OPT_SAMPLE_GROUPS, "blocks_per_sm": _lib.OPT_BLOCKS_PER_SM,
        "squared_dist": _lib.OPT_SQUARED_DIST, "radius_quantile": _lib.OPT_RADIUS_QUANTILE,
        "keep_scratch": _lib.OPT_KEEP_SCRATCH, "sparse_divisor": _lib.OPT_SPARSE_DIVISOR,
        "approx_filter": _lib.OPT_APPROX_FILTER, "output_chunks": _lib.OPT_OUTPUT_CHUNKS,
        "file_order_chunks": _lib.OPT_FILE_ORDER_CHUNKS, "morton_bits": _lib.OPT_MORTON_BITS,
        "tie_pruning": _lib.OPT_TIE_PRUNING, "warp_round_max": _lib.OPT_WARP_ROUND_MAX, "curve": _lib.OPT_CURVE,
        "speculative_max": _lib.OPT_SPECULATIVE_MAX, "sparse_team": _lib.OPT_SPARSE_TEAM, "sort_mode": _lib.OPT_SORT_MODE,
    }

    def set_option(self, name: str, value: int):
        if name not in self._OPTS:
            raise KeyError(name)
        self._check(self._L.tknn_set_option(self._h, self._OPTS[name], int(value)))

    def set_stream(self, cuda_stream: int | None):
        """Run on a caller-owned cudaStream_t (e.g. torch.cuda.current_stream().cuda_stream).

        None = the context's own non-blocking stream.  0 is the LEGACY DEFAULT stream (what torch's default
        current stream is): it is passed as cudaStreamLegacy (0x1) so that the library's work is ordered with the
        caller's kernels and collectives on that stream instead of racing them from a private stream."""
        self._own_stream = cuda_stream is None
        if cuda_stream is None:
            handle = 0
        elif cuda_stream == 0:
            handle = 1  # cudaStreamLegacy
        else:
            handle = cuda_stream
        self._check(self._L.tknn_set_stream(self._h, C.c_void_p(handle)))

    def _in(self, a):
        """Pointer of an INPUT array.  While the context runs on its own stream, a CUDA tensor produced by
        torch on another stream must be complete before the library reads it: settle torch's current stream."""
        if a is not None and _is_tensor(a) and a.is_cuda and self._own_stream:
            import torch

            torch.cuda.current_stream(a.device).synchronize()
        return _ptr(a)

    def stats(self) -> dict:
        s = Stats()
        self._check(self._L.tknn_get_stats(self._h, C.byref(s)))
        return s.as_dict()

    def _out(self, like, rows, k, dtype):
        if _is_tensor(like):
            import torch

            out = torch.empty((rows, k), dtype={np.int32: torch.int32, np.float32: torch.float32}[dtype], device=like.device)
            self._settle_outputs(out)
            return out
        return np.empty((rows, k), dtype)

    def _settle_outputs(self, t):
        """The caching allocator may hand back a block that a kernel still pending on torch's current stream uses.
        While the context runs on its OWN stream nothing orders that kernel before the engine's writes: settle torch's
        stream first (a no-op cost when the context was given torch's stream with set_stream)."""
        if self._own_stream and _is_tensor(t) and t.is_cuda:
            import torch

            torch.cuda.current_stream(t.device).synchronize()

    # ---- the path ----
    def build(self, points, dim: int | None = None):
        """Build the LBVH over `points` ([n, 2|3] float32; extra columns are skipped through the stride)."""
        if not _is_tensor(points):
            points = np.ascontiguousarray(points, dtype=np.float32)
        _check_dtype(points, np.float32, "points")
        if points.ndim != 2:
            raise ValueError("points must be [n, >=dim]")
        stride = int(points.shape[1])
        if dim is None:
            dim = min(stride, 3)
        self._check(self._L.tknn_build(self._h, self._in(points), int(points.shape[0]), int(dim), stride))
        self.n = int(points.shape[0])
        self._like = points
        return self

    def search(self, k: int, start_radius: float = 0.0, out=None, indices_only: bool = False):
        """All-points kNN. Returns (idx [n,k] int32, dist [n,k] float32), rows in build order.
        indices_only (or out = (idx, None)): dist_out = NULL — no distances are returned or copied."""
        if out is None:
            idx = self._out(self._like, self.n, k, np.int32)
            dist = None if indices_only else self._out(self._like, self.n, k, np.float32)
        else:
            idx, dist = out
        _check_dtype(idx, np.int32, "idx_out")
        if dist is not None:
            _check_dtype(dist, np.float32, "dist_out")
        self._check(self._L.tknn_search(self._h, int(k), C.c_float(start_radius), _ptr(idx), _ptr(dist) if dist is not None else None))
        return idx, dist

    def shard_capacity(self, n_shards: int) -> int:
        return int(self._L.tknn_shard_capacity(self.n, int(n_shards)))

    def search_shard(self, k: int, shard: int, n_shards: int, start_radius: float = 0.0, out=None):
        """Query-sharded search: returns (qid [m], idx [m,k], dist [m,k]) for this shard's Morton slice."""
        cap = self.shard_capacity(n_shards)
        if out is None:
            if _is_tensor(self._like):
                import torch

                qid = torch.empty((cap,), dtype=torch.int32, device=self._like.device)
            else:
                qid = np.empty((cap,), np.int32)
            idx = self._out(self._like, cap, k, np.int32)
            dist = self._out(self._like, cap, k, np.float32)
        else:
            qid, idx, dist = out
        m = C.c_uint64(0)
        self._check(self._L.tknn_search_shard(self._h, int(k), C.c_float(start_radius), int(shard), int(n_shards), _ptr(qid),
                                              _ptr(idx), _ptr(dist), C.byref(m)))
        m = int(m.value)
        return qid[:m], idx[:m], dist[:m]

    def query(self, queries, k: int, self_ids=None, init_radius2=None, start_radius: float = 0.0, dim: int | None = None):
        """kNN of a separate query set against the built BVH; rows follow the query order.
        init_radius2: optional per-query cap on the squared distance (closed; negative = none)."""
        if not _is_tensor(queries):
            queries = np.ascontiguousarray(queries, dtype=np.float32)
        _check_dtype(queries, np.float32, "queries")
        nq, stride = int(queries.shape[0]), int(queries.shape[1])
        if dim is None:
            dim = min(stride, 3)
        if self_ids is not None and not _is_tensor(self_ids):
            self_ids = np.ascontiguousarray(self_ids, dtype=np.int32)
        if init_radius2 is not None and not _is_tensor(init_radius2):
            init_radius2 = np.ascontiguousarray(init_radius2, dtype=np.float32)
        idx = self._out(queries, nq, k, np.int32)
        dist = self._out(queries, nq, k, np.float32)
        self._check(self._L.tknn_query(self._h, self._in(queries), nq, int(dim), stride, self._in(self_ids), self._in(init_radius2), int(k),
                                       C.c_float(start_radius), _ptr(idx), _ptr(dist)))
        return idx, dist

    def range_count(self, radius: float):
        """Neighbours within the closed ball of `radius` for every point (DBSCAN core-point test)."""
        if _is_tensor(self._like):
            import torch

            out = torch.empty((self.n,), dtype=torch.int32, device=self._like.device)
        else:
            out = np.empty((self.n,), np.uint32)
        self._check(self._L.tknn_range_count(self._h, C.c_float(radius), _ptr(out)))
        return out

    def estimate_start_radius(self, k: int) -> float:
        r = C.c_float(0)
        self._check(self._L.tknn_estimate_start_radius(self._h, int(k), C.byref(r)))
        return float(r.value)

    def brute_force(self, query_ids, k: int):
        """Exact tiled brute force on the GPU for selected data indices (second oracle)."""
        if not _is_tensor(query_ids):
            query_ids = np.ascontiguousarray(query_ids, dtype=np.int32)
        nq = int(query_ids.shape[0])
        idx = self._out(query_ids, nq, k, np.int32)
        dist = self._out(query_ids, nq, k, np.float32)
        self._check(self._L.tknn_brute_force(self._h, self._in(query_ids), nq, int(k), _ptr(idx), _ptr(dist)))
        return idx, dist

    def merge_topk(self, idx_parts, d2_parts):
        """Merge [parts, nq, k] (idx, d2) device lists on (d2, index); returns (idx, dist=sqrt(d2))."""
        parts, nq, k = (int(x) for x in idx_parts.shape)
        idx = self._out(idx_parts, nq, k, np.int32)
        dist = self._out(idx_parts, nq, k, np.float32)
        self._check(self._L.tknn_merge_topk(self._h, self._in(idx_parts), self._in(d2_parts), parts, nq, k, _ptr(idx), _ptr(dist)))
        return idx, dist

    # ---- multi-GPU, one rank per process (include/trueknn.h "multi-GPU" (b)) ----
    @staticmethod
    def unique_id() -> bytes:
        """128 bytes from ncclGetUniqueId: rank 0 creates them, every rank passes them to comm_init."""
        L = _lib.load()
        buf = C.create_string_buffer(_lib.UNIQUE_ID_BYTES)
        rc = L.tknn_comm_unique_id(buf)
        if rc != _lib.OK:
            raise TrueKNNError(rc, "tknn_comm_unique_id failed (libnccl.so.2 not loadable?)")
        return buf.raw

    def comm_init(self, n_ranks: int, rank: int, unique_id: bytes):
        if len(unique_id) != _lib.UNIQUE_ID_BYTES:
            raise ValueError("unique_id must be 128 bytes")
        self._check(self._L.tknn_comm_init(self._h, int(n_ranks), int(rank), C.c_char_p(unique_id)))
        self.rank, self.n_ranks = int(rank), int(n_ranks)
        return self

    def comm_init_torch(self, group=None):
        """Communicator over the ranks of a torch.distributed group: the unique id travels through the group."""
        import torch.distributed as dist

        rank, world = dist.get_rank(group), dist.get_world_size(group)
        box = [self.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        return self.comm_init(world, rank, box[0])

    def dist_stats(self) -> dict:
        s = _lib.DistStats()
        self._check(self._L.tknn_get_dist_stats(self._h, C.byref(s)))
        return s.as_dict()

    def build_replicated(self, local_points, first: int, n_total: int, dim: int | None = None):
        """TKNN_SHARD_QUERIES build: this rank's contiguous slice in, the replicated LBVH out (one ncclAllGather)."""
        if not _is_tensor(local_points):
            local_points = np.ascontiguousarray(local_points, dtype=np.float32)
        _check_dtype(local_points, np.float32, "local_points")
        stride = int(local_points.shape[1])
        if dim is None:
            dim = min(stride, 3)
        self._check(self._L.tknn_build_replicated(self._h, self._in(local_points), int(local_points.shape[0]), int(first),
                                                  int(n_total), int(dim), stride))
        self.n = int(n_total)
        self._like = local_points
        return self

    def partition_build(self, local_points, first_index: int, dim: int | None = None):
        """TKNN_PARTITION_POINTS build: rows of global indices first_index.. in, one Morton range of the cloud owned."""
        if not _is_tensor(local_points):
            local_points = np.ascontiguousarray(local_points, dtype=np.float32)
        _check_dtype(local_points, np.float32, "local_points")
        stride = int(local_points.shape[1])
        if dim is None:
            dim = min(stride, 3)
        self._check(self._L.tknn_partition_build(self._h, self._in(local_points), int(local_points.shape[0]), int(first_index),
                                                 int(dim), stride))
        self.n = int(self._L.tknn_partition_owned(self._h))
        self._like = local_points
        return self

    def partition_owned(self) -> int:
        return int(self._L.tknn_partition_owned(self._h))

    def partition_search(self, k: int, start_radius: float = 0.0, out=None):
        """Exact kNN of the owned points against the global cloud: (gid [m], idx [m,k] global ids, dist [m,k])."""
        m = self.partition_owned()
        if out is None:
            if _is_tensor(self._like):
                import torch

                gid = torch.empty((m,), dtype=torch.int32, device=self._like.device)
            else:
                gid = np.empty((m,), np.int32)
            idx = self._out(self._like, m, k, np.int32)
            dist = self._out(self._like, m, k, np.float32)
        else:
            gid, idx, dist = out
        got = C.c_uint64(0)
        self._check(self._L.tknn_partition_search(self._h, int(k), C.c_float(start_radius), _ptr(gid), _ptr(idx), _ptr(dist),
                                                  int(gid.shape[0]), C.byref(got)))
        g = int(got.value)
        return gid[:g], idx[:g], dist[:g]

    def partition_verify(self, k: int, samples: int, gid, idx, dist):
        """Distributed brute-force check of `samples` sampled rows per rank; returns GLOBAL (checked, bad)."""
        a, b = C.c_uint64(0), C.c_uint64(0)
        self._check(self._L.tknn_partition_verify(self._h, int(k), int(samples), self._in(gid), self._in(idx), self._in(dist),
                                                  int(gid.shape[0]), C.byref(a), C.byref(b)))
        return int(a.value), int(b.value)

    # ---- introspection (tests / bench) ----
    def sort_pairs(self, keys, values=None):
        n = int(keys.shape[0])
        self._check(self._L.tknn_sort_pairs(self._h, self._in(keys), self._in(values) if values is not None else None, n))
        return keys, values

    def get_bvh(self):
        """Host copies: nodes [n_nodes,16] float32 (bit views for refs), points [n,4] float32, leaf_start [n_leaves+1]."""
        st = self.stats()
        nodes = np.empty((st["n_nodes"], 16), np.float32)
        pts = np.empty((self.n, 4), np.float32)
        leaf_start = np.empty((st["n_leaves"] + 1,), np.uint32)
        self._check(self._L.tknn_get_bvh(self._h, _ptr(nodes), _ptr(pts), _ptr(leaf_start)))
        return nodes, pts, leaf_start

    def morton_codes(self, points, box6, dim: int | None = None):
        """63-bit Morton codes on the cubic grid over `box6` = (lo.xyz, hi.xyz); int64/uint64 [n]."""
        if not _is_tensor(points):
            points = np.ascontiguousarray(points, dtype=np.float32)
        n, stride = int(points.shape[0]), int(points.shape[1])
        if dim is None:
            dim = min(stride, 3)
        if _is_tensor(points):
            import torch

            out = torch.empty((n,), dtype=torch.int64, device=points.device)
            box = torch.as_tensor(box6, dtype=torch.float32).cpu().contiguous().numpy()
        else:
            out = np.empty((n,), np.uint64)
            box = np.ascontiguousarray(box6, dtype=np.float32)
        self._check(self._L.tknn_morton_codes(self._h, self._in(points), n, int(dim), stride, _ptr(box), _ptr(out)))
        return out

    def reach_mask(self, points, reach2, box6, summaries, self_rank: int, cell_bits: int = 3):
        """Bit mask of the remote ranks whose per-cell summary boxes the ball (p, sqrt(reach2)) touches (CUDA tensors)."""
        import torch

        n, stride = int(points.shape[0]), int(points.shape[1])
        n_ranks = int(summaries.shape[0])
        box = torch.as_tensor(box6, dtype=torch.float32, device=points.device).contiguous()
        summ = summaries.contiguous()
        out = torch.empty((n,), dtype=torch.int32, device=points.device)
        self._check(self._L.tknn_reach_mask(self._h, self._in(points), n, stride, self._in(reach2.contiguous()), self._in(box),
                                            self._in(summ), n_ranks, int(cell_bits), int(self_rank), _ptr(out)))
        return out

    def generate_uniform(self, seed: int, first: int, n: int, out=None):
        if out is None:
            out = np.empty((n, 3), np.float32)
        self._check(self._L.tknn_generate_uniform(self._h, int(seed), int(first), int(n), _ptr(out)))
        return out

    def measure_bandwidth(self) -> dict:
        l2, hbm, sms, l2b = C.c_double(0), C.c_double(0), C.c_int(0), C.c_uint64(0)
        self._check(self._L.tknn_measure_bandwidth(self._h, C.byref(l2), C.byref(hbm), C.byref(sms), C.byref(l2b)))
        return {"l2_gbs": l2.value, "hbm_read_gbs": hbm.value, "sm_count": sms.value, "l2_bytes": int(l2b.value)}

    def measure_smem_bandwidth(self) -> dict:
        a, b = C.c_double(0), C.c_double(0)
        self._check(self._L.tknn_measure_smem_bandwidth(self._h, C.byref(a), C.byref(b)))
        return {"smem_conflict_free_gbs": a.value, "smem_broadcast_gbs": b.value}


class _RankView(TrueKNN):
    """A rank's context inside a MultiTrueKNN (owned by it): statistics and options only."""

    def __init__(self, L, handle):  # noqa: D401 — no tknn_create here
        self._L, self._h, self._own_stream, self.n, self._like = L, handle, True, 0, None

    def close(self):
        self._h = None


class MultiTrueKNN:
    """All devices from ONE process (`tknn_create_multi`): the device-list form of the reference's context
    (owlContextCreate(ids, n), owl/include/owl/owl_host.h:360).  Host arrays in, host arrays in file order out.

    mode: "shard" (BVH replicated, queries sharded) or "partition" (points partitioned by Morton range).
    Naming one device several times runs that many ranks on it (how the single-GPU tests exercise the paths)."""

    MODES = {"shard": _lib.SHARD_QUERIES, "partition": _lib.PARTITION_POINTS}

    def __init__(self, device_ids, mode: str = "shard", **options):
        self._L = _lib.load()
        ids = (C.c_int * len(device_ids))(*[int(d) for d in device_ids])
        h = C.c_void_p()
        rc = self._L.tknn_create_multi(ids, len(device_ids), self.MODES[mode], C.byref(h))
        if rc != _lib.OK:
            raise TrueKNNError(rc, f"tknn_create_multi({list(device_ids)}, {mode}) failed")
        self._h, self.mode, self.n = h, mode, 0
        for name, v in options.items():
            self.set_option(name, v)

    def _check(self, rc):
        if rc != _lib.OK:
            raise TrueKNNError(rc, self._L.tknn_multi_last_error(self._h).decode())

    def set_option(self, name: str, value: int):
        self._check(self._L.tknn_multi_set_option(self._h, TrueKNN._OPTS[name], int(value)))

    @property
    def n_ranks(self) -> int:
        return int(self._L.tknn_multi_ranks(self._h))

    def rank(self, r: int) -> TrueKNN:
        return _RankView(self._L, C.c_void_p(self._L.tknn_multi_ctx(self._h, int(r))))

    def build(self, points, dim: int | None = None):
        points = np.ascontiguousarray(points, dtype=np.float32)
        stride = int(points.shape[1])
        if dim is None:
            dim = min(stride, 3)
        self._check(self._L.tknn_multi_build(self._h, _ptr(points), int(points.shape[0]), int(dim), stride))
        self.n = int(points.shape[0])
        return self

    def search(self, k: int, start_radius: float = 0.0, out=None):
        idx, dist = out if out is not None else (np.empty((self.n, k), np.int32), np.empty((self.n, k), np.float32))
        self._check(self._L.tknn_multi_search(self._h, int(k), C.c_float(start_radius), _ptr(idx), _ptr(dist)))
        return idx, dist

    def times(self) -> dict:
        t = (C.c_float * 4)()
        self._check(self._L.tknn_multi_get_times(self._h, t))
        return {"build_ms": t[0], "search_ms": t[1], "exchange_ms": t[2], "d2h_ms": t[3]}

    def close(self):
        if getattr(self, "_h", None):
            self._L.tknn_multi_destroy(self._h)
            self._h = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ------------------------------------------------------------------------------------------------
# The sample's command line (hostCode.cpp:66-73) as a function.
# ------------------------------------------------------------------------------------------------
_FLOAT_RE = re.compile(r"[+-]?(?:\d+\.?\d*|\.\d+)(?:[eE][+-]?\d+)?")


def read_points(path: str, n: int, dim: int) -> np.ndarray:
    """Point-file grammar of the sample (hostCode.cpp:83-124): floats separated by ',' and/or blanks,
    lines read while n*dim floats are still owed (the last line read is consumed whole), then chunked
    by `dim`; dim 2 gives z = 0.  Fast path: a `.f32` file is raw little-endian float32 rows of `dim`."""
    if dim not in (2, 3):
        raise ValueError("dimension must be 2 or 3 (hostCode.cpp:114-124 creates no points otherwise)")
    if path.endswith(".f32"):
        flat = np.fromfile(path, dtype="<f4", count=n * dim)
    else:
        vals: list[float] = []
        owed = n * dim
        with open(path, "rb") as f:
            for line in f:
                if owed <= 0:
                    break
                # operator>> stops at the first token that is not a float; a ',' after a float is skipped
                pos, text = 0, line.decode("latin-1")
                ln = len(text)
                while True:
                    while pos < ln and text[pos].isspace():
                        pos += 1
                    if pos >= ln:
                        break
                    m = _FLOAT_RE.match(text, pos)
                    if not m:
                        break
                    vals.append(float(m.group(0)))
                    owed -= 1
                    pos = m.end()
                    if pos < ln and text[pos] == ",":
                        pos += 1
        flat = np.asarray(vals, dtype=np.float32)
    if flat.size % dim:
        raise ValueError(f"{flat.size} floats is not a multiple of dim={dim} (the reference throws std::out_of_range)")
    pts = flat.reshape(-1, dim)
    if dim == 2:
        pts = np.concatenate([pts, np.zeros((pts.shape[0], 1), np.float32)], axis=1)
    return np.ascontiguousarray(pts, dtype=np.float32)


def read_points_fast(path: str, n: int, dim: int) -> np.ndarray:
    """Same grammar as read_points, parsed in parallel by libtrueknn (tknn_read_points)."""
    L = _lib.load()
    cap = int(n) + 8
    m = C.c_uint64(0)
    for _ in range(2):
        out = np.empty((cap, 3), np.float32)
        rc = L.tknn_read_points(path.encode(), int(n), int(dim), C.c_void_p(out.ctypes.data), cap, C.byref(m))
        if rc == _lib.OK:
            return np.ascontiguousarray(out[: int(m.value)])
        # The last line read is consumed whole (hostCode.cpp:92), so a file with many floats per line yields more
        # than n rows: *n_out then names the row count needed — retry once with that capacity.
        if rc == _lib.EINVAL and int(m.value) > cap:
            cap = int(m.value)
            continue
        break
    raise ValueError(f"tknn_read_points({path!r}) failed with {_lib.ERROR_NAMES.get(rc, rc)}")


def write_neighbours(path: str, idx: np.ndarray, dist: np.ndarray, binary: bool = False):
    """`query,neighbourIndex,distance` lines (hostCode.cpp:316, commented out in the reference) or raw arrays."""
    L = _lib.load()
    idx = np.ascontiguousarray(idx, dtype=np.int32)
    dist = np.ascontiguousarray(dist, dtype=np.float32)
    rc = L.tknn_write_neighbours(path.encode(), C.c_void_p(idx.ctypes.data), C.c_void_p(dist.ctypes.data), idx.shape[0],
                                 idx.shape[1], 1 if binary else 0)
    if rc != _lib.OK:
        raise OSError(f"tknn_write_neighbours({path!r}) failed")


def run_sample(argv: list[str], device: int = 0, neighbours_path: str | None = None) -> dict:
    """`sample01-trueknn file n dim start_radius k outfile` (samples/s01-trueknn/README.md:7-14).

    Appends the total seconds (build + kNN) to `outfile` exactly like hostCode.cpp:346-356, and —
    what the reference leaves commented out at hostCode.cpp:312-319 — optionally writes the
    neighbours as `query,neighbourIndex,distance` lines to `neighbours_path`."""
    if len(argv) != 6:
        raise SystemExit("usage: trueknn <file> <n> <dim> <start radius> <k> <output file>")
    path, n, dim, r0, k, outfile = argv[0], int(argv[1]), int(argv[2]), float(argv[3]), int(argv[4]), argv[5]
    pts = read_points_fast(path, n, dim)
    with TrueKNN(device) as t:
        t0 = time.perf_counter()
        t.build(pts, dim=3)
        t1 = time.perf_counter()
        idx, dist = t.search(k, r0)
        t2 = time.perf_counter()
        st = t.stats()
    total = (t1 - t0) + (t2 - t1)
    with open(outfile, "a") as f:
        f.write(f"{total}\n")
    if neighbours_path:
        write_neighbours(neighbours_path, idx, dist)
    return {"build_s": t1 - t0, "knn_s": t2 - t1, "total_s": total, "rounds": st["rounds"], "idx": idx, "dist": dist,
            "stats": st}
