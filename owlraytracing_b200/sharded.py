"""Query-sharded multi-GPU TrueKNN (BASELINE.json configs[3]).

One process per GPU (`torch.distributed`).  Every rank holds the whole cloud and builds the same
LBVH (the build is deterministic, so the replicas are identical); rank g answers the g-th contiguous
Morton slice of the queries.  There is NO collective on the search path — the reference's own
multi-GPU model is exactly this replicate-and-split (owl/RayGen.cpp:150-200), unused by its sample
(hostCode.cpp:141).  `gather=True` assembles the full [n, k] arrays on every rank with one
all_gather per array (NCCL over NVLink on GPUs; gloo in the CPU tests of the host logic).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


class ShardedTrueKNN:
    """`engine` is a built-or-buildable TrueKNN context (owlraytracing_b200.TrueKNN on this rank's GPU).

    The engine needs: build(points), build_replicated(slice, first, n_total), search_shard(k, shard, n_shards,
    start_radius) -> (qid, idx, dist), shard_capacity(n_shards).  Tests inject a CPU stand-in to exercise the
    bookkeeping; the product path always uses the CUDA engine (there is no fallback), whose collectives run inside
    libtrueknn over NCCL — torch.distributed only carries the NCCL unique id and the optional gather=True."""

    def __init__(self, engine=None, device: int | None = None, group=None):
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        if engine is None:
            from .trueknn import TrueKNN

            if device is None:
                device = torch.cuda.current_device()
            engine = TrueKNN(device)
        self.engine = engine
        self.n = 0

    def build(self, points):
        """points: the FULL cloud on every rank ([n, 2|3] float32, host array or CUDA tensor)."""
        self.engine.build(points)
        self.n = int(points.shape[0])
        return self

    def build_from_slices(self, local_points, n_total: int, device=None):
        """Each rank holds only its contiguous ceil(n/N)-row slice of the cloud (host or device).  The library uploads
        the slice, assembles the cloud on every GPU with ONE ncclAllGather over NVLink and builds the replicated LBVH
        (tknn_build_replicated): n/N points cross PCIe per rank instead of n.  It is the only collective of the
        query-sharded variant, and it is on the build path, not the search path."""
        per = (n_total + self.world - 1) // self.world
        if getattr(self.engine, "n_ranks", None) is None:
            self.engine.comm_init_torch(self.group)
        self.engine.build_replicated(local_points, min(n_total, per * self.rank), n_total)
        self.n = int(n_total)
        return self

    def search(self, k: int, start_radius: float = 0.0, gather: bool = False):
        """Local shard: (qid [m], idx [m, k], dist [m, k]).  gather=True: (idx [n, k], dist [n, k]) on every rank."""
        qid, idx, dst = self.engine.search_shard(k, self.rank, self.world, start_radius=start_radius)
        if not gather:
            return qid, idx, dst
        qid, idx, dst = (torch.as_tensor(a) for a in (qid, idx, dst))
        if self.world == 1:
            out_i = torch.empty((self.n, k), dtype=torch.int32, device=idx.device)
            out_d = torch.empty((self.n, k), dtype=torch.float32, device=idx.device)
            out_i[qid.long()] = idx
            out_d[qid.long()] = dst
            return out_i, out_d
        # shards differ by at most one group of 32: pad to the common capacity, all_gather, scatter by query id
        cap = self.engine.shard_capacity(self.world)
        m = int(qid.shape[0])
        dev = idx.device

        def padded(t, fill, shape, dtype):
            buf = torch.full(shape, fill, dtype=dtype, device=dev)
            buf[:m] = t
            return buf

        pq = padded(qid, -1, (cap,), torch.int32)
        pi = padded(idx, -1, (cap, k), torch.int32)
        pd = padded(dst, 0.0, (cap, k), torch.float32)
        gq = torch.empty((self.world * cap,), dtype=torch.int32, device=dev)
        gi = torch.empty((self.world * cap, k), dtype=torch.int32, device=dev)
        gd = torch.empty((self.world * cap, k), dtype=torch.float32, device=dev)
        dist.all_gather_into_tensor(gq, pq, group=self.group)
        dist.all_gather_into_tensor(gi, pi, group=self.group)
        dist.all_gather_into_tensor(gd, pd, group=self.group)
        keep = gq >= 0
        rows = gq[keep].long()
        out_i = torch.empty((self.n, k), dtype=torch.int32, device=dev)
        out_d = torch.empty((self.n, k), dtype=torch.float32, device=dev)
        out_i[rows] = gi[keep]
        out_d[rows] = gd[keep]
        return out_i, out_d
