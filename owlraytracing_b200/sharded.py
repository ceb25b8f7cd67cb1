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

    The engine needs: build(points), search_shard(k, shard, n_shards, start_radius) -> (qid, idx, dist),
    shard_capacity(n_shards).  Tests inject a CPU stand-in to exercise the collective plumbing;
    the product path always uses the CUDA engine (there is no fallback)."""

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
        """Each rank holds only its contiguous 1/N slice of the cloud (host or device): upload the slice, assemble
        the full cloud on every GPU with ONE all_gather over NVLink, then build the replicated LBVH.  Moves
        n/N points over PCIe per rank instead of n (the only collective of the query-sharded variant, and it is
        on the build path, not the search path)."""
        if not torch.is_tensor(local_points):
            local_points = torch.from_numpy(local_points)
        if device is None:
            device = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else local_points.device
        loc = local_points.to(device, non_blocking=True).contiguous()
        if self.world == 1:
            full = loc
        else:
            per = (n_total + self.world - 1) // self.world          # slices are ceil(n/N) rows, the last one shorter
            pad = torch.zeros((per, loc.shape[1]), dtype=loc.dtype, device=device)
            pad[: loc.shape[0]] = loc
            full = torch.empty((per * self.world, loc.shape[1]), dtype=loc.dtype, device=device)
            dist.all_gather_into_tensor(full, pad, group=self.group)
            full = full[:n_total].contiguous() if per * self.world != n_total else full
        self._full = full
        if full.is_cuda:
            torch.cuda.current_stream(full.device).synchronize()  # the engine may run on a stream of its own
        return self.build(full)

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
