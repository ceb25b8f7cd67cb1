"""CPU stand-in for the TrueKNN context, used ONLY by the world_size>1 gloo tests of the multi-GPU
drivers' host logic (collective plumbing, routing, merging).  It answers through the oracle; the
product drivers construct the CUDA engine by default and have no fallback to this."""
import numpy as np
import torch

from oracle import oracle as O


def _spread21(v):
    x = v.astype(np.uint64) & np.uint64(0x1FFFFF)
    for shift, mask in ((32, 0x001F00000000FFFF), (16, 0x001F0000FF0000FF), (8, 0x100F00F00F00F00F), (4, 0x10C30C30C30C30C3),
                        (2, 0x1249249249249249)):
        x = (x | (x << np.uint64(shift))) & np.uint64(mask)
    return x


class CpuEngine:
    def __init__(self):
        self.squared = 0
        self.x = None

    def set_option(self, name, value):
        if name == "squared_dist":
            self.squared = int(value)

    def _emit(self, idx, dist):
        return torch.from_numpy(idx), torch.from_numpy(dist)

    def morton_codes(self, points, box6):
        p = np.asarray(points, dtype=np.float32)
        box = np.asarray(box6, dtype=np.float32)
        ext = max(float((box[3:] - box[:3]).max()), np.finfo(np.float32).tiny)
        scale = np.float32(2097152.0) / np.float32(ext)
        c = np.clip((p - box[:3]) * scale, 0, 2097151).astype(np.uint32)
        code = (_spread21(c[:, 0]) << np.uint64(2)) | (_spread21(c[:, 1]) << np.uint64(1)) | _spread21(c[:, 2])
        return torch.from_numpy(code.astype(np.int64))

    def build(self, points):
        self.x = np.ascontiguousarray(np.asarray(points, dtype=np.float32))
        self.n = self.x.shape[0]
        return self

    n_ranks = 0  # "communicator" present: the stand-in talks through torch.distributed (gloo) itself

    def build_replicated(self, local_points, first, n_total):
        """Stand-in of tknn_build_replicated: padded all_gather of the slices over the default (gloo) group."""
        import torch.distributed as dist

        world = dist.get_world_size()
        loc = torch.as_tensor(np.asarray(local_points, dtype=np.float32))
        per = (n_total + world - 1) // world
        pad = torch.zeros((per, loc.shape[1]), dtype=loc.dtype)
        pad[: loc.shape[0]] = loc
        full = torch.empty((per * world, loc.shape[1]), dtype=loc.dtype)
        dist.all_gather_into_tensor(full, pad)
        return self.build(full[:n_total].numpy())

    def search(self, k, start_radius=0.0):
        O.set_squared_output(bool(self.squared))
        try:
            return self._emit(*O.knn_brute(self.x, k))
        finally:
            O.set_squared_output(False)

    def query(self, queries, k, self_ids=None, init_radius2=None, start_radius=0.0):
        O.set_squared_output(bool(self.squared))
        try:
            caps = None if init_radius2 is None else np.asarray(init_radius2, dtype=np.float32)
            return self._emit(*O.knn_brute_queries(self.x, np.asarray(queries, dtype=np.float32), k, self_ids=self_ids, caps2=caps))
        finally:
            O.set_squared_output(False)

    def merge_topk(self, idx_parts, d2_parts):
        ip, dp = np.asarray(idx_parts), np.asarray(d2_parts, dtype=np.float32)
        parts, nq, k = ip.shape
        oi = np.full((nq, k), -1, np.int32)
        od = np.full((nq, k), np.finfo(np.float32).max, np.float32)
        for q in range(nq):
            seen, keys = set(), []
            for s in range(parts):
                for i in range(k):
                    j = int(ip[s, q, i])
                    if j < 0:
                        break
                    if j in seen:
                        continue
                    seen.add(j)
                    keys.append((int(dp[s, q, i].view(np.uint32)), j, dp[s, q, i]))
            keys.sort()
            for i, (_, j, d) in enumerate(keys[:k]):
                oi[q, i] = j
                od[q, i] = d if self.squared else np.sqrt(d)
        return torch.from_numpy(oi), torch.from_numpy(od)

    def reach_mask(self, points, reach2, box6, summaries, self_rank, cell_bits=3):
        """numpy stand-in of tknn_reach_mask: every cell of every remote rank (no neighbour-cell shortcut)."""
        p = np.asarray(points, dtype=np.float64)
        r2 = np.asarray(reach2, dtype=np.float64)
        summ = np.asarray(summaries, dtype=np.float64)
        mask = np.zeros(p.shape[0], np.int32)
        lim = r2 * (1 + 1e-5) + 1e-30
        for s in range(summ.shape[0]):
            if s == self_rank:
                continue
            lo, hi = summ[s, :, :3], summ[s, :, 3:]
            ok = np.isfinite(lo[:, 0]) & (lo[:, 0] <= hi[:, 0])
            lo, hi = lo[ok], hi[ok]
            d = np.maximum(np.maximum(lo[None] - p[:, None], p[:, None] - hi[None]), 0.0)
            hit = ((d * d).sum(-1) <= lim[:, None]).any(1) if lo.shape[0] else np.zeros(p.shape[0], bool)
            mask |= hit.astype(np.int32) << s
        return torch.from_numpy(mask)

    # query-sharded interface: contiguous slices of a Morton-ish order in groups of 32
    def shard_capacity(self, n_shards):
        groups = (self.n + 31) // 32
        return ((groups + n_shards - 1) // n_shards + 1) * 32

    def search_shard(self, k, shard, n_shards, start_radius=0.0):
        box = np.concatenate([self.x.min(0), self.x.max(0)])
        order = np.argsort(self.morton_codes(self.x, box).numpy(), kind="stable")
        groups = (self.n + 31) // 32
        g0, g1 = groups * shard // n_shards, groups * (shard + 1) // n_shards
        ids = order[g0 * 32: min(self.n, g1 * 32)].astype(np.int32)
        idx, dist = O.knn_brute_queries(self.x, self.x[ids], k, self_ids=ids)
        return torch.from_numpy(ids), torch.from_numpy(idx), torch.from_numpy(dist)
