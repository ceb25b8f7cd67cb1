"""Executable MODEL of the point-partitioned protocol — TEST INFRASTRUCTURE (tests/test_cpu_multigpu.py).

The product implements this pipeline inside libtrueknn (owlraytracing_b200/csrc/dist.cu, behind
tknn_partition_build / tknn_partition_search, NCCL).  This file restates the same protocol with torch tensors
and torch.distributed so that it runs on CPU over gloo with a stand-in engine answering the local searches
through the oracle (tests/cpu_engine.py): it pins the routing rules — Morton-range ownership, per-cell
summaries, the conservative reach test, the closed k-th-d2 cap on remote searches, the merge on (d2, GLOBAL
index) — against the brute-force oracle at world sizes 2 and 3 without a GPU.

  build   1. global scene box             all_reduce(min/max)
          2. Morton codes on that box     engine.morton_codes
          3. splitters                    pooled code samples, world-1 quantiles (dist.cu: histogram of the top 24
                                          code bits + all-reduce, cell-aligned splitters)
          4. redistribution               all_to_all of (x, y, z, global id) by Morton range
          5. local search structure       over the owned points, ordered by global id (dist.cu: the LBVH stores the
                                          global id of every point, no re-ordering needed)
          6. partition summaries          tight boxes of the owned points per top-level Morton cell (<= 512), all_gather
  search  7. local all-points kNN         squared distances
          8. boundary queries             a query goes to rank s iff its ball (q, d_k) reaches one of s's summary boxes
          9. exchange + remote search     all_to_all; bounded search with the k-th d2 as a closed per-query cap
         10. return + merge               all_to_all back; merge on (d2, GLOBAL index)
"""
from __future__ import annotations

import torch
import torch.distributed as dist

_SUMMARY_BITS = 9                      # summaries live on the 8 x 8 x 8 grid of top-level Morton cells
_SUMMARY_BOXES = 1 << _SUMMARY_BITS    # <= 512 occupied cells per rank
_SAMPLES_PER_RANK = 1024


def _all_to_all_rows(rows: torch.Tensor, send_counts: torch.Tensor, group):
    """rows [sum(send_counts), ...] grouped by destination rank -> (received rows, recv_counts)."""
    recv_counts = torch.empty_like(send_counts)
    dist.all_to_all_single(recv_counts, send_counts, group=group)
    rc = recv_counts.tolist()
    out = torch.empty((int(sum(rc)),) + tuple(rows.shape[1:]), dtype=rows.dtype, device=rows.device)
    dist.all_to_all_single(out, rows.contiguous(), output_split_sizes=rc, input_split_sizes=send_counts.tolist(), group=group)
    return out, recv_counts


def _box_dist2(q: torch.Tensor, lo: torch.Tensor, hi: torch.Tensor) -> torch.Tensor:
    """Squared distance from points q [m,3] to boxes lo/hi (broadcastable to [m,...,3])."""
    d = torch.clamp(torch.maximum(lo - q, q - hi), min=0.0)
    return (d * d).sum(-1)


def _settle(t: torch.Tensor):
    """The engine may run on a stream of its own: torch-produced inputs must be complete before it reads them.
    (Engine calls are synchronous on return, so the other direction needs nothing.)"""
    if t.is_cuda:
        torch.cuda.current_stream(t.device).synchronize()


class PartitionModel:
    def __init__(self, engine=None, device: int | None = None, group=None):
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        if engine is None:
            raise ValueError("the protocol model needs a stand-in engine (tests/cpu_engine.py)")
        self.engine = engine
        self.engine.set_option("squared_dist", 1)  # partial lists travel as d2 (sqrtf is not injective)
        self.stats = {}

    # ------------------------------------------------------------------ build
    def build(self, local_points: torch.Tensor, first_index: int):
        """local_points [n,3] float32 on this rank's device; their global indices are first_index + arange(n)."""
        w, dev = self.world, local_points.device
        pts = local_points.contiguous()
        n = pts.shape[0]
        lo = pts.amin(0) if n else torch.full((3,), float("inf"), device=dev)
        hi = pts.amax(0) if n else torch.full((3,), float("-inf"), device=dev)
        if w > 1:
            dist.all_reduce(lo, op=dist.ReduceOp.MIN, group=self.group)
            dist.all_reduce(hi, op=dist.ReduceOp.MAX, group=self.group)
        self.box = torch.cat([lo, hi]).float()
        gid = torch.arange(first_index, first_index + n, dtype=torch.int64, device=dev)
        if w > 1:
            _settle(pts)
            codes = torch.as_tensor(self.engine.morton_codes(pts, self.box.cpu())).to(dev).long()
            # splitters: evenly spaced samples of the locally sorted codes, pooled, then world-1 quantiles
            sc, _ = torch.sort(codes)
            take = (torch.arange(_SAMPLES_PER_RANK, device=dev, dtype=torch.int64) * max(n - 1, 0)) // (_SAMPLES_PER_RANK - 1)
            sample = sc[take] if n else torch.zeros((_SAMPLES_PER_RANK,), dtype=torch.int64, device=dev)
            pool = torch.empty((w * _SAMPLES_PER_RANK,), dtype=torch.int64, device=dev)
            dist.all_gather_into_tensor(pool, sample.contiguous(), group=self.group)
            pool, _ = torch.sort(pool)
            cut = (torch.arange(1, w, device=dev) * pool.numel()) // w
            self.splitters = pool[cut].contiguous()
            dest = torch.bucketize(codes, self.splitters, right=True)  # Morton range -> owner rank
            order = torch.argsort(dest, stable=True)
            send_counts = torch.bincount(dest, minlength=w).to(torch.int64)
            payload = torch.cat([pts, gid.to(torch.int32).view(-1, 1).view(torch.float32)], 1)[order]
            recv, _ = _all_to_all_rows(payload, send_counts, self.group)
            pts = recv[:, :3].contiguous()
            gid = recv[:, 3].contiguous().view(torch.int32).long()
        # local order = ascending global id, so (d2, local index) ties resolve like (d2, global index)
        gid, o = torch.sort(gid)
        self.pts = pts[o].contiguous()
        self.gid = gid
        self.n_owned = int(self.pts.shape[0])
        if self.n_owned < 2:
            raise ValueError("a rank owns fewer than 2 points; use fewer ranks for this cloud")
        _settle(self.pts)
        self.engine.build(self.pts)
        # Partition summary: the tight box of the owned points inside every top-level Morton cell (9 code bits =
        # the 8 x 8 x 8 grid of the global cube) this rank touches.  Cells are disjoint cubes and a rank's Morton
        # range is a run of whole cells plus at most two partial ones, so the boxes of different ranks only
        # overlap inside those shared end cells — chunks of consecutive points (the first version) straddled
        # Z-curve jumps and made 29 % of the queries "boundary" at 8 ranks.
        if w > 1:
            codes = torch.as_tensor(self.engine.morton_codes(self.pts, self.box.cpu())).to(dev).long()
            cell = codes >> (63 - _SUMMARY_BITS)
            lo_b = torch.full((_SUMMARY_BOXES, 3), float("inf"), device=dev)
            hi_b = torch.full((_SUMMARY_BOXES, 3), float("-inf"), device=dev)
            ix = cell.view(-1, 1).expand(-1, 3)
            lo_b.scatter_reduce_(0, ix, self.pts, "amin")
            hi_b.scatter_reduce_(0, ix, self.pts, "amax")
            mine = torch.cat([lo_b, hi_b], 1).contiguous()            # [512, 6]; untouched cells stay inverted (never hit)
            allb = torch.empty((w * _SUMMARY_BOXES, 6), dtype=torch.float32, device=dev)
            dist.all_gather_into_tensor(allb, mine, group=self.group)
            self.summ = allb.view(w, _SUMMARY_BOXES, 6)
        return self

    # ------------------------------------------------------------------ search
    def search(self, k: int, start_radius: float = 0.0):
        """Exact kNN of the owned points against the GLOBAL cloud.

        Returns (gid [m] int64, idx [m,k] int32 global neighbour indices, dist [m,k] float32)."""
        w, dev, eng = self.world, self.pts.device, self.engine
        import time as _time

        def _tick():
            if dev.type == "cuda":
                torch.cuda.synchronize(dev)
            return _time.perf_counter()

        t_0 = _tick()
        idx, d2 = eng.search(k, start_radius)
        idx, d2 = torch.as_tensor(idx).to(dev), torch.as_tensor(d2).to(dev)
        g32 = self.gid.to(torch.int32)
        gidx = torch.where(idx >= 0, g32[idx.clamp(min=0).long()], idx)
        t_local = _tick()
        self.stats = {"owned": self.n_owned, "boundary_sent": 0, "boundary_received": 0}
        t_detect = t_exchange = t_remote = t_merge = t_local
        if w > 1:
            inf = torch.full((), float("inf"), device=dev)
            dk2 = torch.where(idx[:, k - 1] >= 0, d2[:, k - 1], inf)  # unfilled list: unbounded ball
            reach = dk2  # the reach kernel widens it a little itself (its box test is plain fp32)
            # which remote ranks can each query's ball reach?  (tknn_reach_mask: <= 8 neighbouring cells per query)
            mask = eng.reach_mask(self.pts, reach.contiguous(), self.box, self.summ, self.rank, _SUMMARY_BITS // 3)
            mask = torch.as_tensor(mask).to(dev)
            send_rows, send_counts = [], []
            for s in range(w):
                rows = torch.nonzero((mask >> s) & 1).view(-1) if s != self.rank else torch.empty((0,), dtype=torch.int64, device=dev)
                send_rows.append(rows)
                send_counts.append(int(rows.numel()))
            rows_all = torch.cat(send_rows)
            t_detect = _tick()
            sc = torch.tensor(send_counts, dtype=torch.int64, device=dev)
            self.stats["boundary_sent"] = int(rows_all.numel())
            # query payload: x, y, z, d_k^2
            payload = torch.cat([self.pts[rows_all], dk2[rows_all].view(-1, 1)], 1)
            rq, rc = _all_to_all_rows(payload, sc, self.group)
            self.stats["boundary_received"] = int(rq.shape[0])
            t_exchange = _tick()
            # remote search: closed cap at the asker's k-th d2; local indices -> global
            if rq.shape[0] > 0:
                cap = torch.where(torch.isinf(rq[:, 3]), torch.full_like(rq[:, 3], -1.0), rq[:, 3]).contiguous()
                kk = min(k, self.n_owned)
                rq3 = rq[:, :3].contiguous()
                _settle(rq3)
                ri, rd = eng.query(rq3, kk, init_radius2=cap)
                ri, rd = torch.as_tensor(ri).to(dev), torch.as_tensor(rd).to(dev)
                if kk < k:
                    pad_i = torch.full((ri.shape[0], k - kk), -1, dtype=torch.int32, device=dev)
                    pad_d = torch.full((ri.shape[0], k - kk), 3.4028234663852886e38, dtype=torch.float32, device=dev)
                    ri, rd = torch.cat([ri, pad_i], 1), torch.cat([rd, pad_d], 1)
                rgi = torch.where(ri >= 0, g32[ri.clamp(min=0).long()], ri)
            else:
                rgi = torch.empty((0, k), dtype=torch.int32, device=dev)
                rd = torch.empty((0, k), dtype=torch.float32, device=dev)
            t_remote = _tick()
            # answers travel back along the same routes
            ans = torch.cat([rgi.view(torch.float32), rd], 1)
            back, _ = _all_to_all_rows(ans, rc, self.group)
            bi, bd = back[:, :k].contiguous().view(torch.int32), back[:, k:].contiguous()
            # merge per source rank on (d2, global index)
            off = 0
            for s in range(w):
                m = send_counts[s]
                if m == 0:
                    continue
                rows = send_rows[s]
                ip = torch.stack([gidx[rows], bi[off:off + m]]).contiguous()
                dp = torch.stack([d2[rows], bd[off:off + m]]).contiguous()
                _settle(dp)
                mi, md = eng.merge_topk(ip, dp)
                gidx[rows] = torch.as_tensor(mi).to(dev)
                d2[rows] = torch.as_tensor(md).to(dev)
                off += m
            t_merge = _tick()
        dist_out = torch.where(gidx >= 0, torch.sqrt(d2), d2)
        t_end = _tick()
        self.stats["phase_s"] = {"local_search": t_local - t_0, "boundary_detect": t_detect - t_local,
                                 "exchange_out": t_exchange - t_detect, "remote_search": t_remote - t_exchange,
                                 "return_and_merge": t_merge - t_remote, "finish": t_end - t_merge}
        return self.gid, gidx, dist_out
