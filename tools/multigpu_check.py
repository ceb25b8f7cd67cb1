"""Multi-GPU parity check over NCCL, one rank per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
        tools/multigpu_check.py [n_points] [k]

Runs the library's query-sharded path (tknn_build_replicated + tknn_search_shard) and its point-partitioned path
(tknn_partition_build / _search / _verify) on a uniform and a LiDAR-like cloud and compares EVERY result row with the
CPU oracle (kd-tree) on rank 0.  torch.distributed only carries the NCCL unique id and gathers the results for the
check.  Prints one JSON line per case; exits non-zero on any mismatch.
"""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from owlraytracing_b200 import TrueKNN, datasets  # noqa: E402


def gather_rows(world, dev, k, ids, idx, dst):
    """padded all_gather of (ids [m], idx [m,k], dist [m,k]) -> host arrays of all ranks' rows"""
    m = torch.tensor([ids.numel()], device=dev)
    ms = [torch.zeros_like(m) for _ in range(world)]
    dist.all_gather(ms, m)
    cap = int(max(int(v.item()) for v in ms))
    pg = torch.full((cap,), -1, dtype=torch.int32, device=dev); pg[:ids.numel()] = ids
    pi = torch.full((cap, k), -1, dtype=torch.int32, device=dev); pi[:ids.numel()] = idx
    pd = torch.zeros((cap, k), dtype=torch.float32, device=dev); pd[:ids.numel()] = dst
    G, I, D = (torch.empty((world * cap,) + t.shape[1:], dtype=t.dtype, device=dev) for t in (pg, pi, pd))
    dist.all_gather_into_tensor(G, pg); dist.all_gather_into_tensor(I, pi); dist.all_gather_into_tensor(D, pd)
    keep = (G >= 0).cpu().numpy()
    return G.cpu().numpy()[keep], I.cpu().numpy()[keep], D.cpu().numpy()[keep]


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 400_000
    k = int(sys.argv[2]) if len(sys.argv) > 2 else 10
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    ok_all = True
    eng = TrueKNN(local)
    eng.set_stream(torch.cuda.current_stream().cuda_stream)
    eng.comm_init_torch()
    for kind in ("uniform", "lidar"):
        x = datasets.uniform(n, 42) if kind == "uniform" else datasets.lidar_like(n, 7)
        ref = None
        if rank == 0:
            from oracle import oracle as O

            ref = O.knn_kdtree(x, k)
        # ---- query-sharded: every rank uploads its ceil(n/N)-row slice from the host ----
        per = (n + world - 1) // world
        lo = min(n, per * rank)
        t0 = time.perf_counter()
        eng.build_replicated(x[lo: min(n, lo + per)], lo, n)
        qid, si, sd = eng.search_shard(k, rank, world)
        t_sh = time.perf_counter() - t0
        sh_stats = eng.dist_stats()
        sg, sI, sD = gather_rows(world, dev, k, torch.from_numpy(np.asarray(qid)).to(dev), torch.from_numpy(np.asarray(si)).to(dev),
                                 torch.from_numpy(np.asarray(sd)).to(dev))
        # ---- point-partitioned: rank r starts with an uneven contiguous index slice, on the device ----
        cuts = [n * r // world for r in range(world)] + [n]
        cuts = [0] + [min(n, c + (997 * (i + 1)) % 5000) for i, c in enumerate(cuts[1:-1])] + [n]
        lo, hi = cuts[rank], cuts[rank + 1]
        xd = torch.from_numpy(x[lo:hi]).to(dev)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        eng.partition_build(xd, lo)
        gid, pi, pd = eng.partition_search(k)
        t_pp = time.perf_counter() - t0
        pst = eng.dist_stats()
        checked, bad = eng.partition_verify(k, 512, gid, pi, pd)
        pg, pI, pD = gather_rows(world, dev, k, gid, pi, pd)
        sent = torch.tensor([pst["boundary_sent"]], device=dev)
        dist.all_reduce(sent)
        if rank == 0:
            ri, rd = ref
            ok_sh = bool(np.array_equal(np.sort(sg), np.arange(n)) and (sI == ri[sg]).all() and np.allclose(sD, rd[sg], rtol=1e-6, atol=0))
            ok_pp = bool(np.array_equal(np.sort(pg), np.arange(n)) and (pI == ri[pg]).all() and np.allclose(pD, rd[pg], rtol=1e-6, atol=0))
            ok_v = checked == 512 * world and bad == 0
            ok_all = ok_all and ok_sh and ok_pp and ok_v
            print(json.dumps({"cloud": kind, "n": n, "k": k, "world": world, "transport": "nccl", "sharded_ok": ok_sh,
                              "partitioned_ok": ok_pp, "verify_checked": checked, "verify_bad": bad,
                              "sharded_s": round(t_sh, 4), "partitioned_s": round(t_pp, 4), "boundary_queries": int(sent.item()),
                              "allgather_ms": round(sh_stats["allgather_ms"], 3),
                              "partition_build_ms": round(pst["build_total_ms"], 3),
                              "partition_search_ms": round(pst["search_total_ms"], 3)}), flush=True)
    dist.barrier()
    eng.close()
    dist.destroy_process_group()
    if rank == 0 and not ok_all:
        sys.exit(1)


if __name__ == "__main__":
    main()
