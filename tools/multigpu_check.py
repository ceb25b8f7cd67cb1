"""Multi-GPU parity check, one rank per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
        tools/multigpu_check.py [n_points] [k]

Runs the query-sharded and the point-partitioned drivers over NCCL on a uniform and a LiDAR-like cloud and
compares every result with the CPU oracle (kd-tree) on rank 0.  Prints one JSON line per case.
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
from owlraytracing_b200.partitioned import PartitionedTrueKNN  # noqa: E402
from owlraytracing_b200.sharded import ShardedTrueKNN  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 400_000
    k = int(sys.argv[2]) if len(sys.argv) > 2 else 10
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    ok_all = True
    for kind in ("uniform", "lidar"):
        x = datasets.uniform(n, 42) if kind == "uniform" else datasets.lidar_like(n, 7)
        xd = torch.from_numpy(x).to(dev)
        ref = None
        if rank == 0:
            from oracle import oracle as O

            ref = O.knn_kdtree(x, k)
        # ---- query-sharded ----
        t0 = time.perf_counter()
        drv = ShardedTrueKNN(engine=TrueKNN(local)).build(xd)
        gi, gd = drv.search(k, gather=True)
        torch.cuda.synchronize()
        t_sh = time.perf_counter() - t0
        # ---- point-partitioned: rank r starts with the r-th contiguous index slice ----
        lo, hi = n * rank // world, n * (rank + 1) // world
        t0 = time.perf_counter()
        pdrv = PartitionedTrueKNN(engine=TrueKNN(local)).build(xd[lo:hi].contiguous(), lo)
        gid, pi, pd = pdrv.search(k)
        torch.cuda.synchronize()
        t_pp = time.perf_counter() - t0
        # gather the partitioned results on rank 0 (padded all_gather)
        m = torch.tensor([gid.numel()], device=dev)
        ms = [torch.zeros_like(m) for _ in range(world)]
        dist.all_gather(ms, m)
        cap = int(max(int(v.item()) for v in ms))
        pg = torch.full((cap,), -1, dtype=torch.int64, device=dev); pg[:gid.numel()] = gid
        pii = torch.full((cap, k), -1, dtype=torch.int32, device=dev); pii[:gid.numel()] = pi
        pdd = torch.zeros((cap, k), dtype=torch.float32, device=dev); pdd[:gid.numel()] = pd
        G, I, D = (torch.empty((world * cap,) + t.shape[1:], dtype=t.dtype, device=dev) for t in (pg, pii, pdd))
        dist.all_gather_into_tensor(G, pg); dist.all_gather_into_tensor(I, pii); dist.all_gather_into_tensor(D, pdd)
        sent = torch.tensor([pdrv.stats["boundary_sent"]], device=dev)
        dist.all_reduce(sent)
        if rank == 0:
            ri, rd = ref
            ok_sh = bool((gi.cpu().numpy() == ri).all() and np.allclose(gd.cpu().numpy(), rd, rtol=1e-6, atol=0))
            keep = (G >= 0).cpu().numpy()
            g = G.cpu().numpy()[keep]
            ok_pp = bool(np.array_equal(np.sort(g), np.arange(n)) and (I.cpu().numpy()[keep] == ri[g]).all()
                         and np.allclose(D.cpu().numpy()[keep], rd[g], rtol=1e-6, atol=0))
            ok_all = ok_all and ok_sh and ok_pp
            print(json.dumps({"cloud": kind, "n": n, "k": k, "world": world, "sharded_ok": ok_sh, "partitioned_ok": ok_pp,
                              "sharded_s": round(t_sh, 4), "partitioned_s": round(t_pp, 4), "boundary_queries": int(sent.item())}),
                  flush=True)
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0 and not ok_all:
        sys.exit(1)


if __name__ == "__main__":
    main()
