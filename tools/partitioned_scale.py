"""Point-partitioned kNN at scale (BASELINE.json configs[4] shape), one rank per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29515 \
        tools/partitioned_scale.py <total_points> [k] [verify_samples]

Every rank generates its contiguous index slice of the uniform hash cloud on its GPU, the driver redistributes
by Morton range, builds local LBVHs, searches, exchanges boundary queries and merges partial top-k lists over NCCL.
Verification: `verify_samples` owned queries per rank are re-answered by the exact GPU brute-force kernel over a
replica of the WHOLE cloud (an independent path: no BVH culling, no exchange, no merge).  One JSON line on rank 0.
"""
import json
import os
import sys
import time

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from owlraytracing_b200 import TrueKNN  # noqa: E402
from owlraytracing_b200.partitioned import PartitionedTrueKNN  # noqa: E402


def main():
    n_total = int(sys.argv[1]) if len(sys.argv) > 1 else 50_000_000
    k = int(sys.argv[2]) if len(sys.argv) > 2 else 10
    samples = int(sys.argv[3]) if len(sys.argv) > 3 else 2048
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    eng = TrueKNN(local)
    lo, hi = n_total * rank // world, n_total * (rank + 1) // world
    x = torch.empty((hi - lo, 3), dtype=torch.float32, device=dev)
    eng.generate_uniform(42, lo, hi - lo, out=x)
    drv = PartitionedTrueKNN(engine=eng)
    dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    drv.build(x, lo)
    dist.barrier(); torch.cuda.synchronize()
    t1 = time.perf_counter()
    # integrity of the redistribution: every global id owned exactly once
    g = drv.gid
    own = torch.tensor([g.numel(), int(g.min().item()), int(g.max().item()), int((g[1:] == g[:-1]).sum().item())], device=dev)
    tot = own.clone(); dist.all_reduce(tot)
    if rank == 0:
        print("owned/min/max/dups on rank 0:", own.tolist(), "total owned:", int(tot[0].item()), "total dups:", int(tot[3].item()), flush=True)
    gid, idx, dst = drv.search(k)
    dist.barrier(); torch.cuda.synchronize()
    t2 = time.perf_counter()
    gid, idx, dst = drv.search(k)          # second search: steady state (allocations done)
    dist.barrier(); torch.cuda.synchronize()
    t3 = time.perf_counter()
    stats = dict(drv.stats)
    sent = torch.tensor([stats["boundary_sent"], stats["owned"]], device=dev, dtype=torch.int64)
    dist.all_reduce(sent)
    peak_gb = torch.cuda.max_memory_allocated(dev) / 1e9
    ok = True
    if samples > 0:
        del x
        torch.cuda.empty_cache()
        ver = TrueKNN(local)
        full = torch.empty((n_total, 3), dtype=torch.float32, device=dev)
        ver.generate_uniform(42, 0, n_total, out=full)
        ver.build(full)
        pick = torch.randperm(gid.numel(), device=dev, generator=torch.Generator(device=dev).manual_seed(rank))[:samples]
        ids = gid[pick].to(torch.int32).contiguous()
        bi, bd = ver.brute_force(ids, k)
        ok = bool(torch.equal(bi, idx[pick]) and torch.allclose(bd, dst[pick], rtol=1e-6, atol=0))
        ver.close()
    all_phases = [None] * world
    dist.all_gather_object(all_phases, {a: round(b, 4) for a, b in stats.get("phase_s", {}).items()})
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(json.dumps({"n_total": n_total, "world": world, "k": k, "build_s": round(t1 - t0, 3), "search_first_s": round(t2 - t1, 3),
                          "search_s": round(t3 - t2, 3), "queries_per_s": n_total / (t3 - t2),
                          "boundary_queries": int(sent[0].item()), "boundary_fraction": float(sent[0].item()) / n_total,
                          "verified_samples_per_rank": samples, "verified_ok": bool(flag.item()),
                          "torch_peak_alloc_gb_rank0": round(peak_gb, 2),
                          "phase_s_per_rank": all_phases}), flush=True)
    dist.barrier()
    dist.destroy_process_group()
    if not bool(flag.item()):
        sys.exit(1)


if __name__ == "__main__":
    main()
