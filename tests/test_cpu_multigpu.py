"""world_size > 1 on CPU (gloo): the host logic of the query-sharded driver (shard bookkeeping, slice upload +
all_gather, padded gather of the results) and the protocol MODEL of the point-partitioned variant
(tests/partition_model.py: Morton-range redistribution, boundary-query routing, merge on (d2, global index)) — the
product runs that protocol inside libtrueknn over NCCL (csrc/dist.cu) and is tested on the GPU
(tests/test_gpu_multi.py) — with a CPU stand-in engine answering the local searches through the oracle."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _cloud(kind, n):
    from owlraytracing_b200 import datasets

    if kind == "uniform":
        return datasets.uniform(n, seed=11)
    if kind == "lidar":
        return datasets.lidar_like(n, seed=5)
    if kind == "dups":
        return np.concatenate([datasets.uniform(n // 2, seed=3)] * 2).astype(np.float32)
    raise ValueError(kind)


def _worker(rank, world, port, mode, kind, n, k, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from cpu_engine import CpuEngine

        x = _cloud(kind, n)
        if mode == "sharded":
            from owlraytracing_b200.sharded import ShardedTrueKNN

            if kind == "uniform":
                drv = ShardedTrueKNN(engine=CpuEngine()).build(x)
            else:  # every rank contributes only its 1/N slice; one all_gather replicates the cloud
                per = (n + world - 1) // world
                drv = ShardedTrueKNN(engine=CpuEngine()).build_from_slices(x[rank * per: min(n, (rank + 1) * per)], n)
            qid, idx, dst = drv.search(k)
            gi, gd = drv.search(k, gather=True)
            np.savez(os.path.join(out_dir, f"r{rank}.npz"), qid=np.asarray(qid), idx=np.asarray(idx), dist=np.asarray(dst),
                     gidx=gi.numpy(), gdist=gd.numpy())
        else:
            from partition_model import PartitionModel

            # every rank starts with an arbitrary contiguous slice of the global index space
            lo, hi = n * rank // world, n * (rank + 1) // world
            drv = PartitionModel(engine=CpuEngine()).build(torch.from_numpy(x[lo:hi]), lo)
            gid, idx, dst = drv.search(k)
            np.savez(os.path.join(out_dir, f"r{rank}.npz"), gid=gid.numpy(), idx=idx.numpy(), dist=dst.numpy(),
                     sent=drv.stats["boundary_sent"], owned=drv.stats["owned"])
    finally:
        dist.destroy_process_group()


def _run(world, mode, kind, n, k, tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(world, port, mode, kind, n, k, str(tmp_path)), nprocs=world, join=True)
    return [np.load(os.path.join(str(tmp_path), f"r{r}.npz")) for r in range(world)]


@pytest.mark.parametrize("world,kind", [(2, "uniform"), (3, "uniform"), (3, "lidar")])
def test_query_sharded_driver(world, kind, oracle, tmp_path):
    n, k = 2500, 6
    parts = _run(world, "sharded", kind, n, k, tmp_path)
    ref_i, ref_d = oracle.knn_brute(_cloud(kind, n), k)
    seen = np.zeros(n, bool)
    for p in parts:
        assert not seen[p["qid"]].any()
        seen[p["qid"]] = True
        assert (p["idx"] == ref_i[p["qid"]]).all() and (p["dist"] == ref_d[p["qid"]]).all()
        # gather=True gives every rank the full arrays in build order
        assert (p["gidx"] == ref_i).all() and (p["gdist"] == ref_d).all()
    assert seen.all()
    sizes = [len(p["qid"]) for p in parts]
    assert max(sizes) - min(sizes) <= 32  # contiguous Morton slices differ by at most one group


@pytest.mark.parametrize("world,kind,k", [(2, "uniform", 5), (3, "uniform", 10), (2, "lidar", 8), (2, "dups", 4)])
def test_point_partitioned_driver(world, kind, k, oracle, tmp_path):
    n = 3000
    parts = _run(world, "partitioned", kind, n, k, tmp_path)
    ref_i, ref_d = oracle.knn_brute(_cloud(kind, n), k)
    owned = np.concatenate([p["gid"] for p in parts])
    assert np.array_equal(np.sort(owned), np.arange(n))  # every point has exactly one owner
    for p in parts:
        g = p["gid"]
        bad = np.nonzero((p["idx"] != ref_i[g]).any(1))[0]
        assert bad.size == 0, (kind, g[bad[:3]], p["idx"][bad[:3]], ref_i[g[bad[:3]]])
        assert np.allclose(p["dist"], ref_d[g], rtol=1e-6, atol=0)
    # the exchange really happened, and only for a minority of the queries on non-degenerate clouds
    assert sum(int(p["sent"]) for p in parts) > 0
    if kind == "uniform":
        assert sum(int(p["sent"]) for p in parts) < n
