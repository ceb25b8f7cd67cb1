"""Point-partitioned multi-GPU TrueKNN (BASELINE.json configs[4]) — one rank per process.

Thin caller of the library: the whole pipeline (global box, Morton codes, cell-aligned splitters, all-to-all
redistribution, local LBVH carrying global ids, partition summaries, local search, boundary-query exchange,
remote bounded search, merge on (d2, global index)) runs inside libtrueknn over NCCL
(owlraytracing_b200/csrc/dist.cu: tknn_partition_build / tknn_partition_search / tknn_partition_verify).
`torch.distributed` is used for ONE thing: carrying the 128-byte NCCL unique id from rank 0 to the others.

The reference cannot hold such clouds at all: one accel is limited to 2^29 primitives
(owl/UserGeomGroup.cpp:87-117) and its multi-GPU model only replicates (owl/RayGen.cpp:150-200).
"""
from __future__ import annotations


class PartitionedTrueKNN:
    def __init__(self, engine=None, device: int | None = None, group=None):
        import torch
        import torch.distributed as dist

        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        if engine is None:
            from .trueknn import TrueKNN

            engine = TrueKNN(torch.cuda.current_device() if device is None else device)
        self.engine = engine
        if getattr(engine, "n_ranks", None) is None:
            engine.comm_init_torch(group)
        self.stats = {}

    def build(self, local_points, first_index: int):
        """local_points [n, 3] float32 (CUDA tensor or host array); their global indices are first_index + arange(n)."""
        self.engine.partition_build(local_points, first_index)
        self.n_owned = self.engine.partition_owned()
        self.stats = self.engine.dist_stats()
        return self

    def search(self, k: int, start_radius: float = 0.0, out=None):
        """Exact kNN of the owned points against the GLOBAL cloud: (gid [m] int32, idx [m, k] int32 global neighbour
        indices, dist [m, k] float32)."""
        res = self.engine.partition_search(k, start_radius, out=out)
        self.stats = self.engine.dist_stats()
        return res

    def verify(self, k: int, samples: int, gid, idx, dist):
        """Distributed GPU brute force over `samples` rows per rank (no replica needed): (checked, bad), global sums."""
        return self.engine.partition_verify(k, samples, gid, idx, dist)
