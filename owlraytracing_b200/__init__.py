"""owlraytracing_b200 — B200-native TrueKNN hot path (LBVH build + warp-cooperative traversal).

Only what the path needs: `csrc/` (CUDA kernels + the C ABI of include/trueknn.h), the ctypes
loader, the host-side mirror of the sample's interface, the synthetic clouds of BASELINE.json and
the two multi-GPU drivers.  Importing the package does not require a GPU; creating a `TrueKNN`
does (there is no CPU fallback).
"""
from .trueknn import MultiTrueKNN, TrueKNN, TrueKNNError, read_points, read_points_fast, run_sample, write_neighbours  # noqa: F401
from . import datasets  # noqa: F401

__all__ = ["TrueKNN", "MultiTrueKNN", "TrueKNNError", "read_points", "read_points_fast", "write_neighbours", "run_sample", "datasets"]
