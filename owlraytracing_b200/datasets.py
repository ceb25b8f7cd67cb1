"""Synthetic point clouds of BASELINE.json's configs, CPU side.

`uniform` is the stateless, index-addressable hash cloud — bit-identical to the device generator
`tknn_generate_uniform`, so any rank or verifier can regenerate point i without communication.
`lidar_like` is the clustered cfg3 cloud (host only: it needs log/cos, which are not bit-portable).
"""
from __future__ import annotations

import numpy as np

_PHI = np.uint64(0x9E3779B97F4A7C15)
_M1 = np.uint64(0xBF58476D1CE4E5B9)
_M2 = np.uint64(0x94D049BB133111EB)


def _mix64(z: np.ndarray) -> np.ndarray:
    z = (z ^ (z >> np.uint64(30))) * _M1
    z = (z ^ (z >> np.uint64(27))) * _M2
    return z ^ (z >> np.uint64(31))


def hash_u01(seed: int, counter: np.ndarray) -> np.ndarray:
    """u = (mix64(seed ^ counter * phi64) >> 40) * 2^-24 in [0, 1), float32-exact."""
    with np.errstate(over="ignore"):
        h = _mix64(np.uint64(seed) ^ (counter.astype(np.uint64) * _PHI))
    return (h >> np.uint64(40)).astype(np.float32) * np.float32(2.0 ** -24)


def uniform(n: int, seed: int = 42, first: int = 0) -> np.ndarray:
    """Points first..first+n of the uniform [0,1)^3 cloud: coordinate a of point i uses counter 3i+a."""
    c = np.arange(3 * first, 3 * (first + n), dtype=np.uint64)
    return hash_u01(seed, c).reshape(n, 3)


def _u(seed: int, stream: int, n: int) -> np.ndarray:
    """float64 uniforms in (0, 1) from an independent hash stream."""
    c = np.arange(n, dtype=np.uint64) + (np.uint64(stream) << np.uint64(40))
    with np.errstate(over="ignore"):
        h = _mix64(np.uint64(seed) ^ (c * _PHI))
    return ((h >> np.uint64(11)).astype(np.float64) + 0.5) * (2.0 ** -53)


def _normal(seed: int, stream: int, n: int) -> np.ndarray:
    u1, u2 = _u(seed, stream, n), _u(seed, stream + 1, n)
    return np.sqrt(-2.0 * np.log(u1)) * np.cos(2.0 * np.pi * u2)


def lidar_like(n: int, seed: int = 7) -> np.ndarray:
    """cfg3: skewed-density LiDAR-like cloud (metres).

    70 % ground returns on 64 rings (log-spaced radii 2..100 m, i.e. radial density ~ 1/rho,
    z ~ N(0, 0.02^2)); 25 % in 2 000 Gaussian object clusters (sigma log-uniform 0.05..2 m, centres
    uniform in 200 x 200 x 6 m); 4.9 % uniform clutter; 0.1 % exact duplicates of earlier points.
    """
    n_ground = int(0.70 * n)
    n_obj = int(0.25 * n)
    n_dup = int(0.001 * n)
    n_clut = n - n_ground - n_obj - n_dup
    parts = []
    # ground rings
    ring = np.minimum((_u(seed, 1, n_ground) * 64).astype(np.int64), 63)
    rho = 2.0 * (50.0 ** (ring / 63.0)) + 0.02 * _normal(seed, 2, n_ground)
    ang = 2.0 * np.pi * _u(seed, 4, n_ground)
    parts.append(np.stack([rho * np.cos(ang), rho * np.sin(ang), 0.02 * _normal(seed, 5, n_ground)], 1))
    # object clusters
    n_cl = 2000
    cx = (_u(seed, 10, n_cl) - 0.5) * 200.0
    cy = (_u(seed, 11, n_cl) - 0.5) * 200.0
    cz = _u(seed, 12, n_cl) * 6.0
    sig = 0.05 * (40.0 ** _u(seed, 13, n_cl))
    which = np.minimum((_u(seed, 14, n_obj) * n_cl).astype(np.int64), n_cl - 1)
    parts.append(np.stack([cx[which] + sig[which] * _normal(seed, 15, n_obj),
                           cy[which] + sig[which] * _normal(seed, 17, n_obj),
                           cz[which] + sig[which] * _normal(seed, 19, n_obj)], 1))
    # clutter
    parts.append(np.stack([(_u(seed, 21, n_clut) - 0.5) * 200.0, (_u(seed, 22, n_clut) - 0.5) * 200.0,
                           _u(seed, 23, n_clut) * 6.0], 1))
    pts = np.concatenate(parts, 0).astype(np.float32)
    # interleave the three populations deterministically so that file order is not sorted by kind
    perm = np.argsort(hash_u01(seed + 1, np.arange(pts.shape[0], dtype=np.uint64)), kind="stable")
    pts = pts[perm]
    if n_dup:
        src = np.minimum((_u(seed, 30, n_dup) * pts.shape[0]).astype(np.int64), pts.shape[0] - 1)
        pts = np.concatenate([pts, pts[src]], 0)
    return np.ascontiguousarray(pts, dtype=np.float32)


def lattice(m: int) -> np.ndarray:
    """m^3 unit lattice: the hand-computable known-answer case (shells of equal distance, index ties)."""
    g = np.arange(m, dtype=np.float32)
    return np.ascontiguousarray(np.stack(np.meshgrid(g, g, g, indexing="ij"), -1).reshape(-1, 3))
