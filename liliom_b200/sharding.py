"""Host-side statement of the multi-GPU partition rule implemented in csrc/ (knn_core.cuh
`owner_of`, api.cu `k_shard_flags`): space is cut into 16 m blocks, a block belongs to rank
hash(block) % nranks, a query is processed by the owner of the block its transformed position
falls in, and a rank's map shard holds every point whose +-halo box touches a block it owns (so the
query's whole 1 m search ball is local).  Used by bench.py to report shard sizes and by the gloo tests."""
from __future__ import annotations

import numpy as np


def _block_hash(b: np.ndarray) -> np.ndarray:
    b = b.astype(np.int64)
    with np.errstate(over="ignore"):
        h = ((b[..., 0].astype(np.uint32) * np.uint32(73856093)) ^ (b[..., 1].astype(np.uint32) * np.uint32(19349663))
             ^ (b[..., 2].astype(np.uint32) * np.uint32(83492791))).astype(np.uint32)
        h ^= h >> np.uint32(15)
        h = (h * np.uint32(0x2c1b3c6d)).astype(np.uint32)
        h ^= h >> np.uint32(12)
    return h


def owner_of(xyz: np.ndarray, nranks: int, block: int = 16) -> np.ndarray:
    xyz = np.asarray(xyz, np.float32)
    b = np.floor(xyz * np.float32(1.0 / block)).astype(np.int32)
    return (_block_hash(b) % np.uint32(nranks)).astype(np.int32)


def shard_mask(xyz: np.ndarray, nranks: int, rank: int, halo: float = 1.0, block: int = 16) -> np.ndarray:
    xyz = np.asarray(xyz, np.float32)
    keep = np.zeros(len(xyz), bool)
    h = np.float32(halo)
    for c in range(8):
        off = np.array([h if c & 1 else -h, h if c & 2 else -h, h if c & 4 else -h], np.float32)
        keep |= owner_of(xyz + off, nranks, block) == rank
    return keep


def transform_f32(p_xyz: np.ndarray, pose7) -> np.ndarray:
    """q * p + t in fp64 (Eigen's expression), narrowed to fp32 (L/src/LidarOdometry.cpp:221-238)."""
    q = np.asarray(pose7[:4], np.float64); t = np.asarray(pose7[4:], np.float64)
    p = np.asarray(p_xyz, np.float64)
    uv = 2.0 * np.cross(q[1:], p)
    return (p + q[0] * uv + np.cross(q[1:], uv) + t).astype(np.float32)
