#!/usr/bin/env python
"""Tiny driver for ncu: one map, one sweep, a few GN scan-to-map calls
(argv[1] = map points, argv[2] = 'ds' (down-sampled 24k sweep, the headline shape) | 'dense' (every surf feature) |
 'x8' (8 shuffled copies, 128k queries) | 'hdl' (every return of a 130k-pt HDL-64E sweep: the one-thread-per-query shape))."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import liliom_b200 as L
from liliom_b200 import synth
n_map = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
which = sys.argv[2] if len(sys.argv) > 2 else "ds"
m, _ = synth.make_map(n_map)
T = synth.default_true_pose()
pts, q = synth.make_horizon_sweep(T)
guess = synth.perturbed_pose(T)
c = L.Context(variant=0)
c.map_set_points(m)
surf, edge, cut = c.extract_horizon(pts, q)
feats = c.voxelgrid(surf, 0.4) if which == "ds" else surf
if which == "x8":
    feats = np.concatenate([surf[np.random.default_rng(0).permutation(len(surf))] for _ in range(8)])
if which == "hdl":
    hdl, _ = synth.make_hdl64_sweep(T)
    feats = np.ones((len(hdl), 4), np.float32); feats[:, 0] = hdl["x"]; feats[:, 1] = hdl["y"]; feats[:, 2] = hdl["z"]
c.upload_feats(feats)
c.set_kernel_timing(True)
for _ in range(4):
    pose, _ = c.scan_to_map_resident(guess, 10, mode=L.MODE_GN)
k = c.counters()
nq, c27 = c.knn_block_stats(guess)
print("n", len(feats), "us/pass", 1e3 * k.knn_ms / k.knn_launches, "examined/query", k.knn_candidates / max(k.knn_queries, 1), "block27/query", c27 / max(nq, 1),
      "err", np.linalg.norm(pose[4:] - T[4:]))
