#!/usr/bin/env python
"""Kernel-shape sweep for k_knn_plane (GPU only): CUDA-event time per launch for several
(lanes, rounds) settings, on the down-sampled queries of one sweep and on the dense surf set."""
import os, sys, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import liliom_b200 as L
from liliom_b200 import synth

n_map = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
m, _ = synth.make_map(n_map)
T = synth.default_true_pose()
pts, q = synth.make_horizon_sweep(T)
guess = synth.perturbed_pose(T)
base = L.Context(variant=0)
surf, edge, cut = base.extract_horizon(pts, q)
ds = base.voxelgrid(surf, 0.4)
base.close()
dense = np.concatenate([surf] * 1)
big = np.concatenate([surf[np.random.default_rng(0).permutation(len(surf))] for _ in range(8)])   # 128k queries, unsorted
hdl, q2 = synth.make_hdl64_sweep(T)
hdl48 = np.zeros(len(hdl), L.PT48); hdl48["x"] = hdl["x"]; hdl48["y"] = hdl["y"]; hdl48["z"] = hdl["z"]
big_sorted = np.concatenate([surf] * 8)
for name, feats in (("ds", ds), ("hdl_ring_order", hdl48), ("dense_x8_in_order", big_sorted), ("dense_x8_shuffled", big)):
    for lanes, rounds in ((0, 0), (1, 1)):
        if lanes: os.environ["LILIOM_KNN_LANES"] = str(lanes); os.environ["LILIOM_KNN_ROUNDS"] = str(rounds)
        else: os.environ.pop("LILIOM_KNN_LANES", None); os.environ.pop("LILIOM_KNN_ROUNDS", None)
        c = L.Context(variant=0)
        c.map_set_points(m)
        c.upload_feats(feats)
        c.set_kernel_timing(True)
        for _ in range(3): c.scan_to_map_resident(guess, 10, mode=L.MODE_GN)
        c.counters(reset=True)
        for _ in range(5): pose, _ = c.scan_to_map_resident(guess, 10, mode=L.MODE_GN)
        k = c.counters()
        us = 1e3 * k.knn_ms / k.knn_launches
        cbar = k.knn_candidates / k.knn_queries
        gbs = (len(feats) * (16 + 216 + 16 * cbar)) / (us * 1e-6) / 1e9
        print(json.dumps(dict(feats=name, n=len(feats), lanes=lanes or "auto", rounds=rounds or "auto", us_per_launch=round(us, 2), cbar=round(cbar, 1), alg_GBs=round(gbs, 1),
                              err=float(np.linalg.norm(pose[4:] - T[4:])))), flush=True)
        c.close()
