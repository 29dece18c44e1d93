#!/usr/bin/env python
"""Streamed sequence through the reference's REAL per-scan lifecycle: every scan pushes
its down-sampled surf cloud into the 20-frame local map, the map is re-voxelised and re-indexed, then
scan-to-map runs — on the GPU through the node mirror (csrc/host/nodes.cpp -> C ABI), and on the CPU through
the oracle restatement (concat + VoxelGrid + kd-tree build + Ceres-faithful solve), 1 thread like the reference node.
usage: stream_bench.py [n_scans] [mode 0=CERES|1=GN]"""
import os, sys, time, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import liliom_b200 as L
from liliom_b200 import synth
import oracle_lib as O
from test_gpu_nodes import OracleLO

n_scans = int(sys.argv[1]) if len(sys.argv) > 1 else 40
mode = int(sys.argv[2]) if len(sys.argv) > 2 else 0
T0 = synth.default_true_pose()
scans = []
for k in range(n_scans):
    T = np.array(T0); T[4] += 0.10 * k; T[5] += 0.01 * k
    T[:4] = synth.qmul(synth.q_from_axis_angle([0, 0, 1], np.deg2rad(0.3 * k)), T0[:4])
    pts, q = synth.make_horizon_sweep(T, seed=100 + k)
    scans.append((pts, q))

ctx_pre = L.Context(variant=0); ctx_lo = L.Context(variant=0)
node = L.LidarOdometryNode(ctx_lo, max_num_iter=15, scan_match_cnt=2, if_to_deskew=False, mode=mode)
feats = []
t0 = time.perf_counter()
for pts, q in scans:
    feats.append(ctx_pre.extract_horizon(pts, q))
t_ext_gpu = time.perf_counter() - t0
poses_gpu = []
t0 = time.perf_counter()
for k, (surf, edge, cut) in enumerate(feats):
    node.feed(0.1 * k, edge, surf, cut)
    out, *_ = node.run(want_clouds=False)
    poses_gpu.append(np.array(out.abs_pose))
t_lo_gpu = time.perf_counter() - t0

t0 = time.perf_counter()
feats_cpu = [O.extract_horizon(pts, q) for pts, q in scans]
t_ext_cpu = time.perf_counter() - t0
ref = OracleLO(O, 15, 2, mode)
poses_cpu = []
t0 = time.perf_counter()
for surf, edge, cut in feats_cpu:
    r = ref.run(surf)
    poses_cpu.append(ref.abs.copy())
t_lo_cpu = time.perf_counter() - t0
dmax = max(np.linalg.norm(a[4:] - b[4:]) for a, b in zip(poses_gpu, poses_cpu))
print(json.dumps(dict(scans=n_scans, mode="CERES" if mode == 0 else "GN",
                      gpu_scans_per_s=round(n_scans / (t_ext_gpu + t_lo_gpu), 1), cpu_scans_per_s=round(n_scans / (t_ext_cpu + t_lo_cpu), 1),
                      gpu_ms=dict(extract=round(1e3 * t_ext_gpu / n_scans, 3), odometry=round(1e3 * t_lo_gpu / n_scans, 3)),
                      cpu_ms=dict(extract=round(1e3 * t_ext_cpu / n_scans, 3), odometry=round(1e3 * t_lo_cpu / n_scans, 3)),
                      max_pose_delta_m=float(dmax), final_map_points=int(out.n_map))))
