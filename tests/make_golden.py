#!/usr/bin/env python
"""Generates tests/golden/*.npz from the oracle on seeded synthetic inputs (small, committed).
The reference cannot be run here (ROS/PCL/Ceres missing) and ships no golden vectors; these are
regression pins of the oracle itself and the fixtures the GPU parity tests are checked against."""
import os, sys
import numpy as np
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE); sys.path.insert(0, os.path.dirname(HERE))
import oracle_lib as O
from liliom_b200 import synth

out = os.path.join(HERE, "golden")
os.makedirs(out, exist_ok=True)
T = synth.default_true_pose()

# Horizon: the first 1500 time columns of a sweep (the 6x4000 grid logic needs the real column encoding)
pts, q = synth.make_horizon_sweep(T, seed=11)
frac = pts["intensity"] - np.floor(pts["intensity"])
pts = pts[frac < 0.1 * 1500 / 3999.0]
surf, edge, cut = O.extract_horizon(pts, q, 0.2, 4.0)
np.savez_compressed(os.path.join(out, "horizon_small.npz"), pts=pts.view(np.uint8), q_imu=q, surf_thres=0.2, edge_thres=4.0,
                    surf=surf.view(np.uint8), edge=edge.view(np.uint8), cut=cut.view(np.uint8), surf_ds=O.voxelgrid(surf, 0.4).view(np.uint8))
print("horizon", len(pts), len(surf), len(edge))

# ROT: every 4th azimuth step of an HDL-64E sweep
hdl, q2 = synth.make_hdl64_sweep(T, seed=12, steps=500)
q_lb = np.array([0.9995, 0.01, -0.02, 0.02]); q_lb /= np.linalg.norm(q_lb)
rc, s2, e2, c2, lab, cur = O.extract_rot(hdl, q2, q_lb, 64, 2)
np.savez_compressed(os.path.join(out, "rot_small.npz"), pts=hdl.view(np.uint8), q_imu=q2, q_lb=q_lb, line_num=64, ds_rate=2,
                    surf=s2.view(np.uint8), edge=e2.view(np.uint8), cut=c2.view(np.uint8), label=lab, curv=cur)
print("rot", len(hdl), len(s2), len(e2), len(c2))

# scan-to-map: 20k-point map patch, the down-sampled surf features of a full sweep
m, _ = synth.make_map(20000)
pts, q = synth.make_horizon_sweep(T, seed=13)
surf, _, _ = O.extract_horizon(pts, q)
ds = O.voxelgrid(surf, 0.4)
feats = np.ones((len(ds), 4), np.float32); feats[:, 0] = ds["x"]; feats[:, 1] = ds["y"]; feats[:, 2] = ds["z"]
pose0 = synth.perturbed_pose(T)
tree = O.KdTree(m)
cnt, valid, plane, idx, pw = O.find_surf_corr(tree, feats, pose0)
neq = O.normal_equations(feats, valid, plane, pose0)
rc, pose_gn, _ = O.scan_to_map_gn(tree, feats, pose0, 6)
rc, pose_ce, st = O.scan_to_map_ceres(tree, feats, pose0, 2, 15)
np.savez_compressed(os.path.join(out, "s2m_small.npz"), map=m, feats=feats, pose0=pose0, valid=valid, plane=plane, nn_idx=idx, neq29=neq,
                    pose_gn6=pose_gn, pose_ceres=pose_ce, ceres_lm_iters=np.array([s.lm_iters for s in st]))
print("s2m", len(m), len(feats), cnt, [s.lm_iters for s in st])
