"""Integration parity of the host-side node mirror (csrc/host/nodes.cpp, driving the CUDA library through
the C ABI) against a test-side restatement of the reference's node glue built on the CPU oracle:
Preprocessing::cloudHandler / processIMU (L/src/Preprocessing.cpp:129-234) and LidarOdometry::run
(L/src/LidarOdometry.cpp:280-350, 415-480, 483-586, 652-686)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _f4(c):
    out = np.ones((len(c), 4), np.float32)
    out[:, 0] = c["x"]; out[:, 1] = c["y"]; out[:, 2] = c["z"]
    return out


class OraclePre:
    """processIMU / solveRotation / the two-scans-behind queue, restated with the oracle's deltaQ product."""

    def __init__(self, O):
        self.O = O; self.imu = []; self.idx = 0; self.t_imu = -1.0; self.g0 = None; self.q = np.array([1.0, 0, 0, 0]); self.queue = []

    def imu_msg(self, t, g):
        self.imu.append((t, np.asarray(g, float)))
        if self.t_imu < 0: self.t_imu = t
        if self.g0 is None: self.g0 = np.asarray(g, float)

    def _solve(self, dt, w):
        self.O.lib().orc_solve_rotation(self.O._d(self.q), self.O._d(np.ascontiguousarray(self.g0)), self.O._d(np.ascontiguousarray(w)), float(dt))
        self.g0 = np.asarray(w, float)

    def process_imu(self, t_cur):
        r = np.zeros(3)
        i = self.idx
        if i >= len(self.imu): i -= 1
        while self.imu[i][0] < t_cur:
            t = self.imu[i][0]
            if self.t_imu < 0: self.t_imu = t
            dt = t - self.t_imu
            self.t_imu = t
            r = self.imu[i][1].copy()
            self._solve(dt, r)
            i += 1
            if i >= len(self.imu): break
        if i < len(self.imu):
            dt1 = t_cur - self.t_imu; dt2 = self.imu[i][0] - t_cur
            w1 = dt2 / (dt1 + dt2); w2 = dt1 / (dt1 + dt2)
            r = w1 * r + w2 * self.imu[i][1]
            self._solve(dt1, r)
        self.t_imu = t_cur; self.idx = i

    def cloud(self, stamp, pts):
        self.queue.append((stamp, pts))
        if len(self.queue) <= 2: return None
        st, cur = self.queue.pop(0)
        t_next = self.queue[0][0]
        self.process_imu(t_next)
        q = self.q.copy()
        surf, edge, cut = self.O.extract_horizon(cur, q)
        self.q = np.array([1.0, 0, 0, 0])
        return st, surf, edge, cut, q


def test_preprocessing_node_matches_oracle_glue(oracle, world_small):
    import liliom_b200 as L
    from liliom_b200 import synth
    ctx = L.Context(variant=0)
    node = L.PreprocessingNode(ctx)
    ref = OraclePre(oracle)
    T = world_small["T"]
    sweeps = [synth.make_horizon_sweep(T, seed=40 + k)[0][:6000] for k in range(5)]
    t_imu = 0.0
    got, want = [], []
    for k, sw in enumerate(sweeps):
        stamp = 0.1 * k
        while t_imu < stamp + 0.1501:            # IMU runs ahead of the LiDAR, 200 Hz, slowly varying rate
            g = (0.02 * np.sin(3 * t_imu), -0.01, 0.2 + 0.05 * np.cos(2 * t_imu))
            node.imu(t_imu, g); ref.imu_msg(t_imu, g)
            t_imu += 0.005
        a = node.cloud(stamp, sw); b = ref.cloud(stamp, sw)
        assert (a is None) == (b is None)
        if a is not None:
            got.append(a); want.append(b)
    assert len(got) == 3
    for a, b in zip(got, want):
        assert a[0] == b[0]
        np.testing.assert_array_equal(a[4], b[4])                      # q_iMU bit-identical
        for ca, cb in zip(a[1:4], b[1:4]):
            assert ca.view(np.uint8).tobytes() == cb.view(np.uint8).tobytes()
    node.close(); ctx.close()


class OracleLO:
    def __init__(self, O, max_num_iter, scan_match_cnt, mode):
        self.O = O; self.mni = max_num_iter; self.smc = scan_match_cnt; self.mode = mode
        self.abs = np.array([1.0, 0, 0, 0, 0, 0, 0]); self.rel = self.abs.copy()
        self.poses = []; self.frames = []; self.recent = []; self.latest = 0
        self.init = False; self.kf = True; self.kf_num = 0
        self.t_kf = np.zeros(3); self.q_kf = np.array([1.0, 0, 0, 0])

    def _compose(self, a, b):
        out = np.zeros(7); self.O.lib().orc_pose_compose(self.O._d(np.ascontiguousarray(a)), self.O._d(np.ascontiguousarray(b)), self.O._d(out)); return out

    def _relative(self, a, b):
        out = np.zeros(7); self.O.lib().orc_pose_relative(self.O._d(np.ascontiguousarray(a)), self.O._d(np.ascontiguousarray(b)), self.O._d(out)); return out

    def run(self, surf):
        O = self.O
        if not self.init:
            self.poses.append(self.abs.copy()); self.frames.append(surf[:0].copy()); self.init = True
            return None
        self.abs = self._compose(self.abs, self.rel)                                  # poseInitialization
        if len(self.poses) <= 1:                                                      # buildLocalMap
            raw = surf
        else:
            if len(self.recent) < 20:
                i = len(self.poses) - 1
                self.recent.append(O.transform_cloud(self.frames[i], self.poses[i]))
            elif self.latest != len(self.poses) - 1:
                self.recent.pop(0); self.latest = len(self.poses) - 1
                self.recent.append(O.transform_cloud(self.frames[self.latest], self.poses[self.latest]))
            raw = np.concatenate(self.recent) if self.recent else surf[:0]
        map_ds = O.voxelgrid(raw, 0.4); ds = O.voxelgrid(surf, 0.4)                   # downSampleCloud
        if len(map_ds) >= 10:                                                         # updateTransformationWithCeres
            tree = O.KdTree(_f4(map_ds))
            match_cnt = 8 if len(self.poses) < 2 else self.smc
            if self.mode == 0: rc, pose, _ = O.scan_to_map_ceres(tree, ds, self.abs, match_cnt, self.mni)
            else: rc, pose, _ = O.scan_to_map_gn(tree, ds, self.abs, match_cnt)
            self.abs = pose
            dis = np.linalg.norm(pose[4:] - self.t_kf)
            qi = self.q_kf * np.array([1, -1, -1, -1]) / np.dot(self.q_kf, self.q_kf)
            w = qi[0] * pose[0] - qi[1] * pose[1] - qi[2] * pose[2] - qi[3] * pose[3]
            ang = 2 * np.arccos(w)
            sz = len(self.poses)
            if ((dis > 0.2 or ang > 0.1) and (sz - self.kf_num > 1)) or (sz - self.kf_num > 2) or sz <= 1:
                self.kf = True; self.t_kf = pose[4:].copy(); self.q_kf = pose[:4].copy()
            else:
                self.kf = False
        self.poses.append(self.abs.copy()); self.frames.append(ds)                    # savePoses
        self.rel = self._relative(self.poses[-2], self.abs)                           # computeRelative
        if self.kf: self.kf_num = len(self.poses)
        return dict(abs=self.abs.copy(), rel=self.rel.copy(), kf=self.kf, n_map=len(map_ds), n_ds=len(ds))


@pytest.mark.parametrize("mode", [0, 1])
def test_lidar_odometry_node_sequence(oracle, world_small, mode):
    import liliom_b200 as L
    from liliom_b200 import synth
    ctx = L.Context(variant=0)
    node = L.LidarOdometryNode(ctx, max_num_iter=15, scan_match_cnt=2, if_to_deskew=False, mode=mode)
    ref = OracleLO(oracle, 15, 2, mode)
    T0 = world_small["T"]
    n_kf = 0
    for k in range(7):
        T = np.array(T0); T[4] += 0.12 * k; T[5] += 0.02 * k
        T[:4] = synth.qmul(synth.q_from_axis_angle([0, 0, 1], np.deg2rad(0.6 * k)), T0[:4])
        pts, q = synth.make_horizon_sweep(T, seed=60 + k)
        surf, edge, cut = oracle.extract_horizon(pts, q)
        node.feed(0.1 * k, edge, surf, cut)
        out, ke, ks, kfull = node.run()
        want = ref.run(surf)
        assert out.ran == 1
        if want is None:
            assert out.initialized == 0 and len(ks) == len(surf)          # checkInitialization republishes the clouds
            continue
        assert out.initialized == 1 and out.status == 0
        assert out.n_map == want["n_map"] and out.n_surf_ds == want["n_ds"]
        a = np.array(out.abs_pose); r = np.array(out.rel_pose)
        assert np.linalg.norm(a[4:] - want["abs"][4:]) < 1e-4 and 1 - abs(np.dot(a[:4], want["abs"][:4])) < 1e-9, (k, a, want["abs"])
        assert np.linalg.norm(r[4:] - want["rel"][4:]) < 1e-4
        assert bool(out.kf) == want["kf"], k
        if out.kf:
            n_kf += 1
            assert len(ks) == len(surf) and len(ke) == len(edge) and len(kfull) == len(cut)
        else:
            assert len(ks) == 0
        # the estimate follows the true motion relative to the first frame (first frame defines the world)
    assert n_kf >= 2
    node.close(); ctx.close()


def test_lidar_odometry_deskew_and_sync_gate(oracle, world_small):
    import liliom_b200 as L
    ctx = L.Context(variant=0)
    node = L.LidarOdometryNode(ctx, if_to_deskew=True, mode=1)
    surf, edge, cut = oracle.extract_horizon(world_small["hz"], world_small["q_hz"])
    # stamps more than 0.1 s apart: run() must not fire (L/src/LidarOdometry.cpp:653-660)
    import ctypes
    lib = L._binding.lib()
    lib.liliom_lo_edge(node._h, 0.0, L._binding._ptr(edge), len(edge))
    lib.liliom_lo_surf(node._h, 0.25, L._binding._ptr(surf), len(surf))
    lib.liliom_lo_full(node._h, 0.0, L._binding._ptr(cut), len(cut))
    out, *_ = node.run()
    assert out.ran == 0
    node.close(); ctx.close()
