// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle_math.h).  PARITY UNPINNED.
// BackendFusion correspondence searches restated (kernel-reuse rows a16/a18 of SURVEY.md §8):
//   point-to-line : LiLi-OM/src/BackendFusion.cpp:1531-1599, LiLi-OM-ROT/src/BackendFusion.cpp:1394-1462
//   point-to-plane: LiLi-OM/src/BackendFusion.cpp:1601-1681, LiLi-OM-ROT/src/BackendFusion.cpp:1464-1520
// The callers (sliding-window optimiser) are out of scope.
#include "oracle_api.h"
#include "oracle_math.h"
#include <cstring>

namespace orc { void knn5_one(const void* tree, const float q[3], int idx[5], float sqd[5]); }
using namespace orc;

static inline void xform(const double pose7[7], const float p[3], float out[3]) {
    Quat q{pose7[0], pose7[1], pose7[2], pose7[3]};
    V3 o = qrot(q, V3{p[0], p[1], p[2]}) + V3{pose7[4], pose7[5], pose7[6]};
    out[0] = (float)o.x; out[1] = (float)o.y; out[2] = (float)o.z;
}

extern "C" int orc_correspond_edge(const void* tree, const float* map_xyzw, int m, const float* feats_xyzw, int n,
                                   const double pose7[7], int variant, unsigned char* valid, float* pa, float* pb) {
    (void)m;
    int cnt = 0;
    for (int i = 0; i < n; ++i) {
        valid[i] = 0;
        for (int k = 0; k < 3; ++k) { pa[3 * (size_t)i + k] = 0; pb[3 * (size_t)i + k] = 0; }
        float sel[3];
        xform(pose7, feats_xyzw + 4 * (size_t)i, sel);
        int idx[5]; float sqd[5];
        knn5_one(tree, sel, idx, sqd);
        if (idx[4] < 0 || !(sqd[4] < 1.0)) continue;                               // L:1543
        V3 c{0, 0, 0};
        V3 pt[5];
        for (int j = 0; j < 5; ++j) {
            const float* mp = map_xyzw + 4 * (size_t)idx[j];
            pt[j] = V3{mp[0], mp[1], mp[2]};
            c = c + pt[j];
        }
        c = V3{c.x / 5.0, c.y / 5.0, c.z / 5.0};                                    // L:1555
        double A[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
        for (int j = 0; j < 5; ++j) {                                               // L:1560-1564
            V3 z = pt[j] - c;
            const double zm[3] = {z.x, z.y, z.z};
            for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) A[a][b] += zm[a] * zm[b];
        }
        double ev[3], evec[3][3];
        eigen_sym3(A, ev, evec);                                                    // L:1568
        V3 u{evec[0][2], evec[1][2], evec[2][2]};
        if (!(ev[2] > 3 * ev[1])) continue;                                         // L:1575
        V3 A_ = c + 0.1 * u, B_ = c - 0.1 * u;                                      // L:1579-1580
        if (variant == 1) {                                                         // R:1435-1439
            V3 lp{sel[0], sel[1], sel[2]};
            V3 nu = cross(lp - A_, lp - B_);
            V3 de = A_ - B_;
            double dist = norm(nu) / norm(de);
            if (!(dist < 0.1)) continue;
        }
        pa[3 * (size_t)i] = (float)A_.x; pa[3 * (size_t)i + 1] = (float)A_.y; pa[3 * (size_t)i + 2] = (float)A_.z;
        pb[3 * (size_t)i] = (float)B_.x; pb[3 * (size_t)i + 1] = (float)B_.y; pb[3 * (size_t)i + 2] = (float)B_.z;
        valid[i] = 1;
        ++cnt;
    }
    return cnt;
}

extern "C" int orc_correspond_surf_backend(const void* tree, const float* map_xyzw, int m, const float* feats_xyzw, int n,
                                           const double pose7[7], double kd_max_radius, double surf_dist_thres, double w_gate,
                                           double lidar_const, const float* map_refl, const float* feat_refl, double reflect_thres,
                                           unsigned char* valid, float* plane, double* score) {
    (void)m;
    int cnt = 0;
    for (int i = 0; i < n; ++i) {
        valid[i] = 0; score[i] = 0;
        for (int k = 0; k < 4; ++k) plane[4 * (size_t)i + k] = 0;
        float sel[3];
        xform(pose7, feats_xyzw + 4 * (size_t)i, sel);
        int idx[5]; float sqd[5];
        knn5_one(tree, sel, idx, sqd);
        if (idx[4] < 0 || !((double)sqd[4] < kd_max_radius)) continue;             // L:1615 / R:1476
        double A[15], B[5];
        double sum_w = 0;
        for (int j = 0; j < 5; ++j) B[j] = -1.0;
        if (map_refl) {                                                             // L:1617-1638 (Horizon only)
            double vec_w[5];
            for (int j = 0; j < 5; ++j) {
                double tmp_w = std::fabs(feat_refl[i] - map_refl[idx[j]]);
                sum_w += tmp_w;
                vec_w[j] = 1.0 / tmp_w;      // may be +inf, as in the reference (L:1624-1625)
            }
            for (int j = 0; j < 5; ++j) vec_w[j] /= sum_w;
            if (sum_w > reflect_thres) continue;
            for (int j = 0; j < 5; ++j) {
                const float* mp = map_xyzw + 4 * (size_t)idx[j];
                A[3 * j] = vec_w[j] * mp[0]; A[3 * j + 1] = vec_w[j] * mp[1]; A[3 * j + 2] = vec_w[j] * mp[2];
                B[j] *= vec_w[j];
            }
        } else {
            for (int j = 0; j < 5; ++j) {
                const float* mp = map_xyzw + 4 * (size_t)idx[j];
                A[3 * j] = mp[0]; A[3 * j + 1] = mp[1]; A[3 * j + 2] = mp[2];
            }
        }
        double nv[3];
        colpiv_qr_solve_nx3(5, A, B, nv);
        double nn = std::sqrt(nv[0] * nv[0] + nv[1] * nv[1] + nv[2] * nv[2]);
        double normInverse = 1 / nn;
        if (nn * nn > 0) { nv[0] /= nn; nv[1] /= nn; nv[2] /= nn; }
        bool planeValid = true;
        for (int j = 0; j < 5; ++j) {
            const float* mp = map_xyzw + 4 * (size_t)idx[j];
            if (std::fabs(nv[0] * mp[0] + nv[1] * mp[1] + nv[2] * mp[2] + normInverse) > surf_dist_thres) { planeValid = false; break; }
        }
        if (!planeValid) continue;
        float pd = nv[0] * sel[0] + nv[1] * sel[1] + nv[2] * sel[2] + normInverse;
        float weight = 1 - 0.9 * std::fabs(pd) / std::sqrt(std::sqrt(sel[0] * sel[0] + sel[1] * sel[1] + sel[2] * sel[2]));
        if (!(weight > w_gate)) continue;
        plane[4 * (size_t)i] = weight * nv[0];
        plane[4 * (size_t)i + 1] = weight * nv[1];
        plane[4 * (size_t)i + 2] = weight * nv[2];
        plane[4 * (size_t)i + 3] = weight * normInverse;
        score[i] = map_refl ? lidar_const * (weight + std::exp(-sum_w)) : lidar_const * weight;   // L:1676 / R:1515
        valid[i] = 1;
        ++cnt;
    }
    return cnt;
}

// ---------------------------------------------------------------------------------------
// the LiDAR residual blocks one window keyframe contributes to the backend problem,
// LiLi-OM/src/BackendFusion.cpp:919-979 — LidarEdgeFactor (LidarKeyframeFactor.h:12-62) and LidarPlaneNormFactor
// (:65-108), AutoDiffCostFunction<.,1,3,4> on the parameter blocks (t, q) with QuaternionParameterization on q and
// ceres::CauchyLoss(1.0) (:845).  What Ceres feeds its linear solver for such a block is, per residual, the row
// sqrt(rho') * [dr/dt , dr/dq * PlusJacobian(q)] and the residual sqrt(rho') * r (Corrector, rho'' <= 0 branch); the
// functions below reduce those rows to the keyframe's 6x6 normal-equation block.
// Tangent order follows the parameter-block order: [t(3), rot(3)].  out29 = 21 (upper triangle, row-major) + 6 + cost + count.
// ---------------------------------------------------------------------------------------
namespace {

// d(q*v)/dq_k for Eigen's q*v expression (uv = qv x v; uv += uv; v + w uv + qv x uv), then the from-knowledge
// ceres::QuaternionParameterization::ComputeJacobian (4x3) — same construction as plane_row() in oracle_s2m.cpp.
inline void rot_tangent_rows(const double x[4], V3 v, V3 g, double Jrot[3]) {
    const double w = x[0];
    V3 qv{x[1], x[2], x[3]};
    V3 uv = cross(qv, v); uv = uv + uv;
    double jq[4];
    jq[0] = dot(g, uv);
    const V3 e[3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
    for (int k = 0; k < 3; ++k) {
        V3 duv = 2.0 * cross(e[k], v);
        V3 df = w * duv + cross(e[k], uv) + cross(qv, duv);
        jq[k + 1] = dot(g, df);
    }
    const double P[4][3] = {{-x[1], -x[2], -x[3]}, {x[0], x[3], -x[2]}, {-x[3], x[0], x[1]}, {x[2], -x[1], x[0]}};
    for (int c = 0; c < 3; ++c) Jrot[c] = jq[0] * P[0][c] + jq[1] * P[1][c] + jq[2] * P[2][c] + jq[3] * P[3][c];
}

// from-knowledge ceres::CauchyLoss(b)::Evaluate: rho = b^2 log(1 + s/b^2), rho' = 1/(1 + s/b^2), rho'' < 0, then the
// Corrector's rho'' <= 0 branch: row and residual scaled by sqrt(rho').
inline void add_row(double J[6], double r, double cauchy_b, double out29[29]) {
    const double s = r * r, b2 = cauchy_b * cauchy_b, c2 = 1.0 / b2;
    const double sum = 1.0 + s * c2;
    const double inv = 1.0 / sum;
    const double rho0 = b2 * std::log(sum);
    const double rho1 = std::max(std::numeric_limits<double>::min(), inv);
    const double sr = std::sqrt(rho1);
    for (int c = 0; c < 6; ++c) J[c] *= sr;
    r *= sr;
    int k = 0;
    for (int a = 0; a < 6; ++a)
        for (int b = a; b < 6; ++b) out29[k++] += J[a] * J[b];
    for (int a = 0; a < 6; ++a) out29[21 + a] += J[a] * r;
    out29[27] += 0.5 * rho0;
    out29[28] += 1.0;
}

}  // namespace

extern "C" void orc_backend_edge_block(const float* feats_xyzw, int n, const unsigned char* valid, const float* pa, const float* pb,
                                       double s_weight, const double pose7_body[7], double cauchy_b, double out29[29]) {
    for (int k = 0; k < 29; ++k) out29[k] = 0;
    Quat q{pose7_body[0], pose7_body[1], pose7_body[2], pose7_body[3]};
    V3 t{pose7_body[4], pose7_body[5], pose7_body[6]};
    for (int i = 0; i < n; ++i) {
        if (!valid[i]) continue;
        const float* f = feats_xyzw + 4 * (size_t)i;
        V3 cp{f[0], f[1], f[2]};
        V3 a{pa[3 * (size_t)i], pa[3 * (size_t)i + 1], pa[3 * (size_t)i + 2]};
        V3 b{pb[3 * (size_t)i], pb[3 * (size_t)i + 1], pb[3 * (size_t)i + 2]};
        V3 lp = qrot(q, cp) + t;                                 // LidarKeyframeFactor.h:38 (the extrinsics are NOT applied)
        V3 nu = cross(lp - a, lp - b);                           // :40
        V3 de = a - b;                                           // :41
        const double nn = norm(nu), dn = norm(de);
        double r = nn / dn;                                      // :43
        r *= s_weight;                                           // :44
        // dr/dlp = s ((a-b) x nu^) / |a-b|; a zero cross product has no derivative (autodiff yields NaN): row dropped to 0
        V3 g{0, 0, 0};
        if (nn > 0) g = (s_weight / (dn * nn)) * cross(de, nu);
        double J[6];
        J[0] = g.x; J[1] = g.y; J[2] = g.z;
        rot_tangent_rows(pose7_body, cp, g, J + 3);
        add_row(J, r, cauchy_b, out29);
    }
}

extern "C" void orc_backend_surf_block(const float* feats_xyzw, int n, const unsigned char* valid, const float* plane, const double* score,
                                       const double pose7_body[7], const double q_lb[4], const double t_lb[3], double cauchy_b,
                                       double out29[29]) {
    for (int k = 0; k < 29; ++k) out29[k] = 0;
    Quat q{pose7_body[0], pose7_body[1], pose7_body[2], pose7_body[3]};
    V3 t{pose7_body[4], pose7_body[5], pose7_body[6]};
    Quat qlbi = qinv(Quat{q_lb[0], q_lb[1], q_lb[2], q_lb[3]});
    V3 tlb{t_lb[0], t_lb[1], t_lb[2]};
    for (int i = 0; i < n; ++i) {
        if (!valid[i]) continue;
        const float* f = feats_xyzw + 4 * (size_t)i;
        const float* pl = plane + 4 * (size_t)i;
        V3 cp{f[0], f[1], f[2]};
        V3 pbody = qrot(qlbi, cp - tlb);                         // LidarKeyframeFactor.h:87
        V3 pw = qrot(q, pbody) + t;                              // :88
        V3 nrm{pl[0], pl[1], pl[2]};
        const double sc = score[i];
        const double r = sc * (dot(nrm, pw) + (double)pl[3]);    // :91
        V3 g = sc * nrm;
        double J[6];
        J[0] = g.x; J[1] = g.y; J[2] = g.z;
        rot_tangent_rows(pose7_body, pbody, g, J + 3);
        add_row(J, r, cauchy_b, out29);
    }
}

// ---------------------------------------------------------------------------------------
// FormatConvert's livoxLidarHandler, LiLi-OM/src/FormatConvert.cpp:11-24, on the in-memory
// livox_ros_driver::CustomPoint {uint32 offset_time; float x,y,z; uint8 reflectivity, tag, line} (stride bytes per
// point: 20 for the C++ message struct, 19 for the serialised wire layout).  Output: pcl::PointXYZINormal as
// push_back'ed there (data[3] = 1, normals/curvature pads 0).
// ---------------------------------------------------------------------------------------
extern "C" void orc_convert_livox(const unsigned char* custom_pts, int n, int stride, orc_pt48* out) {
    if (n <= 0) return;
    unsigned int time_end;
    std::memcpy(&time_end, custom_pts + (size_t)(n - 1) * stride, 4);       // :13 points.back().offset_time
    for (int i = 0; i < n; ++i) {
        const unsigned char* p = custom_pts + (size_t)i * stride;
        unsigned int off; float xyz[3];
        std::memcpy(&off, p, 4); std::memcpy(xyz, p + 4, 12);
        const unsigned char refl = p[16], line = p[18];
        orc_pt48 pt;
        std::memset(&pt, 0, sizeof(pt));
        pt.x = xyz[0]; pt.y = xyz[1]; pt.z = xyz[2]; pt.w = 1.0f;
        float s = float(off / (float)time_end);                              // :19 (uint32 -> float, fp32 division)
        pt.intensity = line + s * 0.1;                                       // :20 (double arithmetic, narrowed)
        pt.curvature = 0.1 * refl;                                           // :21
        out[i] = pt;
    }
}
