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
