// TEST INFRASTRUCTURE: host build of the device scalar routines (liliom_b200/csrc/dev_math.cuh) so that the CPU test tier can
// exercise the SAME SOURCE the kernels compile — plane fit (fast path + rank-revealing QR fallback), 6x6 LDL^T, Ceres' Plus,
// the 3x3 symmetric eigensolver — against NumPy and against the oracle's independent copies.  The round-to-nearest intrinsics
// become plain operators (this file is compiled with -ffp-contract=off, so nothing is fused either).
#include <cmath>
#include <cfloat>
#include <cstring>
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline double __dadd_rn(double a, double b) { return a + b; }
static inline double __dsub_rn(double a, double b) { return a - b; }
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __fsub_rn(float a, float b) { return a - b; }
using std::isfinite;
#include <cuda_runtime.h>
#ifndef __noinline__
#define __noinline__
#endif
#include "../liliom_b200/csrc/dev_math.cuh"

using namespace lili;

extern "C" {
int dm_plane_fit5(const float* m20, double nv[3]) {          // returns 1 when the fast path produced the answer, 0 for the QR fallback
    float4 m[5];
    for (int j = 0; j < 5; ++j) m[j] = make_float4(m20[4 * j], m20[4 * j + 1], m20[4 * j + 2], m20[4 * j + 3]);
    if (plane_fit5_fast(m, nv)) return 1;
    plane_fit5_qr(m, nv);
    return 0;
}
int dm_solve6(const double s21[21], const double rhs[6], double x[6]) { return solve6_ldlt(s21, rhs, x) ? 1 : 0; }
int dm_gn_safe_step(const double s21[21], const double rhs[6], double d[6]) { return gn_safe_step(s21, rhs, d) ? 1 : 0; }
void dm_pose_plus(const double x[7], const double d[6], double out[7]) { pose_plus(x, d, out); }
void dm_eigen_sym3(const double a[6], double ev[3], double evec[9]) {     // a = {a00, a10, a20, a11, a21, a22}
    double v[3][3];
    eigen_sym3(a[0], a[1], a[2], a[3], a[4], a[5], ev, v);
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) evec[3 * r + c] = v[r][c];
}
void dm_colpiv_qr(const double A15[15], const double b5[5], double x[3]) {
    double A[5][3], b[5];
    for (int j = 0; j < 5; ++j) { for (int k = 0; k < 3; ++k) A[j][k] = A15[3 * j + k]; b[j] = b5[j]; }
    colpiv_qr_solve_5x3(A, b, x);
}
void dm_qrot(const double q[4], const double v[3], double out[3]) {
    D3 r = qrot_x(Q4{q[0], q[1], q[2], q[3]}, D3{v[0], v[1], v[2]});
    out[0] = r.x; out[1] = r.y; out[2] = r.z;
}
}
