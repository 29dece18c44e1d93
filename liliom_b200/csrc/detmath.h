// Deterministic fp32 atan / atan2 for the spinning-LiDAR extractor.
//
// The reference calls the float overloads atan(float) / atan2(float,float)
// (R/src/Preprocessing.cpp:285-286,315,349 under `using namespace std`), i.e. glibc's atanf /
// atan2f.  CUDA's atanf/atan2f differ from glibc's in the last ulp on a few percent of
// inputs, which would leak into relTime -> intensity -> de-skew -> curvature -> labels.
// These routines follow the classic fdlibm single-precision algorithm (argument reduction
// to [0, 7/16] + odd/even polynomial, hi/lo table constants) using only IEEE-754
// add/mul/div — no FMA contraction (compile with --fmad=false / -ffp-contract=off) — and are
// verified bit-for-bit against this image's glibc 2.39 atanf (all 2^32 inputs) and atan2f
// (10^9 random pairs + special cases) by tests/test_detmath.py.
#pragma once
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define LILI_HD __host__ __device__ __forceinline__
#else
#define LILI_HD static inline
#endif

namespace lili {

LILI_HD int32_t f2i(float f) {
#if defined(__CUDA_ARCH__)
    return __float_as_int(f);
#else
    int32_t i; memcpy(&i, &f, 4); return i;
#endif
}
LILI_HD float i2f(int32_t i) {
#if defined(__CUDA_ARCH__)
    return __int_as_float(i);
#else
    float f; memcpy(&f, &i, 4); return f;
#endif
}

LILI_HD float det_atanf(float x) {
    const float hi0 = i2f(0x3eed6338), hi1 = i2f(0x3f490fda), hi2 = i2f(0x3f7b985e), hi3 = i2f(0x3fc90fda);
    const float lo0 = i2f(0x31ac3769), lo1 = i2f(0x33222168), lo2 = i2f(0x33140fb4), lo3 = i2f(0x33a22168);
    const float a0 = i2f(0x3eaaaaab), a1 = i2f((int32_t)0xbe4ccccd), a2 = i2f(0x3e124925), a3 = i2f((int32_t)0xbde38e38),
                a4 = i2f(0x3dba2e6e), a5 = i2f((int32_t)0xbd9d8795), a6 = i2f(0x3d886b35), a7 = i2f((int32_t)0xbd6ef16b),
                a8 = i2f(0x3d4bda59), a9 = i2f((int32_t)0xbd15a221), a10 = i2f(0x3c8569d7);
    const int32_t hx = f2i(x);
    const int32_t ix = hx & 0x7fffffff;
    int id;
    float hi = 0.f, lo = 0.f;
    if (ix >= 0x4c000000) {            // |x| >= 2^25
        if (ix > 0x7f800000) return x + x;   // NaN
        return hx > 0 ? hi3 + lo3 : -hi3 - lo3;
    }
    if (ix < 0x3ee00000) {             // |x| < 0.4375
        if (ix < 0x31000000) return x; // |x| < 2^-29
        id = -1;
    } else {
        x = i2f(ix);                   // fabsf
        if (ix < 0x3f980000) {         // |x| < 1.1875
            if (ix < 0x3f300000) { id = 0; hi = hi0; lo = lo0; x = (2.0f * x - 1.0f) / (2.0f + x); }
            else { id = 1; hi = hi1; lo = lo1; x = (x - 1.0f) / (x + 1.0f); }
        } else {
            if (ix < 0x401c0000) { id = 2; hi = hi2; lo = lo2; x = (x - 1.5f) / (1.0f + 1.5f * x); }
            else { id = 3; hi = hi3; lo = lo3; x = -1.0f / x; }
        }
    }
    float z = x * x;
    float w = z * z;
    float s1 = z * (a0 + w * (a2 + w * (a4 + w * (a6 + w * (a8 + w * a10)))));
    float s2 = w * (a1 + w * (a3 + w * (a5 + w * (a7 + w * a9))));
    if (id < 0) return x - x * (s1 + s2);
    z = hi - ((x * (s1 + s2) - lo) - x);
    return hx < 0 ? -z : z;
}

LILI_HD float det_atan2f(float y, float x) {
    const float tiny = 1.0e-30f;
    const float pi_o_4 = i2f(0x3f490fdb), pi_o_2 = i2f(0x3fc90fdb), pi = i2f(0x40490fdb), pi_lo = i2f((int32_t)0xb3bbbd2e);
    const int32_t hx = f2i(x), hy = f2i(y);
    const int32_t ix = hx & 0x7fffffff, iy = hy & 0x7fffffff;
    if (ix > 0x7f800000 || iy > 0x7f800000) return x + y;
    if (hx == 0x3f800000) return det_atanf(y);
    const int m = ((hy >> 31) & 1) | ((hx >> 30) & 2);
    if (iy == 0) {
        switch (m) {
            case 0:
            case 1: return y;
            case 2: return pi + tiny;
            default: return -pi - tiny;
        }
    }
    if (ix == 0) return (hy < 0) ? -pi_o_2 - tiny : pi_o_2 + tiny;
    if (ix == 0x7f800000) {
        if (iy == 0x7f800000) {
            switch (m) {
                case 0: return pi_o_4 + tiny;
                case 1: return -pi_o_4 - tiny;
                case 2: return 3.0f * pi_o_4 + tiny;
                default: return -3.0f * pi_o_4 - tiny;
            }
        } else {
            switch (m) {
                case 0: return 0.0f;
                case 1: return -0.0f;
                case 2: return pi + tiny;
                default: return -pi - tiny;
            }
        }
    }
    if (iy == 0x7f800000) return (hy < 0) ? -pi_o_2 - tiny : pi_o_2 + tiny;
    const int32_t k = (iy - ix) >> 23;
    float z;
    if (k > 60) z = pi_o_2 + 0.5f * pi_lo;
    else if (hx < 0 && k < -60) z = 0.0f;
    else z = det_atanf(i2f(f2i(y / x) & 0x7fffffff));
    switch (m) {
        case 0: return z;
        case 1: return i2f(f2i(z) ^ (int32_t)0x80000000);
        case 2: return pi - (z - pi_lo);
        default: return (z - pi_lo) - pi;
    }
}

}  // namespace lili
