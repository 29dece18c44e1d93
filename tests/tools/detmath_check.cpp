// Host-side check of liliom_b200/csrc/detmath.h against glibc (test tool; built by tests/test_detmath.py).
// usage: detmath_check <atanf_stride> <atan2_pairs>
#include "../../liliom_b200/csrc/detmath.h"
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <omp.h>
static inline bool same(float a, float b) {
    if (std::isnan(a) && std::isnan(b)) return true;
    return lili::f2i(a) == lili::f2i(b);
}
int main(int argc, char** argv) {
    long stride = argc > 1 ? atol(argv[1]) : 1;
    long pairs = argc > 2 ? atol(argv[2]) : 100000000L;
    long bad1 = 0, bad2 = 0;
#pragma omp parallel for reduction(+ : bad1) schedule(static)
    for (long long u = 0; u < (1LL << 32); u += stride) {
        float x = lili::i2f((int32_t)(uint32_t)u);
        if (!same(lili::det_atanf(x), atanf(x))) { if (bad1 < 5) fprintf(stderr, "atanf mismatch x=%a det=%a libm=%a\n", x, lili::det_atanf(x), atanf(x)); ++bad1; }
    }
#pragma omp parallel reduction(+ : bad2)
    {
        uint64_t s = 0x9E3779B97F4A7C15ull * (uint64_t)(omp_get_thread_num() + 1);
        long per = pairs / omp_get_num_threads();
        for (long i = 0; i < per; ++i) {
            s ^= s << 13; s ^= s >> 7; s ^= s << 17;
            uint32_t a = (uint32_t)s, b = (uint32_t)(s >> 32);
            float y, x;
            if (i & 1) { y = lili::i2f((int32_t)a); x = lili::i2f((int32_t)b); }
            else {  // lidar-like magnitudes
                y = ((int32_t)a) * (200.0f / 2147483648.0f); x = ((int32_t)b) * (200.0f / 2147483648.0f);
            }
            if (!same(lili::det_atan2f(y, x), atan2f(y, x))) { if (bad2 < 5) fprintf(stderr, "atan2f mismatch y=%a x=%a det=%a libm=%a\n", y, x, lili::det_atan2f(y, x), atan2f(y, x)); ++bad2; }
        }
    }
    const float sp[] = {0.f, -0.f, 1.f, -1.f, INFINITY, -INFINITY, NAN, 1e-40f, -1e-40f, 3e38f, -3e38f, 0.5f, 2.f};
    for (float y : sp) for (float x : sp) if (!same(lili::det_atan2f(y, x), atan2f(y, x))) { fprintf(stderr, "atan2f special mismatch y=%a x=%a\n", y, x); ++bad2; }
    printf("atanf_mismatch=%ld atan2f_mismatch=%ld\n", bad1, bad2);
    return (bad1 || bad2) ? 1 : 0;
}
