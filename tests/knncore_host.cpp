// TEST INFRASTRUCTURE: host build of the pruned search of liliom_b200/csrc/knn_core.cuh (query_cell, cell_bound, row_cells, consider,
// thread_knn5 — the one-thread-per-query shape of the roofline kernel) so that the CPU test tier can run the SAME SOURCE the kernels
// compile against an exhaustive search: the claim under test is that thresholds, cell lower bounds, run trimming and the
// per-thread run list never change the five nearest keys.  Round-to-nearest intrinsics become plain operators (-ffp-contract=off),
// read-only loads become plain loads; the warp primitives the other shapes use are only declared (never instantiated here).
#include <algorithm>
#include <cmath>
#include <cfloat>
#include <cstring>
using std::min;
using std::max;
using std::isfinite;
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline double __dadd_rn(double a, double b) { return a + b; }
static inline double __dsub_rn(double a, double b) { return a - b; }
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __fsub_rn(float a, float b) { return a - b; }
static inline unsigned __float_as_uint(float f) { unsigned u; memcpy(&u, &f, 4); return u; }
static inline int __float_as_int(float f) { int u; memcpy(&u, &f, 4); return u; }
static inline float __uint_as_float(unsigned u) { float f; memcpy(&f, &u, 4); return f; }
static inline float __int_as_float(int u) { float f; memcpy(&f, &u, 4); return f; }
template <class T> static inline T __ldg(const T* p) { return *p; }
template <class T> T __shfl_xor_sync(unsigned, T, int);          // declared only: the multi-lane shapes are not instantiated on the host
template <class T> T __shfl_sync(unsigned, T, int);
long long clock64();
static inline size_t __cvta_generic_to_shared(const void* p) { return (size_t)p; }
#include <cuda_runtime.h>
#ifndef __noinline__
#define __noinline__
#endif
#include "../liliom_b200/csrc/knn_core.cuh"

using namespace lili;

extern "C" {
float kc_gate_tau(double max_sqd) { return knn_gate_tau(max_sqd); }
int kc_owner_of(float x, float y, float z, int nranks, float inv_block) { return owner_of(x, y, z, nranks, inv_block); }

// map_sorted: m x {x, y, z, index bits} in cell order; returns the number of map points examined
unsigned long long kc_thread_knn5(float sx, float sy, float sz, const float* map_sorted, const int* cell_start, float inv_cell,
                                  const int org[3], const int dim[3], float tau0, unsigned long long out5[5]) {
    GridDesc g;
    g.inv_cell = inv_cell;
    for (int k = 0; k < 3; ++k) { g.org[k] = org[k]; g.dim[k] = dim[k]; }
    g.ncells = dim[0] * dim[1] * dim[2];
    Top5 top;
    top5_init(top);
    unsigned long long cand = 0;
    int4 runs[kRunCap];
    thread_knn5<8, 4>(sx, sy, sz, reinterpret_cast<const float4*>(map_sorted), cell_start, g, tau0, runs, 1, top, cand);
    out5[0] = top.k0; out5[1] = top.k1; out5[2] = top.k2; out5[3] = top.k3; out5[4] = top.k4;
    return cand;
}
}
