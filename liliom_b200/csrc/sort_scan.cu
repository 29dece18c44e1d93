// Library plumbing: stable radix sort and exclusive scan through CUB (part of the CUDA
// toolkit).  These are the only non-hand-written device routines in the library; they are
// counted separately (liliom_counters.lib_launches) and never appear in gpu_launches.
#include "ctx.cuh"
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

namespace lili {

int sort_pairs_u32(liliom_ctx* c, const uint32_t* kin, uint32_t* kout, const int* vin, int* vout, int n, int end_bit) {
    if (n <= 0) return LILIOM_OK;
    size_t need = 0;
    LILI_CUDA(c, cub::DeviceRadixSort::SortPairs(nullptr, need, kin, kout, vin, vout, n, 0, end_bit, c->stream));
    LILI_CUDA(c, c->cub_tmp.ensure(need));
    LILI_CUDA(c, cub::DeviceRadixSort::SortPairs(c->cub_tmp.p, need, kin, kout, vin, vout, n, 0, end_bit, c->stream));
    c->cnt.lib_launches++;
    return LILIOM_OK;
}

int sort_pairs_u64(liliom_ctx* c, const unsigned long long* kin, unsigned long long* kout, const int* vin, int* vout, int n, int end_bit) {
    if (n <= 0) return LILIOM_OK;
    size_t need = 0;
    LILI_CUDA(c, cub::DeviceRadixSort::SortPairs(nullptr, need, kin, kout, vin, vout, n, 0, end_bit, c->stream));
    LILI_CUDA(c, c->cub_tmp.ensure(need));
    LILI_CUDA(c, cub::DeviceRadixSort::SortPairs(c->cub_tmp.p, need, kin, kout, vin, vout, n, 0, end_bit, c->stream));
    c->cnt.lib_launches++;
    return LILIOM_OK;
}

// out[0..n] = exclusive prefix sums of in[0..n) with the total at out[n] (in[n] is read as 0:
// callers keep one spare zeroed element).
int exclusive_scan_i32(liliom_ctx* c, const int* in, int* out, int n) {
    size_t need = 0;
    LILI_CUDA(c, cub::DeviceScan::ExclusiveSum(nullptr, need, in, out, n + 1, c->stream));
    LILI_CUDA(c, c->cub_tmp.ensure(need));
    LILI_CUDA(c, cub::DeviceScan::ExclusiveSum(c->cub_tmp.p, need, in, out, n + 1, c->stream));
    c->cnt.lib_launches++;
    return LILIOM_OK;
}

// in-place inclusive running maximum of data[0..n) (used by grid_build: upper bounds of the cell runs)
int inclusive_max_scan_i32(liliom_ctx* c, int* data, int n) {
    if (n <= 0) return LILIOM_OK;
    size_t need = 0;
    LILI_CUDA(c, cub::DeviceScan::InclusiveScan(nullptr, need, data, data, cub::Max(), n, c->stream));
    LILI_CUDA(c, c->cub_tmp.ensure(need));
    LILI_CUDA(c, cub::DeviceScan::InclusiveScan(c->cub_tmp.p, need, data, data, cub::Max(), n, c->stream));
    c->cnt.lib_launches++;
    return LILIOM_OK;
}

}  // namespace lili
