// Library primitives (CUB, shipped with CUDA 12.9) used off the critical kernels:
// prefix sums, one u64 radix sort and one stream compaction.  Kept in their own
// translation unit because the CUB headers dominate compile time.
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <cub/device/device_select.cuh>
#include <cub/iterator/transform_input_iterator.cuh>

#include "nsmh_internal.cuh"

// cub::TransformInputIterator is deprecated in favour of thrust::transform_iterator in CUDA 12.9 but
// still the documented input adaptor of this CUB version; keep the build output clean.
#pragma GCC diagnostic ignored "-Wdeprecated-declarations"

namespace nsmh {

struct U32ToU64 {
    __host__ __device__ __forceinline__ uint64_t operator()(const uint32_t &v) const { return v; }
};
struct Low32 {
    __host__ __device__ __forceinline__ uint32_t operator()(const uint64_t &v) const {
        return (uint32_t)v;
    }
};

cudaError_t cub_exclusive_sum_u32(void *tmp, size_t &tmp_bytes, const uint32_t *in, uint32_t *out,
                                  size_t n, cudaStream_t s) {
    return cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, in, out, n, s);
}

cudaError_t cub_exclusive_sum_u32_to_u64(void *tmp, size_t &tmp_bytes, const uint32_t *in,
                                         uint64_t *out, size_t n, cudaStream_t s) {
    cub::TransformInputIterator<uint64_t, U32ToU64, const uint32_t *> it(in, U32ToU64());
    return cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, it, out, n, s);
}

cudaError_t cub_sort_keys_u64(void *tmp, size_t &tmp_bytes, uint64_t *keys, uint64_t *alt,
                              size_t n, int begin_bit, int end_bit, bool &result_in_alt,
                              cudaStream_t s) {
    cub::DoubleBuffer<uint64_t> db(keys, alt);
    cudaError_t e = cub::DeviceRadixSort::SortKeys(tmp, tmp_bytes, db, n, begin_bit, end_bit, s);
    result_in_alt = (db.Current() == alt);
    return e;
}

cudaError_t cub_select_low32_flagged(void *tmp, size_t &tmp_bytes, const uint64_t *in,
                                     const uint8_t *flags, uint32_t *out, uint64_t *num_selected,
                                     size_t n, cudaStream_t s) {
    cub::TransformInputIterator<uint32_t, Low32, const uint64_t *> it(in, Low32());
    return cub::DeviceSelect::Flagged(tmp, tmp_bytes, it, flags, out, num_selected, n, s);
}

} // namespace nsmh
