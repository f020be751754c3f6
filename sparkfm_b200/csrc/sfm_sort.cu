// sfm_sort.cu -- CUB-backed helpers (library calls, kept in their own translation unit):
//   * stable radix sort of the batch's (feature id, {row, x}) entries -- the transposition the
//     deterministic reduce-by-feature needs.  It is overhead, not counted in
//     the algorithmic bytes.  sort_pairs / sort_pairs32 dispatch to the wide-digit sort of
//     sfm_radix.cu; the library sort stays as the SFM_SORT=cub fallback and as the test oracle.
//   * ascending sort of float scores (AUC).
// (The samplers' stream compaction and the exclusive scans were library calls here in round 1;
// they are own kernels now: sfm_scan.cu.)
#include <cub/cub.cuh>

#include "sfm_common.h"

namespace sfm {

typedef unsigned long long pay_t;  // uint2 payload moved as one 8-byte word

size_t sort_pairs_temp_bytes(int64_t n, int end_bit) {
    size_t bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, bytes, (const uint32_t*)nullptr, (uint32_t*)nullptr,
                                    (const pay_t*)nullptr, (pay_t*)nullptr, n, 0, end_bit);
    const size_t own = radix_temp_bytes(n, end_bit, 8);
    return bytes > own ? bytes : own;
}

cudaError_t sort_pairs(void* tmp, size_t tmp_bytes, const uint32_t* keys_in, uint32_t* keys_out,
                       const uint2* pay_in, uint2* pay_out, int64_t n, int end_bit,
                       cudaStream_t st, int64_t* launches) {
    if (radix_usable(n, end_bit))
        return radix_sort_pairs64(tmp, tmp_bytes, keys_in, keys_out, pay_in, pay_out, n, end_bit, st,
                                  launches);
    // onesweep: 1 histogram + 1 scan + ceil(end_bit/8) passes
    *launches += 2 + (end_bit + 7) / 8;
    return cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, keys_in, keys_out,
                                           reinterpret_cast<const pay_t*>(pay_in),
                                           reinterpret_cast<pay_t*>(pay_out), n, 0, end_bit, st);
}

size_t sort_pairs32_temp_bytes(int64_t n, int end_bit) {
    size_t bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, bytes, (const uint32_t*)nullptr, (uint32_t*)nullptr,
                                    (const uint32_t*)nullptr, (uint32_t*)nullptr, n, 0, end_bit);
    const size_t own = radix_temp_bytes(n, end_bit, 4);
    return bytes > own ? bytes : own;
}

cudaError_t sort_pairs32(void* tmp, size_t tmp_bytes, const uint32_t* keys_in, uint32_t* keys_out,
                         const uint32_t* val_in, uint32_t* val_out, int64_t n, int end_bit,
                         cudaStream_t st, int64_t* launches, int implicit_div) {
    if (radix_usable(n, end_bit))
        return radix_sort_pairs32(tmp, tmp_bytes, keys_in, keys_out, val_in, val_out, n, end_bit,
                                  implicit_div, st, launches);
    if (implicit_div || !val_in) return cudaErrorInvalidValue;   // the library sort needs real values
    *launches += 2 + (end_bit + 7) / 8;
    return cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, keys_in, keys_out, val_in, val_out, n, 0,
                                           end_bit, st);
}

size_t sort_f32_u32_temp_bytes(int64_t n) {
    size_t bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, bytes, (const float*)nullptr, (float*)nullptr,
                                    (const uint32_t*)nullptr, (uint32_t*)nullptr, n);
    return bytes;
}

// ascending sort of float scores with their row numbers (AUC)
cudaError_t sort_f32_u32(void* tmp, size_t tmp_bytes, const float* keys_in, float* keys_out,
                         const uint32_t* val_in, uint32_t* val_out, int64_t n, cudaStream_t st,
                         int64_t* launches) {
    *launches += 6;
    return cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, keys_in, keys_out, val_in, val_out, n, 0,
                                           32, st);
}

}  // namespace sfm
