// Stable device radix sorts of (key, value) pairs over the low `end_bit` key bits (cub::DeviceRadixSort underneath; preparatory
// steps only — the cell order of the tile search, the cluster order of the dynamic-vehicle filter — never a per-iteration kernel).
#pragma once
#include <cuda_runtime.h>

#include <cstddef>
#include <cstdint>

namespace sage {

size_t sort_pairs_tmp_bytes_u32(size_t n, int end_bit);
size_t sort_pairs_tmp_bytes_u64(size_t n, int end_bit);
// returns the number of kernels cub launches for this sort (for the launch counter)
int sort_pairs_u32(void *tmp, size_t tmp_bytes, const uint32_t *keys_in, uint32_t *keys_out, const uint32_t *vals_in, uint32_t *vals_out, size_t n,
                   int end_bit, cudaStream_t stream);
int sort_pairs_u64(void *tmp, size_t tmp_bytes, const unsigned long long *keys_in, unsigned long long *keys_out, const uint32_t *vals_in,
                   uint32_t *vals_out, size_t n, int end_bit, cudaStream_t stream);

}  // namespace sage
