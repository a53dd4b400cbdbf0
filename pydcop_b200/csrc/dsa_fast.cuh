// Shape-specialised DSA kernels (filled in after the generic path is parity-green).
#pragma once
#include <vector>
#include "common.cuh"

template <typename T>
inline bool dsa_fast_step(const fg_dsa_desc_t &, const std::vector<fg_class_t> &, const int32_t *, int32_t *,
                          uint32_t, cudaStream_t, int64_t &) {
  return false;
}
