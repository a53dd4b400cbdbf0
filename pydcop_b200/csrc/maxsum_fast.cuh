// Shape-specialised MaxSum kernels (filled in after the generic path is parity-green).
#pragma once
#include <vector>
#include "common.cuh"

struct MaxSumFastPlan {
  bool enabled = false;
};

inline void maxsum_fast_plan(const fg_maxsum_desc_t &, const std::vector<fg_class_t> &, MaxSumFastPlan &) {}

template <typename T>
inline bool maxsum_fast_f2v(const MaxSumFastPlan &, int, const fg_class_t &, const fg_maxsum_desc_t &,
                            const T *, const T *, T *, const MaxSumParams &, cudaStream_t, int64_t &) {
  return false;
}

template <typename T>
inline bool maxsum_fast_v2f(const MaxSumFastPlan &, const fg_maxsum_desc_t &, const T *, const T *, T *,
                            const MaxSumParams &, cudaStream_t, int64_t &) {
  return false;
}
