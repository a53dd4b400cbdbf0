// Diagnostics exported through the C-ABI (not on any product path).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

// Random-row gather: the access pattern of the DSA / MGM value kernels (one row of `row_bytes` out of a
// table region per incidence, row start = a pseudo-random multiple of `stride_bytes`) without their
// arithmetic and without the dependent index loads — the HBM throughput this pattern can reach is the
// honest ceiling for those kernels (tools/gather_peak.py).
template <int ROWS>
__global__ void __launch_bounds__(128)
k_selftest_gather(const uint8_t *__restrict__ base, uint64_t n_slots, int row_bytes, int stride_bytes, int64_t n_threads,
                  float *__restrict__ out) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_threads) return;
  float acc = 0.f;
  const float4 *rows[ROWS];
#pragma unroll
  for (int i = 0; i < ROWS; ++i) {
    uint64_t h = ((uint64_t)t * ROWS + i) * 0x9E3779B97F4A7C15ull;
    h ^= h >> 29; h *= 0xBF58476D1CE4E5B9ull; h ^= h >> 32;
    rows[i] = reinterpret_cast<const float4 *>(base + (h % n_slots) * (uint64_t)stride_bytes);
  }
  const int nv = row_bytes / 16;
  for (int k = 0; k < nv; ++k) {
#pragma unroll
    for (int i = 0; i < ROWS; ++i) {
      const float4 v = rows[i][k];
      acc += v.x + v.y + v.z + v.w;
    }
  }
  out[t] = acc;
}
