// Row loader shared by the fast DSA-family kernels: D contiguous elements, 16-byte loads when the row
// size allows (rows start at multiples of D elements from a 128-byte aligned base).
// Free of CUDA runtime includes (tests/hostshim/ compiles it on the CPU).
#pragma once
#include <stdint.h>

// D contiguous elements; 16-byte loads when the row size allows (rows start at multiples of D
// elements from a 128-byte aligned base)
template <typename T, int D>
__host__ __device__ __forceinline__ void fg_load_row(const T *__restrict__ p, T (&r)[D]) {
  if constexpr (sizeof(T) == 4 && D % 4 == 0) {
    const float4 *q = reinterpret_cast<const float4 *>(p);
#pragma unroll
    for (int i = 0; i < D / 4; ++i) {
      const float4 v = q[i];
      r[4 * i] = v.x; r[4 * i + 1] = v.y; r[4 * i + 2] = v.z; r[4 * i + 3] = v.w;
    }
  } else if constexpr (sizeof(T) == 8 && D % 2 == 0) {
    const double2 *q = reinterpret_cast<const double2 *>(p);
#pragma unroll
    for (int i = 0; i < D / 2; ++i) {
      const double2 v = q[i];
      r[2 * i] = v.x; r[2 * i + 1] = v.y;
    }
  } else {
#pragma unroll
    for (int i = 0; i < D; ++i) r[i] = p[i];
  }
}
