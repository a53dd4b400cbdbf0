// Philox4x32-10, the injected-draw generator of the DSA parity definition (oracle/philox.py).
#pragma once
#include <stdint.h>

__host__ __device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2,
                                                       uint32_t c3, uint32_t k0, uint32_t k1,
                                                       uint32_t out[4]) {
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

#define FG_PHILOX_INIT_CYCLE 0xFFFFFFFFu

// u in [0,1) with 53 bits
__host__ __device__ __forceinline__ double philox_u53(const uint32_t b[4]) {
  return ((double)(b[0] >> 5) * 67108864.0 + (double)(b[1] >> 6)) / 9007199254740992.0;
}
__host__ __device__ __forceinline__ int philox_choice(const uint32_t b[4], int n) {
  return (int)(((uint64_t)b[2] * (uint64_t)n) >> 32);
}
