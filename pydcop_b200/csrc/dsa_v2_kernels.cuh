// DSA step, chunked variant (opt-in experiment): same inputs, same arithmetic and
// the same results as k_dsa_step_bin (dsa_fast.cuh), but the slot loop handles U incidences per
// trip: their neighbour ids / table offsets are loaded together, their U neighbour values are
// gathered together, and the U table rows are loaded by unconditional (always in-bounds) loads that
// the compiler can hoist — the three dependent loads per incidence (id -> value -> row) overlap
// across U incidences instead of running back to back.  Costs are still added in slot order, so the
// sums are bit-identical.
// Free of CUDA runtime includes: tests/hostshim/ runs this file on the CPU (test infrastructure).
#pragma once
#include <stdint.h>

#include "../../include/pydcop_b200.h"
#include "philox.cuh"

template <typename T> struct DsaV2Inf;
template <> struct DsaV2Inf<float> { __host__ __device__ static float pos() { return __builtin_huge_valf(); } };
template <> struct DsaV2Inf<double> { __host__ __device__ static double pos() { return __builtin_huge_val(); } };

template <typename T> __host__ __device__ __forceinline__ T dsa_v2_abs(T x) { return x < (T)0 ? -x : x; }

// D contiguous elements; 16-byte loads when the row size allows (rows start at multiples of D
// elements from a 128-byte aligned base)
template <typename T, int D>
__host__ __device__ __forceinline__ void dsa_v2_load_row(const T *__restrict__ p, T (&r)[D]) {
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

template <typename T, int D, int U>
__global__ void __launch_bounds__(128)
k_dsa_step_bin_v2(int n_vars, const int32_t *__restrict__ var_ptr, const int32_t *__restrict__ slot_nbr,
                  const int64_t *__restrict__ slot_tab, const T *__restrict__ slot_opt,
                  const T *__restrict__ tables_or, const uint8_t *__restrict__ has_nbr,
                  const double *__restrict__ prob, const int32_t *__restrict__ var_id,
                  const int32_t *__restrict__ val, int32_t *__restrict__ val_next, T *__restrict__ val_cost,
                  int mode_max, int variant, uint64_t seed, uint32_t cycle) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= n_vars) return;
  const int cur = val[v];
  if (!has_nbr[v]) { val_next[v] = cur; return; }
  T cost[D];
#pragma unroll
  for (int x = 0; x < D; ++x) cost[x] = (T)0;
  bool violated = false;
  const int s0 = var_ptr[v], s1 = var_ptr[v + 1];
  for (int s = s0; s < s1; s += U) {
    int nb[U];
    int64_t tb[U];
    bool ok[U];
#pragma unroll
    for (int i = 0; i < U; ++i) {  // slots past the end re-read slot s (in bounds) and are not added
      ok[i] = s + i < s1;
      const int si = ok[i] ? s + i : s;
      nb[i] = slot_nbr[si];
      tb[i] = slot_tab[si];
    }
    int y[U];
#pragma unroll
    for (int i = 0; i < U; ++i) y[i] = val[nb[i]];
#pragma unroll
    for (int i = 0; i < U; ++i) {
      const T *row = tables_or + tb[i] + (int64_t)y[i] * D;
      T r[D];
      dsa_v2_load_row<T, D>(row, r);
      // assignment_cost, relations.py:1479-1532.  Select between the old sum and old + row instead of
      // branching: the row loads stay unconditional, so the rows of the whole chunk can be in flight
      // together; a slot past the end leaves the sum untouched (exactly, no +0.0)
#pragma unroll
      for (int x = 0; x < D; ++x) {
        const T nc = cost[x] + r[x];
        cost[x] = ok[i] ? nc : cost[x];
      }
      if (variant == FG_DSA_B && ok[i] && row[cur] != slot_opt[s + i]) violated = true;  // dsa.py:419-431
    }
  }
  // find_optimal (relations.py:1594-1638)
  T best_cost = mode_max ? -DsaV2Inf<T>::pos() : DsaV2Inf<T>::pos();
  int nbest = 0;
  T cur_cost = (T)0;
#pragma unroll
  for (int x = 0; x < D; ++x) {
    const T c = cost[x];
    if (x == cur) cur_cost = c;
    if (c == best_cost) ++nbest;
    else if (mode_max ? (c > best_cost) : (c < best_cost)) { best_cost = c; nbest = 1; }
  }
  const T delta = dsa_v2_abs<T>(cur_cost - best_cost);
  bool attempt = false, drop_cur = false;
  if (delta > (T)0) {
    attempt = true;
  } else if (delta == (T)0) {
    if (variant == FG_DSA_C || (variant == FG_DSA_B && violated)) {
      attempt = true;
      drop_cur = nbest > 1;
    }
  }
  int nv = cur;
  if (attempt) {  // probabilistic_change, dsa.py:407-417
    uint32_t b[4];
    philox4x32_10((uint32_t)var_id[v], cycle, 0u, 0u, (uint32_t)seed, (uint32_t)(seed >> 32), b);
    if (prob[v] > philox_u53(b)) {
      int pick = philox_choice(b, drop_cur ? nbest - 1 : nbest);
      bool done = false;
#pragma unroll
      for (int x = 0; x < D; ++x) {
        if (!done && cost[x] == best_cost && !(drop_cur && x == cur)) {
          if (pick == 0) { nv = x; done = true; }
          --pick;
        }
      }
      val_cost[v] = best_cost;
    }
  }
  val_next[v] = nv;
}
