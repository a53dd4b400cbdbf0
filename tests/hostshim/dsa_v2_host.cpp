// TEST INFRASTRUCTURE: runs the chunked DSA kernel SOURCE (pydcop_b200/csrc/dsa_v2_kernels.cuh) on
// the CPU, one "thread" after the other.  Nothing in the product links or calls this.
#include <stdint.h>

struct Dim3 { unsigned x, y, z; };
static thread_local Dim3 blockIdx, blockDim, threadIdx;
struct float4 { float x, y, z, w; };
struct double2 { double x, y; };
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(n)

#include "../../pydcop_b200/csrc/dsa_v2_kernels.cuh"

struct dsa_v2_host_arrays {
  const int32_t *var_ptr, *slot_nbr;
  const int64_t *slot_tab;
  const void *slot_opt, *tables_or;
  const uint8_t *has_nbr;
  const double *prob;
  const int32_t *var_id, *val;
  int32_t *val_next;
  void *val_cost;
  int32_t n_vars, precision, dom, chunk, mode_max, variant;
  uint64_t seed;
};

template <typename T, int D, int U>
static void run(const dsa_v2_host_arrays *a, uint32_t cycle) {
  blockDim = Dim3{128, 1, 1};
  for (unsigned b = 0; b < (unsigned)((a->n_vars + 127) / 128); ++b)
    for (unsigned t = 0; t < 128; ++t) {
      blockIdx = Dim3{b, 0, 0};
      threadIdx = Dim3{t, 0, 0};
      k_dsa_step_bin_v2<T, D, U>(a->n_vars, a->var_ptr, a->slot_nbr, a->slot_tab, (const T *)a->slot_opt,
                                 (const T *)a->tables_or, a->has_nbr, a->prob, a->var_id, a->val, a->val_next,
                                 (T *)a->val_cost, a->mode_max, a->variant, a->seed, cycle);
    }
}

template <typename T, int U>
static int by_dom(const dsa_v2_host_arrays *a, uint32_t cycle) {
  switch (a->dom) {
    case 4: run<T, 4, U>(a, cycle); return 0;
    case 8: run<T, 8, U>(a, cycle); return 0;
    case 10: run<T, 10, U>(a, cycle); return 0;
    case 16: run<T, 16, U>(a, cycle); return 0;
    case 20: run<T, 20, U>(a, cycle); return 0;
  }
  return 3;
}

extern "C" int dsa_v2_host_step(const dsa_v2_host_arrays *a, uint32_t cycle) {
  if (a->precision == FG_F64) return a->chunk == 2 ? by_dom<double, 2>(a, cycle) : by_dom<double, 4>(a, cycle);
  return a->chunk == 2 ? by_dom<float, 2>(a, cycle) : by_dom<float, 4>(a, cycle);
}
