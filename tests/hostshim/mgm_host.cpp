// TEST INFRASTRUCTURE: runs the MGM CUDA kernel SOURCE (pydcop_b200/csrc/mgm_kernels.cuh) on the
// CPU, one "thread" after the other, so its logic can be compared with the oracle and the
// reference trajectories without a GPU.  Nothing in the product links or calls this.
//   g++ -O1 -ffp-contract=off -shared -fPIC -o mgm_host.so mgm_host.cpp
#include <stdint.h>

struct Dim3 { unsigned x, y, z; };
static thread_local Dim3 blockIdx, blockDim, threadIdx;
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(n)

#include "../../pydcop_b200/csrc/mgm_kernels.cuh"

template <typename F>
static void launch(int n, int block, F body) {
  blockDim = Dim3{(unsigned)block, 1, 1};
  for (unsigned b = 0; b < (unsigned)((n + block - 1) / block); ++b)
    for (unsigned t = 0; t < (unsigned)block; ++t) {
      blockIdx = Dim3{b, 0, 0};
      threadIdx = Dim3{t, 0, 0};
      body();
    }
}

extern "C" {

struct mgm_host_arrays {
  const fg_class_t *classes;
  const int32_t *dom_size, *var_id, *var_rank, *edge_var, *edge_class, *var_ptr, *slot_edge, *nbr_ptr, *nbr_idx;
  const int64_t *unary_off;
  const int32_t *init_value;
  const void *tables, *unary;
  int32_t *value;
  void *cost;
  uint8_t *has_cost;
  void *gain;
  int32_t *new_value;
  int32_t n_vars, precision, mode_max;
  uint64_t seed;
};

static MgmSide side(const mgm_host_arrays *a) {
  return MgmSide{a->classes, a->dom_size, a->var_id, a->var_rank, a->edge_var, a->edge_class,
                 a->var_ptr, a->slot_edge, a->nbr_ptr, a->nbr_idx, a->unary_off};
}

void mgm_host_init(const mgm_host_arrays *a) {
  launch(a->n_vars, 128, [&] { k_mgm_init(side(a), a->n_vars, a->init_value, a->seed, a->value, a->has_cost); });
}

}  // extern "C"

template <typename T>
static void cycle_t(const mgm_host_arrays *a, uint32_t cycle) {
  launch(a->n_vars, 128, [&] {
    k_mgm_gain<T>(side(a), a->n_vars, (const T *)a->tables, (const T *)a->unary, a->value, (T *)a->cost,
                  a->has_cost, (T *)a->gain, a->new_value, a->mode_max, a->seed, cycle);
  });
  // the decide kernel of the GPU sees a consistent `gain` array (kernel boundary); sequential
  // execution gives the same because gain / new_value are not written in this phase
  launch(a->n_vars, 256, [&] {
    k_mgm_decide<T>(side(a), a->n_vars, (const T *)a->gain, a->new_value, a->value, (T *)a->cost);
  });
}

extern "C" {
void mgm_host_cycle(const mgm_host_arrays *a, uint32_t cycle) {
  if (a->precision == FG_F64) cycle_t<double>(a, cycle);
  else cycle_t<float>(a, cycle);
}
}
