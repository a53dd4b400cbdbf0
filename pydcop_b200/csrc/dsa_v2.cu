// Opt-in experiment: the chunked DSA step (dsa_v2_kernels.cuh) behind its own C-ABI entry, in its
// own translation unit so the validated kernels of engine.cu are not recompiled.  The caller
// (pydcop_b200.engine.DsaEngine with PYDCOP_B200_DSA_V2=<U>) passes the same descriptor it gave
// fg_dsa_create, the current buffer index and cycle from fg_dsa_current, and commits with
// fg_dsa_cycle_commit as usual.
#include "common.cuh"
#include "dsa_v2_kernels.cuh"

template <typename T, int D, int U>
static void launch(const fg_dsa_desc_t &d, const int32_t *val, int32_t *val_next, uint32_t cycle, cudaStream_t st) {
  const unsigned blocks = (unsigned)((d.n_vars + 127) / 128);
  k_dsa_step_bin_v2<T, D, U><<<blocks, 128, 0, st>>>(
      d.n_vars, d.dev_var_ptr, d.dev_slot_nbr, d.dev_slot_tab, (const T *)d.dev_slot_opt, (const T *)d.dev_tables_or,
      d.dev_has_nbr, d.dev_prob, d.dev_var_id, val, val_next, (T *)d.dev_value_cost, d.mode_max, d.variant, d.seed, cycle);
}

template <typename T, int U>
static int dispatch(const fg_dsa_desc_t &d, const int32_t *val, int32_t *val_next, uint32_t cycle, cudaStream_t st) {
  switch (d.fast_dom) {
    case 4: launch<T, 4, U>(d, val, val_next, cycle, st); return FG_OK;
    case 8: launch<T, 8, U>(d, val, val_next, cycle, st); return FG_OK;
    case 10: launch<T, 10, U>(d, val, val_next, cycle, st); return FG_OK;
    case 16: launch<T, 16, U>(d, val, val_next, cycle, st); return FG_OK;
    case 20: launch<T, 20, U>(d, val, val_next, cycle, st); return FG_OK;
  }
  return FG_ERR_UNSUPPORTED;
}

extern "C" int fg_dsa_step_v2(const fg_dsa_desc_t *desc, int32_t cur, int64_t cycle, int32_t chunk, void *stream) {
  if (!desc || (cur != 0 && cur != 1)) return FG_ERR_ARG;
  const fg_dsa_desc_t &d = *desc;
  if (!d.dev_tables_or || !d.dev_slot_nbr || !d.dev_slot_tab || !d.dev_slot_opt || d.fast_dom <= 0) return FG_ERR_UNSUPPORTED;
  if (d.stop_cycle && cycle >= d.stop_cycle) return FG_OK;
  if (!d.n_vars) return FG_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const int32_t *val = d.dev_value[cur];
  int32_t *val_next = d.dev_value[cur ^ 1];
  int rc;
  if (d.precision == FG_F64) {
    rc = chunk == 2 ? dispatch<double, 2>(d, val, val_next, (uint32_t)cycle, st)
                    : (chunk == 4 ? dispatch<double, 4>(d, val, val_next, (uint32_t)cycle, st) : FG_ERR_ARG);
  } else {
    rc = chunk == 2 ? dispatch<float, 2>(d, val, val_next, (uint32_t)cycle, st)
                    : (chunk == 4 ? dispatch<float, 4>(d, val, val_next, (uint32_t)cycle, st) : FG_ERR_ARG);
  }
  if (rc != FG_OK) return rc;
  return cudaGetLastError() == cudaSuccess ? FG_OK : FG_ERR_CUDA;
}
