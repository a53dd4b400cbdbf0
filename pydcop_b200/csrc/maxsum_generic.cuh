// Generic MaxSum kernels: any arity <= FG_MAX_ARITY, any domain sizes.  One thread per directed
// edge on both sides.  These are the always-correct members of the kernel family; the shape-
// specialised kernels in maxsum_fast.cuh take over for the hot classes.
#pragma once
#include "common.cuh"

// QMODE: which incoming v->f messages count in the marginal
//   0 all edges hold a message (every cycle >= 2)   1 consult q_valid (cycle 1)   2 none (on_start)
// factor_costs_for_var (maxsum.py:382-447) for scope position j of factor f of class c; the result
// for value xv is returned one value at a time through `emit(xv, value)`.
template <typename T, int QMODE, typename Emit>
__device__ __forceinline__ void factor_marginal_generic(const fg_class_t &c, const T *__restrict__ table,
                                                        const T *__restrict__ q,
                                                        const int64_t *__restrict__ edge_qoff_f,
                                                        const uint8_t *__restrict__ q_valid_f, int j,
                                                        bool mode_max, Emit emit) {
  const int a = c.arity;
  int64_t stride[FG_MAX_ARITY], qoff[FG_MAX_ARITY];
  bool use[FG_MAX_ARITY];
  {
    int64_t s = 1;
    for (int i = a - 1; i >= 0; --i) { stride[i] = s; s *= c.dom[i]; }
    for (int i = 0; i < a; ++i) {
      use[i] = (i != j) && (QMODE == 0 || (QMODE == 1 && q_valid_f[i] != 0));
      qoff[i] = use[i] ? edge_qoff_f[i] : 0;
    }
  }
  const int dj = c.dom[j];
  for (int xv = 0; xv < dj; ++xv) {
    T opt = mode_max ? -Inf<T>::pos() : Inf<T>::pos();
    int x[FG_MAX_ARITY];
    for (int i = 0; i < a; ++i) x[i] = 0;
    x[j] = xv;
    for (;;) {
      int64_t idx = 0;
      T sum = (T)0;
      for (int i = 0; i < a; ++i) {
        idx += x[i] * stride[i];
        if (use[i]) sum += q[qoff[i] + x[i]];
      }
      T cur = table[idx] + sum;
      opt_update(opt, cur, mode_max);
      int i = a - 1;
      for (; i >= 0; --i) {
        if (i == j) continue;
        if (++x[i] < c.dom[i]) break;
        x[i] = 0;
      }
      if (i < 0) break;
    }
    emit(xv, opt);
  }
}

// One thread per edge (f, j) of one class.  Reads q_cur / r_cur, writes r_next.  The validity
// arrays are only READ here: both sides of cycle 1 must see them as of the end of cycle 0, so the
// host sets them to all-ones after cycle 1 (every edge has posted by then).
template <typename T, int QMODE>
__global__ void __launch_bounds__(128)
k_f2v_generic(const fg_class_t c, const T *__restrict__ tables, const T *__restrict__ q_cur,
              const T *__restrict__ r_cur, T *__restrict__ r_next,
              const int64_t *__restrict__ edge_qoff, const uint8_t *__restrict__ q_valid,
              uint8_t *__restrict__ r_cnt, uint8_t *__restrict__ r_sent, MaxSumParams p) {
  int64_t le = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (le >= (int64_t)c.n_factors * c.arity) return;
  int f = (int)(le / c.arity), j = (int)(le % c.arity);
  int e = c.first_edge + (int)le;
  const T *table = tables + c.table_base + (int64_t)f * c.table_size;
  int64_t row_base = c.msg_base + (int64_t)f * c.row_total;
  int64_t row = row_base + c.row_off[j];
  const int d = c.dom[j];
  const T lam = (T)p.damping, oml = (T)p.one_minus_damping, stab = (T)p.stability;
  uint8_t cnt = r_cnt[e];
  const bool has_prev = cnt & 1;
  const bool damp = p.damp_factors && has_prev;
  bool match = has_prev;
  factor_marginal_generic<T, QMODE>(
      c, table, q_cur, edge_qoff + (c.first_edge + f * c.arity),
      q_valid + (c.first_edge + f * c.arity), j, p.mode_max != 0,
      [&](int xv, T cand) {
        T prev = r_cur[row + xv];
        if (damp) cand = lam * prev + oml * cand;
        if (has_prev && !approx_match1<T>(cand, prev, stab)) match = false;
        r_next[row + xv] = cand;
      });
  bool sent = gate_decide(match, cnt);
  if (!sent)
    for (int xv = 0; xv < d; ++xv) r_next[row + xv] = r_cur[row + xv];
  r_cnt[e] = cnt;
  if (r_sent) r_sent[e] = sent ? 1 : 0;
}

// on_start of factors (maxsum.py:305-328): unary factors (leafs / leafs_vars) or all factors
// (all) post their marginal computed with NO incoming message; not recorded as prev.
template <typename T>
__global__ void __launch_bounds__(128)
k_f2v_start(const fg_class_t c, const T *__restrict__ tables, T *__restrict__ r_cur,
            uint8_t *__restrict__ r_valid, uint8_t *__restrict__ r_sent, int mode_max) {
  int64_t le = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (le >= (int64_t)c.n_factors * c.arity) return;
  int f = (int)(le / c.arity), j = (int)(le % c.arity);
  int e = c.first_edge + (int)le;
  const T *table = tables + c.table_base + (int64_t)f * c.table_size;
  int64_t row_base = c.msg_base + (int64_t)f * c.row_total;
  int64_t row = row_base + c.row_off[j];
  factor_marginal_generic<T, 2>(c, table, (const T *)nullptr, nullptr, nullptr, j, mode_max != 0,
                                [&](int xv, T cand) { r_cur[row + xv] = cand; });
  r_valid[e] = 1;
  if (r_sent) r_sent[e] = 1;
}

// ---------------------------------------------------------------------------------------------
// variable side
// ---------------------------------------------------------------------------------------------
struct VarSide {
  const int32_t *dom_size;
  const int64_t *unary_off;
  const int32_t *var_ptr;
  const int64_t *var_qbase;  // q row of slot s of v = var_qbase[v] + (s - var_ptr[v]) * dom_size[v]
  const int64_t *slot_roff;  // r row of slot s
  const int32_t *slot_edge;
  const int32_t *slot_var;
};

// select_value (maxsum.py:584-620): costs summed in `links` order, first optimum wins.
template <typename T, int RMODE>
__device__ __forceinline__ void select_value_generic(const VarSide &g, const T *__restrict__ unary,
                                                     const T *__restrict__ r,
                                                     const uint8_t *__restrict__ r_valid, int v,
                                                     bool mode_max, int32_t *value, T *value_cost) {
  const int d = g.dom_size[v];
  const int s0 = g.var_ptr[v], s1 = g.var_ptr[v + 1];
  const int64_t u0 = g.unary_off[v];
  int best = 0;
  T best_c = (T)0;
  for (int x = 0; x < d; ++x) {
    T c = unary[u0 + x];
    if (RMODE != 2)
      for (int s = s0; s < s1; ++s) {
        if (RMODE == 1 && !r_valid[g.slot_edge[s]]) continue;
        c += r[g.slot_roff[s] + x];
      }
    if (x == 0 || (mode_max ? (c > best_c) : (c < best_c))) { best = x; best_c = c; }
  }
  value[v] = best;
  value_cost[v] = best_c;
}

// One thread per slot (v, s).  costs_for_factor (maxsum.py:623-676) + damping + gate; the thread
// of a variable's first slot also runs select_value.  RMODE as QMODE above, for f->v messages.
template <typename T, int RMODE>
__global__ void __launch_bounds__(128)
k_v2f_generic(VarSide g, int slot_begin, int n_slots, const T *__restrict__ unary, const T *__restrict__ r_cur,
              const T *__restrict__ q_cur, T *__restrict__ q_next,
              const uint8_t *__restrict__ r_valid, uint8_t *__restrict__ q_cnt, uint8_t *__restrict__ q_sent, int32_t *__restrict__ value,
              T *__restrict__ value_cost, MaxSumParams p) {
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_slots) return;
  s += slot_begin;
  const int v = g.slot_var[s];
  const int d = g.dom_size[v];
  const int s0 = g.var_ptr[v], s1 = g.var_ptr[v + 1];
  const int64_t u0 = g.unary_off[v];
  const int64_t row = g.var_qbase[v] + (int64_t)(s - s0) * d;
  const T lam = (T)p.damping, oml = (T)p.one_minus_damping, stab = (T)p.stability;
  if (s == s0)
    select_value_generic<T, RMODE>(g, unary, r_cur, r_valid, v, p.mode_max != 0, value, value_cost);
  // pass 1: un-normalised sums, in the reference's (value-major, then factor) order
  T sum_cost = (T)0;
  for (int x = 0; x < d; ++x) {
    T m = unary[u0 + x];
    for (int t = s0; t < s1; ++t) {
      if (t == s) continue;
      if (RMODE == 1 && !r_valid[g.slot_edge[t]]) continue;
      T c = r_cur[g.slot_roff[t] + x];
      sum_cost += c;
      m += c;
    }
    q_next[row + x] = m;
  }
  const T avg = sum_cost / (T)d;
  uint8_t cnt = q_cnt[s];
  const bool has_prev = cnt & 1;
  const bool damp = p.damp_vars && has_prev;
  bool match = has_prev;
  for (int x = 0; x < d; ++x) {
    T cand = q_next[row + x] - avg;
    T prev = q_cur[row + x];
    if (damp) cand = lam * prev + oml * cand;
    if (has_prev && !approx_match1<T>(cand, prev, stab)) match = false;
    q_next[row + x] = cand;
  }
  bool sent = gate_decide(match, cnt);
  if (!sent)
    for (int x = 0; x < d; ++x) q_next[row + x] = q_cur[row + x];
  q_cnt[s] = cnt;
  if (q_sent) q_sent[s] = sent ? 1 : 0;
}

// on_start of variables (maxsum.py:495-523): initial value; leaf variables (leafs) or all
// variables (leafs_vars / all) post q = own costs (normalisation subtracts 0/|D| = 0.0).
template <typename T>
__global__ void __launch_bounds__(128)
k_v2f_start(VarSide g, int n_vars, const T *__restrict__ unary, const int32_t *__restrict__ init_value,
            T *__restrict__ q_cur, uint8_t *__restrict__ q_valid, uint8_t *__restrict__ q_sent,
            int32_t *__restrict__ value, T *__restrict__ value_cost, int mode_max,
            int start_messages) {
  int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= n_vars) return;
  select_value_generic<T, 2>(g, unary, (const T *)nullptr, nullptr, v, mode_max != 0, value,
                             value_cost);
  if (init_value && init_value[v] >= 0) { value[v] = init_value[v]; value_cost[v] = (T)0; }
  const int s0 = g.var_ptr[v], s1 = g.var_ptr[v + 1];
  const int d = g.dom_size[v];
  const int64_t u0 = g.unary_off[v];
  if (((s1 - s0) == 1 && start_messages == FG_START_LEAFS) || start_messages >= FG_START_LEAFS_VARS) {
    for (int s = s0; s < s1; ++s) {
      const int64_t row = g.var_qbase[v] + (int64_t)(s - s0) * d;
      const T avg = (T)0 / (T)d;
      for (int x = 0; x < d; ++x) q_cur[row + x] = unary[u0 + x] - avg;
      q_valid[g.slot_edge[s]] = 1;
      if (q_sent) q_sent[s] = 1;
    }
  }
}
