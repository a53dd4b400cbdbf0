/* CPU ORACLE — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C restatement of the reference's MaxSum / DSA hot path (pyDcop, pure Python):
 *   pydcop/algorithms/maxsum.py:305-328,339-379,382-447,495-565,584-710
 *   pydcop/algorithms/dsa.py:277-431, pydcop/dcop/relations.py:1367-1400,1479-1532,1594-1638
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may
 * load this library.  It is pinned against the reference's own lock-step trajectories in
 * tests/golden/*.npz (made by oracle/make_golden.py) by tests/test_oracle_golden.py.
 *
 * This file is included twice by dcop_oracle.c with REAL = double (suffix _f64: the reference's
 * own arithmetic type, Python float) and REAL = float (suffix _f32: the engine's throughput type).
 * Every floating-point expression keeps the reference's operand ORDER so that the f64 build is
 * bit-identical to the Python reference; compile with -ffp-contract=off.
 *
 * Edge state flag byte: bit0 RECV (receiver holds a message), bit1 PREV (sender recorded a
 * previous message in on_new_cycle, maxsum.py:364,372,551,559), bits 2..4 = send count.
 */

#define CAT_(a, b) a##b
#define CAT(a, b) CAT_(a, b)
#define FN(name) CAT(name, SUFFIX)

/* maxsum.py:688-710 */
static int FN(approx_match)(const REAL *costs, const REAL *prev, int d, REAL stability) {
  for (int x = 0; x < d; ++x) {
    REAL c = costs[x], prev_c = prev[x];
    if (prev_c != c) {
      REAL delta = FABS(prev_c - c);
      if (prev_c + c != (REAL)0) {
        if (!(((REAL)2 * delta / FABS(prev_c + c)) < stability)) return 0;
      } else {
        return 0;
      }
    }
  }
  return 1;
}

/* damping (maxsum.py:679-685) + send gate (maxsum.py:356-377 / 545-564), shared by both sides.
 * `cand` holds the freshly computed message; `state` is the receiver-side copy (== sender's
 * prev message whenever PREV is set).  Writes the next state row and flag byte. */
static void FN(damp_and_gate)(REAL *cand, const REAL *state, REAL *next, int d, uint8_t flags,
                              uint8_t *next_flags, uint8_t *sent, int damp, REAL lam, REAL oml,
                              REAL stability) {
  int has_prev = (flags & FLAG_PREV) != 0;
  int cnt = flags >> 2;
  if (damp && has_prev)
    for (int x = 0; x < d; ++x) cand[x] = lam * state[x] + oml * cand[x];
  int match = has_prev && FN(approx_match)(cand, state, d, stability);
  if (!match) {
    for (int x = 0; x < d; ++x) next[x] = cand[x];
    *next_flags = FLAG_RECV | FLAG_PREV | (1u << 2);
    *sent = 1;
  } else if (cnt < SAME_COUNT) {
    for (int x = 0; x < d; ++x) next[x] = cand[x];
    *next_flags = FLAG_RECV | FLAG_PREV | (uint8_t)((cnt + 1) << 2);
    *sent = 1;
  } else {
    for (int x = 0; x < d; ++x) next[x] = state[x];
    *next_flags = flags;
    *sent = 0;
  }
}

/* factor_costs_for_var, maxsum.py:382-447: min/max-marginal of factor f for scope position j. */
static void FN(factor_marginal)(const fg_t *g, const REAL *tables, const REAL *q,
                                const uint8_t *q_flags, int f, int j, int mode_max, REAL *out) {
  int e0 = g->factor_ptr[f], a = g->factor_ptr[f + 1] - e0;
  const REAL *T = tables + g->table_off[f];
  int dsz[MAX_ARITY], x[MAX_ARITY];
  int64_t stride[MAX_ARITY];
  for (int i = 0; i < a; ++i) dsz[i] = g->dom_size[g->edge_var[e0 + i]];
  int64_t s = 1;
  for (int i = a - 1; i >= 0; --i) { stride[i] = s; s *= dsz[i]; }
  for (int xv = 0; xv < dsz[j]; ++xv) {
    REAL opt = mode_max ? -(REAL)INFINITY : (REAL)INFINITY;
    for (int i = 0; i < a; ++i) x[i] = 0;
    x[j] = xv;
    for (;;) {
      int64_t idx = 0;
      REAL sum = (REAL)0;
      for (int i = 0; i < a; ++i) {
        idx += x[i] * stride[i];
        if (i != j && (q_flags[e0 + i] & FLAG_RECV)) sum += q[g->msg_off[e0 + i] + x[i]];
      }
      REAL cur = T[idx] + sum;
      if (mode_max ? (opt < cur) : (opt > cur)) opt = cur;
      /* odometer over the other variables, last axis fastest (order is irrelevant for the min) */
      int i = a - 1;
      for (; i >= 0; --i) {
        if (i == j) continue;
        if (++x[i] < dsz[i]) break;
        x[i] = 0;
      }
      if (i < 0) break;
    }
    out[xv] = opt;
  }
}

/* select_value, maxsum.py:584-620 (received costs summed in `links` order; first optimum wins) */
static void FN(select_value)(const fg_t *g, const REAL *unary, const REAL *r,
                             const uint8_t *r_flags, int v, int mode_max, int32_t *value,
                             REAL *value_cost) {
  int d = g->dom_size[v];
  int best = 0;
  REAL best_c = 0;
  for (int x = 0; x < d; ++x) {
    REAL c = unary[g->unary_off[v] + x];
    for (int s = g->var_ptr[v]; s < g->var_ptr[v + 1]; ++s) {
      int e = g->var_edge[s];
      if (r_flags[e] & FLAG_RECV) c += r[g->msg_off[e] + x];
    }
    if (x == 0 || (mode_max ? (c > best_c) : (c < best_c))) { best = x; best_c = c; }
  }
  *value = best;
  *value_cost = best_c;
}

/* costs_for_factor, maxsum.py:623-676: message of variable v for the edge in slot `slot`. */
static void FN(var_message)(const fg_t *g, const REAL *unary, const REAL *r,
                            const uint8_t *r_flags, int v, int slot, REAL *out) {
  int d = g->dom_size[v];
  REAL sum_cost = (REAL)0;
  for (int x = 0; x < d; ++x) {
    REAL m = unary[g->unary_off[v] + x];
    for (int s = g->var_ptr[v]; s < g->var_ptr[v + 1]; ++s) {
      if (s == slot) continue;
      int e = g->var_edge[s];
      if (!(r_flags[e] & FLAG_RECV)) continue;
      REAL c = r[g->msg_off[e] + x];
      sum_cost += c;
      m += c;
    }
    out[x] = m;
  }
  REAL avg = sum_cost / (REAL)d;
  for (int x = 0; x < d; ++x) out[x] = out[x] - avg;
}

/* Cycle 0: on_start of every computation (maxsum.py:305-328, 495-523).
 * start_messages: 0 leafs, 1 leafs_vars, 2 all.  Start posts set RECV but not PREV. */
void FN(maxsum_oracle_init)(const fg_t *g, const REAL *tables, const REAL *unary,
                            const int32_t *init_value, int mode_max, int start_messages, REAL *q,
                            REAL *r, uint8_t *q_flags, uint8_t *r_flags, uint8_t *q_sent,
                            uint8_t *r_sent, int32_t *value, REAL *value_cost) {
  int64_t M = g->msg_off[g->E];
  memset(q, 0, sizeof(REAL) * M);
  memset(r, 0, sizeof(REAL) * M);
  memset(q_flags, 0, g->E);
  memset(r_flags, 0, g->E);
  memset(q_sent, 0, g->E);
  memset(r_sent, 0, g->E);
  for (int v = 0; v < g->V; ++v) {
    FN(select_value)(g, unary, r, r_flags, v, mode_max, &value[v], &value_cost[v]);
    if (init_value && init_value[v] >= 0) { value[v] = init_value[v]; value_cost[v] = 0; }
    int k = g->var_ptr[v + 1] - g->var_ptr[v];
    if ((k == 1 && start_messages == 0) || start_messages >= 1) {
      for (int s = g->var_ptr[v]; s < g->var_ptr[v + 1]; ++s) {
        int e = g->var_edge[s];
        FN(var_message)(g, unary, r, r_flags, v, s, q + g->msg_off[e]); /* r_flags all 0 here */
        q_sent[e] = 1;
      }
    }
  }
  for (int f = 0; f < g->F; ++f) {
    int e0 = g->factor_ptr[f], a = g->factor_ptr[f + 1] - e0;
    if ((a == 1 && start_messages <= 1) || start_messages == 2) {
      for (int j = 0; j < a; ++j) {
        /* on_start sees no q at all: r_flags is still all-zero here and stands in for q_flags */
        FN(factor_marginal)(g, tables, q, r_flags, f, j, mode_max,
                            r + g->msg_off[e0 + j]);
        r_sent[e0 + j] = 1;
      }
    }
  }
  for (int e = 0; e < g->E; ++e) {
    if (q_sent[e]) q_flags[e] = FLAG_RECV;
    if (r_sent[e]) r_flags[e] = FLAG_RECV;
  }
}

/* one cycle: reads (q, r, q_flags, r_flags), writes (q2, r2, qf2, rf2) */
static void FN(maxsum_cycle_into)(const fg_t *g, const REAL *tables, const REAL *unary, int mode_max,
                                  int damp_vars, int damp_factors, REAL lam, REAL oml, REAL stab,
                                  const REAL *q, const REAL *r, const uint8_t *q_flags, const uint8_t *r_flags,
                                  REAL *q2, REAL *r2, uint8_t *qf2, uint8_t *rf2, uint8_t *q_sent,
                                  uint8_t *r_sent, int32_t *value, REAL *value_cost) {
#pragma omp parallel
  {
    REAL cand[MAX_DOM];
#pragma omp for schedule(static)
    for (int f = 0; f < g->F; ++f) {
      int e0 = g->factor_ptr[f], a = g->factor_ptr[f + 1] - e0;
      for (int j = 0; j < a; ++j) {
        int e = e0 + j;
        int d = g->dom_size[g->edge_var[e]];
        FN(factor_marginal)(g, tables, q, q_flags, f, j, mode_max, cand);
        FN(damp_and_gate)(cand, r + g->msg_off[e], r2 + g->msg_off[e], d, r_flags[e], &rf2[e],
                          &r_sent[e], damp_factors, lam, oml, stab);
      }
    }
#pragma omp for schedule(static)
    for (int v = 0; v < g->V; ++v) {
      int k = g->var_ptr[v + 1] - g->var_ptr[v];
      if (k == 0) continue; /* isolated: never cycles, keeps its cycle-0 value */
      int d = g->dom_size[v];
      FN(select_value)(g, unary, r, r_flags, v, mode_max, &value[v], &value_cost[v]);
      for (int s = g->var_ptr[v]; s < g->var_ptr[v + 1]; ++s) {
        int e = g->var_edge[s];
        FN(var_message)(g, unary, r, r_flags, v, s, cand);
        FN(damp_and_gate)(cand, q + g->msg_off[e], q2 + g->msg_off[e], d, q_flags[e], &qf2[e],
                          &q_sent[e], damp_vars, lam, oml, stab);
      }
    }
  }
}

/* n synchronous cycles k >= 1 for every factor and variable (Jacobi: both sides read the state at the
 * end of cycle k-1, SynchronousComputationMixin, computations.py:633-642,755-788).  The scratch
 * buffers are allocated once per call and the two states swap roles between cycles, so a timed call
 * with n >> 1 measures the arithmetic, not malloc / memcpy. */
void FN(maxsum_oracle_steps)(const fg_t *g, const REAL *tables, const REAL *unary, int mode_max,
                             int damp_vars, int damp_factors, double damping, double stability,
                             REAL *q, REAL *r, uint8_t *q_flags, uint8_t *r_flags, uint8_t *q_sent,
                             uint8_t *r_sent, int32_t *value, REAL *value_cost, int n) {
  int64_t M = g->msg_off[g->E];
  REAL lam = (REAL)damping, oml = (REAL)(1.0 - damping), stab = (REAL)stability;
  REAL *q2 = (REAL *)malloc(sizeof(REAL) * (M ? M : 1));
  REAL *r2 = (REAL *)malloc(sizeof(REAL) * (M ? M : 1));
  uint8_t *qf2 = (uint8_t *)malloc(g->E ? g->E : 1);
  uint8_t *rf2 = (uint8_t *)malloc(g->E ? g->E : 1);
  REAL *qa = q, *ra = r, *qb = q2, *rb = r2;
  uint8_t *qfa = q_flags, *rfa = r_flags, *qfb = qf2, *rfb = rf2;
  for (int i = 0; i < n; ++i) {
    FN(maxsum_cycle_into)(g, tables, unary, mode_max, damp_vars, damp_factors, lam, oml, stab, qa, ra, qfa,
                          rfa, qb, rb, qfb, rfb, q_sent, r_sent, value, value_cost);
    REAL *t;
    uint8_t *u;
    t = qa; qa = qb; qb = t;
    t = ra; ra = rb; rb = t;
    u = qfa; qfa = qfb; qfb = u;
    u = rfa; rfa = rfb; rfb = u;
  }
  if (qa != q) { /* odd n: the current state sits in the scratch buffers */
    memcpy(q, qa, sizeof(REAL) * M);
    memcpy(r, ra, sizeof(REAL) * M);
    memcpy(q_flags, qfa, g->E);
    memcpy(r_flags, rfa, g->E);
  }
  free(q2); free(r2); free(qf2); free(rf2);
}

void FN(maxsum_oracle_step)(const fg_t *g, const REAL *tables, const REAL *unary, int mode_max,
                            int damp_vars, int damp_factors, double damping, double stability,
                            REAL *q, REAL *r, uint8_t *q_flags, uint8_t *r_flags, uint8_t *q_sent,
                            uint8_t *r_sent, int32_t *value, REAL *value_cost) {
  FN(maxsum_oracle_steps)(g, tables, unary, mode_max, damp_vars, damp_factors, damping, stability, q, r,
                          q_flags, r_flags, q_sent, r_sent, value, value_cost, 1);
}

/* ------------------------------------------------------------------------------------------
 * DSA (dsa.py:277-431).  Constraints use the same arrays as factors (factor_ptr / edge_var /
 * table_off); var_ptr / var_edge list, per variable, its incident EDGES (constraint = factor of
 * that edge) in node.constraints order (dsa.py:255).
 * ------------------------------------------------------------------------------------------ */

/* find_optimum, relations.py:1367-1400 */
void FN(dsa_oracle_constraint_optima)(const fg_t *g, const REAL *tables, int mode_max, REAL *opt) {
  for (int f = 0; f < g->F; ++f) {
    const REAL *T = tables + g->table_off[f];
    int64_t n = g->table_off[f + 1] - g->table_off[f];
    REAL o = T[0];
    for (int64_t i = 1; i < n; ++i)
      if (mode_max ? (T[i] > o) : (T[i] < o)) o = T[i];
    opt[f] = o;
  }
}

/* value of constraint `f` with variable at scope position j set to xj, others from `val` */
static REAL FN(con_value)(const fg_t *g, const REAL *tables, const int32_t *val, int f, int j,
                          int xj) {
  int e0 = g->factor_ptr[f], a = g->factor_ptr[f + 1] - e0;
  int64_t idx = 0, s = 1;
  for (int i = a - 1; i >= 0; --i) {
    int u = g->edge_var[e0 + i];
    int xi = (i == j) ? xj : val[u];
    idx += xi * s;
    s *= g->dom_size[u];
  }
  return tables[g->table_off[f] + idx];
}

/* One evaluate_cycle (dsa.py:320-357) for every variable with >=1 neighbour.
 * variant: 0 A, 1 B, 2 C.  prob[v] is the per-variable threshold (p_mode, dsa.py:257-263).
 * edge_fac[e] = factor owning edge e.  has_nbr[v] = variable has at least one neighbour. */
/* var_cost != NULL: A-DSA (adsa.py:344-377 find_best_values adds the variable's own cost to every candidate;
 * current_cost, adsa.py:262, does not).  NULL: DSA (the variable-cost branch of find_optimal is dead code). */
void FN(dsa_oracle_step)(const fg_t *g, const REAL *tables, const int32_t *edge_fac,
                         const uint8_t *has_nbr, const REAL *con_opt, const double *prob,
                         int mode_max, int variant, uint64_t seed, uint32_t cycle,
                         const int32_t *var_id, const int32_t *val, int32_t *val_next,
                         REAL *val_cost, const REAL *var_cost) {
#pragma omp parallel
  {
    REAL cost[MAX_DOM];
    int best[MAX_DOM];
#pragma omp for schedule(static)
    for (int v = 0; v < g->V; ++v) {
      val_next[v] = val[v];
      if (!has_nbr[v]) continue;
      int d = g->dom_size[v];
      int cur = val[v];
      /* find_optimal, relations.py:1594-1638 (variable cost branch is dead code, :1630) */
      int nbest = 0;
      REAL best_cost = mode_max ? -(REAL)INFINITY : (REAL)INFINITY;
      for (int x = 0; x < d; ++x) {
        REAL c = (REAL)0;
        for (int s = g->var_ptr[v]; s < g->var_ptr[v + 1]; ++s) {
          int e = g->var_edge[s];
          int f = edge_fac[e];
          c += FN(con_value)(g, tables, val, f, e - g->factor_ptr[f], x);
        }
        cost[x] = c;
        if (var_cost) c += var_cost[g->unary_off[v] + x];
        if (c == best_cost) {
          best[nbest++] = x;
        } else if (mode_max ? (c > best_cost) : (c < best_cost)) {
          best_cost = c;
          nbest = 0;
          best[nbest++] = x;
        }
      }
      REAL delta = FABS(cost[cur] - best_cost);
      int attempt = 0;
      if (delta > (REAL)0) {
        attempt = 1;
      } else if (delta == (REAL)0) {
        int go = 0;
        if (variant == 2) {
          go = 1;
        } else if (variant == 1) { /* exists_violated_constraint, dsa.py:419-431 */
          for (int s = g->var_ptr[v]; s < g->var_ptr[v + 1] && !go; ++s) {
            int e = g->var_edge[s];
            int f = edge_fac[e];
            if (FN(con_value)(g, tables, val, f, e - g->factor_ptr[f], cur) != con_opt[f]) go = 1;
          }
        }
        if (go) {
          attempt = 1;
          if (nbest > 1) { /* best_values.remove(current_value) */
            int w = 0;
            for (int i = 0; i < nbest; ++i)
              if (best[i] != cur) best[w++] = best[i];
            nbest = w;
          }
        }
      }
      if (attempt) { /* probabilistic_change, dsa.py:407-417 */
        uint32_t b[4];
        philox4x32_10((uint32_t)(var_id ? var_id[v] : v), cycle, 0u, 0u, (uint32_t)seed,
                      (uint32_t)(seed >> 32), b);
        double u = ((double)(b[0] >> 5) * 67108864.0 + (double)(b[1] >> 6)) / 9007199254740992.0;
        if (prob[v] > u) {
          val_next[v] = best[(int)(((uint64_t)b[2] * (uint64_t)nbest) >> 32)];
          val_cost[v] = best_cost;
        }
      }
    }
  }
}

/* ---------------------------------------------------------------------------------------------
 * MGM (pydcop/algorithms/mgm.py).  One cycle = mgm_oracle_gain (every variable's best local gain,
 * :343-397,434-455) then mgm_oracle_decide (who moves, :497-537,574-591).
 *   nbr_ptr / nbr_idx   distinct neighbours of each variable (mgm.py:245-252)
 *   var_rank            position of the variable's name in sorted order (lexic tie break)
 *   cost / has_cost     current_cost, None until the first round (:349)
 * Quirks kept on purpose: the variable's own cost enters at its CURRENT value, not at the
 * candidate (:449); current_cost is computed once and afterwards only updated by the variable's
 * own moves (:349,518), so it goes stale when neighbours move; in max mode the comparison of gains
 * is the same `>` as in min mode (:516) although improving gains are negative; break_mode
 * 'random' is dead code (`self.break_mode == random` compares with the module, :541).
 * The reference adds the concerned variables' costs in the iteration order of a Python set
 * (:360-366,446-452); here: own cost first, then the neighbours in nbr order. */
void FN(mgm_oracle_gain)(const fg_t *g, const REAL *tables, const REAL *unary,
                         const int32_t *edge_fac, const uint8_t *has_nbr, const int32_t *nbr_ptr,
                         const int32_t *nbr_idx, int mode_max, uint64_t seed, uint32_t cycle,
                         const int32_t *val, REAL *cost, uint8_t *has_cost, REAL *gain,
                         int32_t *new_val) {
#pragma omp parallel
  {
    int best[MAX_DOM];
#pragma omp for schedule(static)
    for (int v = 0; v < g->V; ++v) {
      if (!has_nbr[v]) continue;
      int d = g->dom_size[v], cur = val[v];
      int s0 = g->var_ptr[v], s1 = g->var_ptr[v + 1];
      /* costs of the concerned variables at their current values */
      REAL others = unary[g->unary_off[v] + cur];
      if (!has_cost[v]) { /* first round: current_cost, mgm.py:349-368 */
        REAL c = (REAL)0;
        for (int s = s0; s < s1; ++s) {
          int e = g->var_edge[s], f = edge_fac[e];
          REAL t = FN(con_value)(g, tables, val, f, e - g->factor_ptr[f], cur);
          c = (s == s0) ? t : c + t;
        }
        c += unary[g->unary_off[v] + cur];
        for (int i = nbr_ptr[v]; i < nbr_ptr[v + 1]; ++i) {
          int u = nbr_idx[i];
          c += unary[g->unary_off[u] + val[u]];
        }
        cost[v] = c;
        has_cost[v] = 1;
      }
      /* find_arg_optimal, relations.py:1554-1591 */
      REAL best_val = mode_max ? (REAL)-2147483648.0 : (REAL)2147483647.0;
      int nbest = 0;
      for (int x = 0; x < d; ++x) {
        REAL c = (REAL)0;
        for (int s = s0; s < s1; ++s) {
          int e = g->var_edge[s], f = edge_fac[e];
          REAL t = FN(con_value)(g, tables, val, f, e - g->factor_ptr[f], x);
          c = (s == s0) ? t : c + t;
        }
        if (mode_max ? (best_val < c) : (best_val > c)) {
          best_val = c;
          nbest = 0;
          best[nbest++] = x;
        } else if (c == best_val) {
          best[nbest++] = x;
        }
      }
      REAL val_cost = best_val + others; /* own cost at the CURRENT value, mgm.py:449 */
      for (int i = nbr_ptr[v]; i < nbr_ptr[v + 1]; ++i) {
        int u = nbr_idx[i];
        val_cost += unary[g->unary_off[u] + val[u]];
      }
      REAL gn = cost[v] - val_cost;
      gain[v] = gn;
      if (mode_max ? (gn < (REAL)0) : (gn > (REAL)0)) { /* mgm.py:382-385 */
        uint32_t b[4];
        philox4x32_10((uint32_t)v, cycle, 0u, 0u, (uint32_t)seed, (uint32_t)(seed >> 32), b);
        new_val[v] = best[(int)(((uint64_t)b[2] * (uint64_t)nbest) >> 32)];
      } else {
        new_val[v] = cur;
      }
    }
  }
}

void FN(mgm_oracle_decide)(const fg_t *g, const uint8_t *has_nbr, const int32_t *nbr_ptr,
                           const int32_t *nbr_idx, const int32_t *var_rank, const REAL *gain,
                           const int32_t *new_val, int32_t *val, REAL *cost) {
#pragma omp parallel for schedule(static)
  for (int v = 0; v < g->V; ++v) {
    if (!has_nbr[v]) continue;
    REAL mx = gain[nbr_idx[nbr_ptr[v]]];
    for (int i = nbr_ptr[v] + 1; i < nbr_ptr[v + 1]; ++i)
      if (gain[nbr_idx[i]] > mx) mx = gain[nbr_idx[i]];
    int move = 0;
    if (gain[v] > mx) {
      move = 1;
    } else if (gain[v] == mx) { /* lexic order of the names, mgm.py:574-583 */
      move = 1;
      for (int i = nbr_ptr[v]; i < nbr_ptr[v + 1]; ++i) {
        int u = nbr_idx[i];
        if (gain[u] == mx && var_rank[u] < var_rank[v]) move = 0;
      }
    }
    if (move) { /* value_selection(new_value, current_cost - gain), mgm.py:518 */
      /* val is read by no other variable in this phase: in-place update is safe */
      cost[v] = cost[v] - gain[v];
    }
    if (move) val[v] = new_val[v];
  }
}

#undef FN
#undef CAT
#undef CAT_
