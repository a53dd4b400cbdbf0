/* CPU ORACLE — TEST INFRASTRUCTURE, NOT PRODUCT CODE.  See dcop_oracle_impl.h.
 * Build: make -C oracle   (gcc -O2 -ffp-contract=off -fopenmp -shared -fPIC) */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#define MAX_ARITY 8
#define MAX_DOM 256
#define SAME_COUNT 4 /* maxsum.py:106 */
#define FLAG_RECV 1u
#define FLAG_PREV 2u

typedef struct {
  int32_t V, F, E;
  const int32_t *dom_size;   /* [V] */
  const int32_t *factor_ptr; /* [F+1] edges of factor f = [factor_ptr[f], factor_ptr[f+1]) */
  const int32_t *edge_var;   /* [E] variable of edge e (scope order = table axis order) */
  const int64_t *table_off;  /* [F+1] */
  const int32_t *var_ptr;    /* [V+1] */
  const int32_t *var_edge;   /* [E] incident edges of v in `links` order */
  const int64_t *msg_off;    /* [E+1] message row of edge e = [msg_off[e], msg_off[e+1]) */
  const int64_t *unary_off;  /* [V+1] */
} fg_t;

/* Philox4x32-10; same draw definition as oracle/philox.py and csrc/philox.cuh */
static void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                          uint32_t k1, uint32_t out[4]) {
  for (int i = 0; i < 10; ++i) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* OpenMP threads the parallel loops use: set (n > 0) / query.  torchrun exports OMP_NUM_THREADS=1; a
 * timed baseline sets the count explicitly and reports what the runtime says. */
int oracle_threads(int set_to) {
#ifdef _OPENMP
  if (set_to > 0) omp_set_num_threads(set_to);
  return omp_get_max_threads();
#else
  (void)set_to;
  return 1;
#endif
}

void oracle_philox(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                   uint32_t out[4]) {
  philox4x32_10(c0, c1, c2, c3, k0, k1, out);
}

/* DSA on_start (dsa.py:277-295): random initial value for connected variables (injected draw,
 * cycle = 0xffffffff); isolated variables take argopt of (own cost, value) — tuple ordering of
 * relations.py:1641-1669 with value == domain index.  var_id (may be NULL = identity) is the
 * Philox counter of each variable: its id in the whole problem when `g` is one shard of it. */
void dsa_oracle_init(const fg_t *g, const double *unary, const uint8_t *has_nbr, int mode_max,
                     uint64_t seed, const int32_t *var_id, int32_t *val) {
  for (int v = 0; v < g->V; ++v) {
    int d = g->dom_size[v];
    if (has_nbr[v]) {
      uint32_t b[4];
      philox4x32_10((uint32_t)(var_id ? var_id[v] : v), 0xFFFFFFFFu, 0u, 0u, (uint32_t)seed, (uint32_t)(seed >> 32), b);
      val[v] = (int32_t)(((uint64_t)b[2] * (uint64_t)d) >> 32);
    } else {
      int best = 0;
      for (int x = 1; x < d; ++x) {
        double c = unary[g->unary_off[v] + x], bc = unary[g->unary_off[v] + best];
        if (mode_max ? (c >= bc) : (c < bc)) best = x;
      }
      val[v] = best;
    }
  }
}

/* MGM on_start (mgm.py:283-310): a variable without neighbours takes argopt of (own cost, value)
 * (optimal_cost_value, relations.py:1641-1669) and is done; the others take their initial_value or an
 * injected random.choice(domain) (draw keyed cycle = 0xffffffff). */
void mgm_oracle_init(const fg_t *g, const double *unary, const uint8_t *has_nbr,
                     const int32_t *init_value, int mode_max, uint64_t seed, int32_t *val) {
  for (int v = 0; v < g->V; ++v) {
    int d = g->dom_size[v];
    if (has_nbr[v]) {
      if (init_value && init_value[v] >= 0) {
        val[v] = init_value[v];
      } else {
        uint32_t b[4];
        philox4x32_10((uint32_t)v, 0xFFFFFFFFu, 0u, 0u, (uint32_t)seed, (uint32_t)(seed >> 32), b);
        val[v] = (int32_t)(((uint64_t)b[2] * (uint64_t)d) >> 32);
      }
    } else {
      int best = 0;
      for (int x = 1; x < d; ++x) {
        double c = unary[g->unary_off[v] + x], bc = unary[g->unary_off[v] + best];
        if (mode_max ? (c >= bc) : (c < bc)) best = x;
      }
      val[v] = best;
    }
  }
}

#define REAL double
#define SUFFIX _f64
#define FABS fabs
#include "dcop_oracle_impl.h"
#undef REAL
#undef SUFFIX
#undef FABS

#define REAL float
#define SUFFIX _f32
#define FABS fabsf
#include "dcop_oracle_impl.h"
#undef REAL
#undef SUFFIX
#undef FABS
