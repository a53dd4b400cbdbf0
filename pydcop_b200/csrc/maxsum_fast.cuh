// Shape-specialised MaxSum kernels for uniform-domain graphs (compile-time arity A and domain D).
//
//  k_f2v_tile<T,A,D>  factor -> variable.  One CTA = one tile of NF same-shaped factors.
//      HBM side: the tile's cost tables, previous r rows and the produced r rows are CONTIGUOUS
//      (class-major layout) and move with 1-D bulk async copies (TMA, mbarrier completion) when no
//      padding is needed, otherwise with 16-byte cp.async into a bank-conflict-free padded layout;
//      q rows are gathered through edge_qoff with cp.async.  All global traffic is issued as whole
//      16-byte (or row-sized) async transactions; the math runs out of shared memory with one
//      thread per directed edge, lanes mapped (position j, factor f) so a warp is uniform in j.
//  k_v2f_tile<T,D>    variable -> factor (+ select_value).  One CTA = 256 consecutive slots
//      (variable-major).  r rows are gathered through slot_roff into a TRANSPOSED shared tile,
//      q_old / q_next tiles are contiguous (slot order) and move with bulk async copies.
//
// Floating-point operand order is the reference's (see common.cuh); results are bit-identical to
// the generic kernels and to the CPU oracle of the same precision.
#pragma once
#include <vector>

#include "common.cuh"
#include "tma.cuh"

// ------------------------------------------------------------------------------------------------
// small helpers
// ------------------------------------------------------------------------------------------------
constexpr int fg_ipow(int b, int e) { return e == 0 ? 1 : b * fg_ipow(b, e - 1); }
constexpr int fg_gcd(int a, int b) { return b == 0 ? a : fg_gcd(b, a % b); }
constexpr int fg_clamp(int x, int lo, int hi) { return x < lo ? lo : (x > hi ? hi : x); }

template <typename T, int N> struct VecT;
template <> struct VecT<float, 4> { using type = float4; };
template <> struct VecT<float, 2> { using type = float2; };
template <> struct VecT<float, 1> { using type = float; };
template <> struct VecT<double, 2> { using type = double2; };
template <> struct VecT<double, 1> { using type = double; };

// load N contiguous elements with vector width V (V | N), static indexing only
template <typename T, int N, int V>
__device__ __forceinline__ void ld_row(const T *__restrict__ p, T (&out)[N]) {
  using VT = typename VecT<T, V>::type;
#pragma unroll
  for (int i = 0; i < N / V; ++i) {
    VT v = reinterpret_cast<const VT *>(p)[i];
    const T *e = reinterpret_cast<const T *>(&v);
#pragma unroll
    for (int k = 0; k < V; ++k) out[i * V + k] = e[k];
  }
}
template <typename T, int N, int V>
__device__ __forceinline__ void st_row(T *__restrict__ p, const T (&in)[N]) {
  using VT = typename VecT<T, V>::type;
#pragma unroll
  for (int i = 0; i < N / V; ++i) {
    VT v;
    T *e = reinterpret_cast<T *>(&v);
#pragma unroll
    for (int k = 0; k < V; ++k) e[k] = in[i * V + k];
    reinterpret_cast<VT *>(p)[i] = v;
  }
}

// async copy of `bytes` (4, 8 or 16) global -> shared
template <int BYTES> __device__ __forceinline__ void cp_async_b(void *dst, const void *src);
template <> __device__ __forceinline__ void cp_async_b<16>(void *d, const void *s) { cp_async_16(d, s); }
template <> __device__ __forceinline__ void cp_async_b<8>(void *d, const void *s) { cp_async_8(d, s); }
template <> __device__ __forceinline__ void cp_async_b<4>(void *d, const void *s) { cp_async_4(d, s); }

// cooperative contiguous copy global -> shared (both 16-B aligned at the start)
template <typename T>
__device__ __forceinline__ void coop_copy_in(T *dst, const T *__restrict__ src, int n, int tid, int nt) {
  constexpr int E16 = 16 / (int)sizeof(T);
  const int nv = n / E16;
  for (int i = tid; i < nv; i += nt) cp_async_16(dst + i * E16, src + i * E16);
  for (int i = nv * E16 + tid; i < n; i += nt) cp_async_b<(int)sizeof(T)>(dst + i, src + i);
}
// cooperative contiguous copy shared -> global
template <typename T>
__device__ __forceinline__ void coop_copy_out(T *__restrict__ dst, const T *src, int n, int tid, int nt) {
  constexpr int E16 = 16 / (int)sizeof(T);
  using V16 = typename VecT<T, E16>::type;
  const int nv = n / E16;
  for (int i = tid; i < nv; i += nt) reinterpret_cast<V16 *>(dst)[i] = reinterpret_cast<const V16 *>(src)[i];
  for (int i = nv * E16 + tid; i < n; i += nt) dst[i] = src[i];
}

template <typename T> struct MatchEps;
template <> struct MatchEps<float> { __device__ static float lo() { return 1.0f - 9.5367431640625e-07f; } __device__ static float hi() { return 1.0f + 9.5367431640625e-07f; } __device__ static float tiny() { return 1e-30f; } };
template <> struct MatchEps<double> { __device__ static double lo() { return 1.0 - 3.552713678800501e-15; } __device__ static double hi() { return 1.0 + 3.552713678800501e-15; } __device__ static double tiny() { return 1e-290; } };

// Same predicate as approx_match1 (maxsum.py:688-710), bit-for-bit, but the IEEE division is only
// executed when 2*delta is within a few ulps of stab*|s|; otherwise the comparison of the products
// decides (the margins are wider than the worst-case rounding of the quotient).
template <typename T>
__device__ __forceinline__ bool approx_match_fast(T c, T prev_c, T stab) {
  if (prev_c == c) return true;
  const T s = prev_c + c;
  if (s == (T)0) return false;
  const T d2 = (T)2 * fg_abs<T>(prev_c - c);
  const T as = fg_abs<T>(s);
  const T rhs = stab * as;
  if (rhs > MatchEps<T>::tiny() && rhs < Inf<T>::pos()) {
    if (d2 < rhs * MatchEps<T>::lo()) return true;
    if (d2 > rhs * MatchEps<T>::hi()) return false;
  }
  return (d2 / as) < stab;
}

// damping + approx_match + gate for one message row held in registers.
// cand: fresh message (in) / message to store (out).  Returns `sent`.
template <typename T, int D>
__device__ __forceinline__ bool damp_gate_row(T (&cand)[D], const T (&prev)[D], uint8_t &cnt, bool damp_side,
                                              T lam, T oml, T stab) {
  const bool has_prev = cnt & 1;
  bool match = has_prev;
  if (has_prev) {
#pragma unroll
    for (int x = 0; x < D; ++x) {
      T c = cand[x];
      if (damp_side) c = lam * prev[x] + oml * c;
      cand[x] = c;
      if (!approx_match_fast<T>(c, prev[x], stab)) match = false;
    }
  }
  const bool sent = gate_decide(match, cnt);
  if (!sent) {
#pragma unroll
    for (int x = 0; x < D; ++x) cand[x] = prev[x];
  }
  return sent;
}

template <typename T> __device__ __forceinline__ T fg_opt(T a, T b, bool mx);
template <> __device__ __forceinline__ float fg_opt<float>(float a, float b, bool mx) { return mx ? fmaxf(a, b) : fminf(a, b); }
template <> __device__ __forceinline__ double fg_opt<double>(double a, double b, bool mx) { return mx ? fmax(a, b) : fmin(a, b); }

// ------------------------------------------------------------------------------------------------
// factor -> variable
// ------------------------------------------------------------------------------------------------
template <typename T, int A, int D>
struct F2VCfg {
  static constexpr int S = fg_ipow(D, A);
  static constexpr int R = A * D;
  static constexpr int INNER = S / D;
  static constexpr int E16 = 16 / (int)sizeof(T);
  static constexpr bool QUADS = (S * (int)sizeof(T)) % 16 == 0;
  static constexpr int Q = (S * (int)sizeof(T)) / 16;
  static constexpr int PAD = (QUADS && (Q % 2 == 0)) ? E16 : 0;  // odd 16-B stride: no bank conflicts
  static constexpr int SP = S + PAD;
  static constexpr int PER_FACTOR = (SP + 3 * R) * (int)sizeof(T);
  static constexpr int NF = fg_clamp(((40 * 1024) / PER_FACTOR) / 32 * 32, 32, 512);
  static constexpr int NT = fg_clamp(NF * A, 64, 256);
  // vector width (elements) for table rows (D contiguous elements at multiples of D inside a
  // factor whose stride is SP) and for message rows
  static constexpr int VT_BYTES = fg_gcd(16, fg_gcd(D * (int)sizeof(T), SP * (int)sizeof(T)));
  static constexpr int VT = VT_BYTES / (int)sizeof(T);
  static constexpr int VR_BYTES = fg_gcd(16, D * (int)sizeof(T));
  static constexpr int VR = VR_BYTES / (int)sizeof(T);
  static constexpr size_t SMEM = (size_t)(NF * SP + 3 * NF * R) * sizeof(T) + 16;
};

template <typename T, int A, int D, typename OffT>
__global__ void __launch_bounds__(F2VCfg<T, A, D>::NT)
k_f2v_tile(const fg_class_t c, const T *__restrict__ tables, const T *__restrict__ q_cur,
           const T *__restrict__ r_cur, T *__restrict__ r_next, const OffT *__restrict__ edge_qoff,
           uint8_t *__restrict__ r_cnt, uint8_t *__restrict__ r_sent, MaxSumParams p) {
  using C = F2VCfg<T, A, D>;
  constexpr int S = C::S, R = C::R, SP = C::SP, NF = C::NF, NT = C::NT, INNER = C::INNER;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  T *tab = reinterpret_cast<T *>(smem_raw);
  T *qt = tab + NF * SP;   // q rows of the tile's edges  [f][j][x]
  T *rt = qt + NF * R;     // previous r rows (contiguous copy of the class-major array)
  T *ot = rt + NF * R;     // produced r rows
  uint64_t *bar = reinterpret_cast<uint64_t *>(ot + NF * R);

  const int tid = threadIdx.x;
  const int f0 = blockIdx.x * NF;
  const int nf = min(NF, c.n_factors - f0);
  const T *gtab = tables + c.table_base + (int64_t)f0 * S;
  const int64_t rbase = c.msg_base + (int64_t)f0 * R;
  const int e_base = c.first_edge + f0 * A;

  const bool full = (nf == NF);
  const bool tma_tab = full && C::PAD == 0 && ((NF * S * (int)sizeof(T)) % 16 == 0);
  const bool tma_rows = full && ((NF * R * (int)sizeof(T)) % 16 == 0);
  if (tid == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
  }
  __syncthreads();
  if (tid == 0) {
    uint32_t bytes = 0;
    if (tma_tab) bytes += NF * S * (uint32_t)sizeof(T);
    if (tma_rows) bytes += NF * R * (uint32_t)sizeof(T);
    mbar_expect_tx(bar, bytes);  // bytes == 0: plain arrival completes the phase
    if (tma_tab) tma_load_1d(tab, gtab, NF * S * (uint32_t)sizeof(T), bar);
    if (tma_rows) tma_load_1d(rt, r_cur + rbase, NF * R * (uint32_t)sizeof(T), bar);
  }
  if (!tma_tab) {
    if constexpr (C::QUADS) {  // 16-byte units into the (possibly padded) per-factor stride
      constexpr int Q = C::Q, E16 = C::E16;
      for (int u = tid; u < nf * Q; u += NT) {
        const int f = u / Q, w = u - f * Q;
        cp_async_16(tab + f * SP + w * E16, gtab + (int64_t)f * S + w * E16);
      }
    } else {
      for (int i = tid; i < nf * S; i += NT) cp_async_b<(int)sizeof(T)>(tab + i, gtab + i);
    }
  }
  if (!tma_rows) coop_copy_in<T>(rt, r_cur + rbase, nf * R, tid, NT);
  // gather the q rows: thread le <-> edge (f, j) = (le / A, le % A) so edge_qoff reads are coalesced
  for (int le = tid; le < nf * A; le += NT) {
    const int64_t off = (int64_t)edge_qoff[e_base + le];
    const T *src = q_cur + off;
    T *dst = qt + le * D;
#pragma unroll
    for (int i = 0; i < D / C::VR; ++i) cp_async_b<C::VR_BYTES>(dst + i * C::VR, src + i * C::VR);
  }
  cp_async_wait_all();
  mbar_wait(bar, 0);
  __syncthreads();

  const bool mx = p.mode_max != 0;
  const T lam = (T)p.damping, oml = (T)p.one_minus_damping, stab = (T)p.stability;
  const T init = mx ? -Inf<T>::pos() : Inf<T>::pos();

  for (int le = tid; le < NF * A; le += NT) {
    const int j = le / NF, f = le - j * NF;  // warp-uniform j (NF % 32 == 0)
    if (f >= nf) continue;
    const T *tf = tab + f * SP;
    const T *qf = qt + f * R;
    T *of = ot + f * R + j * D;
    T cand[D];
    if (A == 1) {
      ld_row<T, D, C::VT>(tf, cand);
    } else if (A == 2) {
      if (j == 0) {
        T qo[D];
        ld_row<T, D, C::VR>(qf + D, qo);
#pragma unroll
        for (int x0 = 0; x0 < D; ++x0) {
          T row[D];
          ld_row<T, D, C::VT>(tf + x0 * D, row);
          T o = init;
#pragma unroll
          for (int x1 = 0; x1 < D; ++x1) o = fg_opt<T>(o, row[x1] + qo[x1], mx);
          cand[x0] = o;
        }
      } else {
#pragma unroll
        for (int x1 = 0; x1 < D; ++x1) cand[x1] = init;
        T q0r[D];
        ld_row<T, D, C::VR>(qf, q0r);
#pragma unroll
        for (int x0 = 0; x0 < D; ++x0) {
          T row[D];
          ld_row<T, D, C::VT>(tf + x0 * D, row);
#pragma unroll
          for (int x1 = 0; x1 < D; ++x1) cand[x1] = fg_opt<T>(cand[x1], row[x1] + q0r[x0], mx);
        }
      }
    } else {  // A == 3, table[x0][x1][x2]; sum of the two other rows in position order
      if (j == 0) {
        T q1[D], q2[D];
        ld_row<T, D, C::VR>(qf + D, q1);
        ld_row<T, D, C::VR>(qf + 2 * D, q2);
#pragma unroll 1
        for (int x0 = 0; x0 < D; ++x0) {
          T o = init;
#pragma unroll
          for (int x1 = 0; x1 < D; ++x1) {
            T row[D];
            ld_row<T, D, C::VT>(tf + x0 * INNER + x1 * D, row);
#pragma unroll
            for (int x2 = 0; x2 < D; ++x2) o = fg_opt<T>(o, row[x2] + (q1[x1] + q2[x2]), mx);
          }
          of[x0] = o;
        }
        ld_row<T, D, C::VR>(of, cand);
      } else {
#pragma unroll
        for (int x = 0; x < D; ++x) cand[x] = init;
        T qk[D];  // the other non-leading row: position 2 when j == 1, position 1 when j == 2
        ld_row<T, D, C::VR>(qf + (j == 1 ? 2 * D : D), qk);
#pragma unroll 1
        for (int x0 = 0; x0 < D; ++x0) {
          const T q0 = qf[x0];
          T s[D];
#pragma unroll
          for (int x = 0; x < D; ++x) s[x] = q0 + qk[x];
#pragma unroll
          for (int x1 = 0; x1 < D; ++x1) {
            T row[D];
            ld_row<T, D, C::VT>(tf + x0 * INNER + x1 * D, row);
            if (j == 1) {
#pragma unroll
              for (int x2 = 0; x2 < D; ++x2) cand[x1] = fg_opt<T>(cand[x1], row[x2] + s[x2], mx);
            } else {
#pragma unroll
              for (int x2 = 0; x2 < D; ++x2) cand[x2] = fg_opt<T>(cand[x2], row[x2] + s[x1], mx);
            }
          }
        }
      }
    }
    // damping, send gate, state update
    T prev[D];
    ld_row<T, D, C::VR>(rt + f * R + j * D, prev);
    const int e = e_base + f * A + j;
    uint8_t cnt = r_cnt[e];
    const bool sent = damp_gate_row<T, D>(cand, prev, cnt, p.damp_factors != 0, lam, oml, stab);
    st_row<T, D, C::VR>(of, cand);
    r_cnt[e] = cnt;
    if (r_sent) r_sent[e] = sent ? 1 : 0;
  }
  // publish the tile
  if (tma_rows) {
    fence_proxy_async_smem();
    __syncthreads();
    if (tid == 0) {
      tma_store_1d(r_next + rbase, ot, NF * R * (uint32_t)sizeof(T));
      tma_store_commit();
      tma_store_wait_read();
    }
  } else {
    __syncthreads();
    coop_copy_out<T>(r_next + rbase, ot, nf * R, tid, NT);
  }
}

// ------------------------------------------------------------------------------------------------
// variable -> factor
// ------------------------------------------------------------------------------------------------
#define FG_V2F_TS 256       // slots per tile
#define FG_V2F_MAXDEG 32    // fast path requires max_degree <= this
#define FG_V2F_RS (FG_V2F_TS + 2 * FG_V2F_MAXDEG)

template <typename T, int D>
struct V2FCfg {
  static constexpr int VR_BYTES = fg_gcd(16, D * (int)sizeof(T));
  static constexpr int VR = VR_BYTES / (int)sizeof(T);
  static constexpr size_t SMEM = (size_t)(D * FG_V2F_RS + 2 * FG_V2F_TS * D) * sizeof(T) + 16;
};

template <typename T, int D, typename OffT>
__global__ void __launch_bounds__(FG_V2F_TS)
k_v2f_tile(int n_slots, const int32_t *__restrict__ var_ptr, const int32_t *__restrict__ slot_var,
           const OffT *__restrict__ slot_roff, const T *__restrict__ unary, const T *__restrict__ r_cur,
           const T *__restrict__ q_cur, T *__restrict__ q_next, uint8_t *__restrict__ q_cnt,
           uint8_t *__restrict__ q_sent, int32_t *__restrict__ value, T *__restrict__ value_cost,
           MaxSumParams p) {
  using C = V2FCfg<T, D>;
  constexpr int TS = FG_V2F_TS, RS = FG_V2F_RS;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  T *rtT = reinterpret_cast<T *>(smem_raw);  // [D][RS] transposed gathered r rows
  T *qin = rtT + D * RS;                     // [TS][D]
  T *qout = qin + TS * D;                    // [TS][D]
  uint64_t *bar = reinterpret_cast<uint64_t *>(qout + TS * D);

  const int tid = threadIdx.x;
  const int t0 = blockIdx.x * TS;
  const int n = min(TS, n_slots - t0);
  const bool full = (n == TS);
  const bool tma_q = full && ((TS * D * (int)sizeof(T)) % 16 == 0);
  if (tid == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
  }
  __syncthreads();
  if (tid == 0) {
    mbar_expect_tx(bar, tma_q ? TS * D * (uint32_t)sizeof(T) : 0u);
    if (tma_q) tma_load_1d(qin, q_cur + (int64_t)t0 * D, TS * D * (uint32_t)sizeof(T), bar);
  }
  if (!tma_q) coop_copy_in<T>(qin, q_cur + (int64_t)t0 * D, n * D, tid, TS);
  // rows of every variable touching this tile: [s_lo, s_hi)
  const int s_lo = var_ptr[slot_var[t0]];
  const int s_hi = var_ptr[slot_var[t0 + n - 1] + 1];
  for (int i = s_lo + tid; i < s_hi; i += TS) {
    const T *src = r_cur + (int64_t)slot_roff[i];
    T row[D];
    ld_row<T, D, C::VR>(src, row);
#pragma unroll
    for (int x = 0; x < D; ++x) rtT[x * RS + (i - s_lo)] = row[x];
  }
  cp_async_wait_all();
  mbar_wait(bar, 0);
  __syncthreads();

  if (tid < n) {
    const int s = t0 + tid;
    const int v = slot_var[s];
    const int s0 = var_ptr[v], s1 = var_ptr[v + 1];
    const bool mx = p.mode_max != 0;
    const T lam = (T)p.damping, oml = (T)p.one_minus_damping, stab = (T)p.stability;
    T un[D];
    ld_row<T, D, C::VR>(unary + (int64_t)v * D, un);
    const int a = s0 - s_lo, b = s1 - s_lo, me = s - s_lo;
    if (s == s0) {  // select_value (maxsum.py:584-620)
      int best = 0;
      T best_c = (T)0;
#pragma unroll
      for (int x = 0; x < D; ++x) {
        T cst = un[x];
        for (int t = a; t < b; ++t) cst += rtT[x * RS + t];
        if (x == 0 || (mx ? (cst > best_c) : (cst < best_c))) { best = x; best_c = cst; }
      }
      value[v] = best;
      value_cost[v] = best_c;
    }
    // costs_for_factor (maxsum.py:623-676), reference order: value-major, then factor
    T cand[D];
    T sum_cost = (T)0;
#pragma unroll
    for (int x = 0; x < D; ++x) {
      T m = un[x];
      for (int t = a; t < b; ++t) {
        if (t == me) continue;
        const T cst = rtT[x * RS + t];
        sum_cost += cst;
        m += cst;
      }
      cand[x] = m;
    }
    const T avg = sum_cost / (T)D;
#pragma unroll
    for (int x = 0; x < D; ++x) cand[x] = cand[x] - avg;
    T prev[D];
    ld_row<T, D, C::VR>(qin + tid * D, prev);
    uint8_t cnt = q_cnt[s];
    const bool sent = damp_gate_row<T, D>(cand, prev, cnt, p.damp_vars != 0, lam, oml, stab);
    st_row<T, D, C::VR>(qout + tid * D, cand);
    q_cnt[s] = cnt;
    if (q_sent) q_sent[s] = sent ? 1 : 0;
  }
  if (tma_q) {
    fence_proxy_async_smem();
    __syncthreads();
    if (tid == 0) {
      tma_store_1d(q_next + (int64_t)t0 * D, qout, TS * D * (uint32_t)sizeof(T));
      tma_store_commit();
      tma_store_wait_read();
    }
  } else {
    __syncthreads();
    coop_copy_out<T>(q_next + (int64_t)t0 * D, qout, n * D, tid, TS);
  }
}

// ------------------------------------------------------------------------------------------------
// host-side dispatch
// ------------------------------------------------------------------------------------------------
struct MaxSumFastPlan {
  bool v2f = false;               // uniform domain, supported D, max degree small enough
  bool off32 = false;             // 32-bit gather offsets available
  std::vector<uint8_t> f2v;       // per class: fast kernel available
  bool attrs_set = false;
};

#define FG_FAST_DOMS(X) X(2) X(3) X(4) X(5) X(6) X(8) X(10) X(16) X(20)
#define FG_FAST_DOMS_A3(X) X(2) X(3) X(4) X(5) X(8)

inline bool fg_fast_dom(int d) {
#define X(n) if (d == n) return true;
  FG_FAST_DOMS(X)
#undef X
  return false;
}
inline bool fg_fast_dom_a3(int d) {
#define X(n) if (d == n) return true;
  FG_FAST_DOMS_A3(X)
#undef X
  return false;
}

inline bool fg_fast_disabled() {
  const char *e = getenv("PYDCOP_B200_NO_FAST");
  return e && e[0] == '1';
}

inline void maxsum_fast_plan(const fg_maxsum_desc_t &d, const std::vector<fg_class_t> &classes, MaxSumFastPlan &plan) {
  plan.f2v.assign(classes.size(), 0);
  plan.off32 = d.dev_slot_roff32 != nullptr && d.dev_edge_qoff32 != nullptr;
  if (fg_fast_disabled()) return;
  for (size_t i = 0; i < classes.size(); ++i) {
    const fg_class_t &c = classes[i];
    bool uni = true;
    for (int j = 1; j < c.arity; ++j) uni = uni && c.dom[j] == c.dom[0];
    if (!uni) continue;
    if (c.arity <= 2 && fg_fast_dom(c.dom[0])) plan.f2v[i] = 1;
    if (c.arity == 3 && fg_fast_dom_a3(c.dom[0])) plan.f2v[i] = 1;
  }
  plan.v2f = d.uniform_dom > 0 && fg_fast_dom(d.uniform_dom) && d.max_degree <= FG_V2F_MAXDEG && d.n_edges > 0;
}

template <typename T, int A, int D, typename OffT>
inline void launch_f2v_tile(const fg_class_t &c, const fg_maxsum_desc_t &d, const T *q_cur, const T *r_cur, T *r_next,
                            const OffT *edge_qoff, const MaxSumParams &p, cudaStream_t st) {
  using C = F2VCfg<T, A, D>;
  auto kern = k_f2v_tile<T, A, D, OffT>;
  static bool attr_done = false;  // one per instantiation
  if (!attr_done) {
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM);
    attr_done = true;
  }
  const unsigned blocks = (unsigned)((c.n_factors + C::NF - 1) / C::NF);
  kern<<<blocks, C::NT, C::SMEM, st>>>(c, (const T *)d.dev_tables, q_cur, r_cur, r_next, edge_qoff, d.dev_r_cnt,
                                       d.dev_r_sent, p);
}

template <typename T, typename OffT>
inline bool dispatch_f2v(const fg_class_t &c, const fg_maxsum_desc_t &d, const T *q_cur, const T *r_cur, T *r_next,
                         const OffT *edge_qoff, const MaxSumParams &p, cudaStream_t st) {
  const int D = c.dom[0];
  switch (c.arity) {
    case 1:
      switch (D) {
#define X(n) case n: launch_f2v_tile<T, 1, n, OffT>(c, d, q_cur, r_cur, r_next, edge_qoff, p, st); return true;
        FG_FAST_DOMS(X)
#undef X
      }
      return false;
    case 2:
      switch (D) {
#define X(n) case n: launch_f2v_tile<T, 2, n, OffT>(c, d, q_cur, r_cur, r_next, edge_qoff, p, st); return true;
        FG_FAST_DOMS(X)
#undef X
      }
      return false;
    case 3:
      switch (D) {
#define X(n) case n: launch_f2v_tile<T, 3, n, OffT>(c, d, q_cur, r_cur, r_next, edge_qoff, p, st); return true;
        FG_FAST_DOMS_A3(X)
#undef X
      }
      return false;
  }
  return false;
}

template <typename T>
inline bool maxsum_fast_f2v(const MaxSumFastPlan &plan, int ci, const fg_class_t &c, const fg_maxsum_desc_t &d,
                            const T *q_cur, const T *r_cur, T *r_next, const MaxSumParams &p, cudaStream_t st,
                            int64_t &launches) {
  if (!plan.f2v[ci]) return false;
  bool ok = plan.off32 ? dispatch_f2v<T, uint32_t>(c, d, q_cur, r_cur, r_next, d.dev_edge_qoff32, p, st)
                       : dispatch_f2v<T, int64_t>(c, d, q_cur, r_cur, r_next, d.dev_edge_qoff, p, st);
  if (ok) ++launches;
  return ok;
}

template <typename T, int D, typename OffT>
inline void launch_v2f_tile(const fg_maxsum_desc_t &d, const T *r_cur, const T *q_cur, T *q_next,
                            const OffT *slot_roff, const MaxSumParams &p, cudaStream_t st) {
  using C = V2FCfg<T, D>;
  auto kern = k_v2f_tile<T, D, OffT>;
  static bool attr_done = false;
  if (!attr_done) {
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM);
    attr_done = true;
  }
  const unsigned blocks = (unsigned)((d.n_edges + FG_V2F_TS - 1) / FG_V2F_TS);
  kern<<<blocks, FG_V2F_TS, C::SMEM, st>>>(d.n_edges, d.dev_var_ptr, d.dev_slot_var, slot_roff, (const T *)d.dev_unary,
                                            r_cur, q_cur, q_next, d.dev_q_cnt, d.dev_q_sent, d.dev_value,
                                            (T *)d.dev_value_cost, p);
}

template <typename T, typename OffT>
inline bool dispatch_v2f(const fg_maxsum_desc_t &d, const T *r_cur, const T *q_cur, T *q_next, const OffT *slot_roff,
                         const MaxSumParams &p, cudaStream_t st) {
  switch (d.uniform_dom) {
#define X(n) case n: launch_v2f_tile<T, n, OffT>(d, r_cur, q_cur, q_next, slot_roff, p, st); return true;
    FG_FAST_DOMS(X)
#undef X
  }
  return false;
}

template <typename T>
inline bool maxsum_fast_v2f(const MaxSumFastPlan &plan, const fg_maxsum_desc_t &d, const T *r_cur, const T *q_cur,
                            T *q_next, const MaxSumParams &p, cudaStream_t st, int64_t &launches) {
  if (!plan.v2f) return false;
  bool ok = plan.off32 ? dispatch_v2f<T, uint32_t>(d, r_cur, q_cur, q_next, d.dev_slot_roff32, p, st)
                       : dispatch_v2f<T, int64_t>(d, r_cur, q_cur, q_next, d.dev_slot_roff, p, st);
  if (ok) ++launches;
  return ok;
}
