// Factor side for the shapes the compile-time kernels do not cover: mixed domain sizes, domain sizes
// outside the compiled set, arity up to FG_MAX_ARITY.  factor_costs_for_var (maxsum.py:382-447) +
// the factor's on_new_cycle (maxsum.py:339-379) with RUNTIME dimensions and compile-time arity.
//
// The one-thread-per-edge kernel (maxsum_generic.cuh) reads each table `arity` times with a stride of one
// table between neighbouring threads — 3.5 % of the HBM roofline on C2 (profiles/r01_bench_v0_generic_kernels.json).
// Here a CTA owns a run of consecutive factors of one class: their tables, their previous messages and their
// output rows are CONTIGUOUS in the class-major layout, so every byte crosses HBM once, in coalesced
// requests; only the incoming q rows are gathered (one row per edge, via edge_qoff).  Work is then split by
// OUTPUT VALUE — one item = (factor, scope position j, value x_j), a minimum over the table slice —
// so a class of few large tables keeps a whole CTA busy as well as a class of many small ones.
//
// Message rows may be padded (fg_class_t::row_off is aligned, row_total >= sum(dom)): padding elements are carried over.
// Arithmetic is the generic kernel's, operand for operand: sum = 0, += q_i[x_i] for i != j in scope
// order, table + sum, strict optimum; damping / approx_match / send gate per edge.  Bit-identical to
// k_f2v_generic and to the oracle (tests/test_gpu_tiled_rt.py).
#pragma once
#include <algorithm>

#include "common.cuh"

struct TiledRtCfg {
  int nf_tile = 0;   // factors per CTA (0: the class does not fit, use the generic kernel)
  int sp = 0;        // table stride in shared memory (odd: neighbouring factors start in different banks)
  size_t smem = 0;
};

// one CTA's work: `nf` factors of class `cls` from factor `f0` on
struct RtTile { int32_t cls, f0, nf, sp; };

constexpr int FG_TILED_RT_THREADS = 256;
constexpr size_t FG_TILED_RT_SMEM = 72 * 1024;   // 3 CTAs per SM

inline TiledRtCfg tiled_rt_cfg(const fg_class_t &c, size_t elem) {
  TiledRtCfg r;
  const int64_t S = c.table_size, RT = c.row_total;
  if (S <= 0 || S > (1 << 20)) return r;
  const int64_t sp = S | 1;
  const int64_t per = (sp + 3 * RT) * (int64_t)elem;
  int64_t fit = (int64_t)FG_TILED_RT_SMEM / per;
  if (fit < 1) return r;
  // enough items (nf * RT output values) for every thread, enough CTAs for every SM
  int64_t want = std::max<int64_t>(1, (2 * FG_TILED_RT_THREADS + RT - 1) / RT);
  int64_t spread = std::max<int64_t>(1, c.n_factors / (148 * 6));
  int64_t nf = std::min(fit, std::max(want, std::min<int64_t>(spread, 4 * want)));
  r.nf_tile = (int)std::max<int64_t>(1, nf);
  r.sp = (int)sp;
  r.smem = (size_t)(r.nf_tile * per);
  return r;
}

// Launch plan: ONE launch per arity over every class of that arity (a mixed-domain problem has
// |domains|^arity classes — a launch per class would be launch-bound), tiles in class order.
struct TiledRtPlan {
  std::vector<uint8_t> on;            // per class: taken by this kernel
  fg_class_t *dev_classes = nullptr;  // all classes of the problem (indexed by RtTile::cls)
  RtTile *dev_tiles = nullptr;
  int first_tile[FG_MAX_ARITY + 2] = {0};   // tiles of arity A: [first_tile[A], first_tile[A + 1])
  bool any = false;
};

inline void tiled_rt_free(TiledRtPlan &plan) {
  if (plan.dev_classes) cudaFree(plan.dev_classes);
  if (plan.dev_tiles) cudaFree(plan.dev_tiles);
  plan.dev_classes = nullptr;
  plan.dev_tiles = nullptr;
}

// `skip[i]` != 0: class i already has a compile-time kernel
inline int tiled_rt_plan(const fg_maxsum_desc_t &d, const std::vector<fg_class_t> &classes, const std::vector<uint8_t> &skip,
                         bool enabled, TiledRtPlan &plan) {
  plan.on.assign(classes.size(), 0);
  plan.any = false;
  if (!enabled) return FG_OK;
  const size_t elem = d.precision == FG_F64 ? 8 : 4;
  std::vector<RtTile> tiles;
  for (int a = 1; a <= FG_MAX_ARITY; ++a) {
    plan.first_tile[a] = (int)tiles.size();
    for (size_t i = 0; i < classes.size(); ++i) {
      const fg_class_t &c = classes[i];
      if (c.arity != a || c.n_factors == 0 || (c.flags & FG_CLASS_GHOST) || skip[i]) continue;
      const TiledRtCfg cfg = tiled_rt_cfg(c, elem);
      if (cfg.nf_tile == 0) continue;
      plan.on[i] = 1;
      for (int f0 = 0; f0 < c.n_factors; f0 += cfg.nf_tile)
        tiles.push_back(RtTile{(int32_t)i, f0, std::min(cfg.nf_tile, c.n_factors - f0), cfg.sp});
    }
  }
  plan.first_tile[FG_MAX_ARITY + 1] = (int)tiles.size();
  plan.first_tile[0] = 0;
  if (tiles.empty()) return FG_OK;
  plan.any = true;
  if (cudaMalloc(&plan.dev_classes, classes.size() * sizeof(fg_class_t)) != cudaSuccess) return FG_ERR_CUDA;
  if (cudaMalloc(&plan.dev_tiles, tiles.size() * sizeof(RtTile)) != cudaSuccess) return FG_ERR_CUDA;
  if (cudaMemcpy(plan.dev_classes, classes.data(), classes.size() * sizeof(fg_class_t), cudaMemcpyHostToDevice) != cudaSuccess) return FG_ERR_CUDA;
  if (cudaMemcpy(plan.dev_tiles, tiles.data(), tiles.size() * sizeof(RtTile), cudaMemcpyHostToDevice) != cudaSuccess) return FG_ERR_CUDA;
  return FG_OK;
}

template <typename T, int A>
__global__ void __launch_bounds__(FG_TILED_RT_THREADS)
k_f2v_tiled_rt(const fg_class_t *__restrict__ classes, const RtTile *__restrict__ tiles, const T *__restrict__ tables,
               const T *__restrict__ q_cur, const T *__restrict__ r_cur, T *__restrict__ r_next,
               const int64_t *__restrict__ edge_qoff, uint8_t *__restrict__ r_cnt, uint8_t *__restrict__ r_sent, MaxSumParams p) {
  extern __shared__ __align__(16) unsigned char fg_tiled_rt_smem[];
  const RtTile tile = tiles[blockIdx.x];
  const fg_class_t &c = classes[tile.cls];
  const int nf_tile = tile.nf, sp = tile.sp;
  const int S = (int)c.table_size, RT = c.row_total;
  T *tab = reinterpret_cast<T *>(fg_tiled_rt_smem);   // nf_tile x sp
  T *qs = tab + (size_t)nf_tile * sp;                 // nf_tile x RT   incoming v->f rows, scope order
  T *cand = qs + (size_t)nf_tile * RT;                // nf_tile x RT   new f->v rows
  T *prev = cand + (size_t)nf_tile * RT;              // nf_tile x RT   previous f->v rows
  const int f0 = tile.f0, nf = tile.nf;
  const int first_edge = c.first_edge;
  const int64_t table_base = c.table_base, msg_base = c.msg_base;
  const int tid = threadIdx.x, nt = blockDim.x;
  int dom[A], roff[A], stride[A];
  {
    int s = 1;
#pragma unroll
    for (int i = A - 1; i >= 0; --i) { dom[i] = c.dom[i]; roff[i] = c.row_off[i]; stride[i] = s; s *= c.dom[i]; }
  }
  // --- stage: tables and previous rows are contiguous runs; q rows are gathered per edge
  {
    const T *src = tables + table_base + (int64_t)f0 * S;
    const int n = nf * S;
    for (int i = tid; i < n; i += nt) {
      const int f = i / S, k = i - f * S;
      tab[f * sp + k] = src[i];
    }
    const T *rp = r_cur + msg_base + (int64_t)f0 * RT;
    const int64_t *qo = edge_qoff + first_edge + (int64_t)f0 * A;
    const int m = nf * RT;
    for (int i = tid; i < m; i += nt) {
      prev[i] = rp[i];
      const int f = i / RT, k = i - f * RT;
      int j = 0, ro = 0;
#pragma unroll
      for (int t = 1; t < A; ++t)
        if (k >= roff[t]) { j = t; ro = roff[t]; }
      int dj = dom[0];
#pragma unroll
      for (int t = 1; t < A; ++t)
        if (j == t) dj = dom[t];
      qs[i] = (k - ro) < dj ? q_cur[qo[f * A + j] + (k - ro)] : (T)0;   // rows may be padded (row_off is aligned)
    }
  }
  __syncthreads();
  // --- one item per output value: optimum over the slice x_j = xv of the table
  {
    const bool mx = p.mode_max != 0;
    const int m = nf * RT;
    for (int it = tid; it < m; it += nt) {
      const int f = it / RT, k = it - f * RT;
      int j = 0, ro = 0;
#pragma unroll
      for (int t = 1; t < A; ++t)
        if (k >= roff[t]) { j = t; ro = roff[t]; }
      const int xv = k - ro;
      int dj = dom[0];
#pragma unroll
      for (int t = 1; t < A; ++t)
        if (j == t) dj = dom[t];
      if (xv >= dj) { cand[it] = prev[it]; continue; }   // padding between rows: carried over
      const T *tf = tab + f * sp;
      const T *qf = qs + f * RT;
      int x[A];
#pragma unroll
      for (int i = 0; i < A; ++i) x[i] = 0;
      int idx = 0;
#pragma unroll
      for (int i = 0; i < A; ++i)
        if (i == j) idx = xv * stride[i];
      T opt = mx ? -Inf<T>::pos() : Inf<T>::pos();
      for (;;) {
        T sum = (T)0;
#pragma unroll
        for (int i = 0; i < A; ++i)
          if (i != j) sum += qf[roff[i] + x[i]];
        const T cur = tf[idx] + sum;
        opt_update(opt, cur, mx);
        bool more = false;
#pragma unroll
        for (int i = A - 1; i >= 0; --i) {
          if (more || i == j) continue;
          ++x[i];
          idx += stride[i];
          if (x[i] < dom[i]) { more = true; continue; }
          idx -= dom[i] * stride[i];
          x[i] = 0;
        }
        if (!more) break;
      }
      cand[it] = opt;
    }
  }
  __syncthreads();
  // --- per edge: damping, approx_match, send gate (maxsum.py:339-379, 679-710); results replace `cand`
  {
    const T lam = (T)p.damping, oml = (T)p.one_minus_damping, stab = (T)p.stability;
    const int ne = nf * A;
    for (int le = tid; le < ne; le += nt) {
      const int f = le / A, j = le - f * A;
      const int e = first_edge + (f0 + f) * A + j;
      int d = dom[0], ro = roff[0];
#pragma unroll
      for (int t = 1; t < A; ++t)
        if (j == t) { d = dom[t]; ro = roff[t]; }
      T *cr = cand + f * RT + ro;
      const T *pr = prev + f * RT + ro;
      uint8_t cnt = r_cnt[e];
      const bool has_prev = cnt & 1;
      const bool damp = p.damp_factors && has_prev;
      bool match = has_prev;
      for (int xv = 0; xv < d; ++xv) {
        T v = cr[xv];
        const T pv = pr[xv];
        if (damp) v = lam * pv + oml * v;
        if (has_prev && !approx_match1<T>(v, pv, stab)) match = false;
        cr[xv] = v;
      }
      const bool sent = gate_decide(match, cnt);
      if (!sent)
        for (int xv = 0; xv < d; ++xv) cr[xv] = pr[xv];
      r_cnt[e] = cnt;
      if (r_sent) r_sent[e] = sent ? 1 : 0;
    }
  }
  __syncthreads();
  {
    T *dst = r_next + msg_base + (int64_t)f0 * RT;
    const int m = nf * RT;
    for (int i = tid; i < m; i += nt) dst[i] = cand[i];
  }
}

// every class of the plan, one launch per arity present
template <typename T, int A>
inline int launch_f2v_tiled_rt(const TiledRtPlan &plan, const fg_maxsum_desc_t &d, const T *q_cur, const T *r_cur, T *r_next,
                               const MaxSumParams &p, cudaStream_t st) {
  const int t0 = plan.first_tile[A], t1 = plan.first_tile[A + 1];
  if (t1 <= t0) return 0;
  static bool attr_done = false;   // per instantiation
  if (!attr_done) {
    cudaFuncSetAttribute(k_f2v_tiled_rt<T, A>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FG_TILED_RT_SMEM);
    attr_done = true;
  }
  k_f2v_tiled_rt<T, A><<<t1 - t0, FG_TILED_RT_THREADS, FG_TILED_RT_SMEM, st>>>(plan.dev_classes, plan.dev_tiles + t0, (const T *)d.dev_tables,
                                                                            q_cur, r_cur, r_next, d.dev_edge_qoff, d.dev_r_cnt,
                                                                            d.dev_r_sent, p);
  return 1;
}

template <typename T>
inline int dispatch_f2v_tiled_rt(const TiledRtPlan &plan, const fg_maxsum_desc_t &d, const T *q_cur, const T *r_cur, T *r_next,
                                 const MaxSumParams &p, cudaStream_t st) {
  int n = 0;
  if (!plan.any) return 0;
#define X(a) n += launch_f2v_tiled_rt<T, a>(plan, d, q_cur, r_cur, r_next, p, st);
  X(1) X(2) X(3) X(4) X(5) X(6) X(7) X(8)
#undef X
  return n;
}
