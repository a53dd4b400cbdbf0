// Warp-autonomous MaxSum kernels (round 2): every warp runs its OWN software pipeline over its own
// tiles — its own shared-memory stages, its own mbarriers and cp.async groups — so there is no
// block-wide barrier anywhere in the steady state (round 1's k_v2f_pipe spent 45 % of its issue
// slots parked at __syncthreads, profiles/r01_final_ncu_v2f_pipe_c2.txt).
//
//  k_v2f_warp<T,D>   variable -> factor (+ select_value).  A tile = nv variables of one (D, K) class
//      = nv*K <= 32 consecutive slots, ONE SLOT PER LANE.  HBM side: the previous q rows and the unary
//      rows of a tile are contiguous (1-D bulk async copies, TMA, per-warp mbarrier), the K r rows of a
//      variable are gathered through slot_roff with cp.async — lanes mapped (row, 8/16-byte piece) so
//      one LDGSTS instruction touches ~7 rows instead of 32 — and the produced q rows leave with one
//      bulk store per tile (written in place over the staged previous rows).  While tile k is computed
//      the loads of tile k+1 are in flight.
//      Arithmetic per lane (slot f of variable i): costs_for_factor (maxsum.py:623-676) with the
//      reference's operand order — value-major, then factors in `links` order, the own factor skipped —
//      fully unrolled for K = 1..8; the lane of the LAST slot also holds ur + r_0 + ... + r_{K-2}, so
//      select_value (maxsum.py:584-620) costs it one more add per value.
//
// Results are bit-identical to the generic kernels, the round-1 pipelined kernels and the CPU oracle
// of the same precision.
#pragma once
#include "maxsum_fast.cuh"

#define FG_WARP_MAX_ENTRIES 24

struct WTileEntry {
  fg_varclass_t vc;
  int32_t tile_begin;  // first tile of this class in the launch
  int32_t nv_tile;     // variables per tile (nv_tile * degree <= 32)
};
struct WTileTable {
  int32_t n;
  int32_t total_tiles;
  WTileEntry e[FG_WARP_MAX_ENTRIES];
};

struct WTile {
  int K, nv, nslots, slot0, var0, valid;
  uint32_t qoff, uoff;  // element offsets (the fast plans require 32-bit message offsets)
};

// tile t of the launch; `ci` is a cursor that only moves forward (a warp visits its tiles in order)
__device__ __forceinline__ WTile wtile_at(const WTileTable &tab, int t, int D, int &ci) {
  WTile o;
  o.valid = t < tab.total_tiles;
  if (!o.valid) { o.K = 1; o.nv = o.nslots = o.slot0 = o.var0 = 0; o.qoff = o.uoff = 0; return o; }
#pragma unroll 1
  while (ci + 1 < tab.n && t >= tab.e[ci + 1].tile_begin) ++ci;
  const WTileEntry &en = tab.e[ci];
  o.K = en.vc.degree;
  const int v0 = (t - en.tile_begin) * en.nv_tile;
  o.nv = min(en.nv_tile, en.vc.n_vars - v0);
  o.nslots = o.nv * o.K;
  o.slot0 = en.vc.first_slot + v0 * o.K;
  o.var0 = en.vc.first_var + v0;
  o.qoff = (uint32_t)(en.vc.q_base + (int64_t)v0 * o.K * D);
  o.uoff = (uint32_t)(en.vc.unary_base + (int64_t)v0 * D);
  return o;
}

// One lane = one slot f of a variable whose K gathered r rows start at `col` (row g at col + g*D) and
// whose own costs are `ur`.  cand <- un-normalised message (own factor skipped), returns sum_cost;
// best / best_c <- select_value over ur + sum of ALL K rows, meaningful on the lane with f == K-1 only
// (its message chain is exactly the first K-1 terms of that sum).
// K > 0: compile-time degree; K == 0: run-time degree k_rt.
template <typename T, int D, int K>
__device__ __forceinline__ T v2f_lane_msg(const T *__restrict__ col, const T *__restrict__ ur, int f, int k_rt, bool mx,
                                          T (&cand)[D], int &best, T &best_c) {
  constexpr int VR = V2FCfg<T, D>::VR;
  T sum_cost = (T)0;
  best = 0;
  best_c = (T)0;
  if constexpr (K > 0) {
#pragma unroll
    for (int x0 = 0; x0 < D; x0 += VR) {
      T u[VR], c[K][VR];
      ld_row<T, VR, VR>(ur + x0, u);
#pragma unroll
      for (int g = 0; g < K; ++g) ld_row<T, VR, VR>(col + g * D + x0, c[g]);
#pragma unroll
      for (int xx = 0; xx < VR; ++xx) {
        T m = u[xx];
#pragma unroll
        for (int g = 0; g < K; ++g) {
          if (g != f) {
            sum_cost += c[g][xx];
            m += c[g][xx];
          }
        }
        cand[x0 + xx] = m;
        const T tot = m + c[K - 1][xx];  // lane f == K-1: ((ur + r_0) + ...) + r_{K-1}
        const int x = x0 + xx;
        if (x == 0 || (mx ? (tot > best_c) : (tot < best_c))) { best = x; best_c = tot; }
      }
    }
  } else {  // run-time degree (9..16): values unrolled, factors in a loop
#pragma unroll
    for (int x = 0; x < D; ++x) {
      T m = ur[x];
      T last = (T)0;
#pragma unroll 1
      for (int g = 0; g < k_rt; ++g) {
        const T cst = col[g * D + x];
        last = cst;
        if (g != f) {
          sum_cost += cst;
          m += cst;
        }
      }
      cand[x] = m;
      const T tot = m + last;  // lane f == K-1 skipped exactly the last row
      if (x == 0 || (mx ? (tot > best_c) : (tot < best_c))) { best = x; best_c = tot; }
    }
  }
  return sum_cost;
}

#define FG_WARP_K_SWITCH(K_, CALL)                 \
  switch (K_) {                                    \
    case 1: { constexpr int KK = 1; CALL; } break; \
    case 2: { constexpr int KK = 2; CALL; } break; \
    case 3: { constexpr int KK = 3; CALL; } break; \
    case 4: { constexpr int KK = 4; CALL; } break; \
    case 5: { constexpr int KK = 5; CALL; } break; \
    case 6: { constexpr int KK = 6; CALL; } break; \
    case 7: { constexpr int KK = 7; CALL; } break; \
    case 8: { constexpr int KK = 8; CALL; } break; \
  }

#define FG_V2FW_WARPS 4   // warps per CTA (each warp is independent; the CTA is only a container)
#ifndef FG_V2FW_MINB
#define FG_V2FW_MINB 6   // register cap 65536 / (6 * 128) = 85: the factor side runs concurrently on the same SMs
#endif

template <typename T, int D, int NS_>
struct V2FWarpCfg {
  static constexpr int VR = V2FCfg<T, D>::VR;
  static constexpr int VR_BYTES = V2FCfg<T, D>::VR_BYTES;
  static constexpr int PIECES = D / VR;                     // async-copy pieces per row
  static constexpr int STAGE = 3 * 32 * D;                  // rrow | qio | un   (elements)
  static constexpr int NS = NS_;
  static constexpr size_t WARP_SMEM = (size_t)NS * STAGE * sizeof(T);
  static constexpr size_t SMEM = FG_V2FW_WARPS * WARP_SMEM + FG_V2FW_WARPS * NS * sizeof(uint64_t) + 16;
  static constexpr bool OK = SMEM <= FG_SMEM_LIMIT;
};

template <typename T, int D, int NS_, typename OffT>
__global__ void __launch_bounds__(FG_V2FW_WARPS * 32, FG_V2FW_MINB)
k_v2f_warp(const WTileTable tab, const OffT *__restrict__ slot_roff, const T *__restrict__ unary,
           const T *__restrict__ r_cur, const T *__restrict__ q_cur, T *__restrict__ q_next,
           uint8_t *__restrict__ q_cnt, uint8_t *__restrict__ q_sent, int32_t *__restrict__ value,
           T *__restrict__ value_cost, MaxSumParams p) {
  using C = V2FWarpCfg<T, D, NS_>;
  constexpr int VR = C::VR, PIECES = C::PIECES, NS = C::NS;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  T *stage0 = reinterpret_cast<T *>(smem_raw + (size_t)wib * C::WARP_SMEM);
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem_raw + (size_t)FG_V2FW_WARPS * C::WARP_SMEM) + NS * wib;
  if (lane == 0) {
#pragma unroll
    for (int s = 0; s < NS; ++s) mbar_init(&bars[s], 1);
    fence_mbar_init();
  }
  __syncwarp();

  const int gw = (int)blockIdx.x * FG_V2FW_WARPS + wib;       // global warp id
  const int nw = (int)gridDim.x * FG_V2FW_WARPS;
  // one forward-only cursor per role (compute, issue, gather-index loads, counter loads): each role
  // visits this warp's tiles in ascending order
  int cur_c = 0, cur_i = 0, cur_l = 0, cur_n = 0;
  auto tile_at = [&](int k, int &cursor) { return wtile_at(tab, gw + k * nw, D, cursor); };

  auto load_roff = [&](const WTile &t) -> OffT {
    return (t.valid && lane < t.nslots) ? slot_roff[t.slot0 + lane] : (OffT)0;
  };
  auto load_cnt = [&](const WTile &t) -> uint8_t {
    return (t.valid && lane < t.nslots) ? q_cnt[t.slot0 + lane] : (uint8_t)0;
  };
  // start every load of tile t into stage s (all lanes)
  auto issue = [&](int s, const WTile &t, OffT roff) {
    if (t.valid) {
      T *rrow = stage0 + s * C::STAGE;
      T *qio = rrow + 32 * D;
      T *un = qio + 32 * D;
      const uint32_t qbytes = (uint32_t)(t.nslots * D) * (uint32_t)sizeof(T);
      const uint32_t ubytes = (uint32_t)(t.nv * D) * (uint32_t)sizeof(T);
      const bool tma_q = (qbytes % 16 == 0) && ((((int64_t)t.qoff * (int64_t)sizeof(T)) & 15) == 0);
      const bool tma_u = (ubytes % 16 == 0) && ((((int64_t)t.uoff * (int64_t)sizeof(T)) & 15) == 0);
      if (lane == 0) {
        mbar_expect_tx(&bars[s], (tma_q ? qbytes : 0u) + (tma_u ? ubytes : 0u));
        if (tma_q) tma_load_1d(qio, q_cur + t.qoff, qbytes, &bars[s]);
        if (tma_u) tma_load_1d(un, unary + t.uoff, ubytes, &bars[s]);
      }
      if (!tma_q)
        for (int i = lane; i < t.nslots * D; i += 32) cp_async_b<(int)sizeof(T)>(qio + i, q_cur + t.qoff + i);
      if (!tma_u)
        for (int i = lane; i < t.nv * D; i += 32) cp_async_b<(int)sizeof(T)>(un + i, unary + t.uoff + i);
      // r rows: lane <-> (row, piece), consecutive lanes on consecutive pieces of one row
#pragma unroll
      for (int it = 0; it < PIECES; ++it) {
        const int pc = it * 32 + lane;
        const int row = pc / PIECES, piece = pc - row * PIECES;
        const OffT ro = __shfl_sync(0xffffffffu, roff, row & 31);
        if (row < t.nslots)
          cp_async_b<C::VR_BYTES>(rrow + row * D + piece * VR, r_cur + (int64_t)ro + piece * VR);
      }
    }
    cp_async_commit();  // one group per tile, even when empty: uniform accounting
  };

  const bool mx = p.mode_max != 0;
  const T lam = (T)p.damping, oml = (T)p.one_minus_damping, stab = (T)p.stability;

  // prologue: tiles 0 .. NS-2 in flight, gather indices of tile NS-1 and counters of tile 0 on their way
#pragma unroll 1
  for (int k = 0; k < NS - 1; ++k) {
    const WTile t = tile_at(k, cur_i);
    issue(k, t, load_roff(t));
    cur_l = cur_i;
  }
  OffT roff_n = load_roff(tile_at(NS - 1, cur_l));
  uint8_t cnt = load_cnt(tile_at(0, cur_n));

#pragma unroll 1
  for (int k = 0;; ++k) {
    const WTile t = tile_at(k, cur_c);
    if (!t.valid) break;
    const int s = k % NS;
    // tile k+NS-1 goes into the stage tile k-1 held: its bulk store must have finished READING it
    if (lane == 0) tma_store_wait_read();
    __syncwarp();
    issue((k + NS - 1) % NS, tile_at(k + NS - 1, cur_i), roff_n);
    const OffT roff_n2 = load_roff(tile_at(k + NS, cur_l));
    const uint8_t cnt_n = load_cnt(tile_at(k + 1, cur_n));
    cp_async_wait_group<NS - 1>();                  // my gathers of tile k have landed ...
    mbar_wait(&bars[s], (uint32_t)((k / NS) & 1));  // ... and so have its bulk copies
    __syncwarp();                                   // ... and every other lane's gathers

    const int K = t.K;
    T *rrow = stage0 + s * C::STAGE;
    T *qio = rrow + 32 * D;
    const T *un = qio + 32 * D;
    const bool act = lane < t.nslots;
    if (act) {
      const int i = lane / K, f = lane - i * K;
      const T *col = rrow + i * K * D;
      const T *ur = un + i * D;
      T cand[D], prev[D];
      int best = 0;
      T best_c = (T)0, sum_cost = (T)0;
      if (K <= 8) {
        FG_WARP_K_SWITCH(K, (sum_cost = v2f_lane_msg<T, D, KK>(col, ur, f, K, mx, cand, best, best_c)))
      } else {
        sum_cost = v2f_lane_msg<T, D, 0>(col, ur, f, K, mx, cand, best, best_c);
      }
      const T avg = sum_cost / (T)D;
#pragma unroll
      for (int x = 0; x < D; ++x) cand[x] = cand[x] - avg;
      ld_row<T, D, VR>(qio + lane * D, prev);
      uint8_t c8 = cnt;
      const bool sent = damp_gate_row<T, D>(cand, prev, c8, p.damp_vars != 0, lam, oml, stab);
      st_row<T, D, VR>(qio + lane * D, cand);
      q_cnt[t.slot0 + lane] = c8;
      if (q_sent) q_sent[t.slot0 + lane] = sent ? 1 : 0;
      if (f == K - 1) {
        value[t.var0 + i] = best;
        value_cost[t.var0 + i] = best_c;
      }
    }
    const uint32_t obytes = (uint32_t)(t.nslots * D) * (uint32_t)sizeof(T);
    if ((obytes % 16 == 0) && ((((int64_t)t.qoff * (int64_t)sizeof(T)) & 15) == 0)) {
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        tma_store_1d(q_next + t.qoff, qio, obytes);
        tma_store_commit();
      }
    } else {
      __syncwarp();
      for (int i = lane; i < t.nslots * D; i += 32) q_next[t.qoff + i] = qio[i];
      __syncwarp();
    }
    roff_n = roff_n2;
    cnt = cnt_n;
  }
  cp_async_wait_all();
  if (lane == 0) tma_store_wait_read();
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
// Regular variable classes (degree 1..16, one domain size D, not ghosts) -> tile tables of at most
// FG_WARP_MAX_ENTRIES classes.  nv_tile: as many variables as fit 32 lanes, rounded down to the number
// of rows that keeps every tile's q / unary offsets 16-byte aligned (bulk copies) when possible.
inline void v2fw_build_tables(const std::vector<fg_varclass_t> &vcs, int D, size_t elem, std::vector<WTileTable> &out) {
  WTileTable cur;
  cur.n = 0;
  cur.total_tiles = 0;
  auto flush = [&]() {
    if (cur.n) out.push_back(cur);
    cur.n = 0;
    cur.total_tiles = 0;
  };
  std::vector<fg_varclass_t> order(vcs);
  std::stable_sort(order.begin(), order.end(),
                   [](const fg_varclass_t &a, const fg_varclass_t &b) { return a.degree > b.degree; });
  const int unit = 16 / fg_gcd(16, D * (int)elem);  // rows per 16-byte multiple
  for (const fg_varclass_t &vc : order) {
    if (vc.dom != D || vc.degree < 1 || vc.degree > 32 || vc.n_vars == 0) continue;
    int nv = 32 / vc.degree;
    if (nv >= unit) nv = nv / unit * unit;
    WTileEntry e;
    e.vc = vc;
    e.tile_begin = cur.total_tiles;
    e.nv_tile = nv;
    cur.e[cur.n++] = e;
    cur.total_tiles += (vc.n_vars + nv - 1) / nv;
    if (cur.n == FG_WARP_MAX_ENTRIES) flush();
  }
  flush();
}

template <typename T, int D, int NS_>
inline bool launch_v2f_warp_ns(const WTileTable &tab, const fg_maxsum_desc_t &d, const T *r_cur, const T *q_cur, T *q_next,
                               const MaxSumParams &p, cudaStream_t st) {
  using C = V2FWarpCfg<T, D, NS_>;
  if constexpr (!C::OK) return false;
  auto kern = k_v2f_warp<T, D, NS_, uint32_t>;
  static int per_sm = 0, n_sm = 0;  // one per instantiation
  if (!per_sm) {
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM);
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, FG_V2FW_WARPS * 32, C::SMEM);
    if (per_sm < 1) per_sm = 1;
    const int cap = fg_env_int("PYDCOP_B200_V2FW_CPS", 2);  // leaves room for the factor side (runs concurrently)
    if (per_sm > cap) per_sm = cap;
  }
  const int need = (tab.total_tiles + FG_V2FW_WARPS - 1) / FG_V2FW_WARPS;
  const unsigned blocks = (unsigned)std::min(need, n_sm * per_sm);
  kern<<<blocks, FG_V2FW_WARPS * 32, C::SMEM, st>>>(tab, d.dev_slot_roff32, (const T *)d.dev_unary, r_cur, q_cur, q_next,
                                                     d.dev_q_cnt, d.dev_q_sent, d.dev_value, (T *)d.dev_value_cost, p);
  return true;
}

// pipeline depth: PYDCOP_B200_V2FW_NS = 2 | 3 | 4 stages per warp (default 3)
template <typename T, int D>
inline void launch_v2f_warp(const WTileTable &tab, const fg_maxsum_desc_t &d, const T *r_cur, const T *q_cur, T *q_next,
                            const MaxSumParams &p, cudaStream_t st) {
  static const int ns = fg_env_int("PYDCOP_B200_V2FW_NS", 3);
  if (ns >= 4 && launch_v2f_warp_ns<T, D, 4>(tab, d, r_cur, q_cur, q_next, p, st)) return;
  if (ns >= 3 && launch_v2f_warp_ns<T, D, 3>(tab, d, r_cur, q_cur, q_next, p, st)) return;
  launch_v2f_warp_ns<T, D, 2>(tab, d, r_cur, q_cur, q_next, p, st);
}

template <typename T>
inline bool dispatch_v2f_warp(int D, const WTileTable &tab, const fg_maxsum_desc_t &d, const T *r_cur, const T *q_cur,
                              T *q_next, const MaxSumParams &p, cudaStream_t st) {
  switch (D) {
#define X(n) case n: launch_v2f_warp<T, n>(tab, d, r_cur, q_cur, q_next, p, st); return true;
    FG_FAST_DOMS(X)
#undef X
  }
  return false;
}

struct MaxSumWarpPlan {
  std::vector<uint8_t> f2v;          // per factor class: warp kernel available (PYDCOP_B200_F2V != pipe)
  bool v2f_on = false;               // PYDCOP_B200_V2F != pipe
  std::vector<WTileTable> v2f;       // launches over the regular variable classes
  std::vector<int> v2f_dom;
};

// same class selection as maxsum_fast_plan (every regular class of a compiled domain size)
inline void maxsum_warp_plan(const fg_maxsum_desc_t &d, const std::vector<fg_varclass_t> &vcs, const MaxSumFastPlan &fast,
                             MaxSumWarpPlan &plan) {
  plan.v2f.clear();
  plan.v2f_dom.clear();
  const char *e = getenv("PYDCOP_B200_V2F");
  plan.v2f_on = fast.off32 && !fg_fast_disabled() && !(e && e[0] == 'p');
  if (!plan.v2f_on) return;
  const size_t elem = d.precision == FG_F64 ? 8 : 4;
  std::vector<uint8_t> taken(vcs.size(), 0);
  for (size_t i = 0; i < vcs.size(); ++i) {
    if (taken[i] || (vcs[i].flags & FG_CLASS_GHOST) || vcs[i].degree < 1 || !fg_fast_dom(vcs[i].dom)) continue;
    const int D = vcs[i].dom;
    std::vector<fg_varclass_t> same;
    for (size_t j = i; j < vcs.size(); ++j)
      if (!taken[j] && !(vcs[j].flags & FG_CLASS_GHOST) && vcs[j].dom == D && vcs[j].degree >= 1) {
        same.push_back(vcs[j]);
        taken[j] = 1;
      }
    std::vector<WTileTable> ts;
    v2fw_build_tables(same, D, elem, ts);
    for (auto &t : ts) { plan.v2f.push_back(t); plan.v2f_dom.push_back(D); }
  }
}

// ------------------------------------------------------------------------------------------------
//  k_f2v_warp<T,D>   factor -> variable, binary factors over one (even) domain size D.
//      A tile = 16 factors = 32 directed edges, TWO LANES PER FACTOR: lane (f, h) owns the table rows
//      x0 in [h*D/2, (h+1)*D/2) and walks them ONCE, producing from the same shared-memory read
//        - its D/2 values of the marginal towards position 0 (row optimum of T[x0][.] + q1), and
//        - partial optima over its rows of the marginal towards position 1 (T[.][x1] + q0[x0]),
//      which the two lanes of a factor complete with D/2 shuffles (the optimum of a set of floats is
//      exact, so the split changes no bit).  Round 1's kernel read every table twice (once per directed
//      edge) and spent 31 % of its shared-memory wavefronts on bank conflicts; here the lanes of a
//      half-warp read 16 distinct 8-byte bank pairs (factor stride D*D words, half stride D*D/2).
//      HBM side per warp and stage: tables and previous r rows of the tile are contiguous (1-D bulk
//      async copies on a per-warp mbarrier), the 32 q rows are gathered through edge_qoff with
//      cp.async (lanes <-> (row, piece)), the produced r rows are written in place over the staged
//      previous rows and leave with one bulk store.  NS stages: tile k computes while k+1 .. k+NS-1 load.
// ------------------------------------------------------------------------------------------------
template <typename T, int D, int NS_>
struct F2VWarpCfg {
  static constexpr int S = D * D, HD = D / 2, NF = 16;
  static constexpr int VR = V2FCfg<T, D>::VR;          // elements per row vector (row = D elements)
  static constexpr int VR_BYTES = V2FCfg<T, D>::VR_BYTES;
  static constexpr int PIECES = D / VR;
  static constexpr int VH_BYTES = fg_gcd(16, HD * (int)sizeof(T));   // half rows
  static constexpr int VH = VH_BYTES / (int)sizeof(T);
  static constexpr int STAGE = NF * S + 2 * NF * 2 * D;              // tab | rt (in/out) | qt   (elements)
  static constexpr int NS = NS_;
  static constexpr size_t WARP_SMEM = (size_t)NS * STAGE * sizeof(T);
  static constexpr size_t smem_for(int w) { return (size_t)w * WARP_SMEM + (size_t)w * NS * sizeof(uint64_t) + 16; }
  // warps per CTA: the CTA is only a container of independent warps
  static constexpr int WARPS = smem_for(2) <= FG_SMEM_LIMIT ? 2 : (smem_for(1) <= FG_SMEM_LIMIT ? 1 : 0);
  static constexpr size_t SMEM = smem_for(WARPS > 0 ? WARPS : 1);
  static constexpr bool OK = (D % 2 == 0) && D >= 4 && WARPS > 0;
};

template <typename T, int D, int NS_, typename OffT>
__global__ void __launch_bounds__(F2VWarpCfg<T, D, NS_>::WARPS > 0 ? F2VWarpCfg<T, D, NS_>::WARPS * 32 : 32)
k_f2v_warp(const fg_class_t c, const T *__restrict__ tables, const T *__restrict__ q_cur,
           const T *__restrict__ r_cur, T *__restrict__ r_next, const OffT *__restrict__ edge_qoff,
           uint8_t *__restrict__ r_cnt, uint8_t *__restrict__ r_sent, MaxSumParams p) {
  using C = F2VWarpCfg<T, D, NS_>;
  constexpr int S = C::S, HD = C::HD, NF = C::NF, NS = C::NS, VR = C::VR, PIECES = C::PIECES, R = 2 * D;
  constexpr int WARPS = C::WARPS > 0 ? C::WARPS : 1;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  T *stage0 = reinterpret_cast<T *>(smem_raw + (size_t)wib * C::WARP_SMEM);
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem_raw + (size_t)WARPS * C::WARP_SMEM) + NS * wib;
  if (lane == 0) {
#pragma unroll
    for (int s = 0; s < NS; ++s) mbar_init(&bars[s], 1);
    fence_mbar_init();
  }
  __syncwarp();

  const int gw = (int)blockIdx.x * WARPS + wib;
  const int nw = (int)gridDim.x * WARPS;
  const int n_tiles = (c.n_factors + NF - 1) / NF;
  const int n_my = gw < n_tiles ? (n_tiles - 1 - gw) / nw + 1 : 0;

  // lane = local edge le = 2 f + j of my k-th tile: gather offset / gate counter
  auto load_qo = [&](int k) -> OffT {
    if (k < n_my) {
      const int f0 = (gw + k * nw) * NF;
      if (lane < 2 * min(NF, c.n_factors - f0)) return edge_qoff[c.first_edge + f0 * 2 + lane];
    }
    return (OffT)0;
  };
  auto load_cnt = [&](int k) -> uint8_t {
    if (k < n_my) {
      const int f0 = (gw + k * nw) * NF;
      if (lane < 2 * min(NF, c.n_factors - f0)) return r_cnt[c.first_edge + f0 * 2 + lane];
    }
    return (uint8_t)0;
  };
  auto issue = [&](int k, OffT qo) {
    if (k < n_my) {
      const int s = k % NS;
      const int f0 = (gw + k * nw) * NF;
      const int nf = min(NF, c.n_factors - f0);
      T *tab = stage0 + s * C::STAGE;
      T *rt = tab + NF * S;
      T *qt = rt + NF * R;
      const T *gtab = tables + c.table_base + (int64_t)f0 * S;
      const T *grow = r_cur + c.msg_base + (int64_t)f0 * R;
      const uint32_t tb = (uint32_t)(nf * S) * (uint32_t)sizeof(T), rb = (uint32_t)(nf * R) * (uint32_t)sizeof(T);
      const bool tma_t = (tb % 16 == 0), tma_r = (rb % 16 == 0);   // bases are 128-byte aligned, tiles 16 factors
      if (lane == 0) {
        mbar_expect_tx(&bars[s], (tma_t ? tb : 0u) + (tma_r ? rb : 0u));
        if (tma_t) tma_load_1d(tab, gtab, tb, &bars[s]);
        if (tma_r) tma_load_1d(rt, grow, rb, &bars[s]);
      }
      if (!tma_t) for (int i = lane; i < nf * S; i += 32) cp_async_b<(int)sizeof(T)>(tab + i, gtab + i);
      if (!tma_r) for (int i = lane; i < nf * R; i += 32) cp_async_b<(int)sizeof(T)>(rt + i, grow + i);
#pragma unroll
      for (int it = 0; it < PIECES; ++it) {
        const int pc = it * 32 + lane;
        const int row = pc / PIECES, piece = pc - row * PIECES;
        const OffT ro = __shfl_sync(0xffffffffu, qo, row & 31);
        if (row < 2 * nf) cp_async_b<C::VR_BYTES>(qt + row * D + piece * VR, q_cur + (int64_t)ro + piece * VR);
      }
    }
    cp_async_commit();
  };

  const bool mx = p.mode_max != 0;
  const T lam = (T)p.damping, oml = (T)p.one_minus_damping, stab = (T)p.stability;
  const T init = mx ? -Inf<T>::pos() : Inf<T>::pos();
  const int fl = lane >> 1, h = lane & 1;

  // prologue: tiles 0 .. NS-2 in flight, gather offsets of tile NS-1 and counters of tile 0 on their way
#pragma unroll 1
  for (int k = 0; k < NS - 1; ++k) issue(k, load_qo(k));
  OffT qo_next = load_qo(NS - 1);
  uint8_t cnt = load_cnt(0);

#pragma unroll 1
  for (int k = 0; k < n_my; ++k) {
    const int s = k % NS;
    // tile k+NS-1 goes into the stage tile k-1 held: its bulk store must have finished reading it
    if (lane == 0) tma_store_wait_read();
    __syncwarp();
    issue(k + NS - 1, qo_next);
    const OffT qo_n2 = load_qo(k + NS);
    const uint8_t cnt_n = load_cnt(k + 1);
    cp_async_wait_group<NS - 1>();                  // my gathers of tile k have landed ...
    mbar_wait(&bars[s], (uint32_t)((k / NS) & 1));  // ... and its bulk copies
    __syncwarp();                                   // ... and every other lane's gathers

    const int f0 = (gw + k * nw) * NF;
    const int nf = min(NF, c.n_factors - f0);
    T *tab = stage0 + s * C::STAGE;
    T *rt = tab + NF * S;
    const T *qt = rt + NF * R;
    const bool act = fl < nf;
    const uint8_t cnt_other = (uint8_t)__shfl_xor_sync(0xffffffffu, (unsigned)cnt, 1);
    T cand0[HD], cand1[HD], prev0[HD], prev1[HD];
    bool m0 = false, m1 = false;
    uint8_t c0 = h == 0 ? cnt : cnt_other, c1 = h == 0 ? cnt_other : cnt;
    {
      const T *tf = tab + fl * S + h * HD * D;   // my HD rows
      T q0[D], q1[D], part[D];
      if (act) {
        ld_row<T, D, VR>(qt + (2 * fl) * D, q0);
        ld_row<T, D, VR>(qt + (2 * fl + 1) * D, q1);
      }
#pragma unroll
      for (int x = 0; x < D; ++x) part[x] = init;
      if (act) {
#pragma unroll
        for (int i = 0; i < HD; ++i) {
          T row[D], a[D];
          ld_row<T, D, VR>(tf + i * D, row);
          const T qx = h == 0 ? q0[i] : q0[HD + i];   // static register indices
#pragma unroll
          for (int x1 = 0; x1 < D; ++x1) {
            a[x1] = row[x1] + q1[x1];
            part[x1] = fg_opt<T>(part[x1], row[x1] + qx, mx);
          }
          cand0[i] = opt_tree<T, D>(a, mx);
        }
      }
      // complete position 1: I keep x1 in my half, the partner lane sends its partial optima for it
#pragma unroll
      for (int i = 0; i < HD; ++i) {
        const T mine = h == 0 ? part[i] : part[HD + i];
        const T give = h == 0 ? part[HD + i] : part[i];
        const T got = __shfl_xor_sync(0xffffffffu, give, 1);
        cand1[i] = fg_opt<T>(mine, got, mx);
      }
      if (act) {
        ld_row<T, HD, C::VH>(rt + fl * R + h * HD, prev0);
        ld_row<T, HD, C::VH>(rt + fl * R + D + h * HD, prev1);
        m0 = damp_match_row<T, HD>(cand0, prev0, (c0 & 1) != 0, p.damp_factors != 0, lam, oml, stab);
        m1 = damp_match_row<T, HD>(cand1, prev1, (c1 & 1) != 0, p.damp_factors != 0, lam, oml, stab);
      }
    }
    const unsigned mm = (m0 ? 1u : 0u) | (m1 ? 2u : 0u);
    const unsigned mo = __shfl_xor_sync(0xffffffffu, mm, 1);
    if (act) {
      const bool s0 = gate_decide((mm & mo & 1u) != 0, c0);
      const bool s1 = gate_decide((mm & mo & 2u) != 0, c1);
      if (!s0) {
#pragma unroll
        for (int i = 0; i < HD; ++i) cand0[i] = prev0[i];
      }
      if (!s1) {
#pragma unroll
        for (int i = 0; i < HD; ++i) cand1[i] = prev1[i];
      }
      st_row<T, HD, C::VH>(rt + fl * R + h * HD, cand0);
      st_row<T, HD, C::VH>(rt + fl * R + D + h * HD, cand1);
      const int e = c.first_edge + f0 * 2 + lane;   // edge (f, j = h)
      r_cnt[e] = h == 0 ? c0 : c1;
      if (r_sent) r_sent[e] = (h == 0 ? s0 : s1) ? 1 : 0;
    }
    const uint32_t ob = (uint32_t)(nf * R) * (uint32_t)sizeof(T);
    T *gout = r_next + c.msg_base + (int64_t)f0 * R;
    if (ob % 16 == 0) {
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        tma_store_1d(gout, rt, ob);
        tma_store_commit();
      }
    } else {
      __syncwarp();
      for (int i = lane; i < nf * R; i += 32) gout[i] = rt[i];
      __syncwarp();
    }
    qo_next = qo_n2;
    cnt = cnt_n;
  }
  cp_async_wait_all();
  if (lane == 0) tma_store_wait_read();
}

template <typename T, int D, int NS_>
inline bool launch_f2v_warp_ns(bool probe, const fg_class_t &c, const fg_maxsum_desc_t &d, const T *q_cur, const T *r_cur,
                               T *r_next, const MaxSumParams &p, cudaStream_t st) {
  using C = F2VWarpCfg<T, D, NS_>;
  if constexpr (!C::OK) {
    return false;
  } else {
    if (probe) return true;
    auto kern = k_f2v_warp<T, D, NS_, uint32_t>;
    static int per_sm = 0, n_sm = 0;  // one per instantiation
    if (!per_sm) {
      cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM);
      int dev = 0;
      cudaGetDevice(&dev);
      cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, C::WARPS * 32, C::SMEM);
      if (per_sm < 1) per_sm = 1;
      const int cap = fg_env_int("PYDCOP_B200_F2VW_CPS", 2);  // CTAs of 2 warps; the variable side shares the SMs
      if (per_sm > cap) per_sm = cap;
    }
    const int n_tiles = (c.n_factors + C::NF - 1) / C::NF;
    const int need = (n_tiles + C::WARPS - 1) / C::WARPS;
    const unsigned blocks = (unsigned)std::min(need, n_sm * per_sm);
    kern<<<blocks, C::WARPS * 32, C::SMEM, st>>>(c, (const T *)d.dev_tables, q_cur, r_cur, r_next, d.dev_edge_qoff32,
                                                d.dev_r_cnt, d.dev_r_sent, p);
    return true;
  }
}

// pipeline depth: PYDCOP_B200_F2VW_NS = 2 | 3 | 4 stages per warp (default 3), shallower when it does not fit
template <typename T, int D>
inline bool launch_f2v_warp(bool probe, const fg_class_t &c, const fg_maxsum_desc_t &d, const T *q_cur, const T *r_cur,
                            T *r_next, const MaxSumParams &p, cudaStream_t st) {
  static const int ns = fg_env_int("PYDCOP_B200_F2VW_NS", 3);
  if (ns >= 4 && launch_f2v_warp_ns<T, D, 4>(probe, c, d, q_cur, r_cur, r_next, p, st)) return true;
  if (ns >= 3 && launch_f2v_warp_ns<T, D, 3>(probe, c, d, q_cur, r_cur, r_next, p, st)) return true;
  return launch_f2v_warp_ns<T, D, 2>(probe, c, d, q_cur, r_cur, r_next, p, st);
}

// binary classes over one even domain size; probe == true: only report whether the kernel exists
template <typename T>
inline bool dispatch_f2v_warp(bool probe, const fg_class_t &c, const fg_maxsum_desc_t &d, const T *q_cur, const T *r_cur,
                              T *r_next, const MaxSumParams &p, cudaStream_t st) {
  if (c.arity != 2 || c.dom[0] != c.dom[1]) return false;
  switch (c.dom[0]) {
#define X(n) case n: return launch_f2v_warp<T, n>(probe, c, d, q_cur, r_cur, r_next, p, st);
    X(4) X(6) X(8) X(10) X(16) X(20)
#undef X
  }
  return false;
}

// factor classes the warp kernel takes over from the round-1 pipelined kernel
inline void maxsum_warp_plan_f2v(const fg_maxsum_desc_t &d, const std::vector<fg_class_t> &classes,
                                 const MaxSumFastPlan &fast, MaxSumWarpPlan &plan) {
  plan.f2v.assign(classes.size(), 0);
  const char *e = getenv("PYDCOP_B200_F2V");
  if (!fast.off32 || fg_fast_disabled() || (e && e[0] == 'p')) return;
  MaxSumParams dummy{};
  for (size_t i = 0; i < classes.size(); ++i) {
    if (classes[i].flags & FG_CLASS_GHOST) continue;
    const bool ok = d.precision == FG_F64
                        ? dispatch_f2v_warp<double>(true, classes[i], d, nullptr, nullptr, nullptr, dummy, nullptr)
                        : dispatch_f2v_warp<float>(true, classes[i], d, nullptr, nullptr, nullptr, dummy, nullptr);
    if (ok) plan.f2v[i] = 1;
  }
}
