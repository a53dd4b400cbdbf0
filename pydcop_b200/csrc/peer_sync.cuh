// Device-side cycle barrier in NVLink peer memory and the peer push with a fused release
// (include/pydcop_b200.h: fg_peer_sync_t, fg_halo_plan_t).
//
// Ordering argument.  Producer rank A, consumer rank B, epoch e:
//   A: every thread of the push kernel stores its piece of a boundary row into B's buffer (plain
//      st.global to an IPC-mapped address -> NVLink), executes fence.acq_rel.sys
//      (__threadfence_system) and arrives on a device counter; the LAST block to arrive fences again
//      and writes e into its slot of B's flag array with st.release.sys.  The counter's atomics order
//      every other block's fenced stores before the last block's release (the classic
//      threadfence-reduction pattern, at system scope).
//   B: one warp spins with ld.acquire.sys on its own flag array until every peer's slot holds >= e.
//      Kernels launched behind the wait kernel on the same stream then read the rows.
// Flags only grow, so a late observer can never see an older epoch after a newer one.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <type_traits>

#include "../../include/pydcop_b200.h"

struct PeerSlots {
  int32_t n;
  int32_t rank[FG_MAX_PEERS];
  uint64_t *slot[FG_MAX_PEERS];
};

__device__ __forceinline__ void st_release_sys_u64(uint64_t *p, uint64_t v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ uint64_t ld_acquire_sys_u64(const uint64_t *p) {
  uint64_t v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ uint64_t global_timer_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// release `epoch` into my slot on every peer (call from ONE thread, after the data is fenced)
__device__ __forceinline__ void peer_release_all(const PeerSlots &ps, uint64_t epoch) {
  __threadfence_system();
  for (int i = 0; i < ps.n; ++i) st_release_sys_u64(ps.slot[i], epoch);
}

// ONE thread: wait until every peer's slot in MY flag array holds >= epoch (ld.acquire.sys), with a timeout
__device__ __forceinline__ void peer_wait_all(const uint64_t *__restrict__ flags, const PeerSlots &ps, uint64_t epoch,
                                              uint64_t timeout_ns, int32_t *__restrict__ err) {
  const uint64_t t0 = global_timer_ns();
  for (int i = 0; i < ps.n; ++i) {
    const uint64_t *f = flags + ps.rank[i];
    unsigned spins = 0;
    while (ld_acquire_sys_u64(f) < epoch) {
      if ((++spins & 255u) == 0 && global_timer_ns() - t0 > timeout_ns) {
        atomicCAS(err, 0, ps.rank[i] + 1);
        break;
      }
    }
  }
  __threadfence_system();
}

// what the releasing thread does after the release when the launch also closes the cycle (null flags = nothing)
struct PeerWait {
  const uint64_t *flags;
  uint64_t timeout_ns;
  int32_t *err;
};

__global__ void k_peer_signal_wait(PeerSlots ps, uint64_t epoch, PeerWait w) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    peer_release_all(ps, epoch);
    if (w.flags) peer_wait_all(w.flags, ps, epoch, w.timeout_ns, w.err);
  }
}

__global__ void k_peer_signal(PeerSlots ps, uint64_t epoch) {
  if (threadIdx.x == 0 && blockIdx.x == 0) peer_release_all(ps, epoch);
}

// one warp: lane i waits for peer i
__global__ void k_peer_wait(const uint64_t *__restrict__ flags, PeerSlots ps, uint64_t epoch, uint64_t timeout_ns,
                            int32_t *__restrict__ err) {
  const int i = threadIdx.x;
  if (i < ps.n) {
    const uint64_t *f = flags + ps.rank[i];
    const uint64_t t0 = global_timer_ns();
    unsigned spins = 0;
    while (ld_acquire_sys_u64(f) < epoch) {
      if ((++spins & 255u) == 0 && global_timer_ns() - t0 > timeout_ns) {
        atomicCAS(err, 0, ps.rank[i] + 1);
        break;
      }
    }
  }
  __syncwarp();
  __threadfence_system();
}

// rows -> absolute (peer) addresses, one thread per (row, PB-byte piece): r rows first, then q rows.
// SIGNAL: the last block to finish releases `epoch` to every peer.
template <int PB, bool SIGNAL>
__global__ void __launch_bounds__(256)
k_halo_push_sig(const unsigned char *__restrict__ arr_r, const unsigned char *__restrict__ arr_q,
                const int64_t *__restrict__ off_r, const int64_t *__restrict__ off_q,
                const int64_t *__restrict__ dst_r, const int64_t *__restrict__ dst_q, int64_t n_r, int64_t n_q,
                int row_bytes, int elem, uint32_t *__restrict__ counter, PeerSlots ps, uint64_t epoch, PeerWait w) {
  const int ppr = row_bytes / PB;
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t row = t / ppr;
  const int piece = (int)(t - row * ppr);
  if (row < n_r + n_q) {
    const bool is_q = row >= n_r;
    const int64_t i = is_q ? row - n_r : row;
    const unsigned char *a = (is_q ? arr_q : arr_r) + (is_q ? off_q[i] : off_r[i]) * elem + piece * PB;
    unsigned char *b = reinterpret_cast<unsigned char *>(is_q ? dst_q[i] : dst_r[i]) + piece * PB;
    using V = typename std::conditional<PB == 16, uint4, typename std::conditional<PB == 8, uint2, uint32_t>::type>::type;
    *reinterpret_cast<V *>(b) = *reinterpret_cast<const V *>(a);
  }
  if (SIGNAL) {
    __threadfence_system();   // my stores are ordered before whatever follows the barrier below
    __syncthreads();
    if (threadIdx.x == 0) {
      const unsigned prev = atomicAdd(counter, 1u);
      if (prev == gridDim.x - 1) {
        *counter = 0;          // ready for the next launch (stream-ordered)
        peer_release_all(ps, epoch);
        if (w.flags) peer_wait_all(w.flags, ps, epoch, w.timeout_ns, w.err);
      }
    }
  }
}

// Run-based push of ONE list: thread <-> one 16-byte aligned unit of a destination run; the two 8-byte
// halves of a unit come from (possibly different) source rows.  A unit that sticks out of its run at
// either end is stored as its inner 8-byte half.
template <bool SIGNAL>
__global__ void __launch_bounds__(256)
k_halo_push_runs(const unsigned char *__restrict__ arr, const int64_t *__restrict__ src_off,
                 const int64_t *__restrict__ runs, int n_runs, int64_t total_units, uint32_t row_bytes, int elem,
                 uint32_t *__restrict__ counter, PeerSlots ps, uint64_t epoch, PeerWait w) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < total_units) {
    int lo = 0, hi = n_runs - 1;
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (runs[4 * mid + 3] <= t) lo = mid; else hi = mid - 1;
    }
    const int64_t a0 = runs[4 * lo], row0 = runs[4 * lo + 1], len = runs[4 * lo + 2];
    const int64_t unit = (a0 & ~(int64_t)15) + 16 * (t - runs[4 * lo + 3]);
    uint2 v[2];
    bool ok[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int64_t b = unit + 8 * h - a0;  // byte offset inside the run
      ok[h] = b >= 0 && b < len;
      if (ok[h]) {
        const uint32_t bb = (uint32_t)b, r = bb / row_bytes, c = bb - r * row_bytes;
        v[h] = *reinterpret_cast<const uint2 *>(arr + src_off[row0 + r] * elem + c);
      }
    }
    unsigned char *d = reinterpret_cast<unsigned char *>(unit);
    if (ok[0] && ok[1]) *reinterpret_cast<uint4 *>(d) = make_uint4(v[0].x, v[0].y, v[1].x, v[1].y);
    else if (ok[0]) *reinterpret_cast<uint2 *>(d) = v[0];
    else if (ok[1]) *reinterpret_cast<uint2 *>(d + 8) = v[1];
  }
  if (SIGNAL) {
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
      const unsigned prev = atomicAdd(counter, 1u);
      if (prev == gridDim.x - 1) {
        *counter = 0;
        peer_release_all(ps, epoch);
        if (w.flags) peer_wait_all(w.flags, ps, epoch, w.timeout_ns, w.err);   // chained mode: this launch closes the cycle
      }
    }
  }
}

inline PeerSlots peer_slots_of(const fg_peer_sync_t &s) {
  PeerSlots ps;
  ps.n = s.n_peers;
  for (int i = 0; i < FG_MAX_PEERS; ++i) {
    ps.rank[i] = i < s.n_peers ? s.peer_rank[i] : 0;
    ps.slot[i] = i < s.n_peers ? s.peer_slot[i] : nullptr;
  }
  return ps;
}

inline int peer_sync_check(const fg_peer_sync_t *s) {
  if (!s || s->n_peers < 0 || s->n_peers > FG_MAX_PEERS) return FG_ERR_ARG;
  if (s->n_peers && (!s->dev_flags || !s->dev_error)) return FG_ERR_ARG;
  for (int i = 0; i < s->n_peers; ++i)
    if (!s->peer_slot[i] || s->peer_rank[i] < 0) return FG_ERR_ARG;
  return FG_OK;
}

// push rows [r_lo, r_lo + n_r) of list r and [q_lo, q_lo + n_q) of list q of `plan` for buffer b;
// signal != 0: release `epoch` from the last block (or from a one-thread kernel when there is no row)
// wait != 0 (with signal): the releasing thread then waits for every peer's release of the same epoch — the launch
// closes the cycle (chained mode), no separate release / wait kernels
inline int halo_push_launch(const fg_halo_plan_t &p, const void *arr_r, const void *arr_q, int b, int64_t n_r,
                            int64_t n_q, int signal, uint64_t epoch, cudaStream_t st, int64_t &launches, int wait = 0) {
  const PeerSlots ps = peer_slots_of(p.sync);
  PeerWait w{nullptr, 0, nullptr};
  if (wait && signal && ps.n) w = PeerWait{p.sync.dev_flags, p.sync.timeout_ns ? p.sync.timeout_ns : 20000000000ull, p.sync.dev_error};
  // run tables present: one launch per list, 16-byte stores; the release rides on the last launch
  const bool runs_r = n_r > 0 && p.dev_runs_r[b] && p.n_runs_r[b] > 0, runs_q = n_q > 0 && p.dev_runs_q[b] && p.n_runs_q[b] > 0;
  if ((n_r <= 0 || runs_r) && (n_q <= 0 || runs_q) && (n_r > 0 || n_q > 0) && ((p.dom * p.elem_bytes) % 8 == 0)) {
    const bool sig = signal && ps.n;
    for (int list = 0; list < 2; ++list) {
      const bool is_q = list == 1;
      if (is_q ? !runs_q : !runs_r) continue;
      const bool last = is_q || !runs_q;
      const int64_t units = is_q ? p.units_q[b] : p.units_r[b];
      const unsigned blocks = (unsigned)((units + 255) / 256);
      const unsigned char *arr = (const unsigned char *)(is_q ? arr_q : arr_r);
      const int64_t *off = is_q ? p.dev_src_q_off : p.dev_src_r_off;
      const int64_t *runs = is_q ? p.dev_runs_q[b] : p.dev_runs_r[b];
      const int nr = is_q ? p.n_runs_q[b] : p.n_runs_r[b];
      if (sig && last)
        k_halo_push_runs<true><<<blocks, 256, 0, st>>>(arr, off, runs, nr, units, (uint32_t)(p.dom * p.elem_bytes), p.elem_bytes, p.dev_counter, ps, epoch, w);
      else
        k_halo_push_runs<false><<<blocks, 256, 0, st>>>(arr, off, runs, nr, units, (uint32_t)(p.dom * p.elem_bytes), p.elem_bytes, p.dev_counter, ps, epoch, w);
      ++launches;
    }
    return cudaGetLastError() == cudaSuccess ? FG_OK : FG_ERR_CUDA;
  }
  if (n_r + n_q <= 0) {
    if (signal && ps.n) { k_peer_signal_wait<<<1, 32, 0, st>>>(ps, epoch, w); ++launches; }
    return cudaGetLastError() == cudaSuccess ? FG_OK : FG_ERR_CUDA;
  }
  const int row_bytes = p.dom * p.elem_bytes;
  const int pb = (row_bytes % 16 == 0) ? 16 : ((row_bytes % 8 == 0) ? 8 : 4);
  const int64_t threads = (n_r + n_q) * (row_bytes / pb);
  const unsigned blocks = (unsigned)((threads + 255) / 256);
  const unsigned char *r = (const unsigned char *)arr_r, *q = (const unsigned char *)arr_q;
#define FG_PUSH(PB_, SIG_)                                                                                       \
  k_halo_push_sig<PB_, SIG_><<<blocks, 256, 0, st>>>(r, q, p.dev_src_r_off, p.dev_src_q_off, p.dev_dst_r[b],     \
                                                     p.dev_dst_q[b], n_r, n_q, row_bytes, p.elem_bytes,          \
                                                     p.dev_counter, ps, epoch, w)
  if (signal && ps.n) {
    if (pb == 16) FG_PUSH(16, true); else if (pb == 8) FG_PUSH(8, true); else FG_PUSH(4, true);
  } else {
    if (pb == 16) FG_PUSH(16, false); else if (pb == 8) FG_PUSH(8, false); else FG_PUSH(4, false);
  }
#undef FG_PUSH
  ++launches;
  return cudaGetLastError() == cudaSuccess ? FG_OK : FG_ERR_CUDA;
}

inline int peer_wait_launch(const fg_peer_sync_t &s, uint64_t epoch, cudaStream_t st, int64_t &launches) {
  if (!s.n_peers) return FG_OK;
  const uint64_t to = s.timeout_ns ? s.timeout_ns : 20000000000ull;
  k_peer_wait<<<1, 32, 0, st>>>(s.dev_flags, peer_slots_of(s), epoch, to, s.dev_error);
  ++launches;
  return cudaGetLastError() == cudaSuccess ? FG_OK : FG_ERR_CUDA;
}
