/* pydcop_b200 — C-ABI of the B200-native factor-graph message-passing engine.
 *
 * The reference (Orange-OpenSource/pyDcop) is pure Python and has NO FFI: its boundary for this
 * path is the algorithm-module surface `pydcop.algorithms.<name>` (GRAPH_TYPE / algo_params /
 * build_computation / computation_memory / communication_load,
 * pydcop/algorithms/__init__.py:508-566, pydcop/infrastructure/computations.py:1156-1165).
 * pydcop_b200/algorithms/{maxsum_gpu,dsa_gpu}.py implement that surface; THIS header is the
 * boundary one level below it: what a reference maintainer binds with ctypes to
 * replace, for all edges at once, the per-computation Python hot loops cited on each entry point.
 *
 * Conventions
 *   - plain C, opaque handles, int return codes (0 = FG_OK), no exceptions, no torch types;
 *   - every `dev_*` pointer is CALLER-OWNED DEVICE memory (the Python host passes
 *     torch.Tensor.data_ptr()); the only device memory the library owns is the DSA handle's copy
 *     of the class table (a few hundred bytes) and a MaxSum handle's class table of the variable side
 *     (48 bytes per (domain, degree) class); a MaxSum handle also owns one side stream + 2 events;
 *   - `stream` is a cudaStream_t passed as void* (torch.cuda.current_stream().cuda_stream);
 *   - all calls are asynchronous on `stream`; the caller synchronises;
 *   - value type T is float (precision = FG_F32) or double (FG_F64) for every cost/message array;
 *   - there is NO CPU fallback: without a CUDA device every launch entry point returns
 *     FG_ERR_CUDA and fg_last_error() says why.
 *
 * Data layout: factors are grouped into CLASSES of identical shape
 * (arity, domain sizes).  Inside a class everything is affine in the factor index f:
 *   table of f        = dev_tables + table_base + f * table_size          (row-major, axis i <->
 *                                                                           scope position i)
 *   message row (f,j) = msg_base + f * row_total + row_off[j], length dom[j]
 *                       (rows of a class whose positions have DIFFERENT domain sizes start on
 *                       4-element boundaries; uniform classes are dense: row_off[j] = j * dom)
 * (this is the layout of the factor->variable array r, written by the factor side).  The variable
 * side owns a CSR: variable v has slots [var_ptr[v], var_ptr[v+1]) in the reference's `links` order
 * (pydcop/algorithms/maxsum.py:466); the variable->factor array q is stored in SLOT order:
 *   q row of slot s   = var_qbase[v] + (s - var_ptr[v]) * dom_size[v]
 * so each side WRITES its own messages with perfectly coalesced stores and reads the other side's
 * through one gather index: edge_qoff[e] (factor side reads q), slot_roff[s] (variable side reads r).
 * Variables are stored class-major too (classes of identical (domain size, degree), fg_varclass_t);
 * dev_value / dev_value_cost are in that internal variable order (the Python host keeps the
 * permutation, pydcop_b200/layout.py).
 */
#ifndef PYDCOP_B200_H
#define PYDCOP_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FG_ABI_VERSION 1
#define FG_MAX_ARITY 8
#define FG_MAX_DOM 256

enum { FG_OK = 0, FG_ERR_ARG = 1, FG_ERR_CUDA = 2, FG_ERR_UNSUPPORTED = 3 };
enum { FG_F32 = 0, FG_F64 = 1 };
enum { FG_START_LEAFS = 0, FG_START_LEAFS_VARS = 1, FG_START_ALL = 2 }; /* maxsum.py:219 */
enum { FG_DSA_A = 0, FG_DSA_B = 1, FG_DSA_C = 2 };                      /* dsa.py:133 */
enum { FG_CLASS_GHOST = 1, FG_CLASS_BOUNDARY = 2 };

/* One class of same-shaped factors (constraints). */
typedef struct {
  int32_t arity;                 /* 1..FG_MAX_ARITY */
  int32_t dom[FG_MAX_ARITY];     /* domain size per scope position */
  int32_t row_off[FG_MAX_ARITY]; /* offset of position j's message row inside a factor's rows */
  int32_t row_total;             /* elements of one factor's rows: sum(dom), or with every row padded to 4 elements when the scope mixes domain sizes */
  int32_t n_factors;
  int32_t flags;                 /* FG_CLASS_GHOST: rows exist but the class is never computed
                                    (multi-GPU halo stubs, filled by the exchange) */
  int32_t first_factor;          /* global factor index of the first factor of the class */
  int32_t first_edge;            /* global edge index of its first edge; edge(f,j)=first_edge+f*arity+j */
  int64_t table_size;            /* prod(dom) */
  int64_t table_base;            /* element offset into dev_tables */
  int64_t msg_base;              /* element offset into the message arrays */
} fg_class_t;

/* One class of same-shaped variables: identical domain size and degree.  Variables are stored in
 * class-major ("internal") order; inside a class everything is affine in the variable's rank i:
 *   unary row   = dev_unary + unary_base + i * dom
 *   slots       = first_slot + i * degree + g            (g = 0..degree-1, the `links` order)
 *   q row (i,g) = q_base + (i * degree + g) * dom
 * degree == -1 marks the irregular class (variables of degree > 16): CSR through dev_var_ptr,
 * q rows still dense from q_base in slot order. */
typedef struct {
  int32_t dom, degree;
  int32_t n_vars, first_var, first_slot;
  int32_t n_slots;     /* total slots of the class */
  int32_t flags, reserved;  /* FG_CLASS_GHOST as above */
  int64_t unary_base;
  int64_t q_base;
} fg_varclass_t;

/* ------------------------------------------------------------------------------------------
 * MaxSum  (replaces MaxSumFactorComputation.on_new_cycle maxsum.py:339-379 +
 * factor_costs_for_var :382-447, MaxSumVariableComputation.on_new_cycle :525-565 +
 * select_value :584-620 + costs_for_factor :623-676, apply_damping :679-685, approx_match
 * :688-710, and both on_start methods :305-328, :495-523 — for ALL factors/variables at once)
 * ------------------------------------------------------------------------------------------ */
typedef struct {
  int32_t abi_version; /* FG_ABI_VERSION */
  int32_t precision;   /* FG_F32 | FG_F64 */
  int32_t n_vars, n_factors, n_edges, n_classes;
  int64_t n_msg_r;     /* elements of each r buffer (class-major edge order, padded bases) */
  int64_t n_msg_q;     /* elements of each q buffer (slot order) */
  int32_t uniform_dom; /* D if every variable has domain size D, else 0 (enables fast kernels) */
  int32_t max_degree;  /* largest number of factors on one variable */
  const fg_class_t *classes; /* HOST array [n_classes], copied by fg_maxsum_create */
  int32_t n_varclasses, reserved0;
  const fg_varclass_t *varclasses; /* HOST array [n_varclasses], copied by fg_maxsum_create */

  /* problem (device, read-only) */
  const void *dev_tables;       /* T[sum table_size] */
  const void *dev_unary;        /* T[sum dom_size]  variable costs (+noise), maxsum.py:477-487 */
  const int32_t *dev_dom_size;  /* [n_vars] */
  const int64_t *dev_unary_off; /* [n_vars+1] */
  const int32_t *dev_var_ptr;   /* [n_vars+1] */
  const int64_t *dev_var_qbase; /* [n_vars+1] q offset of the first slot of v */
  const int64_t *dev_slot_roff; /* [n_edges] slot order: offset in r of the row of slot s's edge */
  const int64_t *dev_edge_qoff; /* [n_edges] edge order: offset in q of the row of edge e */
  const uint32_t *dev_slot_roff32; /* same as 32-bit offsets, or NULL when they do not fit */
  const uint32_t *dev_edge_qoff32;
  const int32_t *dev_slot_edge; /* [n_edges] edge id of slot s */
  const int32_t *dev_slot_var;  /* [n_edges] variable of slot s */
  const int32_t *dev_init_value; /* [n_vars] initial_value index or -1 (maxsum.py:497-500) */

  /* state (device, read-write).  q = variable->factor, r = factor->variable; [0]/[1] are the
   * Jacobi double buffers, the engine tracks which one is current. */
  void *dev_q[2], *dev_r[2];       /* T[n_msg_q] / T[n_msg_r] each */
  uint8_t *dev_q_valid, *dev_r_valid; /* [n_edges] edge order: receiver holds a message */
  uint8_t *dev_q_cnt;              /* [n_edges] SLOT order: bit0 has-prev, bits1.. send count */
  uint8_t *dev_r_cnt;              /* [n_edges] edge order: same encoding */
  uint8_t *dev_q_sent, *dev_r_sent; /* [n_edges] q_sent SLOT order, r_sent edge order: message
                                       posted in the last cycle (may be NULL: not recorded) */
  int32_t *dev_value;              /* [n_vars] selected value index (select_value) */
  void *dev_value_cost;            /* T[n_vars] its cost */

  /* algorithm parameters (maxsum.py:212-220) */
  int32_t mode_max;      /* 0 'min', 1 'max' */
  int32_t damp_vars, damp_factors; /* damping_nodes in {vars,both} / {factors,both} */
  int32_t start_messages; /* FG_START_* */
  double damping, stability;
} fg_maxsum_desc_t;

typedef struct fg_maxsum *fg_maxsum_t;

int fg_abi_version(void);
/* Number of CUDA devices visible to the library (0 when none / driver missing). */
int fg_device_count(void);

int fg_maxsum_create(const fg_maxsum_desc_t *desc, fg_maxsum_t *out);
int fg_maxsum_destroy(fg_maxsum_t h);
const char *fg_maxsum_last_error(fg_maxsum_t h);
/* cycle 0: every computation's on_start (maxsum.py:305-328, 495-523). */
int fg_maxsum_init(fg_maxsum_t h, void *stream);
/* n_cycles synchronous cycles (one = every factor's and every variable's on_new_cycle). */
int fg_maxsum_step(fg_maxsum_t h, int32_t n_cycles, void *stream);
/* Split-phase variant for the multi-GPU host loop: launch the compute of ONE cycle (writes the
 * `next` buffers), let the caller exchange halos on them, then commit (swap buffers). */
int fg_maxsum_cycle_compute(fg_maxsum_t h, void *stream);
int fg_maxsum_cycle_commit(fg_maxsum_t h);
/* Index (0/1) of the CURRENT message buffers and the number of cycles done so far. */
int fg_maxsum_current(fg_maxsum_t h, int32_t *buf_index, int64_t *cycle);
/* Number of kernel launches issued by this handle so far (bench.py's gpu_launches). */
int64_t fg_maxsum_launch_count(fg_maxsum_t h);
/* Which kernel family computes each factor class from cycle 2 on (diagnostic; -1 = ghost class, never computed):
 * generic one-thread-per-edge, round-1 CTA-pipelined (compile-time shape), warp-autonomous (binary, even domain),
 * runtime-dimension tiled (any other shape whose table fits shared memory). */
enum { FG_KERNEL_GENERIC = 0, FG_KERNEL_PIPE = 1, FG_KERNEL_WARP = 2, FG_KERNEL_TILED_RT = 3 };
int fg_maxsum_kernel_plan(fg_maxsum_t h, int32_t *family, int32_t n_classes);

/* Gather / scatter of boundary message rows for the halo exchange (replaces
 * Messaging.post_msg for cut edges, pydcop/infrastructure/communication.py:588-698).
 * rows: `n_rows` (offset,len) pairs in dev_row_off/dev_row_len; packed back to back. */
int fg_halo_pack(int32_t precision, const void *dev_src, void *dev_packed,
                 const int64_t *dev_row_off, const int64_t *dev_packed_off,
                 const int32_t *dev_row_len, int64_t n_rows, void *stream);
int fg_halo_unpack(int32_t precision, void *dev_dst, const void *dev_packed,
                   const int64_t *dev_row_off, const int64_t *dev_packed_off,
                   const int32_t *dev_row_len, int64_t n_rows, void *stream);

/* Uniform-domain fast form of the two calls above, r and q rows in ONE launch.  The packed
 * buffer is the all_to_all buffer: for peer p the block [r rows for p | q rows for p] starts at
 * element peer_base[p]; rows of the r list [peer_r_start[p], peer_r_start[p+1]) and of the q list
 * [peer_q_start[p], peer_q_start[p+1]) belong to peer p (HOST arrays of n_peers+1 / n_peers
 * entries, n_peers <= 16).  pack != 0: arrays -> packed; pack == 0: packed -> arrays. */
int fg_halo_rows_uniform(int32_t precision, int32_t pack, void *dev_r, void *dev_q, void *dev_packed,
                         const int64_t *dev_row_off_r, const int64_t *dev_row_off_q, int64_t n_r,
                         int64_t n_q, int32_t dom, int32_t n_peers, const int64_t *peer_r_start,
                         const int64_t *peer_q_start, const int64_t *peer_base, void *stream);

/* Direct halo push over NVLink peer memory (uniform domain): row i of the r list
 * (dev_r + row_off_r[i], `dom` elements) is stored at the ABSOLUTE device address dst_r[i] — a
 * location inside the consuming rank's `next` buffer, mapped into this process through CUDA IPC —
 * and likewise for the q list.  One launch replaces pack -> NCCL all_to_all -> unpack; the caller
 * closes the cycle with a barrier. */
/* Enable stores from the CURRENT device into `peer_device`'s memory (cudaDeviceEnablePeerAccess);
 * returns FG_ERR_UNSUPPORTED when the pair has no peer path. */
int fg_enable_peer_access(int32_t peer_device);
/* CUDA IPC for the peer push: export = IPC handle (64 bytes) of the allocation containing dev_ptr
 * and dev_ptr's byte offset inside it; import = map a peer's allocation into THIS process with the
 * current device as accessor (cudaIpcMemLazyEnablePeerAccess), returning its base address. */
int fg_ipc_export(const void *dev_ptr, unsigned char handle_out[64], int64_t *offset_out);
int fg_ipc_import(const unsigned char handle[64], void **base_out);
int fg_ipc_close(void *base);
int fg_halo_push(int32_t precision, const void *dev_r, const void *dev_q, const int64_t *dev_row_off_r,
                 const int64_t *dev_row_off_q, const int64_t *dev_dst_r, const int64_t *dev_dst_q,
                 int64_t n_r, int64_t n_q, int32_t dom, void *stream);

/* ------------------------------------------------------------------------------------------
 * DSA  (replaces DsaComputation.on_start dsa.py:277-299 and evaluate_cycle :320-357 with
 * find_optimal relations.py:1594-1638, assignment_cost :1479-1532, variant_a/b/c dsa.py:359-405,
 * probabilistic_change :407-417, exists_violated_constraint :419-431, find_optimum
 * relations.py:1367-1400 — for ALL variables at once).  Random draws are Philox4x32-10 keyed by
 * (seed; variable, cycle): u = 53-bit uniform, choice = (word*n)>>32 (oracle/philox.py).
 * ------------------------------------------------------------------------------------------ */
typedef struct {
  int32_t abi_version, precision;
  int32_t n_vars, n_factors, n_edges, n_classes;
  const fg_class_t *classes;     /* HOST [n_classes] */
  const void *dev_tables;        /* T[...] */
  const int32_t *dev_dom_size;   /* [n_vars] */
  const int32_t *dev_var_id;     /* [n_vars] caller's (canonical) id of each internal variable: the
                                    Philox counter, so draws do not depend on the internal order */
  const int32_t *dev_edge_var;   /* [n_edges] variable of edge e (class-major edge order) */
  const int32_t *dev_edge_class; /* [n_edges] class of edge e */
  const int32_t *dev_var_ptr;    /* [n_vars+1] */
  const int32_t *dev_slot_edge;  /* [n_edges] incident edges of v in node.constraints order */
  const uint8_t *dev_has_nbr;    /* [n_vars] 1: variable has >= 1 neighbour (dsa.py:278); 0: isolated, its
                                    value is carried over; 2: ghost of a variable another rank owns (multi-GPU):
                                    never written by the kernels, filled by the owner's push / halo unpack */
  const double *dev_prob;        /* [n_vars] change threshold (p_mode fixed|arity, dsa.py:252-263) */
  void *dev_con_opt;             /* T[n_factors] per-constraint optimum, filled by fg_dsa_init */
  /* optional fast path: every constraint binary over one domain size `fast_dom` (else NULL / 0).
   * Per slot (variable v, incident constraint c, neighbour u): the table of c ORIENTED so that
   * row y = value of u is contiguous over v's values (position-0 slots read a transposed copy):
   *   cost_v[x] += dev_tables_or[slot_tab[s] + y * fast_dom + x] */
  const void *dev_tables_or;     /* T[...] original tables followed by the transposed copies */
  const int32_t *dev_slot_nbr;   /* [n_edges] neighbour variable of slot s */
  const int64_t *dev_slot_tab;   /* [n_edges] element offset of the oriented table of slot s */
  const void *dev_slot_opt;      /* T[n_edges] optimum of the slot's constraint (variant B) */
  int32_t fast_dom, reserved1;
  int32_t *dev_value[2];         /* [n_vars] current / next value index (double buffer) */
  void *dev_value_cost;          /* T[n_vars] cost reported with the last selection */
  int32_t mode_max, variant;     /* FG_DSA_* */
  int32_t stop_cycle;            /* 0 = never (dsa.py:134,352) */
  uint64_t seed;
  /* A-DSA (pydcop/algorithms/adsa.py:344-377): non-NULL = the variables' own costs T[sum dom] (internal variable
   * order, row of v at dev_unary_off[v]) are added to every CANDIDATE's cost; the current cost (adsa.py:262) and
   * DSA itself (the variable-cost branch of find_optimal is dead code, relations.py:1630) do not use them. */
  const void *dev_var_cost;
  const int64_t *dev_unary_off;  /* [n_vars+1] */
  /* optional, fast path only (csrc/dsa_cached.cuh): the row each slot read last, slot-major T[n_edges * fast_dom],
   * and the neighbour value it belongs to (0xFF = none yet; reset by fg_dsa_init).  A cycle streams these rows and
   * goes to the oriented tables only for the slots whose neighbour changed; results are identical. */
  void *dev_row_cache;
  uint8_t *dev_slot_last;        /* [n_edges] */
} fg_dsa_desc_t;

typedef struct fg_dsa *fg_dsa_t;

int fg_dsa_create(const fg_dsa_desc_t *desc, fg_dsa_t *out);
int fg_dsa_destroy(fg_dsa_t h);
const char *fg_dsa_last_error(fg_dsa_t h);
/* on_start: per-constraint optima + random initial values (isolated variables keep the value
 * the host stored in dev_value[0], dsa.py:278-289). */
int fg_dsa_init(fg_dsa_t h, void *stream);
/* up to n_cycles evaluate_cycle rounds (stops at stop_cycle). */
int fg_dsa_step(fg_dsa_t h, int32_t n_cycles, void *stream);
int fg_dsa_cycle_compute(fg_dsa_t h, void *stream);
int fg_dsa_cycle_commit(fg_dsa_t h);
int fg_dsa_current(fg_dsa_t h, int32_t *buf_index, int64_t *cycle);
int64_t fg_dsa_launch_count(fg_dsa_t h);

/* ------------------------------------------------------------------------------------------
 * Device-side cycle barrier + fused peer push (multi-GPU; replaces the per-cycle NCCL all_reduce and
 * the host-driven compute -> push -> barrier chain; reference counterpart: the cycle_id handshake of
 * SynchronousComputationMixin, pydcop/infrastructure/computations.py:696-718, and Messaging.post_msg
 * for cut edges, communication.py:588-698).
 *
 * Every rank owns an array of `world` 64-bit EPOCH FLAGS in its own device memory, mapped into the
 * peers through CUDA IPC; slot p is written only by rank p.  After its boundary rows of a cycle are
 * stored in the peers' buffers, a rank releases (fence.sys + st.release.sys) the new epoch into its
 * slot on every peer; before it reads rows the peers wrote it acquires (ld.acquire.sys, spinning)
 * the same epoch from every peer's slot in its own array.  No host round trip, no collective.
 * A wait that does not complete within `timeout_ns` stores peer rank + 1 in *dev_error and returns,
 * so a lost peer cannot hang the device.
 * ------------------------------------------------------------------------------------------ */
#define FG_MAX_PEERS 16
typedef struct {
  int32_t n_peers;                    /* ranks this rank exchanges with (<= FG_MAX_PEERS) */
  int32_t my_rank;
  const uint64_t *dev_flags;          /* MY flag array [world], zeroed by the caller at creation */
  int32_t peer_rank[FG_MAX_PEERS];    /* slot of dev_flags written by peer i */
  uint64_t *peer_slot[FG_MAX_PEERS];  /* address, as mapped into THIS process, of slot my_rank in
                                         peer i's flag array */
  int32_t *dev_error;                 /* one int32, zeroed by the caller */
  uint64_t timeout_ns;                /* 0 = 20 s */
} fg_peer_sync_t;

int fg_peer_signal(const fg_peer_sync_t *ps, uint64_t epoch, void *stream);
int fg_peer_wait(const fg_peer_sync_t *ps, uint64_t epoch, void *stream);

/* Rows a rank pushes every cycle: row i of list r is `dom` elements at dev_r + src_r_off[i] and is
 * stored at the absolute (peer) address dst_r[b][i] when the cycle writes buffer b; likewise q.
 * For DSA the "rows" are single 4-byte values (dom = 1, list r only). */
typedef struct {
  int32_t elem_bytes, dom;
  int64_t n_r, n_q;
  const int64_t *dev_src_r_off, *dev_src_q_off;
  const int64_t *dev_dst_r[2], *dev_dst_q[2];
  uint32_t *dev_counter;              /* one zeroed uint32: last-block detection of the push kernel */
  fg_peer_sync_t sync;
  /* Optional (NULL / 0 = one thread per 4/8/16-byte piece of a row): the push lists cut into RUNS of rows
   * whose destinations are consecutive addresses — the rows one peer receives land in dense blocks of its
   * buffers — so that the kernel can issue whole 16-byte aligned stores over NVLink whatever the row
   * size (rows of 40 bytes would otherwise leave as 8-byte stores).  Per list and buffer index b:
   * dev_runs[b] = int64 [n_runs][4] = (first destination address, index of the run's first row in the
   * list, bytes in the run, index of the run's first 16-byte destination unit), units = total units.
   * Needs row bytes % 8 == 0. */
  const int64_t *dev_runs_r[2], *dev_runs_q[2];
  int32_t n_runs_r[2], n_runs_q[2];
  int64_t units_r[2], units_q[2];
  /* Optional, MaxSum only (NULL = off): FUSED halo.  Per buffer index b, for every edge (class-major order) the
   * address in its consumer's r[b] of the f->v row the factor side produces, and for every slot the address in
   * its consumer's q[b] of the v->f row the variable side produces; 0 = the row stays on this GPU.  When every
   * non-ghost class of the shard runs on the warp-autonomous kernels (csrc/maxsum_warp.cuh), those kernels store
   * each boundary row to its peer from the lane that produced it — the transfer overlaps the rest of the cycle's
   * arithmetic tile by tile and the two push launches disappear; otherwise the push kernels above are used. */
  const int64_t *dev_edge_dst_r[2], *dev_slot_dst_q[2];
} fg_halo_plan_t;

/* Attach a push plan to an engine (copied).  From then on fg_*_shard_step runs whole cycles on the
 * device without the host: compute -> push boundary rows into the peers' `next` buffers, the last
 * block of the push kernel releasing epoch+1 -> wait for every peer's epoch+1 -> commit.  The epoch
 * counter only grows (it survives fg_*_init), so every rank must run the same number of cycles. */
int fg_maxsum_shard_attach(fg_maxsum_t h, const fg_halo_plan_t *plan);
/* 1 when the attached plan runs in FUSED mode (boundary rows stored by the warp kernels), else 0. */
int fg_maxsum_shard_fused(fg_maxsum_t h);
int fg_maxsum_shard_step(fg_maxsum_t h, int32_t n_cycles, void *stream);
/* One phase of a cycle, for timing breakdowns: 0 compute, 1 push + signal, 2 wait, 3 commit. */
int fg_maxsum_shard_phase(fg_maxsum_t h, int32_t phase, void *stream);
/* Diagnostics: n_cycles whole cycles with CUDA timing events INSIDE the cycle (no host work between the
 * kernels), averaged, microseconds: out_us[0] factor side, [1] push of the r rows (split mode),
 * [2] variable side, [3] push of the q rows (split mode), [4] push / release after the join, [5] wait
 * for the peers, [6] whole cycle.  Synchronises the stream after every cycle. */
int fg_maxsum_shard_profile(fg_maxsum_t h, int32_t n_cycles, void *stream, double *out_us);
int fg_dsa_shard_attach(fg_dsa_t h, const fg_halo_plan_t *plan);
int fg_dsa_shard_step(fg_dsa_t h, int32_t n_cycles, void *stream);

/* ------------------------------------------------------------------------------------------
 * MGM  (next-tier row §8f.4; replaces MgmComputation.on_start mgm.py:283-310, the value phase
 * _handle_value_message :343-397 with _compute_best_value :434-455 and find_arg_optimal
 * relations.py:1554-1591, and the gain phase _handle_gain_message :497-537 with the lexicographic
 * tie break :574-591 — for ALL variables at once).  Same constraint-hypergraph arrays as DSA.
 * One cycle = two launches: every variable's best local gain, then every variable's decision.
 * Draws: Philox4x32-10 keyed (seed; canonical variable id, cycle) — cycle 0xffffffff for the
 * initial value, else the 1-based round (the reference's cycle_count when it draws).
 * ------------------------------------------------------------------------------------------ */
typedef struct {
  int32_t abi_version, precision;
  int32_t n_vars, n_factors, n_edges, n_classes;
  const fg_class_t *classes;      /* HOST [n_classes] */
  const void *dev_tables;         /* T[...] */
  const void *dev_unary;          /* T[...] variable costs (cost_for_val) */
  const int64_t *dev_unary_off;   /* [n_vars+1] */
  const int32_t *dev_dom_size;    /* [n_vars] */
  const int32_t *dev_var_id;      /* [n_vars] canonical id = Philox counter */
  const int32_t *dev_var_rank;    /* [n_vars] position of the variable's NAME in sorted order */
  const int32_t *dev_edge_var;    /* [n_edges] */
  const int32_t *dev_edge_class;  /* [n_edges] */
  const int32_t *dev_var_ptr;     /* [n_vars+1] */
  const int32_t *dev_slot_edge;   /* [n_edges] incident edges of v in node.constraints order */
  const int32_t *dev_nbr_ptr;     /* [n_vars+1] distinct neighbours (mgm.py:245-252) */
  const int32_t *dev_nbr_idx;     /* [nbr_ptr[n_vars]] */
  const int32_t *dev_init_value;  /* [n_vars] value index or -1 (random), may be NULL */
  int32_t *dev_value;             /* [n_vars] in/out; variables without neighbours are preset by the
                                     host (optimal_cost_value, relations.py:1641-1669) */
  void *dev_cost;                 /* T[n_vars] current_cost (preset for isolated variables) */
  uint8_t *dev_has_cost;          /* [n_vars] current_cost is not None */
  void *dev_gain;                 /* T[n_vars] */
  int32_t *dev_new_value;         /* [n_vars] */
  int32_t mode_max;
  int32_t stop_cycle;             /* 0 = never; a run ends when round + 1 >= stop_cycle (mgm.py:404) */
  uint64_t seed;
  /* optional fast shape (opt-in experiment): every constraint binary over one domain size `fast_dom`
   * in {4, 8, 10, 16, 20}; the oriented tables of the DSA fast path (fg_dsa_desc_t) and the number of
   * incidences handled per trip (2 | 4).  All NULL / 0 = the generic value-phase kernel. */
  const void *dev_tables_or;      /* T[...] */
  const int32_t *dev_slot_nbr;    /* [n_edges] */
  const int64_t *dev_slot_tab;    /* [n_edges] */
  int32_t fast_dom, fast_chunk;
  /* optional with the fast shape (csrc/mgm_cached_kernels.cuh): the row each slot read last, slot-major
   * T[n_edges * fast_dom], and the neighbour value it belongs to (0xFF = none yet; reset by fg_mgm_init): the value
   * phase streams these rows and reads the oriented tables only where a neighbour moved; results are identical. */
  void *dev_row_cache;
  uint8_t *dev_slot_last;         /* [n_edges] */
} fg_mgm_desc_t;

typedef struct fg_mgm *fg_mgm_t;

int fg_mgm_create(const fg_mgm_desc_t *desc, fg_mgm_t *out);
int fg_mgm_destroy(fg_mgm_t h);
const char *fg_mgm_last_error(fg_mgm_t h);
/* on_start for the variables that have neighbours: initial_value or an injected random choice. */
int fg_mgm_init(fg_mgm_t h, void *stream);
/* up to n_cycles value+gain rounds (stops as the reference does at stop_cycle). */
int fg_mgm_step(fg_mgm_t h, int32_t n_cycles, void *stream);
int fg_mgm_current(fg_mgm_t h, int64_t *cycle, int32_t *finished);
int64_t fg_mgm_launch_count(fg_mgm_t h);

/* Solution cost (next-tier row §8f.1; pydcop/dcop/dcop.py:319-367 `solution_cost`, called by the
 * orchestrator at the end of a run and at every metric tick, orchestrator.py:1229-1231): over the factors
 * of every non-ghost class the table entry at the assignment, over the first n_vars variables
 * dev_unary[unary_off[v] + value[v]]; an entry EQUAL to `infinity` (compared in the tables' precision) is
 * counted as a violation, every other entry is summed: dev_out[0] = cost (double), dev_out[1] =
 * violations.  dev_unary holds the variables' OWN costs in internal variable order (without MaxSum's
 * noise, which is not part of the problem's cost); NULL skips the variable costs.  A shard counts only
 * what it owns — n_vars leading variables, and the optional byte masks dev_factor_skip[internal factor]
 * / dev_var_skip[internal variable] (non-zero = not mine; NULL = none) — and the caller all-reduces
 * dev_out over the ranks. */
int fg_solution_cost(int32_t precision, int32_t n_classes, const fg_class_t *classes,
                     const void *dev_tables, const int32_t *dev_edge_var,
                     const int32_t *dev_value, const void *dev_unary,
                     const int64_t *dev_unary_off, int32_t n_vars, const uint8_t *dev_factor_skip,
                     const uint8_t *dev_var_skip, double infinity, double *dev_out, void *stream);

/* Diagnostic: evaluates the send-gate predicate approx_match (maxsum.py:688-710) on n pairs with
 * the division-free fast form used by the tiled kernels (out_fast) and the literal form
 * (out_exact); the two must agree on every input.  T arrays of n elements, uint8 outputs. */
int fg_selftest_approx_match(int32_t precision, int64_t n, const void *dev_c, const void *dev_prev,
                             double stability, uint8_t *dev_out_fast, uint8_t *dev_out_exact,
                             void *stream);

/* Diagnostic: HBM throughput of the DSA / MGM table access pattern alone — `n_threads` threads each read
 * `rows_per_thread` (1, 2, 3 or 6) rows of `row_bytes` (multiple of 16) at pseudo-random multiples of `stride_bytes`
 * inside [dev_base, dev_base + region_bytes), all loads independent; dev_out[n_threads] receives a checksum.
 * Timed by the caller (tools/gather_peak.py): the ceiling those kernels are measured against in DESIGN.md. */
int fg_selftest_gather(const void *dev_base, int64_t region_bytes, int32_t row_bytes, int32_t stride_bytes,
                       int64_t n_threads, int32_t rows_per_thread, float *dev_out, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* PYDCOP_B200_H */
