"""Python host of the B200 engine: device memory (torch), streams, and the C-ABI calls.

`MaxSumEngine` / `DsaEngine` are what the pyDcop plugin modules (pydcop_b200/algorithms/) and
bench.py drive.  All arithmetic on messages happens in libpydcop_b200.so (hand-written sm_100a
kernels); torch only owns the buffers.  There is no CPU path: constructing an engine without a
CUDA device raises EngineError.
"""
import ctypes as C
import os
from typing import Optional

import numpy as np
import torch

from . import _cabi
from ._cabi import EngineError, FgClass, FgDsaDesc, FgMaxSumDesc, FgMgmDesc, FgVarClass
from .layout import FactorGraphLayout

PRECISIONS = {"f32": (_cabi.FG_F32, torch.float32, np.float32),
              "f64": (_cabi.FG_F64, torch.float64, np.float64)}


def _require_cuda(device):
    if not torch.cuda.is_available():
        raise EngineError("no CUDA device: pydcop_b200 runs the MaxSum/DSA hot path on the GPU "
                          "only (no CPU fallback)")
    dev = torch.device(device if device is not None else "cuda")
    if dev.type != "cuda":
        raise EngineError(f"device must be a CUDA device, got {dev}")
    if dev.index is None:
        dev = torch.device("cuda", torch.cuda.current_device())
    return dev


def _class_array(layout: FactorGraphLayout):
    arr = (FgClass * max(1, len(layout.classes)))()
    for i, c in enumerate(layout.classes):
        fc = arr[i]
        fc.arity = c.arity
        for j in range(c.arity):
            fc.dom[j] = c.dom[j]
            fc.row_off[j] = c.row_off[j]
        fc.row_total = c.row_total
        fc.n_factors, fc.first_factor, fc.first_edge = c.n_factors, c.first_factor, c.first_edge
        fc.flags = 1 if c.tag else 0
        fc.table_size, fc.table_base, fc.msg_base = c.table_size, c.table_base, c.msg_base
    return arr


def _varclass_array(layout: FactorGraphLayout):
    arr = (FgVarClass * max(1, len(layout.var_classes)))()
    for i, c in enumerate(layout.var_classes):
        vc = arr[i]
        vc.dom, vc.degree, vc.n_vars = c.dom, c.degree, c.n_vars
        vc.first_var, vc.first_slot, vc.n_slots = c.first_var, c.first_slot, c.n_slots
        vc.unary_base, vc.q_base = c.unary_base, c.q_base
        vc.flags = (1 if c.tag == 2 else 0) | (2 if c.tag == 1 else 0)   # FG_CLASS_GHOST | FG_CLASS_BOUNDARY (multigpu.build_shard)
    return arr


def _ptr(t: Optional[torch.Tensor]):
    return C.c_void_p(t.data_ptr() if t is not None and t.numel() else 0) if t is not None \
        else C.c_void_p(0)


class _EngineBase:
    def _dev(self, a, dtype):
        a = np.ascontiguousarray(a)
        if not a.flags.writeable:      # e.g. a read-only memory map of the binary container
            a = a.copy()
        t = torch.from_numpy(a).to(device=self.device, dtype=dtype)
        self.h2d_bytes = getattr(self, "h2d_bytes", 0) + t.numel() * t.element_size()   # what crossed PCIe / C2C
        return t

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _check(self, rc, what):
        if rc != _cabi.FG_OK:
            raise EngineError(f"{what} failed (rc={rc}): {self._last_error()}")

    def _solution_cost(self, value_tensor, infinity=float("inf"), unary=None, n_vars=None, factor_skip=None,
                       var_skip=None):
        """(cost, violations) of `value_tensor` (internal variable order) as LOCAL sums, on the device:
        pydcop/dcop/dcop.py:319-367 — an entry equal to `infinity` is a violation, the others are
        summed, constraints and variable costs alike.  `unary`: the variables' own costs in CANONICAL
        order (default: none are added); n_vars: leading variables that count (a shard's own ones);
        factor_skip / var_skip: uint8 device tensors in internal order, non-zero = owned by another rank."""
        L = self.layout
        prec = PRECISIONS[self.precision][0]
        tdt = PRECISIONS[self.precision][1]
        with torch.cuda.device(self.device):
            if not hasattr(self, "_cost_out"):
                self._edge_var_dev = self._dev(L.edge_var, torch.int32)
                self._cost_out = torch.zeros(2, dtype=torch.float64, device=self.device)
                self._cost_unary_off = self._dev(L.unary_off, torch.int64)
            un = None
            if unary is not None:   # canonical (variable-major) -> internal (class-major, padded bases)
                un_host = np.zeros(max(int(L.unary_off[-1]), 1))
                cu = np.asarray(unary, dtype=np.float64).reshape(-1)
                dom = L.dom_size.astype(np.int64)
                c_off = np.concatenate([[0], np.cumsum(dom[L.var_perm])])[:-1]   # canonical offsets
                if len(cu):
                    src = np.repeat(c_off[L.var_order], dom) + (np.arange(int(dom.sum())) - np.repeat(np.cumsum(dom) - dom, dom))
                    dst = np.repeat(L.unary_off[:-1], dom) + (np.arange(int(dom.sum())) - np.repeat(np.cumsum(dom) - dom, dom))
                    un_host[dst] = cu[src]
                un = self._dev(un_host, tdt)
            rc = self.lib.fg_solution_cost(
                prec, len(L.classes), C.cast(self._classes, C.POINTER(FgClass)), _ptr(self.tables),
                _ptr(self._edge_var_dev), _ptr(value_tensor), _ptr(un), _ptr(self._cost_unary_off),
                int(L.n_vars if n_vars is None else n_vars), _ptr(factor_skip), _ptr(var_skip), float(infinity),
                _ptr(self._cost_out), self._stream())
            self._check(rc, "fg_solution_cost")
            return self._cost_out


class MaxSumEngine(_EngineBase):
    """All-edges-at-once synchronous MaxSum.

    Parameters mirror the reference's algo_params (pydcop/algorithms/maxsum.py:212-220); `noise`
    is applied by the caller to the instance's `unary` before packing (pydcop_b200.ingest.add_noise).
    State after `init()` is the reference's cycle 0 (on_start); each `step()` cycle is one
    synchronous round of every factor's and every variable's on_new_cycle.
    """

    def __init__(self, layout: FactorGraphLayout, device=None, precision="f32", mode="min",
                 damping=0.5, damping_nodes="both", stability=0.1, start_messages="leafs",
                 record_sent=True):
        self.lib = _cabi.load()
        self.device = _require_cuda(device)
        self.layout = L = layout
        self.precision = precision
        prec, tdt, self.np_dtype = PRECISIONS[precision]
        if damping_nodes not in ("vars", "factors", "both", "none"):
            raise ValueError(f"invalid damping_nodes {damping_nodes!r}")
        if start_messages not in _cabi.START_MESSAGES:
            raise ValueError(f"invalid start_messages {start_messages!r}")
        if mode not in ("min", "max"):
            raise ValueError(f"invalid mode {mode!r}")
        with torch.cuda.device(self.device):
            self.tables = self._dev(L.tables, tdt)
            self.unary = self._dev(L.unary, tdt)
            self.dom_size = self._dev(L.dom_size, torch.int32)
            self.unary_off = self._dev(L.unary_off, torch.int64)
            self.var_ptr = self._dev(L.var_ptr, torch.int32)
            self.var_qbase = self._dev(L.var_qbase, torch.int64)
            self.slot_roff = self._dev(L.slot_roff, torch.int64)
            self.edge_qoff = self._dev(L.edge_qoff, torch.int64)
            fits32 = max(L.n_msg, L.n_msg_q) < 2 ** 32
            # uint32 offsets, stored as int32 bit patterns (torch has no uint32 arithmetic)
            self.slot_roff32 = self._dev(L.slot_roff.astype(np.uint32).view(np.int32), torch.int32) if fits32 else None
            self.edge_qoff32 = self._dev(L.edge_qoff.astype(np.uint32).view(np.int32), torch.int32) if fits32 else None
            self.slot_edge = self._dev(L.slot_edge, torch.int32)
            self.slot_var = self._dev(L.slot_var, torch.int32)
            self.init_value = self._dev(L.init_value, torch.int32)
            z = lambda n, dt: torch.zeros(max(int(n), 1), dtype=dt, device=self.device)  # noqa
            self.q = [z(L.n_msg_q, tdt), z(L.n_msg_q, tdt)]
            self.r = [z(L.n_msg, tdt), z(L.n_msg, tdt)]
            self.q_valid, self.r_valid = z(L.n_edges, torch.uint8), z(L.n_edges, torch.uint8)
            self.q_cnt, self.r_cnt = z(L.n_edges, torch.uint8), z(L.n_edges, torch.uint8)
            self.q_sent = z(L.n_edges, torch.uint8) if record_sent else None
            self.r_sent = z(L.n_edges, torch.uint8) if record_sent else None
            self.value = z(L.n_vars, torch.int32)
            self.value_cost = z(L.n_vars, tdt)
        self._classes = _class_array(L)
        d = FgMaxSumDesc()
        d.abi_version, d.precision = _cabi.FG_ABI_VERSION, prec
        d.n_vars, d.n_factors, d.n_edges = L.n_vars, L.n_factors, L.n_edges
        d.n_classes, d.n_msg_r, d.n_msg_q = len(L.classes), L.n_msg, L.n_msg_q
        d.uniform_dom, d.max_degree = L.uniform_dom, L.max_degree
        d.classes = C.cast(self._classes, C.POINTER(FgClass))
        self._varclasses = _varclass_array(L)
        d.n_varclasses = len(L.var_classes)
        d.varclasses = C.cast(self._varclasses, C.POINTER(FgVarClass))
        d.dev_tables, d.dev_unary = _ptr(self.tables), _ptr(self.unary)
        d.dev_dom_size, d.dev_unary_off = _ptr(self.dom_size), _ptr(self.unary_off)
        d.dev_var_ptr, d.dev_var_qbase = _ptr(self.var_ptr), _ptr(self.var_qbase)
        d.dev_slot_roff, d.dev_edge_qoff = _ptr(self.slot_roff), _ptr(self.edge_qoff)
        d.dev_slot_roff32, d.dev_edge_qoff32 = _ptr(self.slot_roff32), _ptr(self.edge_qoff32)
        d.dev_slot_edge, d.dev_slot_var = _ptr(self.slot_edge), _ptr(self.slot_var)
        d.dev_init_value = _ptr(self.init_value)
        for b in range(2):
            d.dev_q[b], d.dev_r[b] = self.q[b].data_ptr(), self.r[b].data_ptr()
        d.dev_q_valid, d.dev_r_valid = _ptr(self.q_valid), _ptr(self.r_valid)
        d.dev_q_cnt, d.dev_r_cnt = _ptr(self.q_cnt), _ptr(self.r_cnt)
        d.dev_q_sent, d.dev_r_sent = _ptr(self.q_sent), _ptr(self.r_sent)
        d.dev_value, d.dev_value_cost = _ptr(self.value), _ptr(self.value_cost)
        d.mode_max = int(mode == "max")
        d.damp_vars = int(damping_nodes in ("vars", "both"))
        d.damp_factors = int(damping_nodes in ("factors", "both"))
        d.start_messages = _cabi.START_MESSAGES[start_messages]
        d.damping, d.stability = float(damping), float(stability)
        self._desc = d
        self._h = C.c_void_p()
        rc = self.lib.fg_maxsum_create(C.byref(d), C.byref(self._h))
        self._check(rc, "fg_maxsum_create")

    def _last_error(self):
        return (self.lib.fg_maxsum_last_error(self._h) or b"").decode()

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self.lib.fg_maxsum_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- driving -----------------------------------------------------------------------------
    def init(self):
        with torch.cuda.device(self.device):
            self._check(self.lib.fg_maxsum_init(self._h, self._stream()), "fg_maxsum_init")
        return self

    def step(self, n_cycles=1):
        with torch.cuda.device(self.device):
            self._check(self.lib.fg_maxsum_step(self._h, int(n_cycles), self._stream()),
                        "fg_maxsum_step")
        return self

    def cycle_compute(self):
        with torch.cuda.device(self.device):
            self._check(self.lib.fg_maxsum_cycle_compute(self._h, self._stream()),
                        "fg_maxsum_cycle_compute")

    def cycle_commit(self):
        self._check(self.lib.fg_maxsum_cycle_commit(self._h), "fg_maxsum_cycle_commit")

    @property
    def cur(self):
        b, c = C.c_int32(), C.c_int64()
        self.lib.fg_maxsum_current(self._h, C.byref(b), C.byref(c))
        return b.value

    @property
    def cycle(self):
        b, c = C.c_int32(), C.c_int64()
        self.lib.fg_maxsum_current(self._h, C.byref(b), C.byref(c))
        return c.value

    @property
    def launch_count(self):
        return int(self.lib.fg_maxsum_launch_count(self._h))

    KERNEL_FAMILIES = {-1: "ghost", 0: "generic", 1: "pipe", 2: "warp", 3: "tiled_rt"}

    def kernel_plan(self):
        """Kernel family of every factor class (fg_maxsum_kernel_plan), as a list of names."""
        n = len(self.layout.classes)
        out = (C.c_int32 * max(n, 1))()
        rc = self.lib.fg_maxsum_kernel_plan(self._h, out, n)
        if rc != 0:
            raise EngineError(f"fg_maxsum_kernel_plan rc={rc}")
        return [self.KERNEL_FAMILIES[int(out[i])] for i in range(n)]

    # -- readback (canonical edge order) -----------------------------------------------------
    def messages(self):
        """(q, r) receiver-side message state as float64 numpy arrays in canonical edge order."""
        g = torch.from_numpy(self.layout.message_gather_index()).to(self.device)
        gq = torch.from_numpy(self.layout.message_gather_index_q()).to(self.device)
        cur = self.cur
        q = self.q[cur][gq].double().cpu().numpy() if g.numel() else np.zeros(0)
        r = self.r[cur][g].double().cpu().numpy() if g.numel() else np.zeros(0)
        return q, r

    def flags(self):
        """dict of canonical-edge-order numpy arrays: q_valid, r_valid, q_sent, r_sent, q_cnt, r_cnt."""
        L = self.layout
        out = {"q_valid": L.edges_to_canonical(self.q_valid.cpu().numpy()[:L.n_edges]),
               "r_valid": L.edges_to_canonical(self.r_valid.cpu().numpy()[:L.n_edges]),
               "r_cnt": L.edges_to_canonical(self.r_cnt.cpu().numpy()[:L.n_edges]),
               "q_cnt": L.slots_to_canonical_edges(self.q_cnt.cpu().numpy()[:L.n_edges])}
        if self.q_sent is not None:
            out["q_sent"] = L.slots_to_canonical_edges(self.q_sent.cpu().numpy()[:L.n_edges])
            out["r_sent"] = L.edges_to_canonical(self.r_sent.cpu().numpy()[:L.n_edges])
        return out

    def solution_cost(self, infinity=float("inf"), unary=None):
        """(cost, violations) of the currently selected assignment, reduced on the device
        (pydcop/dcop/dcop.py:319-367; `infinity` as `pydcop solve -i`).  `unary`: the variables' own
        costs in canonical order WITHOUT MaxSum's noise (default: variable costs are not added)."""
        out = self._solution_cost(self.value, infinity, unary).cpu().numpy()
        return float(out[0]), int(out[1])

    def values(self):
        """(value index, reported cost) per variable, in the caller's canonical variable order."""
        L = self.layout
        n = L.n_vars
        return (L.vars_to_canonical(self.value[:n].cpu().numpy()),
                L.vars_to_canonical(self.value_cost[:n].double().cpu().numpy()))


def dsa_row_stride(d: int, elem: int) -> int:
    """Elements between two rows of an oriented table: the next power of two >= d, whole 128-byte lines above 128
    bytes (csrc/row_load.cuh::fg_row_stride, the same rule): the one row a slot reads per cycle then costs one or two
    whole DRAM lines."""
    p2 = 1
    while p2 < d:
        p2 *= 2
    line = 128 // elem
    return -(-d // line) * line if p2 * elem > 128 else p2


def dsa_fast_arrays(layout: FactorGraphLayout, tables: torch.Tensor, mode="min"):
    """Arrays of the DSA fast path, or None when the instance does not qualify (every constraint
    binary over ONE domain size).  Per slot (variable v, incident constraint c, neighbour u) the
    table of c is read ORIENTED so that row y = value of u is contiguous over v's values: slots at
    scope position 1 read the table as stored, slots at position 0 read a transposed copy; rows are
    stored with stride dsa_row_stride (each row on its own cache-line slot).
    Returns (tables_or, slot_tab, slot_nbr, slot_opt, D) as tensors on `tables.device`."""
    L = layout
    if not (len(L.classes) == 1 and L.classes[0].arity == 2 and L.classes[0].dom[0] == L.classes[0].dom[1]
            and L.n_edges and not L.classes[0].tag):
        return None
    dev = tables.device
    c0 = L.classes[0]
    D = c0.dom[0]
    S = D * D
    nF = c0.n_factors
    t = tables[c0.table_base:c0.table_base + nF * S].view(nF, D, D)
    RS = dsa_row_stride(D, t.element_size())
    SP = D * RS                                     # elements per padded table
    tables_or = torch.zeros(2 * nF * SP, dtype=t.dtype, device=dev)
    tv = tables_or.view(2, nF, D, RS)
    tv[0, :, :, :D] = t
    tv[1, :, :, :D] = t.transpose(1, 2)
    opt = (t.reshape(nF, S).max(dim=1).values if mode == "max" else t.reshape(nF, S).min(dim=1).values)
    e = L.slot_edge.astype(np.int64) - c0.first_edge
    f, j = e // 2, e % 2
    # position 1 (me second): T[y][x] is already row-contiguous; position 0: transposed copy
    slot_tab = np.where(j == 1, f * SP, nF * SP + f * SP).astype(np.int64)
    slot_nbr = L.edge_var[c0.first_edge + f * 2 + (1 - j)].astype(np.int32)
    return (tables_or, torch.from_numpy(np.ascontiguousarray(slot_tab)).to(dev),
            torch.from_numpy(np.ascontiguousarray(slot_nbr)).to(dev),
            opt[torch.from_numpy(f).to(dev)].contiguous(), D)


class DsaEngine(_EngineBase):
    """All-variables-at-once DSA-A/B/C (pydcop/algorithms/dsa.py:130-135 parameters)."""

    def __init__(self, layout: FactorGraphLayout, device=None, precision="f32", mode="min",
                 probability=0.7, p_mode="fixed", variant="B", stop_cycle=0, seed=0,
                 isolated_value=None, var_global_id=None, frozen=None, var_costs=False):
        """var_costs: A-DSA's decision (pydcop/algorithms/adsa.py:344-377: every candidate value carries the
        variable's own cost, the current cost does not); False: DSA.
        var_global_id / frozen (canonical order) serve the multi-GPU partition
        (pydcop_b200/multigpu_dsa.py): the Philox counter of each variable when the layout is a
        shard of a larger problem, and the ghost variables whose value is only copied through."""
        self.lib = _cabi.load()
        self.device = _require_cuda(device)
        self.layout = L = layout
        self.precision = precision
        prec, tdt, self.np_dtype = PRECISIONS[precision]
        if variant not in _cabi.DSA_VARIANTS:
            raise ValueError(f"invalid variant {variant!r}")
        if p_mode not in ("fixed", "arity"):
            raise ValueError(f"invalid p_mode {p_mode!r}")
        arity = np.array([c.arity for c in L.classes], dtype=np.int64)
        n_count = np.zeros(L.n_vars, dtype=np.int64)
        if L.n_edges:
            np.add.at(n_count, L.slot_var, arity[L.edge_class[L.slot_edge]] - 1)
        has_nbr = (n_count > 0).astype(np.uint8)
        if frozen is not None:   # 2 = ghost of another rank's variable: never written by the local kernels
            has_nbr[np.asarray(frozen, dtype=bool)[L.var_order]] = 2
        var_id = (L.var_order if var_global_id is None
                  else np.asarray(var_global_id, dtype=np.int32)[L.var_order])
        if p_mode == "arity":  # dsa.py:257-260: 1 / n_count * 1.2
            with np.errstate(divide="ignore"):
                prob = np.where(n_count > 0, 1.0 / np.maximum(n_count, 1) * 1.2, 0.0)
        else:
            prob = np.full(L.n_vars, float(probability))
        # isolated variables (dsa.py:278-289): argopt of (own cost, value), tuple order
        if isolated_value is not None:   # given in canonical order
            isolated_value = np.asarray(isolated_value, dtype=np.int32)[L.var_order]
        else:
            isolated_value = np.zeros(L.n_vars, dtype=np.int32)
            for v in np.nonzero(n_count == 0)[0]:  # (frozen ghosts are filled by the exchange)
                c = L.unary[L.unary_off[v]:L.unary_off[v] + L.dom_size[v]]
                if mode == "min":
                    isolated_value[v] = int(np.argmin(c))
                else:
                    isolated_value[v] = int(len(c) - 1 - np.argmax(c[::-1]))
        self.has_nbr_host, self.n_count = has_nbr, n_count
        with torch.cuda.device(self.device):
            self.tables = self._dev(L.tables, tdt)
            self.dom_size = self._dev(L.dom_size, torch.int32)
            self.var_id = self._dev(var_id, torch.int32)
            self.edge_var = self._dev(L.edge_var, torch.int32)
            self.edge_class = self._dev(L.edge_class, torch.int32)
            self.var_ptr = self._dev(L.var_ptr, torch.int32)
            self.slot_edge = self._dev(L.slot_edge, torch.int32)
            self.has_nbr = self._dev(has_nbr, torch.uint8)
            self.prob = self._dev(prob, torch.float64)
            z = lambda n, dt: torch.zeros(max(int(n), 1), dtype=dt, device=self.device)  # noqa
            self.con_opt = z(L.n_factors, tdt)
            self.value = [self._dev(isolated_value, torch.int32) if L.n_vars else z(1, torch.int32),
                          z(L.n_vars, torch.int32)]
            self.value_cost = z(L.n_vars, tdt)
            self.var_cost = self._dev(L.unary, tdt) if var_costs else None
            self.unary_off = self._dev(L.unary_off, torch.int64) if var_costs else None
        # fast path: every constraint binary over ONE domain size -> oriented tables, one contiguous
        # row per incidence (transposed copy for scope position 0)
        self.tables_or = self.slot_nbr = self.slot_tab = self.slot_opt = None
        fast_dom = 0
        with torch.cuda.device(self.device):
            fast = dsa_fast_arrays(L, self.tables, mode)
        self.row_cache = self.slot_last = None
        if fast is not None:
            self.tables_or, self.slot_tab, self.slot_nbr, self.slot_opt, fast_dom = fast
            with torch.cuda.device(self.device):   # active-row array (csrc/dsa_cached.cuh): L.n_edges rows of fast_dom costs
                self.row_cache = torch.zeros(max(L.n_edges * fast_dom, 4), dtype=tdt, device=self.device)
                self.slot_last = torch.full((max(L.n_edges, 1),), 255, dtype=torch.uint8, device=self.device)
        self._classes = _class_array(L)
        d = FgDsaDesc()
        d.abi_version, d.precision = _cabi.FG_ABI_VERSION, prec
        d.n_vars, d.n_factors, d.n_edges, d.n_classes = L.n_vars, L.n_factors, L.n_edges, len(L.classes)
        d.classes = C.cast(self._classes, C.POINTER(FgClass))
        d.dev_tables, d.dev_dom_size = _ptr(self.tables), _ptr(self.dom_size)
        d.dev_var_id = _ptr(self.var_id)
        d.dev_edge_var, d.dev_edge_class = _ptr(self.edge_var), _ptr(self.edge_class)
        d.dev_var_ptr, d.dev_slot_edge = _ptr(self.var_ptr), _ptr(self.slot_edge)
        d.dev_has_nbr, d.dev_prob, d.dev_con_opt = _ptr(self.has_nbr), _ptr(self.prob), _ptr(self.con_opt)
        d.dev_value[0], d.dev_value[1] = self.value[0].data_ptr(), self.value[1].data_ptr()
        d.dev_value_cost = _ptr(self.value_cost)
        d.dev_tables_or, d.dev_slot_nbr = _ptr(self.tables_or), _ptr(self.slot_nbr)
        d.dev_slot_tab, d.dev_slot_opt = _ptr(self.slot_tab), _ptr(self.slot_opt)
        d.fast_dom = fast_dom
        d.mode_max, d.variant = int(mode == "max"), _cabi.DSA_VARIANTS[variant]
        d.stop_cycle, d.seed = int(stop_cycle), int(seed) & (2 ** 64 - 1)
        d.dev_var_cost, d.dev_unary_off = _ptr(self.var_cost), _ptr(self.unary_off)
        d.dev_row_cache, d.dev_slot_last = _ptr(self.row_cache), _ptr(self.slot_last)
        self._desc = d
        self._h = C.c_void_p()
        self._check(self.lib.fg_dsa_create(C.byref(d), C.byref(self._h)), "fg_dsa_create")

    def _last_error(self):
        return (self.lib.fg_dsa_last_error(self._h) or b"").decode()

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self.lib.fg_dsa_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def init(self):
        with torch.cuda.device(self.device):
            self._check(self.lib.fg_dsa_init(self._h, self._stream()), "fg_dsa_init")
        return self

    def step(self, n_cycles=1):
        with torch.cuda.device(self.device):
            self._check(self.lib.fg_dsa_step(self._h, int(n_cycles), self._stream()), "fg_dsa_step")
        return self

    def cycle_compute(self):
        with torch.cuda.device(self.device):
            self._check(self.lib.fg_dsa_cycle_compute(self._h, self._stream()), "fg_dsa_cycle_compute")

    def cycle_commit(self):
        self._check(self.lib.fg_dsa_cycle_commit(self._h), "fg_dsa_cycle_commit")

    def _current(self):
        b, c = C.c_int32(), C.c_int64()
        self.lib.fg_dsa_current(self._h, C.byref(b), C.byref(c))
        return b.value, c.value

    @property
    def cur(self):
        return self._current()[0]

    @property
    def cycle(self):
        return self._current()[1]

    @property
    def launch_count(self):
        return int(self.lib.fg_dsa_launch_count(self._h))

    def solution_cost(self, infinity=float("inf"), unary=None):
        """(cost, violations) of the current assignment on the device (dcop.py:319-367)."""
        out = self._solution_cost(self.value[self.cur], infinity, unary).cpu().numpy()
        return float(out[0]), int(out[1])

    def values(self):
        """Current value index per variable, canonical variable order."""
        L = self.layout
        return L.vars_to_canonical(self.value[self.cur][:L.n_vars].cpu().numpy())


def distinct_neighbours(layout: FactorGraphLayout):
    """CSR (nbr_ptr, nbr_idx) of every variable's DISTINCT neighbours (the other variables of its
    constraints, mgm.py:245-252) in internal variable ids, each list in order of first appearance
    over the variable's constraints (node.constraints order, then scope order) — the order the
    oracle uses, so neighbour costs are added in the same sequence."""
    L = layout
    V, E = L.n_vars, L.n_edges
    if not E:
        return np.zeros(V + 1, dtype=np.int32), np.zeros(1, dtype=np.int32)
    arity = np.array([c.arity for c in L.classes], dtype=np.int64)
    first_edge = np.array([c.first_edge for c in L.classes], dtype=np.int64)
    e = L.slot_edge.astype(np.int64)
    cl = L.edge_class[e].astype(np.int64)
    a = arity[cl]
    e0 = first_edge[cl] + ((e - first_edge[cl]) // a) * a
    A = int(arity.max())
    vs, seqs, us = [], [], []
    slot = np.arange(E, dtype=np.int64)
    for j in range(A):
        m = a > j
        u = L.edge_var[(e0 + j)[m]].astype(np.int64)
        v = L.slot_var[m].astype(np.int64)
        keep = u != v
        vs.append(v[keep])
        us.append(u[keep])
        seqs.append((slot[m] * A + j)[keep])
    v, u, seq = np.concatenate(vs), np.concatenate(us), np.concatenate(seqs)
    o = np.lexsort((seq, u, v))            # by (v, u), earliest occurrence first
    v, u, seq = v[o], u[o], seq[o]
    first = np.ones(len(v), dtype=bool)
    first[1:] = (v[1:] != v[:-1]) | (u[1:] != u[:-1])
    v, u, seq = v[first], u[first], seq[first]
    o = np.lexsort((seq, v))               # back to first-appearance order inside each variable
    v, u = v[o], u[o]
    ptr = np.zeros(V + 1, dtype=np.int32)
    np.cumsum(np.bincount(v, minlength=V), out=ptr[1:])
    return ptr, (u.astype(np.int32) if len(u) else np.zeros(1, dtype=np.int32))


def mgm_host_arrays(layout: FactorGraphLayout, mode="min", var_rank=None, isolated_value=None):
    """Everything MgmEngine derives on the host, in the layout's INTERNAL variable order: the
    neighbour CSR, name ranks, and the preset value / cost of the variables without neighbours
    (mgm.py:285-294: optimal_cost_value, then finished)."""
    L = layout
    nbr_ptr, nbr_idx = distinct_neighbours(L)
    has_nbr = np.diff(nbr_ptr) > 0
    rank = (np.arange(L.n_vars, dtype=np.int32) if var_rank is None
            else np.asarray(var_rank, dtype=np.int32))[L.var_order]
    value0 = np.zeros(L.n_vars, dtype=np.int32)
    cost0 = np.zeros(L.n_vars, dtype=np.float64)
    iso = np.nonzero(~has_nbr)[0]
    if isolated_value is not None:
        value0[iso] = np.asarray(isolated_value, dtype=np.int32)[L.var_order][iso]
    for v in iso:
        c = L.unary[L.unary_off[v]:L.unary_off[v] + L.dom_size[v]]
        if isolated_value is None:
            value0[v] = int(np.argmin(c)) if mode == "min" else int(len(c) - 1 - np.argmax(c[::-1]))
        cost0[v] = c[value0[v]]
    return dict(nbr_ptr=nbr_ptr, nbr_idx=nbr_idx, has_nbr=has_nbr, var_rank=rank.astype(np.int32),
                value0=value0, cost0=cost0)


class MgmEngine(_EngineBase):
    """All-variables-at-once MGM (pydcop/algorithms/mgm.py; parameters :78-81).

    var_rank       canonical order: position of each variable's NAME in sorted order, the
                   lexicographic tie break of mgm.py:574-583 (default: the canonical index)
    isolated_value canonical order: value index of the variables without neighbours
                   (optimal_cost_value over the real domain values, relations.py:1641-1669);
                   default: argopt of the own cost with ties broken on the value index
    `break_mode` is accepted for the reference's signature; its 'random' branch is dead code in
    the reference (mgm.py:541 compares with the `random` module) and both settings act as 'lexic'.
    """

    def __init__(self, layout: FactorGraphLayout, device=None, precision="f32", mode="min",
                 stop_cycle=0, seed=0, break_mode="lexic", var_rank=None, isolated_value=None):
        self.lib = _cabi.load()
        self.device = _require_cuda(device)
        self.layout = L = layout
        self.precision = precision
        prec, tdt, self.np_dtype = PRECISIONS[precision]
        if break_mode not in ("lexic", "random"):
            raise ValueError(f"invalid break_mode {break_mode!r}")
        if mode not in ("min", "max"):
            raise ValueError(f"invalid mode {mode!r}")
        h = mgm_host_arrays(L, mode, var_rank, isolated_value)
        nbr_ptr, nbr_idx, has_nbr, rank = h["nbr_ptr"], h["nbr_idx"], h["has_nbr"], h["var_rank"]
        value0, cost0 = h["value0"], h["cost0"]
        self.has_nbr_host = has_nbr
        with torch.cuda.device(self.device):
            self.tables = self._dev(L.tables, tdt)
            self.unary = self._dev(L.unary, tdt)
            self.unary_off = self._dev(L.unary_off, torch.int64)
            self.dom_size = self._dev(L.dom_size, torch.int32)
            self.var_id = self._dev(L.var_order, torch.int32)
            self.var_rank = self._dev(rank, torch.int32)
            self.edge_var = self._dev(L.edge_var, torch.int32)
            self.edge_class = self._dev(L.edge_class, torch.int32)
            self.var_ptr = self._dev(L.var_ptr, torch.int32)
            self.slot_edge = self._dev(L.slot_edge, torch.int32)
            self.nbr_ptr = self._dev(nbr_ptr, torch.int32)
            self.nbr_idx = self._dev(nbr_idx, torch.int32)
            self.init_value = self._dev(L.init_value, torch.int32)
            n = max(L.n_vars, 1)
            self.value = self._dev(np.resize(value0, n) if L.n_vars else np.zeros(1, np.int32), torch.int32)
            self.cost = self._dev(np.resize(cost0, n) if L.n_vars else np.zeros(1), tdt)
            self.has_cost = self._dev((~np.resize(has_nbr, n)).astype(np.uint8) if L.n_vars
                                      else np.zeros(1, np.uint8), torch.uint8)
            self.gain = torch.zeros(n, dtype=tdt, device=self.device)
            self.new_value = torch.zeros(n, dtype=torch.int32, device=self.device)
        self._classes = _class_array(L)
        d = FgMgmDesc()
        d.abi_version, d.precision = _cabi.FG_ABI_VERSION, prec
        d.n_vars, d.n_factors, d.n_edges, d.n_classes = L.n_vars, L.n_factors, L.n_edges, len(L.classes)
        d.classes = C.cast(self._classes, C.POINTER(FgClass))
        d.dev_tables, d.dev_unary, d.dev_unary_off = _ptr(self.tables), _ptr(self.unary), _ptr(self.unary_off)
        d.dev_dom_size, d.dev_var_id, d.dev_var_rank = _ptr(self.dom_size), _ptr(self.var_id), _ptr(self.var_rank)
        d.dev_edge_var, d.dev_edge_class = _ptr(self.edge_var), _ptr(self.edge_class)
        d.dev_var_ptr, d.dev_slot_edge = _ptr(self.var_ptr), _ptr(self.slot_edge)
        d.dev_nbr_ptr, d.dev_nbr_idx = _ptr(self.nbr_ptr), _ptr(self.nbr_idx)
        d.dev_init_value = _ptr(self.init_value)
        d.dev_value, d.dev_cost, d.dev_has_cost = _ptr(self.value), _ptr(self.cost), _ptr(self.has_cost)
        d.dev_gain, d.dev_new_value = _ptr(self.gain), _ptr(self.new_value)
        d.mode_max, d.stop_cycle, d.seed = int(mode == "max"), int(stop_cycle), int(seed) & (2 ** 64 - 1)
        # the fast value-phase kernel on the DSA fast-path arrays (binary constraints over one domain size
        # in {4, 8, 10, 16, 20}); 2 incidences per trip measured fastest on the B200 (C4 instance: 605 us
        # per cycle vs 1869 us generic, profiles/r02_call1_pending_summary.txt).  PYDCOP_B200_MGM_FAST=0|2|4
        self.fast_chunk = 0
        chunk = int(os.environ.get("PYDCOP_B200_MGM_FAST", "2") or 0)
        if chunk in (2, 4):
            with torch.cuda.device(self.device):
                fast = dsa_fast_arrays(L, self.tables, mode)
            if fast is not None and fast[4] in (4, 8, 10, 16, 20):
                self.tables_or, self.slot_tab, self.slot_nbr, _, fast_dom = fast
                d.dev_tables_or, d.dev_slot_nbr, d.dev_slot_tab = (_ptr(self.tables_or), _ptr(self.slot_nbr),
                                                                   _ptr(self.slot_tab))
                d.fast_dom, d.fast_chunk = fast_dom, chunk
                self.fast_chunk = chunk
                self.row_cache = self.slot_last = None
                if os.environ.get("PYDCOP_B200_MGM_CACHE", "1") != "0":   # active-row array (csrc/mgm_cached_kernels.cuh)
                    with torch.cuda.device(self.device):
                        self.row_cache = torch.zeros(max(L.n_edges * fast_dom, 4), dtype=tdt, device=self.device)
                        self.slot_last = torch.full((max(L.n_edges, 1),), 255, dtype=torch.uint8, device=self.device)
                    d.dev_row_cache, d.dev_slot_last = _ptr(self.row_cache), _ptr(self.slot_last)
        self._desc = d
        self._h = C.c_void_p()
        self._check(self.lib.fg_mgm_create(C.byref(d), C.byref(self._h)), "fg_mgm_create")

    def _last_error(self):
        return (self.lib.fg_mgm_last_error(self._h) or b"").decode()

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self.lib.fg_mgm_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def init(self):
        with torch.cuda.device(self.device):
            self._check(self.lib.fg_mgm_init(self._h, self._stream()), "fg_mgm_init")
        return self

    def step(self, n_cycles=1):
        with torch.cuda.device(self.device):
            self._check(self.lib.fg_mgm_step(self._h, int(n_cycles), self._stream()), "fg_mgm_step")
        return self

    def _current(self):
        c, f = C.c_int64(), C.c_int32()
        self.lib.fg_mgm_current(self._h, C.byref(c), C.byref(f))
        return c.value, bool(f.value)

    @property
    def cycle(self):
        return self._current()[0]

    @property
    def finished(self):
        return self._current()[1]

    @property
    def launch_count(self):
        return int(self.lib.fg_mgm_launch_count(self._h))

    def values(self):
        """(value index, current cost or NaN while the reference's is None) per variable,
        canonical variable order."""
        L = self.layout
        n = L.n_vars
        cost = self.cost[:n].double().cpu().numpy()
        cost[self.has_cost[:n].cpu().numpy() == 0] = np.nan
        return L.vars_to_canonical(self.value[:n].cpu().numpy()), L.vars_to_canonical(cost)

    def solution_cost(self, infinity=float("inf"), unary=None):
        """(cost, violations) of the current assignment on the device (dcop.py:319-367)."""
        out = self._solution_cost(self.value, infinity, unary).cpu().numpy()
        return float(out[0]), int(out[1])

    def gains(self):
        """(gain, intended value) of the last round per variable, canonical order."""
        L = self.layout
        n = L.n_vars
        return (L.vars_to_canonical(self.gain[:n].double().cpu().numpy()),
                L.vars_to_canonical(self.new_value[:n].cpu().numpy()))
