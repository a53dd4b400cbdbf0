"""ctypes front-end of the CPU oracle (oracle/dcop_oracle.c).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` leg may
import this module; the product package (pydcop_b200/) never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "_build", "libdcop_oracle.so")

FLAG_RECV, FLAG_PREV = 1, 2
START_MESSAGES = {"leafs": 0, "leafs_vars": 1, "all": 2}
VARIANTS = {"A": 0, "B": 1, "C": 2}


class _FG(C.Structure):
    _fields_ = [("V", C.c_int32), ("F", C.c_int32), ("E", C.c_int32),
                ("dom_size", C.c_void_p), ("factor_ptr", C.c_void_p), ("edge_var", C.c_void_p),
                ("table_off", C.c_void_p), ("var_ptr", C.c_void_p), ("var_edge", C.c_void_p),
                ("msg_off", C.c_void_p), ("unary_off", C.c_void_p)]


def build(force: bool = False) -> str:
    src = [os.path.join(HERE, f) for f in ("dcop_oracle.c", "dcop_oracle_impl.h", "Makefile")]
    if force or not os.path.exists(LIB_PATH) or any(
            os.path.getmtime(s) > os.path.getmtime(LIB_PATH) for s in src):
        subprocess.run(["make", "-C", HERE, "-B"], check=True, capture_output=True)
    return LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
    return _lib


def _p(a):
    return C.c_void_p(a.ctypes.data)


def threads(set_to: int = 0) -> int:
    """OpenMP threads of the oracle's parallel loops: set (set_to > 0) and / or query."""
    fn = lib().oracle_threads
    fn.restype, fn.argtypes = C.c_int, [C.c_int]
    return int(fn(int(set_to)))


class _Graph:
    """Flat factor-graph arrays shared by both oracles (same arrays as the engine front door)."""

    def __init__(self, inst, dtype):
        self.dtype = np.dtype(dtype)
        self.sfx = "_f64" if self.dtype == np.float64 else "_f32"
        c = np.ascontiguousarray
        self.dom_size = c(inst["dom_size"], dtype=np.int32)
        self.factor_ptr = c(inst["factor_ptr"], dtype=np.int32)
        self.edge_var = c(inst["edge_var"], dtype=np.int32)
        if "table_off" in inst and inst["table_off"] is not None:
            self.table_off = c(inst["table_off"], dtype=np.int64)
        else:  # dense tables back to back, sizes from the scopes' domain sizes
            ar = np.diff(self.factor_ptr)
            ts = np.ones(len(ar), dtype=np.int64)
            np.multiply.at(ts, np.repeat(np.arange(len(ar)), ar),
                           self.dom_size[self.edge_var].astype(np.int64))
            self.table_off = np.concatenate([[0], np.cumsum(ts)]).astype(np.int64)
        self.var_ptr = c(inst["var_ptr"], dtype=np.int32)
        self.var_edge = c(inst["var_edge"], dtype=np.int32)
        self.tables = c(inst["tables"], dtype=self.dtype)
        self.V, self.F, self.E = len(self.dom_size), len(self.factor_ptr) - 1, len(self.edge_var)
        d = self.dom_size[self.edge_var] if self.E else np.zeros(0, np.int32)
        self.msg_off = np.zeros(self.E + 1, dtype=np.int64)
        np.cumsum(d, out=self.msg_off[1:])
        self.unary_off = np.zeros(self.V + 1, dtype=np.int64)
        np.cumsum(self.dom_size, out=self.unary_off[1:])
        self.M = int(self.msg_off[-1])
        assert int(self.dom_size.max(initial=1)) <= 256, "oracle MAX_DOM"
        assert int(np.diff(self.factor_ptr).max(initial=1)) <= 8, "oracle MAX_ARITY"
        self.fg = _FG(self.V, self.F, self.E, _p(self.dom_size), _p(self.factor_ptr),
                      _p(self.edge_var), _p(self.table_off), _p(self.var_ptr), _p(self.var_edge),
                      _p(self.msg_off), _p(self.unary_off))


class MaxSumOracle(_Graph):
    """Lock-step MaxSum.  State after `init()` == golden state 0, after n `step()` == state n."""

    def __init__(self, inst, dtype=np.float64, mode="min", damping=0.5, damping_nodes="both",
                 stability=0.1, start_messages="leafs", **_ignored):
        super().__init__(inst, dtype)
        self.unary = np.ascontiguousarray(inst["unary"], dtype=self.dtype)
        iv = inst.get("init_value") if hasattr(inst, "get") else (
            inst["init_value"] if "init_value" in inst else None)
        self.init_value = (np.ascontiguousarray(iv, dtype=np.int32) if iv is not None
                           else np.full(self.V, -1, np.int32))
        self.mode_max = int(mode == "max")
        self.damping, self.stability = float(damping), float(stability)
        self.damp_vars = int(damping_nodes in ("vars", "both"))
        self.damp_factors = int(damping_nodes in ("factors", "both"))
        self.start_messages = START_MESSAGES[start_messages]
        self.q = np.zeros(self.M, self.dtype)
        self.r = np.zeros(self.M, self.dtype)
        self.q_flags = np.zeros(self.E, np.uint8)
        self.r_flags = np.zeros(self.E, np.uint8)
        self.q_sent = np.zeros(self.E, np.uint8)
        self.r_sent = np.zeros(self.E, np.uint8)
        self.value = np.zeros(self.V, np.int32)
        self.value_cost = np.zeros(self.V, self.dtype)
        self.cycle = 0

    def init(self):
        getattr(lib(), "maxsum_oracle_init" + self.sfx)(
            C.byref(self.fg), _p(self.tables), _p(self.unary), _p(self.init_value),
            self.mode_max, self.start_messages, _p(self.q), _p(self.r), _p(self.q_flags),
            _p(self.r_flags), _p(self.q_sent), _p(self.r_sent), _p(self.value),
            _p(self.value_cost))
        self.cycle = 0
        return self

    def step(self, n=1):
        """n cycles in ONE C call (scratch allocated once, states swapped between cycles)."""
        if n <= 0:
            return self
        fn = getattr(lib(), "maxsum_oracle_steps" + self.sfx)
        fn(C.byref(self.fg), _p(self.tables), _p(self.unary), self.mode_max, self.damp_vars,
           self.damp_factors, C.c_double(self.damping), C.c_double(self.stability),
           _p(self.q), _p(self.r), _p(self.q_flags), _p(self.r_flags), _p(self.q_sent),
           _p(self.r_sent), _p(self.value), _p(self.value_cost), C.c_int(int(n)))
        self.cycle += int(n)
        return self


class DsaOracle(_Graph):
    """Lock-step DSA with injected Philox draws (oracle/philox.py)."""

    def __init__(self, inst, dtype=np.float64, mode="min", probability=0.7, p_mode="fixed",
                 variant="B", stop_cycle=0, seed=0, var_id=None, frozen=None, var_costs=False, **_ignored):
        """var_id: Philox counter per variable (default: its index); frozen: variables that are
        never evaluated (ghosts of a partition, pydcop_b200/multigpu_dsa.py); var_costs: A-DSA's
        decision (adsa.py:344-377: candidates carry the variable's own cost, the current cost does not)."""
        if "var_edge" not in inst:  # DSA fixtures carry var_con (constraint ids), derive edges
            inst = dict(inst)
            inst["var_edge"] = var_con_to_edges(inst)
        super().__init__(inst, dtype)
        self.unary = np.ascontiguousarray(inst["unary"], dtype=np.float64)
        self.var_cost = np.ascontiguousarray(inst["unary"], dtype=self.dtype) if var_costs else None
        self.mode_max = int(mode == "max")
        self.variant = VARIANTS[variant]
        self.stop_cycle = int(stop_cycle)
        self.seed = int(seed)
        arity = np.diff(self.factor_ptr)
        self.edge_fac = np.repeat(np.arange(self.F, dtype=np.int32), arity).astype(np.int32)
        # neighbours = other variables of incident constraints
        n_count = np.zeros(self.V, np.int64)
        for v in range(self.V):
            for s in range(self.var_ptr[v], self.var_ptr[v + 1]):
                n_count[v] += arity[self.edge_fac[self.var_edge[s]]] - 1
        self.has_nbr = (n_count > 0).astype(np.uint8)
        if frozen is not None:
            self.has_nbr[np.asarray(frozen, dtype=bool)] = 0
        self.var_id = None if var_id is None else np.ascontiguousarray(var_id, dtype=np.int32)
        if p_mode == "arity":  # dsa.py:257-260
            self.prob = np.array([1 / n * 1.2 if n else 0.0 for n in n_count], dtype=np.float64)
        else:
            self.prob = np.full(self.V, float(probability), dtype=np.float64)
        self.con_opt = np.zeros(self.F, self.dtype)
        self.val = np.zeros(self.V, np.int32)
        self.val_next = np.zeros(self.V, np.int32)
        self.val_cost = np.zeros(self.V, self.dtype)
        self.cycle = 0

    def init(self):
        getattr(lib(), "dsa_oracle_constraint_optima" + self.sfx)(
            C.byref(self.fg), _p(self.tables), self.mode_max, _p(self.con_opt))
        lib().dsa_oracle_init(C.byref(self.fg), _p(self.unary), _p(self.has_nbr), self.mode_max,
                              C.c_uint64(self.seed), self._vid(), _p(self.val))
        self.cycle = 0
        return self

    def _vid(self):
        return _p(self.var_id) if self.var_id is not None else C.c_void_p(0)

    @property
    def stopped(self):
        return bool(self.stop_cycle and self.cycle >= self.stop_cycle)

    def compute(self):
        """One evaluate_cycle of every variable: reads `val`, fills `val_next`."""
        if self.stopped:
            return self
        getattr(lib(), "dsa_oracle_step" + self.sfx)(
            C.byref(self.fg), _p(self.tables), _p(self.edge_fac), _p(self.has_nbr),
            _p(self.con_opt), _p(self.prob), self.mode_max, self.variant,
            C.c_uint64(self.seed), C.c_uint32(self.cycle), self._vid(), _p(self.val),
            _p(self.val_next), _p(self.val_cost),
            _p(self.var_cost) if self.var_cost is not None else C.c_void_p(0))
        return self

    def commit(self):
        if self.stopped:
            return self
        self.val, self.val_next = self.val_next, self.val
        self.cycle += 1
        return self

    def step(self, n=1):
        for _ in range(n):
            if self.stopped:
                break
            self.compute().commit()
        return self


def distinct_neighbours(factor_ptr, edge_var, var_ptr, var_edge, edge_fac):
    """CSR of each variable's distinct neighbours, in order of first appearance over its
    constraints (set semantics of mgm.py:245-252; any fixed order is admissible)."""
    V = len(var_ptr) - 1
    ptr, idx = [0], []
    for v in range(V):
        seen = []
        for s in range(var_ptr[v], var_ptr[v + 1]):
            f = edge_fac[var_edge[s]]
            for e in range(factor_ptr[f], factor_ptr[f + 1]):
                u = int(edge_var[e])
                if u != v and u not in seen:
                    seen.append(u)
        idx.extend(seen)
        ptr.append(len(idx))
    return np.array(ptr, dtype=np.int32), np.array(idx, dtype=np.int32)


class MgmOracle(_Graph):
    """Lock-step MGM with injected Philox draws.  State after `init()` == golden state 0, after
    n `step()` == state n (one step = value phase + gain phase)."""

    def __init__(self, inst, dtype=np.float64, mode="min", stop_cycle=0, seed=0, break_mode="lexic",
                 **_ignored):
        if "var_edge" not in inst:
            inst = dict(inst)
            inst["var_edge"] = var_con_to_edges(inst)
        super().__init__(inst, dtype)
        self.unary64 = np.ascontiguousarray(inst["unary"], dtype=np.float64)
        self.unary = np.ascontiguousarray(inst["unary"], dtype=self.dtype)
        iv = inst["init_value"] if "init_value" in inst else None
        self.init_value = (np.ascontiguousarray(iv, dtype=np.int32) if iv is not None
                           else np.full(self.V, -1, np.int32))
        rank = inst["var_rank"] if "var_rank" in inst else None
        self.var_rank = (np.ascontiguousarray(rank, dtype=np.int32) if rank is not None
                         else np.arange(self.V, dtype=np.int32))
        self.mode_max = int(mode == "max")
        self.stop_cycle, self.seed = int(stop_cycle), int(seed)
        arity = np.diff(self.factor_ptr)
        self.edge_fac = np.repeat(np.arange(self.F, dtype=np.int32), arity).astype(np.int32)
        self.nbr_ptr, self.nbr_idx = distinct_neighbours(self.factor_ptr, self.edge_var,
                                                         self.var_ptr, self.var_edge, self.edge_fac)
        if len(self.nbr_idx) == 0:
            self.nbr_idx = np.zeros(1, np.int32)
        self.has_nbr = (np.diff(self.nbr_ptr) > 0).astype(np.uint8)
        self.val = np.zeros(self.V, np.int32)
        self.cost = np.zeros(self.V, self.dtype)
        self.has_cost = np.zeros(self.V, np.uint8)
        self.gain = np.zeros(self.V, self.dtype)
        self.new_val = np.zeros(self.V, np.int32)
        self.cycle = 0

    def init(self):
        lib().mgm_oracle_init(C.byref(self.fg), _p(self.unary64), _p(self.has_nbr),
                              _p(self.init_value), self.mode_max, C.c_uint64(self.seed),
                              _p(self.val))
        self.has_cost[:] = 0
        self.cost[:] = 0
        for v in np.nonzero(self.has_nbr == 0)[0]:  # value_selection(value, cost), mgm.py:292
            self.cost[v] = self.unary[self.unary_off[v] + self.val[v]]
            self.has_cost[v] = 1
        self.cycle = 0
        return self

    @property
    def finished(self):
        """mgm.py:404-407: cycle_count (= rounds done + 1) reached stop_cycle."""
        return bool(self.stop_cycle and self.cycle + 1 >= self.stop_cycle)

    def step(self, n=1):
        gain_fn = getattr(lib(), "mgm_oracle_gain" + self.sfx)
        decide_fn = getattr(lib(), "mgm_oracle_decide" + self.sfx)
        for _ in range(n):
            if self.finished:
                break
            gain_fn(C.byref(self.fg), _p(self.tables), _p(self.unary), _p(self.edge_fac),
                    _p(self.has_nbr), _p(self.nbr_ptr), _p(self.nbr_idx), self.mode_max,
                    C.c_uint64(self.seed), C.c_uint32(self.cycle + 1), _p(self.val), _p(self.cost),
                    _p(self.has_cost), _p(self.gain), _p(self.new_val))
            decide_fn(C.byref(self.fg), _p(self.has_nbr), _p(self.nbr_ptr), _p(self.nbr_idx),
                      _p(self.var_rank), _p(self.gain), _p(self.new_val), _p(self.val),
                      _p(self.cost))
            self.cycle += 1
        return self


def var_con_to_edges(inst):
    """DSA fixtures list constraint ids per variable (var_con); turn them into edge ids."""
    factor_ptr, edge_var = inst["factor_ptr"], inst["edge_var"]
    var_ptr, var_con = inst["var_ptr"], inst["var_con"]
    out = np.zeros(len(var_con), np.int32)
    for v in range(len(var_ptr) - 1):
        for s in range(var_ptr[v], var_ptr[v + 1]):
            f = var_con[s]
            scope = edge_var[factor_ptr[f]:factor_ptr[f + 1]]
            out[s] = factor_ptr[f] + int(np.nonzero(scope == v)[0][0])
    return out


def load_golden(path):
    import json
    z = np.load(path)
    inst = {k: z[k] for k in z.files if k != "meta"}
    meta = json.loads(str(z["meta"]))
    return inst, meta
