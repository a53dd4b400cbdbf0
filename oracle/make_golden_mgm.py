#!/usr/bin/env python
"""Generate tests/golden/mgm_*.npz by running the UNMODIFIED reference MGM in lock-step.

TEST INFRASTRUCTURE (build container only: needs /root/reference).  Run:
    python oracle/make_golden_mgm.py [fixture names]

Same method as oracle/make_golden.py: the reference's own `MgmComputation`
(pydcop/algorithms/mgm.py:214) is built through `build_computation(ComputationDef(node, algo))`,
`message_sender` is replaced by a recorder and all messages are delivered batch by batch.  One MGM
cycle is TWO delivery batches: the value messages (every variable then computes its best local
gain and posts it, mgm.py:343-397) and the gain messages (every variable then decides whether it
moves and posts its value for the next cycle, mgm.py:497-537,593-609).

Random draws (mgm.py:301 initial value, :385 choice among equally good values) are injected from
oracle/philox.py, keyed (seed; variable, cycle) — cycle = INIT_CYCLE for the initial value, else
the variable's `cycle_count` when it draws.

Fixture keys: the constraint-hypergraph arrays of the DSA fixtures (dom_size, factor_ptr, edge_var,
table_off, tables, var_ptr, var_con, unary, init_value) + var_rank (position of the variable's name
in sorted order: the lexicographic tie break, mgm.py:574-583) + per-cycle value, cost (NaN while
the reference's current_cost is None), gain, new_value, cycle_count, finished.
"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as G  # noqa: E402  (installs the reference shims)
import philox  # noqa: E402

from pydcop.algorithms import AlgorithmDef, ComputationDef  # noqa: E402
from pydcop.computations_graph import constraints_hypergraph  # noqa: E402
from pydcop.dcop.objects import Domain, Variable, VariableWithCostDict  # noqa: E402
from pydcop.dcop.relations import NAryMatrixRelation  # noqa: E402
from pydcop.dcop.yamldcop import load_dcop_from_file  # noqa: E402
from pydcop.infrastructure.computations import build_computation  # noqa: E402
import pydcop.algorithms.mgm as ref_mgm  # noqa: E402


class _MgmRandom:
    """Replaces `random` in pydcop/algorithms/mgm.py (stdlib; used :301, :385, :419)."""

    @staticmethod
    def _cycle():
        c = G._Ctx.comp
        return philox.INIT_CYCLE if c._state == "starting" else c.cycle_count

    @staticmethod
    def choice(seq):
        seq = list(seq)
        _, w = philox.draw(G._Ctx.seed, G._Ctx.var, _MgmRandom._cycle())
        return seq[philox.choice_index(w, len(seq))]

    @staticmethod
    def random():  # MgmGainMessage.random_nb: only read by the dead `break_mode == random` branch
        u, _ = philox.draw(G._Ctx.seed, G._Ctx.var, _MgmRandom._cycle())
        return u


def run_mgm(variables, constraints, params, mode, n_cycles, seed):
    variables, constraints = list(variables), list(constraints)
    arr, vidx = G.instance_arrays(variables, constraints)
    V = len(variables)
    g = constraints_hypergraph.build_computation_graph(
        None, variables=variables, constraints=constraints)
    algo = AlgorithmDef.build_with_default_param("mgm", dict(params), mode=mode)
    comps, outbox, finished = {}, [], set()

    def sender(s, d, m, prio=None, on_error=None):
        outbox.append((s, d, m))

    for node in g.nodes:
        c = build_computation(ComputationDef(node, algo))
        c.message_sender = sender
        c.finished = (lambda name=c.name: finished.add(name))
        comps[c.name] = c
    cidx = {c.name: i for i, c in enumerate(constraints)}
    var_ptr, var_con = [0], []
    for v in variables:
        for c in comps[v.name].utilities:  # node.constraints order, mgm.py:233
            var_con.append(cidx[c.name])
        var_ptr.append(len(var_con))
    arr["var_ptr"] = np.array(var_ptr, dtype=np.int32)
    arr["var_con"] = np.array(var_con, dtype=np.int32)
    arr["unary"] = np.array([float(v.cost_for_val(x)) for v in variables for x in v.domain])
    doms = [list(v.domain) for v in variables]
    arr["init_value"] = np.array(
        [doms[i].index(v.initial_value) if v.initial_value is not None else -1
         for i, v in enumerate(variables)], dtype=np.int32)
    order = sorted(range(V), key=lambda i: variables[i].name)
    rank = np.zeros(V, dtype=np.int32)
    rank[order] = np.arange(V, dtype=np.int32)
    arr["var_rank"] = rank

    saved = ref_mgm.random
    ref_mgm.random = _MgmRandom
    G._Ctx.seed = seed
    N = n_cycles
    value = np.zeros((N + 1, V), dtype=np.int32)
    cost = np.full((N + 1, V), np.nan)
    gain = np.full((N + 1, V), np.nan)
    new_value = np.full((N + 1, V), -1, dtype=np.int32)
    cycle_count = np.zeros((N + 1, V), dtype=np.int32)

    def record(k):
        for i, v in enumerate(variables):
            c = comps[v.name]
            value[k, i] = doms[i].index(c.current_value)
            if c.current_cost is not None:
                cost[k, i] = c.current_cost
            cycle_count[k, i] = c.cycle_count

    def deliver():
        batch, outbox[:] = list(outbox), []
        for s, d, m in batch:
            G._Ctx.var, G._Ctx.comp = vidx[d], comps[d]
            comps[d].on_message(s, m, 0)
        return len(batch)

    try:
        t0 = time.perf_counter()
        for v in variables:
            G._Ctx.var, G._Ctx.comp = vidx[v.name], comps[v.name]
            comps[v.name].start()
        record(0)
        for k in range(1, N + 1):
            if deliver():  # value messages -> gains computed and posted
                for i, v in enumerate(variables):
                    c = comps[v.name]
                    if c._gain is not None and c.neighbors:
                        gain[k, i] = c._gain
                        new_value[k, i] = doms[i].index(c._new_value)
                deliver()  # gain messages -> decisions, next cycle's values posted
            record(k)
        dt = time.perf_counter() - t0
    finally:
        ref_mgm.random = saved
    arr.update(value=value, cost=cost, gain=gain, new_value=new_value, cycle_count=cycle_count)
    arr["finished"] = np.array([v.name in finished for v in variables])
    meta = dict(algo="mgm", mode=mode, params=dict(algo.params), n_cycles=N, seed=seed,
                ref_seconds=dt, ref_var_updates_per_s=V * N / dt if dt > 0 else 0.0)
    return arr, meta


def dyadic_costs(rng, variables):
    """Variables with own costs that are multiples of 1/8: the reference sums them in the
    iteration order of a Python set (mgm.py:351-366,446-452), exact sums make that order
    irrelevant."""
    out = []
    for v in variables:
        costs = {x: float(rng.integers(0, 17)) / 8.0 for x in v.domain}
        out.append(VariableWithCostDict(v.name, v.domain, costs, initial_value=v.initial_value))
    return out


def rebuild(constraints, variables):
    by_name = {v.name: v for v in variables}
    return [NAryMatrixRelation([by_name[v.name] for v in c.dimensions], c._m, name=c.name)
            for c in constraints]


def main():
    only = set(sys.argv[1:])

    def want(n):
        return not only or n in only

    cases = {
        # name: (params, mode, arities, n_levels of integer cost, var costs, initial values)
        "mgm_default": ({}, "min", [2], 10, False, False),
        "mgm_ties": ({}, "min", [2], 3, False, False),
        "mgm_max": ({}, "max", [2], 4, False, False),
        "mgm_ternary": ({}, "min", [3, 2, 2, 1], 5, False, False),
        "mgm_var_costs": ({}, "min", [2, 2, 3], 6, True, False),
        "mgm_initial_values": ({}, "min", [2], 6, False, True),
        "mgm_stop8": ({"stop_cycle": 8}, "min", [2], 6, False, False),
        "mgm_break_random": ({"break_mode": "random"}, "min", [2], 3, False, False),
    }
    for name, (params, mode, arities, levels, var_costs, init) in cases.items():
        if not want(name):
            continue
        rng = np.random.default_rng(40)
        vs, cs = G.random_instance(rng, 28, [5, 4, 3], 44, arities)
        for c in cs:
            c._m[...] = rng.integers(0, levels, size=c._m.shape)
        if init:
            vs2 = [Variable(v.name, v.domain, initial_value=(list(v.domain)[i % len(v.domain)]
                                                             if i % 3 else None))
                   for i, v in enumerate(vs)]
            cs, vs = rebuild(cs, vs2), vs2
        if var_costs:
            vs2 = dyadic_costs(rng, vs)
            cs, vs = rebuild(cs, vs2), vs2
        vs.append(VariableWithCostDict("iso", Domain("diso", "", [0, 1, 2]), {0: 0.5, 1: 0.25, 2: 0.25}))
        arr, meta = run_mgm(vs, cs, params, mode, 20, seed=4321)
        G.save(name, arr, meta)
    if want("mgm_gc10"):
        dcop = load_dcop_from_file([os.path.join(G.INSTANCES, "graph_coloring_10_4_15_0.1.yml")])
        arr, meta = run_mgm(dcop.variables.values(), dcop.constraints.values(),
                            {}, dcop.objective, 20, seed=78)
        G.save("mgm_gc10", arr, meta)
    if want("mgm_rand_d12"):
        rng = np.random.default_rng(41)
        vs, cs = G.random_instance(rng, 40, [12], 110, [2])
        arr, meta = run_mgm(vs, cs, {}, "min", 15, seed=6)
        G.save("mgm_rand_d12", arr, meta)


if __name__ == "__main__":
    main()
