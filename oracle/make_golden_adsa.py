#!/usr/bin/env python
"""Reference trajectories of the UNMODIFIED `ADsaComputation` (pydcop/algorithms/adsa.py:131-392) under a
lock-step schedule: every computation's `delayed_start` runs, its value messages are delivered, then per TICK every
computation's `tick()` runs on the values of the previous tick and all value messages are delivered — one of the
executions the timer-driven algorithm allows (all periods equal, all ticks aligned), and the one a batched engine
runs.  A-DSA's decision differs from DSA's in one place: `find_best_values` (:344-377) adds the VARIABLE's own cost
to every candidate while `current_cost` (:262) does not.

Random draws are injected (Philox keyed by (seed, variable, tick), oracle/philox.py) exactly like make_golden.run_dsa.
TEST INFRASTRUCTURE (build container only); writes tests/golden/adsa_*.npz.

    python oracle/make_golden_adsa.py
"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as G  # noqa: E402
import philox  # noqa: E402
import pydcop.algorithms.adsa as ref_adsa  # noqa: E402
import pydcop.infrastructure.computations as ref_computations  # noqa: E402
from pydcop.algorithms import AlgorithmDef, ComputationDef  # noqa: E402
from pydcop.computations_graph import constraints_hypergraph  # noqa: E402
from pydcop.infrastructure.computations import build_computation  # noqa: E402


class _Tick:
    n = 0


class _AdsaRandom:
    """Replaces `random` in pydcop/algorithms/adsa.py (used :214 start delay, :340-341 change / choice)."""

    @staticmethod
    def random():
        u, _ = philox.draw(G._Ctx.seed, G._Ctx.var, _Tick.n)
        return u

    @staticmethod
    def choice(seq):
        _, w = philox.draw(G._Ctx.seed, G._Ctx.var, _Tick.n)
        return seq[philox.choice_index(w, len(seq))]


class _Timers:
    def set_periodic_action(self, period, cb):
        return cb

    def remove_periodic_action(self, handle):
        pass


def run_adsa(variables, constraints, params, mode, n_ticks, seed):
    variables, constraints = list(variables), list(constraints)
    arr, vidx = G.instance_arrays(variables, constraints)
    V = len(variables)
    g = constraints_hypergraph.build_computation_graph(None, variables=variables, constraints=constraints)
    algo = AlgorithmDef.build_with_default_param("adsa", dict(params), mode=mode)
    comps, outbox, finished = {}, [], set()

    def sender(s, d, m, prio=None, on_error=None):
        outbox.append((s, d, m))

    cidx = {c.name: i for i, c in enumerate(constraints)}
    var_ptr, var_con = [0], []
    for node in g.nodes:
        c = build_computation(ComputationDef(node, algo))
        c.message_sender = sender
        c.periodic_action_handler = _Timers()
        c.finished = (lambda name=c.name: finished.add(name))
        comps[c.name] = c
    for v in variables:
        for c in comps[v.name].constraints:
            var_con.append(cidx[c.name])
        var_ptr.append(len(var_con))
    arr["var_ptr"] = np.array(var_ptr, dtype=np.int32)
    arr["var_con"] = np.array(var_con, dtype=np.int32)
    arr["unary"] = np.array([float(v.cost_for_val(x)) for v in variables for x in v.domain])

    saved = ref_adsa.random, ref_computations.random
    ref_adsa.random, ref_computations.random = _AdsaRandom, G._InitRandom
    G._Ctx.seed = seed
    doms = [list(v.domain) for v in variables]
    value = np.zeros((n_ticks + 1, V), dtype=np.int32)

    def record(k):
        for i, v in enumerate(variables):
            value[k, i] = doms[i].index(comps[v.name].current_value)

    def deliver():
        batch, outbox[:] = list(outbox), []
        for s, d, m in batch:
            comps[d].on_message(s, m, 0)

    try:
        t0 = time.perf_counter()
        _Tick.n = philox.INIT_CYCLE
        for v in variables:
            G._Ctx.var, G._Ctx.comp = vidx[v.name], comps[v.name]
            comps[v.name].start()            # registers the delayed start (start delay draw: unused by the schedule)
            comps[v.name].delayed_start()    # initial value (injected choice), posts it
        deliver()
        record(0)
        for k in range(1, n_ticks + 1):
            _Tick.n = k - 1                  # tick index = the engine's cycle counter when it draws
            for v in variables:
                if v.name in finished:
                    continue
                G._Ctx.var, G._Ctx.comp = vidx[v.name], comps[v.name]
                comps[v.name].tick()
            deliver()
            record(k)
        dt = time.perf_counter() - t0
    finally:
        ref_adsa.random, ref_computations.random = saved
    arr["value"] = value
    arr["finished"] = np.array([v.name in finished for v in variables])
    meta = dict(algo="adsa", mode=mode, params=dict(algo.params), n_cycles=n_ticks, seed=seed, ref_seconds=dt)
    return arr, meta


def connected(vs, cs):
    """Drop the variables without neighbours: `delayed_start` (adsa.py:236-253) unpacks optimal_cost_value's
    (value, cost) pair the wrong way round and selects the COST as the value of such a variable — there is no
    behaviour to pin for them (adsa_gpu gives them the argopt value, the evident intent)."""
    nb = set()
    for c in cs:
        if len(c.dimensions) > 1:
            nb.update(v.name for v in c.dimensions)
    keep = [v for v in vs if v.name in nb]
    names = {v.name for v in keep}
    return keep, [c for c in cs if all(v.name in names for v in c.dimensions)]


def main():
    rng = np.random.default_rng(21)
    cases = []
    vs, cs = G.random_instance(rng, 30, [4], 50, [2], with_var_costs=True)
    cases.append(("adsa_B_varcosts", vs, cs, {}, "min", 25, 3))
    vs, cs = G.random_instance(rng, 30, [3, 5], 45, [2, 2, 3], with_var_costs=True, int_tables=False)
    cases.append(("adsa_A_mixed", vs, cs, {"variant": "A", "probability": 0.5}, "min", 25, 4))
    vs, cs = G.random_instance(rng, 24, [4], 40, [2, 1], with_var_costs=True)
    cases.append(("adsa_C_max", vs, cs, {"variant": "C"}, "max", 25, 5))
    vs, cs = G.random_instance(rng, 40, [6], 50, [2], with_var_costs=False)
    for c in cs:     # few cost levels: ties, delta == 0 branches
        c._m[...] = np.floor(c._m / 4.0)
    cases.append(("adsa_B_ties", vs, cs, {}, "min", 30, 6))
    for name, vs, cs, params, mode, n, seed in cases:
        vs, cs = connected(vs, cs)
        arr, meta = run_adsa(vs, cs, params, mode, n, seed)
        G.save(name, arr, meta)
        print(name, "moved", int((np.diff(arr["value"], axis=0) != 0).sum()), "isolated/finished", int(arr["finished"].sum()))


if __name__ == "__main__":
    main()
