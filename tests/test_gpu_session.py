"""GPU: the plugin's session core (graph assembly from node objects -> layout -> REAL engine ->
published values) with duck-typed stand-ins for pyDcop's Variable / Constraint objects (the
reference itself is not on the GPU box).  Values must equal the oracle's on the same instance."""
import time

import numpy as np
import pytest

import oracle as orc

pytestmark = pytest.mark.gpu


class _Var:
    def __init__(self, name, domain, costs=None, initial_value=None):
        self.name, self.domain, self._costs, self.initial_value = name, list(domain), costs, initial_value

    def cost_for_val(self, v):
        return self._costs[v] if self._costs else 0.0


class _Con:
    def __init__(self, name, dims, table):
        self.name, self.dimensions, self._t = name, dims, np.asarray(table, dtype=np.float64)

    def __call__(self, **asg):
        idx = tuple(v.domain.index(asg[v.name]) for v in self.dimensions)
        return float(self._t[idx])


def _graph(seed=0):
    rng = np.random.default_rng(seed)
    doms = {3: ["R", "G", "B"], 4: [0, 1, 2, 3]}
    vs = [_Var(f"v{i}", doms[3 if i % 3 else 4],
               {x: float(rng.uniform(0, 0.3)) for x in doms[3 if i % 3 else 4]}) for i in range(40)]
    cs = []
    for j in range(70):
        a, b = rng.choice(40, 2, replace=False)
        cs.append(_Con(f"c{j}", [vs[a], vs[b]], rng.integers(0, 6, (len(vs[a].domain), len(vs[b].domain)))))
    for j in range(8):
        a, b, c = rng.choice(40, 3, replace=False)
        cs.append(_Con(f"t{j}", [vs[a], vs[b], vs[c]],
                       rng.integers(0, 6, (len(vs[a].domain), len(vs[b].domain), len(vs[c].domain)))))
    links = {v.name: [c.name for c in cs if v in c.dimensions] for v in vs}
    return vs, cs, links


def _wait(session, cycle, timeout=60):
    t0 = time.time()
    while time.time() - t0 < timeout:
        snap = session.poll()
        if snap is not None and snap.cycle >= cycle:
            return snap
        time.sleep(0.01)
    raise AssertionError(f"session did not reach cycle {cycle}: {session.error!r}")


def test_maxsum_session_runs_real_engine_and_matches_oracle():
    from pydcop_b200.algorithms._session import GpuSession
    GpuSession.reset()
    GpuSession.engine_factory = None
    vs, cs, links = _graph(1)
    params = dict(damping=0.5, damping_nodes="both", stability=0.1, noise=0.0, start_messages="leafs",
                  stop_cycle=20, precision="f64", seed=0)
    s = GpuSession.get("maxsum:t1", "maxsum")
    for c in cs:
        s.add_factor(c.name, c, params, "min")
    for v in vs:
        s.add_variable(v.name, v, links[v.name], None, params, "min")
    assert s.is_complete()
    for n in [c.name for c in cs] + [v.name for v in vs]:
        s.notify_started(n)
    snap = _wait(s, 20)
    assert snap.finished
    inst = s.build_instance()
    o = orc.MaxSumOracle(inst, np.float64, noise=0.0).init().step(20)
    for i, n in enumerate(s.var_order):
        dom = list(s.variables[n].domain)
        assert snap.values[n][0] == dom[int(o.value[i])], n
        assert snap.values[n][1] == float(o.value_cost[i]), n
    GpuSession.reset()


def test_dsa_session_runs_real_engine_and_matches_oracle():
    from pydcop_b200.algorithms._session import GpuSession
    GpuSession.reset()
    GpuSession.engine_factory = None
    vs, cs, links = _graph(2)
    vs.append(_Var("lonely", [5, 7, 6], {5: 0.3, 7: 0.1, 6: 0.1}))   # isolated: argmin (cost, value) -> 6
    links["lonely"] = []
    params = dict(probability=0.7, p_mode="fixed", variant="B", stop_cycle=15, precision="f64", seed=99)
    s = GpuSession.get("dsa:t2", "dsa")
    for v in vs:
        s.add_variable(v.name, v, links[v.name], [c for c in cs if v in c.dimensions], params, "min")
    assert s.is_complete()
    for v in vs:
        s.notify_started(v.name)
    snap = _wait(s, 15)
    assert snap.finished and snap.values["lonely"][0] == 6
    inst = s.build_instance()
    inst["unary"] = inst["unary"]
    o = orc.DsaOracle(inst, np.float64, seed=99, stop_cycle=15).init()
    o.val[s.var_order.index("lonely")] = 2          # the session's tuple-ordered pick (value 6)
    o.step(15)
    for i, n in enumerate(s.var_order):
        assert snap.values[n][0] == list(s.variables[n].domain)[int(o.val[i])], n
    GpuSession.reset()
