"""GPU: tiled kernels in cycle 1 (the default) against the generic first-cycle kernels
(PYDCOP_B200_FAST_FIRST=0) and the oracle: same messages, send decisions and values at every
cycle, for every start protocol."""
import numpy as np
import pytest

import oracle as orc
from bench import oracle_instance
from pydcop_b200.generators import ising_grid, random_factor_graph

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("kind", ["binary10", "grid", "arity3"])
@pytest.mark.parametrize("start", ["leafs", "leafs_vars", "all"])
@pytest.mark.parametrize("precision", ["f32", "f64"])
def test_fast_first_cycle_equals_default_and_oracle(monkeypatch, kind, start, precision):
    from pydcop_b200 import MaxSumEngine, build_layout
    if kind == "binary10":
        inst = random_factor_graph(4000, 10, 8000, 2, seed=5)
    elif kind == "grid":
        inst = ising_grid(48, 40, seed=6)
    else:
        inst = random_factor_graph(3000, 8, 3000, 3, seed=7)
    L = build_layout(**inst)
    dt = np.float64 if precision == "f64" else np.float32
    monkeypatch.setenv("PYDCOP_B200_FAST_FIRST", "0")
    base = MaxSumEngine(L, precision=precision, start_messages=start).init()
    monkeypatch.setenv("PYDCOP_B200_FAST_FIRST", "1")
    fast = MaxSumEngine(L, precision=precision, start_messages=start).init()
    o = orc.MaxSumOracle(oracle_instance(inst, L), dt, start_messages=start).init()
    for k in range(6):
        if k:
            o.step()
            base.step()
            fast.step()
        for e in (base, fast):
            q, r = e.messages()
            assert np.array_equal(q.astype(dt), o.q) and np.array_equal(r.astype(dt), o.r), (k, e is fast)
            assert np.array_equal(e.values()[0], o.value), (k, e is fast)
            f = e.flags()
            assert np.array_equal(f["q_sent"], o.q_sent) and np.array_equal(f["r_sent"], o.r_sent), (k, e is fast)
    assert fast.launch_count <= base.launch_count     # cycle 1 took the tiled launches (equal when there is one class per side)
