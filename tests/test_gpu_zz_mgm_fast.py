"""GPU: the opt-in fast MGM value-phase kernel (PYDCOP_B200_MGM_FAST=2|4, csrc/mgm_fast_kernels.cuh)
against the default kernel and the oracle: values, costs, gains and intended moves identical."""
import numpy as np
import pytest

import oracle as orc
from pydcop_b200.generators import random_factor_graph
from pydcop_b200.layout import build_layout, default_var_csr

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("d,mode,precision,chunk", [(20, "min", "f32", 4), (20, "min", "f32", 2), (20, "max", "f64", 4),
                                                    (10, "min", "f32", 4), (16, "min", "f64", 2), (8, "max", "f32", 4),
                                                    (4, "min", "f64", 2)])
def test_fast_value_phase_equals_default_and_oracle(monkeypatch, d, mode, precision, chunk):
    from pydcop_b200.engine import MgmEngine
    n = 20000
    inst = random_factor_graph(n, d, n * 3, 2, seed=d + chunk, noise=0.5, int_tables=False)
    rng = np.random.default_rng(3)
    rank = rng.permutation(n).astype(np.int32)
    L = build_layout(**inst)
    dt = np.float64 if precision == "f64" else np.float32
    monkeypatch.setenv("PYDCOP_B200_MGM_FAST", "0")
    base = MgmEngine(L, precision=precision, mode=mode, seed=5, var_rank=rank).init()
    monkeypatch.setenv("PYDCOP_B200_MGM_FAST", str(chunk))
    fast = MgmEngine(L, precision=precision, mode=mode, seed=5, var_rank=rank).init()
    assert fast.fast_chunk == chunk and base.fast_chunk == 0
    vp, ve = default_var_csr(n, inst["edge_var"])
    o = orc.MgmOracle(dict(inst, var_ptr=vp, var_edge=ve, var_rank=rank), dt, mode=mode, seed=5).init()
    for k in range(1, 9):
        o.step()
        base.step()
        fast.step()
        for e in (base, fast):
            val, cost = e.values()
            assert np.array_equal(val, o.val), k
            assert np.array_equal(cost.astype(dt), o.cost), k
            gain, new_value = e.gains()
            assert np.array_equal(gain.astype(dt), o.gain) and np.array_equal(new_value, o.new_val), k


def test_shapes_outside_the_fast_set_use_the_default_kernel(monkeypatch):
    from pydcop_b200.engine import MgmEngine
    monkeypatch.setenv("PYDCOP_B200_MGM_FAST", "4")
    inst = random_factor_graph(400, 3, 700, 2, seed=2, noise=0.25)       # d=3: not compiled
    e = MgmEngine(build_layout(**inst), precision="f64", seed=1).init().step(4)
    assert e.fast_chunk == 0 and e.cycle == 4
    t = random_factor_graph(400, 4, 300, 3, seed=3)                       # arity 3: no fast arrays
    e = MgmEngine(build_layout(**t), precision="f32", seed=1).init().step(2)
    assert e.fast_chunk == 0
