"""GPU: MGM kernels (pydcop_b200/csrc/mgm.cu) through the C-ABI against the reference's own
lock-step trajectories (tests/golden/mgm_*.npz) and against the oracle on larger random graphs.
Named to run after the MaxSum / DSA GPU files."""
import os

import numpy as np
import pytest

import oracle as orc
from conftest import GOLDEN_DIR, golden_names
from pydcop_b200.generators import random_factor_graph
from pydcop_b200.layout import layout_from_instance

pytestmark = pytest.mark.gpu


def _engine(inst, precision, meta):
    from pydcop_b200.engine import MgmEngine
    L = layout_from_instance(inst)
    rank = inst["var_rank"] if "var_rank" in inst else None
    return MgmEngine(L, precision=precision, mode=meta["mode"], seed=meta["seed"], var_rank=rank,
                     **meta["params"])


@pytest.mark.parametrize("name", golden_names("mgm_"))
@pytest.mark.parametrize("precision", ["f64", "f32"])
def test_mgm_matches_reference_trajectory(name, precision):
    inst, meta = orc.load_golden(os.path.join(GOLDEN_DIR, name + ".npz"))
    dt = np.float64 if precision == "f64" else np.float32
    eng = _engine(inst, precision, meta).init()
    has_nbr = ~np.isnan(inst["gain"][1])
    for k in range(meta["n_cycles"] + 1):
        ran = False
        if k:
            before = eng.cycle
            eng.step()
            ran = eng.cycle > before
        val, cost = eng.values()
        assert np.array_equal(val, inst["value"][k]), (name, k)
        known = ~np.isnan(inst["cost"][k])
        assert np.array_equal(~np.isnan(cost), known), (name, k)
        assert np.array_equal(cost[known].astype(dt), inst["cost"][k][known].astype(dt)), (name, k)
        if ran:
            gain, new_value = eng.gains()
            assert np.array_equal(gain[has_nbr].astype(dt), inst["gain"][k][has_nbr].astype(dt)), (name, k)
            assert np.array_equal(new_value[has_nbr], inst["new_value"][k][has_nbr]), (name, k)
    assert eng.cycle + 1 == int(inst["cycle_count"][-1].max())
    assert eng.launch_count == 1 + 2 * eng.cycle


@pytest.mark.parametrize("precision,d,arity,mode", [("f64", 10, 2, "min"), ("f32", 10, 2, "min"),
                                                    ("f64", 5, 3, "max"), ("f32", 20, 2, "min")])
def test_mgm_matches_oracle_on_random_graphs(precision, d, arity, mode):
    """20k variables, float tables and float variable costs: same operand order as the oracle, so
    values, costs and gains are bit-identical in both precisions."""
    from pydcop_b200.engine import MgmEngine
    n = 20000
    inst = random_factor_graph(n, d, n * 2 if arity == 2 else n, arity, seed=3, noise=0.5, int_tables=False)
    rng = np.random.default_rng(5)
    inst["var_rank"] = rng.permutation(n).astype(np.int32)
    inst["init_value"] = np.where(rng.random(n) < 0.3, rng.integers(0, d, n), -1).astype(np.int32)
    dt = np.float64 if precision == "f64" else np.float32
    o = orc.MgmOracle(_with_csr(inst), dt, mode=mode, seed=17).init()
    eng = MgmEngine(layout_from_instance(inst), precision=precision, mode=mode, seed=17,
                    var_rank=inst["var_rank"]).init()
    assert np.array_equal(eng.values()[0], o.val)
    for k in range(1, 13):
        o.step()
        eng.step()
        val, cost = eng.values()
        assert np.array_equal(val, o.val), k
        assert np.array_equal(cost.astype(dt), o.cost), k
        gain, new_value = eng.gains()
        assert np.array_equal(gain.astype(dt), o.gain) and np.array_equal(new_value, o.new_val), k


def _with_csr(inst):
    from pydcop_b200.layout import default_var_csr
    out = dict(inst)
    if out.get("var_ptr") is None:
        out["var_ptr"], out["var_edge"] = default_var_csr(len(inst["dom_size"]), inst["edge_var"])
    return out


def test_mgm_stop_cycle_and_isolated_variables():
    from pydcop_b200.engine import MgmEngine
    inst = random_factor_graph(500, 4, 600, 2, seed=1, noise=0.25, int_tables=True)
    L = layout_from_instance(inst)
    eng = MgmEngine(L, precision="f64", stop_cycle=6, seed=2).init()
    eng.step(50)
    assert eng.cycle == 5 and eng.finished       # rounds run = stop_cycle - 1 (mgm.py:404)
    o = orc.MgmOracle(_with_csr(inst), np.float64, stop_cycle=6, seed=2).init().step(50)
    val, cost = eng.values()
    assert np.array_equal(val, o.val) and o.cycle == 5
    iso = np.nonzero(o.has_nbr == 0)[0]
    assert len(iso) > 0
    u = inst["unary"].reshape(500, 4)
    assert np.array_equal(val[iso], u[iso].argmin(axis=1))
    assert np.array_equal(cost[iso], u[iso].min(axis=1))
