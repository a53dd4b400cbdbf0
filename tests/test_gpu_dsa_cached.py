"""GPU: the DSA kernel with the active-row array (csrc/dsa_cached.cuh) — rows are re-read from the
oriented tables only where a neighbour's value changed.  Against the CPU oracle over enough cycles
for the assignment to go from 'everything changes' to 'almost nothing changes', against the
uncached kernel, across re-init, with hub variables (more slots than one chunk) and isolated ones."""
import numpy as np
import pytest

import oracle as orc
from bench import oracle_instance
from pydcop_b200.generators import random_factor_graph

pytestmark = pytest.mark.gpu


def _inst(n_vars, d, n_factors, seed, levels=4):
    inst = random_factor_graph(n_vars, d, n_factors, 2, seed=seed, noise=0.0)
    rng = np.random.default_rng(seed + 1)
    inst["tables"] = rng.integers(0, levels, len(inst["tables"])).astype(np.float32)
    return inst


@pytest.mark.parametrize("precision", ["f32", "f64"])
@pytest.mark.parametrize("variant,mode,d", [("B", "min", 20), ("A", "min", 10), ("C", "max", 6), ("B", "max", 16), ("C", "min", 3)])
def test_cached_rows_vs_oracle_over_many_cycles(variant, mode, d, precision):
    from pydcop_b200 import DsaEngine, build_layout
    inst = _inst(6000, d, 17000, seed=3 + d)
    L = build_layout(**inst)
    npdt = np.float64 if precision == "f64" else np.float32
    o = orc.DsaOracle(oracle_instance(inst, L), npdt, mode=mode, variant=variant, seed=5).init()
    eng = DsaEngine(L, precision=precision, mode=mode, variant=variant, seed=5).init()
    assert eng.row_cache is not None
    changed = []
    for k in range(40):
        before = o.val.copy()
        o.step()
        eng.step()
        assert np.array_equal(eng.values(), o.val), k
        changed.append(float((before != o.val).mean()))
    assert changed[0] > 0.2 and changed[-1] < changed[0]      # both regimes were exercised
    # slot_last mirrors the neighbour values the LAST executed cycle read (= the assignment before it)
    last = eng.slot_last.cpu().numpy()[:L.n_edges]
    assert (last != 255).all()


def test_cached_equals_uncached_and_reinit(monkeypatch):
    from pydcop_b200 import DsaEngine, build_layout
    inst = _inst(30000, 20, 90000, seed=9, levels=10)
    L = build_layout(**inst)
    a = DsaEngine(L, precision="f32", variant="B", seed=2).init().step(7)
    va = a.values().copy()
    a.init().step(3)            # re-init: the array is invalidated, not reused
    a.step(4)
    assert np.array_equal(a.values(), va)
    monkeypatch.setenv("PYDCOP_B200_DSA_CACHE", "0")
    b = DsaEngine(build_layout(**inst), precision="f32", variant="B", seed=2).init().step(7)
    assert np.array_equal(b.values(), va)
    assert np.array_equal(a.value_cost.cpu().numpy(), b.value_cost.cpu().numpy())


def test_hub_and_isolated_variables():
    """a variable with more slots than one chunk (256) and variables without any constraint"""
    from pydcop_b200 import DsaEngine, build_layout
    inst = _inst(3000, 10, 5000, seed=21)
    ev = inst["edge_var"].reshape(-1, 2).copy()
    ev[:700, 0] = 5                    # hub of degree ~700 (3 chunks)
    ev[ev[:, 0] == ev[:, 1], 1] = 6
    used = np.zeros(3000, bool)
    used[ev.reshape(-1)] = True
    inst["edge_var"] = ev.reshape(-1)
    L = build_layout(**inst)
    o = orc.DsaOracle(oracle_instance(inst, L), np.float32, variant="C", seed=4).init()
    eng = DsaEngine(L, precision="f32", variant="C", seed=4).init()
    assert (~used).any()
    for k in range(15):
        o.step()
        eng.step()
        assert np.array_equal(eng.values(), o.val), k
