"""GPU: A-DSA (`adsa_gpu`, DsaEngine(var_costs=True)) — the DSA kernels with the variables' own costs added to
the candidates (pydcop/algorithms/adsa.py:344-377) — against the trajectories recorded from the UNMODIFIED
ADsaComputation under aligned ticks (tests/golden/adsa_*.npz, oracle/make_golden_adsa.py) and against the oracle
on a graph large enough for the oriented-table fast kernel."""
import glob
import os

import numpy as np
import pytest

import oracle as orc
from bench import oracle_instance
from pydcop_b200.generators import random_factor_graph

pytestmark = pytest.mark.gpu

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
NAMES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, "adsa_*.npz")))


@pytest.mark.parametrize("precision", ["f64", "f32"])
@pytest.mark.parametrize("name", NAMES)
def test_adsa_engine_matches_reference_trajectory(name, precision):
    from pydcop_b200 import DsaEngine, layout_from_instance
    inst, meta = orc.load_golden(os.path.join(GOLDEN_DIR, name + ".npz"))
    p = meta["params"]
    eng = DsaEngine(layout_from_instance(inst), precision=precision, mode=meta["mode"], probability=p["probability"],
                    variant=p["variant"], seed=meta["seed"], var_costs=True).init()
    assert np.array_equal(eng.values(), inst["value"][0])
    for k in range(1, meta["n_cycles"] + 1):
        eng.step()
        assert np.array_equal(eng.values(), inst["value"][k]), k


@pytest.mark.parametrize("precision", ["f32", "f64"])
@pytest.mark.parametrize("variant,mode,d", [("B", "min", 20), ("A", "max", 10), ("C", "min", 8)])
def test_adsa_fast_kernel_vs_oracle(variant, mode, d, precision):
    from pydcop_b200 import DsaEngine, build_layout
    rng = np.random.default_rng(5)
    inst = random_factor_graph(3000, d, 8000, 2, seed=20 + d, noise=0.0)
    inst["tables"] = rng.integers(0, 3, len(inst["tables"])).astype(np.float32)
    inst["unary"] = np.round(rng.uniform(0, 2, len(inst["unary"])), 2)     # decisive variable costs
    L = build_layout(**inst)
    dt = np.float64 if precision == "f64" else np.float32
    o = orc.DsaOracle(oracle_instance(inst, L), dt, mode=mode, variant=variant, seed=9, var_costs=True).init()
    plain = orc.DsaOracle(oracle_instance(inst, L), dt, mode=mode, variant=variant, seed=9).init()
    eng = DsaEngine(L, precision=precision, mode=mode, variant=variant, seed=9, var_costs=True).init()
    assert eng.tables_or is not None
    for k in range(10):
        o.step()
        plain.step()
        eng.step()
        assert np.array_equal(eng.values(), o.val), k
    assert not np.array_equal(o.val, plain.val)      # the variable costs really decide
