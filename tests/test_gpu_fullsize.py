"""GPU, BASELINE.json's full sizes: the bench workload itself (C2: 100k variables, d=10, 200k
binary factors, 400k edges) checked bit-for-bit against the CPU oracle, plus a size-independent
property (negation symmetry: MaxSum in 'max' mode on negated costs is the exact mirror image of
'min' mode — every IEEE operation on the path commutes with negation), and DSA at 100k variables
d=20 (degree 6, the C4 shape) against the oracle."""
import numpy as np
import pytest

import oracle as orc
from bench import oracle_instance
from pydcop_b200.generators import config_c2, random_factor_graph

pytestmark = pytest.mark.gpu


def test_c2_full_size_bit_exact_vs_oracle():
    from pydcop_b200 import MaxSumEngine, build_layout
    inst = config_c2(seed=0)
    L = build_layout(**inst)
    assert L.n_vars == 100_000 and L.n_edges == 400_000
    cycles = 5
    o = orc.MaxSumOracle(oracle_instance(inst, L), np.float32).init().step(cycles)
    eng = MaxSumEngine(L, precision="f32").init().step(cycles)
    q, r = eng.messages()
    assert np.array_equal(q, o.q.astype(np.float64))
    assert np.array_equal(r, o.r.astype(np.float64))
    val, cost = eng.values()
    assert np.array_equal(val, o.value)
    assert np.array_equal(cost, o.value_cost.astype(np.float64))
    fl = eng.flags()
    assert np.array_equal(fl["q_sent"], o.q_sent) and np.array_equal(fl["r_sent"], o.r_sent)


def test_c2_full_size_negation_symmetry():
    from pydcop_b200 import MaxSumEngine, build_layout
    inst = config_c2(seed=1)
    neg = dict(inst, tables=-np.asarray(inst["tables"]), unary=-np.asarray(inst["unary"]))
    a = MaxSumEngine(build_layout(**inst), precision="f32", mode="min").init().step(8)
    b = MaxSumEngine(build_layout(**neg), precision="f32", mode="max").init().step(8)
    qa, ra = a.messages()
    qb, rb = b.messages()
    assert np.array_equal(qa, -qb) and np.array_equal(ra, -rb)
    assert np.array_equal(a.values()[0], b.values()[0])
    assert np.array_equal(a.values()[1], -b.values()[1])
    fa, fb = a.flags(), b.flags()
    assert np.array_equal(fa["q_sent"], fb["q_sent"]) and np.array_equal(fa["r_sent"], fb["r_sent"])


def test_multi_step_call_equals_single_steps_full_size():
    from pydcop_b200 import MaxSumEngine, build_layout
    L = build_layout(**config_c2(seed=2))
    a = MaxSumEngine(L, precision="f32").init().step(7)
    b = MaxSumEngine(L, precision="f32").init()
    for _ in range(7):
        b.step(1)
    for x, y in zip(a.messages(), b.messages()):
        assert np.array_equal(x, y)
    assert np.array_equal(a.values()[0], b.values()[0])


def test_dsa_100k_vars_d20_exact_vs_oracle():
    from pydcop_b200 import DsaEngine, build_layout
    rng = np.random.default_rng(3)
    inst = random_factor_graph(100_000, 20, 300_000, 2, seed=4, noise=0.0)
    inst["tables"] = rng.integers(0, 10, len(inst["tables"])).astype(np.float32)
    L = build_layout(**inst)
    o = orc.DsaOracle(oracle_instance(inst, L), np.float32, seed=11).init()
    eng = DsaEngine(L, precision="f32", seed=11).init()
    assert eng.tables_or is not None
    assert np.array_equal(eng.values(), o.val)
    for k in range(4):
        o.step()
        eng.step()
        assert np.array_equal(eng.values(), o.val), k


def _assert_state_equals_oracle(eng, o):
    q, r = eng.messages()
    assert np.array_equal(q, o.q.astype(np.float64)), "q"
    assert np.array_equal(r, o.r.astype(np.float64)), "r"
    val, cost = eng.values()
    assert np.array_equal(val, o.value), "value"
    assert np.array_equal(cost, o.value_cost.astype(np.float64)), "value_cost"
    fl = eng.flags()
    assert np.array_equal(fl["q_sent"], o.q_sent) and np.array_equal(fl["r_sent"], o.r_sent), "sent"


def test_c3_ising_1024x1024_full_size_bit_exact_vs_oracle():
    """BASELINE configs[2] at its full size (1 048 576 variables, d=2, 2 097 152 binary + 1 048 576 unary factors,
    E = 5 242 880): three cycles, every message / send decision / value against the CPU oracle."""
    from pydcop_b200 import MaxSumEngine, build_layout
    from pydcop_b200.generators import config_c3
    inst = config_c3(seed=0)
    L = build_layout(**inst)
    assert L.n_vars == 1024 * 1024 and L.n_edges == 5 * 1024 * 1024
    o = orc.MaxSumOracle(oracle_instance(inst, L), np.float32).init().step(3)
    eng = MaxSumEngine(L, precision="f32").init().step(3)
    _assert_state_equals_oracle(eng, o)


def test_c5_arity3_50k_d8_full_size_bit_exact_vs_oracle():
    """BASELINE configs[4] at its full size (50 000 ternary factors over d=8: 512-entry tables)."""
    from pydcop_b200 import MaxSumEngine, build_layout
    from pydcop_b200.generators import config_c5
    inst = config_c5(seed=0)
    L = build_layout(**inst)
    assert L.n_factors == 50_000 and L.n_edges == 150_000
    o = orc.MaxSumOracle(oracle_instance(inst, L), np.float32).init().step(4)
    eng = MaxSumEngine(L, precision="f32").init().step(4)
    _assert_state_equals_oracle(eng, o)


def test_target_1m_vars_d10_full_size_bit_exact_vs_oracle():
    """The north-star instance (1M variables, d=10, 2M binary factors) for two cycles."""
    from pydcop_b200 import MaxSumEngine, build_layout
    from pydcop_b200.generators import config_target
    inst = config_target(seed=0)
    L = build_layout(**inst)
    assert L.n_vars == 1_000_000 and L.n_edges == 4_000_000
    o = orc.MaxSumOracle(oracle_instance(inst, L), np.float32).init().step(2)
    eng = MaxSumEngine(L, precision="f32").init().step(2)
    _assert_state_equals_oracle(eng, o)


def test_c4_dsa_1m_vars_d20_full_size_exact_vs_oracle():
    """BASELINE configs[3] at its full size (1M variables, d=20, 3M binary constraints): two DSA-B cycles."""
    from pydcop_b200 import DsaEngine, build_layout
    from pydcop_b200.generators import config_c4
    inst = config_c4(seed=0)
    L = build_layout(**inst)
    assert L.n_vars == 1_000_000 and L.n_factors == 3_000_000
    o = orc.DsaOracle(oracle_instance(inst, L), np.float32, seed=1).init()
    eng = DsaEngine(L, precision="f32", seed=1).init()
    assert np.array_equal(eng.values(), o.val)
    for k in range(2):
        o.step()
        eng.step()
        assert np.array_equal(eng.values(), o.val), k
    moved = int((eng.values() != DsaEngine(L, precision="f32", seed=1).init().values()).sum())
    assert moved > 100_000
