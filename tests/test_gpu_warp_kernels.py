"""GPU: the warp-autonomous kernels of round 2 (csrc/maxsum_warp.cuh, the default) against the
round-1 CTA-pipelined kernels (PYDCOP_B200_V2F=pipe / PYDCOP_B200_F2V=pipe) and the CPU oracle:
messages, send decisions, counters, values and reported costs of every cycle, for degree
distributions that reach every code path of the variable side (compile-time degrees 1..8, run-time
degrees 9..16, the irregular class above 16, ragged last tiles, domain sizes whose rows are not
16-byte multiples) and both precisions."""
import numpy as np
import pytest

import oracle as orc
from bench import oracle_instance
from pydcop_b200.generators import random_factor_graph

pytestmark = pytest.mark.gpu


def _skewed_instance(n_vars, d, n_factors, seed, hub=True):
    """Binary factors whose first endpoints follow a heavy-tailed distribution: degrees 1..16 and
    a few variables far above (the irregular class), plus unary factors."""
    rng = np.random.default_rng(seed)
    inst = random_factor_graph(n_vars, d, n_factors, 2, seed=seed, int_tables=False)
    ev = inst["edge_var"].reshape(-1, 2).copy()
    heavy = rng.zipf(1.6, n_factors) % (n_vars // 4)
    pick = rng.random(n_factors) < 0.5
    ev[pick, 0] = heavy[pick]
    if hub:
        ev[: n_factors // 20, 0] = 3
    same = ev[:, 0] == ev[:, 1]
    ev[same, 1] = (ev[same, 0] + 1) % n_vars
    inst["edge_var"] = ev.reshape(-1).astype(np.int32)
    n_unary = n_vars // 3
    uv = rng.integers(0, n_vars, n_unary).astype(np.int32)
    inst["edge_var"] = np.concatenate([inst["edge_var"], uv])
    inst["factor_ptr"] = np.concatenate([inst["factor_ptr"], inst["factor_ptr"][-1] + np.arange(1, n_unary + 1)])
    inst["tables"] = np.concatenate([inst["tables"], rng.uniform(0, 3, n_unary * d).astype(np.float32)])
    return inst


def _check(eng, o, tag):
    q, r = eng.messages()
    val, cost = eng.values()
    fl = eng.flags()
    assert np.array_equal(q, o.q.astype(np.float64)), (tag, "q")
    assert np.array_equal(r, o.r.astype(np.float64)), (tag, "r")
    assert np.array_equal(val, o.value), (tag, "value")
    assert np.array_equal(cost, o.value_cost.astype(np.float64)), (tag, "value_cost")
    assert np.array_equal(fl["q_sent"], o.q_sent) and np.array_equal(fl["r_sent"], o.r_sent), (tag, "sent")
    assert np.array_equal(fl["q_cnt"] >> 1, o.q_flags >> 2) and np.array_equal(fl["r_cnt"] >> 1, o.r_flags >> 2), (tag, "cnt")


@pytest.mark.parametrize("precision", ["f32", "f64"])
@pytest.mark.parametrize("d", [2, 3, 4, 5, 6, 8, 10, 16, 20])
def test_warp_kernels_equal_pipelined_kernels_and_oracle(monkeypatch, d, precision):
    from pydcop_b200 import MaxSumEngine, build_layout
    inst = _skewed_instance(1500 if d <= 10 else 700, d, 4000 if d <= 10 else 1800, seed=40 + d)
    L = build_layout(**inst)
    degs = sorted({c.degree for c in L.var_classes})
    assert any(1 <= k <= 8 for k in degs) and any(9 <= k <= 16 for k in degs) and -1 in degs, degs
    dt = np.float64 if precision == "f64" else np.float32
    for k in ("PYDCOP_B200_V2F", "PYDCOP_B200_F2V"):
        monkeypatch.delenv(k, raising=False)
    warp = MaxSumEngine(L, precision=precision).init()
    monkeypatch.setenv("PYDCOP_B200_V2F", "pipe")
    monkeypatch.setenv("PYDCOP_B200_F2V", "pipe")
    pipe = MaxSumEngine(L, precision=precision).init()
    o = orc.MaxSumOracle(oracle_instance(inst, L), dt).init()
    for k in range(7):
        if k:
            o.step()
            warp.step()
            pipe.step()
        _check(warp, o, ("warp", k))
        _check(pipe, o, ("pipe", k))


@pytest.mark.parametrize("params", [dict(mode="max"), dict(damping_nodes="none"), dict(damping_nodes="factors"),
                                    dict(stability=1e-4), dict(start_messages="all"), dict(start_messages="leafs_vars")])
def test_warp_kernels_parameter_sweep(params):
    from pydcop_b200 import MaxSumEngine, build_layout
    inst = _skewed_instance(2000, 10, 5000, seed=8, hub=False)
    L = build_layout(**inst)
    o = orc.MaxSumOracle(oracle_instance(inst, L), np.float32, **params).init().step(9)
    eng = MaxSumEngine(L, precision="f32", **params).init().step(9)
    _check(eng, o, params)


def test_warp_kernels_many_tiles_per_warp_and_multi_step():
    """More tiles than resident warps (every warp walks its pipeline several times), one step(n) call."""
    from pydcop_b200 import MaxSumEngine, build_layout
    inst = random_factor_graph(120_000, 10, 240_000, 2, seed=77)
    L = build_layout(**inst)
    o = orc.MaxSumOracle(oracle_instance(inst, L), np.float32).init().step(5)
    eng = MaxSumEngine(L, precision="f32").init().step(5)
    _check(eng, o, "large")
