"""Pins the CPU oracle (oracle/dcop_oracle.c) to the reference's own lock-step trajectories.

tests/golden/*.npz were produced by oracle/make_golden.py from the unmodified reference
(pydcop/algorithms/maxsum.py, dsa.py; mgm.py by oracle/make_golden_mgm.py) — see those scripts.  The f64 oracle keeps the reference's
floating-point operand order, so messages, send flags and values must agree BIT FOR BIT at every
cycle; only the reported selection cost (whose summation order in the reference is message
*arrival* order, maxsum.py:608-610) is compared to 1e-12.
"""
import os

import numpy as np
import pytest

import oracle as orc
from conftest import GOLDEN_DIR, golden_names


def _assert_values_match(o, inst, k, what):
    """Values must be identical, except on ulp-level ties: the reference sums the received costs
    in message ARRIVAL order (maxsum.py:608-610; neighbours come from a `set`,
    computations_graph/objects.py:93-94, so that order is not even stable across runs) while the
    oracle sums in `links` order.  A differing pick is accepted only if both candidates' totals
    (from the reference's own pre-cycle messages) agree to 1e-12 relative, i.e. the reference's
    choice is itself rounding-order dependent."""
    ref_value = inst["value"][k]
    diff = np.nonzero(o.value != ref_value)[0]
    assert len(diff) == 0 or k > 0, what
    for v in diff:
        d = int(o.dom_size[v])
        tot = np.array(inst["unary"][o.unary_off[v]:o.unary_off[v] + d], dtype=np.float64)
        for s in range(o.var_ptr[v], o.var_ptr[v + 1]):
            e = o.var_edge[s]
            if inst["r_valid"][k - 1][e]:
                tot += inst["r_state"][k - 1][o.msg_off[e]:o.msg_off[e] + d]
        a, b = tot[o.value[v]], tot[ref_value[v]]
        assert abs(a - b) <= 1e-12 * max(1.0, abs(a), abs(b)), (what, int(v), a, b)
    return len(diff)


@pytest.mark.parametrize("name", golden_names("ms_"))
def test_maxsum_oracle_f64_bit_exact_vs_reference(name):
    inst, meta = orc.load_golden(os.path.join(GOLDEN_DIR, name + ".npz"))
    o = orc.MaxSumOracle(inst, np.float64, mode=meta["mode"], **meta["params"]).init()
    n = meta["n_cycles"]
    has_edges = np.diff(inst["var_ptr"]) > 0
    for k in range(n + 1):
        if k:
            o.step()
        assert np.array_equal(o.r_flags & 1, inst["r_valid"][k]), (name, k, "r_valid")
        assert np.array_equal(o.q_flags & 1, inst["q_valid"][k]), (name, k, "q_valid")
        assert np.array_equal(o.r_sent.astype(bool), inst["r_sent"][k]), (name, k, "r_sent")
        assert np.array_equal(o.q_sent.astype(bool), inst["q_sent"][k]), (name, k, "q_sent")
        assert np.array_equal(o.r, inst["r_state"][k]), (name, k, "r")
        assert np.array_equal(o.q, inst["q_state"][k]), (name, k, "q")
        _assert_values_match(o, inst, k, (name, k, "value"))
        ref_cost = inst["value_cost"][k]
        ok = has_edges if k else np.zeros_like(has_edges)  # cycle-0 cost of initial_value vars
        np.testing.assert_allclose(o.value_cost[ok], ref_cost[ok], rtol=1e-12, atol=1e-12)


@pytest.mark.parametrize("name", golden_names("ms_"))
def test_maxsum_oracle_f32_tracks_f64(name):
    """The f32 build (the engine's throughput type) stays within 1e-5 relative of the f64
    trajectory for as long as no send-gate decision flips (north_star tolerance)."""
    inst, meta = orc.load_golden(os.path.join(GOLDEN_DIR, name + ".npz"))
    o = orc.MaxSumOracle(inst, np.float32, mode=meta["mode"], **meta["params"]).init()
    scale = max(1.0, float(np.abs(inst["tables"]).max(initial=0)))
    for k in range(meta["n_cycles"] + 1):
        if k:
            o.step()
        if not (np.array_equal(o.r_sent.astype(bool), inst["r_sent"][k])
                and np.array_equal(o.q_sent.astype(bool), inst["q_sent"][k])):
            assert k >= 5, (name, k, "gate flipped too early for rounding to explain")
            break
        np.testing.assert_allclose(o.r, inst["r_state"][k], rtol=1e-5, atol=2e-5 * scale)
        np.testing.assert_allclose(o.q, inst["q_state"][k], rtol=1e-5, atol=2e-5 * scale)


@pytest.mark.parametrize("name", golden_names("dsa_"))
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_dsa_oracle_exact_vs_reference(name, dtype):
    inst, meta = orc.load_golden(os.path.join(GOLDEN_DIR, name + ".npz"))
    o = orc.DsaOracle(inst, dtype, mode=meta["mode"], seed=meta["seed"], **meta["params"]).init()
    for k in range(meta["n_cycles"] + 1):
        if k:
            o.step()
        assert np.array_equal(o.val, inst["value"][k]), (name, k)
    assert o.cycle == int(inst["cycle_count"][-1].max())


@pytest.mark.parametrize("name", golden_names("msx_"))
def test_maxsum_oracle_with_infinite_costs(name):
    """Hard constraints (+/-inf table entries, oracle/make_golden_extra.py): the reference's messages
    become inf and NaN (inf - inf in the normalisation and in approx_match); the oracle must produce
    the same inf / NaN pattern, the same send decisions and the same values at every cycle."""
    inst, meta = orc.load_golden(os.path.join(GOLDEN_DIR, name + ".npz"))
    params = {k: v for k, v in meta["params"].items() if k != "noise"}
    o = orc.MaxSumOracle(inst, np.float64, mode=meta["mode"], **params).init()
    assert np.isinf(inst["tables"]).any()
    for k in range(meta["n_cycles"] + 1):
        if k:
            o.step()
        assert np.array_equal(o.q, inst["q_state"][k], equal_nan=True), (name, k)
        assert np.array_equal(o.r, inst["r_state"][k], equal_nan=True), (name, k)
        assert np.array_equal(o.q_sent, inst["q_sent"][k]) and np.array_equal(o.r_sent, inst["r_sent"][k]), (name, k)
        assert np.array_equal(o.value, inst["value"][k]), (name, k)
    assert np.isnan(inst["q_state"][-1]).any() and np.isinf(inst["r_state"][-1]).any()


@pytest.mark.parametrize("name", golden_names("mgm_"))
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_mgm_oracle_exact_vs_reference(name, dtype):
    """Values, current costs, gains and intended moves of every cycle (fixtures written by
    oracle/make_golden_mgm.py from the unmodified pydcop/algorithms/mgm.py).  All fixture costs
    are integers or multiples of 1/8, so float32 is exact as well."""
    inst, meta = orc.load_golden(os.path.join(GOLDEN_DIR, name + ".npz"))
    o = orc.MgmOracle(inst, dtype, mode=meta["mode"], seed=meta["seed"], **meta["params"]).init()
    for k in range(meta["n_cycles"] + 1):
        ran = False
        if k:
            before = o.cycle
            o.step()
            ran = o.cycle > before
        assert np.array_equal(o.val, inst["value"][k]), (name, k)
        known = ~np.isnan(inst["cost"][k])
        assert np.array_equal(known, o.has_cost.astype(bool)), (name, k)
        assert np.array_equal(o.cost[known], inst["cost"][k][known].astype(dtype)), (name, k)
        if ran:
            m = ~np.isnan(inst["gain"][k])
            assert np.array_equal(m, o.has_nbr.astype(bool))
            assert np.array_equal(o.gain[m], inst["gain"][k][m].astype(dtype)), (name, k)
            assert np.array_equal(o.new_val[m], inst["new_value"][k][m]), (name, k)
    # cycle_count of the reference = rounds done + 1 for connected variables (mgm.py:404)
    assert o.cycle + 1 == int(inst["cycle_count"][-1].max())
    assert not inst["finished"][o.has_nbr.astype(bool)].any() or o.finished


def test_philox_matches_python_definition():
    import ctypes as C
    import philox
    out = (C.c_uint32 * 4)()
    for ctr, key in [((0, 0, 0, 0), (0, 0)), ((7, 3, 0, 0), (1234, 0)),
                     ((0xFFFFFFFF,) * 4, (0xFFFFFFFF,) * 2)]:
        orc.lib().oracle_philox(*[C.c_uint32(x) for x in ctr], *[C.c_uint32(x) for x in key], out)
        assert tuple(out) == philox.philox4x32_10(ctr, key)
    # Random123 known-answer vectors
    assert philox.philox4x32_10((0, 0, 0, 0), (0, 0)) == (
        0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8)
    assert philox.philox4x32_10((0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344),
                                (0xA4093822, 0x299F31D0)) == (
        0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1)


def test_reference_unit_test_vectors():
    """Known answers held by the reference's own unit tests for this path:
    tests/unit/test_algorithms_amaxsum.py:77-150 (factor_costs_for_var, unary and |x1-x2|/2),
    tests/unit/test_algorithms_maxsum.py:103-127 (empty costs, select_value 0.1 -> value 3),
    tests/unit/test_algorithms_amaxsum.py:160-200 (approx_match)."""
    # binary |x1 - x2| / 2, x1 in 0..9, x2 in 0..4, no incoming costs, start_messages=all
    t = np.abs(np.arange(10)[:, None] - np.arange(5)[None, :]) / 2.0
    inst = dict(dom_size=[10, 5], factor_ptr=[0, 2], edge_var=[0, 1], table_off=[0, 50],
                tables=t.reshape(-1), var_ptr=[0, 1, 2], var_edge=[0, 1],
                unary=np.zeros(15))
    o = orc.MaxSumOracle(inst, start_messages="all").init()
    costs = o.r[:10]
    assert costs[5] == (5 - 4) / 2 and costs[9] == (9 - 4) / 2 and costs[2] == 0
    # unary factor x*2
    inst = dict(dom_size=[10], factor_ptr=[0, 1], edge_var=[0], table_off=[0, 10],
                tables=np.arange(10) * 2.0, var_ptr=[0, 1], var_edge=[0], unary=np.zeros(10))
    o = orc.MaxSumOracle(inst).init()
    assert o.r[0] == 0 and o.r[5] == 10
    # select_value with cost (4 - v) / 10 over [1, 2, 3] -> value 3 (index 2), cost 0.1
    inst = dict(dom_size=[3], factor_ptr=[0], edge_var=[], table_off=[0], tables=[],
                var_ptr=[0, 0], var_edge=[], unary=[(4 - v) / 10 for v in (1, 2, 3)])
    o = orc.MaxSumOracle(inst).init()
    assert o.value[0] == 2 and o.value_cost[0] == 0.1


@pytest.mark.parametrize("name", golden_names("adsa_"))
def test_adsa_oracle_matches_reference_trajectory(name):
    """The DSA oracle with var_costs=True reproduces, tick by tick, the values of the UNMODIFIED
    ADsaComputation (pydcop/algorithms/adsa.py) recorded by oracle/make_golden_adsa.py."""
    inst, meta = orc.load_golden(os.path.join(GOLDEN_DIR, name + ".npz"))
    p = meta["params"]
    o = orc.DsaOracle(inst, np.float64, mode=meta["mode"], probability=p["probability"], variant=p["variant"],
                      seed=meta["seed"], var_costs=True).init()
    assert np.array_equal(o.val, inst["value"][0])
    for k in range(1, meta["n_cycles"] + 1):
        o.step()
        assert np.array_equal(o.val, inst["value"][k]), k
