"""GPU: the direct solve entry (ingestion -> layout -> REAL engine -> result dict) against the
oracle driven on the same arrays."""
import os

import numpy as np
import pytest

import oracle as orc
from pydcop_b200 import ingest, solve as S
from pydcop_b200.generators import random_factor_graph

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
SPLIT = [os.path.join(HERE, "golden", "yaml", f) for f in ("split_problem.yaml", "split_agents.yaml")]
MIXED = os.path.join(HERE, "golden", "yaml", "mixed_grammar.yaml")


def _index(dcop, assignment):
    return [dcop.values_of(i).index(assignment[n]) for i, n in enumerate(dcop.var_names)]


@pytest.mark.parametrize("files", [SPLIT, MIXED], ids=["split", "mixed"])
def test_maxsum_f64_from_yaml_matches_oracle(files):
    d = ingest.load_yaml(files)
    res = S.solve(files, "maxsum", {"stop_cycle": 30, "noise": 0}, precision="f64")
    o = orc.MaxSumOracle(d.instance(), np.float64, mode=d.objective).init().step(30)
    assert _index(d, res["assignment"]) == o.value.tolist()
    assert res["status"] == "FINISHED" and res["cycle"] == 30
    assert res["cost"] == pytest.approx(d.cost(o.value), rel=1e-12)


def test_maxsum_f32_from_binary_container_with_seeded_noise(tmp_path):
    inst = random_factor_graph(3000, 10, 6000, 2, seed=7, noise=0.0)
    p = tmp_path / "g.fgb"
    ingest.save_instance(p, ingest.from_arrays(inst), names=False)
    res = S.solve(str(p), "maxsum", {"stop_cycle": 20}, precision="f32", seed=11)
    d = ingest.load_instance(p)
    ref = d.instance()
    ref["unary"] = ingest.add_noise(ref["unary"], 0.01, seed=11)
    o = orc.MaxSumOracle(ref, np.float32, mode="min").init().step(20)
    assert _index(d, res["assignment"]) == o.value.tolist()
    assert res["violation"] == 0 and res["cost"] == pytest.approx(d.cost(o.value), rel=1e-9)


@pytest.mark.parametrize("variant", ["A", "B", "C"])
def test_dsa_from_yaml_matches_oracle(variant):
    d = ingest.load_yaml(SPLIT)
    res = S.solve(SPLIT, "dsa", {"stop_cycle": 25, "variant": variant}, precision="f64", seed=5)
    o = orc.DsaOracle(d.instance(), np.float64, mode=d.objective, variant=variant, stop_cycle=25,
                      seed=5).init().step(25)
    assert _index(d, res["assignment"]) == o.val.tolist()
    assert res["cycle"] == 25


@pytest.mark.parametrize("precision", ["f32", "f64"])
@pytest.mark.parametrize("kind", ["maxsum", "dsa", "mgm"])
def test_device_solution_cost_equals_the_references_known_answers(kind, precision):
    """fg_solution_cost (pydcop/dcop/dcop.py:319-367) against (violations, cost) computed by the UNMODIFIED
    reference for assignments of problems whose tables AND variable costs hold entries equal to `infinity`
    (tests/golden/solution_cost.json, oracle/make_golden_cost.py) — incl. the reference's own hard
    graph-colouring instance with `-i 10000`."""
    import json
    import torch
    from pydcop_b200 import build_layout
    from pydcop_b200.engine import DsaEngine, MaxSumEngine, MgmEngine
    probs = json.load(open(os.path.join(HERE, "golden", "solution_cost.json")))
    n_viol = 0
    for prob in probs:
        inst = {k: np.asarray(prob[k]) for k in ("dom_size", "factor_ptr", "edge_var", "tables")}
        L = build_layout(**inst)
        eng = {"maxsum": MaxSumEngine, "dsa": DsaEngine, "mgm": MgmEngine}[kind](L, precision=precision)
        for case in prob["cases"]:
            idx = torch.from_numpy(np.asarray(case["value_index"], dtype=np.int32)[L.var_order]).to(eng.device)
            out = eng._solution_cost(idx, prob["infinity"], np.asarray(prob["unary"])).cpu().numpy()
            assert int(out[1]) == case["violation"], (prob["name"], case)
            tol = 1e-12 if precision == "f64" else 2e-6
            assert float(out[0]) == pytest.approx(case["cost"], rel=tol, abs=tol), (prob["name"], case)
            n_viol += case["violation"]
    assert n_viol > 50


def test_solve_reports_device_cost_and_violations_on_a_hard_instance():
    """solve() ends with the device reduction: a hard graph colouring (10000 if equal) solved by DSA —
    cost / violation of the returned assignment equal the host evaluation of the same assignment."""
    import json
    prob = [p for p in json.load(open(os.path.join(HERE, "golden", "solution_cost.json")))
            if p["name"].startswith("graph_coloring")][0]
    arrays = {k: np.asarray(prob[k]) for k in ("dom_size", "factor_ptr", "edge_var", "tables", "unary")}
    d = ingest.from_arrays(arrays)
    for algo, params in (("dsa", {"stop_cycle": 3}), ("maxsum", {"stop_cycle": 2, "noise": 0}), ("mgm", {"stop_cycle": 3})):
        res = S.solve(arrays, algo, params, precision="f64", seed=3, infinity=10000)
        viol, cost = S.solution_cost(d, _index(d, res["assignment"]), 10000)
        assert (res["violation"], res["cost"]) == (viol, pytest.approx(cost)), algo
