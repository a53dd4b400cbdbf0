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
