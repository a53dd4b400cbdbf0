"""The reference's own unit tests of the algorithm-module surface, re-run against the GPU modules:
tests/unit/test_algorithms_maxsum.py:49-101 (computation creation), test_algorithms_amaxsum.py:205-300
(memory / communication models of the factor graph), test_algorithms_dsa.py:43-72,128-206 and
test_algorithms_mgm.py:38-70 (load, memory, neighbour sets of the constraints hypergraph).
Needs the reference (build container)."""
import pytest

import ref_shim

pytestmark = pytest.mark.skipif(not ref_shim.reference_available(), reason="reference tree not present")


@pytest.fixture(scope="module")
def mods():
    ref_shim.install()
    from pydcop_b200 import launcher
    launcher.install()
    from pydcop.algorithms import load_algorithm_module
    from pydcop_b200.algorithms._session import GpuSession
    GpuSession.reset()
    yield {n: load_algorithm_module(n) for n in ("maxsum_gpu", "dsa_gpu", "mgm_gpu", "maxsum", "dsa", "mgm")}
    GpuSession.reset()


def test_maxsum_computation_creation(mods):
    from pydcop.algorithms import AlgorithmDef, ComputationDef
    from pydcop.computations_graph.factor_graph import build_computation_graph
    from pydcop.dcop.objects import Domain, Variable
    from pydcop.dcop.relations import constraint_from_str
    d = Domain("d", "", ["R", "G"])
    v1, v2 = Variable("v1", d), Variable("v2", d)
    c1 = constraint_from_str("c1", "10 if v1 == v2 else 0", [v1, v2])
    graph = build_computation_graph(None, constraints=[c1], variables=[v1, v2])
    algo = AlgorithmDef.build_with_default_param("maxsum_gpu", {"session": "unit_creation"})
    comp = mods["maxsum_gpu"].build_computation(ComputationDef(graph.computation("c1"), algo))
    assert comp is not None and comp.name == "c1" and comp.factor == c1
    comp = mods["maxsum_gpu"].build_computation(ComputationDef(graph.computation("v1"), algo))
    assert comp is not None and comp.name == "v1" and comp.variable.name == "v1" and comp.factors == ["c1"]


def test_maxsum_memory_and_communication_known_answers(mods):
    from pydcop.computations_graph.factor_graph import FactorComputationNode, VariableComputationNode
    from pydcop.dcop.objects import Variable, VariableDomain
    from pydcop.dcop.relations import relation_from_str
    m, ref = mods["maxsum_gpu"], mods["maxsum"]
    assert (m.FACTOR_UNIT_SIZE, m.VARIABLE_UNIT_SIZE, m.UNIT_SIZE, m.HEADER_SIZE) == \
        (ref.FACTOR_UNIT_SIZE, ref.VARIABLE_UNIT_SIZE, ref.UNIT_SIZE, ref.HEADER_SIZE)
    d1 = VariableDomain("d1", "", [1, 2, 3, 5])
    v1 = Variable("v1", d1)
    assert m.computation_memory(VariableComputationNode(v1, [])) == 0
    f1 = relation_from_str("f1", "v1 * 0.5", [v1])
    cv1, cf1 = VariableComputationNode(v1, ["f1"]), FactorComputationNode(f1)
    assert m.computation_memory(cv1) == m.VARIABLE_UNIT_SIZE * 4
    assert m.computation_memory(cf1) == m.FACTOR_UNIT_SIZE * 4
    assert m.computation_memory(VariableComputationNode(v1, ["f1", "f2"])) == m.VARIABLE_UNIT_SIZE * 4 * 2
    d5, d3 = VariableDomain("d1", "", [1, 2, 3, 4, 5]), VariableDomain("d1", "", [1, 2, 3])
    w1, w2 = Variable("v1", d5), Variable("v2", d3)
    f2 = relation_from_str("f1", "v1 * 0.5 + v2", [w1, w2])
    assert m.computation_memory(FactorComputationNode(f2)) == m.FACTOR_UNIT_SIZE * (5 + 3)
    # communication: one cost per value of the receiving / sending variable
    assert m.communication_load(cv1, "f1") == ref.communication_load(cv1, "f1") == m.UNIT_SIZE * 4 + m.HEADER_SIZE
    assert m.communication_load(cf1, "v1") == ref.communication_load(cf1, "v1") == m.UNIT_SIZE * 4 + m.HEADER_SIZE
    cf2 = FactorComputationNode(f2)
    assert m.communication_load(cf2, "v1") == m.UNIT_SIZE * 5 + m.HEADER_SIZE
    assert m.communication_load(cf2, "v2") == m.UNIT_SIZE * 3 + m.HEADER_SIZE
    with pytest.raises(ValueError):
        m.communication_load(cf2, "v9")


@pytest.mark.parametrize("name", ["dsa", "mgm"])
def test_hypergraph_modules_known_answers(mods, name):
    from pydcop.algorithms import AlgorithmDef, ComputationDef
    from pydcop.computations_graph.constraints_hypergraph import VariableComputationNode
    from pydcop.dcop.objects import Variable
    from pydcop.dcop.relations import constraint_from_str
    m, ref = mods[name + "_gpu"], mods[name]
    assert (m.UNIT_SIZE, m.HEADER_SIZE) == (ref.UNIT_SIZE, ref.HEADER_SIZE)
    v = Variable("v1", list(range(10)))
    assert m.communication_load(VariableComputationNode(v, []), "f1") == m.UNIT_SIZE + m.HEADER_SIZE
    v1, v2, v3, v4 = (Variable(n, list(range(10))) for n in ("v1", "v2", "v3", "v4"))
    c1 = constraint_from_str("c1", " v1 + v2 == v3", [v1, v2, v3])
    assert m.computation_memory(VariableComputationNode(v1, [c1])) == m.UNIT_SIZE * 2      # one hyper-edge, 3 vertices
    cs = [constraint_from_str("c1", " v1 == v2", [v1, v2]), constraint_from_str("c2", " v1 == v3", [v1, v3]),
          constraint_from_str("c3", " v1 == v4", [v1, v4])]
    assert m.computation_memory(VariableComputationNode(v1, cs)) == m.UNIT_SIZE * 3
    # computations: neighbour sets as in test_algorithms_dsa.py:128-206
    algo = AlgorithmDef.build_with_default_param(name + "_gpu", {"session": "unit_" + name})
    u1 = constraint_from_str("u1", " v1 * 0.5", [v1])
    comp = m.build_computation(ComputationDef(VariableComputationNode(v1, [u1]), algo))
    assert comp.name == "v1" and len(comp.neighbors) == 0
    b12 = constraint_from_str("b12", " v1 - v2", [v1, v2])
    comp = m.build_computation(ComputationDef(VariableComputationNode(v1, [b12, cs[0]]), algo))
    assert set(comp.neighbors) == {"v2"}
    comp = m.build_computation(ComputationDef(VariableComputationNode(v1, [c1]), algo))
    assert set(comp.neighbors) == {"v2", "v3"}
    assert comp.footprint() == m.computation_memory(VariableComputationNode(v1, [c1]))
