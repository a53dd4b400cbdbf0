"""Host side of the MGM engine (no GPU): neighbour lists in the engine's internal order against
the oracle's, rank handling."""
import os

import numpy as np
import pytest

import oracle as orc
from conftest import GOLDEN_DIR, golden_names
from pydcop_b200.engine import distinct_neighbours
from pydcop_b200.generators import random_factor_graph
from pydcop_b200.layout import layout_from_instance


@pytest.mark.parametrize("name", golden_names("mgm_"))
def test_distinct_neighbours_match_oracle_order(name):
    inst, meta = orc.load_golden(os.path.join(GOLDEN_DIR, name + ".npz"))
    o = orc.MgmOracle(inst, np.float64, mode=meta["mode"])
    L = layout_from_instance(inst)
    ptr, idx = distinct_neighbours(L)
    assert ptr[-1] == o.nbr_ptr[-1]
    for vi in range(L.n_vars):           # internal variable vi is canonical variable var_order[vi]
        vc = int(L.var_order[vi])
        mine = [int(L.var_order[u]) for u in idx[ptr[vi]:ptr[vi + 1]]]
        assert mine == o.nbr_idx[o.nbr_ptr[vc]:o.nbr_ptr[vc + 1]].tolist(), (name, vc)


def test_distinct_neighbours_large_random_graph():
    inst = random_factor_graph(3000, 4, 5000, 3, seed=9)
    L = layout_from_instance(inst)
    ptr, idx = distinct_neighbours(L)
    ev, fp = inst["edge_var"], inst["factor_ptr"]
    want = [set() for _ in range(3000)]
    for f in range(5000):
        scope = ev[fp[f]:fp[f + 1]]
        for a in scope:
            want[a].update(int(b) for b in scope if b != a)
    for vi in range(L.n_vars):
        got = [int(L.var_order[u]) for u in idx[ptr[vi]:ptr[vi + 1]]]
        assert len(got) == len(set(got)) and set(got) == want[int(L.var_order[vi])]
    empty = layout_from_instance(dict(dom_size=np.array([2, 3]), factor_ptr=np.array([0]),
                                      edge_var=np.zeros(0, np.int32), tables=np.zeros(0)))
    p, i = distinct_neighbours(empty)
    assert p.tolist() == [0, 0, 0] and len(i) == 1
