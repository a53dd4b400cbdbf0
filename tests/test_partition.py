"""The variable partitioner behind the multi-GPU paths (pydcop_b200/partition.py)."""
import numpy as np
import pytest

from pydcop_b200 import partition as P
from pydcop_b200.generators import ising_grid, random_factor_graph


def _loads(owner, edge_var, world):
    deg = np.bincount(edge_var, minlength=len(owner))
    return np.bincount(owner, weights=1 + deg, minlength=world)


@pytest.mark.parametrize("world", [2, 4, 8])
def test_multilevel_cuts_far_fewer_edges_of_a_random_graph_and_stays_balanced(world):
    inst = random_factor_graph(20000, 10, 40000, 2, seed=1)
    fp, ev = inst["factor_ptr"], inst["edge_var"]
    blocks = P.block_owner(20000, world)
    ml = P.multilevel_owner(20000, fp, ev, world)
    assert ml.shape == (20000,) and ml.min() >= 0 and ml.max() == world - 1
    cb, cm = P.edge_cut(blocks, fp, ev), P.edge_cut(ml, fp, ev)
    assert cb == pytest.approx(40000 * (world - 1) / world, rel=0.05)   # blocks: (N-1)/N of the cut-able edges
    assert cm < 0.6 * cb, (cb, cm)
    load = _loads(ml, ev, world)
    assert load.max() <= 1.035 * load.mean()
    assert np.array_equal(ml, P.multilevel_owner(20000, fp, ev, world))      # deterministic
    assert np.array_equal(P.partition_variables(20000, fp, ev, world, "auto"), ml)


def test_auto_keeps_row_strips_on_a_raster_grid_when_they_are_better():
    inst = ising_grid(64, 64, seed=0)
    fp, ev = inst["factor_ptr"], inst["edge_var"]
    n = 64 * 64
    blocks = P.block_owner(n, 4)
    auto = P.partition_variables(n, fp, ev, 4, "auto")
    assert P.edge_cut(auto, fp, ev) <= P.edge_cut(blocks, fp, ev)
    load = _loads(auto, ev, 4)
    assert load.max() <= 1.035 * load.mean()


def test_edge_cases():
    inst = random_factor_graph(50, 3, 60, 3, seed=2)     # arity 3: star edges from the first scope variable
    fp, ev = inst["factor_ptr"], inst["edge_var"]
    assert np.array_equal(P.partition_variables(50, fp, ev, 1), np.zeros(50, np.int32))
    assert np.array_equal(P.partition_variables(7, fp[:1], ev[:0], 4, "auto"), P.block_owner(7, 4))
    ml = P.multilevel_owner(50, fp, ev, 3)
    assert sorted(set(ml.tolist())) == [0, 1, 2]
    assert P.edge_cut(np.zeros(50, np.int32), fp, ev) == 0
    A = P.star_graph(50, fp, ev)
    assert (A != A.T).nnz == 0 and A.diagonal().sum() == 0
    with pytest.raises(ValueError):
        P.partition_variables(50, fp, ev, 2, "metis")
    from pydcop_b200.multigpu import resolve_owner
    with pytest.raises(ValueError):
        resolve_owner(inst, 2, np.full(50, 2))
