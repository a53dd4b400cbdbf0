"""GPU: the sharded MaxSum path with a NON-contiguous variable partition (the multilevel split that
bench.py uses at N > 1) emulated on one device, against the single-GPU engine: messages on every
real edge and the assignment must be bit-identical whatever the owner array is."""
import numpy as np
import pytest

from pydcop_b200.generators import random_factor_graph
from test_gpu_sharded import _run_sharded

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("world,partition", [(2, "multilevel"), (4, "multilevel"), (3, "scattered")])
def test_partition_does_not_change_the_trajectory(world, partition):
    from pydcop_b200 import MaxSumEngine, build_layout
    inst = random_factor_graph(3000, 10, 6000, 2, seed=8)
    V = 3000
    if partition == "scattered":
        partition = np.random.default_rng(1).integers(0, world, V).astype(np.int32)
    cycles = 9
    ref = MaxSumEngine(build_layout(**inst), precision="f32").init().step(cycles)
    ref_q, ref_r = ref.messages()
    shards = _run_sharded(inst, world, cycles, partition=partition)
    dom = inst["dom_size"][inst["edge_var"]]
    off = np.concatenate([[0], np.cumsum(dom)])
    val = np.full(V, -1)
    n_cut = shards[0].plan.n_cut_edges
    for s in shards:
        ids, v = s.local_values()
        val[ids] = v
        q, r = s.engine.messages()
        Ls = s.plan.layout
        loff = np.concatenate([[0], np.cumsum(Ls.canon_dom_size[Ls.canon_edge_var])])
        canon = np.concatenate([s.plan.own_factor_edges, s.plan.stub_edges])
        for le, ge in enumerate(canon):
            d = int(dom[ge])
            assert np.array_equal(r[loff[le]:loff[le] + d], ref_r[off[ge]:off[ge] + d]), ("r", ge)
            assert np.array_equal(q[loff[le]:loff[le] + d], ref_q[off[ge]:off[ge] + d]), ("q", ge)
    assert np.array_equal(val, ref.values()[0])
    if isinstance(partition, str):   # the multilevel split cuts far fewer edges than blocks would
        assert n_cut < 0.7 * 6000 * (world - 1) / world
