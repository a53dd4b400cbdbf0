"""GPU: the sharded (multi-GPU) MaxSum path emulated on ONE device: `world` shard engines live in
one process on cuda:0 and exchange their halos through an in-process copy that follows exactly the
send/recv splits of the NCCL all_to_all.  The result must be bit-identical to the single-GPU engine
(and therefore to the oracle): same messages on every real edge, same assignment."""
import numpy as np
import pytest

import oracle as orc
from bench import oracle_instance
from pydcop_b200.generators import ising_grid, random_factor_graph

pytestmark = pytest.mark.gpu


class _LocalFabric:
    """all_to_all_single between in-process shards."""

    def __init__(self, shards):
        self.shards = shards

    def exchange(self, which):
        for a, sa in enumerate(self.shards):
            so = np.concatenate([[0], np.cumsum(sa.halo.send_split)]).astype(int)
            for b, sb in enumerate(self.shards):
                ro = np.concatenate([[0], np.cumsum(sb.halo.recv_split)]).astype(int)
                n = so[b + 1] - so[b]
                assert n == ro[a + 1] - ro[a]
                if n:
                    sb.halo.buf_recv[ro[a]:ro[a + 1]] = sa.halo.buf_send[so[b]:so[b + 1]]

    def exchange_flags(self, attr_send, attr_recv, rows_s, rows_r, arr):
        pass


def _run_sharded(inst, world, cycles, precision="f32", **params):
    import torch
    from pydcop_b200.multigpu import ShardedMaxSum
    shards = [ShardedMaxSum(inst, r, world, "cuda:0", precision=precision, **params) for r in range(world)]
    fab = _LocalFabric(shards)

    def halo(bufsel):
        for s in shards:
            q, r = bufsel(s.engine)
            s.halo.pack_rows(q, r)
        torch.cuda.synchronize()
        fab.exchange(None)
        for s in shards:
            q, r = bufsel(s.engine)
            s.halo.unpack_rows(q, r)

    for s in shards:
        s.engine.init()
    halo(lambda e: (e.q[0], e.r[0]))
    # validity flags along the cut edges
    for name_send, name_recv, rows in (("send_r_edge", "recv_r_edge", "r_valid"),
                                       ("send_q_edge", "recv_q_edge", "q_valid")):
        outs = []
        for s in shards:
            idx = torch.from_numpy(getattr(s.plan, name_send).astype(np.int64)).cuda()
            outs.append(getattr(s.engine, rows)[idx].clone())
        rs = "send_r_rows" if rows == "r_valid" else "send_q_rows"
        rr = "recv_r_rows" if rows == "r_valid" else "recv_q_rows"
        for b, sb in enumerate(shards):
            parts = []
            for a, sa in enumerate(shards):
                so = np.concatenate([[0], np.cumsum(getattr(sa.plan, rs))]).astype(int)
                parts.append(outs[a][so[b]:so[b + 1]])
            inn = torch.cat(parts) if parts else None
            idx = torch.from_numpy(getattr(sb.plan, name_recv).astype(np.int64)).cuda()
            assert len(idx) == len(inn) == sum(getattr(sb.plan, rr))
            if len(idx):
                getattr(sb.engine, rows)[idx] = inn
    for _ in range(cycles):
        for s in shards:
            s.engine.cycle_compute()
        halo(lambda e: (e.q[e.cur ^ 1], e.r[e.cur ^ 1]))
        for s in shards:
            s.engine.cycle_commit()
    torch.cuda.synchronize()
    return shards


@pytest.mark.parametrize("world", [2, 4])
@pytest.mark.parametrize("kind", ["grid", "random", "mixed"])
def test_sharded_equals_single_gpu_and_oracle(kind, world):
    from pydcop_b200 import MaxSumEngine, build_layout
    if kind == "grid":
        inst = ising_grid(24, 16, seed=1)
    elif kind == "random":
        inst = random_factor_graph(1500, 10, 3000, 2, seed=2)
    else:
        inst = random_factor_graph(600, 4, 900, 2, seed=3)
        t = random_factor_graph(600, 4, 200, 3, seed=4)
        inst["edge_var"] = np.concatenate([inst["edge_var"], t["edge_var"]])
        inst["factor_ptr"] = np.concatenate([inst["factor_ptr"], inst["factor_ptr"][-1] + t["factor_ptr"][1:]])
        inst["tables"] = np.concatenate([inst["tables"], t["tables"]])
    cycles = 9
    L = build_layout(**inst)
    ref = MaxSumEngine(L, precision="f32").init().step(cycles)
    ref_q, ref_r = ref.messages()
    ref_val = ref.values()[0]
    o = orc.MaxSumOracle(oracle_instance(inst, L), np.float32).init().step(cycles)
    assert np.array_equal(ref_val, o.value)
    shards = _run_sharded(inst, world, cycles)
    dom = inst["dom_size"][inst["edge_var"]]
    off = np.concatenate([[0], np.cumsum(dom)])
    val = np.full(len(inst["dom_size"]), -1)
    for s in shards:
        ids, v = s.local_values()
        val[ids] = v
        q, r = s.engine.messages()             # canonical LOCAL edge order: real edges then stubs
        Ls = s.plan.layout
        loff = np.concatenate([[0], np.cumsum(Ls.canon_dom_size[Ls.canon_edge_var])])
        canon = np.concatenate([s.plan.own_factor_edges, s.plan.stub_edges])
        for le, ge in enumerate(canon):
            d = int(dom[ge])
            assert np.array_equal(r[loff[le]:loff[le] + d], ref_r[off[ge]:off[ge] + d]), ("r", ge)
            assert np.array_equal(q[loff[le]:loff[le] + d], ref_q[off[ge]:off[ge] + d]), ("q", ge)
    assert np.array_equal(val, ref_val)
