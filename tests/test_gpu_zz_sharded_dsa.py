"""GPU: the sharded (multi-GPU) DSA path emulated on ONE device: `world` shard engines live in one
process on cuda:0 and exchange their boundary values through an in-process copy that follows
exactly the send/recv splits of the NCCL all_to_all (the pack / unpack kernels are the real ones).
The assignment of every cycle must equal the single-GPU engine's and the oracle's."""
import numpy as np
import pytest

import oracle as orc
from pydcop_b200.generators import random_factor_graph
from pydcop_b200.layout import default_var_csr

pytestmark = pytest.mark.gpu


def _fabric(shards):
    for a, sa in enumerate(shards):
        so = np.concatenate([[0], np.cumsum(sa.halo.send_split)]).astype(int)
        for b, sb in enumerate(shards):
            ro = np.concatenate([[0], np.cumsum(sb.halo.recv_split)]).astype(int)
            n = so[b + 1] - so[b]
            assert n == ro[a + 1] - ro[a]
            if n:
                sb.halo.buf_recv[ro[a]:ro[a + 1]] = sa.halo.buf_send[so[b]:so[b + 1]]


def _exchange(shards, sel):
    import torch
    for s in shards:
        s.halo.pack_values(sel(s.engine))
    torch.cuda.synchronize()
    _fabric(shards)
    for s in shards:
        s.halo.unpack_values(sel(s.engine))


def _assignment(shards, n):
    val = np.full(n, -1, dtype=np.int64)
    for s in shards:
        ids, v = s.local_values()
        val[ids] = v
    assert (val >= 0).all()
    return val


@pytest.mark.parametrize("world", [2, 4])
@pytest.mark.parametrize("kind,precision,params", [
    ("binary20", "f32", dict(variant="B", seed=3)),                       # the fast DSA kernel
    ("binary20", "f64", dict(variant="C", probability=0.4, seed=4)),
    ("mixed", "f64", dict(variant="A", seed=5, mode="max")),               # generic kernel
    ("mixed", "f32", dict(variant="B", seed=6, p_mode="arity")),
])
def test_sharded_dsa_equals_single_gpu_and_oracle(kind, precision, params, world):
    import torch
    from pydcop_b200.engine import DsaEngine
    from pydcop_b200.layout import build_layout
    from pydcop_b200.multigpu_dsa import ShardedDsa
    if kind == "binary20":
        inst = random_factor_graph(3000, 20, 9000, 2, seed=2, noise=0.0)
    else:
        inst = random_factor_graph(900, 4, 1300, 2, seed=3, noise=0.0)
        t = random_factor_graph(900, 4, 250, 3, seed=4)
        inst["edge_var"] = np.concatenate([inst["edge_var"], t["edge_var"]])
        inst["factor_ptr"] = np.concatenate([inst["factor_ptr"], inst["factor_ptr"][-1] + t["factor_ptr"][1:]])
        inst["tables"] = np.concatenate([inst["tables"], t["tables"]])
    inst["tables"] = np.floor(inst["tables"] / 3.0).astype(np.float32)   # ties: the draws matter
    V = len(inst["dom_size"])
    vp, ve = default_var_csr(V, inst["edge_var"])
    full = dict(inst, var_ptr=vp, var_edge=ve)
    dt = np.float64 if precision == "f64" else np.float32
    cycles = 8
    o = orc.DsaOracle(full, dt, **params).init()
    single = DsaEngine(build_layout(**inst), precision=precision, **params).init()
    shards = [ShardedDsa(inst, r, world, "cuda:0", precision=precision, **params) for r in range(world)]
    for s in shards:
        s.engine.init()
    _exchange(shards, lambda e: e.value[e.cur])
    assert np.array_equal(_assignment(shards, V), o.val)
    assert np.array_equal(single.values(), o.val)
    moved = 0
    for k in range(cycles):
        prev = o.val.copy()
        o.step()
        single.step()
        for s in shards:
            s.engine.cycle_compute()
        _exchange(shards, lambda e: e.value[e.cur ^ 1])
        for s in shards:
            s.engine.cycle_commit()
        torch.cuda.synchronize()
        assert np.array_equal(single.values(), o.val), k
        assert np.array_equal(_assignment(shards, V), o.val), k
        moved += int((prev != o.val).sum())
    assert moved > 0
    assert sum(s.shard.n_own_vars for s in shards) == V and shards[0].shard.n_boundary > 0
