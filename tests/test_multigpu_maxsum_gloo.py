"""ShardedMaxSum itself on CPU: the class's own init / compute / exchange / commit loop over
torch.distributed (gloo, world 2 and 3), each rank stepping the generic kernel SOURCE through the
host shim (tests/hostshim/maxsum_generic_host.cpp) instead of a GPU, with index-based stand-ins
for the pack kernels.  The all-gathered assignment of every cycle must equal the single-process
oracle's.  Needs the CUDA toolkit headers (not a GPU)."""
import ctypes as C
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle as orc
from pydcop_b200.generators import ising_grid, random_factor_graph
from pydcop_b200.layout import default_var_csr
from pydcop_b200.multigpu import ShardedMaxSum
from test_maxsum_generic_hostshim import CUDA_INC, HostMaxSum

pytestmark = pytest.mark.skipif(not os.path.exists(os.path.join(CUDA_INC, "cuda_runtime.h")),
                                reason="CUDA toolkit headers not present")


class HostEngine(HostMaxSum):
    """HostMaxSum with MaxSumEngine's attribute surface (torch views of the same buffers)."""

    def __init__(self, layout, precision="f64", **params):
        super().__init__(layout, precision, **params)
        k = self.keep
        self.q = [torch.from_numpy(k["q0"]), torch.from_numpy(k["q1"])]
        self.r = [torch.from_numpy(k["r0"]), torch.from_numpy(k["r1"])]
        self.q_valid, self.r_valid = torch.from_numpy(k["q_valid"]), torch.from_numpy(k["r_valid"])
        self.device = torch.device("cpu")
        self.launch_count = 0
        self.layout = layout

        self.value = torch.from_numpy(k["value"])

    def cycle_compute(self):
        self.lib.ms_host_cycle(C.byref(self.h), self.cur, int(self.cycle == 0))

    def _solution_cost(self, value_tensor, infinity=float("inf"), unary=None, n_vars=None, factor_skip=None,
                       var_skip=None):
        """Host restatement of fg_solution_cost's contract (same arguments as _EngineBase._solution_cost):
        non-ghost classes, the first n_vars internal variables, entries equal to `infinity` counted."""
        L = self.layout
        val = value_tensor.numpy().astype(np.int64)
        ev = np.asarray(L.edge_var, dtype=np.int64)
        tdt = np.asarray(self.keep["tables"]).dtype
        inf_t = np.array(infinity).astype(tdt)
        cost, viol = 0.0, 0
        for c in L.classes:
            if c.tag or not c.n_factors:
                continue
            a, S = c.arity, c.table_size
            scope = ev[c.first_edge:c.first_edge + c.n_factors * a].reshape(c.n_factors, a)
            idx = np.zeros(c.n_factors, dtype=np.int64)
            for j in range(a):
                idx = idx * c.dom[j] + val[scope[:, j]]
            ent = np.asarray(self.keep["tables"])[c.table_base + np.arange(c.n_factors) * S + idx]
            bad = ent == inf_t
            viol += int(bad.sum())
            cost += float(ent[~bad].astype(np.float64).sum())
        if unary is not None:
            n = L.n_vars if n_vars is None else n_vars
            cu = np.asarray(unary, dtype=np.float64)
            dom = L.dom_size.astype(np.int64)               # internal order
            c_dom = dom[L.var_perm]                         # canonical order
            c_off = np.concatenate([[0], np.cumsum(c_dom)])[:-1]
            for v in range(n):                               # internal variable v = canonical L.var_order[v]
                u = cu[c_off[L.var_order[v]] + val[v]].astype(tdt)
                if u == inf_t:
                    viol += 1
                else:
                    cost += float(u)
        return torch.tensor([cost, float(viol)], dtype=torch.float64)

    def cycle_commit(self):
        self.cur ^= 1
        self.cycle += 1


def _pack(src, packed, row_off, packed_off, row_len, n):
    for i in range(n):
        a, b, ln = int(row_off[i]), int(packed_off[i]), int(row_len[i])
        packed[b:b + ln] = src[a:a + ln]


def _unpack(dst, packed, row_off, packed_off, row_len, n):
    for i in range(n):
        a, b, ln = int(row_off[i]), int(packed_off[i]), int(row_len[i])
        dst[a:a + ln] = packed[b:b + ln]


def _instance(kind):
    if kind == "grid":
        return ising_grid(6, 7, seed=2)
    inst = random_factor_graph(36, 3, 50, 2, seed=4)
    t = random_factor_graph(36, 3, 9, 3, seed=5)
    inst["edge_var"] = np.concatenate([inst["edge_var"], t["edge_var"]])
    inst["factor_ptr"] = np.concatenate([inst["factor_ptr"], inst["factor_ptr"][-1] + t["factor_ptr"][1:]])
    inst["tables"] = np.concatenate([inst["tables"], t["tables"]])
    return inst


def _cost_infinity(inst):
    """a table value that occurs: entries equal to it count as violations (dcop.py:355-367)"""
    return float(np.asarray(inst["tables"]).reshape(-1)[3])


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, kind, params, partition, n_cycles, q):
    try:
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        dist.init_process_group("gloo", rank=rank, world_size=world)
        sm = ShardedMaxSum(_instance(kind), rank, world, torch.device("cpu"), precision="f64",
                           partition=partition, engine_factory=HostEngine, pack=_pack, unpack=_unpack,
                           **params).init()
        traj = [sm.values()]
        for _ in range(n_cycles):
            sm.step()
            traj.append(sm.values())
        inst = _instance(kind)
        cost = sm.solution_cost(_cost_infinity(inst), inst["unary"])
        dist.barrier()
        dist.destroy_process_group()
        q.put((rank, "ok", (np.stack(traj), cost)))
    except Exception as e:  # noqa: BLE001
        import traceback
        q.put((rank, "FAIL " + repr(e) + traceback.format_exc(), None))


@pytest.mark.parametrize("kind,world,params,partition", [
    ("random", 2, {}, "blocks"),
    ("random", 3, dict(mode="max", damping_nodes="vars", start_messages="all"), "multilevel"),
    ("grid", 2, dict(start_messages="leafs_vars", stability=0.01), "blocks"),
])
def test_sharded_maxsum_class_over_gloo_with_the_kernel_source(kind, world, params, partition):
    n_cycles = 8
    inst = _instance(kind)
    vp, ve = default_var_csr(len(inst["dom_size"]), inst["edge_var"])
    o = orc.MaxSumOracle(dict(inst, var_ptr=vp, var_edge=ve), np.float64, **params).init()
    want = [o.value.copy()]
    for _ in range(n_cycles):
        o.step()
        want.append(o.value.copy())
    want = np.stack(want)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, kind, params, partition, n_cycles, q))
             for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(timeout=30)
    from pydcop_b200 import ingest, solve as S
    dcop = ingest.from_arrays(inst)
    viol, cost = S.solution_cost(dcop, want[-1], _cost_infinity(inst))
    for rank, status, out in res:
        assert status == "ok", (rank, status)
        traj, got = out
        assert np.array_equal(traj, want), rank
        assert got[1] == viol and viol > 0 and abs(got[0] - cost) <= 1e-9 * max(1.0, abs(cost)), (got, cost, viol)
    assert (np.diff(want, axis=0) != 0).any()
