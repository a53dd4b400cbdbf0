"""Sharded DSA host logic on CPU: the partition, the closed local problems, and the whole
init / compute / exchange / commit loop over torch.distributed (gloo, world 2 and 3) with the
oracle standing in for the GPU engine and index-based stand-ins for the CUDA pack kernels.
The sharded trajectory must equal the single-process one bit for bit."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle as orc
from pydcop_b200.generators import random_factor_graph
from pydcop_b200.layout import default_var_csr
from pydcop_b200.multigpu import variable_owner
from pydcop_b200.multigpu_dsa import ShardedDsa, boundary_pairs, build_dsa_shard


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _instance(kind):
    if kind == "binary":       # one domain size, binary: the engine's fast DSA shape
        inst = random_factor_graph(60, 4, 130, 2, seed=5, noise=0.0)
    else:                      # mixed arities incl. unary constraints, isolated variables
        inst = random_factor_graph(50, 3, 45, 2, seed=7, noise=0.3)
        t = random_factor_graph(50, 3, 14, 3, seed=8)
        u = random_factor_graph(50, 3, 6, 1, seed=9)
        for extra in (t, u):
            inst["edge_var"] = np.concatenate([inst["edge_var"], extra["edge_var"]])
            inst["factor_ptr"] = np.concatenate([inst["factor_ptr"],
                                                 inst["factor_ptr"][-1] + extra["factor_ptr"][1:]])
            inst["tables"] = np.concatenate([inst["tables"], extra["tables"]])
    # few cost levels -> many ties -> the random choices matter
    inst["tables"] = np.floor(inst["tables"] / 4.0).astype(np.float32)
    vp, ve = default_var_csr(len(inst["dom_size"]), inst["edge_var"])
    return dict(inst, var_ptr=vp, var_edge=ve)


class FakeDsaEngine:
    """DsaEngine's driving surface (value double buffer in the layout's INTERNAL variable order,
    cur, cycle, init / cycle_compute / cycle_commit / values) on top of the oracle."""

    def __init__(self, layout, inst, precision="f64", var_global_id=None, frozen=None,
                 isolated_value=None, **params):
        self.layout = layout
        self.perm = torch.from_numpy(np.asarray(layout.var_perm, dtype=np.int64))
        self.o = orc.DsaOracle(inst, np.float64 if precision == "f64" else np.float32,
                               var_id=var_global_id, frozen=frozen, **params)
        n = max(layout.n_vars, 1)
        self.value = [torch.zeros(n, dtype=torch.int32), torch.zeros(n, dtype=torch.int32)]
        self.cur = 0
        self.launch_count = 0

    @property
    def cycle(self):
        return self.o.cycle

    def _store(self, b, canonical):
        self.value[b][self.perm] = torch.from_numpy(np.ascontiguousarray(canonical))

    def _load(self, b):
        return self.value[b][self.perm].numpy().astype(np.int32)

    def init(self):
        self.o.init()
        self._store(0, self.o.val)
        self._store(1, self.o.val)
        self.cur = 0
        return self

    def cycle_compute(self):
        if self.o.stopped:
            return
        self.o.val = self._load(self.cur)          # includes the ghosts the exchange filled
        self.o.compute()
        self._store(self.cur ^ 1, self.o.val_next)

    def cycle_commit(self):
        if self.o.stopped:
            return
        self.o.val_next = self._load(self.cur ^ 1)  # ghosts of the next buffer come from the peers
        self.o.commit()
        self.cur ^= 1

    def values(self):
        return self._load(self.cur)


def _pack(src, packed, row_off, packed_off, row_len, n):
    for i in range(n):
        packed[int(packed_off[i])] = src[int(row_off[i])]


def _unpack(dst, packed, row_off, packed_off, row_len, n):
    for i in range(n):
        dst[int(row_off[i])] = packed[int(packed_off[i])]


@pytest.mark.parametrize("kind", ["binary", "mixed"])
@pytest.mark.parametrize("world", [2, 3, 8])
def test_dsa_shards_are_closed_and_halo_lists_pair_up(kind, world):
    inst = _instance(kind)
    V = len(inst["dom_size"])
    owner = variable_owner(V, world)
    shards = [build_dsa_shard(inst, r, world) for r in range(world)]
    assert sum(s.n_own_vars for s in shards) == V
    deg = np.bincount(inst["edge_var"], minlength=V)
    u, r = boundary_pairs(np.asarray(inst["edge_var"], np.int64), np.asarray(inst["factor_ptr"], np.int64), owner)
    assert len(u) == shards[0].n_boundary == sum(len(s.recv_var) for s in shards)
    for a, s in enumerate(shards):
        L = s.layout
        # owned variables keep every constraint, in order; ghosts are exactly the remote neighbours
        own_deg = np.diff(s.inst["var_ptr"])[:s.n_own_vars]
        assert np.array_equal(own_deg, deg[s.own_vars])
        assert np.array_equal(s.local_global_id[:s.n_own_vars], s.own_vars)
        ghosts = s.local_global_id[s.n_own_vars:]
        assert (owner[ghosts] != a).all() and s.frozen[s.n_own_vars:].all() and not s.frozen[:s.n_own_vars].any()
        assert sorted(ghosts.tolist()) == sorted(u[r == a].tolist())
        assert L.n_vars == len(s.local_global_id)
        for b in range(world):
            assert s.send_split[b] == shards[b].recv_split[a]
        assert s.send_split[a] == 0 and s.recv_split[a] == 0
        # what a sends to b are the global ids b expects, in the same order
        pos = 0
        for b in range(world):
            mine = s.local_global_id[s.send_var[pos:pos + s.send_split[b]]]
            rb = shards[b]
            off = sum(rb.recv_split[:a])
            theirs = rb.local_global_id[rb.recv_var[off:off + rb.recv_split[a]]]
            assert np.array_equal(mine, theirs)
            pos += s.send_split[b]


def _worker(rank, world, port, kind, params, n_cycles, q, partition="blocks"):
    try:
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        dist.init_process_group("gloo", rank=rank, world_size=world)
        inst = _instance(kind)
        sd = ShardedDsa(inst, rank, world, torch.device("cpu"), precision="f64",
                        engine_factory=FakeDsaEngine, pack=_pack, unpack=_unpack, partition=partition,
                        **params).init()
        traj = [sd.values()]
        for _ in range(n_cycles):
            sd.step()
            traj.append(sd.values())
        dist.barrier()
        dist.destroy_process_group()
        q.put((rank, "ok", np.stack(traj), sd.cycle))
    except Exception as e:  # noqa: BLE001
        import traceback
        q.put((rank, "FAIL " + repr(e) + traceback.format_exc(), None, None))


@pytest.mark.parametrize("kind,world,params,partition", [
    ("binary", 2, dict(variant="B", seed=11), "blocks"),
    ("mixed", 2, dict(variant="A", seed=12, mode="max"), "blocks"),
    ("mixed", 3, dict(variant="C", probability=0.5, seed=13, p_mode="arity"), "blocks"),
    ("binary", 3, dict(variant="B", seed=14, stop_cycle=5), "blocks"),
    ("binary", 3, dict(variant="B", seed=15), "multilevel"),
])
def test_sharded_dsa_equals_single_process_over_gloo(kind, world, params, partition):
    n_cycles = 9
    inst = _instance(kind)
    if params.get("p_mode") == "arity":   # (divides by zero for isolated variables, dsa.py:258-260)
        pass
    o = orc.DsaOracle(inst, np.float64, **params).init()
    want = [o.val.copy()]
    for _ in range(n_cycles):
        o.step()
        want.append(o.val.copy())
    want = np.stack(want)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, kind, params, n_cycles, q, partition))
             for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=30)
    for rank, status, traj, cycle in res:
        assert status == "ok", (rank, status)
        assert np.array_equal(traj, want), rank
        assert cycle == o.cycle
    assert len({tuple(map(tuple, w)) for w in [want]}) == 1 and (np.diff(want, axis=0) != 0).any()


@pytest.mark.parametrize("world,partition", [(3, "blocks"), (4, "multilevel")])
def test_value_push_tables_fill_every_ghost(world, partition):
    """The address tables of the DSA peer push in a simulated address space: after executing the
    push as plain copies every ghost holds its owner's value."""
    from pydcop_b200.multigpu_dsa import value_push_tables
    inst = _instance("binary")
    V = len(inst["dom_size"])
    shards = [build_dsa_shard(inst, r, world, partition) for r in range(world)]
    stride = 1 << 40
    base = np.array([[r * 2 * stride + b * stride for b in range(2)] for r in range(world)], dtype=np.int64)
    truth = np.arange(V) * 7 + 3                       # "value" of every global variable
    bufs = []
    for s in shards:                                   # internal order; ghosts start out wrong
        perm = np.asarray(s.layout.var_perm, dtype=np.int64)
        b = np.full(s.layout.n_vars, -1, dtype=np.int64)
        b[perm[:s.n_own_vars]] = truth[s.own_vars]
        bufs.append(b)
    for a, sa in enumerate(shards):
        dst_idx = []
        for b, sb in enumerate(shards):
            off = sum(sb.recv_split[:a])
            perm_b = np.asarray(sb.layout.var_perm, dtype=np.int64)
            dst_idx.append(perm_b[sb.recv_var[off:off + sb.recv_split[a]]])
        t = value_push_tables(sa, base, np.concatenate(dst_idx) if dst_idx else np.zeros(0, np.int64))
        for src, addr in zip(t["src"], t["dst"][1]):
            rank, rem = divmod(int(addr), 2 * stride)
            which, byte = divmod(rem, stride)
            assert which == 1 and rank != a and byte % 4 == 0
            bufs[rank][byte // 4] = bufs[a][src]
    for s, b in zip(shards, bufs):
        perm = np.asarray(s.layout.var_perm, dtype=np.int64)
        assert np.array_equal(b[perm], truth[s.local_global_id])
