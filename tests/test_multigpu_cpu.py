"""Multi-GPU host logic on CPU: variable-cut partition, closed local graphs with ghost stubs, and
the halo exchange plumbing run for real over torch.distributed (gloo, world_size 2 and 3) with
index-based pack/unpack stand-ins for the CUDA pack kernels."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT
from pydcop_b200.generators import ising_grid, random_factor_graph
from pydcop_b200.multigpu import HaloExchange, build_shard, variable_owner


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _instance(kind):
    if kind == "grid":
        return ising_grid(6, 8, seed=3)
    inst = random_factor_graph(40, 4, 70, 2, seed=5)
    t = random_factor_graph(40, 4, 12, 3, seed=6)
    inst["edge_var"] = np.concatenate([inst["edge_var"], t["edge_var"]])
    inst["factor_ptr"] = np.concatenate([inst["factor_ptr"], inst["factor_ptr"][-1] + t["factor_ptr"][1:]])
    inst["tables"] = np.concatenate([inst["tables"], t["tables"]])
    return inst


@pytest.mark.parametrize("kind", ["grid", "random"])
@pytest.mark.parametrize("world", [2, 3, 8])
def test_shards_cover_the_graph(kind, world):
    inst = _instance(kind)
    V, E = len(inst["dom_size"]), len(inst["edge_var"])
    owner = variable_owner(V, world)
    assert np.all(np.diff(owner) >= 0) and owner.max() < world
    plans = [build_shard(inst, r, world) for r in range(world)]
    assert sum(p.n_own_vars for p in plans) == V
    # every global edge is a real edge on exactly one rank
    seen = np.concatenate([p.own_factor_edges for p in plans])
    assert sorted(seen.tolist()) == list(range(E))
    # the stub of a cut edge lives on the variable's owner, and send/recv lists pair up
    for a in range(world):
        for b in range(world):
            assert plans[a].send_r_split[b] == plans[b].recv_r_split[a]
            assert plans[a].send_q_split[b] == plans[b].recv_q_split[a]
    n_cut = plans[0].n_cut_edges
    assert sum(len(p.stub_edges) for p in plans) == n_cut
    assert sum(len(p.send_r_len) for p in plans) == n_cut == sum(len(p.send_q_len) for p in plans)
    for p in plans:
        L = p.layout
        ghost_f = [c for c in L.classes if c.tag]
        ghost_v = [c for c in L.var_classes if c.tag == 2]
        bound_v = [c for c in L.var_classes if c.tag == 1]
        # every q row this rank sends belongs to a variable of a boundary class, and every boundary variable sends
        assert sum(c.n_vars for c in bound_v) == len(np.unique(inst["edge_var"][p.stub_edges]))
        assert sum(c.n_factors for c in ghost_f) == len(p.stub_edges)
        assert sum(c.n_slots for c in ghost_v) == len(p.send_r_len)
        # own variables keep their true degree (real + stub edges)
        deg = np.bincount(inst["edge_var"], minlength=V)
        own_deg = np.diff(L.canon_var_ptr)[:p.n_own_vars]
        assert np.array_equal(own_deg, deg[p.own_vars])


def _worker(rank, world, port, kind, q):
    try:
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        dist.init_process_group("gloo", rank=rank, world_size=world)
        inst = _instance(kind)
        plan = build_shard(inst, rank, world)
        L = plan.layout
        E = len(inst["edge_var"])
        dom = inst["dom_size"][inst["edge_var"]]

        def truth(sign, e):   # value of element x of the message on global edge e
            return [sign * (1000.0 * e + x) for x in range(int(dom[e]))]

        qbuf = torch.zeros(L.n_msg_q, dtype=torch.float64)
        rbuf = torch.zeros(L.n_msg, dtype=torch.float64)
        # fill what this rank PRODUCES: r of its real edges, q of its own variables' slots
        canon = np.concatenate([plan.own_factor_edges, plan.stub_edges])   # local canonical -> global
        owner = variable_owner(len(inst["dom_size"]), world)
        for le, ge in enumerate(canon):
            ie = L.edge_perm[le]
            d = int(dom[ge])
            is_real = le < len(plan.own_factor_edges)
            var_mine = owner[inst["edge_var"][ge]] == rank
            if is_real:
                rbuf[L.edge_msg_off[ie]:L.edge_msg_off[ie] + d] = torch.tensor(truth(1, ge))
            if var_mine:
                qbuf[L.edge_qoff[ie]:L.edge_qoff[ie] + d] = torch.tensor(truth(-1, ge))

        def pack(src, packed, row_off, packed_off, row_len, n):
            for i in range(n):
                a, b, ln = int(row_off[i]), int(packed_off[i]), int(row_len[i])
                packed[b:b + ln] = src[a:a + ln]

        def unpack(dst, packed, row_off, packed_off, row_len, n):
            for i in range(n):
                a, b, ln = int(row_off[i]), int(packed_off[i]), int(row_len[i])
                dst[a:a + ln] = packed[b:b + ln]

        halo = HaloExchange(plan, torch.float64, torch.device("cpu"), pack, unpack)
        halo.exchange(qbuf, rbuf)
        # after the exchange EVERY local edge holds both truths
        for le, ge in enumerate(canon):
            ie = L.edge_perm[le]
            d = int(dom[ge])
            assert rbuf[L.edge_msg_off[ie]:L.edge_msg_off[ie] + d].tolist() == truth(1, ge), ("r", rank, ge)
            assert qbuf[L.edge_qoff[ie]:L.edge_qoff[ie] + d].tolist() == truth(-1, ge), ("q", rank, ge)
        # validity flags: producers mark their own, the exchange must deliver them to the ghosts
        qv = torch.zeros(L.n_edges, dtype=torch.uint8)
        rv = torch.zeros(L.n_edges, dtype=torch.uint8)
        for le, ge in enumerate(canon):
            ie = L.edge_perm[le]
            if le < len(plan.own_factor_edges):
                rv[ie] = ge % 2
            if owner[inst["edge_var"][ge]] == rank:
                qv[ie] = (ge // 2) % 2
        halo.exchange_flags(qv, rv)
        for le, ge in enumerate(canon):
            ie = L.edge_perm[le]
            assert int(rv[ie]) == ge % 2 and int(qv[ie]) == (ge // 2) % 2, (rank, ge)
        dist.barrier()
        dist.destroy_process_group()
        q.put((rank, "ok"))
    except Exception as e:  # noqa: BLE001
        import traceback
        q.put((rank, "FAIL " + repr(e) + traceback.format_exc()))


@pytest.mark.parametrize("kind,world", [("grid", 2), ("random", 2), ("random", 3)])
def test_halo_exchange_over_gloo(kind, world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, kind, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=30)
    assert all(r[1] == "ok" for r in results), results


def test_more_ranks_than_variables():
    """Ranks that own nothing get an empty (but valid) shard and empty halo lists."""
    inst = random_factor_graph(5, 3, 6, 2, seed=1)
    plans = [build_shard(inst, r, 8) for r in range(8)]
    assert sum(p.n_own_vars for p in plans) == 5
    assert sorted(np.concatenate([p.own_factor_edges for p in plans]).tolist()) == list(range(12))
    for p in plans[5:]:
        assert p.n_own_vars == 0 and p.layout.n_edges == 0 and sum(p.send_r_split) == 0
    for a in range(8):
        for b in range(8):
            assert plans[a].send_r_split[b] == plans[b].recv_r_split[a]


def test_rows_of_one_peer_land_in_one_block_and_push_order_follows_destinations():
    """Stub factors are grouped by producing rank, so the r rows rank a sends to rank b occupy one
    contiguous block of b's buffer; `destination_order` then makes a's push walk that block in
    address order (what the peer-push kernel needs to coalesce its NVLink stores)."""
    from pydcop_b200.multigpu import destination_order
    inst = random_factor_graph(400, 10, 800, 2, seed=11)
    world = 4
    plans = [build_shard(inst, r, world) for r in range(world)]
    d = 10
    for b, pb in enumerate(plans):
        pos = 0
        for a in range(world):
            n = pb.recv_r_rows[a]
            off = np.asarray(pb.recv_r_off[pos:pos + n])
            if n:
                assert np.array_equal(off, off[0] + d * np.arange(n)), (a, b)   # dense, ascending
            pos += n
    # the sender's permutation: grouped by peer as before, ascending destination inside a group
    rng = np.random.default_rng(0)
    rows = [3, 0, 5, 2]
    dst = rng.permutation(100)[:10] * 40
    perm = destination_order(dst, rows)
    assert sorted(perm.tolist()) == list(range(10))
    pos = 0
    for n in rows:
        grp = perm[pos:pos + n]
        assert set(grp.tolist()) == set(range(pos, pos + n))
        assert (np.diff(dst[grp]) > 0).all()
        pos += n
    # q rows: what a sends to b, taken in destination order, is dense inside every ghost class
    for a, pa in enumerate(plans):
        pos = 0
        for b in range(world):
            n = pa.send_q_rows[b]
            if n:
                pbq = plans[b]
                rpos = sum(pbq.recv_q_rows[:a])
                off = np.sort(np.asarray(pbq.recv_q_off[rpos:rpos + n]))
                gaps = int((np.diff(off) != d).sum())
                assert gaps <= len([c for c in pbq.layout.var_classes if c.tag == 2]), (a, b, gaps)
            pos += n


@pytest.mark.parametrize("kind,world,params,partition", [
    ("random", 2, {}, "blocks"),
    ("random", 4, dict(damping_nodes="vars", start_messages="leafs_vars"), "blocks"),
    ("grid", 3, dict(mode="max", start_messages="all"), "blocks"),
    ("random", 3, {}, "scattered"),          # an arbitrary owner array
    ("random", 4, {}, "multilevel"),
    ("grid", 2, dict(start_messages="all"), "multilevel"),
])
def test_sharded_maxsum_emulated_with_the_oracle_is_bit_identical(kind, world, params, partition):
    """The partition itself (ghost variables, stub factors grouped by producing rank, `links` order
    of every own variable) run cycle by cycle with one oracle per shard and an in-process exchange
    of the boundary rows: messages on every real edge and the assignment must equal the
    single-process oracle bit for bit.  (The GPU version of this test is tests/test_gpu_sharded.py.)"""
    import oracle as orc
    from pydcop_b200.layout import default_var_csr
    inst = _instance(kind)
    V, E = len(inst["dom_size"]), len(inst["edge_var"])
    vp, ve = default_var_csr(V, inst["edge_var"])
    ref = orc.MaxSumOracle(dict(inst, var_ptr=vp, var_edge=ve), np.float64, **params).init()
    if partition == "scattered":
        partition = np.random.default_rng(4).integers(0, world, V).astype(np.int32)
    plans = [build_shard(inst, r, world, partition) for r in range(world)]
    assert sorted(np.concatenate([p.own_vars for p in plans]).tolist()) == list(range(V))
    shards = [orc.MaxSumOracle(p.local_inst, np.float64, **params).init() for p in plans]
    inv = []
    for p in plans:
        L = p.layout
        i = np.zeros(L.n_edges, dtype=np.int64)
        i[L.edge_perm] = np.arange(L.n_edges)
        inv.append(i)

    def exchange():
        for a, pa in enumerate(plans):
            for arr, flg, s_edge, r_edge, s_rows, r_rows in (
                    ("r", "r_flags", "send_r_edge", "recv_r_edge", "send_r_rows", "recv_r_rows"),
                    ("q", "q_flags", "send_q_edge", "recv_q_edge", "send_q_rows", "recv_q_rows")):
                so = np.concatenate([[0], np.cumsum(getattr(pa, s_rows))]).astype(int)
                for b, pb in enumerate(plans):
                    ro = np.concatenate([[0], np.cumsum(getattr(pb, r_rows))]).astype(int)
                    src = inv[a][getattr(pa, s_edge)[so[b]:so[b + 1]]]
                    dst = inv[b][getattr(pb, r_edge)[ro[a]:ro[a + 1]]]
                    assert len(src) == len(dst)
                    oa, ob = shards[a], shards[b]
                    for es, ed in zip(src, dst):
                        d = int(oa.msg_off[es + 1] - oa.msg_off[es])
                        getattr(ob, arr)[ob.msg_off[ed]:ob.msg_off[ed] + d] = \
                            getattr(oa, arr)[oa.msg_off[es]:oa.msg_off[es] + d]
                        fb = getattr(ob, flg)
                        fb[ed] = (int(fb[ed]) & (0xFF ^ orc.FLAG_RECV)) | (int(getattr(oa, flg)[es]) & orc.FLAG_RECV)

    def check(k):
        dom = inst["dom_size"][inst["edge_var"]]
        goff = np.concatenate([[0], np.cumsum(dom)])
        val = np.full(V, -1)
        for p, o in zip(plans, shards):
            val[p.own_vars] = o.value[:p.n_own_vars]
            canon = np.concatenate([p.own_factor_edges, p.stub_edges])
            for le, ge in enumerate(canon):
                d = int(dom[ge])
                assert np.array_equal(o.r[o.msg_off[le]:o.msg_off[le] + d], ref.r[goff[ge]:goff[ge] + d]), (k, "r", ge)
                assert np.array_equal(o.q[o.msg_off[le]:o.msg_off[le] + d], ref.q[goff[ge]:goff[ge] + d]), (k, "q", ge)
        assert np.array_equal(val, ref.value), k

    exchange()
    check(0)
    for k in range(1, 9):
        ref.step()
        for o in shards:
            o.step()
        exchange()
        check(k)


@pytest.mark.parametrize("world,partition", [(2, "blocks"), (4, "multilevel")])
def test_peer_push_tables_move_every_boundary_row_to_its_ghost_row(world, partition):
    """The address tables of the NVLink peer push (multigpu.push_tables) in a simulated address
    space: every rank's q/r buffers are numpy arrays at fake base addresses; executing the push as
    plain copies must fill exactly the rows the all_to_all exchange fills, with the same data."""
    from pydcop_b200.multigpu import push_tables
    inst = random_factor_graph(300, 10, 600, 2, seed=13)
    plans = [build_shard(inst, r, world, partition) for r in range(world)]
    elem, stride = 4, 1 << 40
    base = np.array([[r * 4 * stride + b * stride for b in range(4)] for r in range(world)], dtype=np.int64)
    rng = np.random.default_rng(0)
    bufs = [[rng.integers(1, 1 << 30, size=max(p.layout.n_msg_q, 1)).astype(np.int64),
             None, rng.integers(1, 1 << 30, size=max(p.layout.n_msg, 1)).astype(np.int64), None] for p in plans]
    for b in bufs:
        b[1], b[3] = b[0].copy(), b[2].copy()
    want = [[x.copy() for x in b] for b in bufs]
    d = 10
    # reference: what pack -> all_to_all -> unpack does on buffer index 1 (q[1], r[1])
    for a, pa in enumerate(plans):
        so_r = np.concatenate([[0], np.cumsum(pa.send_r_rows)]).astype(int)
        so_q = np.concatenate([[0], np.cumsum(pa.send_q_rows)]).astype(int)
        for bb, pb in enumerate(plans):
            ro_r = np.concatenate([[0], np.cumsum(pb.recv_r_rows)]).astype(int)
            ro_q = np.concatenate([[0], np.cumsum(pb.recv_q_rows)]).astype(int)
            for s_off, r_off, arr in ((pa.send_r_off[so_r[bb]:so_r[bb + 1]], pb.recv_r_off[ro_r[a]:ro_r[a + 1]], 3),
                                      (pa.send_q_off[so_q[bb]:so_q[bb + 1]], pb.recv_q_off[ro_q[a]:ro_q[a + 1]], 1)):
                for s, t in zip(s_off, r_off):
                    want[bb][arr][t:t + d] = bufs[a][arr][s:s + d]
    # the push: dst offsets as the all_to_all of recv offsets delivers them, then plain copies
    got = [[x.copy() for x in b] for b in bufs]
    for a, pa in enumerate(plans):
        dst_r, dst_q = [], []
        for bb, pb in enumerate(plans):
            ro_r = np.concatenate([[0], np.cumsum(pb.recv_r_rows)]).astype(int)
            ro_q = np.concatenate([[0], np.cumsum(pb.recv_q_rows)]).astype(int)
            dst_r.append(np.asarray(pb.recv_r_off[ro_r[a]:ro_r[a + 1]], dtype=np.int64))
            dst_q.append(np.asarray(pb.recv_q_off[ro_q[a]:ro_q[a + 1]], dtype=np.int64))
        t = push_tables(pa, base, np.concatenate(dst_r), np.concatenate(dst_q), elem)
        # the FUSED halo's per-edge / per-slot destinations describe the same set of (source row -> address) moves:
        # edge e's r row sits at edge_msg_off[e], slot s's q row at edge_qoff[slot_edge[s]]
        from pydcop_b200.multigpu import fused_destinations
        ed, sd = fused_destinations(pa, base, np.concatenate(dst_r), np.concatenate(dst_q), elem)
        La = pa.layout
        for b in range(2):
            e_idx = np.nonzero(ed[b])[0]
            assert sorted(zip(La.edge_msg_off[e_idx].tolist(), ed[b][e_idx].tolist())) == \
                sorted(zip(np.asarray(t["src_r_off"]).tolist(), np.asarray(t["dst_r"][b]).tolist()))
            s_idx = np.nonzero(sd[b])[0]
            assert sorted(zip(La.edge_qoff[La.slot_edge[s_idx]].tolist(), sd[b][s_idx].tolist())) == \
                sorted(zip(np.asarray(t["src_q_off"]).tolist(), np.asarray(t["dst_q"][b]).tolist()))
        for src, dst, arr in ((t["src_r_off"], t["dst_r"][1], 3), (t["src_q_off"], t["dst_q"][1], 1)):
            assert len(src) == len(dst)
            for s, addr in zip(src, dst):
                rank, rem = divmod(int(addr), 4 * stride)
                which, byte = divmod(rem, stride)
                assert which == arr and rank != a and byte % elem == 0
                o = byte // elem
                got[rank][arr][o:o + d] = bufs[a][arr][s:s + d]
            if len(dst) > 1:   # ascending addresses inside every peer group
                peer = np.asarray(dst) // (4 * stride)
                assert (np.diff(np.asarray(dst))[np.diff(peer) == 0] > 0).all()
    for r in range(world):
        for arr in (1, 3):
            assert np.array_equal(got[r][arr], want[r][arr]), (r, arr)


@pytest.mark.parametrize("row_bytes", [40, 80, 16, 8])
def test_push_runs_cover_every_destination_byte_exactly_once(row_bytes):
    """multigpu.push_runs + the unit addressing of k_halo_push_runs (restated here in numpy): executing every
    16-byte unit — full, or its inner 8-byte half at the ends of a run — writes exactly the bytes the per-row
    push writes, from the same source bytes."""
    from pydcop_b200.multigpu import push_runs
    rng = np.random.default_rng(row_bytes)
    # destination rows: a few dense blocks at 8-byte aligned bases, some single rows
    dst = []
    addr = 1 << 20
    for n in (7, 1, 12, 1, 1, 30, 2):
        addr += int(rng.integers(1, 50)) * 8 + n * row_bytes
        dst += [addr + i * row_bytes for i in range(n)]
        addr += n * row_bytes
    dst = np.array(dst, dtype=np.int64)
    n = len(dst)
    src_off = rng.permutation(n * 3)[:n].astype(np.int64) * row_bytes      # byte offsets of the source rows
    src = rng.integers(0, 255, size=3 * n * row_bytes + 64).astype(np.uint8)
    want = {}
    for i in range(n):
        for k in range(row_bytes):
            want[int(dst[i]) + k] = src[src_off[i] + k]
    runs, units = push_runs(dst, row_bytes)
    assert runs[:, 2].sum() == n * row_bytes and runs[0, 3] == 0
    got = {}
    for t in range(units):
        lo = int(np.searchsorted(runs[:, 3], t, side="right") - 1)
        a0, row0, ln, u0 = (int(x) for x in runs[lo])
        unit = (a0 & ~15) + 16 * (t - u0)
        for h in range(2):
            b = unit + 8 * h - a0
            if 0 <= b < ln:
                r, c = divmod(b, row_bytes)
                for k in range(8):
                    key = unit + 8 * h + k
                    assert key not in got
                    got[key] = src[src_off[row0 + r] + c + k]
    assert got == want
