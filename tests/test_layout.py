"""Host layout (pydcop_b200/layout.py): class-major packing is a pure re-ordering.

Checked without a GPU by running the CPU oracle on the canonical instance and on the instance
re-expressed in the layout's internal order, then mapping back."""
import os

import numpy as np
import pytest

import oracle as orc
from conftest import GOLDEN_DIR, golden_names
from pydcop_b200.layout import build_layout, layout_from_instance, default_var_csr


def _internal_instance(L):
    """Re-express the packed layout as a canonical-style instance (factor-major, no padding)."""
    factor_ptr, edge_var, tables, table_off = [0], [], [], [0]
    for c in L.classes:
        for f in range(c.n_factors):
            e0 = c.first_edge + f * c.arity
            edge_var.extend(L.edge_var[e0:e0 + c.arity])
            factor_ptr.append(len(edge_var))
            t0 = c.table_base + f * c.table_size
            tables.append(L.tables[t0:t0 + c.table_size])
            table_off.append(table_off[-1] + c.table_size)
    return dict(dom_size=L.dom_size, factor_ptr=np.array(factor_ptr), edge_var=np.array(edge_var),
                tables=np.concatenate(tables) if tables else np.zeros(0),
                table_off=np.array(table_off),
                unary=np.concatenate([L.unary[L.unary_off[v]:L.unary_off[v] + L.dom_size[v]]
                                      for v in range(L.n_vars)]) if L.n_vars else np.zeros(0),
                var_ptr=L.var_ptr, var_edge=L.slot_edge, init_value=L.init_value)


@pytest.mark.parametrize("name", ["ms_rand_mixed", "ms_arity4_mixed", "ms_secp_simple1", "ms_ising_4x4"])
def test_layout_is_a_pure_permutation(name):
    inst, meta = orc.load_golden(os.path.join(GOLDEN_DIR, name + ".npz"))
    L = layout_from_instance(inst)
    assert sum(c.n_factors for c in L.classes) == L.n_factors
    assert sorted(L.edge_perm.tolist()) == list(range(L.n_edges))
    # class bases are aligned for bulk copies
    for c in L.classes:
        assert c.table_base % 32 == 0 and c.msg_base % 32 == 0
    a = orc.MaxSumOracle(inst, mode=meta["mode"], **meta["params"]).init().step(8)
    b = orc.MaxSumOracle(_internal_instance(L), mode=meta["mode"], **meta["params"]).init().step(8)
    # map b's (internal, unpadded) messages back to canonical order
    d_int = L.dom_size[L.edge_var]
    off_int = np.concatenate([[0], np.cumsum(d_int)])
    d_can = np.diff(L.canon_msg_off)
    start = off_int[L.edge_perm] - L.canon_msg_off[:-1]
    g = np.repeat(start, d_can) + np.arange(L.n_msg_canonical)
    assert np.array_equal(a.q, b.q[g]) and np.array_equal(a.r, b.r[g])
    assert np.array_equal(a.value, L.vars_to_canonical(b.value))
    assert np.array_equal(L.canonical_unary(), np.asarray(inst["unary"], dtype=np.float64))
    assert np.array_equal(a.r_sent, b.r_sent[L.edge_perm])
    # the padded gather index used for device readback addresses the same rows
    gi = L.message_gather_index()
    assert len(gi) == L.n_msg_canonical and gi.max(initial=-1) < L.n_msg
    assert np.array_equal(np.diff(gi)[np.diff(np.repeat(np.arange(L.n_edges), d_can)) == 0], 
                          np.ones((np.diff(np.repeat(np.arange(L.n_edges), d_can)) == 0).sum()))


def test_default_var_csr_matches_reference_links_order():
    for name in golden_names("ms_"):
        inst, _ = orc.load_golden(os.path.join(GOLDEN_DIR, name + ".npz"))
        vp, ve = default_var_csr(len(inst["dom_size"]), inst["edge_var"])
        assert np.array_equal(vp, inst["var_ptr"]) and np.array_equal(ve, inst["var_edge"]), name
        L = layout_from_instance(inst)
        assert sum(c.n_slots for c in L.var_classes) == L.n_edges
        assert sum(c.n_vars for c in L.var_classes) == L.n_vars


def test_empty_and_degenerate_graphs():
    L = build_layout([3, 2], [0], [], [])            # no factors at all
    assert L.n_edges == 0 and L.n_msg == 0 and len(L.classes) == 0
    L = build_layout([3], [0, 1], [0], [1.0, 2.0, 3.0])  # single unary factor
    assert L.classes[0].arity == 1 and L.classes[0].dom == (3,)
    with pytest.raises(ValueError):
        build_layout([3], [0, 1], [0], [1.0, 2.0])   # wrong table size
    with pytest.raises(ValueError):
        build_layout([3], [0, 1], [1], [1.0, 2.0, 3.0])  # variable out of range
    with pytest.raises(ValueError):
        build_layout([300], [0, 1], [0], np.zeros(300))  # domain too large


def test_layout_scales_vectorised():
    rng = np.random.default_rng(0)
    V, F, d = 20000, 40000, 10
    ev = np.stack([rng.integers(0, V, F), rng.integers(0, V, F)], 1).reshape(-1)
    L = build_layout(np.full(V, d), np.arange(F + 1) * 2, ev, rng.integers(0, 10, F * d * d))
    assert L.n_edges == 2 * F and len(L.classes) == 1 and L.n_msg >= 2 * F * d


def test_table_packing_takes_slices_for_contiguous_runs_and_gathers_otherwise():
    """layout._gather_tables: one run -> a view of the caller's array, a few runs -> slices, scattered ->
    index gather; the packed tables are the same whichever route was taken."""
    from pydcop_b200.layout import _gather_tables
    t = np.arange(1000, dtype=np.float32)
    one = _gather_tables(t, np.arange(5, dtype=np.int64) * 20 + 40, 20)
    assert one.base is t and np.array_equal(one, t[40:140])
    starts = np.concatenate([np.arange(130, dtype=np.int64) * 4, [800, 804, 808]])
    few = _gather_tables(t, starts, 4)
    assert np.array_equal(few, np.concatenate([t[0:520], t[800:812]]))
    scat = _gather_tables(t, np.array([900, 10, 500, 20], dtype=np.int64), 10)
    assert np.array_equal(scat, np.concatenate([t[900:910], t[10:20], t[500:510], t[20:30]]))
    assert len(_gather_tables(t, np.zeros(0, np.int64), 7)) == 0
    # end to end: two interleaved classes (scattered) and one class (view) give the tables the engine expects
    rng = np.random.default_rng(3)
    dom = np.array([2, 3] * 10, dtype=np.int32)
    fp, ev, tabs = [0], [], []
    for k in range(12):
        a, b = (0, 2) if k % 2 == 0 else (1, 3)          # (2,2) and (3,3) tables alternate in the input
        ev += [a + 4 * (k % 3), b + 4 * (k % 3)]
        fp.append(len(ev))
        tabs.append(rng.uniform(size=int(dom[ev[-2]] * dom[ev[-1]])).astype(np.float32) + k)
    L = build_layout(dom, np.array(fp), np.array(ev, np.int32), np.concatenate(tabs))
    off = np.concatenate([[0], np.cumsum([len(x) for x in tabs])])
    for c in L.classes:
        for i in range(c.n_factors):
            canon = int(np.nonzero(L.factor_perm == c.first_factor + i)[0][0])
            assert np.array_equal(L.tables[c.table_base + i * c.table_size:c.table_base + (i + 1) * c.table_size],
                                  tabs[canon]), (c.dom, i)


def test_stable_group_order_equals_stable_argsort():
    """the counting sort behind the packing (scipy's COO -> CSR kernel) == np.argsort(kind='stable') + group pointers,
    and so does the fallback taken when scipy's private kernel is missing"""
    from pydcop_b200 import layout as LY
    rng = np.random.default_rng(3)
    for n, g in ((0, 0), (1, 1), (17, 5), (5000, 37), (20000, 20000)):
        keys = rng.integers(0, max(g, 1), n).astype(np.int32)
        order, ptr = LY.stable_group_order(keys, g)
        want = np.argsort(keys, kind="stable")
        assert np.array_equal(order, want) and order.dtype == np.int32
        assert np.array_equal(np.diff(ptr), np.bincount(keys, minlength=g)) and ptr[0] == 0
    # fallback path
    import scipy.sparse._sparsetools as st
    saved = st.coo_tocsr
    try:
        st.coo_tocsr = None
        keys = rng.integers(0, 9, 300).astype(np.int64)
        order, ptr = LY.stable_group_order(keys, 9)
        assert np.array_equal(order, np.argsort(keys, kind="stable")) and ptr[-1] == 300
    finally:
        st.coo_tocsr = saved
