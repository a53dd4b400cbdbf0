"""Which GPU owns which variable: a communication-minimising, balanced k-way split of the factor
graph for the multi-GPU paths (pydcop_b200/multigpu.py, multigpu_dsa.py).

The halo a rank exchanges every cycle is one row per CUT edge — an edge whose factor lives with
another rank than its variable (a factor lives with its first scope variable) — so the exchange
volume is the edge cut of the variable partition.  Contiguous blocks of variable ids are ideal for
raster-ordered grids (row strips) and the worst case for random graphs ((N-1)/N of the edges cut).
`partition_variables` runs a small multilevel scheme (heavy-edge matching, greedy growing on the
coarsest graph, balanced label-propagation refinement on the way back up) in vectorised
numpy / scipy.sparse and keeps whichever of {contiguous blocks, multilevel} cuts fewer edges.

pyDcop has the same concern one level up: its distribution methods place computations on agents to
minimise communication load under capacity constraints (pydcop/distribution/ilp_fgdp.py,
heur_comhost.py, with the algorithms' `communication_load`, maxsum.py:166-209).  This is that
idea for ranks of one box; it does not touch the algorithm: any owner array gives bit-identical
results (tests/test_multigpu_cpu.py runs the partition emulation with arbitrary owners).
"""
import numpy as np


def star_graph(n_vars, factor_ptr, edge_var):
    """Symmetric weighted adjacency (scipy CSR) whose edge weights count the factor->variable edges
    that are cut when the two endpoints are separated: (first scope variable, other scope variable)
    for every factor (a factor lives with its first scope variable)."""
    import scipy.sparse as sp
    factor_ptr = np.asarray(factor_ptr, dtype=np.int64)
    edge_var = np.asarray(edge_var, dtype=np.int64)
    arity = np.diff(factor_ptr)
    first = np.repeat(edge_var[factor_ptr[:-1]], arity) if len(arity) else np.zeros(0, np.int64)
    other = edge_var
    keep = first != other
    a, b = first[keep], other[keep]
    w = np.ones(len(a), dtype=np.float64)
    A = sp.coo_matrix((np.concatenate([w, w]), (np.concatenate([a, b]), np.concatenate([b, a]))),
                      shape=(n_vars, n_vars)).tocsr()
    A.sum_duplicates()
    return A


def edge_cut(owner, factor_ptr, edge_var) -> int:
    """Number of cut edges (= halo rows per direction per cycle) of an owner array."""
    factor_ptr = np.asarray(factor_ptr, dtype=np.int64)
    edge_var = np.asarray(edge_var, dtype=np.int64)
    owner = np.asarray(owner)
    arity = np.diff(factor_ptr)
    if not len(arity):
        return 0
    fo = np.repeat(owner[edge_var[factor_ptr[:-1]]], arity)
    return int((fo != owner[edge_var]).sum())


def block_owner(n_vars: int, world: int) -> np.ndarray:
    """Contiguous blocks of ceil(V / world) variables (the default of round 1)."""
    per = -(-n_vars // world) if world > 0 else n_vars
    return (np.arange(n_vars, dtype=np.int64) // max(per, 1)).astype(np.int32)


def _match(A, w, max_w, rng, rounds=3):
    """Heavy-edge handshake matching: every unmatched vertex proposes to its heaviest unmatched
    neighbour (random tie break); mutual proposals are matched.  Returns coarse ids.
    Linear in the number of edges per round: the best neighbour of a row is a segmented maximum over the
    CSR entries (np.maximum.reduceat), not a sort."""
    import scipy.sparse as sp
    A = A.tocsr()
    n = A.shape[0]
    indptr, indices, data = A.indptr.astype(np.int64), A.indices.astype(np.int64), A.data
    deg = np.diff(indptr)
    row = np.repeat(np.arange(n, dtype=np.int64), deg)
    nonempty = np.nonzero(deg > 0)[0]
    mate = np.full(n, -1, dtype=np.int64)
    wsum_ok = (w[row] + w[indices]) <= max_w
    for _ in range(rounds):
        free = mate < 0
        if free.sum() < 2 or not len(indices):
            break
        # restrict to edges between free vertices whose merged weight stays bounded
        ok = free[row] & free[indices] & wsum_ok
        if not ok.any():
            break
        score = np.where(ok, data + rng.random(len(data)) * 0.5, -1.0)     # heavy edge first, random ties
        rowmax = np.full(n, -1.0)
        rowmax[nonempty] = np.maximum.reduceat(score, indptr[:-1][nonempty])
        pos = np.flatnonzero(ok & (score == rowmax[row]))                    # sorted by row
        if not len(pos):
            break
        r = row[pos]
        first = np.ones(len(pos), dtype=bool)
        first[1:] = r[1:] != r[:-1]
        prop = np.full(n, -1, dtype=np.int64)
        prop[r[first]] = indices[pos[first]]
        v = np.nonzero(prop >= 0)[0]
        mutual = v[prop[prop[v]] == v]
        mate[mutual] = prop[mutual]
    rep = np.where((mate >= 0) & (mate < np.arange(n)), mate, np.arange(n))   # smaller id represents
    uniq, cid = np.unique(rep, return_inverse=True)
    P = sp.csr_matrix((np.ones(n), (np.arange(n), cid)), shape=(n, len(uniq)))
    return cid, P


def _refine(A, w, label, k, cap, rng, iters):
    """Balanced label propagation: a vertex moves to the part it is most connected to when that
    lowers the cut and the target has room; half of the candidates move per sweep (no swaps of
    adjacent vertices in lock-step)."""
    import scipy.sparse as sp
    n = A.shape[0]
    for _ in range(iters):
        H = sp.csr_matrix((np.ones(n), (np.arange(n), label)), shape=(n, k))
        S = np.asarray((A @ H).todense())
        cur = S[np.arange(n), label]
        S[np.arange(n), label] = -1.0
        best = S.argmax(axis=1)
        gain = S[np.arange(n), best] - cur
        cand = np.nonzero((gain > 0) & (rng.random(n) < 0.5))[0]
        if not len(cand):
            break
        load = np.bincount(label, weights=w, minlength=k)
        moved = 0
        order = cand[np.argsort(-gain[cand], kind="stable")]
        tgt = best[order]
        for p in range(k):
            mv = order[tgt == p]
            if not len(mv):
                continue
            room = cap - load[p]
            take = mv[np.cumsum(w[mv]) <= room]
            if len(take):
                np.subtract.at(load, label[take], w[take])
                load[p] += w[take].sum()
                label[take] = p
                moved += len(take)
        if not moved:
            break
    return label


def _grow(A, w, k, cap, rng):
    """Greedy growing on the coarsest graph: k random seeds, unassigned vertices join the part they
    are most connected to (room permitting), the rest is filled by lightest part."""
    import scipy.sparse as sp
    n = A.shape[0]
    label = np.full(n, -1, dtype=np.int64)
    seeds = rng.choice(n, size=min(k, n), replace=False)
    label[seeds] = np.arange(len(seeds))
    load = np.bincount(label[seeds], weights=w[seeds], minlength=k).astype(np.float64)
    for _ in range(4 * int(np.ceil(np.log2(max(n, 2)))) + 8):
        un = np.nonzero(label < 0)[0]
        if not len(un):
            break
        assigned = label >= 0
        H = sp.csr_matrix((np.ones(assigned.sum()), (np.nonzero(assigned)[0], label[assigned])), shape=(n, k))
        S = np.asarray((A[un] @ H).todense())
        S[:, load >= cap] = 0.0
        best = S.argmax(axis=1)
        conn = S[np.arange(len(un)), best]
        front = np.nonzero(conn > 0)[0]
        if not len(front):
            break
        order = front[np.argsort(-conn[front], kind="stable")]
        for p in range(k):
            mv = un[order[best[order] == p]]
            if not len(mv):
                continue
            take = mv[np.cumsum(w[mv]) <= (cap - load[p]) * 0.5 + w[mv].min()]   # grow in steps
            label[take] = p
            load[p] += w[take].sum()
    left = np.nonzero(label < 0)[0]         # disconnected leftovers: fill the parts up to the mean load
    if len(left):
        room = np.maximum((w.sum() / k) - load, 0.0)
        if room.sum() <= 0:
            room = np.ones(k)
        fill = np.cumsum(room) * (w[left].sum() / room.sum())
        label[left] = np.minimum(np.searchsorted(fill, np.cumsum(w[left]), side="left"), k - 1)
    return label


def _rebalance(A, w, label, k, cap, rng):
    """Force every part under `cap`: overweight parts give away the vertices that lose the least."""
    import scipy.sparse as sp
    n = A.shape[0]
    for _ in range(4 * k):
        load = np.bincount(label, weights=w, minlength=k)
        over = np.nonzero(load > cap)[0]
        if not len(over):
            break
        p = int(over[np.argmax(load[over])])
        H = sp.csr_matrix((np.ones(n), (np.arange(n), label)), shape=(n, k))
        mine = np.nonzero(label == p)[0]
        S = np.asarray((A[mine] @ H).todense())
        stay = S[:, p].copy()
        S[:, load >= cap] = -1.0
        S[:, p] = -1.0
        best = S.argmax(axis=1)
        loss = stay - S[np.arange(len(mine)), best]
        order = np.argsort(loss, kind="stable")
        excess = load[p] - cap
        cum = np.cumsum(w[mine[order]])
        n_move = int(np.searchsorted(cum, excess) + 1)
        mv, tg = mine[order[:n_move]], best[order[:n_move]]
        room = cap - load
        for q in range(k):
            sel = mv[tg == q]
            if len(sel) and q != p:
                sel = sel[np.cumsum(w[sel]) <= max(room[q], 0)]
                label[sel] = q
        if (np.bincount(label, weights=w, minlength=k)[p] >= load[p]):   # nothing fitted: spill to lightest
            q = int(np.argmin(load))
            label[mv[:max(1, n_move // 2)]] = q
    return label


def multilevel_owner(n_vars, factor_ptr, edge_var, world, seed=0, imbalance=0.03,
                     coarse_target=None, refine_iters=8) -> np.ndarray:
    """k-way multilevel partition of the variables (k = world).  Vertex weight = 1 + degree (a
    variable's own work plus its share of the factors), parts within (1 + imbalance) of the mean."""
    import scipy.sparse as sp
    rng = np.random.default_rng(seed)
    k = int(world)
    A0 = star_graph(n_vars, factor_ptr, edge_var)
    deg = np.bincount(np.asarray(edge_var, dtype=np.int64), minlength=n_vars).astype(np.float64)
    w0 = 1.0 + deg
    cap = (1.0 + imbalance) * w0.sum() / k
    target = coarse_target or max(200 * k, 4000)
    levels, graphs = [], [(A0, w0)]     # levels[i] maps graphs[i] -> graphs[i + 1]
    A, w = A0, w0
    while A.shape[0] > target:
        _, P = _match(A, w, max_w=cap / 20.0, rng=rng)
        if P.shape[1] > 0.92 * A.shape[0]:
            break
        A = (P.T @ A @ P).tocsr()
        A.setdiag(0)
        A.eliminate_zeros()
        w = np.asarray(P.T @ w).ravel()
        levels.append(P)
        graphs.append((A, w))
    label = _grow(A, w, k, cap, rng)
    label = _refine(A, w, label, k, cap, rng, iters=4 * refine_iters)
    for lvl in range(len(levels) - 1, -1, -1):
        label = label[levels[lvl].indices]      # one entry per row: the coarse id of every vertex
        Ai, wi = graphs[lvl]
        label = _refine(Ai, wi, label, k, cap, rng, iters=refine_iters)
    label = _rebalance(A0, w0, label, k, cap, rng)
    return label.astype(np.int32)


def partition_variables(n_vars, factor_ptr, edge_var, world, method="auto", seed=0,
                        imbalance=0.03) -> np.ndarray:
    """Owner rank of every variable.  method: 'blocks' | 'multilevel' | 'auto' (the one of the two
    that cuts fewer edges; blocks on ties, and always for world == 1 or tiny graphs)."""
    blocks = block_owner(n_vars, world)
    if method == "blocks" or world <= 1 or n_vars < 4 * world:
        return blocks
    ml = multilevel_owner(n_vars, factor_ptr, edge_var, world, seed=seed, imbalance=imbalance)
    if method == "multilevel":
        return ml
    if method != "auto":
        raise ValueError(f"unknown partition method {method!r}")
    return ml if edge_cut(ml, factor_ptr, edge_var) < edge_cut(blocks, factor_ptr, edge_var) else blocks
