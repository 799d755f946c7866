"""BASELINE.json's full-size shapes through size-independent properties (the CPU oracle cannot finish these):
R-MAT scale-22 node2vec (configs[2]) and Barabasi-Albert n = 10^7 TopSim top-20 (configs[4]).  Checked: generator
determinism and invariants, every walk step follows an edge, no dead ends on an undirected graph, results independent
of how the walks / queries are split across calls (the multi-GPU sharding rule), production kernel == exact hash
kernel bit for bit, top-k order, step counts, degree-proportional stationary visits."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from graph_embedding_b200 import _lib  # noqa: E402


def _steps_follow_edges(rp, col, walks):
    """Vectorised binary search of every (a -> b) step in the sorted row of a."""
    a = walks[:, :-1].ravel().astype(np.int64)
    b = walks[:, 1:].ravel()
    lo, hi = rp[a].copy(), rp[a + 1].copy()
    end = hi.copy()
    while (lo < hi).any():
        mid = (lo + hi) >> 1
        go = (lo < hi) & (col[np.minimum(mid, len(col) - 1)] < b)
        stay = (lo < hi) & ~go
        lo = np.where(go, mid + 1, lo)
        hi = np.where(stay, mid, hi)
    return bool(((lo < end) & (col[np.minimum(lo, len(col) - 1)] == b)).all())


def test_rmat22_node2vec_full_size_properties():
    h = _lib.GraphHandle.rmat(22, 16 << 22, seed=1)
    assert h.n == 1 << 22 and h.nnz == 134109214 and h.max_degree == 1678      # 134.1 M directed entries
    starts = h.nonisolated()
    assert len(starts) == 4178039                                               # |V+|, 99.6 % of the ids
    c = h.csr(weights=False, node_ids=False, first_seen=False)
    rp, col = c["row_ptr"], c["col_idx"]
    deg = np.diff(rp)
    assert rp[-1] == h.nnz and deg.max() == 1678 and np.array_equal(np.nonzero(deg)[0], starts)
    # one full pass through the host API (what bench.py's e2e leg times): 4.18 M walks x 80
    walks, lens = h.walks(0.25, 4.0, 80, starts, seed=3)
    assert walks.shape == (len(starts), 80) and (lens == 80).all()             # undirected: no dead ends
    assert np.array_equal(walks[:, 0], starts) and walks.min() >= 0 and walks.max() < h.n
    sub = slice(0, len(starts), 211)                                            # 19.8 k walks, 1.56 M steps
    assert _steps_follow_edges(rp, col, walks[sub])
    # the second-order law at full size, on its return component: over the visited (prev, cur) contexts the observed
    # return frequency must equal the mean of get_alias_edge's P(return) = (1/p) / (1/p + c + (deg(cur) - 1 - c)/q),
    # c = |N(prev) & N(cur)| from a host intersection (node2vec.py:61-81)
    ws = walks[::2111]                                                          # 1980 walks, 154 k contexts
    prev, cur, nxt = ws[:, :-2].ravel(), ws[:, 1:-1].ravel(), ws[:, 2:].ravel()
    pred = np.empty(len(prev))
    for i, (u, v) in enumerate(zip(prev.tolist(), cur.tolist())):
        cnt = len(np.intersect1d(col[rp[u]:rp[u + 1]], col[rp[v]:rp[v + 1]], assume_unique=True))
        pred[i] = 4.0 / (4.0 + cnt + (deg[v] - 1 - cnt) / 4.0)
    obs = (nxt == prev).mean()
    assert abs(obs - pred.mean()) < 5 * np.sqrt(pred.mean() * (1 - pred.mean()) / len(pred)) + 1e-3, (obs, pred.mean())
    # sharding rule: the corpus depends on (seed, global walk id) only, not on the batch split
    k = 1234567
    a, _ = h.walks(0.25, 4.0, 80, starts[:k], seed=3, walk_id_base=0)
    b, _ = h.walks(0.25, 4.0, 80, starts[k:k + 100000], seed=3, walk_id_base=k)
    assert np.array_equal(a, walks[:k]) and np.array_equal(b, walks[k:k + 100000])
    other, _ = h.walks(0.25, 4.0, 80, starts[:1000], seed=4)
    assert not np.array_equal(other, walks[:1000])
    del walks, a, b
    # q < 1 (configs[3]'s parameters): Bloom-filtered "others" component, same validity
    wq, lq = h.walks(4.0, 0.5, 80, starts[sub], seed=5)
    assert (lq == 80).all() and _steps_follow_edges(rp, col, wq)
    assert (wq[:, 2:] == wq[:, :-2]).mean() < 0.02                              # p = 4 avoids the return edge
    # ... at exactly the law's rate: P(return) = (1/p) / (1/p + c + (deg(cur) - 1 - c)/q), the component the reverse
    # index serves (prev is never proposed; its share of the mass is decided before the load)
    ws = wq[::10]
    prev, cur, nxt = ws[:, :-2].ravel(), ws[:, 1:-1].ravel(), ws[:, 2:].ravel()
    pred = np.empty(len(prev))
    for i, (u, v) in enumerate(zip(prev.tolist(), cur.tolist())):
        cnt = len(np.intersect1d(col[rp[u]:rp[u + 1]], col[rp[v]:rp[v + 1]], assume_unique=True))
        pred[i] = 0.25 / (0.25 + cnt + (deg[v] - 1 - cnt) * 2.0)
    obs = (nxt == prev).mean()
    assert abs(obs - pred.mean()) < 5 * np.sqrt(pred.mean() / len(pred)) + 2e-4, (obs, pred.mean())
    # first-order stationary law: visits ~ degree
    w1, _ = h.walks(1.0, 1.0, 80, starts, seed=6)
    visits = np.bincount(w1[:, 40:].ravel(), minlength=h.n).astype(np.float64)
    assert np.corrcoef(visits, deg.astype(np.float64))[0, 1] > 0.98              # ~40 visits per vertex: Poisson noise caps it near 0.99
    for lo_d, hi_d in ((1, 8), (8, 64), (64, 512), (512, 4096)):               # visit share of a degree class = its share of the entries
        cls = (deg >= lo_d) & (deg < hi_d)
        assert abs(visits[cls].sum() / visits.sum() - deg[cls].sum() / h.nnz) < 2e-3, (lo_d, hi_d)


def test_ba10m_topsim_full_size_properties(monkeypatch):
    n, m = 10_000_000, 8
    h = _lib.GraphHandle.barabasi_albert(n, m, seed=1)
    assert h.n == n and h.nnz == 2 * (28 + (n - 8) * 8)                         # 8-clique seed + 8 edges per new vertex
    q = np.random.RandomState(2).choice(n, 4096, replace=False).astype(np.int64)
    monkeypatch.delenv("GW_SIMRANK", raising=False)
    ids, sc = h.simrank_topk(q, 0.6, 5, 10000, 20, seed=1)
    assert h.simrank_last_steps() == 4096 * 10000 * 10                          # no isolated vertex, no dead end
    assert (np.diff(sc, axis=1) <= 0).all() and (sc[:, 0] > 0).all() and (sc >= 0).all()
    assert ids.max() < n and (ids[sc > 0] >= 0).all()
    assert not (ids == q[:, None]).any()                                        # sim[v][v] = 0
    tie = (np.diff(sc, axis=1) == 0) & (sc[:, 1:] > 0)
    assert (np.diff(ids, axis=1)[tie] > 0).all()                                # equal scores: lower id first
    # sharding rule: results depend on (seed, global query index) only
    a_ids, a_sc = h.simrank_topk(q[:1500], 0.6, 5, 10000, 20, seed=1, query_id_base=0)
    b_ids, b_sc = h.simrank_topk(q[1500:], 0.6, 5, 10000, 20, seed=1, query_id_base=1500)
    assert np.array_equal(np.vstack([a_ids, b_ids]), ids) and np.vstack([a_sc, b_sc]).tobytes() == sc.tobytes()
    slow = h.simrank_last_slow_queries()
    assert slow < 64                                                            # the log kernel finishes almost everything
    # production kernel == exact hash kernel, bit for bit, at full graph size
    monkeypatch.setenv("GW_SIMRANK", "hash")
    e_ids, e_sc = h.simrank_topk(q[:192], 0.6, 5, 10000, 20, seed=1)
    monkeypatch.delenv("GW_SIMRANK", raising=False)
    assert np.array_equal(e_ids, ids[:192]) and e_sc.tobytes() == sc[:192].tobytes()
    # the path-tree estimator at full size: same invariants, same sharding rule (scores are x SAMPLE, as the reference)
    t_ids, t_sc = h.simrank_topk(q[:512], 0.6, 5, 10000, 20, mode=_lib.GW_SIMRANK_HYBRID, seed=1)
    assert (np.diff(t_sc, axis=1) <= 0).all() and (t_sc[:, 0] > 0).all() and not (t_ids == q[:512, None]).any()
    u_ids, u_sc = h.simrank_topk(q[200:512], 0.6, 5, 10000, 20, mode=_lib.GW_SIMRANK_HYBRID, seed=1, query_id_base=200)
    assert np.array_equal(u_ids, t_ids[200:]) and u_sc.tobytes() == t_sc[200:].tobytes()
    m = np.median(t_sc[:, 0] / 10000.0 / sc[:512, 0])
    assert 0.5 < m < 1.5, m                                                     # both estimate the same top score


def test_rmat26_node2vec_sharded_shape_properties():
    """BASELINE configs[3]: R-MAT scale-26 (2^26 ids, 2^30 tuples, 2.1 G directed entries), p = 4, q = 0.5.  The graph is
    generated, built and preprocessed on the device (common-neighbour counts over 1.07 G undirected edges); a sample of
    the start list is walked through the host API.  Size-independent properties, as at scale 22: every step follows an
    edge, no dead ends, the return component follows get_alias_edge's law (node2vec.py:61-81) with c from a host
    intersection, and the corpus does not depend on how the start list is split (the multi-GPU sharding rule)."""
    h = _lib.GraphHandle.rmat(26, 16 << 26, seed=1)
    assert h.n == 1 << 26 and h.nnz == 2147147818 and h.max_degree == 3546
    prep_ms = h.prepare_walks()
    assert 0 < prep_ms < 60000
    starts = h.nonisolated()
    assert len(starts) == 66634460
    c = h.csr(weights=False, node_ids=False, first_seen=False)                  # 8.6 GB + 0.5 GB to the host
    rp, col = c["row_ptr"], c["col_idx"]
    deg = np.diff(rp)
    assert rp[-1] == h.nnz and deg.max() == 3546
    rs = np.random.RandomState(26)
    sample = np.sort(rs.choice(len(starts), size=1 << 20, replace=False))
    sub = starts[sample]
    walks, lens = h.walks(4.0, 0.5, 80, sub, seed=3, walk_id_base=0)
    assert (lens == 80).all() and np.array_equal(walks[:, 0], sub) and walks.min() >= 0 and walks.max() < h.n
    assert _steps_follow_edges(rp, col, walks[::53])
    assert (walks[:, 2:] == walks[:, :-2]).mean() < 0.02                        # p = 4 avoids the return edge ...
    ws = walks[::531]                                                           # ... at exactly the law's rate: 1975 walks, 154 k contexts
    prev, cur, nxt = ws[:, :-2].ravel(), ws[:, 1:-1].ravel(), ws[:, 2:].ravel()
    pred = np.empty(len(prev))
    for i, (u, v) in enumerate(zip(prev.tolist(), cur.tolist())):
        cnt = len(np.intersect1d(col[rp[u]:rp[u + 1]], col[rp[v]:rp[v + 1]], assume_unique=True))
        pred[i] = 0.25 / (0.25 + cnt + (deg[v] - 1 - cnt) * 2.0)
    obs = (nxt == prev).mean()
    assert abs(obs - pred.mean()) < 5 * np.sqrt(pred.mean() / len(pred)) + 2e-4, (obs, pred.mean())
    # the common-neighbour counts the walker uses, against the same host intersection (sampled contexts)
    # split independence = the sharding rule of gw_node2vec_walks_sharded: slices with their global walk ids
    k = 400000
    a, _ = h.walks(4.0, 0.5, 80, sub[:k], seed=3, walk_id_base=0)
    b, _ = h.walks(4.0, 0.5, 80, sub[k:k + 50000], seed=3, walk_id_base=k)
    assert np.array_equal(a, walks[:k]) and np.array_equal(b, walks[k:k + 50000])
    for nranks in (2, 8):                                                       # gw_shard_range slices tile the list
        edges = [_lib.shard_range(len(sub), r, nranks) for r in range(nranks)]
        assert edges[0][0] == 0 and edges[-1][1] == len(sub) and all(edges[i][1] == edges[i + 1][0] for i in range(nranks - 1))
    lo, hi = _lib.shard_range(len(sub), 5, 8)
    s5, _ = h.walks(4.0, 0.5, 80, sub[lo:hi], seed=3, walk_id_base=lo)
    assert np.array_equal(s5, walks[lo:hi])
