"""CPU oracle for the node2vec walk path.  TEST INFRASTRUCTURE ONLY.

This file restates, on flat CSR arrays, the algorithm of the reference's
``node2vec/src/node2vec.py`` (and the loader call in ``node2vec/src/main.py``).
It is the *checker* for the CUDA path: only ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s cpu_baseline / ``--impl reference`` leg may import it.  The
product package (``graph_embedding_b200``) never does.

Parity status: PINNED.  ``tests/golden/make_golden_node2vec.py`` imports the real
reference module from ``/root/reference/node2vec/src`` in the build container and
writes fixtures (CSR, alias tables, recorded uniforms, walks) under
``tests/golden/``; ``tests/test_oracle_node2vec.py`` checks every function below
against those fixtures bit for bit.

Arithmetic notes (all fp64, no FMA):
  * ``sum()`` over unnormalised probabilities is the NAIVE left-to-right sum of
    the reference's pinned interpreter era (cpython-35 .pyc files, numpy 1.11.2),
    not CPython >= 3.12's compensated sum.  The golden generator installs the same
    naive ``sum`` in the reference module's namespace, next to the ``np.int`` shim.
  * dense vertex index = rank of the original id in ascending order, so that
    ``col_idx`` rows (ascending) address the same neighbour as
    ``sorted(G.neighbors(cur))[k]`` (node2vec.py:25).
"""
from __future__ import annotations

import math
import numpy as np


# ----------------------------------------------------------------------------
# graph loading: nx.read_edgelist(..., create_using=nx.DiGraph()) + weight=1 +
# to_undirected()     (reference node2vec/src/main.py:76-89)
# ----------------------------------------------------------------------------
def parse_edgelist(path, delimiter=None, weighted=False):
    """Parse like networkx.read_edgelist/parse_edgelist (main.py:81,83):
    strip comments after '#', split on delimiter, skip lines with < 2 fields,
    nodetype=int.  Returns (src, dst, w) int64/int64/float64 arrays in file order."""
    src, dst, w = [], [], []
    with open(path, "r") as f:
        for line in f:
            p = line.find("#")
            if p >= 0:
                line = line[:p]
            if not len(line):
                continue
            s = line.strip().split(delimiter)
            if len(s) < 2:
                continue
            u = int(s[0])
            v = int(s[1])
            if weighted:
                # data=(('weight', float),) : exactly one data column is required
                if len(s) - 2 != 1:
                    raise IndexError("edge data %r and data_keys (weight,) are not the same length" % (s[2:],))
                w.append(float(s[2]))
            else:
                w.append(1.0)
            src.append(u)
            dst.append(v)
    return (np.asarray(src, dtype=np.int64), np.asarray(dst, dtype=np.int64),
            np.asarray(w, dtype=np.float64))


def build_simple_graph(src, dst, w, directed):
    """networkx semantics of read_graph (main.py:76-89).

    DiGraph: duplicate (u,v) lines collapse, last weight wins, adjacency position is
    the first occurrence.  Undirected: ``to_undirected()`` walks nodes in insertion
    order and each node's successors in insertion order and the LAST directed edge
    met in that order fixes the weight of the undirected pair.

    Returns dict(node_ids (ascending original ids), first_seen (dense indices in
    ``list(G.nodes())`` order), row_ptr int64, col_idx int32, weights float64)."""
    node_order = {}
    for u, v in zip(src.tolist(), dst.tolist()):
        if u not in node_order:
            node_order[u] = len(node_order)
        if v not in node_order:
            node_order[v] = len(node_order)
    # directed adjacency with insertion order
    dadj = {}
    for u, v, ww in zip(src.tolist(), dst.tolist(), w.tolist()):
        dadj.setdefault(u, {})[v] = ww      # dict keeps first-insert position, last value
    ids = np.array(sorted(node_order), dtype=np.int64)
    rank = {int(x): i for i, x in enumerate(ids.tolist())}
    adj = {u: {} for u in node_order}
    if directed:
        for u in node_order:
            for v, ww in dadj.get(u, {}).items():
                adj[u][v] = ww
    else:
        for u in node_order:                 # G.to_undirected(): later edge data wins
            for v, ww in dadj.get(u, {}).items():
                adj[u][v] = ww
                adj[v][u] = ww
    n = len(ids)
    row_ptr = np.zeros(n + 1, dtype=np.int64)
    cols, ws = [], []
    for i, u in enumerate(ids.tolist()):
        nb = sorted(adj[u])
        row_ptr[i + 1] = row_ptr[i] + len(nb)
        cols.extend(rank[x] for x in nb)
        ws.extend(adj[u][x] for x in nb)
    first_seen = np.array([rank[u] for u in node_order], dtype=np.int64)
    return dict(node_ids=ids, first_seen=first_seen, row_ptr=row_ptr,
                col_idx=np.asarray(cols, dtype=np.int32),
                weights=np.asarray(ws, dtype=np.float64))


def load_graph(path, delimiter=None, weighted=False, directed=False):
    s, d, w = parse_edgelist(path, delimiter, weighted)
    return build_simple_graph(s, d, w, directed)


# ----------------------------------------------------------------------------
# alias tables                                      (node2vec.py:116-147, 150-160)
# ----------------------------------------------------------------------------
def alias_setup(probs):
    """node2vec.py:116-147.  probs: sequence of python/np floats.  -> (J int64, q f64)."""
    K = len(probs)
    q = np.zeros(K, dtype=np.float64)
    J = np.zeros(K, dtype=np.int64)
    smaller, larger = [], []
    for kk in range(K):
        q[kk] = K * float(probs[kk])                    # :130
        if q[kk] < 1.0:
            smaller.append(kk)
        else:
            larger.append(kk)
    while smaller and larger:                           # :136
        small = smaller.pop()
        large = larger.pop()
        J[small] = large                                # :140
        q[large] = (q[large] + q[small]) - 1.0          # :141  (left-to-right)
        if q[large] < 1.0:
            smaller.append(large)
        else:
            larger.append(large)
    return J, q


def alias_draw(J, q, u1, u2):
    """node2vec.py:150-160 with the two np.random.rand() values passed in."""
    K = len(J)
    kk = int(math.floor(u1 * K))
    if u2 < q[kk]:
        return kk
    return int(J[kk])


def _naive_sum(xs):
    t = 0.0
    for x in xs:
        t = t + x
    return t


def node_probs(g, v):
    """normalized_probs of node2vec.py:93-96 for dense vertex v."""
    a, b = int(g["row_ptr"][v]), int(g["row_ptr"][v + 1])
    un = [float(x) for x in g["weights"][a:b]]
    norm = _naive_sum(un)
    return [u / norm for u in un]


def has_edge(g, a, b):
    """G.has_edge(a, b): b in out-neighbours of a (binary search in the sorted row)."""
    lo, hi = int(g["row_ptr"][a]), int(g["row_ptr"][a + 1])
    col = g["col_idx"]
    i = int(np.searchsorted(col[lo:hi], b))
    return i < hi - lo and int(col[lo + i]) == b


def edge_probs(g, src, dst, p, q):
    """normalized_probs of get_alias_edge(src, dst)  (node2vec.py:61-79)."""
    a, b = int(g["row_ptr"][dst]), int(g["row_ptr"][dst + 1])
    un = []
    for e in range(a, b):
        nbr = int(g["col_idx"][e])
        wt = float(g["weights"][e])
        if nbr == src:
            un.append(wt / p)                            # :71-72
        elif has_edge(g, nbr, src):                      # :73-74
            un.append(wt)
        else:
            un.append(wt / q)                            # :75-76
    norm = _naive_sum(un)                                # :78
    return [u / norm for u in un]                        # :79


def alias_nodes_flat(g):
    """alias_nodes of preprocess_transition_probs (node2vec.py:91-97), flattened with
    offsets = row_ptr.  -> (J int32[nnz], q f64[nnz])"""
    nnz = int(g["row_ptr"][-1])
    J = np.zeros(nnz, dtype=np.int32)
    q = np.zeros(nnz, dtype=np.float64)
    n = len(g["row_ptr"]) - 1
    for v in range(n):
        a, b = int(g["row_ptr"][v]), int(g["row_ptr"][v + 1])
        if b > a:
            j, qq = alias_setup(node_probs(g, v))
            J[a:b] = j
            q[a:b] = qq
    return J, q


def alias_edges_flat(g, p, q):
    """alias_edges (node2vec.py:99-108): one table of length deg(v) per directed CSR
    entry e=(u->v), in CSR order.  -> (off int64[nnz+1], J int32, q f64)."""
    row_ptr, col = g["row_ptr"], g["col_idx"]
    n = len(row_ptr) - 1
    nnz = int(row_ptr[-1])
    deg = np.diff(row_ptr)
    off = np.zeros(nnz + 1, dtype=np.int64)
    off[1:] = np.cumsum(deg[col])
    J = np.zeros(int(off[-1]), dtype=np.int32)
    Q = np.zeros(int(off[-1]), dtype=np.float64)
    for u in range(n):
        for e in range(int(row_ptr[u]), int(row_ptr[u + 1])):
            v = int(col[e])
            j, qq = alias_setup(edge_probs(g, u, v, p, q))
            J[off[e]:off[e + 1]] = j
            Q[off[e]:off[e + 1]] = qq
    return off, J, Q


# ----------------------------------------------------------------------------
# walks                                                   (node2vec.py:13-59)
# ----------------------------------------------------------------------------
def walk_replay(g, an, ae, walk_length, start, uniforms, pos):
    """node2vec_walk (node2vec.py:13-39) consuming recorded uniforms from ``pos``.
    Returns (walk as dense indices, new pos)."""
    row_ptr, col = g["row_ptr"], g["col_idx"]
    anJ, anq = an
    aoff, aJ, aq = ae
    walk = [int(start)]
    e_prev = -1
    while len(walk) < walk_length:
        cur = walk[-1]
        a, b = int(row_ptr[cur]), int(row_ptr[cur + 1])
        if b - a > 0:
            u1, u2 = uniforms[pos], uniforms[pos + 1]
            pos += 2
            if len(walk) == 1:
                k = alias_draw(anJ[a:b], anq[a:b], u1, u2)           # :28-29
            else:
                o0, o1 = int(aoff[e_prev]), int(aoff[e_prev + 1])
                k = alias_draw(aJ[o0:o1], aq[o0:o1], u1, u2)         # :32-34
            e_prev = a + k
            walk.append(int(col[a + k]))
        else:
            break                                                     # :36-37
    return walk, pos


def simulate_walks_replay(g, an, ae, walk_length, starts, uniforms):
    """simulate_walks (node2vec.py:41-59) with the shuffled start order and the
    np.random.rand() stream supplied.  Returns list of walks (dense indices)."""
    walks, pos = [], 0
    for s in starts:
        w, pos = walk_replay(g, an, ae, walk_length, int(s), uniforms, pos)
        walks.append(w)
    return walks, pos


def simulate_walks_cpu(g, p, q, walk_length, num_walks, seed=0):
    """Free-running CPU baseline with the reference's structure (per-step python loop,
    two numpy draws per step, materialised tables).  Used by bench.py's cpu_baseline
    leg on small graphs only.  Returns (walks, steps)."""
    import random
    rng = np.random.RandomState(seed)
    random.seed(seed)
    an = alias_nodes_flat(g)
    ae = alias_edges_flat(g, p, q)
    nodes = [int(x) for x in g["first_seen"]]
    walks, steps = [], 0
    for _ in range(num_walks):
        random.shuffle(nodes)
        for s in nodes:
            u = rng.rand(2 * (walk_length - 1))
            w, pos = walk_replay(g, an, ae, walk_length, s, u, 0)
            steps += len(w) - 1
            walks.append(w)
    return walks, steps


# exact transition laws for the chi-square tests --------------------------------
def first_step_law(g, cur):
    return np.asarray(node_probs(g, cur))


def second_order_law(g, prev, cur, p, q):
    return np.asarray(edge_probs(g, prev, cur, p, q))
