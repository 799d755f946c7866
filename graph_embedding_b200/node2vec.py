"""Drop-in for the reference module ``node2vec/src/node2vec.py`` backed by libgraphwalk (B200).

Same names, argument meaning and error behaviour as the reference:

    G = node2vec.Graph(nx_G, is_directed, p, q)         # node2vec.py:7-11
    G.preprocess_transition_probs()                      # :83-113  -> G.alias_nodes, G.alias_edges
    walks = G.simulate_walks(num_walks, walk_length)     # :41-59   -> list of lists of node ids
    walk  = G.node2vec_walk(walk_length, start_node)     # :13-39
    J, q  = alias_setup(probs); k = alias_draw(J, q)     # :116-160

``nx_G`` may be a networkx (Di)Graph with 'weight' edge attributes — what ``read_graph``
(node2vec/src/main.py:76-89) returns — or the light ``EdgeListGraph`` returned by this
package's ``main.read_graph`` (no networkx needed, graph built on the device).

Differences that are inherent to the device path and documented in DESIGN.md:
  * alias_edges needs sum(deg^2) entries; it is materialised (bit-exact) only when it fits the
    budget, and is NOT needed by ``simulate_walks``: the free-running walker samples the same
    second-order law on the fly (statistically identical, not the same RNG stream).
  * the global-RNG "seed interface" of the reference is kept: start order comes from
    ``random.shuffle`` and the Philox seed is drawn from ``np.random`` (so ``random.seed`` /
    ``np.random.seed`` make a run reproducible); ``simulate_walks_replay`` consumes a recorded
    ``np.random.rand`` stream and reproduces the reference's walks bit for bit.
"""
import random

import numpy as np

from . import _lib


class EdgeListGraph:
    """What this package's read_graph returns: a device graph + the reference's node order."""

    def __init__(self, handle):
        self.handle = handle
        c = handle.csr(weights=False)
        self.node_ids = c["node_ids"]            # dense index -> original id (ascending)
        self.first_seen = c["first_seen"]        # dense indices in list(G.nodes()) order
        self._csr = c

    def nodes(self):
        return self.node_ids[self.first_seen].tolist()

    def number_of_nodes(self):
        return self.handle.n

    def number_of_edges(self):
        loops = 0
        if not (self.handle.flags & _lib.GW_F_DIRECTED):
            rp, col = self._csr["row_ptr"], self._csr["col_idx"]
            rows = np.repeat(np.arange(self.handle.n), np.diff(rp))
            loops = int((rows == col).sum())
            return (self.handle.nnz + loops) // 2
        return self.handle.nnz


def _from_networkx(nx_G, is_directed):
    nodes = list(nx_G.nodes())
    ids = np.array(sorted(nodes), dtype=np.int64)
    rank = {int(x): i for i, x in enumerate(ids.tolist())}
    src, dst, w = [], [], []
    weighted = False
    for u, v, d in nx_G.edges(data=True):
        ww = d.get("weight", 1)
        weighted |= (ww != 1)
        src.append(rank[u]); dst.append(rank[v]); w.append(float(ww))
    directed = bool(nx_G.is_directed())
    h = _lib.GraphHandle.from_edges(src, dst, w if weighted else None, directed=directed,
                                    mode=_lib.GW_MODE_SIMPLE, n_slots=max(len(ids), 1))
    g = EdgeListGraph.__new__(EdgeListGraph)
    g.handle = h
    g.node_ids = ids
    g.first_seen = np.array([rank[x] for x in nodes], dtype=np.int64)
    g._csr = None
    return g


class _AliasNodes:
    """Mapping node -> (J, q) over the flat device-built tables (node2vec.py:110)."""

    def __init__(self, owner, J, q, row_ptr):
        self._o, self._J, self._q, self._rp = owner, J, q, row_ptr

    def __getitem__(self, node):
        i = self._o._dense(node)
        a, b = self._rp[i], self._rp[i + 1]
        return self._J[a:b].astype(np.int64), self._q[a:b]

    def __len__(self):
        return len(self._rp) - 1

    def __contains__(self, node):
        return node in self._o._rank


class _AliasEdges:
    """Mapping (u, v) -> (J, q) (node2vec.py:111); KeyError for a non-edge, as a dict would."""

    def __init__(self, owner, off, J, q, row_ptr, col):
        self._o, self._off, self._J, self._q, self._rp, self._col = owner, off, J, q, row_ptr, col

    def _entry(self, edge):
        u, v = self._o._dense(edge[0]), self._o._dense(edge[1])
        a, b = self._rp[u], self._rp[u + 1]
        k = int(np.searchsorted(self._col[a:b], v))
        if k >= b - a or self._col[a + k] != v:
            raise KeyError(edge)
        return a + k

    def __getitem__(self, edge):
        e = self._entry(edge)
        a, b = self._off[e], self._off[e + 1]
        return self._J[a:b].astype(np.int64), self._q[a:b]

    def __contains__(self, edge):
        try:
            self._entry(edge)
            return True
        except KeyError:
            return False

    def __len__(self):
        return len(self._off) - 1


class Graph():
    def __init__(self, nx_G, is_directed, p, q):
        self.G = nx_G
        self.is_directed = is_directed
        self.p = p
        self.q = q
        self._g = nx_G if isinstance(nx_G, EdgeListGraph) else _from_networkx(nx_G, is_directed)
        self._h = self._g.handle
        self._rank = None
        self._tables_for = None

    # ---- id plumbing ----
    def _dense(self, node):
        ids = self._g.node_ids
        i = int(np.searchsorted(ids, node))
        if i >= len(ids) or ids[i] != node:
            raise KeyError(node)
        return i

    def _dense_many(self, nodes):
        ids = self._g.node_ids
        nodes = np.asarray(nodes, dtype=np.int64)
        i = np.searchsorted(ids, nodes)
        bad = (i >= len(ids)) | (ids[np.minimum(i, len(ids) - 1)] != nodes)
        if bad.any():
            raise KeyError(int(nodes[bad][0]))
        return i.astype(np.int64)

    def _to_lists(self, walks, lens):
        ids = self._g.node_ids
        out = []
        for w, l in zip(walks, lens):
            out.append(ids[w[:l]].tolist())
        return out

    # ---- reference API ----
    def preprocess_transition_probs(self, materialize_edges=True, budget_bytes=0):
        """node2vec.py:83-113.  Builds alias_nodes (always) and alias_edges (when it fits) on the
        device, bit-exact with the reference tables."""
        c = self._h.csr(weights=False, node_ids=False, first_seen=False)
        J, q = self._h.alias_nodes()
        self.alias_nodes = _AliasNodes(self, J, q, c["row_ptr"])
        self.alias_edges = None
        if materialize_edges:
            try:
                off, eJ, eq = self._h.alias_edges(self.p, self.q, budget_bytes)
                self.alias_edges = _AliasEdges(self, off, eJ, eq, c["row_ptr"], c["col_idx"])
                self._tables_for = (self.p, self.q)
            except MemoryError:
                self.alias_edges = None      # too large: walks still run (on-the-fly bias)
        return

    def node2vec_walk(self, walk_length, start_node):
        """node2vec.py:13-39: one walk from start_node (original id)."""
        seed = int(np.random.randint(0, 2 ** 31 - 1)) | (int(np.random.randint(0, 2 ** 31 - 1)) << 31)
        w, l = self._h.walks(self.p, self.q, walk_length, [self._dense(start_node)], seed=seed)
        return self._to_lists(w, l)[0]

    def simulate_walks(self, num_walks, walk_length, as_array=False):
        """node2vec.py:41-59: num_walks passes over the cumulatively shuffled node list."""
        # the cumulatively shuffled node list, as dense indices: shuffling indices instead of ids is the same permutation,
        # and gw_py_random_shuffle is random.shuffle itself (same draws from Python's global generator, same state left)
        order = self._dense_many(self._g.nodes())
        print('Walk iteration:')
        seed = int(np.random.randint(0, 2 ** 31 - 1)) | (int(np.random.randint(0, 2 ** 31 - 1)) << 31)
        starts = []
        for walk_iter in range(num_walks):
            print(str(walk_iter + 1), '/', str(num_walks))
            _lib.py_random_shuffle(order)
            starts.append(order.copy())
        starts = np.concatenate(starts) if starts else np.zeros(0, dtype=np.int64)
        w, l = self._h.walks(self.p, self.q, walk_length, starts, seed=seed)
        if as_array:
            ids = self._g.node_ids
            if len(ids) and ids[0] == 0 and ids[-1] == len(ids) - 1:        # ids are the dense indices (generated graphs): nothing to map
                return w, l
            return np.where(w >= 0, ids[np.maximum(w, 0)], -1), l
        return self._to_lists(w, l)

    def simulate_walks_replay(self, walk_length, start_nodes, uniforms):
        """Replays recorded reference randomness: start_nodes = the post-shuffle node order of
        every iteration (original ids), uniforms = every np.random.rand() value in order."""
        if self._tables_for != (self.p, self.q):
            self.preprocess_transition_probs()
        if self.alias_edges is None:
            raise MemoryError("replay needs materialised alias_edges")
        starts = self._dense_many(start_nodes)
        c = self._h.csr(weights=False, node_ids=False, first_seen=False)
        deg = np.diff(c["row_ptr"])
        draw_offset = None
        if (deg == 0).any():     # ragged walks: offsets cannot be closed-form; take them from a dry pass
            draw_offset = _draw_offsets(c["row_ptr"], c["col_idx"], self, starts, uniforms, walk_length)
        w, l = self._h.walks_replay(walk_length, starts, uniforms, draw_offset)
        return self._to_lists(w, l)


def _draw_offsets(row_ptr, col, G, starts, uniforms, walk_length):
    """Walk i consumes 2*(len_i - 1) draws; lengths depend on the draws themselves only through
    dead ends, so a sequential host pass over the CSR + tables yields the offsets (directed graphs)."""
    anJ, anq = G.alias_nodes._J, G.alias_nodes._q
    ae = G.alias_edges
    off = np.zeros(len(starts) + 1, dtype=np.int64)
    pos = 0
    for i, s in enumerate(starts.tolist()):
        cur, e_prev, ln = s, -1, 1
        while ln < walk_length:
            a, b = row_ptr[cur], row_ptr[cur + 1]
            K = b - a
            if K == 0:
                break
            u1, u2 = uniforms[pos], uniforms[pos + 1]
            pos += 2
            if ln == 1:
                J, q = anJ[a:b], anq[a:b]
            else:
                J, q = ae._J[ae._off[e_prev]:ae._off[e_prev + 1]], ae._q[ae._off[e_prev]:ae._off[e_prev + 1]]
            kk = int(np.floor(u1 * K))
            k = kk if u2 < q[kk] else int(J[kk])
            e_prev = a + k
            cur = int(col[e_prev])
            ln += 1
        off[i + 1] = pos
    return off


def alias_setup(probs):
    """node2vec.py:116-147 on the device (bit-exact J, q)."""
    J, q = _lib.alias_setup(np.asarray(list(probs), dtype=np.float64))
    return J.astype(np.int64), q


def alias_draw(J, q):
    """node2vec.py:150-160 (host; two np.random.rand() draws, always)."""
    K = len(J)
    kk = int(np.floor(np.random.rand() * K))
    if np.random.rand() < q[kk]:
        return kk
    else:
        return J[kk]
