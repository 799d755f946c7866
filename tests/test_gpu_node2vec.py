"""GPU parity tests of the node2vec path, through the C ABI (ctypes) and the drop-in module.
Bit-exact: CSR, alias tables, replayed walks.  Statistical: free-running walks (chi-square
against the exact law from the oracle, alpha = 1e-4 on the pooled statistic)."""
import json
import os

import numpy as np
import pytest
from scipy import stats

from conftest import GOLDEN, DATA
from helpers import CASES, case, load_npz, data_path, chi2_transitions
from oracle import n2v_oracle as O

pytestmark = pytest.mark.gpu

from graph_embedding_b200 import _lib, node2vec as n2v  # noqa: E402


def open_graph(meta):
    return _lib.GraphHandle.from_file(data_path(meta), delimiter=meta["delimiter"], weighted=meta["weighted"],
                                      directed=meta["directed"])


@pytest.mark.parametrize("meta", CASES, ids=[c["name"] for c in CASES])
def test_csr_bit_exact(meta):
    z = load_npz(meta)
    h = open_graph(meta)
    c = h.csr(weights=True)
    for k in ("node_ids", "first_seen", "row_ptr", "col_idx"):
        assert np.array_equal(c[k], z[k]), k
    assert c["weights"].tobytes() == z["weights"].tobytes()
    assert h.n == meta["n_nodes"] and h.nnz == meta["nnz"]


def test_alias_setup_kats():
    for kat in json.load(open(os.path.join(GOLDEN, "alias_setup_kat.json"))):
        J, q = n2v.alias_setup(kat["probs"])
        assert [int(x) for x in J] == kat["J"]
        assert [float(x).hex() for x in q] == kat["q_hex"]
    rs = np.random.RandomState(3)
    for K in (2, 7, 33, 257, 4000):
        p = rs.rand(K); p /= p.sum()
        J, q = n2v.alias_setup(p)
        Jo, qo = O.alias_setup(list(p))
        assert np.array_equal(J, Jo) and q.tobytes() == qo.tobytes()


@pytest.mark.parametrize("meta", CASES, ids=[c["name"] for c in CASES])
def test_alias_tables_bit_exact(meta):
    z = load_npz(meta)
    h = open_graph(meta)
    J, q = h.alias_nodes()
    assert np.array_equal(J, z["an_J"])
    assert q.tobytes() == z["an_q"].tobytes()
    assert h.alias_edges_size() == meta["n_alias_edge_entries"]
    off, J, q = h.alias_edges(meta["p"], meta["q"])
    assert np.array_equal(off, z["ae_off"])
    assert np.array_equal(J, z["ae_J"])
    assert q.tobytes() == z["ae_q"].tobytes()


def test_alias_edges_large_tables_bit_exact():
    """Tables of every size class of the warp-per-table builder (<= 32 one thread per table; 256 / 1024 /
    4096 / 18000-entry shared-memory stages), weighted so that the fp64 operation order matters:
    sampled tables must equal get_alias_edge + alias_setup of the oracle bit for bit."""
    rs = np.random.RandomState(11)
    hubs = {0: 5000, 1: 1500, 2: 400, 3: 100}                 # hub id -> degree
    n = 6000
    src, dst = [], []
    for hub, d in hubs.items():
        leaves = rs.choice(np.arange(10, n), size=d, replace=False)
        src += [hub] * d
        dst += leaves.tolist()
    extra = rs.randint(10, n, size=(4000, 2))
    extra = extra[extra[:, 0] != extra[:, 1]]
    src += extra[:, 0].tolist() + [0, 0, 1]
    dst += extra[:, 1].tolist() + [1, 2, 2]                     # hub-hub edges: common neighbours exist
    src, dst = np.array(src, dtype=np.int64), np.array(dst, dtype=np.int64)
    key = np.minimum(src, dst) * n + np.maximum(src, dst)      # one line per undirected pair
    _, first = np.unique(key, return_index=True)
    src, dst = src[np.sort(first)], dst[np.sort(first)]
    w = rs.rand(len(src)) * 3.0 + 0.1
    src = np.concatenate([src, np.arange(n)[:-1]]); dst = np.concatenate([dst, np.arange(n)[1:]])   # a path keeps every id present
    w = np.concatenate([w, rs.rand(n - 1) + 0.5])
    key = np.minimum(src, dst) * n + np.maximum(src, dst)
    _, first = np.unique(key, return_index=True)
    keep = np.sort(first)
    src, dst, w = src[keep], dst[keep], w[keep]
    g = O.build_simple_graph(src, dst, w, directed=False)
    h = _lib.GraphHandle.from_edges(src, dst, w)
    c = h.csr(weights=True)
    assert np.array_equal(c["row_ptr"], g["row_ptr"]) and np.array_equal(c["col_idx"], g["col_idx"])
    assert c["weights"].tobytes() == g["weights"].tobytes()
    p, q = 0.7, 1.9
    off, J, Q = h.alias_edges(p, q, budget_bytes=8 << 30)
    deg = np.diff(g["row_ptr"])
    assert off[-1] == int((deg[g["col_idx"]].astype(np.int64)).sum())
    rows = np.repeat(np.arange(n), deg)
    checked = set()
    for target in (0, 1, 2, 3, int(g["col_idx"][g["row_ptr"][0]]), 4000):        # tables INTO the hubs and into small vertices
        ent = np.nonzero(g["col_idx"] == target)[0]
        for e in ent[[0, len(ent) // 2, -1]].tolist():
            u = int(rows[e])
            j, qq = O.alias_setup(O.edge_probs(g, u, target, p, q))
            assert np.array_equal(J[off[e]:off[e + 1]], j), (u, target)
            assert Q[off[e]:off[e + 1]].tobytes() == np.asarray(qq, dtype=np.float64).tobytes(), (u, target)
            checked.add(int(off[e + 1] - off[e]))
    assert max(checked) > 4096 and any(1024 < k <= 4096 for k in checked) and any(256 < k <= 1024 for k in checked)


@pytest.mark.parametrize("meta", CASES, ids=[c["name"] for c in CASES])
def test_replay_walks_bit_exact(meta):
    z = load_npz(meta)
    h = open_graph(meta)
    G = n2v.Graph(n2v.EdgeListGraph(h), meta["directed"], meta["p"], meta["q"])
    G.preprocess_transition_probs()
    ids = z["node_ids"]
    walks = G.simulate_walks_replay(meta["walk_length"], ids[z["starts"]], z["uniforms"])
    assert len(walks) == meta["n_walks"]
    for i, w in enumerate(walks):
        assert w == ids[z["walks"][i, :z["lens"][i]]].tolist()
    # dict-like table access of the reference (node2vec.py:110-111)
    u = int(ids[0]); v = int(ids[z["col_idx"][0]])
    J, q = G.alias_edges[(u, v)]
    assert np.array_equal(J, z["ae_J"][z["ae_off"][0]:z["ae_off"][1]])
    with pytest.raises(KeyError):
        G.alias_edges[(u, 10 ** 9)]


def test_alias_edges_too_large_is_reported():
    h = open_graph(case("g333_p025_q4"))
    with pytest.raises(MemoryError):
        h.alias_edges(0.25, 4.0, budget_bytes=1000)


@pytest.mark.parametrize("walker", ["mixture", "mixture-noridx", "rejection"])
@pytest.mark.parametrize("name,reps,pq", [("karate_p025_q4", 3000, None), ("karate_p3_q07", 3000, None),
                                          ("karate_p1_q1", 1500, None), ("karate_p1_q1", 3000, (4.0, 0.5)),
                                          ("karate_p1_q1", 3000, (0.5, 0.5)), ("karate_p1_q1", 3000, (2.0, 2.0)),
                                          ("wdir_p05_q2", 3000, None), ("wund_p2_q05", 3000, None),
                                          ("moreno_p025_q4", 40, None), ("g333_p025_q4", 100, None),
                                          ("g333_p025_q4", 100, (4.0, 0.5))])
def test_free_running_chi_square(name, reps, pq, walker, monkeypatch):
    """>= 1e6 GPU steps per case; H0: transitions follow get_alias_edge's law (node2vec.py:61-81).
    Both device samplers are tested: the common-neighbour mixture walker (walk_cn.cu, the default
    on undirected unweighted graphs) and the rejection walker (walk.cu, everything else)."""
    meta = dict(case(name))
    if pq is not None:
        meta["p"], meta["q"] = pq
    if walker == "rejection":
        if meta["weighted"] or meta["directed"]:
            pytest.skip("these graphs always use the rejection walker")
        monkeypatch.setenv("GW_WALKER", "rejection")
    if walker == "mixture-noridx":                               # the instantiation graphs with a degree >= 65536 get:
        if meta["weighted"] or meta["directed"]:                 # plain counts in nbr4[].y, proposals of prev rejected
            pytest.skip("these graphs always use the rejection walker")
        monkeypatch.setenv("GW_CN_RIDX", "0")
    h = open_graph(meta)
    g = O.load_graph(data_path(meta), meta["delimiter"], meta["weighted"], meta["directed"])
    L = 40
    starts = np.tile(np.arange(h.n, dtype=np.int64), reps)
    walks, lens = h.walks(meta["p"], meta["q"], L, starts, seed=1234)
    assert (walks[:, 0] == starts).all()
    chi2, df = chi2_transitions(walks, lens, g, lambda c: O.first_step_law(g, c),
                                lambda a, b: O.second_order_law(g, a, b, meta["p"], meta["q"]))
    pval = stats.chi2.sf(chi2, df)
    assert df > 50
    assert pval > 1e-4, (chi2, df, pval)
    # a deliberately wrong law must be rejected by the same statistic (power check)
    chi2w, dfw = chi2_transitions(walks, lens, g, lambda c: O.first_step_law(g, c),
                                  lambda a, b: O.second_order_law(g, a, b, meta["p"] * 1.5, meta["q"] / 1.5))
    if not (meta["p"] == 1.0 and meta["q"] == 1.0):
        assert stats.chi2.sf(chi2w, dfw) < 1e-6


def test_hub_pairs_chi_square(monkeypatch):
    """Heavy-tailed shape: hubs that share most of their (long) rows.  The common-neighbour draw of such
    pairs goes through per-lane rejection instead of the warp-cooperative intersection; the transition
    frequencies must still follow get_alias_edge's law (pooled chi-square, alpha = 1e-4)."""
    monkeypatch.setenv("GW_CN_HUB", "1")          # the kernel instantiation graphs with rows > 2048 entries get
    n_leaf = 330                                  # hub rows ~285 entries: both rows > 256 -> whole-step rejection (hub-hub), shorter -> mixture
    hubs = [0, 1, 2]
    src, dst = [], []
    for h_ in hubs:
        for leaf in range(3, 3 + n_leaf):
            if (leaf + h_) % 7:                                   # rows differ a little: common counts < degree
                src.append(h_); dst.append(leaf)
    src += [0, 0, 1]; dst += [1, 2, 2]                            # hub-hub edges
    for leaf in range(3, 3 + n_leaf - 1, 2):
        src.append(leaf); dst.append(leaf + 1)                    # some leaf-leaf edges
    src, dst = np.array(src, dtype=np.int64), np.array(dst, dtype=np.int64)
    g = O.build_simple_graph(src, dst, np.ones(len(src)), directed=False)
    h = _lib.GraphHandle.from_edges(src, dst)
    deg = np.diff(g["row_ptr"])
    assert deg[:3].min() > 256                                    # long enough for both rejection paths
    for p, q in ((0.25, 4.0), (2.0, 3.0)):
        starts = np.tile(np.arange(h.n, dtype=np.int64), 400)
        walks, lens = h.walks(p, q, 30, starts, seed=77)
        chi2, df = chi2_transitions(walks, lens, g, lambda c: O.first_step_law(g, c),
                                    lambda a, b: O.second_order_law(g, a, b, p, q))
        pval = stats.chi2.sf(chi2, df)
        assert df > 300 and pval > 1e-4, (p, q, chi2, df, pval)
        bad, dfb = chi2_transitions(walks, lens, g, lambda c: O.first_step_law(g, c),
                                    lambda a, b: O.second_order_law(g, a, b, p * 1.5, q / 1.5))
        assert stats.chi2.sf(bad, dfb) < 1e-6                     # power: a nearby law is rejected


@pytest.mark.parametrize("walker", ["mixture", "rejection"])
def test_walks_independent_of_batch_split(walker, monkeypatch):
    if walker == "rejection":
        monkeypatch.setenv("GW_WALKER", "rejection")
    meta = case("karate_p025_q4")
    h = open_graph(meta)
    starts = np.tile(np.arange(h.n, dtype=np.int64), 20)
    whole, _ = h.walks(0.25, 4.0, 30, starts, seed=7, walk_id_base=100)
    a, _ = h.walks(0.25, 4.0, 30, starts[:123], seed=7, walk_id_base=100)
    b, _ = h.walks(0.25, 4.0, 30, starts[123:], seed=7, walk_id_base=223)
    assert np.array_equal(whole, np.concatenate([a, b]))
    other, _ = h.walks(0.25, 4.0, 30, starts, seed=8, walk_id_base=100)
    assert not np.array_equal(whole, other)


def test_large_edgelist_is_parsed_in_parallel_chunks(tmp_path):
    """Files above 1 MB are cut at line boundaries and parsed by several host threads: same CSR and same
    first-appearance order as the edge arrays, comments / blank lines / CRLF handled per line, and an error
    deep in the file is reported with the file's line number."""
    rs = np.random.RandomState(5)
    m = 400000
    src = rs.randint(0, 50000, m); dst = rs.randint(0, 50000, m)
    keep = src != dst
    src, dst = src[keep].astype(np.int64), dst[keep].astype(np.int64)
    lines = []
    for i, (u, v) in enumerate(zip(src.tolist(), dst.tolist())):
        if i % 1000 == 0:
            lines.append("# comment %d" % i)
        if i % 1777 == 0:
            lines.append("")
        lines.append("%d %d%s" % (u, v, "\r" if i % 3 == 0 else ""))
    path = str(tmp_path / "big.edgelist")
    with open(path, "w", newline="") as f:
        f.write("\n".join(lines) + "\n")
    assert os.path.getsize(path) > 3 << 20
    h = _lib.GraphHandle.from_file(path, delimiter=" ")
    g = _lib.GraphHandle.from_edges(src, dst)
    a, b = h.csr(), g.csr()
    for k in ("node_ids", "first_seen", "row_ptr", "col_idx"):
        assert np.array_equal(a[k], b[k]), k
    bad_line = len(lines) - 5
    lines[bad_line - 1] = "12 notanumber"
    with open(path, "w", newline="") as f:
        f.write("\n".join(lines) + "\n")
    with pytest.raises(IOError) as ei:
        _lib.GraphHandle.from_file(path, delimiter=" ")
    assert ":%d:" % bad_line in str(ei.value)


def test_bloom_filter_does_not_change_walks(monkeypatch):
    """q < 1: the edge Bloom filter only short-cuts adjacency tests whose answer is "not adjacent";
    positives are verified exactly, so the walks are identical with the filter switched off."""
    monkeypatch.setenv("GW_BLOOM", "0")
    h0 = _lib.GraphHandle.rmat(14, 16 << 14, seed=3)
    starts = h0.nonisolated()
    w0, _ = h0.walks(4.0, 0.5, 40, starts, seed=5)
    import torch
    d_starts = torch.from_numpy(starts).cuda()
    t0 = h0.walk_traffic_dev(4.0, 0.5, 40, d_starts.data_ptr(), len(starts), seed=5)
    monkeypatch.delenv("GW_BLOOM")
    h1 = _lib.GraphHandle.rmat(14, 16 << 14, seed=3)
    w1, _ = h1.walks(4.0, 0.5, 40, starts, seed=5)
    assert np.array_equal(w0, w1)
    assert len(np.unique(w1[:, 1:])) > 1000
    t1 = h1.walk_traffic_dev(4.0, 0.5, 40, d_starts.data_ptr(), len(starts), seed=5)
    assert t1["steps"] == t0["steps"] and t1["random_accesses"] < t0["random_accesses"]   # the byte model sees the saving


def test_dead_ends_and_errors():
    meta = case("wdir_p05_q2")
    h = open_graph(meta)
    c = h.csr()
    deg = np.diff(c["row_ptr"])
    sink = int(np.nonzero(deg == 0)[0][0])
    w, l = h.walks(0.5, 2.0, 10, [sink], seed=1)
    assert l[0] == 1 and w[0, 0] == sink and (w[0, 1:] == -1).all()
    with pytest.raises(KeyError):
        h.walks(0.5, 2.0, 10, [h.n + 5])
    with pytest.raises(ValueError):
        h.walks(0.0, 2.0, 10, [0])
    w, l = h.walks(0.5, 2.0, 10, np.zeros(0, dtype=np.int64))
    assert w.shape == (0, 10)
    # empty graph and missing file
    e = _lib.GraphHandle.from_edges([], [])
    assert e.n == 0 and e.nnz == 0
    with pytest.raises(IOError):
        _lib.GraphHandle.from_file("/nonexistent/file.txt")


def test_drop_in_module_and_cli(tmp_path):
    """The reference's call sequence (node2vec/src/main.py:104-112) on the karate anchor config."""
    import random
    from graph_embedding_b200 import main as cli
    out = str(tmp_path / "walks.txt")
    args = cli.parse_args(["--input", os.path.join(DATA, "karate.edgelist"), "--delimiter", " ", "--output", out, "--emit", "walks",
                           "--p", "1", "--q", "1", "--walk-length", "80", "--num-walks", "10"])
    random.seed(0); np.random.seed(0)
    walks = cli.main(args)
    assert len(walks) == 340 and all(len(w) == 80 for w in walks)
    # iteration structure of simulate_walks: every iteration starts once from every node
    for it in range(10):
        assert sorted(w[0] for w in walks[it * 34:(it + 1) * 34]) == list(range(1, 35))
    edges = set()
    for line in open(os.path.join(DATA, "karate.edgelist")):
        a, b = map(int, line.split()); edges.add((a, b)); edges.add((b, a))
    assert all((a, b) in edges for w in walks for a, b in zip(w, w[1:]))
    assert cli.read_list(out)[0] == [str(x) for x in walks[0]]
    random.seed(0); np.random.seed(0)
    assert cli.main(args) == walks             # the global-RNG seed interface reproduces the run
    # networkx entry (what the reference passes)
    nx = pytest.importorskip("networkx")
    Gx = nx.read_edgelist(os.path.join(DATA, "karate.edgelist"), nodetype=int, create_using=nx.DiGraph())
    for e in Gx.edges():
        Gx[e[0]][e[1]]['weight'] = 1
    Gx = Gx.to_undirected()
    G = n2v.Graph(Gx, False, 0.25, 4.0)
    G.preprocess_transition_probs()
    z = load_npz(case("karate_p025_q4"))
    J, q = G.alias_nodes[1]
    assert np.array_equal(J, z["an_J"][:len(J)]) and q.tobytes() == z["an_q"][:len(q)].tobytes()
    w = G.node2vec_walk(80, 12)
    assert w[0] == 12 and len(w) == 80


def test_rmat_graph_and_walk_validity():
    h = _lib.GraphHandle.rmat(14, 16 << 14, seed=1)
    c = h.csr()
    rp, col = c["row_ptr"], c["col_idx"]
    n = h.n
    assert n == 1 << 14 and rp[-1] == h.nnz == len(col)
    rows = np.repeat(np.arange(n), np.diff(rp))
    assert (rows != col).all()                                   # no self loops
    key = rows.astype(np.int64) * n + col
    assert (np.diff(key) > 0).all()                              # sorted rows, no duplicates
    assert np.array_equal(np.sort(col.astype(np.int64) * n + rows), key)   # symmetric
    assert h.max_degree == np.diff(rp).max()
    starts = h.nonisolated()
    assert np.array_equal(starts, np.nonzero(np.diff(rp) > 0)[0])
    h2 = _lib.GraphHandle.rmat(14, 16 << 14, seed=1)
    assert np.array_equal(h2.csr()["col_idx"], col)              # deterministic in the seed
    for L in (80, 37):                                           # vector-store path and scalar path
        walks, lens = h.walks(0.25, 4.0, L, starts, seed=3)
        assert (lens == L).all()
        a, b = walks[:, :-1].astype(np.int64), walks[:, 1:].astype(np.int64)
        assert np.isin((a * n + b).ravel(), key).all()           # every step follows an edge
    # common-neighbour counts against a host intersection on sampled edges
    assert h.prepare_walks() >= 0.0
    walks_q, _ = h.walks(4.0, 0.5, 80, starts, seed=4)           # q < 1: "others" component
    a, b = walks_q[:, :-1].astype(np.int64), walks_q[:, 1:].astype(np.int64)
    assert np.isin((a * n + b).ravel(), key).all()
    # size-independent property: stationary first-order walks visit vertices ~ degree
    w1, _ = h.walks(1.0, 1.0, 40, np.tile(starts, 4), seed=5)
    visits = np.bincount(w1[:, 20:].ravel(), minlength=n).astype(np.float64)
    deg = np.diff(rp).astype(np.float64)
    assert np.corrcoef(visits, deg)[0, 1] > 0.98
