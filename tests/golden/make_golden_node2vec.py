#!/usr/bin/env python
"""Generate node2vec golden fixtures by running the UNMODIFIED reference module.

Run in the build container only (needs /root/reference and networkx):

    python tests/golden/make_golden_node2vec.py

Imports ``/root/reference/node2vec/src/node2vec.py`` with two era shims and nothing else:
  * ``np.int = int``       (node2vec.py:125 uses the alias numpy removed in 1.24)
  * ``node2vec.sum = naive left-to-right sum`` (CPython >= 3.12 made the builtin
    compensated; the reference's pinned stack, cpython-35 / numpy 1.11.2, is naive)
and records, per case, into ``tests/golden/n2v_<case>.npz``:
  CSR (node_ids, first_seen, row_ptr, col_idx, weights), alias_nodes / alias_edges
  flattened in CSR order, the post-shuffle start order of every walk iteration, every
  ``np.random.rand()`` value consumed, and the walks (ragged, -1 padded).
Data files are copied verbatim into tests/golden/data/ (they are inputs, not source).
"""
import functools
import gzip
import hashlib
import json
import operator
import os
import random
import shutil
import sys

import numpy as np
import networkx as nx

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
DATA = os.path.join(HERE, "data")

np.int = int                                            # shim 1
sys.path.insert(0, os.path.join(REF, "node2vec", "src"))
import node2vec as ref                                  # noqa: E402  (the reference itself)

ref.sum = lambda xs: functools.reduce(operator.add, xs, 0)   # shim 2 (naive sum)
ref.print = lambda *a, **k: None                        # silence progress prints


def read_graph(path, delimiter, weighted, directed):
    """Verbatim call sequence of node2vec/src/main.py:76-89 (main.py itself imports gensim)."""
    if weighted:
        G = nx.read_edgelist(path, nodetype=int, data=(('weight', float),),
                             create_using=nx.DiGraph(), delimiter=delimiter)
    else:
        G = nx.read_edgelist(path, nodetype=int, create_using=nx.DiGraph(), delimiter=delimiter)
        for edge in G.edges():
            G[edge[0]][edge[1]]['weight'] = 1
    if not directed:
        G = G.to_undirected()
    return G


def flatten(G, an, ae):
    ids = np.array(sorted(G.nodes()), dtype=np.int64)
    rank = {int(x): i for i, x in enumerate(ids)}
    first_seen = np.array([rank[u] for u in G.nodes()], dtype=np.int64)
    row_ptr = [0]
    col, wts = [], []
    for u in ids.tolist():
        nb = sorted(G.neighbors(u))
        col += [rank[x] for x in nb]
        wts += [float(G[u][x]['weight']) for x in nb]
        row_ptr.append(len(col))
    row_ptr = np.array(row_ptr, dtype=np.int64)
    col = np.array(col, dtype=np.int32)
    nnz = len(col)
    anJ = np.zeros(nnz, dtype=np.int32)
    anq = np.zeros(nnz, dtype=np.float64)
    for i, u in enumerate(ids.tolist()):
        J, q = an[u]
        anJ[row_ptr[i]:row_ptr[i + 1]] = J
        anq[row_ptr[i]:row_ptr[i + 1]] = q
    off = [0]
    aeJ, aeq = [], []
    n_tables = 0
    for i, u in enumerate(ids.tolist()):
        for e in range(row_ptr[i], row_ptr[i + 1]):
            v = int(ids[col[e]])
            J, q = ae[(u, v)]
            n_tables += 1
            aeJ += [int(x) for x in J]
            aeq += [float(x) for x in q]
            off.append(len(aeJ))
    assert n_tables == len(ae), (n_tables, len(ae))
    return dict(node_ids=ids, first_seen=first_seen, row_ptr=row_ptr, col_idx=col,
                weights=np.array(wts, dtype=np.float64), an_J=anJ, an_q=anq,
                ae_off=np.array(off, dtype=np.int64), ae_J=np.array(aeJ, dtype=np.int32),
                ae_q=np.array(aeq, dtype=np.float64)), rank


def run_case(name, path, delimiter, weighted, directed, p, q, walk_length, num_walks, seed,
             store_uniforms=True):
    G = read_graph(path, delimiter, weighted, directed)
    g = ref.Graph(G, directed, p, q)
    g.preprocess_transition_probs()
    flat, rank = flatten(G, g.alias_nodes, g.alias_edges)

    draws, orders = [], []
    real_rand, real_shuffle = np.random.rand, random.shuffle

    def rec_rand(*a):
        v = real_rand(*a)
        draws.append(float(v))
        return v

    def rec_shuffle(lst):
        real_shuffle(lst)
        orders.append([rank[x] for x in lst])

    random.seed(seed)
    np.random.seed(seed)
    np.random.rand = rec_rand
    ref.random.shuffle = rec_shuffle
    try:
        walks = g.simulate_walks(num_walks, walk_length)
    finally:
        np.random.rand = real_rand
        ref.random.shuffle = real_shuffle
    W = np.full((len(walks), walk_length), -1, dtype=np.int32)
    lens = np.zeros(len(walks), dtype=np.int32)
    for i, w in enumerate(walks):
        lens[i] = len(w)
        W[i, :len(w)] = [rank[x] for x in w]
    starts = np.array(orders, dtype=np.int64).reshape(-1)
    assert (W[:, 0] == starts).all()
    uni = np.array(draws, dtype=np.float64)
    assert len(uni) == 2 * int((lens - 1).sum())
    h = hashlib.sha256()
    for t in range(len(flat["ae_off"]) - 1):            # per table (u,v) ascending: J then q
        a, b = flat["ae_off"][t], flat["ae_off"][t + 1]
        h.update(flat["ae_J"][a:b].astype("<i8").tobytes())
        h.update(flat["ae_q"][a:b].astype("<f8").tobytes())
    meta = dict(name=name, file=os.path.basename(path), delimiter=delimiter, weighted=weighted,
                directed=directed, p=p, q=q, walk_length=walk_length, num_walks=num_walks,
                seed=seed, n_nodes=int(len(flat["node_ids"])), nnz=int(len(flat["col_idx"])),
                n_alias_edge_entries=int(len(flat["ae_J"])), n_walks=int(len(walks)),
                n_steps=int((lens - 1).sum()), n_uniforms=int(len(uni)),
                sha256_walks=hashlib.sha256(W.astype("<i4").tobytes()).hexdigest(),
                sha256_alias_edges=h.hexdigest(),
                first_walk_ids=[int(x) for x in walks[0][:16]])
    out = dict(flat)
    out.update(starts=starts, walks=W, lens=lens, meta=np.array(json.dumps(meta)))
    if store_uniforms:
        out["uniforms"] = uni
    np.savez_compressed(os.path.join(HERE, "n2v_%s.npz" % name), **out)
    print(name, json.dumps(meta)[:300])
    return meta


def write_weighted_directed(path):
    """Small weighted digraph with sinks, duplicate lines, a self loop and ids with gaps."""
    rs = np.random.RandomState(7)
    lines = []
    n = 40
    for _ in range(220):
        u, v = int(rs.randint(0, n)) * 3 + 5, int(rs.randint(0, n)) * 3 + 5
        w = float(rs.randint(1, 9)) * 0.25               # dyadic: sums exact either way
        lines.append("%d %d %.2f" % (u, v, w))
    lines.append("5 5 1.50")
    lines.append(lines[3])                               # duplicate edge, same weight
    lines.append("500 8 2.00")                           # 500 has out-edges only
    lines.append("8 700 0.75")                           # 700 is a sink
    with open(path, "w") as f:
        f.write("\n".join(lines) + "\n")


def main():
    os.makedirs(DATA, exist_ok=True)
    karate = os.path.join(DATA, "karate.edgelist")
    shutil.copyfile(os.path.join(REF, "node2vec/graph/karate.edgelist"), karate)
    g333 = os.path.join(DATA, "0_333_5038.txt")
    shutil.copyfile(os.path.join(REF, "IsoMap_LE/data/0_333_5038.txt"), g333)
    shutil.copyfile(os.path.join(REF, "IsoMap_LE/data/0_333_5038_simrank_navie_top10.txt.sim.txt"),
                    os.path.join(DATA, "0_333_5038_simrank_navie_top10.txt.sim.txt"))
    moreno = os.path.join(DATA, "moreno_crime_crime.txt")
    shutil.copyfile(os.path.join(REF, "DeepSim/lshrank_data/realdata/moreno_crime_crime.txt"), moreno)
    with open(os.path.join(REF, "DeepSim/lshrank_data/realdata/blog.txt"), "rb") as fi, \
            gzip.GzipFile(os.path.join(DATA, "blog.txt.gz"), "wb", mtime=0) as fo:
        shutil.copyfileobj(fi, fo)
    wd = os.path.join(DATA, "wdir_small.txt")
    write_weighted_directed(wd)

    metas = []
    # config 1 (anchor): karate p=1 q=1 L=80 r=10 — full run, uniforms stored
    metas.append(run_case("karate_p1_q1", karate, " ", False, False, 1.0, 1.0, 80, 10, 0))
    # SURVEY §8(c) walk KAT: karate p=0.25 q=4 seeds 0
    metas.append(run_case("karate_p025_q4", karate, " ", False, False, 0.25, 4.0, 80, 10, 0))
    # non-dyadic p,q : exercises fp64 rounding order in normalisation
    metas.append(run_case("karate_p3_q07", karate, " ", False, False, 3.0, 0.7, 40, 2, 1))
    # weighted + directed, ragged walks (sinks)
    metas.append(run_case("wdir_p05_q2", wd, " ", True, True, 0.5, 2.0, 30, 3, 2))
    # weighted + undirected (to_undirected weight merge)
    metas.append(run_case("wund_p2_q05", wd, " ", True, False, 2.0, 0.5, 30, 2, 3))
    # 333-vertex graph (listed in both directions), TAB graph moreno
    metas.append(run_case("g333_p025_q4", g333, " ", False, False, 0.25, 4.0, 20, 1, 4))
    metas.append(run_case("moreno_p025_q4", moreno, "\t", False, False, 0.25, 4.0, 30, 1, 5))

    # known-answer tests of alias_setup itself
    kats = []
    for probs in ([0.1, 0.2, 0.3, 0.4], [0.5, 0.3, 0.2], [], [1.0], [0.25] * 4,
                  [1.0 / 3] * 3, [0.7, 0.1, 0.1, 0.05, 0.05]):
        J, q = ref.alias_setup(probs)
        kats.append(dict(probs=probs, J=[int(x) for x in J], q_hex=[float(x).hex() for x in q]))
    with open(os.path.join(HERE, "alias_setup_kat.json"), "w") as f:
        json.dump(kats, f, indent=1)
    with open(os.path.join(HERE, "n2v_cases.json"), "w") as f:
        json.dump(metas, f, indent=1)


if __name__ == "__main__":
    main()
