"""Skip-gram on the walk corpus (node2vec/src/main.py:92-101 = gensim 0.13.3 Word2Vec, sg=1, negative
sampling).  gensim is absent from the image and the reference ships no trained embedding: PARITY UNPINNED.  What is
pinned: the device kernel to the restatement in oracle/sgns_oracle.py -- one warp walking the sentences in order
reproduces the restatement's vectors to rounding -- and the restatement's formulas to known answers worked out by hand
from gensim's published source; the parallel (Hogwild) mode is held to the same embedding quality."""
import os

import numpy as np
import pytest

from conftest import DATA
from oracle import n2v_oracle as O
from oracle import sgns_oracle as G


def _corpus(g, n_walks_per_node, L, seed):
    """uniform random walks on a CSR graph (numpy): a small corpus for the oracle"""
    rs = np.random.RandomState(seed)
    rp, col = g["row_ptr"], g["col_idx"]
    n = len(rp) - 1
    starts = np.tile(np.arange(n), n_walks_per_node)
    w = np.full((len(starts), L), -1, dtype=np.int32)
    w[:, 0] = starts
    for t in range(1, L):
        cur = w[:, t - 1]
        deg = rp[cur + 1] - rp[cur]
        k = (rs.random_sample(len(cur)) * deg).astype(np.int64)
        w[:, t] = col[rp[cur] + np.minimum(k, np.maximum(deg - 1, 0))]
    return w


def test_vocabulary_formulas_known_answers():
    # scale_vocab (gensim 0.13.3 word2vec.py): counts 900, 90, 10 of 1000 words, sample = 1e-2 -> threshold 10
    keep = G.scale_vocab([900, 90, 10, 0], 1e-2)
    p = [(np.sqrt(900 / 10.0) + 1) * 10 / 900.0, (np.sqrt(90 / 10.0) + 1) * 10 / 90.0, 1.0]
    assert keep[0] == int(round(p[0] * 2 ** 32)) and keep[1] == int(round(p[1] * 2 ** 32)) and keep[2] == 0xFFFFFFFF
    assert keep[3] == 0xFFFFFFFF and abs(p[0] - 0.11654) < 1e-4                     # a frequent word survives 11.7 % of the time
    assert (G.scale_vocab([5, 5], 0) == 0xFFFFFFFF).all()                           # sample = 0: nothing is dropped
    # make_cum_table: shares proportional to count ** 0.75
    tab = G.negative_table([16, 1, 0, 81], 1 << 10)
    share = np.bincount(tab, minlength=4) / float(len(tab))
    want = np.array([8.0, 1.0, 0.0, 27.0]) / 36.0
    assert np.abs(share - want).max() <= 1.0 / 1024 and share[2] == 0
    # EXP_TABLE: sigmoid sampled at 1000 points of [-6, 6)
    t = G.exp_table()
    assert abs(t[500] - 0.5) < 1e-6 and abs(t[0] - 1 / (1 + np.exp(6.0))) < 1e-6 and t[999] > 0.997
    # Philox4x32-10 known answer (Random123 kat_vectors: counter 0, key 0)
    assert G.philox4x32((0, 0, 0, 0), (0, 0)) == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]


def test_restated_trainer_learns_the_graph():
    """Two 8-cliques joined by one edge: after training, vertices of a clique are closer to each other than to the other
    clique, and edges score above non-edges (AUC)."""
    src, dst = [], []
    for base in (0, 8):
        for i in range(8):
            for j in range(i + 1, 8):
                src.append(base + i); dst.append(base + j)
    src.append(7); dst.append(8)
    g = O.build_simple_graph(np.array(src), np.array(dst), np.ones(len(src)), directed=False)
    walks = _corpus(g, 6, 20, 1)
    counts = np.bincount(walks[walks >= 0], minlength=16)
    keep, tab = G.scale_vocab(counts, 0.0), G.negative_table(counts, 1 << 12)
    syn0, syn1 = G.init_vectors(16, 32, 7)
    init = syn0.copy()
    total = float(counts.sum()) * 3
    done = 0.0
    for e in range(3):
        pairs = G.train(walks, syn0, syn1, keep, tab, 5, 5, 0.05, 0.0001, done, total, 7, sentence_id_base=e * len(walks), subsample=False)
        done += walks.size
    assert pairs > 5000 and np.abs(syn0 - init).max() > 0.05
    v = syn0 / np.linalg.norm(syn0, axis=1, keepdims=True)
    sim = v @ v.T
    inside = (sim[:8, :8].sum() - 8) / 56 + (sim[8:, 8:].sum() - 8) / 56
    across = sim[:8, 8:].mean() * 2
    assert inside > across + 0.5, (inside, across)


@pytest.mark.gpu
@pytest.mark.parametrize("dim,sample", [(32, 0.0), (128, 1e-2)])
def test_one_warp_in_order_reproduces_the_restatement(dim, sample):
    import torch
    from graph_embedding_b200 import _lib
    h = _lib.GraphHandle.from_file(os.path.join(DATA, "karate.edgelist"), delimiter=" ")
    g = O.load_graph(os.path.join(DATA, "karate.edgelist"), " ")
    walks = _corpus(g, 2, 24, 3)
    walks[::7, 15:] = -1                                                            # ragged sentences (dead ends pad with -1)
    n = h.n
    counts = np.bincount(walks[walks >= 0], minlength=n)
    d_w = torch.from_numpy(walks).cuda()
    m = _lib.SkipGram(h, dim, seed=11)
    m.count_dev(d_w.data_ptr(), len(walks), walks.shape[1])
    m.finalize_vocab(sample=sample, negative=5)
    v0, w0, c0 = m.vectors(syn1neg=True, counts=True)
    assert np.array_equal(c0, counts) and m.info()["total_words"] == counts.sum()
    syn0, syn1 = G.init_vectors(n, dim, 11)
    assert np.array_equal(v0, syn0) and not w0.any()                                # same initial vectors, bit for bit
    keep, tab = G.scale_vocab(counts, sample), G.negative_table(counts, G.table_size(n))
    total = float(counts.sum())
    m.train_dev(d_w.data_ptr(), len(walks), walks.shape[1], window=5, alpha=0.025, min_alpha=0.0001, words_before=0.0,
                total_words=total, sentence_id_base=100, subsample=sample > 0, sequential=True)
    pairs = G.train(walks, syn0, syn1, keep, tab, 5, 5, 0.025, 0.0001, 0.0, total, 11, sentence_id_base=100, subsample=sample > 0)
    v1, w1 = m.vectors(syn1neg=True)
    assert m.info()["trained_pairs"] == pairs and pairs > 1000
    assert np.abs(v1 - v0).max() > 1e-3                                             # it trained
    assert np.abs(v1 - syn0).max() <= 2e-6 and np.abs(w1 - syn1).max() <= 2e-6, (np.abs(v1 - syn0).max(), np.abs(w1 - syn1).max())


@pytest.mark.gpu
def test_parallel_training_reaches_the_restatements_quality_and_the_pipeline_call():
    import torch
    from graph_embedding_b200 import _lib
    h = _lib.GraphHandle.from_file(os.path.join(DATA, "karate.edgelist"), delimiter=" ")
    g = O.load_graph(os.path.join(DATA, "karate.edgelist"), " ")
    walks = _corpus(g, 10, 40, 5)
    n = h.n
    counts = np.bincount(walks[walks >= 0], minlength=n)
    # restatement, sequential, 2 epochs
    keep, tab = G.scale_vocab(counts, 0.0), G.negative_table(counts, G.table_size(n))
    syn0, syn1 = G.init_vectors(n, 32, 3)
    total = 2.0 * counts.sum()
    for e in range(2):
        G.train(walks, syn0, syn1, keep, tab, 5, 5, 0.025, 0.0001, e * counts.sum(), total, 3, sentence_id_base=e * len(walks), subsample=False)
    auc_ref = G.edge_auc(syn0, g["row_ptr"], g["col_idx"], np.random.RandomState(0))
    # device, sentences concurrently (Hogwild; the launcher keeps at most one concurrent warp per 16 words)
    d_w = torch.from_numpy(walks).cuda()
    m = _lib.SkipGram(h, 32, seed=3)
    m.count_dev(d_w.data_ptr(), len(walks), walks.shape[1])
    m.finalize_vocab(sample=0.0, negative=5)
    for e in range(2):
        m.train_dev(d_w.data_ptr(), len(walks), walks.shape[1], window=5, words_before=e * counts.sum(), total_words=total,
                    sentence_id_base=e * len(walks), subsample=False)
    auc_dev = G.edge_auc(m.vectors(), g["row_ptr"], g["col_idx"], np.random.RandomState(0))
    assert auc_ref > 0.75 and auc_dev > auc_ref - 0.05, (auc_ref, auc_dev)
    # gw_node2vec_embeddings: walks (p = 0.25, q = 4) -> vocabulary -> skip-gram in one call, nothing leaves the device
    rs = np.random.RandomState(1)
    starts = np.stack([rs.permutation(n) for _ in range(10)])
    vec, cnt, sec = _lib.node2vec_embeddings(h, 0.25, 4.0, 80, 10, starts, dimensions=128, window=10, iter=3, seed=5)
    assert cnt.sum() == 10 * n * 80 and vec.shape == (n, 128) and sec["training"] > 0
    assert G.edge_auc(vec, g["row_ptr"], g["col_idx"], np.random.RandomState(0)) > 0.8
    with pytest.raises(ValueError):
        _lib.SkipGram(h, 100)                                                       # dimensions: 32 / 64 / 128 / 256


@pytest.mark.gpu
def test_cli_writes_the_reference_output_and_learn_embeddings_takes_walk_lists(tmp_path):
    """node2vec/src/main.py:104-114 end to end: --output receives a word2vec text file with one row per node, most
    frequent node first (gensim's order); learn_embeddings(walks) is the reference's own function signature."""
    import random
    from graph_embedding_b200 import main as cli
    out = str(tmp_path / "karate.emb")
    args = cli.parse_args(["--input", os.path.join(DATA, "karate.edgelist"), "--delimiter", " ", "--output", out,
                           "--p", "0.25", "--q", "4", "--walk-length", "80", "--num-walks", "10", "--iter", "2", "--dimensions", "64"])
    random.seed(0); np.random.seed(0)
    words, vec = cli.main(args)
    w2, v2 = cli.load_word2vec_format(out)
    assert w2 == words and v2.shape == (34, 64) and np.abs(v2 - vec).max() < 1e-6
    assert sorted(int(w) for w in words) == list(range(1, 35)) and words[0] in ("34", "1")   # the two hubs are the most visited
    g = O.load_graph(os.path.join(DATA, "karate.edgelist"), " ")
    by_id = np.zeros((34, 64), dtype=np.float32)
    by_id[[int(w) - 1 for w in words]] = vec
    assert G.edge_auc(by_id, g["row_ptr"], g["col_idx"], np.random.RandomState(0)) > 0.75
    # the reference's function: a list of walks (lists of node ids) in, the same file format out
    args2 = cli.parse_args(["--input", os.path.join(DATA, "karate.edgelist"), "--delimiter", " ", "--output", str(tmp_path / "w.txt"),
                            "--emit", "walks", "--walk-length", "40", "--num-walks", "5"])
    walks = cli.main(args2)
    args3 = cli.parse_args(["--output", str(tmp_path / "k2.emb"), "--iter", "2", "--dimensions", "32", "--window-size", "5"])
    words3, vec3 = cli.learn_embeddings(walks, args3)
    w4, v4 = cli.load_word2vec_format(str(tmp_path / "k2.emb"))
    assert w4 == words3 and v4.shape == (34, 32) and sorted(int(w) for w in words3) == list(range(1, 35))
    cnt = np.bincount(np.array([t for w in walks for t in w]), minlength=35)
    assert [cnt[int(w)] for w in words3] == sorted((cnt[int(w)] for w in words3), reverse=True)      # descending count


@pytest.mark.gpu
def test_pipeline_at_scale_learns_the_graph():
    """R-MAT scale-16 (65 k vertices, 1 M undirected edges), 10 passes of walks, one epoch, 9 472 warps racing on the rows
    (Hogwild): the endpoints of an edge end up closer than random vertex pairs (AUC 0.93 measured; 0.50 untrained)."""
    from graph_embedding_b200 import _lib
    h = _lib.GraphHandle.rmat(16, 16 << 16, seed=1)
    c = h.csr(weights=False, node_ids=False, first_seen=False)
    rs = np.random.RandomState(1)
    nodes = h.nonisolated()
    starts = np.stack([rs.permutation(nodes) for _ in range(10)])
    vec, cnt, sec = _lib.node2vec_embeddings(h, 1.0, 1.0, 80, 10, starts, dimensions=128, window=10, iter=1, seed=3)
    assert cnt.sum() == 10 * len(nodes) * 80 and np.isfinite(vec).all()
    deg = np.diff(c["row_ptr"])
    assert np.corrcoef(cnt[deg > 0], deg[deg > 0])[0, 1] > 0.95                      # first-order walks visit ~ degree
    assert G.edge_auc(vec, c["row_ptr"], c["col_idx"], np.random.RandomState(0)) > 0.85
    m = _lib.SkipGram(h, 128, seed=3)
    assert abs(G.edge_auc(m.vectors(), c["row_ptr"], c["col_idx"], np.random.RandomState(0)) - 0.5) < 0.05


@pytest.mark.gpu
@pytest.mark.parametrize("dim,negative", [(64, 1), (256, 8)])
def test_other_dimensions_and_negative_counts_match_the_restatement(dim, negative):
    """64 and 256 dimensions (2 and 8 floats per lane), 1 and 8 negatives (less than one group of rows, more than one)."""
    import torch
    from graph_embedding_b200 import _lib
    h = _lib.GraphHandle.from_file(os.path.join(DATA, "karate.edgelist"), delimiter=" ")
    g = O.load_graph(os.path.join(DATA, "karate.edgelist"), " ")
    walks = _corpus(g, 1, 16, 9)
    n = h.n
    counts = np.bincount(walks[walks >= 0], minlength=n)
    d_w = torch.from_numpy(walks).cuda()
    m = _lib.SkipGram(h, dim, seed=21)
    m.count_dev(d_w.data_ptr(), len(walks), walks.shape[1])
    m.finalize_vocab(sample=0.0, negative=negative)
    syn0, syn1 = G.init_vectors(n, dim, 21)
    keep, tab = G.scale_vocab(counts, 0.0), G.negative_table(counts, G.table_size(n))
    total = float(counts.sum())
    m.train_dev(d_w.data_ptr(), len(walks), walks.shape[1], window=4, alpha=0.05, min_alpha=0.001, total_words=total,
                sentence_id_base=7, subsample=False, sequential=True)
    pairs = G.train(walks, syn0, syn1, keep, tab, 4, negative, 0.05, 0.001, 0.0, total, 21, sentence_id_base=7, subsample=False)
    v1, w1 = m.vectors(syn1neg=True)
    assert m.info()["trained_pairs"] == pairs
    assert np.abs(v1 - syn0).max() <= 2e-6 and np.abs(w1 - syn1).max() <= 2e-6
    with pytest.raises(ValueError):
        m.finalize_vocab(sample=0.0, negative=0)
