"""Drop-in for ``node2vec/src/main.py``: same flags, the whole pipeline on the B200.

    python -m graph_embedding_b200.main --input graph/karate.edgelist --delimiter " " \
           --output karate.emb --p 0.25 --q 4 --walk-length 80 --num-walks 10

The reference pipeline is read_graph -> Graph -> preprocess_transition_probs -> simulate_walks -> Word2Vec ->
save_word2vec_format (main.py:104-114, :92-101).  ``--emit embeddings`` (default, what the reference writes to
--output) runs walks, vocabulary scan and skip-gram with negative sampling on the device in one call
(gw_node2vec_embeddings): the corpus is regenerated pass by pass from its seed and never leaves the GPU.
``--emit walks`` stops after the walks and writes them in the corpus format of DeepSim/src/main.py:237-244 (ids joined
by TAB, trailing TAB, one walk per line), which is what the DeepSim driver re-reads.  --workers is accepted and ignored
(the device runs every walk of a pass concurrently).
"""
import argparse

from . import _lib
from . import node2vec


def parse_args(argv=None, p=1, q=1):
    parser = argparse.ArgumentParser(description="Run node2vec walk generation on B200.")
    parser.add_argument('--input', nargs='?', default='graph/karate.edgelist', help='Input graph path')
    parser.add_argument('--output', nargs='?', default='walks.txt', help='Walk corpus path')
    parser.add_argument('--groups', nargs='?', default=None, help='(ignored: classification input)')
    parser.add_argument('--dimensions', type=int, default=128, help='Number of dimensions (32, 64, 128 or 256). Default is 128.')
    parser.add_argument('--walk-length', type=int, default=80, help='Length of walk per source. Default is 80.')
    parser.add_argument('--num-walks', type=int, default=10, help='Number of walks per source. Default is 10.')
    parser.add_argument('--window-size', type=int, default=10, help='Context size for optimization. Default is 10.')
    parser.add_argument('--iter', default=10, type=int, help='Number of epochs in SGD')
    parser.add_argument('--workers', type=int, default=8, help='(ignored: the device runs all walks of a pass concurrently)')
    parser.add_argument('--emit', choices=['embeddings', 'walks'], default='embeddings',
                        help='what --output receives: the word2vec-format embeddings (reference behaviour) or the walk corpus')
    parser.add_argument('--p', type=float, default=p, help='Return hyperparameter. Default is 1.')
    parser.add_argument('--q', type=float, default=q, help='Inout hyperparameter. Default is 1.')
    parser.add_argument('--delimiter', type=str, default=',', help='the delimiter of a graph. Default is ",".')
    parser.add_argument('--weighted', dest='weighted', action='store_true')
    parser.add_argument('--unweighted', dest='unweighted', action='store_false')
    parser.set_defaults(weighted=False)
    parser.add_argument('--directed', dest='directed', action='store_true')
    parser.add_argument('--undirected', dest='undirected', action='store_false')
    parser.set_defaults(directed=False)
    return parser.parse_args(argv)


def read_graph(args):
    """main.py:76-89 without networkx: the edge list goes straight to the device loader."""
    h = _lib.GraphHandle.from_file(args.input, delimiter=args.delimiter, weighted=args.weighted,
                                   directed=args.directed, mode=_lib.GW_MODE_SIMPLE)
    return node2vec.EdgeListGraph(h)


def save_list(walks, file_path):
    """DeepSim/src/main.py:237-244."""
    with open(file_path, "w") as f:
        for walk in walks:
            f.write("".join(str(t) + "\t" for t in walk))
            f.write("\n")


def read_list(file_path):
    """DeepSim/src/main.py:246-254 (tokens stay strings, as there)."""
    walks = []
    with open(file_path, "r") as f:
        for line in f.readlines():
            walks.append([w for w in line.strip().split("\t")])
    return walks


def save_word2vec_format(file_path, words, vectors):
    """The `.emb` file node2vec/src/main.py:98 writes (gensim 0.13.3 `save_word2vec_format`, text mode): header
    `<count> <dimensions>`, then one `<word> <%f> <%f> ...` line per vector (node2vec/emb/karate.emb).  Walk
    generation ends before Word2Vec; the format is kept so a trainer fed with this engine's corpus stays a drop-in."""
    with open(file_path, "w") as f:
        f.write("%d %d\n" % (len(words), len(vectors[0]) if len(words) else 0))
        for w, row in zip(words, vectors):
            f.write("%s %s\n" % (w, " ".join("%f" % float(x) for x in row)))


def load_word2vec_format(file_path):
    """-> (list of words as strings, float32 array [count, dimensions]); the reader side of the same format
    (node2vec/src/classify.py loads it through gensim)."""
    import numpy as np
    with open(file_path, "r") as f:
        count, dim = (int(x) for x in f.readline().split())
        words, vecs = [], np.zeros((count, dim), dtype=np.float32)
        for i in range(count):
            tok = f.readline().rstrip("\n").split(" ")
            if len(tok) != dim + 1:
                raise ValueError("line %d of %s has %d fields, expected %d" % (i + 2, file_path, len(tok), dim + 1))
            words.append(tok[0])
            vecs[i] = [float(x) for x in tok[1:]]
    return words, vecs


def _by_count(counts, order_hint):
    """gensim's output order: words sorted by descending count; ties keep the order of first appearance."""
    keep = [i for i in order_hint if counts[i] > 0]
    return sorted(keep, key=lambda i: -int(counts[i]))          # sorted() is stable


def learn_embeddings(walks, args):
    """node2vec/src/main.py:92-101: Word2Vec(walks, size=dimensions, window=window_size, min_count=0, sg=1, iter=iter) +
    save_word2vec_format(args.output).  `walks`: list of lists of node ids (what simulate_walks returns) -- uploaded
    once; vocabulary scan and training run on the device.  Returns (words, vectors) in the file's order."""
    import numpy as np
    import torch
    ids = sorted({t for w in walks for t in w})
    rank = {t: i for i, t in enumerate(ids)}
    L = max(len(w) for w in walks)
    arr = np.full((len(walks), L), -1, dtype=np.int32)
    for r, w in enumerate(walks):
        arr[r, :len(w)] = [rank[t] for t in w]
    d_w = torch.from_numpy(arr).cuda()
    m = _lib.SkipGram(len(ids), args.dimensions, seed=1, device=torch.cuda.current_device())
    m.count_dev(d_w.data_ptr(), arr.shape[0], L)
    m.finalize_vocab(sample=1e-3, negative=5)
    total = float(m.info()["total_words"]) * args.iter
    for e in range(args.iter):
        m.train_dev(d_w.data_ptr(), arr.shape[0], L, window=args.window_size, words_before=e * arr.size, total_words=total,
                    sentence_id_base=e * arr.shape[0])
    vec, cnt = m.vectors(counts=True)
    first = list(dict.fromkeys(rank[t] for w in walks for t in w))
    order = _by_count(cnt, first)
    words = [str(ids[i]) for i in order]
    save_word2vec_format(args.output, words, vec[order])
    print("Save.")
    return words, vec[order]


def main(args):
    nx_G = read_graph(args)
    G = node2vec.Graph(nx_G, args.directed, args.p, args.q)
    G.preprocess_transition_probs(materialize_edges=False)
    if args.emit == 'walks':
        walks = G.simulate_walks(args.num_walks, args.walk_length)
        save_list(walks, args.output)
        return walks
    # walks -> vocabulary -> skip-gram without the corpus leaving the device; the start orders keep the reference's
    # contract (cumulative random.shuffle of list(G.nodes()), node2vec.py:47-51; np.random seeds the device streams)
    import numpy as np
    order = G._dense_many(nx_G.nodes())
    seed = int(np.random.randint(0, 2 ** 31 - 1)) | (int(np.random.randint(0, 2 ** 31 - 1)) << 31)
    starts = []
    for _ in range(args.num_walks):
        _lib.py_random_shuffle(order)                     # random.shuffle(nodes), on Python's own generator
        starts.append(order.copy())
    vec, cnt, _ = _lib.node2vec_embeddings(G._h, args.p, args.q, args.walk_length, args.num_walks, np.stack(starts),
                                           dimensions=args.dimensions, window=args.window_size, iter=args.iter, seed=seed)
    order = _by_count(cnt, G._dense_many(nx_G.nodes()).tolist())
    words = [str(int(nx_G.node_ids[i])) for i in order]
    save_word2vec_format(args.output, words, vec[order])
    print("Save.")
    return words, vec[order]


if __name__ == "__main__":
    main(parse_args())
