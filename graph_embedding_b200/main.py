"""Drop-in for ``node2vec/src/main.py``: same flags, walk generation on the B200.

    python -m graph_embedding_b200.main --input graph/karate.edgelist --delimiter " " \
           --output karate.walks --p 0.25 --q 4 --walk-length 80 --num-walks 10

The reference pipeline is read_graph -> Graph -> preprocess_transition_probs ->
simulate_walks -> Word2Vec (main.py:104-114).  Word2Vec training is outside the hot path
(SURVEY.md §8); this CLI stops after the walks and writes them to --output in the corpus
format of DeepSim/src/main.py:237-244 (ids joined by TAB, trailing TAB, one walk per line),
which is what the DeepSim driver re-reads.  The Word2Vec flags are accepted and ignored.
"""
import argparse

from . import _lib
from . import node2vec


def parse_args(argv=None, p=1, q=1):
    parser = argparse.ArgumentParser(description="Run node2vec walk generation on B200.")
    parser.add_argument('--input', nargs='?', default='graph/karate.edgelist', help='Input graph path')
    parser.add_argument('--output', nargs='?', default='walks.txt', help='Walk corpus path')
    parser.add_argument('--groups', nargs='?', default=None, help='(ignored: classification input)')
    parser.add_argument('--dimensions', type=int, default=128, help='(ignored: Word2Vec)')
    parser.add_argument('--walk-length', type=int, default=80, help='Length of walk per source. Default is 80.')
    parser.add_argument('--num-walks', type=int, default=10, help='Number of walks per source. Default is 10.')
    parser.add_argument('--window-size', type=int, default=10, help='(ignored: Word2Vec)')
    parser.add_argument('--iter', default=10, type=int, help='(ignored: Word2Vec)')
    parser.add_argument('--workers', type=int, default=8, help='(ignored: Word2Vec)')
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


def main(args):
    nx_G = read_graph(args)
    G = node2vec.Graph(nx_G, args.directed, args.p, args.q)
    G.preprocess_transition_probs(materialize_edges=False)
    walks = G.simulate_walks(args.num_walks, args.walk_length)
    save_list(walks, args.output)
    return walks


if __name__ == "__main__":
    main(parse_args())
