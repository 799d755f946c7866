"""Generates tests/golden/blog_exact_s5.npz: exact SimRank truncated at 5 sweeps (C = 0.6) on blog.txt for a fixed set
of 64 query vertices -- the EXPECTATION of SingleRandomWalk / TopSim_singleSample at STEP = 5 (BASELINE configs[1]).

Computed by the oracle's pinned exact routine (oracle/simrank_oracle.py simrank_exact_matrix: the array form of
SimRank.java:36-77, itself pinned to 5e-9 on the repository's shipped golden vector), ~100 s on one core for the
10313 x 10313 matrix; kept per query row: the 64 largest entries (the exact top-20 and the near-ties just below it)
and the 20 largest entries among targets of degree >= 16 (the well-conditioned part of the row), ids + fp64 scores +
target degrees -- all the parity tests read.

    python tests/golden/make_golden_blog.py
"""
import gzip
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import simrank_oracle as S  # noqa: E402

V = 10313
raw = gzip.open(os.path.join(HERE, "data", "blog.txt.gz"), "rb").read()
with tempfile.NamedTemporaryFile(suffix=".txt", delete=False) as f:
    f.write(raw)
g = S.load_multigraph(f.name, V, ",")
os.unlink(f.name)
deg = np.diff(g["row_ptr"])
assert deg[0] == 0 and deg.max() == 3992
rs = np.random.RandomState(20261018)
special = [0, int(np.argmax(deg))]                                   # the isolated slot, the 3992-degree hub
special += [int(np.nonzero(deg == d)[0][0]) for d in (1, 2, 64, 256, 1024) if (deg == d).any()]   # leaves, power-of-two degrees
rest = [int(v) for v in rs.permutation(np.arange(1, V)) if int(v) not in special]
queries = np.array(sorted(special + rest[:64 - len(special)]), dtype=np.int64)
exact = S.simrank_exact_matrix(g, 0.6, 5)
ids = np.zeros((64, 64), dtype=np.int32)
sc = np.zeros((64, 64), dtype=np.float64)
# The exact top-20 of a blog vertex are degree-1 / degree-2 targets (leaves of the hubs) whose scores are near-ties and
# whose Monte-Carlo increments are huge (C * deg(mid) / deg(target)): the estimator is heavy-tailed exactly there.  A
# second list keeps the 20 best targets of degree >= 16 per query, where a 1e-3 absolute criterion is meaningful.
wid = np.zeros((64, 20), dtype=np.int32)
wsc = np.zeros((64, 20), dtype=np.float64)
for r, v in enumerate(queries):
    order = np.lexsort((np.arange(V), -exact[v]))                     # score desc, id asc
    ids[r], sc[r] = order[:64], exact[v][order[:64]]
    wc = order[(deg[order] >= 16) & (order != v)][:20]
    wid[r], wsc[r] = wc, exact[v][wc]
np.savez_compressed(os.path.join(HERE, "blog_exact_s5.npz"), queries=queries, degrees=deg[queries].astype(np.int32),
                    top_ids=ids, top_scores=sc, top_degrees=deg[ids].astype(np.int32), wc_ids=wid, wc_scores=wsc,
                    wc_degrees=deg[wid].astype(np.int32), row_sums=exact[queries].sum(axis=1), c=0.6, sweeps=5)
print("wrote blog_exact_s5.npz; degrees of the queries:", sorted(deg[queries].tolist())[-5:], "sum of row sums", exact[queries].sum())
