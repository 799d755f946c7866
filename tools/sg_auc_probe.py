"""Which link-prediction score describes a one-epoch embedding of the bench shape honestly?  Global cosine AUC mixes in the
frequency effect (hubs are subsampled and pulled by many contexts, leaves keep a common early direction)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from graph_embedding_b200 import _lib
scale = int(os.environ.get("SCALE", 22))
g = _lib.GraphHandle.rmat(scale, 16 << scale, a=0.45, b=0.15, c=0.15, seed=1)
g.prepare_walks()
nodes = g.nonisolated()
c = g.csr(weights=False, node_ids=False, first_seen=False)
deg = np.diff(c["row_ptr"])
for passes in (1, 3):
    starts = np.stack([np.random.RandomState(7 + i).permutation(nodes) for i in range(passes)])
    vec, cnt, sec = _lib.node2vec_embeddings(g, 0.25, 4.0, 80, passes, starts, dimensions=128, window=10, iter=1, negative=5, sample=1e-3, seed=11)
    rs = np.random.RandomState(0)
    e = rs.randint(0, g.nnz, size=50000)
    eu = np.searchsorted(c["row_ptr"], e, side="right") - 1
    ev = c["col_idx"][e]
    ru, rv = rs.choice(nodes, 50000), rs.choice(nodes, 50000)
    def cos(a, b, v=vec):
        x, y = v[a], v[b]
        return (x * y).sum(1) / np.maximum(np.linalg.norm(x, axis=1) * np.linalg.norm(y, axis=1), 1e-20)
    def auc(pos, neg):
        pos = np.sort(pos)
        return 1.0 - np.searchsorted(pos, neg, side="left").sum() / (len(pos) * len(neg))
    mean = vec[nodes].mean(0)
    vc = vec - mean
    low = (deg[eu] <= 64) & (deg[ev] <= 64)
    print("passes %d: global cos %.3f | same-source cos %.3f | same-source dot %.3f | centred global cos %.3f | centred same-source %.3f | low-degree edges (%d) global %.3f | |mean|/mean|v| %.3f"
          % (passes, auc(cos(eu, ev), cos(ru, rv)), float((cos(eu, ev) > cos(eu, rv)).mean()),
             float(((vec[eu] * vec[ev]).sum(1) > (vec[eu] * vec[rv]).sum(1)).mean()),
             auc(cos(eu, ev, vc), cos(ru, rv, vc)), float((cos(eu, ev, vc) > cos(eu, rv, vc)).mean()),
             int(low.sum()), auc(cos(eu[low], ev[low]), cos(ru, rv)),
             float(np.linalg.norm(mean) / np.linalg.norm(vec[nodes], axis=1).mean())), flush=True)
