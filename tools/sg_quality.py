import sys, os, time
sys.path.insert(0, '/root/repo'); sys.path.insert(0, os.getcwd())
import numpy as np
from graph_embedding_b200 import _lib
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from _auc import edge_auc
for scale, dim, nw, L, it in ((14, 64, 10, 40, 1), (14, 64, 10, 40, 3), (16, 128, 10, 80, 1), (18, 128, 10, 80, 1)):
    h = _lib.GraphHandle.rmat(scale, 16 << scale, seed=1)
    c = h.csr(weights=False, node_ids=False, first_seen=False)
    rs = np.random.RandomState(1)
    nodes = h.nonisolated()
    starts = np.stack([rs.permutation(nodes) for _ in range(nw)])
    t0 = time.perf_counter()
    vec, cnt, sec = _lib.node2vec_embeddings(h, 1.0, 1.0, L, nw, starts, dimensions=dim, window=10, iter=it, seed=3)
    dt = time.perf_counter() - t0
    auc = edge_auc(vec, c["row_ptr"], c["col_idx"], np.random.RandomState(0), n_neg=20000)
    # same with the untrained initial vectors, for reference
    m = _lib.SkipGram(h, dim, seed=3)
    auc0 = edge_auc(m.vectors(), c["row_ptr"], c["col_idx"], np.random.RandomState(0), n_neg=20000)
    print("rmat%d dim %d walks %dx%d iter %d: AUC %.3f (untrained %.3f), %.2f s, training %.2f s" % (scale, dim, nw, L, it, auc, auc0, dt, sec["training"]), flush=True)
