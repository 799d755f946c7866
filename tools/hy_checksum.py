"""Checksums of the path-tree estimator's output on fixed inputs: a refactoring of the kernel must not change a bit."""
import hashlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from graph_embedding_b200 import _lib
for n, nq, sample, step in ((1000000, 1024, 10000, 5), (200000, 512, 1000, 3), (50000, 256, 20000, 4), (3000, 300, 100000, 5)):
    g = _lib.GraphHandle.barabasi_albert(n, 8, seed=1)
    q = np.random.RandomState(3).choice(g.n, nq, replace=False).astype(np.int64)
    ids, sc = g.simrank_topk(q, 0.6, step, sample, 20, mode=_lib.GW_SIMRANK_HYBRID, seed=9)
    h = hashlib.sha256(ids.tobytes() + sc.tobytes()).hexdigest()[:16]
    print("BA n=%d nq=%d sample=%d step=%d: %s steps=%d" % (n, nq, sample, step, h, g.simrank_last_steps()), flush=True)
