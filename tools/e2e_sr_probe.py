"""Times the blocking host API gw_simrank_topk call by call (e2e variance probe)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from graph_embedding_b200 import _lib
g = _lib.GraphHandle.barabasi_albert(int(os.environ.get("BA", 10000000)), 8, seed=1)
q = np.random.RandomState(2).choice(g.n, size=8192, replace=False).astype(np.int64)
for i in range(8):
    t0 = time.perf_counter()
    g.simrank_topk(q, 0.6, 5, 10000, 20, seed=9 + i)
    dt = time.perf_counter() - t0
    print("call %d: %.2f ms  slow=%d" % (i, dt * 1e3, g.simrank_last_slow_queries()), flush=True)
