"""Times the TopSim_singleSample production kernel (GW_SIMRANK_HYBRID) on a BA graph."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from graph_embedding_b200 import _lib
n = int(os.environ.get("BA", 1000000))
g = _lib.GraphHandle.barabasi_albert(n, 8, seed=1)
q = np.random.RandomState(3).choice(g.n, int(os.environ.get("NQ", 2048)), replace=False).astype(np.int64)
for i in range(3):
    t0 = time.perf_counter()
    ids, sc = g.simrank_topk(q, 0.6, 5, 10000, 20, mode=_lib.GW_SIMRANK_HYBRID, seed=1 + i)
    dt = time.perf_counter() - t0
    print("hybrid: %d queries in %.1f ms = %.0f queries/s, %d tree steps per query" % (len(q), dt * 1e3, len(q) / dt, g.simrank_last_steps() / len(q)), flush=True)
print("slow queries (HY_PROFILE builds print their phase totals here):", g.simrank_last_slow_queries(), flush=True)
