"""Dump SimRank top-k of a fixed query set (for bit-comparison of kernel variants): python tools/sr_compare.py out.npz"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from graph_embedding_b200 import _lib
g = _lib.GraphHandle.barabasi_albert(int(os.environ.get("BA", 1000000)), 8, seed=1)
q = np.random.RandomState(5).choice(g.n, size=2048, replace=False).astype(np.int64)
ids, sc = g.simrank_topk(q, 0.6, 5, 10000, 20, seed=11)
print("slow", g.simrank_last_slow_queries(), "steps", g.simrank_last_steps())
np.savez(sys.argv[1], ids=ids, sc=sc)
