"""Small end-to-end pass over every kernel family for compute-sanitizer (memcheck / racecheck):
   compute-sanitizer --tool memcheck python tools/sanitize_smoke.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from graph_embedding_b200 import _lib
DATA = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "data")
h = _lib.GraphHandle.from_file(os.path.join(DATA, "karate.edgelist"), delimiter=" ")
h.alias_nodes(); h.alias_edges(0.25, 4.0)
st = np.tile(np.arange(h.n, dtype=np.int64), 8)
for p, q in ((0.25, 4.0), (4.0, 0.5), (1.0, 1.0)):
    h.walks(p, q, 40, st, seed=1)
os.environ["GW_CN_HUB"] = "1"
g = _lib.GraphHandle.rmat(12, 16 << 12, a=0.57, b=0.19, c=0.19, seed=2)
s2 = g.nonisolated()
for p, q in ((0.25, 4.0), (4.0, 0.5)):
    g.walks(p, q, 24, s2, seed=3)
os.environ.pop("GW_CN_HUB")
os.environ["GW_WALKER"] = "rejection"
g.walks(0.5, 2.0, 16, s2[:2000], seed=4)
os.environ.pop("GW_WALKER")
m = _lib.GraphHandle.from_file(os.path.join(DATA, "0_333_5038.txt"), delimiter=" ", mode=_lib.GW_MODE_MULTI, n_slots=333)
q = np.arange(0, 333, 7, dtype=np.int64)
m.simrank_topk(q, 0.6, 5, 1500, 20, seed=1)
m.simrank_topk(q, 0.6, 3, 700, 100, seed=1)
m.simrank_topk(q, 0.6, 7, 300, 20, seed=1)
m.simrank_topk(q, 0.6, 5, 1500, 20, mode=_lib.GW_SIMRANK_HYBRID, seed=1)
m.simrank_rows(q[:8], 0.6, 5, 500, seed=1)
m.simrank_rows(q[:8], 0.6, 5, 500, mode=_lib.GW_SIMRANK_HYBRID, seed=1)
m.simrank_rows_javarng(q[:4], 0.6, 5, 200, [1, 2, 3, 4])
m.topsim_rows_javarng(q[:4], 0.6, 3, 200, [1, 2, 3, 4])
m.simrank_exact(0.6, 3, rows=q[:4])
print("sanitize smoke done, launches =", _lib.kernel_launches())
