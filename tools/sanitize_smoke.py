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
m.simrank_cache_javarng(q[:4], 0.6, 5, 300, 40, [1, 2, 3, 4], mode=0)
m.simrank_cache_javarng(q[:4], 0.6, 5, 300, 7, [1, 2, 3, 4], mode=1)
pp = m.double_walk_paths(np.arange(333), 30, 3, seed=1)
m.double_walk_sims(pp, 0.6, rows=q[:6], exact_order=True)
m.double_walk_sims(pp, 0.6, rows=q[:6], exact_order=False)
pj, _ = m.double_walk_paths(np.arange(20), 10, 3, rng_states=[5] * 20)
mm = m.topsim_mass(q[:6], 500.0, 3, seed=2)
m.topsim_mass_sims(mm, 0.6, [0, 1, 2], [3, 4, 5], exact_order=True)
m.topsim_mass_sims(mm, 0.6, [0, 1, 2], [3, 4, 5], exact_order=False)
m.topsim_mass(q[:3], 40.0, 4, rng_states=[1, 2, 3])
for ridx in ("0", "1"):                                        # both walker instantiations on a fresh handle
    os.environ["GW_CN_RIDX"] = ridx
    k2 = _lib.GraphHandle.from_file(os.path.join(DATA, "karate.edgelist"), delimiter=" ")
    for p_, q_ in ((0.25, 4.0), (4.0, 0.5), (1.0, 1.0), (8.0, 4.0)):
        w_, l_ = k2.walks(p_, q_, 33, st[:200], seed=2)
        assert (l_ == 33).all() and w_.min() >= 0
os.environ.pop("GW_CN_RIDX")
c1 = _lib.Comm(0, 1, _lib.Comm.unique_id(), 0)
c1.walks(h, 0.25, 4.0, 16, st[:100], seed=1)
c1.simrank_topk(m, q[:8], 0.6, 5, 500, 20, seed=1)
c1.close()
print("sanitize smoke done, launches =", _lib.kernel_launches())
