"""Precision-vs-SAMPLE curves of the reference's drivers (benchmark/Test_u_u_SingleRandomWalk_Sample.java:35,
Test_u_u_TopSim_singleSample.java, Test_u_u_doubleRandomWalk_Sample.java, Test_u_u_TopSim_doubleSample.java,
Test_u_u_TopSim_Dev.java) through the device estimators.  Gold = exact SimRank truncated at the estimator's STEP
(utils/Eval.java:81-131 semantics: |gold top-k ids with score >= MIN  ∩  estimated top-k ids| / min(k, |gold|), mean
over vertices).  Prints one JSON object; run on a GPU box:  python tools/precision_curve.py > gpurun_out/precision.json"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from graph_embedding_b200 import _lib, simrank as sr

DATA = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "data")
K, MIN = 20, sr.MyConfiguration.MIN


def top_sets_dense(rows, k=K):
    out = []
    for r in rows:
        idx = np.argpartition(-r, min(k, len(r) - 1))[:k]
        out.append({int(i) for i in idx if r[i] >= MIN})
    return out


def top_sets_ids(ids, sc):
    return [{int(i) for i, x in zip(a, b) if i >= 0 and x >= MIN} for a, b in zip(ids, sc)]


def precision(gold, est):
    p = []
    for g, e in zip(gold, est):
        rk = min(K, len(g))
        p.append(1.0 if rk == 0 else len(g & e) / rk)
    return float(np.mean(p))


def timed(f):
    t0 = time.perf_counter()
    r = f()
    return r, time.perf_counter() - t0


out = {"k": K, "C": sr.MyConfiguration.C}

# ---- blog.txt (BASELINE configs[1]): every vertex is a query, STEP = 5 ----
g = sr.Graph(os.path.join(DATA, "blog.txt.gz"), 10313)
exact, t_exact = timed(lambda: g.handle.simrank_exact(0.6, 5))
gold = top_sets_dense(exact)
del exact
blog = {"graph": "blog.txt (10313 slots, 333983 edges)", "step": 5, "gold": "exact SimRank, 5 sweeps, %.3f s on the device" % t_exact,
        "SingleRandomWalk": [], "TopSim_singleSample": [], "SingleRandomWalk_M(M=2)": []}
q = np.arange(10313, dtype=np.int64)
for sample in (1000, 2500, 5000, 10000, 20000, 40000):
    for name, mode in (("SingleRandomWalk", _lib.GW_SIMRANK_MC), ("TopSim_singleSample", _lib.GW_SIMRANK_HYBRID)):
        g.handle.simrank_topk(q[:64], 0.6, 5, sample, K, mode, seed=1)            # warm-up (allocations)
        (ids, sc), t = timed(lambda: g.handle.simrank_topk(q, 0.6, 5, sample, K, mode, seed=1))
        blog[name].append({"sample": sample, "precision": precision(gold, top_sets_ids(ids, sc)), "seconds": t,
                           "queries_per_s": len(q) / t})
m, t = timed(lambda: sr.SingleRandomWalk_M(g, 2, 10000, seed=1).compute())
blog["SingleRandomWalk_M(M=2)"].append({"sample": 10000, "seconds": t, "precision": precision(
    gold, [{k for k, v in list(c)[-K:] if v >= MIN} for c in m.getResult()]), "note": "includes building 10313 host-side FixedCacheMap objects"})
out["blog"] = blog

# ---- 0_333_5038.txt: the pair estimators at the reference's own scale, STEP = 3 ----
h = sr.Graph(os.path.join(DATA, "0_333_5038.txt"), 333, separator=" ")
exact3 = h.handle.simrank_exact(0.6, 3)
gold3 = top_sets_dense(exact3)
small = {"graph": "0_333_5038.txt (333 vertices, 5038 lines)", "step": 3, "DoubleRandomWalk": [], "TopSim_doubleSample": [], "TopSim_Dev": []}
for sample in (50, 100, 200, 400, 800):
    d, t = timed(lambda: sr.DoubleRandomWalk(h, sample, 3, seed=1).compute())
    small["DoubleRandomWalk"].append({"sample": sample, "precision": precision(gold3, top_sets_dense(d.getResult())), "seconds": t})
for sample in (200, 1000, 10000):
    d, t = timed(lambda: sr.TopSim_doubleSample(h, sample, 3, seed=1).compute())
    small["TopSim_doubleSample"].append({"sample": sample, "precision": precision(gold3, top_sets_dense(d.getResult())), "seconds": t})
for single_step in (1, 2):
    cand = sr.TopSim_singleSample(h, 10000, single_step, seed=1).compute().getResult()
    d, t = timed(lambda: sr.TopSim_Dev(h, 10000, 3, K, single_step, seed=1).compute(cand))
    small["TopSim_Dev"].append({"sample": 10000, "singleStep": single_step, "tree_weight": d.SAMPLE,
                                "candidate_precision": precision(gold3, top_sets_dense(cand)),
                                "precision": precision(gold3, top_sets_dense(d.getResult())), "seconds": t})
out["g333"] = small
out["kernel_launches"] = _lib.kernel_launches()
print(json.dumps(out, indent=1))
