#!/usr/bin/env python
"""Per-row measurements for SURVEY.md §8(a): every function of the hot path timed on one B200 through
the C ABI, with the CPU oracle timed beside it on a bounded sample (this is a measurement tool like
bench.py: it may execute oracle/ as the baseline, never as the thing measured).

    python tools/row_bench.py > gpurun_out/rows.json

Rows: a1 loader (edgelist -> CSR), a2-a4 alias tables (node + edge tables, entries/s), a5-a7 replay
walker (bit-exact mode), a9-a10 SingleRandomWalk (dense rows API = hash kernel; top-k API = log
kernel), a11 TopSim_singleSample hybrid tree, a13 exact SimRank.
"""
import gzip
import json
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

from graph_embedding_b200 import _lib  # noqa: E402

DATA = os.path.join(ROOT, "tests", "golden", "data")
out = {}


def timed(fn, reps=3):
    fn()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        r = fn()
        ts.append(time.perf_counter() - t0)
    return min(ts), r


# ---- a1 / a8: loaders on blog.txt (333 983 lines) ------------------------------------------------
with gzip.open(os.path.join(DATA, "blog.txt.gz"), "rb") as f:
    raw = f.read()
tmp = tempfile.NamedTemporaryFile(suffix=".txt", delete=False)
tmp.write(raw)
tmp.close()
t_simple, h_blog = timed(lambda: _lib.GraphHandle.from_file(tmp.name, delimiter=","))
t_multi, h_blog_m = timed(lambda: _lib.GraphHandle.from_file(tmp.name, delimiter=",", mode=_lib.GW_MODE_MULTI, n_slots=10313))
out["a1_read_graph_blog"] = {"seconds": t_simple, "edges_per_s": 333983 / t_simple, "what": "text edgelist -> sorted CSR (SIMPLE mode), host parse + device build"}
out["a8_Graph_java_blog"] = {"seconds": t_multi, "edges_per_s": 333983 / t_multi, "what": "text edgelist -> file-order multigraph CSR (MULTI mode)"}
t_rmat, h_rmat = timed(lambda: _lib.GraphHandle.rmat(22, 16 << 22, seed=1), reps=2)
out["a1_device_build_rmat22"] = {"seconds": t_rmat, "directed_entries": int(h_rmat.nnz), "entries_per_s": h_rmat.nnz / t_rmat,
                                 "what": "R-MAT generator + symmetrise + radix sort + unique + row pointers, all on device"}
del h_rmat

# ---- a2-a4: alias tables on blog (sum deg = 667 966 node entries, sum deg^2 = 3.69e8 edge entries) ----
t_an, _ = timed(lambda: h_blog.alias_nodes(), reps=2)
n_edge_entries = h_blog.alias_edges_size()
_pp = [0.25]
def _build_edges():                       # the tables are cached per (p, q): change p every call so that each call rebuilds
    _pp[0] += 0.125
    return h_blog.alias_edges(_pp[0], 4.0, budget_bytes=16 << 30, fetch=False)
t_ae, _ = timed(_build_edges, reps=2)
out["a4_alias_nodes_blog"] = {"seconds": t_an, "entries": int(h_blog.nnz), "entries_per_s": h_blog.nnz / t_an}
out["a3_a4_alias_edges_blog"] = {"seconds": t_ae, "entries": int(n_edge_entries), "entries_per_s": n_edge_entries / t_ae,
                                 "what": "preprocess_transition_probs' alias_edges, one bit-exact Vose table per directed edge (p=0.25, q=4), tables stay on the device"}

# CPU: the oracle's restatement of alias_setup/get_alias_edge on the 333-vertex graph (bounded)
from oracle import n2v_oracle as O  # noqa: E402
og = O.load_graph(os.path.join(DATA, "0_333_5038.txt"), " ")
t0 = time.perf_counter()
off, J, q = O.alias_edges_flat(og, 0.25, 4.0)
t_cpu = time.perf_counter() - t0
out["a3_a4_alias_edges_cpu_port"] = {"seconds": t_cpu, "entries": int(len(J)), "entries_per_s": len(J) / t_cpu, "cores": 1,
                                     "what": "oracle/n2v_oracle.py on 0_333_5038.txt (sum deg^2 = 156 352)"}

# ---- a5-a7: replay walker (bit-exact mode) ---------------------------------------------------------
z = np.load(os.path.join(ROOT, "tests", "golden", "n2v_moreno_p025_q4.npz"))

cases = json.load(open(os.path.join(ROOT, "tests", "golden", "n2v_cases.json")))
meta = [c for c in (cases["cases"] if isinstance(cases, dict) and "cases" in cases else cases) if c["name"] == "moreno_p025_q4"][0]
hm = _lib.GraphHandle.from_file(os.path.join(DATA, meta["file"]), delimiter=meta["delimiter"], weighted=meta["weighted"], directed=meta["directed"])
hm.alias_nodes()
hm.alias_edges(meta["p"], meta["q"], fetch=False)
starts, uni = z["starts"].astype(np.int64), z["uniforms"]
L = int(z["walks"].shape[1])
t_rep, (w, ln) = timed(lambda: hm.walks_replay(L, starts, uni))
steps = int((ln - 1).sum())
out["a5_a7_replay_moreno"] = {"seconds": t_rep, "walk_steps": steps, "steps_per_s": steps / t_rep,
                              "what": "k_walk_replay fed the reference's recorded uniforms (host buffers in and out); bit-exact walks"}

# ---- a9/a10 dense-rows API (hash kernel), a11 hybrid, on the blog multigraph ------------------------
qs = np.arange(0, 10313, 41, dtype=np.int64)[:252]
t_rows, _ = timed(lambda: h_blog_m.simrank_rows(qs, 0.6, 5, 10000, seed=1), reps=2)
out["a9_a10_getResult_rows_blog"] = {"seconds": t_rows, "queries": len(qs), "queries_per_s": len(qs) / t_rows,
                                     "what": "SingleRandomWalk.compute()/getResult() dense rows (hash kernel, 10313 doubles per query copied to host)"}
t_topk, _ = timed(lambda: h_blog_m.simrank_topk(np.arange(10313, dtype=np.int64), 0.6, 5, 10000, 20, seed=1), reps=2)
out["a9_a12_topk_blog_all_vertices"] = {"seconds": t_topk, "queries": 10313, "queries_per_s": 10313 / t_topk,
                                        "slow_path_queries": int(h_blog_m.simrank_last_slow_queries()),
                                        "what": "BASELINE configs[1]: every vertex of blog.txt, C=0.6 STEP=5 SAMPLE=10000, top-20 (log kernel + hash hand-over)"}
t_hy, _ = timed(lambda: h_blog_m.simrank_topk(qs, 0.6, 5, 10000, 20, mode=_lib.GW_SIMRANK_HYBRID, seed=1), reps=2)
out["a11_TopSim_singleSample_blog"] = {"seconds": t_hy, "queries": len(qs), "queries_per_s": len(qs) / t_hy,
                                       "what": "hybrid enumerate/sample path tree, SAMPLE=10000 STEP=5, top-20"}
hb = _lib.GraphHandle.barabasi_albert(1000000, 8, seed=1)
qb = np.random.RandomState(3).choice(hb.n, 2048, replace=False).astype(np.int64)
t_hyb, _ = timed(lambda: hb.simrank_topk(qb, 0.6, 5, 10000, 20, mode=_lib.GW_SIMRANK_HYBRID, seed=1), reps=2)
out["a11_TopSim_singleSample_ba1m"] = {"seconds": t_hyb, "queries": len(qb), "queries_per_s": len(qb) / t_hyb}

# ---- a13 exact SimRank on blog (dense 10313^2 fp64, 5 sweeps) ---------------------------------------
t_ex, _ = timed(lambda: h_blog_m.simrank_exact(0.6, 5, rows=np.arange(8, dtype=np.int64)), reps=1)
out["a13_exact_simrank_blog"] = {"seconds": t_ex, "iters": 5, "n": 10313, "what": "S <- c P S P^T, dense fp64, 5 sweeps, 8 rows copied out"}

# CPU: C restatement of SingleRandomWalk on blog (1 thread, bounded sample of queries)
from oracle import simrank_oracle as S  # noqa: E402
ogb = S.load_multigraph(tmp.name, 10313, ",")
st = S.java_seed(1)
t0 = time.perf_counter()
nqd = 0
while time.perf_counter() - t0 < 10.0:
    row, _, st = S.single_random_walk_row(ogb, int(qs[nqd % len(qs)]), 10000, 5, 0.6, st)
    S.fixedmaxpq_topk(row, 20)
    nqd += 1
t_cpu = time.perf_counter() - t0
out["a9_a12_cpu_port_blog"] = {"seconds": t_cpu, "queries": nqd, "queries_per_s": nqd / t_cpu, "cores": 1,
                               "what": "oracle/simrank_oracle.c SingleRandomWalk.walk + FixedMaxPQ on blog.txt"}
# CPU: the other rows' oracle routines on bounded samples (1 core each)
t0 = time.perf_counter()
nq2 = 0
st = S.java_seed(2)
while time.perf_counter() - t0 < 5.0:
    row, made, st = S.topsim_row(ogb, int(qs[nq2 % len(qs)]), 10000, 5, 0.6, mode=0, seed_state=st)
    S.fixedmaxpq_topk(row, 20)
    nq2 += 1
t_cpu = time.perf_counter() - t0
out["a11_cpu_port_blog"] = {"seconds": t_cpu, "queries": nq2, "queries_per_s": nq2 / t_cpu, "cores": 1,
                            "what": "oracle/simrank_oracle.c TopSim_singleSample path tree + FixedMaxPQ on blog.txt"}
o333 = S.load_multigraph(os.path.join(DATA, "0_333_5038.txt"), 333, " ")
t0 = time.perf_counter()
S.simrank_exact_matrix(o333, 0.6, 5)
t_cpu = time.perf_counter() - t0
t_g333, _ = timed(lambda: _lib.GraphHandle.from_file(os.path.join(DATA, "0_333_5038.txt"), delimiter=" ", mode=_lib.GW_MODE_MULTI,
                                                      n_slots=333).simrank_exact(0.6, 5), reps=2)
out["a13_exact_simrank_g333"] = {"gpu_seconds": t_g333, "cpu_port_seconds": t_cpu, "n": 333, "iters": 5,
                                 "what": "SimRank.compute on 0_333_5038.txt: device sweeps (incl. loading the file) vs the oracle's C restatement"}
t0 = time.perf_counter()
og2 = O.load_graph(tmp.name, ",")
t_cpu = time.perf_counter() - t0
out["a1_cpu_port_blog"] = {"seconds": t_cpu, "edges_per_s": 333983 / t_cpu, "cores": 1,
                           "what": "oracle/n2v_oracle.py parse_edgelist + build_simple_graph (networkx semantics) on blog.txt"}
an = O.alias_nodes_flat(og)
ae = O.alias_edges_flat(og, 0.25, 4.0)
rng = np.random.RandomState(1)
t0 = time.perf_counter()
steps_cpu = 0
i = 0
while time.perf_counter() - t0 < 5.0:
    wlk, _ = O.walk_replay(og, an, ae, 80, i % 333, rng.rand(2 * 79), 0)
    steps_cpu += len(wlk) - 1
    i += 1
t_cpu = time.perf_counter() - t0
out["a5_a7_cpu_port_g333"] = {"seconds": t_cpu, "walk_steps": steps_cpu, "steps_per_s": steps_cpu / t_cpu, "cores": 1,
                              "what": "oracle/n2v_oracle.py node2vec_walk/alias_draw on 0_333_5038.txt (materialised alias tables)"}
os.unlink(tmp.name)
print(json.dumps(out, indent=1))
