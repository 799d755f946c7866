"""Python face of the SimRank CPU oracle (oracle/simrank_oracle.c).  TEST INFRASTRUCTURE ONLY.

Also restates, in numpy, the pieces that are array algebra:
  * the multigraph loader of structures/Graph.java:28-57,
  * exact SimRank as the matrix iteration S <- C * P S P^T, diag := 1 (SimRank.java:36-77),
  * utils/Print.java:25-84 (.sim.txt writer) and utils/Eval.java:81-131 (precision).
Parity status: see the header of simrank_oracle.c.
"""
from __future__ import annotations

import ctypes
import gzip
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

SEPARATOR = ","        # conf/MyConfiguration.java:16
SEPARATOR_KV = ":"     # :18
TOPK = 20              # :19
MIN = 0.000000001      # :20
C_DEFAULT = 0.6        # :21


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(HERE, "libsimrank_oracle.so")
        src = os.path.join(HERE, "simrank_oracle.c")
        if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
            subprocess.check_call(["make", "-s", "-C", HERE])
        L = ctypes.CDLL(so)
        i64p = ctypes.POINTER(ctypes.c_int64)
        i32p = ctypes.POINTER(ctypes.c_int32)
        f64p = ctypes.POINTER(ctypes.c_double)
        u64p = ctypes.POINTER(ctypes.c_uint64)
        L.or_single_random_walk_row.restype = ctypes.c_int64
        L.or_single_random_walk_row.argtypes = [ctypes.c_int64, i64p, i32p, ctypes.c_int32,
                                                ctypes.c_int32, ctypes.c_int32, ctypes.c_double,
                                                u64p, f64p]
        L.or_topsim_row.restype = ctypes.c_int64
        L.or_topsim_row.argtypes = [ctypes.c_int64, i64p, i32p, ctypes.c_int32, ctypes.c_int32,
                                    ctypes.c_int32, ctypes.c_double, ctypes.c_int32,
                                    ctypes.c_int64, u64p, f64p]
        L.or_simrank_exact.restype = None
        L.or_simrank_exact.argtypes = [ctypes.c_int64, i64p, i32p, ctypes.c_double,
                                       ctypes.c_int32, f64p]
        L.or_fixedmaxpq_topk.restype = ctypes.c_int32
        L.or_fixedmaxpq_topk.argtypes = [f64p, ctypes.c_int64, ctypes.c_int32, i32p, f64p]
        f32p = ctypes.POINTER(ctypes.c_float)
        L.or_fcm_drain.restype = None
        L.or_fcm_drain.argtypes = [ctypes.c_int32, i32p, f32p, i32p, f32p]
        L.or_fcm_puts.restype = ctypes.c_int32
        L.or_fcm_puts.argtypes = [ctypes.c_int32, ctypes.c_int32, i32p, f32p, ctypes.c_int32, i32p, f32p]
        L.or_single_random_walk_cache.restype = ctypes.c_int64
        L.or_single_random_walk_cache.argtypes = [ctypes.c_int64, i64p, i32p, ctypes.c_int32, ctypes.c_int32,
                                                  ctypes.c_int32, ctypes.c_double, ctypes.c_int32, u64p,
                                                  i32p, f32p, i32p]
        L.or_topsim_cache.restype = ctypes.c_int64
        L.or_topsim_cache.argtypes = [ctypes.c_int64, i64p, i32p, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32,
                                      ctypes.c_double, ctypes.c_int32, ctypes.c_int64, u64p, i32p, f32p, i32p]
        L.or_double_walk_paths.restype = None
        L.or_double_walk_paths.argtypes = [ctypes.c_int64, i64p, i32p, ctypes.c_int32, ctypes.c_int32, u64p, i32p]
        L.or_double_walk_matrix.restype = None
        L.or_double_walk_matrix.argtypes = [ctypes.c_int64, i32p, ctypes.c_int32, ctypes.c_int32, ctypes.c_double, f64p]
        L.or_mass_tree.restype = ctypes.c_int64
        L.or_mass_tree.argtypes = [ctypes.c_int64, i64p, i32p, ctypes.c_int32, ctypes.c_double, ctypes.c_int32,
                                   ctypes.c_int64, u64p, f64p]
        L.or_mass_sim.restype = ctypes.c_double
        L.or_mass_sim.argtypes = [f64p, f64p, ctypes.c_int64, ctypes.c_int32, ctypes.c_double]
        L.or_fixedmaxpq_topk_min.restype = ctypes.c_int32
        L.or_fixedmaxpq_topk_min.argtypes = [f64p, ctypes.c_int64, ctypes.c_int32, ctypes.c_double, i32p, f64p]
        L.jr_fill.restype = None
        L.jr_fill.argtypes = [ctypes.c_int64, ctypes.c_int32, ctypes.c_int32, i32p]
        L.or_node2vec_walks.restype = ctypes.c_int64
        L.or_node2vec_walks.argtypes = [ctypes.c_int64, i64p, i32p, ctypes.c_double,
                                        ctypes.c_double, ctypes.c_int32, i64p, ctypes.c_int64,
                                        ctypes.c_uint64, i32p]
        _LIB = L
    return _LIB


def _p(a, t):
    return a.ctypes.data_as(ctypes.POINTER(t))


def java_seed(s):
    """java.util.Random(long seed) initial scramble."""
    return (int(s) ^ 0x5DEECE66D) & ((1 << 48) - 1)


# ---------------- structures/Graph.java ----------------
def read_edge_file(path, separator=SEPARATOR):
    """Graph(String, int) :35-41 : every line, ids[0], ids[1] (String.split drops nothing
    in front; Integer.valueOf rejects blanks, so files are assumed clean)."""
    op = gzip.open if path.endswith(".gz") else open
    src, dst = [], []
    with op(path, "rt") as f:
        for line in f:
            line = line.rstrip("\r\n")
            if not line:
                continue
            ids = line.split(separator)
            src.append(int(ids[0]))
            dst.append(int(ids[1]))
    return np.asarray(src, dtype=np.int64), np.asarray(dst, dtype=np.int64)


def build_multigraph(src, dst, V):
    """addEdge (:53-57): adjs[from].add(to); adjs[to].add(from) per line, file order kept."""
    m = len(src)
    a = np.empty(2 * m, dtype=np.int64)
    b = np.empty(2 * m, dtype=np.int64)
    a[0::2], b[0::2] = src, dst
    a[1::2], b[1::2] = dst, src
    order = np.argsort(a, kind="stable")
    col = b[order].astype(np.int32)
    row_ptr = np.zeros(V + 1, dtype=np.int64)
    np.add.at(row_ptr, a + 1, 1)
    row_ptr = np.cumsum(row_ptr)
    return dict(V=int(V), row_ptr=row_ptr, col=np.ascontiguousarray(col))


def load_multigraph(path, V, separator=SEPARATOR):
    s, d = read_edge_file(path, separator)
    return build_multigraph(s, d, V)


# ---------------- estimators (C) ----------------
def single_random_walk_row(g, v, sample, step, C=C_DEFAULT, seed_state=None):
    """SingleRandomWalk.walk(v) + sim[v][v]=0.  Returns (row, steps, new_seed_state)."""
    row = np.zeros(g["V"], dtype=np.float64)
    st = ctypes.c_uint64(java_seed(0) if seed_state is None else seed_state)
    steps = lib().or_single_random_walk_row(g["V"], _p(g["row_ptr"], ctypes.c_int64),
                                            _p(g["col"], ctypes.c_int32), int(v), int(sample),
                                            int(step), float(C), ctypes.byref(st),
                                            _p(row, ctypes.c_double))
    return row, int(steps), int(st.value)


def topsim_row(g, v, sample, step, C=C_DEFAULT, mode=0, seed_state=None, max_paths=1 << 22):
    """TopSim_singleSample (mode 0) / TopSim_Enumerate (mode 1) for one source; x SAMPLE."""
    row = np.zeros(g["V"], dtype=np.float64)
    st = ctypes.c_uint64(java_seed(0) if seed_state is None else seed_state)
    made = lib().or_topsim_row(g["V"], _p(g["row_ptr"], ctypes.c_int64),
                               _p(g["col"], ctypes.c_int32), int(v), int(sample), int(step),
                               float(C), int(mode), int(max_paths), ctypes.byref(st),
                               _p(row, ctypes.c_double))
    if made < 0:
        raise MemoryError("path tree exceeded max_paths")
    return row, int(made), int(st.value)


# ---------------- lxctools/FixedCacheMap.java and the _M estimators ----------------
def fcm_puts(nmax, keys, vals, key_space=None):
    """A put() sequence into one FixedCacheMap(nmax); returns the heap arrays (keys, float32 values) in heap
    order (slot 1 first)."""
    pk = np.ascontiguousarray(keys, dtype=np.int32)
    pv = np.ascontiguousarray(vals, dtype=np.float32)
    hk = np.zeros(nmax + 1, dtype=np.int32)
    hv = np.zeros(nmax + 1, dtype=np.float32)
    ks = int(pk.max()) + 1 if key_space is None and len(pk) else int(key_space or 1)
    n = lib().or_fcm_puts(int(nmax), len(pk), _p(pk, ctypes.c_int32), _p(pv, ctypes.c_float), ks,
                          _p(hk, ctypes.c_int32), _p(hv, ctypes.c_float))
    return hk[1:n + 1].copy(), hv[1:n + 1].copy()


def fcm_drain(hkeys, hvals):
    """`for (Pair p : cacheMap)` (FixedCacheMap.java:102-132): delMin until empty -> ascending (keys, values)."""
    n = len(hkeys)
    hk = np.zeros(n + 1, dtype=np.int32); hk[1:] = hkeys
    hv = np.zeros(n + 1, dtype=np.float32); hv[1:] = hvals
    ok = np.zeros(max(n, 1), dtype=np.int32)
    ov = np.zeros(max(n, 1), dtype=np.float32)
    lib().or_fcm_drain(n, _p(hk, ctypes.c_int32), _p(hv, ctypes.c_float), _p(ok, ctypes.c_int32), _p(ov, ctypes.c_float))
    return ok[:n].copy(), ov[:n].copy()


def single_random_walk_cache(g, v, sample, step, capacity, C=C_DEFAULT, seed_state=None):
    """SingleRandomWalk_M.walk(v): returns (heap keys, heap float32 values, steps, new_seed_state)."""
    hk = np.zeros(capacity + 1, dtype=np.int32)
    hv = np.zeros(capacity + 1, dtype=np.float32)
    st = ctypes.c_uint64(java_seed(0) if seed_state is None else seed_state)
    n = ctypes.c_int32()
    steps = lib().or_single_random_walk_cache(g["V"], _p(g["row_ptr"], ctypes.c_int64), _p(g["col"], ctypes.c_int32),
                                              int(v), int(sample), int(step), float(C), int(capacity),
                                              ctypes.byref(st), _p(hk, ctypes.c_int32), _p(hv, ctypes.c_float),
                                              ctypes.byref(n))
    return hk[1:n.value + 1].copy(), hv[1:n.value + 1].copy(), int(steps), int(st.value)


def topsim_cache(g, v, sample, step, capacity, C=C_DEFAULT, seed_state=None, max_paths=1 << 22):
    """TopSim_singleSample_M.walk(v): returns (heap keys, heap float32 values, paths made, new_seed_state)."""
    hk = np.zeros(capacity + 1, dtype=np.int32)
    hv = np.zeros(capacity + 1, dtype=np.float32)
    st = ctypes.c_uint64(java_seed(0) if seed_state is None else seed_state)
    n = ctypes.c_int32()
    made = lib().or_topsim_cache(g["V"], _p(g["row_ptr"], ctypes.c_int64), _p(g["col"], ctypes.c_int32), int(v),
                                 int(sample), int(step), float(C), int(capacity), int(max_paths), ctypes.byref(st),
                                 _p(hk, ctypes.c_int32), _p(hv, ctypes.c_float), ctypes.byref(n))
    if made < 0:
        raise MemoryError("path tree exceeded max_paths")
    return hk[1:n.value + 1].copy(), hv[1:n.value + 1].copy(), int(made), int(st.value)


def print_by_order_cache(caches, out_path, topk=TOPK):
    """Print.printByOrder(FixedCacheMap[], outPath, topk) (utils/Print.java:94-123): per vertex the LAST topk
    entries of the ascending iteration, `%.6f` of the float value."""
    with open(out_path, "w", newline="") as out, open(out_path + ".sim.txt", "w", newline="") as outsim:
        for v, (hk, hv) in enumerate(caches):
            ks, vs = fcm_drain(hk, hv)
            out.write(str(v)); outsim.write(str(v))
            for key, val in list(zip(ks.tolist(), vs.tolist()))[max(0, len(ks) - topk):]:
                out.write(SEPARATOR + str(key))
                outsim.write(SEPARATOR + str(key) + SEPARATOR_KV + java_fmt(val, 6))
            out.write("\r\n"); outsim.write("\r\n")


# ---------------- simrank/DoubleRandomWalk.java ----------------
def double_walk_paths(g, sample, step, seed_state):
    """samplePaths(): int32 [V, sample, step]; returns (paths, new_seed_state)."""
    paths = np.zeros((g["V"], sample, step), dtype=np.int32)
    st = ctypes.c_uint64(seed_state)
    lib().or_double_walk_paths(g["V"], _p(g["row_ptr"], ctypes.c_int64), _p(g["col"], ctypes.c_int32), int(sample),
                               int(step), ctypes.byref(st), _p(paths, ctypes.c_int32))
    return paths, int(st.value)


def double_walk_matrix(paths, C=C_DEFAULT):
    V, sample, step = paths.shape
    sim = np.zeros((V, V), dtype=np.float64)
    lib().or_double_walk_matrix(V, _p(np.ascontiguousarray(paths), ctypes.c_int32), sample, step, float(C),
                                _p(sim, ctypes.c_double))
    return sim


# ---------------- simrank/TopSim_doubleSample.java, simrank/TopSim_Dev.java ----------------
def mass_tree(g, src, weight0, step, seed_state, max_paths=1 << 22):
    """sample(src) + computePath: mass[V, step+1] (-1 = unset), new rng state."""
    mass = np.full((g["V"], step + 1), -1.0, dtype=np.float64)
    st = ctypes.c_uint64(seed_state)
    made = lib().or_mass_tree(g["V"], _p(g["row_ptr"], ctypes.c_int64), _p(g["col"], ctypes.c_int32), int(src),
                              float(weight0), int(step), int(max_paths), ctypes.byref(st), _p(mass, ctypes.c_double))
    if made < 0:
        raise MemoryError("path tree exceeded max_paths")
    return mass, int(st.value)


def mass_sim(ma, mb, C=C_DEFAULT):
    V, s1 = ma.shape
    return float(lib().or_mass_sim(_p(np.ascontiguousarray(ma), ctypes.c_double), _p(np.ascontiguousarray(mb), ctypes.c_double),
                                   V, s1 - 1, float(C)))


def topsim_double_sample(g, sample, step, seed_state, C=C_DEFAULT):
    """TopSim_doubleSample.compute(): samplePaths over all vertices in order (one RNG stream), then getSim for
    i < j, mirrored; the diagonal stays 0.  Returns (sim, masses [V, V, step+1], new rng state)."""
    V = g["V"]
    masses = np.empty((V, V, step + 1), dtype=np.float64)
    st = seed_state
    for v in range(V):
        masses[v], st = mass_tree(g, v, float(sample), step, st)
    sim = np.zeros((V, V), dtype=np.float64)
    for i in range(V):
        for j in range(i + 1, V):
            sim[i, j] = sim[j, i] = mass_sim(masses[i], masses[j], C)
    return sim, masses, st


def fixedmaxpq_topk_min(row, k, min_value):
    row = np.ascontiguousarray(row, dtype=np.float64)
    ids = np.zeros(max(k, 1), dtype=np.int32)
    vals = np.zeros(max(k, 1), dtype=np.float64)
    n = lib().or_fixedmaxpq_topk_min(_p(row, ctypes.c_double), len(row), int(k), float(min_value),
                                     _p(ids, ctypes.c_int32), _p(vals, ctypes.c_double))
    return ids[:n].copy(), vals[:n].copy()


def topsim_dev_sample_count(sample, step, topK, singleStep):
    """TopSim_Dev constructor (:35-36)."""
    return int(((step - singleStep) * sample * 2.0) / (float(step) * (topK + 1.0)))


def topsim_dev(g, candidate, sample, step, topK, singleStep, seed_state, C=C_DEFAULT, rows=None):
    """TopSim_Dev.compute(candidate) (:57-98): per vertex i one tree from i, then for each of the top `topK`
    candidates j (candidate[i][j] >= MIN, FixedMaxPQ order) one tree from j and sim[i][j] = getSim; one RNG stream."""
    V = g["V"]
    S_ = topsim_dev_sample_count(sample, step, topK, singleStep)
    sim = np.zeros((V, V), dtype=np.float64)
    st = seed_state
    for i in (range(V) if rows is None else rows):
        m0, st = mass_tree(g, i, float(S_), step, st)
        ids, _ = fixedmaxpq_topk_min(candidate[i], topK, MIN)
        for j in ids.tolist():
            m1, st = mass_tree(g, j, float(S_), step, st)
            sim[i, j] = mass_sim(m0, m1, C)
        sim[i, i] = 0
    return sim, st


def simrank_exact_naive(g, C, iters):
    """SimRank.compute, scalar loops (small graphs)."""
    V = g["V"]
    sim = np.zeros((V, V), dtype=np.float64)
    lib().or_simrank_exact(V, _p(g["row_ptr"], ctypes.c_int64), _p(g["col"], ctypes.c_int32),
                           float(C), int(iters), _p(sim, ctypes.c_double))
    return sim


def simrank_exact_matrix(g, C, iters):
    """Same iteration as array algebra: S <- C * P S P^T off the diagonal, diag = 1, then
    diag := 0 (SimRank.java:36-65); P row-stochastic over the multigraph adjacency."""
    import scipy.sparse as sp
    V = g["V"]
    deg = np.diff(g["row_ptr"]).astype(np.float64)
    rows = np.repeat(np.arange(V), np.diff(g["row_ptr"]))
    inv = np.zeros(V)
    inv[deg > 0] = 1.0 / deg[deg > 0]
    P = sp.csr_matrix((inv[rows], (rows, g["col"])), shape=(V, V))
    S = np.eye(V)
    for _ in range(iters):
        S = C * (P @ (P @ S).T).T
        np.fill_diagonal(S, 1.0)
    np.fill_diagonal(S, 0.0)
    return S


def fixedmaxpq_topk(row, k=TOPK):
    """FixedMaxPQ offers of a dense row + sortedElement (Print.java:31-41)."""
    row = np.ascontiguousarray(row, dtype=np.float64)
    ids = np.zeros(max(k, 1), dtype=np.int32)
    vals = np.zeros(max(k, 1), dtype=np.float64)
    n = lib().or_fixedmaxpq_topk(_p(row, ctypes.c_double), len(row), int(k),
                                 _p(ids, ctypes.c_int32), _p(vals, ctypes.c_double))
    return ids[:n].copy(), vals[:n].copy()


def java_random_ints(seed, bound, n):
    out = np.zeros(n, dtype=np.int32)
    lib().jr_fill(int(seed), int(bound), int(n), _p(out, ctypes.c_int32))
    return out


def node2vec_walks_c(row_ptr, col, p, q, walk_length, starts, seed=1):
    """C-speed free-running node2vec port (CPU baseline only)."""
    row_ptr = np.ascontiguousarray(row_ptr, dtype=np.int64)
    col = np.ascontiguousarray(col, dtype=np.int32)
    starts = np.ascontiguousarray(starts, dtype=np.int64)
    out = np.empty((len(starts), walk_length), dtype=np.int32)
    steps = lib().or_node2vec_walks(len(row_ptr) - 1, _p(row_ptr, ctypes.c_int64),
                                    _p(col, ctypes.c_int32), float(p), float(q),
                                    int(walk_length), _p(starts, ctypes.c_int64), len(starts),
                                    int(seed), _p(out, ctypes.c_int32))
    return out, int(steps)


# ---------------- utils/Print.java, utils/Eval.java ----------------
def java_fmt(x, nd):
    """String.format("%.<nd>f"): java.util.Formatter rounds HALF_UP on the shortest-repr decimal
    digits of the double (FormattedFloatingDecimal.applyPrecision), not on its exact binary
    value: 0.0000005 -> "0.000001" where C printf gives "0.000000"."""
    from decimal import Decimal, ROUND_HALF_UP
    return format(Decimal(repr(float(x))).quantize(Decimal(1).scaleb(-nd), rounding=ROUND_HALF_UP), "f")


def print_by_order(sim, out_path, topk=TOPK, digits=6):
    """Print.printByOrder (:25-53, digits=6) / printByOrderAll (:55-84, digits=7)."""
    with open(out_path, "w", newline="") as out, open(out_path + ".sim.txt", "w", newline="") as outsim:
        for v in range(sim.shape[0]):
            ids, vals = fixedmaxpq_topk(sim[v], topk)
            out.write(str(v))
            outsim.write(str(v))
            for i, x in zip(ids, vals):
                out.write(SEPARATOR + str(int(i)))
                outsim.write(SEPARATOR + str(int(i)) + SEPARATOR_KV + java_fmt(x, digits))
            out.write("\r\n")
            outsim.write("\r\n")


def read_sim_file(path, separator=SEPARATOR):
    """-> list of (v, [(id, score), ...]) as Eval.precision tokenises (:93-111)."""
    rows = []
    with open(path, "r", newline="") as f:
        for line in f.read().split("\n"):
            line = line.rstrip("\r")
            if not line:
                continue
            tok = [t for t in line.split(separator) if t != ""] if separator == " " else line.split(separator)
            rows.append((int(tok[0]), [(int(t.split(SEPARATOR_KV)[0]), float(t.split(SEPARATOR_KV)[1]))
                                       for t in tok[1:]]))
    return rows


def precision_rows(gold_rows, out_rows, topk=TOPK):
    """Eval.precision (:81-131) on parsed rows: per-vertex |gold ∩ out| / min(TOPK, |gold|) over
    ids whose score >= MIN; 1.0 when gold is empty.  Returns (mean, per-vertex list)."""
    pres = []
    for (v1, r1), (v2, r2) in zip(gold_rows, out_rows):
        if v1 != v2:
            continue
        s1 = {i for i, x in r1 if x >= MIN}
        s2 = {i for i, x in r2 if x >= MIN}
        real_k = min(topk, len(s1))
        pres.append(1.0 if real_k == 0 else len(s1 & s2) / real_k)
    return (sum(pres) / len(pres) if pres else 0.0), pres
