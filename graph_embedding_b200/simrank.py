"""Python host mirror of the reference's Java TopSim surface (DeepSim/TopSimAll/src), backed by
libgraphwalk.  Class and method names follow the Java so the reference's drivers
(benchmark/Test_u_u_SingleRandomWalk_Sample.java:21-68) translate line by line:

    g   = Graph(path, V)                              # structures/Graph.java:28
    srw = SingleRandomWalk(g, sample, step)           # simrank/SingleRandomWalk.java:28
    srw.compute(); sim = srw.getResult()              # :39-45 (dense V x V, small graphs)
    ids, scores = srw.topk(k)                         # per-query top-k straight from the device
    Print.printByOrder(sim, outPath, TOPK, k)         # utils/Print.java:25-53
    Eval.precision(gold, outPath + ".sim.txt", prePath, k)   # utils/Eval.java:81-131

The Java binding of the same C ABI (Panama FFM) is in graph_embedding_b200/java/.
"""
import gzip
from decimal import Decimal, ROUND_HALF_UP

import numpy as np

from . import _lib


class MyConfiguration:
    """conf/MyConfiguration.java:16-22 (static mutable, as in the reference)."""
    SEPARATOR = ","
    SEPARATOR_KV = ":"
    TOPK = 20
    MIN = 0.000000001
    C = 0.6
    testTopK = [20]


class Graph:
    """structures/Graph.java: undirected unweighted multigraph, V vertex slots, both directions
    appended per line, duplicates and file order kept."""

    def __init__(self, graphPath, V, separator=None):
        sep = MyConfiguration.SEPARATOR if separator is None else separator
        if str(graphPath).endswith(".gz"):
            src, dst = [], []
            with gzip.open(graphPath, "rt") as f:
                for line in f:
                    line = line.rstrip("\r\n")
                    if line:
                        ids = line.split(sep)
                        src.append(int(ids[0])); dst.append(int(ids[1]))
            self.handle = _lib.GraphHandle.from_edges(src, dst, None, directed=False, mode=_lib.GW_MODE_MULTI,
                                                      n_slots=V)
        else:
            self.handle = _lib.GraphHandle.from_file(graphPath, delimiter=sep, weighted=False, directed=False,
                                                     mode=_lib.GW_MODE_MULTI, n_slots=V)
        self.vCount = self.handle.n
        self.eCount = self.handle.nnz // 2
        self._csr = None

    @classmethod
    def from_handle(cls, handle):
        g = cls.__new__(cls)
        g.handle = handle
        g.vCount = handle.n
        g.eCount = handle.nnz // 2
        g._csr = None
        return g

    def _c(self):
        if self._csr is None:
            self._csr = self.handle.csr(weights=False, node_ids=False, first_seen=False)
        return self._csr

    def degree(self, v):
        c = self._c()
        return int(c["row_ptr"][v + 1] - c["row_ptr"][v])

    def neighbors(self, v):
        c = self._c()
        return c["col_idx"][c["row_ptr"][v]:c["row_ptr"][v + 1]].tolist()

    def getVCount(self):
        return self.vCount

    def getECount(self):
        return self.eCount


_JR_MULT, _JR_ADD, _JR_MASK = 0x5DEECE66D, 0xB, (1 << 48) - 1


def _jr_jump(state, n):
    """java.util.Random state after n calls of next(): s -> a^n s + c (a^n - 1)/(a - 1)  (mod 2^48)."""
    a, c, acc_a, acc_c = _JR_MULT, _JR_ADD, 1, 0
    while n:
        if n & 1:
            acc_a, acc_c = (acc_a * a) & _JR_MASK, (acc_c * a + c) & _JR_MASK
        a, c = (a * a) & _JR_MASK, (c * a + c) & _JR_MASK
        n >>= 1
    return (acc_a * state + acc_c) & _JR_MASK


class SingleRandomWalk:
    """simrank/SingleRandomWalk.java: pure Monte-Carlo single-walk estimator (scores / SAMPLE)."""
    MODE = _lib.GW_SIMRANK_MC
    SAMPLE = 10000
    JAVA_CHUNK = 4096            # queries replayed per launch in java_seed mode

    def __init__(self, g, sample, step, seed=None, java_seed=None):
        """seed: Philox key of the production kernels.  java_seed: replay mode -- the walks are drawn from
        java.util.Random(java_seed) exactly as a JVM whose `Graph.rand` (structures/Graph.java:17) was
        seeded that way would draw them, one stream shared by all queries in order."""
        self.topk_k = MyConfiguration.TOPK
        self.STEP = step
        self.SAMPLE = sample
        self.g = g
        self.COUNT = g.getVCount()
        self.sim = None
        self.seed = int(np.random.randint(0, 2 ** 31 - 1)) if seed is None else int(seed)
        self.java_state = None if java_seed is None else (int(java_seed) ^ _JR_MULT) & _JR_MASK

    def compute(self, queries=None):
        """compute() (:39-45): every vertex 0..COUNT-1 is a query; dense result like double[][]."""
        q = np.arange(self.COUNT, dtype=np.int64) if queries is None else np.asarray(queries, dtype=np.int64)
        self._queries = q
        if self.java_state is not None and self.MODE == _lib.GW_SIMRANK_MC:
            self.sim = self._compute_java_stream(q)
            return self
        self.sim = self.g.handle.simrank_rows(q, MyConfiguration.C, self.STEP, self.SAMPLE, self.MODE, self.seed)
        return self

    def _chain_java_stream(self, q, per_query, run):
        """All queries share ONE sequential java.util.Random stream.  A query normally consumes exactly `per_query`
        draws, so the state in front of every query is predicted by an LCG jump and all queries replay in parallel;
        where the prediction fails (nextInt's rejection loop fired, or the vertex is isolated and drew nothing) the
        tail is replayed again from the true state.  run(queries, states) -> (per-query results, states after)."""
        out = [None] * len(q)
        state, lo = self.java_state, 0
        while lo < len(q):
            hi = min(len(q), lo + self.JAVA_CHUNK)
            states = [state]
            for _ in range(lo + 1, hi):
                states.append(_jr_jump(states[-1], per_query))
            res, after = run(q[lo:hi], states)
            after = [int(x) for x in after]
            good = 1                                             # res[0] started from a true state
            while good < len(states) and after[good - 1] == states[good]:
                good += 1
            out[lo:lo + good] = list(res[:good])
            state = after[good - 1]
            lo += good
        self.java_state = state
        return out

    def _compute_java_stream(self, q):
        rows = self._chain_java_stream(
            q, self.SAMPLE * 2 * self.STEP,
            lambda qs, st: self.g.handle.simrank_rows_javarng(qs, MyConfiguration.C, self.STEP, self.SAMPLE, st))
        return np.asarray(rows, dtype=np.float64).reshape(len(q), self.COUNT)

    def getResult(self):
        return self.sim

    def topk(self, k=None, queries=None):
        """Fused walk-and-meet + per-query top-k on the device (no V x V matrix)."""
        k = self.topk_k if k is None else k
        q = np.arange(self.COUNT, dtype=np.int64) if queries is None else np.asarray(queries, dtype=np.int64)
        return self.g.handle.simrank_topk(q, MyConfiguration.C, self.STEP, self.SAMPLE, k, self.MODE, self.seed)


class TopSim_singleSample(SingleRandomWalk):
    """simrank/TopSim_singleSample.java: hybrid enumerate-while-weight>=degree else sample
    (scores x SAMPLE, unnormalised, as the reference :189).  java_seed: replay mode -- the reference's queue
    order with java.util.Random(java_seed), one stream over all queries (sequential, parity path)."""
    MODE = _lib.GW_SIMRANK_HYBRID

    def compute(self, queries=None):
        if self.java_state is None:
            return super().compute(queries)
        q = np.arange(self.COUNT, dtype=np.int64) if queries is None else np.asarray(queries, dtype=np.int64)
        self._queries = q
        out = np.zeros((len(q), self.COUNT), dtype=np.float64)
        for i, v in enumerate(q.tolist()):                       # draws per query vary: the stream is chained query by query
            row, after = self.g.handle.topsim_rows_javarng([v], MyConfiguration.C, self.STEP, self.SAMPLE, [self.java_state], mode=0)
            out[i] = row[0]
            self.java_state = int(after[0])
        self.sim = out
        return self


class TopSim_Enumerate(SingleRandomWalk):
    """simrank/TopSim_Enumerate.java: the path tree with EVERY path split into all neighbours (deterministic;
    equals SAMPLE x SimRank truncated at STEP sweeps).  The reference runs it for vertex 0 only (:47); level
    sizes are products of degrees, so `max_paths` bounds the queue (MemoryError beyond)."""

    def __init__(self, g, sample, step, max_paths=1 << 22):
        super().__init__(g, sample, step, seed=0)
        self.max_paths = max_paths

    def compute(self, queries=None):
        q = np.array([0], dtype=np.int64) if queries is None else np.asarray(queries, dtype=np.int64)
        self._queries = q
        self.sim, _ = self.g.handle.topsim_rows_javarng(q, MyConfiguration.C, self.STEP, self.SAMPLE, [0] * len(q), mode=1,
                                                        max_paths=self.max_paths)
        return self

    def topk(self, k=None, queries=None):
        raise NotImplementedError("TopSim_Enumerate is the deterministic parity path; use SimRank or SingleRandomWalk.topk")


class FixedCacheMap:
    """lxctools/FixedCacheMap.java: bounded key -> float cache that evicts the entry with the MINIMUM value; 1-based
    binary min-heap (keys/values) plus key -> slot map; float32 arithmetic as Java's `float`.  Iterating drains the
    heap with delMin (:102-132): ascending by value, destructive -- exactly what Print.printByOrder consumes."""

    def __init__(self, NMAX, keys=None, values=None):
        self.NMAX = int(NMAX)
        self.keys = [0] + ([] if keys is None else [int(k) for k in keys])
        self.values = [np.float32(0)] + ([] if values is None else [np.float32(x) for x in values])
        self.key2Index = {k: i for i, k in enumerate(self.keys) if i > 0}

    def size(self):
        return len(self.keys) - 1

    def isEmpty(self):
        return self.size() == 0

    def _exch(self, a, b):
        k, v, m = self.keys, self.values, self.key2Index
        m[k[a]] = b
        m[k[b]] = a
        k[a], k[b] = k[b], k[a]
        v[a], v[b] = v[b], v[a]

    def _sink(self, i):
        n, v = self.size(), self.values
        while 2 * i <= n:
            j = 2 * i
            if j < n and v[j] > v[j + 1]:
                j += 1
            if not v[i] > v[j]:
                break
            self._exch(i, j)
            i = j

    def _swim(self, i):
        v = self.values
        while i > 1 and v[i // 2] > v[i]:
            self._exch(i, i // 2)
            i //= 2

    def put(self, key, value):
        value = np.float32(value)
        idx = self.key2Index.get(key)
        if idx is not None:
            self.values[idx] = np.float32(self.values[idx] + value)
            self._sink(idx)
        elif self.size() < self.NMAX:
            self.keys.append(int(key)); self.values.append(value)
            self.key2Index[key] = self.size()
            self._swim(self.size())
        elif value > self.values[1]:
            del self.key2Index[self.keys[1]]
            self.keys[1], self.values[1] = int(key), value
            self.key2Index[key] = 1
            self._sink(1)

    def __iter__(self):
        while self.size() > 0:
            kv = (self.keys[1], self.values[1])
            self.key2Index.pop(self.keys[1], None)
            self._exch(1, self.size())
            self.keys.pop(); self.values.pop()
            self._sink(1)
            yield kv


class SingleRandomWalk_M(SingleRandomWalk):
    """simrank/SingleRandomWalk_M.java: the walks of SingleRandomWalk, similarities kept per vertex in a
    FixedCacheMap(TOPK * M) as float increments; STEP is the class constant 5 (:19).  getResult() -> list of
    FixedCacheMap.  java_seed: replay mode, bit-exact with a JVM whose Graph.rand was seeded that way (evictions
    included).  Without it the production kernel runs: the device accumulates every target exactly (nothing is
    evicted), and each cache receives the top min(capacity, 128) entries of that result."""
    CACHE_MODE = 0

    def __init__(self, g, M, sample, seed=None, java_seed=None):
        super().__init__(g, sample, 5, seed=seed, java_seed=java_seed)
        self.capacity = self.topk_k * M

    def _draws_per_query(self):
        return self.SAMPLE * 2 * self.STEP

    def compute(self, queries=None):
        q = np.arange(self.COUNT, dtype=np.int64) if queries is None else np.asarray(queries, dtype=np.int64)
        self._queries = q
        if self.java_state is not None:
            heaps = self._chain_java_stream(
                q, self._draws_per_query(),
                lambda qs, st: self.g.handle.simrank_cache_javarng(qs, MyConfiguration.C, self.STEP, self.SAMPLE, self.capacity,
                                                                   st, mode=self.CACHE_MODE))
            self.sim = [FixedCacheMap(self.capacity, k, v) for k, v in heaps]
            return self
        k = min(self.capacity, 128)
        ids, sc = self.g.handle.simrank_topk(q, MyConfiguration.C, self.STEP, self.SAMPLE, k, self.MODE, self.seed)
        if self.MODE == _lib.GW_SIMRANK_HYBRID:
            sc = sc / self.SAMPLE                                # TopSim_singleSample_M.java:224 divides by SAMPLE
        self.sim = []
        for r in range(len(q)):
            m = FixedCacheMap(self.capacity)
            for c in range(k - 1, -1, -1):                       # ascending score
                if ids[r, c] >= 0:
                    m.put(int(ids[r, c]), sc[r, c])
            self.sim.append(m)
        return self


class TopSim_singleSample_M(SingleRandomWalk_M):
    """simrank/TopSim_singleSample_M.java: the path tree of TopSim_singleSample with FixedCacheMap accumulation
    (increments / SAMPLE, float).  java_seed replays the reference's queue order; the draws per query vary, so the
    stream is chained query by query."""
    MODE = _lib.GW_SIMRANK_HYBRID
    CACHE_MODE = 1
    JAVA_CHUNK = 1

    def _draws_per_query(self):
        return 0


class DoubleRandomWalk:
    """simrank/DoubleRandomWalk.java: SAMPLE walks of STEP steps from every vertex (samplePaths :50-65), pair score =
    sum over all SAMPLE^2 path pairs of C^(first common position + 1) / SAMPLE^2 (getSim :77-91).
    java_seed: replay mode -- paths from one java.util.Random(java_seed) stream over the vertices in order and fp64
    adds in the reference's order (bit-exact).  Otherwise Philox paths and integer first-meeting counts."""
    SAMPLE = 200
    JAVA_CHUNK = 4096

    def __init__(self, g, sample, step, seed=None, java_seed=None):
        self.SAMPLE = sample
        self.STEP = step
        self.g = g
        self.COUNT = g.getVCount()
        self.paths = None
        self.sim = None
        self.seed = int(np.random.randint(0, 2 ** 31 - 1)) if seed is None else int(seed)
        self.java_state = None if java_seed is None else (int(java_seed) ^ _JR_MULT) & _JR_MASK

    def samplePaths(self):
        verts = np.arange(self.COUNT, dtype=np.int64)
        if self.java_state is None:
            self.paths = self.g.handle.double_walk_paths(verts, self.SAMPLE, self.STEP, seed=self.seed)
            return self
        per = SingleRandomWalk._chain_java_stream(
            self, verts, self.SAMPLE * self.STEP,
            lambda vs, st: self.g.handle.double_walk_paths(vs, self.SAMPLE, self.STEP, rng_states=st))
        self.paths = np.stack(per).astype(np.int32)
        return self

    def computeSims(self, rows=None):
        self.sim = self.g.handle.double_walk_sims(self.paths, MyConfiguration.C, rows=rows,
                                                  exact_order=self.java_state is not None)
        return self

    def compute(self):
        self.samplePaths()
        return self.computeSims()

    def getResult(self):
        return self.sim


class TopSim_doubleSample:
    """simrank/TopSim_doubleSample.java: one enumerate-or-sample path-mass tree of STEP levels per vertex (sample :66-151,
    computePath :153-178), pair score = sum over targets and levels of C^level * mass_i * mass_j (getSim :189-199;
    unnormalised, x SAMPLE^2).  java_seed: replay mode -- one java.util.Random stream over the vertices in order
    (draws per tree vary, so trees are chained one by one) and getSim's fp64 order; otherwise Philox trees, all in one
    launch, and the warp-parallel product kernel."""
    SAMPLE = 200

    def __init__(self, g, sample, step, seed=None, java_seed=None):
        self.SAMPLE = sample
        self.STEP = step
        self.g = g
        self.COUNT = g.getVCount()
        self.paths = None
        self.sim = None
        self.seed = int(np.random.randint(0, 2 ** 31 - 1)) if seed is None else int(seed)
        self.java_state = None if java_seed is None else (int(java_seed) ^ _JR_MULT) & _JR_MASK

    def samplePaths(self):
        verts = np.arange(self.COUNT, dtype=np.int64)
        if self.java_state is None:
            self.paths = self.g.handle.topsim_mass(verts, self.SAMPLE, self.STEP, seed=self.seed)
            return self
        self.paths = np.empty((self.COUNT, self.COUNT, self.STEP + 1), dtype=np.float64)
        for v in range(self.COUNT):
            m, after = self.g.handle.topsim_mass([v], self.SAMPLE, self.STEP, rng_states=[self.java_state])
            self.paths[v] = m[0]
            self.java_state = int(after[0])
        return self

    def computeSims(self):
        n = self.COUNT
        iu, ju = np.triu_indices(n, 1)
        vals = self.g.handle.topsim_mass_sims(self.paths, MyConfiguration.C, iu, ju, exact_order=self.java_state is not None)
        self.sim = np.zeros((n, n), dtype=np.float64)
        self.sim[iu, ju] = vals
        self.sim[ju, iu] = vals
        return self

    def compute(self):
        self.samplePaths()
        return self.computeSims()

    def getResult(self):
        return self.sim


class TopSim_Dev:
    """simrank/TopSim_Dev.java: two-stage refinement.  For every vertex i: one path-mass tree from i, then for each of
    the `topK` best candidates j of candidate[i] (>= MIN, FixedMaxPQ order, :72-83) one tree from j and
    sim[i][j] = getSim (:233-244).  The per-tree weight is the constructor's
    (int)((step - singleStep) * sample * 2.0 / (step * (topK + 1.0))) (:35).  java_seed: replay of a seeded JVM (one
    stream through i and its candidates in order); otherwise Philox trees in one launch."""

    def __init__(self, g, sample, step, topK, singleStep, seed=None, java_seed=None):
        self.SAMPLE = int(((step - singleStep) * sample * 2.0) / (float(step) * (topK + 1.0)))
        self.STEP = step
        self.singleK = topK
        self.g = g
        self.COUNT = g.getVCount()
        self.sim = None
        self.seed = int(np.random.randint(0, 2 ** 31 - 1)) if seed is None else int(seed)
        self.java_state = None if java_seed is None else (int(java_seed) ^ _JR_MULT) & _JR_MASK

    def _candidates(self, row):
        pq = FixedMaxPQ(self.singleK)
        for j in np.nonzero(np.asarray(row) >= MyConfiguration.MIN)[0].tolist():
            pq.offer(j, float(row[j]))
        return [k for k, _ in pq.sortedElement()]

    def compute(self, candidate, rows=None):
        n = self.COUNT
        rows = range(n) if rows is None else list(rows)
        self.sim = np.zeros((n, n), dtype=np.float64)
        cands = {i: self._candidates(candidate[i]) for i in rows}
        h = self.g.handle
        if self.java_state is not None:
            for i in rows:
                m0, after = h.topsim_mass([i], self.SAMPLE, self.STEP, rng_states=[self.java_state])
                self.java_state = int(after[0])
                for j in cands[i]:
                    m1, after = h.topsim_mass([j], self.SAMPLE, self.STEP, rng_states=[self.java_state])
                    self.java_state = int(after[0])
                    self.sim[i, j] = h.topsim_mass_sims(np.concatenate([m0, m1]), MyConfiguration.C, [0], [1], exact_order=True)[0]
                self.sim[i, i] = 0
            return self
        budget = max(1, (1 << 28) // (n * (self.STEP + 1) * 8))      # trees per launch (256 MB of mass rows)
        srcs, pa, pb, where = [], [], [], []
        def flush():
            if not srcs:
                return
            m = h.topsim_mass(srcs, self.SAMPLE, self.STEP, seed=self.seed, call_id_base=flush.base)
            flush.base += len(srcs)
            vals = h.topsim_mass_sims(m, MyConfiguration.C, pa, pb)
            for (i, j), x in zip(where, vals):
                self.sim[i, j] = x
            del srcs[:], pa[:], pb[:], where[:]
        flush.base = 0
        for i in rows:
            if len(srcs) + 1 + len(cands[i]) > budget:
                flush()
            a = len(srcs)
            srcs.append(i)
            for j in cands[i]:
                pa.append(a); pb.append(len(srcs)); where.append((i, j))
                srcs.append(j)
        flush()
        return self

    def getResult(self):
        return self.sim


class SimRank:
    """simrank/SimRank.java: naive exact SimRank, STEP Jacobi sweeps (STEP = 3 as committed)."""

    def __init__(self, g, step=3):
        self.g = g
        self.STEP = step
        self.COUNT = g.getVCount()
        self.sim = None

    def compute(self):
        self.sim = self.g.handle.simrank_exact(MyConfiguration.C, self.STEP)
        return self

    def getResult(self):
        return self.sim


# ---------------- lxctools/FixedMaxPQ.java + Pair.java ----------------
class FixedMaxPQ:
    """Size-k min-heap with java.util.PriorityQueue's exact sift rules; replaces the minimum only
    when the offer is strictly greater (FixedMaxPQ.java:30-39); sortedElement() is a stable
    descending sort of the heap array (:72-76)."""

    def __init__(self, capacity):
        self.capacity = capacity
        self.q = []

    def _up(self, k, x):
        q = self.q
        while k > 0:
            parent = (k - 1) >> 1
            if x[1] >= q[parent][1]:
                break
            q[k] = q[parent]
            k = parent
        q[k] = x

    def _down(self, k, x):
        q = self.q
        size = len(q)
        half = size >> 1
        while k < half:
            child = 2 * k + 1
            right = child + 1
            if right < size and q[child][1] > q[right][1]:
                child = right
            if x[1] <= q[child][1]:
                break
            q[k] = q[child]
            k = child
        q[k] = x

    def offer(self, key, value):
        e = (key, value)
        if len(self.q) < self.capacity:
            self.q.append(e)
            self._up(len(self.q) - 1, e)
        elif self.capacity > 0 and self.q[0][1] < value:
            last = self.q.pop()
            if self.q:
                self._down(0, last)
            self.q.append(e)
            self._up(len(self.q) - 1, e)

    def sortedElement(self):
        return sorted(self.q, key=lambda kv: -kv[1])          # python's sort is stable, like Collections.sort


def java_format(x, digits):
    """String.format("%.<digits>f"): java.util.Formatter rounds HALF_UP on the shortest decimal
    representation of the double."""
    return format(Decimal(repr(float(x))).quantize(Decimal(1).scaleb(-digits), rounding=ROUND_HALF_UP), "f")


def _row_topk_exact(row, topk):
    """FixedMaxPQ fed with every column of a dense row in ascending id (Print.java:31-37).  Once the heap is full an
    offer changes it only when it is strictly greater than the current minimum (FixedMaxPQ.java:33-37), so the offers
    that would be refused are skipped in bulk: the heap goes through exactly the states the reference's does."""
    pq = FixedMaxPQ(topk)
    row = np.asarray(row)
    n = len(row)
    head = min(topk, n)
    for i in range(head):
        pq.offer(i, float(row[i]))
    i = head
    while i < n and topk > 0:
        ahead = np.nonzero(row[i:] > pq.q[0][1])[0]
        if len(ahead) == 0:
            break
        j = i + int(ahead[0])
        pq.offer(j, float(row[j]))
        i = j + 1
    return pq.sortedElement()


class Print:
    @staticmethod
    def _write(sim, outPath, topk, digits):
        sep, kv = MyConfiguration.SEPARATOR, MyConfiguration.SEPARATOR_KV
        with open(outPath, "w", newline="") as out, open(outPath + ".sim.txt", "w", newline="") as outsim:
            for v in range(len(sim)):
                out.write(str(v)); outsim.write(str(v))
                for key, val in _row_topk_exact(np.asarray(sim[v]), topk):
                    out.write(sep + str(key))
                    outsim.write(sep + str(key) + kv + java_format(val, digits))
                out.write("\r\n"); outsim.write("\r\n")

    @staticmethod
    def printByOrder(sim, outPath, topk, testTopK=None):
        """utils/Print.java:25-53: `<v>,<id>,...` and `<v>,<id>:<%.6f>,...`, CRLF.  A list of FixedCacheMap selects
        the overload of :94-123: per vertex the LAST topk entries of the cache's ascending (destructive) iteration."""
        if len(sim) and isinstance(sim[0], FixedCacheMap):
            sep, kv = MyConfiguration.SEPARATOR, MyConfiguration.SEPARATOR_KV
            with open(outPath, "w", newline="") as out, open(outPath + ".sim.txt", "w", newline="") as outsim:
                for v in range(len(sim)):
                    size = sim[v].size()
                    out.write(str(v)); outsim.write(str(v))
                    for i, (key, val) in enumerate(sim[v]):
                        if i < size - topk:
                            continue
                        out.write(sep + str(key))
                        outsim.write(sep + str(key) + kv + java_format(val, 6))
                    out.write("\r\n"); outsim.write("\r\n")
            return
        Print._write(sim, outPath, topk, 6)

    @staticmethod
    def printByOrderAll(sim, outPath, topk, testTopK=None):
        """utils/Print.java:55-84 (%.7f)."""
        Print._write(sim, outPath, topk, 7)

    @staticmethod
    def printTopk(ids, scores, outPath, queries=None, digits=6):
        """Same files from the device top-k (ids/scores [nq, k]; id -1 = no further positive
        score): rows are padded with the lowest unused ids at score 0, which downstream readers
        drop (Eval.java:99-108 filters < MIN, DeepSim/src/main.py:99 filters <= 1e-8)."""
        sep, kv = MyConfiguration.SEPARATOR, MyConfiguration.SEPARATOR_KV
        nq, k = ids.shape
        with open(outPath, "w", newline="") as out, open(outPath + ".sim.txt", "w", newline="") as outsim:
            for r in range(nq):
                v = r if queries is None else int(queries[r])
                out.write(str(v)); outsim.write(str(v))
                used = set(int(x) for x in ids[r] if x >= 0)
                pad = 0
                for c in range(k):
                    key, val = int(ids[r, c]), float(scores[r, c])
                    if key < 0:
                        while pad in used:
                            pad += 1
                        key, val = pad, 0.0
                        used.add(pad)
                    out.write(sep + str(key))
                    outsim.write(sep + str(key) + kv + java_format(val, digits))
                out.write("\r\n"); outsim.write("\r\n")


class Eval:
    @staticmethod
    def precision(path1, path2, prePath, K):
        """utils/Eval.java:81-131: mean over vertices of |gold ∩ out| / min(TOPK, |gold|) on ids
        whose score >= MIN; writes `<v>,<precision>` lines; returns the mean as a string."""
        sep, kv = MyConfiguration.SEPARATOR, MyConfiguration.SEPARATOR_KV
        total, ssum, mn = 0, 0.0, float("inf")
        with open(prePath, "w", newline="") as out, open(path1, newline="") as f1, open(path2, newline="") as f2:
            for line1 in f1:
                line2 = f2.readline()
                t1 = line1.rstrip("\r\n").split(sep)
                t2 = line2.rstrip("\r\n").split(sep)
                if t1[0] != t2[0]:
                    print("error !" + t1[0] + "\t" + t2[0])
                    continue
                s1 = {t.split(kv)[0] for t in t1[1:] if float(t.split(kv)[1]) >= MyConfiguration.MIN}
                s2 = {t.split(kv)[0] for t in t2[1:] if float(t.split(kv)[1]) >= MyConfiguration.MIN}
                realK = min(MyConfiguration.TOPK, len(s1))
                pre = 1.0 if realK == 0 else 1.0 * len(s1 & s2) / realK
                ssum += pre
                out.write(t1[0] + sep + repr(pre) + "\r\n")
                total += 1
                mn = min(mn, pre)
        print("total nodes:" + str(total) + "\tavg precision: " + str(ssum / total) + "\tmin pre: " + str(mn))
        return str(ssum / total)


def read_simrank(path):
    """DeepSim/src/main.py:83-107: the consumer of `.sim.txt` (entries <= 1e-8 dropped)."""
    simrank = []
    with open(path) as f:
        for line in f.readlines():
            words = line.split(",")
            sim = []
            for i in range(1, len(words)):
                if i == len(words) - 1:
                    words[i] = words[i][:-1]
                ts = words[i].split(":")
                if float(ts[1]) <= 0.00000001:
                    continue
                sim.append((ts[0], ts[1]))
            simrank.append(sim)
    return simrank
