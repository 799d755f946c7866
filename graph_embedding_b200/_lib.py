"""ctypes binding of libgraphwalk.so (include/graphwalk.h).  No torch types cross this boundary.

The library is built in-tree by ``graph_embedding_b200/build.py`` (nvcc, sm_100a).  There is no
CPU fallback: if the shared object is missing and cannot be built, or no CUDA device is present,
compute calls raise ``GraphWalkError``.
"""
import ctypes
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libgraphwalk.so")

GW_OK, GW_E_INVALID, GW_E_CUDA, GW_E_IO, GW_E_TOO_LARGE, GW_E_STATE, GW_E_KEY = 0, -1, -2, -3, -4, -5, -6
GW_MODE_SIMPLE, GW_MODE_MULTI = 0, 1
GW_F_DIRECTED, GW_F_WEIGHTED, GW_F_MULTI = 1, 2, 4
GW_SIMRANK_MC, GW_SIMRANK_HYBRID, GW_SIMRANK_MC_F64 = 0, 1, 2


class GraphWalkError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("libgraphwalk error %d: %s" % (code, msg))
        self.code = code


c_i64p = ctypes.POINTER(ctypes.c_int64)
c_i32p = ctypes.POINTER(ctypes.c_int32)
c_f64p = ctypes.POINTER(ctypes.c_double)
c_vp = ctypes.c_void_p

# name -> (restype, argtypes); every symbol graphwalk.h declares
SIGNATURES = {
    "gw_version": (ctypes.c_int, []),
    "gw_last_error": (ctypes.c_char_p, []),
    "gw_device_count": (ctypes.c_int, [ctypes.POINTER(ctypes.c_int)]),
    "gw_set_device": (ctypes.c_int, [ctypes.c_int]),
    "gw_kernel_launches": (ctypes.c_int64, []),
    "gw_graph_from_edges": (ctypes.c_int, [c_i64p, c_i64p, c_f64p, ctypes.c_int64, ctypes.c_int, ctypes.c_int,
                                           ctypes.c_int64, ctypes.POINTER(c_vp)]),
    "gw_graph_load_edgelist": (ctypes.c_int, [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_int, ctypes.c_int,
                                              ctypes.c_int, ctypes.c_int64, ctypes.POINTER(c_vp)]),
    "gw_graph_rmat": (ctypes.c_int, [ctypes.c_int, ctypes.c_int64, ctypes.c_double, ctypes.c_double,
                                     ctypes.c_double, ctypes.c_uint64, ctypes.POINTER(c_vp)]),
    "gw_graph_barabasi_albert": (ctypes.c_int, [ctypes.c_int64, ctypes.c_int, ctypes.c_uint64, ctypes.POINTER(c_vp)]),
    "gw_graph_free": (ctypes.c_int, [c_vp]),
    "gw_graph_info": (ctypes.c_int, [c_vp, c_i64p, c_i64p, c_i32p, c_i32p, c_i32p]),
    "gw_graph_csr": (ctypes.c_int, [c_vp, c_i64p, c_i32p, c_f64p, c_i64p, c_i64p]),
    "gw_graph_device_views": (ctypes.c_int, [c_vp, ctypes.POINTER(c_vp), ctypes.POINTER(c_vp)]),
    "gw_graph_nonisolated": (ctypes.c_int, [c_vp, c_i64p, c_i64p]),
    "gw_alias_setup": (ctypes.c_int, [c_f64p, ctypes.c_int64, c_i32p, c_f64p]),
    "gw_alias_nodes": (ctypes.c_int, [c_vp, c_i32p, c_f64p]),
    "gw_alias_edges_size": (ctypes.c_int, [c_vp, c_i64p]),
    "gw_alias_edges": (ctypes.c_int, [c_vp, ctypes.c_double, ctypes.c_double, ctypes.c_int64, c_i64p, c_i32p, c_f64p]),
    "gw_node2vec_walks": (ctypes.c_int, [c_vp, ctypes.c_double, ctypes.c_double, ctypes.c_int32, c_i64p,
                                         ctypes.c_int64, ctypes.c_uint64, ctypes.c_uint64, c_i32p, c_i32p]),
    "gw_node2vec_walks_dev": (ctypes.c_int, [c_vp, ctypes.c_double, ctypes.c_double, ctypes.c_int32, c_vp,
                                             ctypes.c_int64, ctypes.c_uint64, ctypes.c_uint64, c_vp, c_vp, c_vp]),
    "gw_graph_last_handoff": (ctypes.c_int, [c_vp, c_i32p, c_i32p]),
    "gw_corpus_unpack24": (ctypes.c_int, [c_vp, c_i32p, ctypes.c_int64, ctypes.c_int32, ctypes.c_int32, c_i32p]),
    "gw_py_random_shuffle": (ctypes.c_int, [ctypes.POINTER(ctypes.c_uint32), c_i32p, c_i64p, ctypes.c_int64]),
    "gw_graph_prepare_walks": (ctypes.c_int, [c_vp, c_f64p]),
    "gw_graph_common_counts": (ctypes.c_int, [c_vp, c_i32p, c_i32p]),
    "gw_node2vec_walks_replay": (ctypes.c_int, [c_vp, ctypes.c_int32, c_i64p, ctypes.c_int64, c_f64p,
                                                ctypes.c_int64, c_i64p, c_i32p, c_i32p]),
    "gw_node2vec_walk_traffic_dev": (ctypes.c_int, [c_vp, ctypes.c_double, ctypes.c_double, ctypes.c_int32, c_vp,
                                                    ctypes.c_int64, ctypes.c_uint64, ctypes.c_uint64, c_i64p, c_vp]),
    "gw_walks_byte_model_dev": (ctypes.c_int, [c_vp, c_vp, ctypes.c_int64, ctypes.c_int32, ctypes.c_int, c_i64p,
                                               c_i64p, c_vp]),
    "gw_simrank_topk": (ctypes.c_int, [c_vp, c_i64p, ctypes.c_int64, ctypes.c_double, ctypes.c_int32,
                                       ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, ctypes.c_uint64,
                                       ctypes.c_uint64, c_i32p, c_f64p]),
    "gw_simrank_topk_dev": (ctypes.c_int, [c_vp, c_vp, ctypes.c_int64, ctypes.c_double, ctypes.c_int32,
                                           ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, ctypes.c_uint64,
                                           ctypes.c_uint64, c_vp, c_vp, c_vp]),
    "gw_simrank_rows": (ctypes.c_int, [c_vp, c_i64p, ctypes.c_int64, ctypes.c_double, ctypes.c_int32,
                                       ctypes.c_int32, ctypes.c_int32, ctypes.c_uint64, ctypes.c_uint64, c_f64p]),
    "gw_simrank_rows_javarng": (ctypes.c_int, [c_vp, c_i64p, ctypes.c_int64, ctypes.c_double, ctypes.c_int32,
                                               ctypes.c_int32, ctypes.POINTER(ctypes.c_uint64), c_f64p]),
    "gw_topsim_rows_javarng": (ctypes.c_int, [c_vp, c_i64p, ctypes.c_int64, ctypes.c_double, ctypes.c_int32,
                                              ctypes.c_int32, ctypes.c_int32, ctypes.c_int64,
                                              ctypes.POINTER(ctypes.c_uint64), c_f64p]),
    "gw_simrank_cache_javarng": (ctypes.c_int, [c_vp, c_i64p, ctypes.c_int64, ctypes.c_double, ctypes.c_int32,
                                                ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, ctypes.c_int64,
                                                ctypes.POINTER(ctypes.c_uint64), c_i32p,
                                                ctypes.POINTER(ctypes.c_float), c_i32p]),
    "gw_double_walk_paths": (ctypes.c_int, [c_vp, c_i64p, ctypes.c_int64, ctypes.c_int32, ctypes.c_int32,
                                            ctypes.c_uint64, ctypes.POINTER(ctypes.c_uint64), c_i32p]),
    "gw_double_walk_sims": (ctypes.c_int, [c_vp, c_i32p, ctypes.c_int64, ctypes.c_int32, ctypes.c_int32,
                                           ctypes.c_double, c_i64p, ctypes.c_int64, ctypes.c_int32, c_f64p]),
    "gw_topsim_mass": (ctypes.c_int, [c_vp, c_i64p, ctypes.c_int64, ctypes.c_double, ctypes.c_int32, ctypes.c_int64,
                                      ctypes.c_uint64, ctypes.c_uint64, ctypes.POINTER(ctypes.c_uint64), c_f64p]),
    "gw_topsim_mass_sims": (ctypes.c_int, [c_vp, c_f64p, ctypes.c_int64, ctypes.c_int32, ctypes.c_double, c_i64p,
                                           c_i64p, ctypes.c_int64, ctypes.c_int32, c_f64p]),
    "gw_sgns_create": (ctypes.c_int, [ctypes.c_int64, ctypes.c_int32, ctypes.c_uint64, ctypes.c_int32, ctypes.POINTER(c_vp)]),
    "gw_sgns_free": (ctypes.c_int, [c_vp]),
    "gw_sgns_count_dev": (ctypes.c_int, [c_vp, c_vp, ctypes.c_int64, ctypes.c_int32, c_vp]),
    "gw_sgns_finalize_vocab": (ctypes.c_int, [c_vp, ctypes.c_double, ctypes.c_int32]),
    "gw_sgns_train_dev": (ctypes.c_int, [c_vp, c_vp, ctypes.c_int64, ctypes.c_int32, ctypes.c_int32, ctypes.c_double,
                                         ctypes.c_double, ctypes.c_double, ctypes.c_double, ctypes.c_uint64, ctypes.c_int32,
                                         ctypes.c_int32, c_vp]),
    "gw_sgns_info": (ctypes.c_int, [c_vp, c_i64p, c_i32p, c_f64p, c_i64p]),
    "gw_sgns_vectors": (ctypes.c_int, [c_vp, ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_float), c_i64p]),
    "gw_sgns_set_vectors": (ctypes.c_int, [c_vp, ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_float)]),
    "gw_node2vec_embeddings": (ctypes.c_int, [c_vp, ctypes.c_double, ctypes.c_double, ctypes.c_int32, ctypes.c_int32, c_i64p,
                                              ctypes.c_int64, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32,
                                              ctypes.c_double, ctypes.c_double, ctypes.c_double, ctypes.c_uint64,
                                              ctypes.POINTER(ctypes.c_float), c_i64p, c_f64p]),
    "gw_comm_unique_id": (ctypes.c_int, [c_vp]),
    "gw_comm_init": (ctypes.c_int, [ctypes.c_int32, ctypes.c_int32, c_vp, ctypes.c_int32, ctypes.POINTER(c_vp)]),
    "gw_comm_info": (ctypes.c_int, [c_vp, c_i32p, c_i32p, c_i32p]),
    "gw_comm_free": (ctypes.c_int, [c_vp]),
    "gw_shard_range": (ctypes.c_int, [ctypes.c_int64, ctypes.c_int32, ctypes.c_int32, c_i64p, c_i64p]),
    "gw_comm_last_times": (ctypes.c_int, [c_vp, c_f64p, c_f64p]),
    "gw_node2vec_walks_sharded": (ctypes.c_int, [c_vp, c_vp, ctypes.c_double, ctypes.c_double, ctypes.c_int32, c_i64p,
                                                 ctypes.c_int64, ctypes.c_uint64, ctypes.c_int32, c_i32p, c_i32p]),
    "gw_simrank_topk_sharded": (ctypes.c_int, [c_vp, c_vp, c_i64p, ctypes.c_int64, ctypes.c_double, ctypes.c_int32,
                                               ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, ctypes.c_uint64,
                                               c_i32p, c_f64p]),
    "gw_simrank_last_steps": (ctypes.c_int, [c_vp, c_i64p]),
    "gw_simrank_last_slow_queries": (ctypes.c_int, [c_vp, c_i64p]),
    "gw_simrank_last_error": (ctypes.c_int, [c_vp, c_i32p]),
    "gw_simrank_check_args": (ctypes.c_int, [c_vp, ctypes.c_double, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32]),
    "gw_simrank_exact": (ctypes.c_int, [c_vp, ctypes.c_double, ctypes.c_int32, c_i64p, ctypes.c_int64, c_f64p]),
}

_lib = None


def load():
    """Loads (building first if needed) libgraphwalk.so and types every entry point."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        from . import build as _build
        _build.build()
    L = ctypes.CDLL(os.environ.get("GW_LIB_OVERRIDE") or LIB_PATH)    # override: kernel-variant experiments (tools/)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(L, name)          # AttributeError here = header/library mismatch: fail loudly
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


def check(rc):
    if rc != GW_OK:
        raise_for(rc)


def raise_for(rc):
    msg = load().gw_last_error().decode("utf-8", "replace")
    if rc == GW_E_KEY:
        raise KeyError(msg)             # the reference raises KeyError on unknown nodes / edges
    if rc == GW_E_IO:
        raise IOError(msg)
    if rc == GW_E_INVALID:
        raise ValueError(msg)
    if rc == GW_E_TOO_LARGE:
        raise MemoryError(msg)
    raise GraphWalkError(rc, msg)


def ptr(a, ctype):
    if a is None:
        return None
    return a.ctypes.data_as(ctypes.POINTER(ctype))


def as_c(a, dtype):
    return np.ascontiguousarray(a, dtype=dtype)


class GraphHandle:
    """Owns one gw_graph*; frees it on garbage collection."""

    def __init__(self, raw):
        self._h = c_vp(raw)
        L = load()
        n, nnz = ctypes.c_int64(), ctypes.c_int64()
        flags, maxd, dev = ctypes.c_int32(), ctypes.c_int32(), ctypes.c_int32()
        check(L.gw_graph_info(self._h, ctypes.byref(n), ctypes.byref(nnz), ctypes.byref(flags), ctypes.byref(maxd),
                              ctypes.byref(dev)))
        self.n, self.nnz, self.flags, self.max_degree, self.device = n.value, nnz.value, flags.value, maxd.value, dev.value

    @property
    def h(self):
        if self._h is None:
            raise GraphWalkError(GW_E_STATE, "graph already freed")
        return self._h

    def close(self):
        if self._h is not None:
            load().gw_graph_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- constructors ----
    @staticmethod
    def from_edges(src, dst, w=None, directed=False, mode=GW_MODE_SIMPLE, n_slots=-1):
        L = load()
        src, dst = as_c(src, np.int64), as_c(dst, np.int64)
        if len(src) != len(dst):
            raise ValueError("src/dst length mismatch")
        wv = None if w is None else as_c(w, np.float64)
        out = c_vp()
        check(L.gw_graph_from_edges(ptr(src, ctypes.c_int64), ptr(dst, ctypes.c_int64), ptr(wv, ctypes.c_double),
                                    len(src), int(bool(directed)), int(mode), int(n_slots), ctypes.byref(out)))
        return GraphHandle(out.value)

    @staticmethod
    def from_file(path, delimiter=None, weighted=False, directed=False, mode=GW_MODE_SIMPLE, n_slots=-1):
        L = load()
        out = c_vp()
        d = None if delimiter is None else delimiter.encode()
        check(L.gw_graph_load_edgelist(os.fsencode(path), d, int(bool(weighted)), int(bool(directed)), int(mode),
                                       int(n_slots), ctypes.byref(out)))
        return GraphHandle(out.value)

    @staticmethod
    def rmat(scale, n_tuples, a=0.45, b=0.15, c=0.15, seed=1):
        out = c_vp()
        check(load().gw_graph_rmat(int(scale), int(n_tuples), a, b, c, int(seed), ctypes.byref(out)))
        return GraphHandle(out.value)

    @staticmethod
    def barabasi_albert(n, m, seed=1):
        out = c_vp()
        check(load().gw_graph_barabasi_albert(int(n), int(m), int(seed), ctypes.byref(out)))
        return GraphHandle(out.value)

    # ---- export ----
    def csr(self, weights=False, node_ids=True, first_seen=True):
        rp = np.empty(self.n + 1, dtype=np.int64)
        col = np.empty(self.nnz, dtype=np.int32)
        w = np.empty(self.nnz, dtype=np.float64) if weights else None
        ids = np.empty(self.n, dtype=np.int64) if node_ids else None
        fs = np.empty(self.n, dtype=np.int64) if first_seen else None
        check(load().gw_graph_csr(self.h, ptr(rp, ctypes.c_int64), ptr(col, ctypes.c_int32), ptr(w, ctypes.c_double),
                                  ptr(ids, ctypes.c_int64), ptr(fs, ctypes.c_int64)))
        return dict(row_ptr=rp, col_idx=col, weights=w, node_ids=ids, first_seen=fs)

    def nonisolated(self):
        cnt = ctypes.c_int64()
        out = np.empty(self.n, dtype=np.int64)
        check(load().gw_graph_nonisolated(self.h, ptr(out, ctypes.c_int64), ctypes.byref(cnt)))
        return out[:cnt.value].copy()

    # ---- alias tables ----
    def alias_nodes(self):
        J = np.empty(self.nnz, dtype=np.int32)
        q = np.empty(self.nnz, dtype=np.float64)
        check(load().gw_alias_nodes(self.h, ptr(J, ctypes.c_int32), ptr(q, ctypes.c_double)))
        return J, q

    def alias_edges_size(self):
        t = ctypes.c_int64()
        check(load().gw_alias_edges_size(self.h, ctypes.byref(t)))
        return t.value

    def alias_edges(self, p, q, budget_bytes=0, fetch=True):
        L = load()
        if not fetch:
            check(L.gw_alias_edges(self.h, float(p), float(q), int(budget_bytes), None, None, None))
            return None
        total = self.alias_edges_size()
        off = np.empty(self.nnz + 1, dtype=np.int64)
        J = np.empty(total, dtype=np.int32)
        qq = np.empty(total, dtype=np.float64)
        check(L.gw_alias_edges(self.h, float(p), float(q), int(budget_bytes), ptr(off, ctypes.c_int64),
                               ptr(J, ctypes.c_int32), ptr(qq, ctypes.c_double)))
        return off, J, qq

    # ---- walks ----
    def walks(self, p, q, walk_length, starts, seed=0, walk_id_base=0, lens=True):
        starts = as_c(starts, np.int64)
        out = np.empty((len(starts), walk_length), dtype=np.int32)
        ln = np.empty(len(starts), dtype=np.int32) if lens else None
        check(load().gw_node2vec_walks(self.h, float(p), float(q), int(walk_length), ptr(starts, ctypes.c_int64),
                                       len(starts), int(seed), int(walk_id_base), ptr(out, ctypes.c_int32),
                                       ptr(ln, ctypes.c_int32)))
        return (out, ln) if lens else out

    def walks_dev(self, p, q, walk_length, d_starts, n_starts, d_out, d_lens=0, seed=0, walk_id_base=0, stream=0):
        check(load().gw_node2vec_walks_dev(self.h, float(p), float(q), int(walk_length), c_vp(d_starts),
                                           int(n_starts), int(seed), int(walk_id_base), c_vp(d_out),
                                           c_vp(d_lens) if d_lens else None, c_vp(stream) if stream else None))

    def last_handoff(self):
        m, t = ctypes.c_int32(), ctypes.c_int32()
        check(load().gw_graph_last_handoff(self.h, ctypes.byref(m), ctypes.byref(t)))
        return {"mode": {0: "none", 1: "direct", 2: "ring", 3: "packed"}[m.value], "copy_threads": t.value}

    def prepare_walks(self):
        ms = ctypes.c_double()
        check(load().gw_graph_prepare_walks(self.h, ctypes.byref(ms)))
        return ms.value

    def common_counts(self):
        """(counts[nnz], reverse_index[nnz]) of the walker's preprocessing, in CSR entry order."""
        cnt = np.empty(self.nnz, dtype=np.int32)
        rix = np.empty(self.nnz, dtype=np.int32)
        check(load().gw_graph_common_counts(self.h, ptr(cnt, ctypes.c_int32), ptr(rix, ctypes.c_int32)))
        return cnt, rix

    def walks_replay(self, walk_length, starts, uniforms, draw_offset=None):
        starts = as_c(starts, np.int64)
        uniforms = as_c(uniforms, np.float64)
        do = None if draw_offset is None else as_c(draw_offset, np.int64)
        out = np.empty((len(starts), walk_length), dtype=np.int32)
        ln = np.empty(len(starts), dtype=np.int32)
        check(load().gw_node2vec_walks_replay(self.h, int(walk_length), ptr(starts, ctypes.c_int64), len(starts),
                                              ptr(uniforms, ctypes.c_double), len(uniforms),
                                              ptr(do, ctypes.c_int64), ptr(out, ctypes.c_int32),
                                              ptr(ln, ctypes.c_int32)))
        return out, ln

    def walk_traffic_dev(self, p, q, walk_length, d_starts, n_starts, seed=0, walk_id_base=0, stream=0):
        out = np.zeros(5, dtype=np.int64)
        check(load().gw_node2vec_walk_traffic_dev(self.h, float(p), float(q), int(walk_length), c_vp(d_starts),
                                                  int(n_starts), int(seed), int(walk_id_base),
                                                  ptr(out, ctypes.c_int64), c_vp(stream) if stream else None))
        return dict(steps=int(out[0]), random_accesses=int(out[1]), streamed_bytes=int(out[2]),
                    intersections=int(out[3]), extra_proposals=int(out[4]))

    def byte_model_dev(self, d_walks, n_walks, walk_length, second_order, stream=0):
        steps, sec = ctypes.c_int64(), ctypes.c_int64()
        check(load().gw_walks_byte_model_dev(self.h, c_vp(d_walks), int(n_walks), int(walk_length),
                                             int(bool(second_order)), ctypes.byref(steps), ctypes.byref(sec),
                                             c_vp(stream) if stream else None))
        return steps.value, sec.value

    # ---- SimRank ----
    def simrank_topk(self, queries, c, step, sample, k, mode=GW_SIMRANK_MC, seed=0, query_id_base=0):
        queries = as_c(queries, np.int64)
        ids = np.empty((len(queries), k), dtype=np.int32)
        sc = np.empty((len(queries), k), dtype=np.float64)
        check(load().gw_simrank_topk(self.h, ptr(queries, ctypes.c_int64), len(queries), float(c), int(step),
                                     int(sample), int(k), int(mode), int(seed), int(query_id_base),
                                     ptr(ids, ctypes.c_int32), ptr(sc, ctypes.c_double)))
        return ids, sc

    def simrank_topk_dev(self, d_queries, nq, c, step, sample, k, d_ids, d_scores, mode=GW_SIMRANK_MC, seed=0,
                         query_id_base=0, stream=0):
        check(load().gw_simrank_topk_dev(self.h, c_vp(d_queries), int(nq), float(c), int(step), int(sample), int(k),
                                         int(mode), int(seed), int(query_id_base), c_vp(d_ids), c_vp(d_scores),
                                         c_vp(stream) if stream else None))

    def simrank_rows(self, queries, c, step, sample, mode=GW_SIMRANK_MC, seed=0, query_id_base=0):
        queries = as_c(queries, np.int64)
        out = np.empty((len(queries), self.n), dtype=np.float64)
        check(load().gw_simrank_rows(self.h, ptr(queries, ctypes.c_int64), len(queries), float(c), int(step),
                                     int(sample), int(mode), int(seed), int(query_id_base),
                                     ptr(out, ctypes.c_double)))
        return out

    def simrank_rows_javarng(self, queries, c, step, sample, rng_states):
        """Replay mode: java.util.Random per query from the given 48-bit states; returns (rows, states after)."""
        queries = as_c(queries, np.int64)
        st = np.ascontiguousarray(np.asarray(rng_states, dtype=np.uint64)).copy()
        if len(st) != len(queries):
            raise ValueError("one rng state per query")
        out = np.empty((len(queries), self.n), dtype=np.float64)
        check(load().gw_simrank_rows_javarng(self.h, ptr(queries, ctypes.c_int64), len(queries), float(c), int(step),
                                             int(sample), ptr(st, ctypes.c_uint64), ptr(out, ctypes.c_double)))
        return out, st

    def topsim_rows_javarng(self, queries, c, step, sample, rng_states, mode=0, max_paths=0):
        """Replay mode of TopSim_singleSample (mode 0) / TopSim_Enumerate (mode 1); returns (rows x SAMPLE, states after)."""
        queries = as_c(queries, np.int64)
        st = np.ascontiguousarray(np.asarray(rng_states, dtype=np.uint64)).copy()
        if len(st) != len(queries):
            raise ValueError("one rng state per query")
        if max_paths <= 0:
            max_paths = 2 * int(step) * int(sample) + 1           # paths(l) <= 1 + l * SAMPLE for the hybrid tree
        out = np.empty((len(queries), self.n), dtype=np.float64)
        check(load().gw_topsim_rows_javarng(self.h, ptr(queries, ctypes.c_int64), len(queries), float(c), int(step),
                                            int(sample), int(mode), int(max_paths), ptr(st, ctypes.c_uint64),
                                            ptr(out, ctypes.c_double)))
        return out, st

    def simrank_cache_javarng(self, queries, c, step, sample, capacity, rng_states, mode=0, max_paths=0):
        """Replay mode of SingleRandomWalk_M (mode 0) / TopSim_singleSample_M (mode 1): per query the heap arrays of its
        FixedCacheMap(capacity).  Returns (list of (keys, float32 values) in heap order, states after)."""
        queries = as_c(queries, np.int64)
        st = np.ascontiguousarray(np.asarray(rng_states, dtype=np.uint64)).copy()
        if len(st) != len(queries):
            raise ValueError("one rng state per query")
        if max_paths <= 0:
            max_paths = 2 * int(step) * int(sample) + 1
        keys = np.zeros((len(queries), capacity), dtype=np.int32)
        vals = np.zeros((len(queries), capacity), dtype=np.float32)
        sizes = np.zeros(len(queries), dtype=np.int32)
        check(load().gw_simrank_cache_javarng(self.h, ptr(queries, ctypes.c_int64), len(queries), float(c), int(step),
                                              int(sample), int(mode), int(capacity), int(max_paths),
                                              ptr(st, ctypes.c_uint64), ptr(keys, ctypes.c_int32),
                                              ptr(vals, ctypes.c_float), ptr(sizes, ctypes.c_int32)))
        return [(keys[i, :sizes[i]].copy(), vals[i, :sizes[i]].copy()) for i in range(len(queries))], st

    def double_walk_paths(self, vertices, sample, step, seed=0, rng_states=None):
        """DoubleRandomWalk.samplePaths: int32 [nv, sample, step]; with rng_states (one 48-bit java.util.Random state per
        vertex) the replay kernel runs and the states after are returned too."""
        vertices = as_c(vertices, np.int64)
        out = np.zeros((len(vertices), sample, step), dtype=np.int32)
        st = None
        if rng_states is not None:
            st = np.ascontiguousarray(np.asarray(rng_states, dtype=np.uint64)).copy()
            if len(st) != len(vertices):
                raise ValueError("one rng state per vertex")
        check(load().gw_double_walk_paths(self.h, ptr(vertices, ctypes.c_int64), len(vertices), int(sample), int(step),
                                          int(seed), ptr(st, ctypes.c_uint64), ptr(out, ctypes.c_int32)))
        return out if st is None else (out, st)

    def double_walk_sims(self, paths, c, rows=None, exact_order=False):
        """DoubleRandomWalk.getSim over a path set [nv, sample, step]: rows x nv fp64."""
        paths = as_c(paths, np.int32)
        nv, sample, step = paths.shape
        rows = np.arange(nv, dtype=np.int64) if rows is None else as_c(rows, np.int64)
        out = np.empty((len(rows), nv), dtype=np.float64)
        check(load().gw_double_walk_sims(self.h, ptr(paths, ctypes.c_int32), nv, sample, step, float(c),
                                         ptr(rows, ctypes.c_int64), len(rows), int(bool(exact_order)),
                                         ptr(out, ctypes.c_double)))
        return out

    def topsim_mass(self, sources, weight, step, seed=0, call_id_base=0, rng_states=None, max_paths=0):
        """TopSim_doubleSample.sample / TopSim_Dev.sample: fp64 [ns, n, step+1] path masses (-1 = unset); with rng_states
        (one java.util.Random state per tree) the replay kernel runs and the states after are returned too."""
        sources = as_c(sources, np.int64)
        if max_paths <= 0:
            max_paths = max(int(step) * (int(weight) + 2) + 1, self.max_degree + 1)
        st = None
        if rng_states is not None:
            st = np.ascontiguousarray(np.asarray(rng_states, dtype=np.uint64)).copy()
            if len(st) != len(sources):
                raise ValueError("one rng state per tree")
        out = np.empty((len(sources), self.n, step + 1), dtype=np.float64)
        check(load().gw_topsim_mass(self.h, ptr(sources, ctypes.c_int64), len(sources), float(weight), int(step),
                                    int(max_paths), int(seed), int(call_id_base), ptr(st, ctypes.c_uint64),
                                    ptr(out, ctypes.c_double)))
        return out if st is None else (out, st)

    def topsim_mass_sims(self, mass, c, pair_a, pair_b, exact_order=False):
        mass = as_c(mass, np.float64)
        ns, n, s1 = mass.shape
        pa, pb = as_c(pair_a, np.int64), as_c(pair_b, np.int64)
        out = np.empty(len(pa), dtype=np.float64)
        check(load().gw_topsim_mass_sims(self.h, ptr(mass, ctypes.c_double), ns, s1 - 1, float(c), ptr(pa, ctypes.c_int64),
                                         ptr(pb, ctypes.c_int64), len(pa), int(bool(exact_order)), ptr(out, ctypes.c_double)))
        return out

    def simrank_last_steps(self):
        s = ctypes.c_int64()
        check(load().gw_simrank_last_steps(self.h, ctypes.byref(s)))
        return s.value

    def simrank_last_error(self):
        v = ctypes.c_int32()
        check(load().gw_simrank_last_error(self.h, ctypes.byref(v)))
        return v.value

    def simrank_last_slow_queries(self):
        s = ctypes.c_int64()
        check(load().gw_simrank_last_slow_queries(self.h, ctypes.byref(s)))
        return s.value

    def simrank_exact(self, c, iters, rows=None):
        rows = np.arange(self.n, dtype=np.int64) if rows is None else as_c(rows, np.int64)
        out = np.empty((len(rows), self.n), dtype=np.float64)
        check(load().gw_simrank_exact(self.h, float(c), int(iters), ptr(rows, ctypes.c_int64), len(rows),
                                      ptr(out, ctypes.c_double)))
        return out


class SkipGram:
    """gw_sgns: skip-gram with negative sampling over walks that sit in device memory (node2vec/src/main.py:92-101)."""

    def __init__(self, n_words, dimensions=128, seed=1, device=None):
        """n_words: the vocabulary size, or a GraphHandle (its vertices, on its device)."""
        if isinstance(n_words, GraphHandle):
            n_words, device = n_words.n, n_words.device if device is None else device
        out = c_vp()
        check(load().gw_sgns_create(int(n_words), int(dimensions), int(seed), int(device or 0), ctypes.byref(out)))
        self._m = c_vp(out.value)
        self.n, self.dimensions = int(n_words), int(dimensions)

    def close(self):
        if getattr(self, "_m", None) is not None and self._m.value:
            load().gw_sgns_free(self._m)
            self._m = c_vp(None)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def count_dev(self, d_walks, n_walks, walk_length, stream=0):
        check(load().gw_sgns_count_dev(self._m, c_vp(d_walks), int(n_walks), int(walk_length), c_vp(stream) if stream else None))

    def finalize_vocab(self, sample=1e-3, negative=5):
        check(load().gw_sgns_finalize_vocab(self._m, float(sample), int(negative)))

    def train_dev(self, d_walks, n_walks, walk_length, window=10, alpha=0.025, min_alpha=0.0001, words_before=0.0,
                  total_words=None, sentence_id_base=0, subsample=True, sequential=False, stream=0):
        tw = float(total_words if total_words is not None else self.info()["total_words"])
        check(load().gw_sgns_train_dev(self._m, c_vp(d_walks), int(n_walks), int(walk_length), int(window), float(alpha),
                                       float(min_alpha), float(words_before), tw, int(sentence_id_base), int(bool(subsample)),
                                       int(bool(sequential)), c_vp(stream) if stream else None))

    def info(self):
        n, d, tw, tp = ctypes.c_int64(), ctypes.c_int32(), ctypes.c_double(), ctypes.c_int64()
        check(load().gw_sgns_info(self._m, ctypes.byref(n), ctypes.byref(d), ctypes.byref(tw), ctypes.byref(tp)))
        return {"n": n.value, "dimensions": d.value, "total_words": tw.value, "trained_pairs": tp.value}

    def vectors(self, syn1neg=False, counts=False):
        v = np.empty((self.n, self.dimensions), dtype=np.float32)
        w = np.empty((self.n, self.dimensions), dtype=np.float32) if syn1neg else None
        c = np.empty(self.n, dtype=np.int64) if counts else None
        check(load().gw_sgns_vectors(self._m, ptr(v, ctypes.c_float), ptr(w, ctypes.c_float), ptr(c, ctypes.c_int64)))
        return (v,) + ((w,) if syn1neg else ()) + ((c,) if counts else ()) if (syn1neg or counts) else v

    def set_vectors(self, syn0=None, syn1neg=None):
        a = None if syn0 is None else as_c(syn0, np.float32)
        b = None if syn1neg is None else as_c(syn1neg, np.float32)
        check(load().gw_sgns_set_vectors(self._m, ptr(a, ctypes.c_float), ptr(b, ctypes.c_float)))


def node2vec_embeddings(handle, p, q, walk_length, num_walks, starts, dimensions=128, window=10, iter=1, negative=5,
                        sample=1e-3, alpha=0.025, min_alpha=0.0001, seed=1):
    """gw_node2vec_embeddings: walks -> vocabulary -> skip-gram, all on the device.  starts: [num_walks, n_starts] dense
    indices (the shuffled node list of every pass).  -> (vectors [n, dim] float32, counts [n] int64, seconds dict)."""
    starts = as_c(starts, np.int64).reshape(int(num_walks), -1)
    vec = np.empty((handle.n, int(dimensions)), dtype=np.float32)
    cnt = np.empty(handle.n, dtype=np.int64)
    sec = np.zeros(3, dtype=np.float64)
    check(load().gw_node2vec_embeddings(handle.h, float(p), float(q), int(walk_length), int(num_walks),
                                        ptr(starts, ctypes.c_int64), starts.shape[1], int(dimensions), int(window), int(iter),
                                        int(negative), float(sample), float(alpha), float(min_alpha), int(seed),
                                        ptr(vec, ctypes.c_float), ptr(cnt, ctypes.c_int64), ptr(sec, ctypes.c_double)))
    return vec, cnt, {"walks": sec[0], "vocabulary_scan": sec[1], "training": sec[2]}


def py_random_shuffle(items):
    """random.shuffle(items) for a contiguous int64 numpy array: same permutation, same state of Python's global `random`
    afterwards, 100x faster than the interpreter's loop (gw_py_random_shuffle)."""
    import random
    assert items.dtype == np.int64 and items.flags.c_contiguous
    ver, st, gauss = random.getstate()
    if ver != 3 or len(st) != 625:                       # an interpreter with another generator: keep the contract, lose the speed
        lst = items.tolist()
        random.shuffle(lst)
        items[:] = lst
        return items
    mt = np.array(st[:624], dtype=np.uint32)
    idx = ctypes.c_int32(st[624])
    rc = load().gw_py_random_shuffle(mt.ctypes.data_as(ctypes.POINTER(ctypes.c_uint32)), ctypes.byref(idx), ptr(items, ctypes.c_int64),
                                     len(items))
    if rc != 0:
        raise ValueError("gw_py_random_shuffle rejected its arguments")
    random.setstate((3, tuple(int(x) for x in mt) + (idx.value,), gauss))
    return items


def shard_range(n, rank, nranks):
    lo, hi = ctypes.c_int64(), ctypes.c_int64()
    check(load().gw_shard_range(int(n), int(rank), int(nranks), ctypes.byref(lo), ctypes.byref(hi)))
    return lo.value, hi.value


class Comm:
    """gw_comm: the C-ABI multi-GPU communicator (NCCL loaded at run time).  `unique_id()` on rank 0, ship the 128
    bytes to the other ranks, then `Comm(rank, nranks, id, device)` everywhere."""

    @staticmethod
    def unique_id():
        buf = ctypes.create_string_buffer(128)
        check(load().gw_comm_unique_id(ctypes.cast(buf, c_vp)))
        return buf.raw

    def __init__(self, rank, nranks, unique_id, device=0):
        if len(unique_id) != 128:
            raise ValueError("a NCCL unique id is 128 bytes")
        out = c_vp()
        buf = ctypes.create_string_buffer(bytes(unique_id), 128)
        check(load().gw_comm_init(int(rank), int(nranks), ctypes.cast(buf, c_vp), int(device), ctypes.byref(out)))
        self._c = c_vp(out.value)
        self.rank, self.nranks, self.device = int(rank), int(nranks), int(device)

    def close(self):
        if getattr(self, "_c", None) is not None and self._c.value:
            load().gw_comm_free(self._c)
            self._c = c_vp(None)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def walks(self, handle, p, q, walk_length, starts_all, seed=0, gather=True, out=None, lens=None):
        """gw_node2vec_walks_sharded: the whole [n, L] corpus on every rank (gather = True / 1), on rank 0 only
        (gather = 2) or this rank's rows only, at their global offsets (gather = False / 0)."""
        starts_all = as_c(starts_all, np.int64)
        if out is None:
            out = np.full((len(starts_all), walk_length), -1, dtype=np.int32)
        ln = np.zeros(len(starts_all), dtype=np.int32) if lens is None else lens
        check(load().gw_node2vec_walks_sharded(handle.h, self._c, float(p), float(q), int(walk_length),
                                               ptr(starts_all, ctypes.c_int64), len(starts_all), int(seed),
                                               int(gather), ptr(out, ctypes.c_int32), ptr(ln, ctypes.c_int32)))
        return out, ln

    def last_times(self):
        """(own-slice kernel ms, NCCL exchange ms) of the last sharded call."""
        a, b = ctypes.c_double(), ctypes.c_double()
        check(load().gw_comm_last_times(self._c, ctypes.byref(a), ctypes.byref(b)))
        return a.value, b.value

    def simrank_topk(self, handle, queries_all, c, step, sample, k, mode=GW_SIMRANK_MC, seed=0, out=None):
        queries_all = as_c(queries_all, np.int64)
        if out is not None:
            ids, sc = out
        else:
            ids = np.empty((len(queries_all), k), dtype=np.int32)
            sc = np.empty((len(queries_all), k), dtype=np.float64)
        check(load().gw_simrank_topk_sharded(handle.h, self._c, ptr(queries_all, ctypes.c_int64), len(queries_all),
                                             float(c), int(step), int(sample), int(k), int(mode), int(seed),
                                             ptr(ids, ctypes.c_int32), ptr(sc, ctypes.c_double)))
        return ids, sc


def alias_setup(probs):
    probs = as_c(probs, np.float64)
    J = np.zeros(len(probs), dtype=np.int32)
    q = np.zeros(len(probs), dtype=np.float64)
    check(load().gw_alias_setup(ptr(probs, ctypes.c_double), len(probs), ptr(J, ctypes.c_int32),
                                ptr(q, ctypes.c_double)))
    return J, q


def device_count():
    c = ctypes.c_int()
    rc = load().gw_device_count(ctypes.byref(c))
    return c.value if rc == GW_OK else 0


def set_device(i):
    check(load().gw_set_device(int(i)))


def kernel_launches():
    return int(load().gw_kernel_launches())
