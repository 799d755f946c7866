"""Corpus hand-off into PAGEABLE host memory as a function of the copy-thread count (GW_HOST_THREADS) and the path
(GW_E2E = ring | packed): R-MAT-22, one pass of 4.18 M walks x 80 (1.34 GB), 3 reps each.  One JSON line per setting."""
import ctypes
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from graph_embedding_b200 import _lib

L_ = _lib.load()
out = None
for mode in ("ring", "packed"):
    for threads in (1, 2, 3, 4, 6, 8, 12, 16):
        os.environ["GW_HOST_THREADS"] = str(threads)
        os.environ["GW_E2E"] = mode
        g = _lib.GraphHandle.rmat(22, 16 << 22, seed=1)          # a fresh handle: the pool is sized when it is first used
        g.prepare_walks()
        starts = g.nonisolated()
        if out is None:
            out = np.zeros((len(starts), 80), dtype=np.int32)

        def one(i):
            _lib.check(L_.gw_node2vec_walks(g.h, 0.25, 4.0, 80, starts.ctypes.data_as(_lib.c_i64p), len(starts), 5, i * len(starts),
                                            out.ctypes.data_as(_lib.c_i32p), None))
        one(0)
        t0 = time.perf_counter()
        for i in range(3):
            one(1 + i)
        dt = (time.perf_counter() - t0) / 3
        print(json.dumps({"mode": g.last_handoff()["mode"], "copy_threads": g.last_handoff()["copy_threads"],
                          "G_steps_per_s": round(len(starts) * 79 / dt / 1e9, 2), "corpus_GB_per_s": round(out.nbytes / dt / 1e9, 1)}), flush=True)
        del g
