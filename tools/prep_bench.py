"""Times the walker's one-off preprocessing (per-edge common-neighbour counts) for both kernels:
python tools/prep_bench.py [scale ...]   -> one JSON line per (scale, build)"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from graph_embedding_b200 import _lib

for scale in [int(x) for x in (sys.argv[1:] or ["22"])]:
    for build in ("v2", "v1"):
        if build == "v1":
            os.environ["GW_CN_BUILD"] = "v1"
        else:
            os.environ.pop("GW_CN_BUILD", None)
        h = _lib.GraphHandle.rmat(scale, 16 << scale, seed=1)
        t0 = time.perf_counter()
        ms = h.prepare_walks()
        wall = time.perf_counter() - t0
        print(json.dumps({"scale": scale, "build": build, "device_ms": round(ms, 2), "wall_ms": round(wall * 1e3, 2),
                          "directed_entries": h.nnz, "max_degree": h.max_degree}), flush=True)
        del h
