"""Edge-list loader throughput: writes an R-MAT-20 text edge list (16 M lines) and times gw_graph_load_edgelist."""
import os, sys, time, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from graph_embedding_b200 import _lib
rs = np.random.RandomState(1)
m = 16 << 20
u = rs.randint(0, 1 << 20, m); v = rs.randint(0, 1 << 20, m)
path = os.path.join(tempfile.gettempdir(), "loader_bench.edgelist")
t0 = time.perf_counter()
np.savetxt(path, np.stack([u, v], 1), fmt="%d", delimiter=",")
print("wrote %.0f MB in %.1f s" % (os.path.getsize(path) / 1e6, time.perf_counter() - t0), flush=True)
for i in range(2):
    t0 = time.perf_counter()
    h = _lib.GraphHandle.from_file(path, delimiter=",")
    dt = time.perf_counter() - t0
    print("load: %d lines in %.2f s = %.1f M lines/s (%d vertices, %d directed entries)" % (m, dt, m / dt / 1e6, h.n, h.nnz), flush=True)
    del h
os.unlink(path)
