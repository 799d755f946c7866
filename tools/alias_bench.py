"""alias_edges build time on the blog graph (sum deg^2 = 3.69e8 entries), bit pattern check against a sha."""
import gzip, os, sys, tempfile, time, hashlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from graph_embedding_b200 import _lib
DATA = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "data")
tmp = tempfile.NamedTemporaryFile(suffix=".txt", delete=False)
tmp.write(gzip.open(os.path.join(DATA, "blog.txt.gz"), "rb").read()); tmp.close()
h = _lib.GraphHandle.from_file(tmp.name, delimiter=",")
n = h.alias_edges_size()
for p in (0.5, 0.25, 2.0):
    t0 = time.perf_counter()
    h.alias_edges(p, 4.0, budget_bytes=16 << 30, fetch=False)
    dt = time.perf_counter() - t0
    print("p=%g: %d entries in %.1f ms = %.2f G entries/s" % (p, n, dt * 1e3, n / dt / 1e9), flush=True)
off, J, q = h.alias_edges(0.25, 4.0, budget_bytes=16 << 30)
print("sha256(J,q) p=0.25 q=4:", hashlib.sha256(J.tobytes() + q.tobytes()).hexdigest())
os.unlink(tmp.name)
