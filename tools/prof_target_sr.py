"""Short profiling target (ncu --set full): BA n=1e7 m=8, two batches of NQ (default 8192) TopSim queries (C=0.6 STEP=5 SAMPLE=1e4, k=20)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from graph_embedding_b200 import _lib

g = _lib.GraphHandle.barabasi_albert(10_000_000, 8, seed=1)
nq = int(os.environ.get("NQ", 8192))
q = torch.from_numpy(np.random.RandomState(2).choice(g.n, size=2 * nq, replace=False).astype(np.int64)).cuda()
ids = torch.empty((nq, 20), dtype=torch.int32, device="cuda")
sc = torch.empty((nq, 20), dtype=torch.float64, device="cuda")
for i in range(2):
    g.simrank_topk_dev(q.data_ptr() + 8 * nq * i, nq, 0.6, 5, 10000, 20, ids.data_ptr(), sc.data_ptr(), seed=7, query_id_base=i * nq)
torch.cuda.synchronize()
print("steps", g.simrank_last_steps(), "slow", g.simrank_last_slow_queries(), "err", g.simrank_last_error())
