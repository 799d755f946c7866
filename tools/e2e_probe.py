import time, numpy as np, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
torch.cuda.set_device(0)
from graph_embedding_b200 import _lib
_lib.set_device(0)
g = _lib.GraphHandle.barabasi_albert(1000000, 8, seed=1)
q = np.random.RandomState(0).choice(g.n, 2048, replace=False).astype(np.int64)
dq = torch.from_numpy(q).cuda(); di = torch.empty((2048, 20), dtype=torch.int32, device='cuda'); ds = torch.empty((2048, 20), dtype=torch.float64, device='cuda')
for i in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    g.simrank_topk_dev(dq.data_ptr(), 2048, 0.6, 5, 10000, 20, di.data_ptr(), ds.data_ptr(), seed=i)
    torch.cuda.synchronize(); print('dev call', i, round((time.perf_counter() - t0) * 1e3, 2), 'ms')
for i in range(5):
    t0 = time.perf_counter(); ids, sc = g.simrank_topk(q, 0.6, 5, 10000, 20, seed=i); t1 = time.perf_counter()
    print('topk call', i, round((t1 - t0) * 1e3, 2), 'ms', g.simrank_last_steps(), 'slow', g.simrank_last_slow_queries())
