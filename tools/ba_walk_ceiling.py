"""What the memory system delivers for THE access stream of the SimRank kernels: plain first-order walks (p = q = 1: one random
16-byte nbr4 entry per step, no rejection, no accumulation) on the BA n = 1e7 m = 8 graph, 1280 walkers per SM."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from graph_embedding_b200 import _lib
g = _lib.GraphHandle.barabasi_albert(10_000_000, 8, seed=1)
g.prepare_walks()
for nw, L in ((1 << 22, 81), (1 << 23, 41), (1 << 24, 11)):
    starts = torch.from_numpy(np.random.RandomState(1).randint(0, g.n, size=nw).astype(np.int64)).cuda()
    out = torch.empty((nw, L), dtype=torch.int32, device="cuda")
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    g.walks_dev(1.0, 1.0, L, starts.data_ptr(), nw, out.data_ptr(), seed=3)
    torch.cuda.synchronize()
    ev[0].record()
    for r in range(3):
        g.walks_dev(1.0, 1.0, L, starts.data_ptr(), nw, out.data_ptr(), seed=4 + r)
    ev[1].record()
    torch.cuda.synchronize()
    ms = ev[0].elapsed_time(ev[1]) / 3
    print("BA-10M p=q=1: %d walks x %d steps in %.2f ms = %.1f G steps/s" % (nw, L - 1, ms, nw * (L - 1) / ms / 1e6), flush=True)
    del out, starts
