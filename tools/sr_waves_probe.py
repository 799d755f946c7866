"""k_simrank_log (SingleRandomWalk top-k) on BA n=1e7: queries/s as a function of GW_SR_WAVES (CTAs launched per SM slot:
1 = one persistent CTA per SM looping over its queries, W = W x 148 CTAs in the launch, each with 1/W of the queries)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from graph_embedding_b200 import _lib
g = _lib.GraphHandle.barabasi_albert(int(os.environ.get("BA", 10_000_000)), 8, seed=1)
nq = int(os.environ.get("NQ", 16384))
q = torch.from_numpy(np.random.RandomState(3).choice(g.n, nq, replace=False).astype(np.int64)).cuda()
ids = torch.empty((nq, 20), dtype=torch.int32, device="cuda")
sc = torch.empty((nq, 20), dtype=torch.float64, device="cuda")
ref = None
for waves in os.environ.get("WAVES", "1,2,4,8,16,32,111").split(","):
    os.environ["GW_SR_WAVES"] = waves
    best = 1e9
    for rep in range(3):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        ev[0].record()
        g.simrank_topk_dev(q.data_ptr(), nq, 0.6, 5, 10000, 20, ids.data_ptr(), sc.data_ptr(), seed=7)
        ev[1].record()
        torch.cuda.synchronize()
        best = min(best, ev[0].elapsed_time(ev[1]))
    out = (ids.cpu().numpy().copy(), sc.cpu().numpy().copy())
    same = True if ref is None else bool(np.array_equal(out[0], ref[0]) and np.array_equal(out[1], ref[1]))
    ref = ref or out
    print("GW_SR_WAVES=%-4s %d queries in %.2f ms = %.0f queries/s, %.1f G steps/s; results equal to waves=1: %s" % (
        waves, nq, best, nq / best * 1e3, g.simrank_last_steps() / best / 1e6, same), flush=True)
