"""Skip-gram training kernel on a 2^20-walk slice of the R-MAT-22 corpus: pairs/s (GW_SG_PIPE=0 selects the unpipelined kernel)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from graph_embedding_b200 import _lib

g = _lib.GraphHandle.rmat(22, 16 << 22, a=0.45, b=0.15, c=0.15, seed=1)
g.prepare_walks()
starts = torch.from_numpy(np.random.RandomState(7).permutation(g.nonisolated())).cuda()
nw = len(starts)
d_w = torch.empty((nw, 80), dtype=torch.int32, device="cuda")
g.walks_dev(0.25, 4.0, 80, starts.data_ptr(), nw, d_w.data_ptr(), seed=11)
for dim in (128, 64, 256):
    m = _lib.SkipGram(g, dim, seed=11)
    m.count_dev(d_w.data_ptr(), nw, 80)
    m.finalize_vocab(sample=1e-3, negative=5)
    part = 1 << 20
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    m.train_dev(d_w.data_ptr(), 1 << 16, 80, window=10, total_words=float(nw * 80))      # warm-up
    p0 = m.info()["trained_pairs"]
    ev[0].record()
    m.train_dev(d_w.data_ptr() + 4 * 80 * (1 << 16), part, 80, window=10, total_words=float(nw * 80), sentence_id_base=1 << 16)
    ev[1].record()
    torch.cuda.synchronize()
    ms = ev[0].elapsed_time(ev[1])
    pairs = m.info()["trained_pairs"] - p0
    print("dim %d pipe=%s: %d pairs in %.1f ms = %.1f M pairs/s (%.2f TB/s by the row model)" % (
        dim, os.environ.get("GW_SG_PIPE", "1"), pairs, ms, pairs / ms / 1e3, pairs * 12.0 * dim * 4 / ms / 1e9), flush=True)
    del m
