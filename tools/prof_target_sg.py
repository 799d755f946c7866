"""Short profiling target (ncu --set full): skip-gram training kernel on a 1 M-walk slice of the R-MAT-22 corpus."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from graph_embedding_b200 import _lib

g = _lib.GraphHandle.rmat(22, 16 << 22, seed=1)
g.prepare_walks()
starts = torch.from_numpy(np.random.RandomState(7).permutation(g.nonisolated())).cuda()
nw = len(starts)
d_w = torch.empty((nw, 80), dtype=torch.int32, device="cuda")
g.walks_dev(0.25, 4.0, 80, starts.data_ptr(), nw, d_w.data_ptr(), seed=11)
m = _lib.SkipGram(g, 128, seed=11)
m.count_dev(d_w.data_ptr(), nw, 80)
m.finalize_vocab(sample=1e-3, negative=5)
part = 1 << 19
m.train_dev(d_w.data_ptr(), part, 80, window=10, total_words=float(nw * 80))
torch.cuda.synchronize()
print("pairs", m.info()["trained_pairs"])
