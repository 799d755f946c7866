"""Short profiling target (ncu --set full): R-MAT-22, the preprocessing kernels and two walk passes of the headline
config, nothing else.  python tools/prof_target.py [p q]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from graph_embedding_b200 import _lib

p, q = (float(sys.argv[1]), float(sys.argv[2])) if len(sys.argv) > 2 else (0.25, 4.0)
g = _lib.GraphHandle.rmat(22, 16 << 22, seed=1)
ms = g.prepare_walks()
starts = torch.from_numpy(np.random.RandomState(1).permutation(g.nonisolated())).cuda()
out = torch.empty((len(starts), 80), dtype=torch.int32, device="cuda")
for i in range(2):
    g.walks_dev(p, q, 80, starts.data_ptr(), len(starts), out.data_ptr(), seed=42, walk_id_base=i * len(starts))
torch.cuda.synchronize()
print("prepare_walks %.2f ms; %d walks" % (ms, len(starts)))
