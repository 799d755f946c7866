"""Karate, 2 epochs, dim 32: edge AUC of the parallel (Hogwild) mode against the sequential restatement, per kernel
(GW_SG_PIPE) and per concurrency cap (GW_SG_WARPS), several runs each -- how much do lost updates cost on a 34-word vocabulary?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from graph_embedding_b200 import _lib
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from _auc import edge_auc
DATA = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "data")
path = os.path.join(DATA, "karate.edgelist")
h = _lib.GraphHandle.from_file(path, delimiter=" ")
g = h.csr(weights=False, node_ids=False, first_seen=False)
n = h.n
rs = np.random.RandomState(5)
starts = np.concatenate([rs.permutation(n) for _ in range(10)])
walks = np.full((len(starts), 40), -1, dtype=np.int32)
d_s = torch.from_numpy(starts.astype(np.int64)).cuda()
d_w = torch.empty((len(starts), 40), dtype=torch.int32, device="cuda")
h.prepare_walks()
h.walks_dev(1.0, 1.0, 40, d_s.data_ptr(), len(starts), d_w.data_ptr(), seed=5)
walks = d_w.cpu().numpy()
counts = np.bincount(walks[walks >= 0], minlength=n)
total = 2.0 * counts.sum()
for pipe in ("1", "0"):
    for warps in ("1", "2", "4", "8"):
        os.environ["GW_SG_PIPE"], os.environ["GW_SG_WARPS"] = pipe, warps
        aucs = []
        for run in range(6):
            m = _lib.SkipGram(h, 32, seed=3)
            m.count_dev(d_w.data_ptr(), len(walks), 40)
            m.finalize_vocab(sample=0.0, negative=5)
            for e in range(2):
                m.train_dev(d_w.data_ptr(), len(walks), 40, window=5, words_before=e * counts.sum(), total_words=total,
                            sentence_id_base=e * len(walks), subsample=False)
            aucs.append(edge_auc(m.vectors(), g["row_ptr"], g["col_idx"], np.random.RandomState(0)))
        print("pipe=%s warps=%s: AUC %s" % (pipe, warps, " ".join("%.3f" % a for a in aucs)), flush=True)
