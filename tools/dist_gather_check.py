"""Multi-GPU end-to-end check (torchrun, NCCL): every rank walks / queries its contiguous shard on its own
GPU, the corpus and the top-k tiles are all-gathered over NVLink, and the gathered result must equal what
ONE GPU produces for the whole input (walk / query ids are global, so the result does not depend on the
world size).  Also times the NCCL gather of a scale-22 corpus.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29520 tools/dist_gather_check.py
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

from graph_embedding_b200 import _lib, dist as gd

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
_lib.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
dev = torch.device("cuda", local)
out = {"world": world}

g = _lib.GraphHandle.rmat(18, 16 << 18, seed=1)
starts = g.nonisolated()
corpus = gd.sharded_walks(g, 0.25, 4.0, 80, starts, seed=11, torch_device=dev)
whole, _ = g.walks(0.25, 4.0, 80, starts, seed=11)                      # the same on one GPU
out["walks_equal_single_gpu"] = bool(np.array_equal(corpus, whole))
out["walks"] = int(len(starts))

b = _lib.GraphHandle.barabasi_albert(200000, 8, seed=1)
q = np.random.RandomState(4).choice(b.n, 1001, replace=False).astype(np.int64)
ids, sc = gd.sharded_simrank_topk(b, q, 0.6, 5, 10000, 20, seed=3, torch_device=dev)
ids1, sc1 = b.simrank_topk(q, 0.6, 5, 10000, 20, seed=3)
out["topk_equal_single_gpu"] = bool(np.array_equal(ids, ids1) and sc.tobytes() == sc1.tobytes())

# NCCL gather of a device-resident scale-22 corpus block (what a single-corpus consumer would pay)
n_total = 4178039
lo, hi = gd.shard_range(n_total, rank, world)
blk = torch.randint(0, 1 << 22, (hi - lo, 80), dtype=torch.int32, device=dev)
gd.gather_rows(blk, n_total)
torch.cuda.synchronize(); dist.barrier()
t0 = time.perf_counter()
for _ in range(3):
    full = gd.gather_rows(blk, n_total)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / 3
out["gather_scale22_corpus_ms"] = dt * 1e3
out["gather_GBps_per_rank_received"] = n_total * 80 * 4 / dt / 1e9
if rank == 0:
    print(json.dumps(out))
dist.destroy_process_group()
