"""torchrun --nproc-per-node N tools/shard_sr_probe.py: per-rank times of gw_simrank_topk_sharded on the configs[4] shape,
repeated, beside a plain gw_simrank_topk_dev call on the same slice -- is a slow first call a sizing effect?"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
from graph_embedding_b200 import _lib
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
_lib.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
uid = torch.zeros(128, dtype=torch.uint8, device=dev)
if rank == 0:
    uid = torch.tensor(list(_lib.Comm.unique_id()), dtype=torch.uint8, device=dev)
if world > 1:
    dist.broadcast(uid, 0)
comm = _lib.Comm(rank, world, bytes(uid.cpu().numpy().tobytes()), local)
b = _lib.GraphHandle.barabasi_albert(10_000_000, 8, seed=1)
nq = int(os.environ.get("NQ", 1_000_000))
queries = np.random.RandomState(2).choice(b.n, size=nq, replace=False).astype(np.int64)
ids = np.zeros((nq, 20), dtype=np.int32)
sc = np.zeros((nq, 20), dtype=np.float64)
comm.simrank_topk(b, queries[:8192 * world], 0.6, 5, 10000, 20, seed=7)
for rep in range(3):
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    comm.simrank_topk(b, queries, 0.6, 5, 10000, 20, seed=7, out=(ids, sc))
    dt = time.perf_counter() - t0
    comp, gath = comm.last_times()
    print("rank %d rep %d: sharded call %.1f ms wall, compute %.1f ms, gather %.2f ms" % (rank, rep, dt * 1e3, comp, gath), flush=True)
lo, hi = _lib.shard_range(nq, rank, world)
d_q = torch.from_numpy(queries[lo:hi]).to(dev)
d_ids = torch.empty((hi - lo, 20), dtype=torch.int32, device=dev)
d_sc = torch.empty((hi - lo, 20), dtype=torch.float64, device=dev)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
for rep in range(3):
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ev[0].record()
    b.simrank_topk_dev(d_q.data_ptr(), hi - lo, 0.6, 5, 10000, 20, d_ids.data_ptr(), d_sc.data_ptr(), seed=7, query_id_base=lo,
                       stream=torch.cuda.current_stream().cuda_stream)
    ev[1].record()
    torch.cuda.synchronize()
    print("rank %d rep %d: plain slice call %.1f ms (%d queries)" % (rank, rep, ev[0].elapsed_time(ev[1]), hi - lo), flush=True)
comm.close()
if world > 1:
    dist.destroy_process_group()
