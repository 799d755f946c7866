"""What bounds the corpus hand-off when N ranks share one box?

torchrun --nproc-per-node N tools/pcie_probe_nway.py     (or plain python for N = 1)

Every rank owns one GPU and measures, alone and then ALL RANKS AT ONCE (barrier-aligned):
  * pinned D2H of a 1.3 GB block (what one pass of the R-MAT-22 corpus is),
  * the same while 1..T host threads per rank copy pinned -> pageable memory (the ring's drain),
  * host memcpy bandwidth alone (T threads per rank, no DMA).
Rank 0 prints one JSON line: per-rank and aggregate GB/s for each case, NUMA / CPU facts of the box."""
import ctypes
import json
import os
import sys
import threading
import time

import numpy as np
import torch
import torch.distributed as dist

rank = int(os.environ.get("RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
dev = torch.device("cuda", local)
N = 1337 << 20
d = torch.empty(N, dtype=torch.uint8, device=dev)
h = torch.empty(N, dtype=torch.uint8).pin_memory()
page = np.empty(N, dtype=np.uint8)
page[:] = 1                                            # commit the pages
libc = ctypes.CDLL(None)
libc.memcpy.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t]
libc.memcpy.restype = ctypes.c_void_p


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def allgather(x):
    t = torch.tensor([float(x)], device=dev, dtype=torch.float64)
    if world == 1:
        return [float(x)]
    out = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(out, t)
    return [float(o.item()) for o in out]


def d2h(reps=4, together=True):
    h.copy_(d, non_blocking=True)
    if together:
        barrier()
    else:
        torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        h.copy_(d, non_blocking=True)
    torch.cuda.synchronize()
    return N * reps / (time.perf_counter() - t0) / 1e9


def host_copy(threads, seconds=0.6, with_dma=False):
    """threads copy disjoint slices pinned -> pageable in a loop (ctypes memcpy releases the GIL)."""
    stop = threading.Event()
    moved = [0] * threads
    per = N // threads

    def work(i):
        src = h.data_ptr() + i * per
        dst = page.ctypes.data + i * per
        while not stop.is_set():
            libc.memcpy(dst, src, per)
            moved[i] += per
    ths = [threading.Thread(target=work, args=(i,)) for i in range(threads)]
    barrier()
    t0 = time.perf_counter()
    for t in ths:
        t.start()
    dma = 0
    if with_dma:
        while time.perf_counter() - t0 < seconds:
            h.copy_(d, non_blocking=True)
            torch.cuda.synchronize()
            dma += N
    else:
        time.sleep(seconds)
    dt = time.perf_counter() - t0
    stop.set()
    for t in ths:
        t.join()
    dt2 = time.perf_counter() - t0
    return sum(moved) / dt2 / 1e9, dma / dt / 1e9


res = {"n_ranks": world, "cpus": os.cpu_count(), "block_MB": N >> 20}
try:
    res["numa_nodes"] = len([x for x in os.listdir("/sys/devices/system/node") if x.startswith("node")])
except OSError:
    res["numa_nodes"] = None
hw = os.cpu_count() or 1
tmax = max(1, min(16, hw // world))
v = allgather(d2h())
res["d2h_all_ranks_at_once"] = {"per_rank_GBs": [round(x, 1) for x in v], "sum_GBs": round(sum(v), 1)}
if world > 1:                                           # one rank at a time: the per-link ceiling
    alone = []
    for r in range(world):
        barrier()
        x = d2h(together=False) if r == rank else 0.0
        barrier()
        alone.append(max(allgather(x)))
    res["d2h_one_rank_at_a_time_GBs"] = [round(x, 1) for x in alone]
for t in sorted({1, max(1, tmax // 4), max(1, tmax // 2), tmax}):
    c, _ = host_copy(t)
    cs = allgather(c)
    c2, dm = host_copy(t, with_dma=True)
    cs2, dms = allgather(c2), allgather(dm)
    res["threads_%d" % t] = {"memcpy_only_sum_GBs": round(sum(cs), 1), "memcpy_with_dma_sum_GBs": round(sum(cs2), 1),
                             "dma_with_memcpy_sum_GBs": round(sum(dms), 1)}
if rank == 0:
    print(json.dumps(res), flush=True)
if world > 1:
    dist.destroy_process_group()
