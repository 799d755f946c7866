"""One rank of the C-ABI multi-GPU check (tests/test_gpu_comm.py, tools): python comm_worker.py RANK NRANKS IDFILE OUT.
Rank 0 creates the NCCL unique id and publishes it through IDFILE (the bootstrap any launcher can do); every rank
builds the same graphs on its own GPU, runs the sharded entry points and compares the gathered result with the
single-GPU call on the same inputs.  Writes "ok" to OUT.<rank>."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from graph_embedding_b200 import _lib

rank, nranks, idfile, out = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3], sys.argv[4]
ndev = _lib.device_count()
device = rank % ndev
_lib.set_device(device)
if rank == 0:
    uid = _lib.Comm.unique_id()
    with open(idfile + ".tmp", "wb") as f:
        f.write(uid)
    os.replace(idfile + ".tmp", idfile)
else:
    t0 = time.time()
    while not os.path.exists(idfile):
        if time.time() - t0 > 120:
            raise SystemExit("no unique id after 120 s")
        time.sleep(0.05)
    uid = open(idfile, "rb").read()
comm = _lib.Comm(rank, nranks, uid, device)

h = _lib.GraphHandle.rmat(16, 16 << 16, seed=1)
starts = np.tile(h.nonisolated(), 2)[:100003]                    # odd count: ragged shards
whole, wl = h.walks(0.25, 4.0, 40, starts, seed=9)
got, gl = comm.walks(h, 0.25, 4.0, 40, starts, seed=9)
assert np.array_equal(got, whole) and np.array_equal(gl, wl), "gathered corpus differs from the single-GPU corpus"
ct, gt = comm.last_times()
assert ct > 0 and (gt > 0 or nranks == 1), (ct, gt)               # own slice | NCCL exchange, device-timed
root, rl = comm.walks(h, 0.25, 4.0, 40, starts, seed=9, gather=2)  # gather to rank 0 only (grouped ncclSend / ncclRecv)
if rank == 0:
    assert np.array_equal(root, whole) and np.array_equal(rl, wl)
else:
    assert (root == -1).all()                                      # nothing is written on the other ranks
mine, ml = comm.walks(h, 4.0, 0.5, 40, starts, seed=9, gather=False)
lo, hi = _lib.shard_range(len(starts), rank, nranks)
ref, _ = h.walks(4.0, 0.5, 40, starts[lo:hi], seed=9, walk_id_base=lo)
assert np.array_equal(mine[lo:hi], ref) and (mine[:lo] == -1).all() and (mine[hi:] == -1).all()

b = _lib.GraphHandle.barabasi_albert(200000, 8, seed=3)
q = np.random.RandomState(1).choice(b.n, 1001, replace=False).astype(np.int64)
for mode in (_lib.GW_SIMRANK_MC, _lib.GW_SIMRANK_HYBRID):
    ids, sc = b.simrank_topk(q, 0.6, 5, 2000, 20, mode=mode, seed=4)
    gi, gs = comm.simrank_topk(b, q, 0.6, 5, 2000, 20, mode=mode, seed=4)
    assert np.array_equal(gi, ids) and gs.tobytes() == sc.tobytes(), "gathered top-k differs (mode %d)" % mode
try:
    comm.walks(h, 1.0, 1.0, 10, np.array([0, h.n], dtype=np.int64))
    raise SystemExit("an out-of-range start node was accepted")
except KeyError:
    pass                                                         # every rank refuses together, nobody waits inside NCCL
# fewer units than ranks + an argument every rank must refuse: ranks with an empty slice return too, nobody hangs
for bad in (dict(p=-1.0), dict(L=0)):
    try:
        comm.walks(h, bad.get("p", 1.0), 1.0, bad.get("L", 10), np.array([0], dtype=np.int64))
        raise SystemExit("an invalid argument was accepted")
    except ValueError:
        pass
try:
    comm.simrank_topk(b, q[:1], 0.6, 11, 100, 20)                  # step out of range, one query for nranks ranks
    raise SystemExit("step = 11 was accepted")
except ValueError:
    pass
try:
    comm.simrank_topk(b, np.array([b.n + 3], dtype=np.int64), 0.6, 5, 100, 20)
    raise SystemExit("an out-of-range query was accepted")
except KeyError:
    pass
one, _ = comm.walks(h, 1.0, 1.0, 10, np.array([5], dtype=np.int64), seed=1)   # one unit: all but one slice are empty
ref1, _ = h.walks(1.0, 1.0, 10, np.array([5], dtype=np.int64), seed=1)
assert np.array_equal(one, ref1)
comm.close()
with open("%s.%d" % (out, rank), "w") as f:
    f.write("ok")
