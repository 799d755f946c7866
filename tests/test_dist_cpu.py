"""Host-side multi-GPU logic on CPU: world_size-2 (and 3) gloo process groups exercise the
sharding arithmetic and the padded all-gather used for walk corpora / top-k tiles.  The device
compute is replaced by a deterministic stub keyed by GLOBAL unit index, which is exactly the
property the Philox keying gives the real kernels."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from graph_embedding_b200 import dist as gd


def test_shard_ranges_partition_exactly():
    for n in (0, 1, 7, 8, 1000, 4177839):
        for world in (1, 2, 3, 4, 8):
            spans = [gd.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
            assert gd.shard_counts(n, world) == sizes


class StubHandle:
    """Stands in for GraphHandle: outputs depend only on (start, global walk id)."""

    def walks(self, p, q, L, starts, seed=0, walk_id_base=0, lens=True):
        ids = np.arange(len(starts), dtype=np.int64) + walk_id_base
        w = (starts[:, None] * 31 + ids[:, None] * 7 + np.arange(L)[None, :] + seed).astype(np.int32)
        return w, np.full(len(starts), L, dtype=np.int32)

    def simrank_topk(self, queries, c, step, sample, k, mode=0, seed=0, query_id_base=0):
        gid = np.arange(len(queries), dtype=np.int64) + query_id_base
        ids = ((queries[:, None] + gid[:, None] * 3 + np.arange(k)[None, :]) % 1000).astype(np.int32)
        sc = 1.0 / (1.0 + np.arange(k)[None, :] + gid[:, None] * 0.0)
        return ids, np.ascontiguousarray(np.broadcast_to(sc, ids.shape), dtype=np.float64)


def _worker(rank, world, port, n_units, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        h = StubHandle()
        starts = (np.arange(n_units, dtype=np.int64) * 13) % 101
        corpus = gd.sharded_walks(h, 0.25, 4.0, 8, starts, seed=5)
        ids, sc = gd.sharded_simrank_topk(h, starts, 0.6, 5, 100, 4, seed=5)
        np.savez(os.path.join(out_dir, "r%d.npz" % rank), corpus=corpus, ids=ids, sc=sc)
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("world,n_units", [(2, 11), (2, 8), (3, 10)])
def test_sharded_results_equal_single_process(tmp_path, world, n_units):
    mp.spawn(_worker, args=(world, _free_port(), n_units, str(tmp_path)), nprocs=world, join=True)
    h = StubHandle()
    starts = (np.arange(n_units, dtype=np.int64) * 13) % 101
    want_w, _ = h.walks(0.25, 4.0, 8, starts, seed=5)
    want_i, want_s = h.simrank_topk(starts, 0.6, 5, 100, 4, seed=5)
    for r in range(world):
        z = np.load(os.path.join(str(tmp_path), "r%d.npz" % r))
        assert np.array_equal(z["corpus"], want_w)
        assert np.array_equal(z["ids"], want_i) and np.array_equal(z["sc"], want_s)


def test_reference_arm_line_says_what_it_ran():
    """bench.py --impl reference (the CPU arm the driver runs beside ours): the JSON line must carry the samples really
    taken as `steps`, a timed region that fits the run, and the graph really walked in config.workload."""
    import json
    import subprocess
    import sys
    import time
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    t0 = time.time()
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1",
                          "--cpu-seconds", "1", "--no-secondary"], capture_output=True, text=True, timeout=300)
    wall = time.time() - t0
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["steps"] == 2 and line["warmup"] == 1 and line["value"] > 1e4
    assert line["steps"] * line["ms_per_step"] * 1e-3 <= wall                       # the timed region fits the run
    assert "R-MAT scale-10" in line["config"]["workload"] and line["cpu_baseline"]["kind"] == "port"
    assert line["preprocess"]["entries_per_s"] > 1e4 and line["e2e"]["h2d_bytes_per_step"] == 0
    assert line["unit"] == "walk-steps/s" and line["higher_is_better"] is True


def test_traffic_captures_cover_the_kernels_bench_reports():
    """roofline.traffic is looked up by the exact kernel instantiation and workload key bench.py is about to report
    (profiles/traffic.json); every headline kernel of the default run has a capture, and an unknown kernel gets None."""
    import importlib.util
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(root, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    walk = bench.ncu_traffic("k_walk_cn<VEC8=1,COUNT=0,MINB=5,HUB=0,RIDX=1>", "rmat22/ef16/abc=0.45,0.15,0.15/p=0.25/q=4/L=80", 330065081)
    assert walk and 40 < walk / 330065081 < 70                                   # ~51 B of DRAM traffic per walk step
    sr = bench.ncu_traffic("k_simrank_log<5>", "ba10000000/m8/sample=10000/step=5/k=20", 8192)
    assert sr and 4.5e6 < sr / 8192 < 5.5e6                                      # ~5.05 MB per query
    hy = bench.ncu_traffic("k_topsim_hybrid<5,true>", "ba10000000/m8/sample=10000/step=5/k=20", 2048)
    assert hy and 6e6 < hy / 2048 < 7.5e6                                        # scaled to another batch size
    sg = bench.ncu_traffic("k_sgns_pipe<4>", "rmat22/ef16/abc=0.45,0.15,0.15/p=0.25/q=4/L=80/dim=128/window=10/negative=5/sample=0.001", 10 ** 9)
    assert sg and 6000 < sg / 10 ** 9 < 6600                                     # ~6.3 kB per trained pair
    assert bench.ncu_traffic("k_walk_cn<VEC8=1,COUNT=0,MINB=6,HUB=0,RIDX=1>", "rmat22/ef16/abc=0.45,0.15,0.15/p=0.25/q=4/L=80") is None
    assert bench.ncu_traffic("k_simrank_log<5>", "ba1000000/m8/sample=10000/step=5/k=20") is None
