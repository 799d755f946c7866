"""CPU-only checks of the boundary: the shared library loads, exports every symbol that
include/graphwalk.h declares, and refuses to compute without a GPU (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import ROOT
from graph_embedding_b200 import _lib


def header_symbols():
    src = open(os.path.join(ROOT, "include", "graphwalk.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(gw_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    L = _lib.load()
    syms = header_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(L, s), "libgraphwalk.so does not export %s" % s
    assert sorted(_lib.SIGNATURES) == syms, "ctypes signature table and header disagree"
    assert L.gw_version() >= 100


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(_lib.GraphWalkError) as e:
        _lib.GraphHandle.from_edges([1, 2], [2, 3])
    assert "no CPU fallback" in str(e.value)
    with pytest.raises(_lib.GraphWalkError):
        _lib.alias_setup([0.5, 0.5])
    assert _lib.device_count() == 0


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "graph_embedding_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".java")):
                txt = open(os.path.join(dp, f), errors="replace").read()
                assert "oracle" not in txt.replace("oracle helper", ""), "%s mentions the oracle" % f


def test_host_side_java_heap_and_format():
    from graph_embedding_b200 import simrank as sr
    from oracle import simrank_oracle as S
    rs = np.random.RandomState(0)
    for _ in range(50):
        n, k = int(rs.randint(1, 60)), int(rs.randint(1, 25))
        row = rs.rand(n) * (rs.rand(n) < 0.4)
        row[rs.randint(n)] = row[rs.randint(n)]                # force some ties
        got = sr._row_topk_exact(row, k)
        ids, vals = S.fixedmaxpq_topk(row, k)
        assert [g[0] for g in got] == ids.tolist() and [g[1] for g in got] == vals.tolist()
    for x, d in [(0.0000005, 6), (0.125, 2), (0.05161244, 8), (1e-9, 6), (0.8, 8)]:
        assert sr.java_format(x, d) == S.java_fmt(x, d)


def test_emb_wire_format_round_trips_the_shipped_file(tmp_path):
    """node2vec/emb/karate.emb (word2vec text format written by main.py:98): reading it and writing it back through
    the host shim reproduces the file byte for byte."""
    from conftest import DATA
    from graph_embedding_b200 import main as cli
    src = os.path.join(DATA, "karate.emb")
    words, vecs = cli.load_word2vec_format(src)
    assert len(words) == 34 and vecs.shape == (34, 128) and words[:2] == ["34", "33"]
    out = str(tmp_path / "k.emb")
    cli.save_word2vec_format(out, words, vecs)
    assert open(out, "rb").read() == open(src, "rb").read()


def test_shard_range_matches_the_host_rule():
    from graph_embedding_b200 import dist
    for n in (0, 1, 7, 100003):
        for world in (1, 2, 3, 8):
            assert [_lib.shard_range(n, r, world) for r in range(world)] == [dist.shard_range(n, r, world) for r in range(world)]


def test_corpus_unpack24_host_half_of_the_packed_handoff():
    """gw_corpus_unpack24 (copy-thread pool + AVX2 / scalar widening) against numpy, ragged rows included: the host
    half of gw_node2vec_walks' packed hand-off runs without a device."""
    L_ = _lib.load()
    rs = np.random.RandomState(5)
    for n_walks, L, threads in [(1, 1, 1), (7, 5, 3), (8, 80, 2), (1000, 80, 4), (4099, 37, 16), (333, 80, 0)]:
        ids = rs.randint(0, 1 << 24, size=(n_walks, L)).astype(np.int32)
        ids[rs.randint(n_walks), :] = (1 << 24) - 1                          # the largest id is not a pad
        lens = np.full(n_walks, L, dtype=np.int32)
        short = rs.rand(n_walks) < 0.2
        lens[short] = rs.randint(0, L + 1, size=int(short.sum()))
        want = ids.copy()
        want[np.arange(L)[None, :] >= lens[:, None]] = -1
        b = np.where(want < 0, 0xFFFFFF, want).astype("<u4").view(np.uint8).reshape(-1, 4)[:, :3].copy().reshape(-1)
        out = np.full((n_walks, L), 12345, dtype=np.int32)
        rc = L_.gw_corpus_unpack24(b.ctypes.data_as(ctypes.c_void_p), lens.ctypes.data_as(_lib.c_i32p), n_walks, L, threads,
                                   out.ctypes.data_as(_lib.c_i32p))
        assert rc == 0 and np.array_equal(out, want), (n_walks, L, threads)
        out2 = np.empty((n_walks, L), dtype=np.int32)                         # lens == NULL: every row is full
        b2 = ids.astype("<u4").view(np.uint8).reshape(-1, 4)[:, :3].copy().reshape(-1)
        assert L_.gw_corpus_unpack24(b2.ctypes.data_as(ctypes.c_void_p), None, n_walks, L, threads,
                                     out2.ctypes.data_as(_lib.c_i32p)) == 0
        assert np.array_equal(out2, ids)
    assert L_.gw_corpus_unpack24(None, None, 4, 80, 1, None) == _lib.GW_E_INVALID


def test_native_shuffle_is_random_shuffle():
    """gw_py_random_shuffle: the permutation and the state of Python's global generator afterwards are those of
    random.shuffle (node2vec.py:51 shuffles ONE list cumulatively; the drop-in must consume the same draws)."""
    import random
    for seed, n in [(0, 34), (1, 1), (2, 2), (3, 1000), (4, 4097), (5, 65536), (6, 100003), (7, 0)]:
        random.seed(seed)
        for _ in range(3):                                          # position inside the 624-word block varies
            random.random()
        ref = list(range(100, 100 + n))
        st0 = random.getstate()
        random.shuffle(ref); random.shuffle(ref)                    # cumulative, as simulate_walks does
        tail_ref = [random.random() for _ in range(5)]
        random.setstate(st0)
        a = np.arange(100, 100 + n, dtype=np.int64)
        _lib.py_random_shuffle(a); _lib.py_random_shuffle(a)
        assert a.tolist() == ref, (seed, n)
        assert [random.random() for _ in range(5)] == tail_ref      # the generator was left where random.shuffle leaves it
