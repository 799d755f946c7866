"""Pin the SimRank CPU oracle.  CPU only.
 * exact SimRank vs the repo's only golden vector (IsoMap_LE/data/0_333_5038_simrank_navie_top10)
 * TopSim_Enumerate restatement == SAMPLE x exact truncated SimRank (deterministic)
 * SingleRandomWalk / TopSim_singleSample restatements converge to the same expectation
 * java.util.Random clone vs published known answers
"""
import ctypes
import os

import numpy as np
import pytest

from oracle import simrank_oracle as S
from conftest import DATA


@pytest.fixture(scope="module")
def g333():
    return S.load_multigraph(os.path.join(DATA, "0_333_5038.txt"), 333, separator=" ")


def test_java_random_known_answers():
    out = np.zeros(2, dtype=np.int32)
    S.lib().jr_fill32.argtypes = [ctypes.c_int64, ctypes.c_int32, ctypes.POINTER(ctypes.c_int32)]
    S.lib().jr_fill32(42, 2, out.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)))
    assert out.tolist() == [-1170105035, 234785527]          # new Random(42).nextInt() x2
    S.lib().jr_fill32(0, 1, out.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)))
    assert int(out[0]) == -1155484576                        # new Random(0).nextInt()
    # bounded draws: range, power-of-two shortcut == high bits, uniformity
    r = S.java_random_ints(7, 10, 20000)
    assert r.min() == 0 and r.max() == 9
    cnt = np.bincount(r, minlength=10)
    assert ((cnt - 2000) ** 2 / 2000).sum() < 27.9           # chi2 df=9, alpha=1e-3
    r16 = S.java_random_ints(7, 16, 100)
    S.lib().jr_fill32(7, 2, out.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)))
    assert int(r16[0]) == ((int(out[0]) & 0xFFFFFFFF) >> 1) * 16 >> 31


def test_multigraph_loader(g333):
    assert g333["V"] == 333 and g333["row_ptr"][-1] == 2 * 5038      # both directions per line
    # file lists each undirected edge in both directions -> every neighbour appears twice
    a, b = int(g333["row_ptr"][0]), int(g333["row_ptr"][1])
    nb, cnt = np.unique(g333["col"][a:b], return_counts=True)
    assert (cnt == 2).all()


def test_exact_simrank_matches_golden(g333):
    """C=0.8, 30 sweeps reproduces every id:score of the shipped golden to 5e-8 (file is %.8f)."""
    sim = S.simrank_exact_matrix(g333, 0.8, 30)
    gold = S.read_sim_file(os.path.join(DATA, "0_333_5038_simrank_navie_top10.txt.sim.txt"), separator=" ")
    assert len(gold) == 333
    worst = 0.0
    for v, row in gold:
        for i, x in row:
            worst = max(worst, abs(sim[v, i] - x))
        # the golden row is the top-10 of the exact row; the printer that wrote the golden
        # dropped zero scores (9 rows are shorter than 10), so shorter rows must be padded by zeros
        ids, vals = S.fixedmaxpq_topk(sim[v], 10)
        assert np.allclose(vals[:len(row)], [x for _, x in row], atol=5e-8)
        assert (vals[len(row):] < S.MIN).all()
    assert worst <= 5e-8


def test_naive_loop_equals_matrix_form(g333):
    a = S.simrank_exact_naive(g333, 0.6, 3)       # the committed defaults C=0.6 STEP=3
    b = S.simrank_exact_matrix(g333, 0.6, 3)
    assert np.abs(a - b).max() < 1e-12


@pytest.fixture(scope="module")
def karate_multi():
    s, d = S.read_edge_file(os.path.join(DATA, "karate.edgelist"), " ")
    return S.build_multigraph(s, d, 35)            # slot 0 isolated, duplicate pair (9,33) kept


@pytest.mark.parametrize("v,step", [(1, 2), (12, 2), (34, 2), (0, 2), (5, 3)])
def test_enumerate_is_sample_times_truncated_exact(karate_multi, v, step):
    sample = 1000
    exact = S.simrank_exact_matrix(karate_multi, 0.6, step)
    row, made, _ = S.topsim_row(karate_multi, v, sample, step, 0.6, mode=1, max_paths=1 << 21)
    assert np.abs(row / sample - exact[v]).max() < 1e-9


def test_single_random_walk_converges(g333):
    step = 3
    exact = S.simrank_exact_matrix(g333, 0.6, step)
    v, sample = 5, 200000
    row, steps, _ = S.single_random_walk_row(g333, v, sample, step, 0.6, seed_state=S.java_seed(1))
    assert steps == sample * 2 * step
    top = np.argsort(-exact[v])[:20]
    assert np.abs(row[top] - exact[v][top]).max() < 4e-3
    assert np.sqrt(np.mean((row[top] - exact[v][top]) ** 2)) < 1.5e-3
    assert abs(row.sum() - exact[v].sum()) < 2e-2 * exact[v].sum()


def test_topsim_single_sample_converges(g333):
    step = 3
    exact = S.simrank_exact_matrix(g333, 0.6, step)
    v, sample = 5, 200000
    row, made, _ = S.topsim_row(g333, v, sample, step, 0.6, mode=0, seed_state=S.java_seed(2))
    top = np.argsort(-exact[v])[:20]
    assert np.abs(row[top] / sample - exact[v][top]).max() < 4e-3


def test_fixedmaxpq_semantics():
    # replace only if strictly greater: the earlier id keeps its place on ties at the boundary
    ids, vals = S.fixedmaxpq_topk(np.array([0.5, 0.1, 0.5, 0.1, 0.1]), 3)
    assert sorted(ids.tolist()) == [0, 1, 2] and vals.tolist() == [0.5, 0.5, 0.1]
    ids, vals = S.fixedmaxpq_topk(np.zeros(50), 20)
    assert sorted(ids.tolist()) == list(range(20))
    ids, vals = S.fixedmaxpq_topk(np.array([0.3, 0.9]), 20)
    assert ids.tolist() == [1, 0]
    # FixedMaxPQ.main : capacity 2, offers luo, liu, zhang -> descending zhang, luo (numbers stand in)
    ids, vals = S.fixedmaxpq_topk(np.array([2.0, 1.0, 3.0]), 2)
    assert ids.tolist() == [2, 0]


def test_print_and_eval_roundtrip(tmp_path, g333):
    sim = S.simrank_exact_matrix(g333, 0.6, 3)
    p = str(tmp_path / "out.txt")
    S.print_by_order(sim, p, 20, 6)
    raw = open(p + ".sim.txt", "rb").read()
    assert raw.count(b"\r\n") == 333
    rows = S.read_sim_file(p + ".sim.txt")
    assert rows[0][0] == 0 and len(rows[0][1]) == 20
    mean, pres = S.precision_rows(rows, rows)
    assert mean == 1.0
    assert S.java_fmt(0.0000005, 6) == "0.000001" and S.java_fmt(0.125, 2) == "0.13"


def test_fixed_cache_map_known_answer_and_host_mirror():
    """lxctools/FixedCacheMap.java: the put() sequence of its own main() (:134-148), traced by hand from the source
    (capacity 3; key 3 is evicted by (4, 8), key 4 accumulates to 16, key 1 to 1.1f, (5, 1) is refused because
    1 <= min 1.1f) -> size 3, iteration (1, 1.1) (2, 3.0) (4, 16.0), size 0 afterwards.  The host mirror class
    (graph_embedding_b200.simrank.FixedCacheMap) must agree with the C restatement on random put sequences, heap
    arrays included (ties and evictions exercise sink/swim order)."""
    hk, hv = S.fcm_puts(3, [1, 2, 3, 4, 4, 1, 5], [0.5, 3, 0.1, 8, 8, 0.6, 1])
    assert hk.tolist() == [1, 2, 4] and hv.tolist() == [np.float32(0.5) + np.float32(0.6), 3.0, 16.0]
    ks, vs = S.fcm_drain(hk, hv)
    assert ks.tolist() == [1, 2, 4] and vs.tolist() == [np.float32(1.1), 3.0, 16.0]
    from graph_embedding_b200 import simrank as sr
    rs = np.random.RandomState(3)
    for trial in range(40):
        cap = int(rs.randint(1, 12))
        n = int(rs.randint(0, 80))
        keys = rs.randint(0, 25, size=n)
        vals = (rs.randint(1, 6, size=n) * np.float32(0.1)).astype(np.float32)       # few distinct values: many ties
        hk, hv = S.fcm_puts(cap, keys, vals, key_space=25)
        m = sr.FixedCacheMap(cap)
        for k, v in zip(keys.tolist(), vals.tolist()):
            m.put(k, v)
        assert m.keys[1:] == hk.tolist() and [float(x) for x in m.values[1:]] == hv.tolist(), trial
        ks, vs = S.fcm_drain(hk, hv)
        got = list(m)
        assert [k for k, _ in got] == ks.tolist() and [float(v) for _, v in got] == vs.tolist()
        assert m.size() == 0
        assert (np.diff(vs) >= 0).all()


def test_cache_estimators_equal_dense_ones_when_nothing_is_evicted(g333):
    """SingleRandomWalk_M / TopSim_singleSample_M with a cache as large as the graph keep every target: same RNG
    stream as the dense classes, and each cached value is the float32 accumulation of the same increments."""
    st0 = S.java_seed(11)
    hk, hv, steps, st = S.single_random_walk_cache(g333, 5, 3000, 5, 333, seed_state=st0)
    row, steps2, st2 = S.single_random_walk_row(g333, 5, 3000, 5, seed_state=st0)
    assert (st, steps) == (st2, steps2)
    assert sorted(hk.tolist()) == np.nonzero(row)[0].tolist()
    assert np.abs(hv - row[hk]).max() < 2e-6
    hk, hv, made, st = S.topsim_cache(g333, 5, 3000, 5, 333, seed_state=st0)
    row, made2, st2 = S.topsim_row(g333, 5, 3000, 5, seed_state=st0)
    assert (st, made) == (st2, made2)
    assert sorted(hk.tolist()) == np.nonzero(row)[0].tolist()
    assert np.abs(hv - row[hk] / 3000).max() < 2e-6
    # a small cache is lossy but keeps the heavy hitters
    hk2, hv2, _, _ = S.single_random_walk_cache(g333, 5, 3000, 5, 40, seed_state=st0)
    assert len(hk2) == 40
    top = np.argsort(-row)[:5]
    assert set(top.tolist()) <= set(hk2.tolist())


def test_double_random_walk_converges_to_truncated_exact(g333):
    """DoubleRandomWalk.getSim is an unbiased estimator of SimRank truncated at STEP sweeps (first meeting of two
    independent walks at position s weighs C^(s+1)); anchored on the pinned exact routine."""
    paths, st = S.double_walk_paths(g333, 400, 3, S.java_seed(5))
    assert paths.shape == (333, 400, 3) and paths.min() >= 0
    pv = paths[[0, 7, 100]]
    lib = S.lib()
    lib.or_double_walk_sim.restype = ctypes.c_double
    lib.or_double_walk_sim.argtypes = [ctypes.POINTER(ctypes.c_int32), ctypes.c_int32, ctypes.c_int32, ctypes.c_double,
                                       ctypes.c_int64, ctypes.c_int64]
    exact = S.simrank_exact_matrix(g333, 0.6, 3)
    pp = np.ascontiguousarray(paths)
    errs = []
    for v, w in [(0, 7), (7, 100), (5, 166), (5, 37), (10, 200)]:
        est = lib.or_double_walk_sim(pp.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)), 400, 3, 0.6, v, w)
        errs.append(est - exact[v, w])
    assert np.abs(errs).max() < 6e-3, errs


def test_path_mass_tree_hand_traced_known_answer():
    """TopSim_doubleSample.sample + computePath (TopSim_doubleSample.java:66-178) traced by hand on the path graph
    0-1-2-3 from vertex 1 with SAMPLE = 8 (every weight >= degree: pure enumeration, no random draw):
      level 1: queue [0:4, 2:4]                          -> mass[0][1] = 4, mass[2][1] = 4
      level 2: 0 -> [1:4]; 2 -> [1:2, 3:2]               -> target 1 is the source (skipped), mass[3][2] = 2
      level 3: 1(w4) -> [0:2, 2:2]; 1(w2) -> [0:1, 2:1]; 3(w2) -> [2:2]
               queue order [0:2, 2:2, 0:1, 2:1, 2:2], the LAST path on a target wins (an overwrite, :175)
                                                          -> mass[0][3] = 1, mass[2][3] = 2
    and getSim (:189-199) of two such trees is the level-wise product sum."""
    g = S.build_multigraph(np.array([0, 1, 2]), np.array([1, 2, 3]), 4)
    st0 = S.java_seed(1)
    m, st = S.mass_tree(g, 1, 8.0, 3, st0)
    assert st == st0                                             # no draw consumed
    want = -np.ones((4, 4))
    want[0, 1] = 4; want[2, 1] = 4; want[3, 2] = 2; want[0, 3] = 1; want[2, 3] = 2
    assert np.array_equal(m, want)
    # from vertex 2 (rows keep file order: N(2) = [1, 3]): level 1 [1:4, 3:4]; level 2 1 -> [0:2, 2:2], 3 -> [2:4]
    # (target 2 is the source) -> mass[0][2] = 2; level 3 0 -> [1:2], 2(w2) -> [1:1, 3:1], 2(w4) -> [1:2, 3:2]
    # -> queue [1:2, 1:1, 3:1, 1:2, 3:2]: mass[1][3] = 2, mass[3][3] = 2.  NOT the mirror image of the tree from 1:
    # the queue order follows the adjacency order, and the last writer wins.
    m2, _ = S.mass_tree(g, 2, 8.0, 3, st0)
    want2 = -np.ones((4, 4))
    want2[1, 1] = 4; want2[3, 1] = 4; want2[0, 2] = 2; want2[1, 3] = 2; want2[3, 3] = 2
    assert np.array_equal(m2, want2)
    # getSim(1, 2): only targets reached by BOTH trees at the same level count: (0, level 3): 1 * 2, (3, 3)... by hand:
    # tree 1: {0: l1, 2: l1, 3: l2, 0: l3, 2: l3}; tree 2: {3: l1, 1: l1, 0: l2, 3: l3, 1: l3} -> no common (target, level)
    assert S.mass_sim(m, m2, 0.6) == 0.0
    m0, _ = S.mass_tree(g, 3, 8.0, 3, st0)                       # from the pendant vertex: 3 -> [2:8] -> [1:4, 3:4] -> 1 -> [0:2, 2:2], 3 -> [2:4]
    want0 = -np.ones((4, 4)); want0[2, 1] = 8; want0[1, 2] = 4; want0[0, 3] = 2; want0[2, 3] = 4
    assert np.array_equal(m0, want0)
    # getSim(1, 3): common (target, level) pairs: (2, 1): 4 * 8, (0, 3): 1 * 2, (2, 3): 2 * 4
    assert S.mass_sim(m, m0, 0.6) == 0.6 * 4 * 8 + 0.6 ** 3 * 1 * 2 + 0.6 ** 3 * 2 * 4
    # the MIN-filtered candidate queue of TopSim_Dev.compute (:72-83): zeros are never offered
    ids, vals = S.fixedmaxpq_topk_min(np.array([0.0, 0.5, 1e-12, 0.25, 0.5]), 2, S.MIN)
    assert sorted(ids.tolist()) == [1, 4] and vals.tolist() == [0.5, 0.5]
    ids, _ = S.fixedmaxpq_topk_min(np.array([0.0, 0.0, 1e-12]), 5, S.MIN)
    assert len(ids) == 0
    assert S.topsim_dev_sample_count(10000, 3, 20, 1) == 634 and S.topsim_dev_sample_count(10000, 3, 20, 3) == 0


def test_double_random_walk_hand_traced_known_answer():
    """DoubleRandomWalk.getSim (:77-91) on hand-made path sets: SAMPLE = 2, STEP = 2, C = 0.6.
    v: paths [5, 7], [6, 8]; w: paths [5, 9], [4, 7].  Pairs (i, j): (0,0) meet at position 0 -> C; (0,1): 5!=4, 7==7
    -> C^2; (1,0): 6!=5, 8!=9 -> nothing; (1,1): nothing.  Sum = C + C^2, divided by SAMPLE^2 = 4.  A dead end (-1)
    stops the comparison of that pair."""
    paths = np.zeros((2, 2, 2), dtype=np.int32)
    paths[0] = [[5, 7], [6, 8]]
    paths[1] = [[5, 9], [4, 7]]
    sim = S.double_walk_matrix(paths, 0.6)
    assert sim[0, 1] == sim[1, 0] == (0.6 + 0.6 ** 2) / 4 and sim[0, 0] == sim[1, 1] == 0
    paths[1, 1] = [-1, 0]                                        # w's second walk died at its first step
    assert S.double_walk_matrix(paths, 0.6)[0, 1] == 0.6 / 4


def test_blog_exact_fixture_is_the_expectation_of_the_restated_estimator(tmp_path):
    """tests/golden/blog_exact_s5.npz (exact SimRank, 5 sweeps, by the pinned exact routine) against the C restatement of
    SingleRandomWalk.walk (SingleRandomWalk.java:53-92) on blog.txt, for three mid-degree queries, on the
    well-conditioned entries of their rows (the 20 best targets of degree >= 16): the mean of 6 runs of SAMPLE = 1e5
    sits within 3e-4 + 6 standard errors of the fixture, and the row sums agree.  (The exact top-20 themselves are
    leaves of the hubs; one hit there adds 0.6 * 3992 / deg / SAMPLE and six CPU runs say nothing about the standard
    error -- the GPU test resolves them at SAMPLE = 1e7.)  The same fixture is what the GPU kernels are held against."""
    import gzip
    from conftest import DATA, GOLDEN
    gold = np.load(os.path.join(GOLDEN, "blog_exact_s5.npz"))
    f = tmp_path / "blog.txt"
    f.write_bytes(gzip.open(os.path.join(DATA, "blog.txt.gz"), "rb").read())
    g = S.load_multigraph(str(f), 10313, ",")
    deg = gold["degrees"]
    picks = np.nonzero((deg > 20) & (deg < 300))[0][:3]
    used = 0
    for r in picks:
        v = int(gold["queries"][r])
        top, val = gold["wc_ids"][r], gold["wc_scores"][r]
        st = S.java_seed(1000 + v)
        est, sums = [], []
        for _ in range(6):
            row, _, st = S.single_random_walk_row(g, v, 100000, 5, 0.6, st)
            est.append(row[top]); sums.append(row.sum())
        est = np.array(est)
        ok = (est > 0).all(axis=0)
        se = est.std(axis=0, ddof=1) / np.sqrt(len(est))
        assert (np.abs(est.mean(axis=0) - val)[ok] <= 3e-4 + 6 * se[ok]).all(), (v, (est.mean(axis=0) - val)[ok], se[ok])
        assert abs(np.mean(sums) - gold["row_sums"][r]) <= 0.03 * gold["row_sums"][r]
        used += int(ok.sum())
    assert used >= 40
