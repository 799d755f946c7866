"""The reference's second, independent statement of the single-walk estimator: RandomWalkTest.testPairSimRank
(DeepSim/TopSimAll/src/simrank/random_test/RandomWalkTest.java:177-205) with its own isFirstMeet (:213-219).
It pins the increment C^(step/2) * deg(path[step/2]) / deg(path[step]) and the first-meeting rule without going
through SingleRandomWalk.java.  Restated here (numpy, vectorised over samples) and held against
  * the pinned exact routine (its expectation: SimRank truncated at L sweeps),
  * the oracle's restatement of SingleRandomWalk (CPU, -m "not gpu"),
  * the device kernels (-m gpu)."""
import os

import numpy as np
import pytest

from conftest import DATA
from oracle import simrank_oracle as S

G333 = os.path.join(DATA, "0_333_5038.txt")
PAIRS = [(5, 17), (5, 100), (0, 3), (200, 287), (332, 1)]


def pair_simrank_restated(g, src, des, C, L, sample, rs):
    """RandomWalkTest.java:177-205: per sample walk 2L uniform steps from src (:191-193, Graph.randNeighbor); whenever
    the walk sits on `des` at an even step and isFirstMeet holds (:195-197), add C^(step/2) * deg(path[step/2]) /
    deg(path[step]); the estimate is the sum / SAMPLE (:201).  (The Java probe averages 30 such runs; one run of
    30 x the samples is the same estimator.)"""
    rp, col = g["row_ptr"], g["col"]
    deg = np.diff(rp)
    path = np.empty((2 * L + 1, sample), dtype=np.int64)
    path[0] = src
    total = 0.0
    for step in range(1, 2 * L + 1):
        cur = path[step - 1]
        assert (deg[cur] > 0).all(), "the probe prints 'cur:-1' on a dead end; the test graphs have none"
        k = (rs.random_sample(sample) * deg[cur]).astype(np.int64)          # rand.nextInt(degree)
        path[step] = col[rp[cur] + k]
        if step % 2 == 0:
            hit = path[step] == des
            for i in range(step // 2):                                       # isFirstMeet :213-219
                hit &= path[i] != path[step - i]
            mid = path[step // 2][hit]
            total += float((C ** (step // 2) * deg[mid] / deg[des]).sum())
    return total / sample


@pytest.fixture(scope="module")
def o333():
    return S.load_multigraph(G333, 333, separator=" ")


def test_pair_probe_agrees_with_exact_and_with_the_oracle(o333):
    C, L, sample = 0.6, 3, 400000
    exact = S.simrank_exact_matrix(o333, C, L)
    rs = np.random.RandomState(7)
    rows = {}
    for src, des in PAIRS:
        est = pair_simrank_restated(o333, src, des, C, L, sample, rs)
        # one-sample standard deviation is below max increment ~ C * maxdeg/mindeg; bound it empirically by 6 sigma of
        # a Bernoulli-scaled increment: 6 * sqrt(exact * w_max / sample)
        w_max = C * np.diff(o333["row_ptr"]).max() / max(1, np.diff(o333["row_ptr"])[des])
        tol = 6.0 * np.sqrt(max(exact[src, des], 1e-6) * w_max / sample) + 1e-5
        assert abs(est - exact[src, des]) <= tol, (src, des, est, exact[src, des], tol)
        if src not in rows:
            rows[src], _, _ = S.single_random_walk_row(o333, src, 200000, L, C, seed_state=S.java_seed(100 + src))
        tol2 = 6.0 * np.sqrt(max(exact[src, des], 1e-6) * w_max / 200000) + tol
        assert abs(rows[src][des] - est) <= tol2, (src, des, rows[src][des], est)


@pytest.mark.gpu
def test_pair_probe_agrees_with_the_device_kernels(o333):
    from graph_embedding_b200 import simrank as sr
    C, L = 0.6, 3
    g = sr.Graph(G333, 333, separator=" ")
    exact = S.simrank_exact_matrix(o333, C, L)
    rs = np.random.RandomState(8)
    srcs = sorted({p[0] for p in PAIRS})
    dev = dict(zip(srcs, g.handle.simrank_rows(np.array(srcs, dtype=np.int64), C, L, 1000000, seed=21)))
    for src, des in PAIRS:
        est = pair_simrank_restated(o333, src, des, C, L, 400000, rs)
        w_max = C * np.diff(o333["row_ptr"]).max() / max(1, np.diff(o333["row_ptr"])[des])
        tol = 6.0 * np.sqrt(max(exact[src, des], 1e-6) * w_max / 400000) + 1e-5
        assert abs(dev[src][des] - est) <= 2 * tol, (src, des, dev[src][des], est)
        assert abs(dev[src][des] - exact[src, des]) <= tol
