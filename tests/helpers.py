import json
import os

import numpy as np

from conftest import GOLDEN, DATA

CASES = json.load(open(os.path.join(GOLDEN, "n2v_cases.json")))


def case(name):
    return [c for c in CASES if c["name"] == name][0]


def load_npz(meta):
    return np.load(os.path.join(GOLDEN, "n2v_%s.npz" % meta["name"]))


def data_path(meta):
    return os.path.join(DATA, meta["file"])


def chi2_transitions(walks, lens, g, law_first, law_second, min_expected=5.0):
    """Pooled chi-square of observed next-vertex counts against the exact first/second-order law.
    walks: int32 [n, L] dense ids (-1 padded).  Returns (chi2, df)."""
    n = len(g["row_ptr"]) - 1
    rp, col = g["row_ptr"], g["col_idx"]
    chi2, df = 0.0, 0
    # first steps
    ok = lens >= 2
    key = walks[ok, 0].astype(np.int64) * n + walks[ok, 1]
    cnt = np.bincount(key, minlength=n * n).reshape(n, n)
    for cur in range(n):
        tot = cnt[cur].sum()
        if tot == 0 or rp[cur + 1] == rp[cur]:
            continue
        nb = col[rp[cur]:rp[cur + 1]]
        assert cnt[cur].sum() == cnt[cur][nb].sum(), "step to a non-neighbour"
        c, d = _chi(cnt[cur][nb], law_first(cur) * tot, min_expected)
        chi2 += c; df += d
    # second-order steps (sparse counting: contexts are the directed edges)
    L = walks.shape[1]
    keys = []
    for i in range(2, L):
        ok = lens > i
        keys.append((walks[ok, i - 2].astype(np.int64) * n + walks[ok, i - 1]) * n + walks[ok, i])
    keys = np.concatenate(keys)
    uk, uc = np.unique(keys, return_counts=True)
    ctx = uk // n
    nxt = uk % n
    bounds = np.flatnonzero(np.diff(ctx)) + 1
    for lo, hi in zip(np.r_[0, bounds], np.r_[bounds, len(ctx)]):
        prev, cur = int(ctx[lo] // n), int(ctx[lo] % n)
        nb = col[rp[cur]:rp[cur + 1]]
        pos = np.searchsorted(nb, nxt[lo:hi])
        assert (pos < len(nb)).all() and (nb[pos] == nxt[lo:hi]).all(), "step to a non-neighbour"
        obs = np.zeros(len(nb))
        obs[pos] = uc[lo:hi]
        c, d = _chi(obs, law_second(prev, cur) * obs.sum(), min_expected)
        chi2 += c; df += d
    return chi2, df


def _chi(obs, exp, min_expected):
    obs = np.asarray(obs, dtype=np.float64)
    exp = np.asarray(exp, dtype=np.float64)
    big = exp >= min_expected
    if big.sum() < 1:
        return 0.0, 0
    o = np.append(obs[big], obs[~big].sum())
    e = np.append(exp[big], exp[~big].sum())
    if e[-1] < min_expected:       # fold the small remainder into the smallest kept cell
        if len(o) == 2 and e[-1] == 0:
            return 0.0, 0
        j = int(np.argmin(e[:-1]))
        o[j] += o[-1]; e[j] += e[-1]
        o, e = o[:-1], e[:-1]
    if len(o) < 2:
        return 0.0, 0
    return float(((o - e) ** 2 / e).sum()), len(o) - 1
