"""Link-prediction score of an embedding for the probes under tools/: P(cos(u, v) of an edge > cos of a random non-adjacent pair)."""
import numpy as np


def edge_auc(vectors, row_ptr, col, rs, n_neg=20000):
    v = vectors / np.maximum(np.linalg.norm(vectors, axis=1, keepdims=True), 1e-12)
    n = len(row_ptr) - 1
    rows = np.repeat(np.arange(n), np.diff(row_ptr))
    pos = np.sort((v[rows] * v[col]).sum(axis=1))
    if len(pos) > 200000:
        pos = np.sort(pos[rs.randint(0, len(pos), size=200000)])
    key = rows.astype(np.int64) * n + col
    key.sort()
    a, b = rs.randint(0, n, size=n_neg), rs.randint(0, n, size=n_neg)
    k2 = a.astype(np.int64) * n + b
    hit = np.searchsorted(key, k2)
    adj = (hit < len(key)) & (key[np.minimum(hit, len(key) - 1)] == k2)
    ok = (a != b) & ~adj
    neg = (v[a[ok]] * v[b[ok]]).sum(axis=1)
    below = np.searchsorted(pos, neg, side="right")            # edges with cosine <= this non-edge's
    return float(1.0 - below.sum() / (len(pos) * len(neg)))
