"""CPU oracle of the skip-gram consumer of the walk corpus.  TEST INFRASTRUCTURE ONLY.

Restates the algorithm behind the reference's ``learn_embeddings`` (node2vec/src/main.py:92-101):

    Word2Vec(walks, size=dimensions, window=window_size, min_count=0, sg=1, workers=workers, iter=iter)

i.e. gensim 0.13.3 (node2vec/requirements.txt:3), skip-gram with negative sampling and gensim's defaults
(alpha=0.025, min_alpha=0.0001, sample=1e-3, negative=5, hs=0).  gensim is a third-party dependency that is NOT under
/root/reference and is not installed in this image, so its published algorithm (gensim/models/word2vec.py
``scale_vocab`` / ``make_cum_table`` / ``train_batch_sg`` and word2vec_inner.pyx ``fast_sentence_sg_neg``) is restated
from its source as published:

  * ``scale_vocab``: threshold = sample * total; p(w) = (sqrt(cnt/threshold) + 1) * threshold / cnt, capped at 1;
    ``sample_int = round(p * 2**32)``; during training a word is skipped when ``sample_int < random 32-bit number``.
  * ``make_cum_table``: negatives are drawn with probability proportional to cnt**0.75.
  * ``train_batch_sg``: per sentence, after subsampling, for every position i: ``b = random % window``; every j in
    ``[max(0, i - window + b), min(len, i + window + 1 - b))``, j != i, trains the pair (word = sent[i], word2 = sent[j]).
  * ``fast_sentence_sg_neg``: row1 = syn0[word2]; target 0 is ``word`` with label 1, then ``negative`` draws with label 0
    (a draw equal to ``word`` is skipped); f = row1 . syn1neg[target]; ``|f| >= MAX_EXP (6)`` skips the target;
    g = (label - EXP_TABLE[int((f + 6) * (1000 / 6 / 2))]) * alpha; work += g * syn1neg[target];
    syn1neg[target] += g * row1; after the targets row1 += work.
  * alpha decreases linearly from ``alpha`` to ``min_alpha`` with the raw words processed.
  * syn0 starts as (uniform - 0.5) / size, syn1neg as zeros.

Parity status: **parity unpinned** -- the reference ships no trained embedding (node2vec/emb/karate.emb holds
near-initial values of unknown p/q) and gensim's own random streams (per-thread LCG seeds, numpy ``randint`` for the
windows, thread interleaving) are not reproducible even by gensim itself.  The oracle therefore takes its random numbers
from the SAME counter-based generator as the device (Philox4x32-10 keyed by (seed, sentence id); one 48-bit LCG per
sentence for windows and negatives, the recurrence gensim uses) and restates the arithmetic in the device's operation order
(fp32, fused multiply-adds emulated through fp64, the 32-lane butterfly of the dot product), so that ONE device warp
walking the sentences in order must reproduce it to rounding -- that pins the kernel to this restatement; the restatement
to gensim is pinned only by reading the two side by side.  Only ``tests/`` may import this file.
"""
import numpy as np

M0, M1, W0, W1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85
MAX_EXP, EXP_TABLE_SIZE = 6.0, 1000


def philox4x32(ctr, key):
    """Philox4x32-10 (Salmon et al., SC'11): ctr = 4 uint32, key = 2 uint32 -> 4 uint32."""
    c = [int(x) & 0xFFFFFFFF for x in ctr]
    k0, k1 = int(key[0]) & 0xFFFFFFFF, int(key[1]) & 0xFFFFFFFF
    for _ in range(10):
        p0, p1 = M0 * c[0], M1 * c[2]
        c = [((p1 >> 32) ^ c[1] ^ k0) & 0xFFFFFFFF, p1 & 0xFFFFFFFF, ((p0 >> 32) ^ c[3] ^ k1) & 0xFFFFFFFF, p0 & 0xFFFFFFFF]
        k0, k1 = (k0 + W0) & 0xFFFFFFFF, (k1 + W1) & 0xFFFFFFFF
    return c


def exp_table():
    """word2vec_inner.pyx init(): EXP_TABLE[i] = exp((i / 1000 * 2 - 1) * 6) / (exp(..) + 1), in fp32."""
    t = np.empty(EXP_TABLE_SIZE, dtype=np.float32)
    for i in range(EXP_TABLE_SIZE):
        e = np.float32(np.exp((i / float(EXP_TABLE_SIZE) * 2.0 - 1.0) * MAX_EXP))
        t[i] = e / (e + np.float32(1.0))
    return t


def scale_vocab(counts, sample):
    """-> sample_int per word (uint32).  counts: raw occurrences per word (0 for words that never occur)."""
    counts = np.asarray(counts, dtype=np.float64)
    total = counts.sum()
    keep = np.full(len(counts), 0xFFFFFFFF, dtype=np.uint32)
    if sample > 0:
        thr = sample * total
        with np.errstate(divide="ignore", invalid="ignore"):
            p = (np.sqrt(counts / thr) + 1.0) * (thr / counts)
        p = np.where(counts > 0, np.minimum(p, 1.0), 1.0)
        si = np.rint(p * 4294967296.0)
        keep = np.where(si >= 4294967295.0, 0xFFFFFFFF, si).astype(np.uint32)
    return keep


def negative_table(counts, size):
    """Inverse lookup of make_cum_table: word i owns the slots [round(cum(i-1)/total * size), round(cum(i)/total * size))."""
    pw = np.where(np.asarray(counts) > 0, np.asarray(counts, dtype=np.float64) ** 0.75, 0.0)
    cum = np.cumsum(pw)                                          # left to right, fp64
    total = cum[-1]
    edges = np.rint(np.concatenate([[0.0], cum]) / total * size).astype(np.int64)
    edges[-1] = size
    tab = np.empty(size, dtype=np.int32)
    for i in range(len(counts)):
        tab[edges[i]:edges[i + 1]] = i
    return tab


def table_size(n):
    t = 1 << 20
    while t < 32 * n and t < (1 << 28):
        t <<= 1
    return t


def init_vectors(n, dim, seed):
    """syn0 = (uniform - 0.5) / size from Philox(counter = index of the 4-float group, key = seed); syn1neg = 0."""
    key = (seed & 0xFFFFFFFF, seed >> 32)
    total = n * dim
    syn0 = np.zeros(total, dtype=np.float32)
    for i in range((total + 3) // 4):
        r = philox4x32((i & 0xFFFFFFFF, i >> 32, 0x5347, 0), key)
        for k in range(4):
            if 4 * i + k < total:
                syn0[4 * i + k] = (np.float32(r[k] >> 8) * np.float32(1.0 / 16777216.0) - np.float32(0.5)) / np.float32(dim)
    return syn0.reshape(n, dim), np.zeros((n, dim), dtype=np.float32)


def _fma32(a, b, c):
    """fmaf on float32 arrays: one rounding of a*b + c (the product is exact in fp64)."""
    return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(np.float32)


def _warp_dot(x, y):
    """The device's dot product: every lane accumulates its dim/32 consecutive elements with fmaf, then the 32 partial
    sums are combined by the xor butterfly (offsets 16, 8, 4, 2, 1)."""
    per = len(x) // 32
    xs, ys = x.reshape(32, per), y.reshape(32, per)
    f = np.zeros(32, dtype=np.float32)
    for k in range(per):
        f = _fma32(xs[:, k], ys[:, k], f)
    idx = np.arange(32)
    for o in (16, 8, 4, 2, 1):
        f = (f + f[idx ^ o]).astype(np.float32)
    return f[0]


def train(walks, syn0, syn1, keep, negtab, window, negative, alpha, min_alpha, words_before, total_words, seed,
          sentence_id_base=0, subsample=True):
    """train_batch_sg over `walks` ([n_walks, L] int, -1 padded), sentences in order, in place.  Returns the number of
    trained (word, word2) pairs."""
    table = exp_table()
    key = (seed & 0xFFFFFFFF, seed >> 32)
    n_walks, L = walks.shape
    mask = len(negtab) - 1
    pairs = 0
    for s in range(n_walks):
        sid = sentence_id_base + s
        sent = []
        for pos in range(L):
            w = int(walks[s, pos])
            if w < 0:
                continue
            if subsample:
                r = philox4x32((sid & 0xFFFFFFFF, sid >> 32, pos >> 2, 0x5342), key)[pos & 3]
                if int(keep[w]) < r:
                    continue
            sent.append(w)
        m = len(sent)
        if m < 2:
            continue
        prog = min(1.0, (words_before + float(s) * float(L)) / total_words)
        al = np.float32(max(float(min_alpha), float(alpha) - (float(alpha) - float(min_alpha)) * prog))
        r = philox4x32((sid & 0xFFFFFFFF, sid >> 32, 0, 0x5353), key)
        rs = (r[0] << 32) | r[1]

        def nxt():
            nonlocal rs
            rs = (rs * 25214903917 + 11) & 0xFFFFFFFFFFFFFFFF
            return (rs >> 16) & 0xFFFFFFFF
        for i in range(m):
            center = sent[i]
            b = nxt() % window
            for j in range(max(0, i - window + b), min(m, i + window + 1 - b)):
                if j == i:
                    continue
                x = syn0[sent[j]].copy()
                work = np.zeros_like(x)
                for d in range(negative + 1):
                    target, label = center, np.float32(1.0)
                    if d > 0:
                        target = int(negtab[nxt() & mask])
                        if target == center:
                            continue
                        label = np.float32(0.0)
                    y = syn1[target].copy()
                    f = _warp_dot(x, y)
                    if f <= -MAX_EXP or f >= MAX_EXP:
                        continue
                    sig = table[int(np.float32(f + np.float32(MAX_EXP)) * np.float32(EXP_TABLE_SIZE / MAX_EXP / 2.0))]
                    g = np.float32(np.float32(label - sig) * al)
                    gv = np.full_like(x, g)
                    work = _fma32(gv, y, work)
                    syn1[target] = _fma32(gv, x, y)
                syn0[sent[j]] = (x + work).astype(np.float32)
                pairs += 1
    return pairs


def edge_auc(vectors, row_ptr, col, rs, n_neg=20000):
    """Quality of an embedding as a link predictor: P(cos(u, v) of an edge > cos of a random non-adjacent pair)."""
    v = vectors / np.maximum(np.linalg.norm(vectors, axis=1, keepdims=True), 1e-12)
    n = len(row_ptr) - 1
    rows = np.repeat(np.arange(n), np.diff(row_ptr))
    pos = (v[rows] * v[col]).sum(axis=1)
    adj = set(zip(rows.tolist(), col.tolist()))
    a, b = rs.randint(0, n, size=n_neg), rs.randint(0, n, size=n_neg)
    ok = np.array([x != y and (x, y) not in adj for x, y in zip(a.tolist(), b.tolist())])
    neg = (v[a[ok]] * v[b[ok]]).sum(axis=1)
    return float((pos[:, None] > neg[None, :]).mean()) if len(pos) * len(neg) < 5e7 else \
        float(np.mean([(p > neg).mean() for p in pos[:: max(1, len(pos) // 2000)]]))
