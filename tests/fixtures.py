"""Deterministic inputs shared by the CPU and GPU suites.

sincos_dataset restates the reference's fixed fixture generator
(tests/recall.test.ts:26-54 un-normalised; tests/recall-common.ts:112-138 normalised):
    v[j] = sin(seed)*0.5 + cos(seed*0.7)*0.3, seed = i*1000 + j  (queries: i + 1000)
computed in f64 then stored as f32.  (V8 uses fdlibm sin/cos; glibc may differ by <= 1 ulp in
f64, which after the f32 store is almost always invisible — SURVEY §8c.)
"""
import numpy as np


def sincos_dataset(dim, base, queries, normalise=False):
    def gen(i0, n):
        i = (np.arange(n, dtype=np.float64)[:, None] + i0) * 1000.0
        seed = i + np.arange(dim, dtype=np.float64)[None, :]
        return (np.sin(seed) * 0.5 + np.cos(seed * 0.7) * 0.3).astype(np.float32)

    b, q = gen(0, base), gen(1000, queries)
    if normalise:
        from oracle import oracle as O
        b = np.stack([O.normalize_vector(r) for r in b])
        q = np.stack([O.normalize_vector(r) for r in q])
    return b, q


def gaussian(n, dim, seed):
    """Synthetic N(0,1) f32 rows."""
    return np.random.default_rng(seed).standard_normal((n, dim), dtype=np.float32)


def true_topk_cosine(q, base, k):
    """tests/recall.test.ts:111-117 ground truth: exact cosine, stable descending sort."""
    b = base.astype(np.float64)
    qq = q.astype(np.float64)
    s = (b @ qq) / (np.linalg.norm(b, axis=1) * np.linalg.norm(qq))
    return np.argsort(-s, kind="stable")[:k]


def edge_dataset(n, dim, nq, seed):
    """Rows and queries with the degenerate shapes the reference has to cope with: an all-zero row, duplicated rows
    (exact score ties: the order the reference's MinHeap returns them in is part of its behaviour), a constant row, a
    negated pair; a query equal to a duplicated row and an all-zero query."""
    base, queries = gaussian(n, dim, seed), gaussian(nq, dim, seed + 1)
    base[7] = 0
    base[9] = base[8]
    base[20] = base[8]
    base[11] = 0.5
    base[13] = -base[12]
    if nq > 1:
        queries[1] = base[8]
    if nq > 2:
        queries[2] = 0
    return base, queries
