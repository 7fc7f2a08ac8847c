"""The reference's recall suite on the GPU path: tests/recall-all-dimensions.test.ts with the configurations and
thresholds of tests/recall-common.ts:43-107 (normalised sin/cos fixture :112-138, lambda=0.001, iters=20, COSINE,
base 1000, 20 queries, k=10; 1-bit, 4-bit and 3x-oversampled 4-bit queries), and the cross-dimension trend check
(:92-137).  The quantised results must also equal the oracle's, list for list."""
import numpy as np
import pytest

from oracle import oracle as O
from tests.fixtures import sincos_dataset, true_topk_cosine
from tests.test_gpu_parity import bbq, make_format  # noqa: F401  (fixture)

pytestmark = pytest.mark.gpu

CONFIGS = {  # dim: (recallThreshold1bit, recallThreshold4bit, recallThresholdOversample)
    384: (0.60, 0.75, 0.80),
    768: (0.55, 0.70, 0.75),
    1024: (0.50, 0.65, 0.70),
    1536: (0.45, 0.60, 0.65),
}


def _recall(found, base, queries, k):
    return float(np.mean([len(set(np.asarray(f).tolist()) & set(true_topk_cosine(q, base, k).tolist())) / k
                          for f, q in zip(found, queries)]))


_results = {}


@pytest.mark.parametrize("dim", sorted(CONFIGS))
def test_recall_all_dimensions(bbq, dim):
    t1, t4, t_over = CONFIGS[dim]
    base, queries = sincos_dataset(dim, 1000, 20, normalise=True)
    for qb, thr in ((1, t1), (4, t4)):
        fmt = make_format(bbq, "COSINE", qb=qb, lam=0.001, iters=20)
        qv = fmt.quantizeVectors(base)["quantizedVectors"]
        idx, sc = fmt.searchBatch(queries, qv, 10)
        assert idx.shape == (20, 10) and np.all(np.diff(sc, axis=1) <= 0)
        r = _recall(idx, base, queries, 10)
        assert r >= thr, (dim, qb, r)
        if qb == 4:
            _results[dim] = r
            # and the lists are the oracle's (restated reference), not merely good enough
            oidx = O.quantize_vectors(base, sim="COSINE", index_bits=1, lam=0.001, iters=20)
            for j in (0, 7, 19):
                want, wsc = O.search_nearest_neighbors(queries[j], oidx, 10, query_bits=4, lam=0.001, iters=20,
                                                       mode="canonical")
                assert idx[j].tolist() == want.tolist()
                assert np.array_equal(sc[j].view(np.uint32), np.asarray(wsc, np.float32).view(np.uint32))
            # oversampled 4-bit search, exact cosine re-rank (recall-common.ts:254-289)
            fmt.attachOriginalVectors(qv, base)
            oi, _, ts = fmt.searchOversampledBatch(queries, qv, 10, 3)
            assert oi.shape == (20, 10) and np.all(np.diff(ts, axis=1) <= 0)
            ro = _recall(oi, base, queries, 10)
            assert ro >= t_over, (dim, "oversample", ro)
            assert ro >= r - 1e-12   # re-ranking a superset never loses true neighbours


def test_recall_trend_across_dimensions(bbq):
    """recall-all-dimensions.test.ts:92-137: recall at a higher dimension is at most 0.1 above the previous one."""
    if len(_results) < len(CONFIGS):
        pytest.skip("needs the per-dimension results of this module's first test")
    dims = sorted(_results)
    for a, b in zip(dims, dims[1:]):
        assert _results[b] <= _results[a] + 0.1
