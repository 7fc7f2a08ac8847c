"""GPU parity: the CUDA path, called through the C ABI (libbbq_b200.so) behind the reference-shaped host
interface, against the CPU oracle on the same seeded inputs.

Bars (BASELINE.json north_star): integer bit dot products BIT-EXACT; returned top-k index lists identical
(ties broken by lower index); corrected f32 scores within 1e-5 relative — in fact asserted BIT-EXACT here,
because the epilogue replays the reference's f64 operation order (the 1e-5 tolerance is also written out).
"""
import os

import numpy as np
import pytest

from oracle import oracle as O
from tests.fixtures import gaussian, sincos_dataset, true_topk_cosine
from tests.golden_util import golden_cases, load_golden

pytestmark = pytest.mark.gpu

REL_TOL = 1e-5  # north_star tolerance for corrected float scores
SIMS = ["EUCLIDEAN", "COSINE", "MAXIMUM_INNER_PRODUCT"]


@pytest.fixture(scope="module")
def bbq():
    import bbq_b200
    bbq_b200.build_library()
    return bbq_b200


def make_format(bbq, sim, qb=4, lam=0.1, iters=5, force_path=None, scan=None, qquant=None, dyntau=None,
                popc_form=None):
    """force_path / scan / qquant are test knobs read by bbq_create
    (BBQ_FORCE_PATH, BBQ_SCAN=popc|mma, BBQ_QQUANT=thread|warp|cta)."""
    saved = {k: os.environ.pop(k, None) for k in ("BBQ_FORCE_PATH", "BBQ_SCAN", "BBQ_QQUANT", "BBQ_DYNTAU", "BBQ_POPC_FORM")}
    if popc_form is not None:
        os.environ["BBQ_POPC_FORM"] = popc_form
    if dyntau is not None:
        os.environ["BBQ_DYNTAU"] = str(dyntau)
    if force_path is not None:
        os.environ["BBQ_FORCE_PATH"] = str(force_path)
    if scan is not None:
        os.environ["BBQ_SCAN"] = scan
    if qquant is not None:
        os.environ["BBQ_QQUANT"] = qquant
    try:
        return bbq.createBinaryQuantizationFormat(
            {"queryBits": qb, "indexBits": 1, "quantizer": {"similarityFunction": sim, "lambda": lam, "iters": iters}})
    finally:
        for k, v in saved.items():
            os.environ.pop(k, None)
            if v is not None:
                os.environ[k] = v


def bits_equal(a, b):
    a, b = np.ascontiguousarray(a), np.ascontiguousarray(b)
    assert a.dtype == b.dtype and a.shape == b.shape, (a.dtype, b.dtype, a.shape, b.shape)
    u = {4: np.uint32, 8: np.uint64, 1: np.uint8}[a.dtype.itemsize]
    return np.array_equal(a.view(u), b.view(u))


def assert_scores_close(got, want):
    ok = np.isclose(got.astype(np.float64), want.astype(np.float64), rtol=REL_TOL, atol=0) | (got == want) | \
        (np.isnan(got) & np.isnan(want))
    assert ok.all(), f"score rel err > {REL_TOL}"
    assert bits_equal(got, want), "scores are within tolerance but not bit-identical to the f64-replayed reference"


# ---- K5: index build --------------------------------------------------------------------------------
@pytest.mark.parametrize("sim", SIMS)
@pytest.mark.parametrize("n,dim", [(1000, 128), (257, 100), (3, 8), (5000, 768), (40000, 64)])
def test_index_build_bit_exact(bbq, sim, n, dim):
    rows = gaussian(n, dim, 11 + n + dim)
    want = O.quantize_vectors(rows, sim=sim, want_unpacked=False)
    fmt = make_format(bbq, sim)
    qv = fmt.quantizeVectors(rows)["quantizedVectors"]
    assert qv.size() == n and qv.dimension() == dim
    assert bits_equal(qv.getCentroid(), want.centroid)
    assert qv.getCentroidDP() == O.centroid_dp(want.centroid)
    packed, corr = qv.exportAll()
    assert np.array_equal(packed, want.packed)
    assert bits_equal(corr, want.corr)
    # accessor forms of the reference interface (src/types.ts:32-49)
    assert np.array_equal(qv.vectorValue(n - 1), want.packed[n - 1])
    assert qv.getCorrectiveTerms(0)["lowerInterval"] == want.corr[0, 0]
    assert np.array_equal(qv.getUnpackedVector(0), np.unpackbits(want.packed[0])[:dim])


def test_index_build_explicit_centroid_and_list_input(bbq):
    rows = gaussian(300, 64, 5)
    cen = np.zeros(64, np.float32)
    want = O.quantize_vectors(rows, sim="EUCLIDEAN", centroid=cen, want_unpacked=False)
    fmt = make_format(bbq, "EUCLIDEAN")
    qv = fmt.quantizeVectors([r for r in rows], centroid=cen)["quantizedVectors"]   # Float32Array[] form
    packed, corr = qv.exportAll()
    assert np.array_equal(packed, want.packed) and bits_equal(corr, want.corr)


def test_index_build_degenerate_rows(bbq):
    # single vector == centroid -> constant centred vector -> NaN interval in the reference; zero vectors under COSINE
    for sim in SIMS:
        rows = gaussian(1, 32, 1)
        want = O.quantize_vectors(rows, sim=sim, want_unpacked=False)
        qv = make_format(bbq, sim).quantizeVectors(rows)["quantizedVectors"]
        packed, corr = qv.exportAll()
        assert np.array_equal(packed, want.packed) and bits_equal(corr, want.corr)
    rows = gaussian(50, 40, 2)
    rows[7] = 0
    rows[9] = rows[8]
    want = O.quantize_vectors(rows, sim="COSINE", want_unpacked=False)
    qv = make_format(bbq, "COSINE").quantizeVectors(rows)["quantizedVectors"]
    packed, corr = qv.exportAll()
    assert np.array_equal(packed, want.packed) and bits_equal(corr, want.corr)


# ---- K4: query quantisation ----------------------------------------------------------------------------
@pytest.mark.parametrize("sim", SIMS)
@pytest.mark.parametrize("qb", [1, 2, 4, 7, 8])
@pytest.mark.parametrize("qquant", [None, "cta", "warp", "thread"])   # default (CTA per query for a narrow batch), and each K4 form forced
def test_query_quantize_bit_exact(bbq, sim, qb, qquant):
    rows, qs = gaussian(400, 200, 21), gaussian(5, 200, 22)
    qs[3] = 0                      # zero query: norm == 0 branch under COSINE
    qs[4] = rows[:400].mean(0)     # ~ centroid: tiny centred vector
    fmt = make_format(bbq, sim, qb=qb, qquant=qquant)
    qv = fmt.quantizeVectors(rows)["quantizedVectors"]
    for q in qs:
        codes, corr = O.quantize_query_vector(q, qv.getCentroid(), sim=sim, query_bits=qb)
        got = fmt.quantizeQueryVector(q, qv)
        assert np.array_equal(got["quantizedQuery"], codes)
        gc = got["queryCorrections"]
        assert bits_equal(np.array([gc["lowerInterval"], gc["upperInterval"], gc["additionalCorrection"],
                                    gc["quantizedComponentSum"]]), corr)


# ---- K1: integer dots and corrected scores over the whole index ------------------------------------------
@pytest.mark.parametrize("sim", SIMS)
@pytest.mark.parametrize("qb", [1, 4, 8])
@pytest.mark.parametrize("n,dim", [(3000, 128), (1500, 100), (2000, 768), (700, 1536), (300, 2048), (900, 40)])
@pytest.mark.parametrize("form", [None, "tile"])   # streaming (register/shuffle) and shared-memory tile kernels
def test_qcdist_and_scores_bit_exact(bbq, sim, qb, n, dim, form):
    rows, qs = gaussian(n, dim, 31 + dim), gaussian(2, dim, 32 + dim)
    idx = O.quantize_vectors(rows, sim=sim, want_unpacked=False)
    fmt = make_format(bbq, sim, qb=qb, popc_form=form)
    qv = fmt.adoptQuantized(idx.packed, idx.corr, idx.centroid)
    for q in qs:
        _, _, alls, alld = O.search_nearest_neighbors(q, idx, 5, query_bits=qb, want_all=True)
        assert np.array_equal(fmt.debugQcDist(q, qv), alld)          # integers: bit-exact
        assert_scores_close(fmt.debugScores(q, qv), alls)


# ---- K1+K3: top-k ------------------------------------------------------------------------------------------
def _check_search(fmt, qv, idx, queries, k, qb, lam=0.1, iters=5):
    gi, gs = fmt.searchBatch(queries, qv, k)
    for qi, q in enumerate(queries):
        wi, ws, alls, _ = O.search_nearest_neighbors(q, idx, k, query_bits=qb, lam=lam, iters=iters,
                                                     mode="canonical", want_all=True)
        assert gi[qi].tolist() == wi.tolist(), f"query {qi}: top-{k} index list differs"
        assert_scores_close(gs[qi], ws)
        hi, _ = O.topk(alls, k, "heap")
        ties = len(np.unique(alls[wi])) < len(wi) or (len(wi) < len(alls) and np.sort(alls)[::-1][len(wi)] == ws[-1])
        if not ties:  # reference heap == canonical unless an exact f32 tie straddles the k-th place
            assert set(hi.tolist()) == set(wi.tolist())
        # single-query form returns the same row
        one = fmt.searchNearestNeighbors(q, qv, k)
        assert [r["index"] for r in one] == wi.tolist()


@pytest.mark.parametrize("sim", SIMS)
@pytest.mark.parametrize("n,dim,k,force", [
    (1000, 128, 10, None),       # BASELINE configs[0] shape (direct path)
    (20000, 256, 10, None),      # sampled threshold + filtered scan
    (20000, 256, 100, None),
    (20000, 256, 10, 2),         # exact chunked path forced
    (9000, 96, 100, 1),          # filtered path forced on a small index
    (33000, 100, 10, None),      # dim % 8 != 0, ragged last tile
])
def test_search_topk_identical(bbq, sim, n, dim, k, force):
    rows, qs = gaussian(n, dim, 41 + n), gaussian(6, dim, 42 + n)
    idx = O.quantize_vectors(rows, sim=sim, want_unpacked=False)
    fmt = make_format(bbq, sim, force_path=force)
    qv = fmt.quantizeVectors(rows)["quantizedVectors"]
    _check_search(fmt, qv, idx, qs, k, 4)
    if force is not None:
        assert fmt.stats()["last_path"] == force


@pytest.mark.parametrize("qb", [1, 8])
def test_search_other_query_bits(bbq, qb):
    rows, qs = gaussian(18000, 128, 51), gaussian(4, 128, 52)
    for sim in SIMS:
        idx = O.quantize_vectors(rows, sim=sim, want_unpacked=False)
        fmt = make_format(bbq, sim, qb=qb, popc_form="tile" if qb == 8 else None)
        qv = fmt.quantizeVectors(rows)["quantizedVectors"]
        _check_search(fmt, qv, idx, qs, 10, qb)


def test_search_ties_lower_index_first(bbq):
    """Duplicate rows give exact f32 ties; the contract is (score desc, index asc)."""
    rows = gaussian(40, 64, 61)
    rows = np.concatenate([rows] * 500)           # 20000 rows, every score repeated 500 times
    for sim in SIMS:
        idx = O.quantize_vectors(rows[:40], sim=sim, want_unpacked=False, centroid=np.zeros(64, np.float32))
        packed, corr = np.concatenate([idx.packed] * 500), np.concatenate([idx.corr] * 500)
        big = O.OracleIndex(idx.centroid, packed, None, corr, 64, sim, 1)
        for force in (None, 2):
            fmt = make_format(bbq, sim, force_path=force)
            qv = fmt.adoptQuantized(packed, corr, idx.centroid)
            for q in gaussian(3, 64, 62):
                wi, ws = O.search_nearest_neighbors(q, big, 25, mode="canonical")
                gi, gs = fmt.searchBatch(q[None], qv, 25)
                assert gi[0].tolist() == wi.tolist() and bits_equal(gs[0], ws)
                assert np.all(np.diff(gi[0][:20]) == 40)   # the best row's duplicates, ascending ids


@pytest.mark.parametrize("scan", ["popc", "mma"])
def test_candidate_overflow_falls_back_to_exact_path(bbq, scan):
    """Adversarial index: every row is the same vector, so every score ties and no sampled threshold can prune —
    each query's candidate list overflows and the exact chunked path must take over, still returning the canonical
    answer (the lowest row ids)."""
    n, dim, nq, k = 40000, 64, 64, 10
    row = gaussian(1, dim, 131)
    idx = O.quantize_vectors(np.concatenate([row, gaussian(3, dim, 132)]), sim="COSINE", want_unpacked=False,
                             centroid=np.zeros(dim, np.float32))
    packed, corr = np.repeat(idx.packed[:1], n, 0), np.repeat(idx.corr[:1], n, 0)
    big = O.OracleIndex(idx.centroid, packed, None, corr, dim, "COSINE", 1)
    fmt = make_format(bbq, "COSINE", scan=scan)
    qv = fmt.adoptQuantized(packed, corr, idx.centroid)
    qs = gaussian(nq, dim, 133)
    gi, gs = fmt.searchBatch(qs, qv, k)
    st = fmt.stats()
    assert st["last_overflow"] == 1 and st["last_path"] == 2
    for qi in (0, 17, 63):
        wi, ws = O.search_nearest_neighbors(qs[qi], big, k, mode="canonical")
        assert gi[qi].tolist() == wi.tolist() == list(range(k)) and bits_equal(gs[qi], ws)


def test_search_edge_cases(bbq):
    rows = gaussian(7, 32, 71)
    fmt = make_format(bbq, "COSINE")
    qv = fmt.quantizeVectors(rows)["quantizedVectors"]
    q = gaussian(1, 32, 72)[0]
    assert fmt.searchNearestNeighbors(q, qv, 0) == []                     # k == 0 -> []
    res = fmt.searchNearestNeighbors(q, qv, 10)                           # k > N -> N results
    idx = O.quantize_vectors(rows, sim="COSINE")
    wi, ws = O.search_nearest_neighbors(q, idx, 10, mode="canonical")
    assert [r["index"] for r in res] == wi.tolist() and len(res) == 7
    assert [np.float32(r["score"]) for r in res] == ws.tolist()
    with pytest.raises(bbq.BbqError, match="k值不能为负数"):
        fmt.searchNearestNeighbors(q, qv, -1)
    with pytest.raises(bbq.BbqError, match="查询向量维度与目标向量维度不匹配"):
        fmt.searchNearestNeighbors(q[:16], qv, 3)
    with pytest.raises(bbq.BbqError, match="查询向量不能为空"):
        fmt.searchNearestNeighbors(None, qv, 3)
    with pytest.raises(bbq.BbqError, match="向量集合不能为空"):
        fmt.quantizeVectors([])
    with pytest.raises(bbq.BbqError, match="维度"):
        fmt.quantizeVectors([np.zeros(4, np.float32), np.zeros(5, np.float32)])
    bad = rows.copy()
    bad[3, 5] = np.nan
    with pytest.raises(bbq.BbqError, match="包含NaN值"):
        fmt.quantizeVectors(bad)
    bad = rows.copy()
    bad[2, 1] = np.inf
    with pytest.raises(bbq.BbqError, match="向量 2 位置 1 包含Infinity值"):
        make_format(bbq, "EUCLIDEAN").quantizeVectors(bad)
    qn = q.copy()
    qn[0] = np.nan
    with pytest.raises(bbq.BbqError, match="包含NaN值"):
        fmt.searchNearestNeighbors(qn, qv, 3)
    # single-row index
    one = make_format(bbq, "EUCLIDEAN").quantizeVectors(rows[:1])["quantizedVectors"]
    assert one.size() == 1


def test_quick_search_and_recall_fixture(bbq):
    """tests/recall.test.ts: 128-d sin/cos fixture, lambda=0.001, iters=20, COSINE: recall@10 >= 0.60 (4b x 1b)
    and >= 0.70 (1b x 1b); quickSearch == oracle quick_search."""
    base, queries = sincos_dataset(128, 100, 10)
    for qb, thr in ((4, 0.60), (1, 0.70)):
        fmt = make_format(bbq, "COSINE", qb=qb, lam=0.001, iters=20)
        qv = fmt.quantizeVectors(base)["quantizedVectors"]
        tot = 0.0
        for q in queries:
            got = [r["index"] for r in fmt.searchNearestNeighbors(q, qv, 10)]
            assert len(got) == 10
            tot += len(set(got) & set(true_topk_cosine(q, base, 10).tolist())) / 10
        assert tot / len(queries) >= thr
    res = bbq.quickSearch(queries[0], base, 10, bbq.VectorSimilarityFunction.COSINE)
    wi, ws = O.quick_search(queries[0], base, 10, "COSINE")
    assert [r["index"] for r in res] == wi.tolist()
    scores = [r["score"] for r in res]
    assert scores == sorted(scores, reverse=True)


@pytest.mark.parametrize("name", golden_cases())
def test_against_golden_fixtures(bbq, name):
    g = load_golden(name)
    fmt = make_format(bbq, g["sim"], qb=g["query_bits"], lam=g["lam"], iters=g["iters"])
    qv = fmt.quantizeVectors(g["base"])["quantizedVectors"]
    assert bits_equal(qv.getCentroid(), g["centroid"])
    packed, corr = qv.exportAll()
    assert np.array_equal(packed[:16], g["packed_head"]) and bits_equal(corr[:16], g["corr_head"])
    assert np.frombuffer(packed.tobytes(), np.uint8).astype(np.uint64).sum() == g["packed_crc"]
    assert np.bitwise_xor.reduce(corr.view(np.uint64).ravel()) == g["corr_bits_xor"]
    gi, gs = fmt.searchBatch(g["queries"], qv, g["k"])
    assert np.array_equal(gi, g["top_idx"]) and bits_equal(gs, g["top_score"])
    for qi, q in enumerate(g["queries"]):
        assert np.array_equal(fmt.debugQcDist(q, qv)[:32], g["dots_head"][qi])
        assert np.bitwise_xor.reduce(fmt.debugScores(q, qv).view(np.uint32)) == g["score_xor"][qi]
        got = fmt.quantizeQueryVector(q, qv)
        assert np.array_equal(got["quantizedQuery"], g["qcodes"][qi])


@pytest.mark.parametrize("dim", [8, 33, 1024, 1536])
def test_query_quantize_dims_and_long_iterations(bbq, dim):
    rows, qs = gaussian(64, dim, 23 + dim), gaussian(3, dim, 24 + dim)
    for sim in SIMS:
        fmt = make_format(bbq, sim, qb=4, lam=0.001, iters=20)
        qv = fmt.quantizeVectors(rows)["quantizedVectors"]
        for q in qs:
            codes, corr = O.quantize_query_vector(q, qv.getCentroid(), sim=sim, query_bits=4, lam=0.001, iters=20)
            got = fmt.quantizeQueryVector(q, qv)
            gc = got["queryCorrections"]
            assert np.array_equal(got["quantizedQuery"], codes)
            assert bits_equal(np.array([gc["lowerInterval"], gc["upperInterval"], gc["additionalCorrection"],
                                        gc["quantizedComponentSum"]]), corr)


def test_batch_equals_single_and_is_deterministic(bbq):
    rows, qs = gaussian(30000, 128, 81), gaussian(70, 128, 82)
    fmt = make_format(bbq, "EUCLIDEAN")
    qv = fmt.quantizeVectors(rows)["quantizedVectors"]
    gi, gs = fmt.searchBatch(qs, qv, 10)
    gi2, gs2 = fmt.searchBatch(qs, qv, 10)
    assert np.array_equal(gi, gi2) and bits_equal(gs, gs2)
    for qi in (0, 33, 69):
        one = fmt.searchNearestNeighbors(qs[qi], qv, 10)
        assert [r["index"] for r in one] == gi[qi].tolist()


# ---- K2: the tcgen05 batched scan must give the same answers as the popcount scan and the oracle -------------
@pytest.mark.parametrize("sim", SIMS)
@pytest.mark.parametrize("n,dim,nq,k,qb", [
    (20000, 256, 70, 10, 4),       # one pass, ragged query block
    (33000, 100, 40, 10, 4),       # dim % 128 != 0, ragged last tile
    (40000, 1024, 300, 10, 4),     # two passes of resident query blocks (BASELINE dims)
    (20000, 768, 64, 100, 4),      # k = 100
    (20000, 128, 96, 10, 1),       # 1-bit queries
    (20000, 128, 96, 10, 5),       # widest query the weighted expansion supports
    (30000, 384, 80, 10, 4),       # odd number of 128-dim chunks (3): the two MMA issuers split 2 + 1
    (26000, 640, 130, 10, 4),      # 5 chunks, two passes
])
def test_mma_scan_matches_oracle_and_popcount(bbq, sim, n, dim, nq, k, qb):
    rows, qs = gaussian(n, dim, 101 + n + dim), gaussian(nq, dim, 102 + n)
    idx = O.quantize_vectors(rows, sim=sim, want_unpacked=False)
    fm = make_format(bbq, sim, qb=qb, scan="mma")
    fp = make_format(bbq, sim, qb=qb, scan="popc")
    qm = fm.adoptQuantized(idx.packed, idx.corr, idx.centroid)
    qp = fp.adoptQuantized(idx.packed, idx.corr, idx.centroid)
    mi, ms = fm.searchBatch(qs, qm, k)
    pi, ps = fp.searchBatch(qs, qp, k)
    assert fm.stats()["last_engine"] == 2 and fm.stats()["last_path"] == 1 and fm.stats()["last_overflow"] == 0
    assert fp.stats()["last_engine"] == 1
    assert np.array_equal(mi, pi) and bits_equal(ms, ps)
    for qi in range(0, nq, max(1, nq // 6)):
        wi, ws = O.search_nearest_neighbors(qs[qi], idx, k, query_bits=qb, mode="canonical")
        assert mi[qi].tolist() == wi.tolist() and bits_equal(ms[qi], ws)


@pytest.mark.parametrize("sim", SIMS)
def test_mma_running_threshold_does_not_change_results(bbq, sim):
    """The in-kernel tightening of tau only prunes work: same answers with it off, far fewer candidates with it on."""
    rows, qs = gaussian(150000, 128, 121), gaussian(128, 128, 122)
    on = make_format(bbq, sim, scan="mma")
    off = make_format(bbq, sim, scan="mma", dyntau=0)
    cen = np.zeros(128, np.float32)
    qa = on.quantizeVectors(rows, centroid=cen)["quantizedVectors"]
    qb_ = off.quantizeVectors(rows, centroid=cen)["quantizedVectors"]
    ai, asc = on.searchBatch(qs, qa, 10)
    bi, bsc = off.searchBatch(qs, qb_, 10)
    assert np.array_equal(ai, bi) and bits_equal(asc, bsc)
    assert on.stats()["last_engine"] == 2 and off.stats()["last_engine"] == 2
    assert on.stats()["last_candidates"] <= off.stats()["last_candidates"]
    pp = make_format(bbq, sim, scan="popc")
    qc = pp.quantizeVectors(rows, centroid=cen)["quantizedVectors"]
    ci, csc = pp.searchBatch(qs, qc, 10)
    assert np.array_equal(ai, ci) and bits_equal(asc, csc)


def test_mma_scan_ties_and_degenerate_rows(bbq):
    """Exact f32 ties (duplicate rows) and rows whose interval is degenerate (always sent to the exact replay)."""
    rows = gaussian(50, 128, 111)
    rows[7] = 0
    rows = np.concatenate([rows] * 400)     # 20000 rows, every score 400 times
    for sim in SIMS:
        idx = O.quantize_vectors(rows[:50], sim=sim, want_unpacked=False, centroid=np.zeros(128, np.float32))
        corr = idx.corr.copy()
        corr[11, 1] = corr[11, 0]           # lx == 0
        corr[12, 1] = np.nan                # non-finite interval
        packed, corr = np.concatenate([idx.packed] * 400), np.concatenate([corr] * 400)
        big = O.OracleIndex(idx.centroid, packed, None, corr, 128, sim, 1)
        fm = make_format(bbq, sim, scan="mma")
        qm = fm.adoptQuantized(packed, corr, idx.centroid)
        qs = gaussian(64, 128, 112)
        mi, ms = fm.searchBatch(qs, qm, 30)
        assert fm.stats()["last_engine"] == 2
        for qi in (0, 13, 63):
            wi, ws = O.search_nearest_neighbors(qs[qi], big, 30, mode="canonical")
            assert mi[qi].tolist() == wi.tolist() and bits_equal(ms[qi], ws)


# ---- "next" row: oversampled search + exact re-rank (src/topKSelector.ts) ---------------------------------------
@pytest.mark.parametrize("n,dim,k,factor", [(100, 128, 10, 3), (5000, 96, 10, 3), (30000, 256, 20, 5), (7, 32, 10, 3)])
def test_oversampled_rerank_matches_oracle(bbq, n, dim, k, factor):
    if n == 100:
        rows, qs = sincos_dataset(128, 100, 10)
        lam, iters = 0.001, 20
    else:
        rows, qs = gaussian(n, dim, 141 + n), gaussian(6, dim, 142 + n)
        lam, iters = 0.1, 5
    idx = O.quantize_vectors(rows, sim="COSINE", want_unpacked=False, lam=lam, iters=iters)
    fmt = make_format(bbq, "COSINE", lam=lam, iters=iters)
    qv = fmt.quantizeVectors(rows)["quantizedVectors"]
    fmt.attachOriginalVectors(qv, rows)
    gi, gq, gt = fmt.searchOversampledBatch(qs, qv, k, factor)
    rec = 0.0
    for qi, q in enumerate(qs):
        wi, wq, wt = O.oversampled_topk(q, rows, idx, k, factor, lam=lam, iters=iters)
        assert gi[qi].tolist() == wi.tolist()
        assert bits_equal(gq[qi], wq) and bits_equal(gt[qi], wt)      # f32 quantised scores, f64 true scores
        rec += len(set(gi[qi].tolist()) & set(true_topk_cosine(q, rows, min(k, n)).tolist())) / min(k, n)
    if n == 100:
        assert rec / len(qs) >= 0.75                                   # tests/recall.test.ts:519,635
        res = bbq.getOversampledTopKWithSort(qs[0], qv, rows, 10, 3, fmt)
        assert [r["index"] for r in res] == gi[0].tolist() and res[0]["trueScore"] == gt[0][0]


# ---- sharding: G shards + deterministic merge == one index --------------------------------------
@pytest.mark.parametrize("shards", [2, 3, 8])
def test_sharded_merge_equals_single(bbq, shards):
    import ctypes as C
    import torch
    n, dim, k, nq = 48000, 128, 10, 9   # 2 shards -> filtered path, 3/8 shards -> direct path
    rows, qs = gaussian(n, dim, 91), gaussian(nq, dim, 92)
    sim = "COSINE"
    idx = O.quantize_vectors(rows, sim=sim, want_unpacked=False)
    fmt = make_format(bbq, sim)
    L = bbq._native.load()
    bounds = np.linspace(0, n, shards + 1).astype(int)
    dq = torch.from_numpy(qs).cuda()
    all_idx = torch.empty((shards, nq, k), dtype=torch.int32, device="cuda")
    all_sc = torch.empty((shards, nq, k), dtype=torch.float32, device="cuda")
    keep = []
    for s in range(shards):
        a, b = bounds[s], bounds[s + 1]
        qv = fmt.adoptQuantized(idx.packed[a:b], idx.corr[a:b], idx.centroid)   # same centroid on every shard
        assert L.bbq_index_set_base(qv._h, int(a)) == 0
        st = L.bbq_search_device(qv._h, dq.data_ptr(), nq, k, all_idx[s].data_ptr(), all_sc[s].data_ptr(), None)
        assert st == 0, L.bbq_last_error()
        keep.append(qv)
    torch.cuda.synchronize()
    out_idx = torch.empty((nq, k), dtype=torch.int32, device="cuda")
    out_sc = torch.empty((nq, k), dtype=torch.float32, device="cuda")
    st = L.bbq_merge_topk_device(fmt._ctx, all_idx.data_ptr(), all_sc.data_ptr(), shards, nq, k, out_idx.data_ptr(),
                                 out_sc.data_ptr(), None)
    assert st == 0, L.bbq_last_error()
    torch.cuda.synchronize()
    gi, gs = out_idx.cpu().numpy(), out_sc.cpu().numpy()
    for qi, q in enumerate(qs):
        wi, ws = O.search_nearest_neighbors(q, idx, k, mode="canonical")
        assert gi[qi].tolist() == wi.tolist() and bits_equal(gs[qi], ws)
