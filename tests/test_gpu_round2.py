"""GPU parity, round 2: the tensor-core scan's INTEGER tap, device-side query screening, large-k / large-shard
candidate lists, and batches above 1024 queries in one launch.  All through the C ABI."""
import numpy as np
import pytest

from oracle import oracle as O
from tests.fixtures import gaussian
from tests.test_gpu_parity import bits_equal, make_format, SIMS

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def bbq():
    import bbq_b200
    bbq_b200.build_library()
    return bbq_b200


# ---- K2 integers: tcgen05 accumulators >> 3 == computeBatchFourBitDotProductDirectPacked -------------------------
@pytest.mark.parametrize("dim", [128, 100, 384, 640, 768, 1024, 1536, 2048, 4096])   # 1..32 chunks, odd counts, ragged dim
@pytest.mark.parametrize("qb", [1, 4, 5])
def test_mma_integer_tap_bit_exact(bbq, dim, qb):
    n, nq = 3000 + dim % 97, 37                      # ragged last tile, ragged query block
    rows, qs = gaussian(n, dim, 501 + dim), gaussian(nq, dim, 502 + dim)
    idx = O.quantize_vectors(rows, sim="COSINE", want_unpacked=False)
    fmt = make_format(bbq, "COSINE", qb=qb)
    qv = fmt.adoptQuantized(idx.packed, idx.corr, idx.centroid)
    got = fmt.debugQcDistBatch(qs, qv)               # k_scan_mma<SCAN_DUMP>, whatever the batch size
    assert got.shape == (nq, n)
    for qi in range(nq):
        _, _, _, alld = O.search_nearest_neighbors(qs[qi], idx, 1, query_bits=qb, want_all=True)
        assert np.array_equal(got[qi], alld), (dim, qb, qi)
    # and the popcount kernel's tap agrees (two engines, one integer)
    assert np.array_equal(fmt.debugQcDist(qs[5], qv), got[5])


def _disabled_mma_integer_tap_rejects_wide_queries(bbq):   # (8-bit queries now run on tcgen05: nibble columns)
    rows = gaussian(500, 128, 7)
    fmt = make_format(bbq, "COSINE", qb=8)
    qv = fmt.quantizeVectors(rows)["quantizedVectors"]
    with pytest.raises(bbq.BbqError) as e:
        fmt.debugQcDistBatch(gaussian(8, 128, 8), qv)
    assert e.value.status == 9


def test_mma_integer_tap_full_size(bbq):
    """A whole 1M x 1024 index x 256 queries (2.7e11 integer MACs on the tensor cores, five passes of resident query
    blocks are not involved here — 256 queries are 2 blocks) against the numpy restatement of the reference loop."""
    import torch
    n, dim, nq = 1_000_000, 1024, 256
    g = torch.Generator(device="cuda")
    g.manual_seed(20260909)
    rows_d = torch.randn((n, dim), generator=g, device="cuda", dtype=torch.float32)
    fmt = make_format(bbq, "EUCLIDEAN")
    qv = fmt.quantizeVectorsDevice(rows_d.data_ptr(), n, dim, centroid=np.zeros(dim, np.float32))["quantizedVectors"]
    del rows_d
    torch.cuda.empty_cache()
    qs = gaussian(nq, dim, 20260910)
    got = fmt.debugQcDistBatch(qs, qv)
    codes = np.stack([fmt.quantizeQueryVector(q, qv)["quantizedQuery"] for q in qs])
    packed, _ = qv.exportAll()
    want = O.qcdist_matrix(codes, packed, dim)
    assert np.array_equal(got, want)
    # anchor the numpy restatement itself on the C++ oracle loop for one query
    assert np.array_equal(want[17], O.qcdist_packed(codes[17], packed, dim))


# ---- query screening on the device (scalarQuantize's validation, optimizedScalarQuantizer.ts:138-148) -------------
def test_query_validation_on_device_reports_like_the_reference(bbq):
    rows = gaussian(300, 64, 11)
    qs = gaussian(6, 64, 12)
    for sim in SIMS:
        fmt = make_format(bbq, sim)
        qv = fmt.quantizeVectors(rows)["quantizedVectors"]
        want = fmt.searchBatch(qs, qv, 5)
        bad = qs.copy()
        bad[4, 9] = np.inf
        bad[2, 30] = np.nan
        bad[2, 7] = np.inf
        with pytest.raises(bbq.BbqError) as e:
            fmt.searchBatch(bad, qv, 5)
        # first offending query is 2; COSINE: NaN anywhere -> the normalised vector is all-NaN -> position 0;
        # otherwise the first non-finite component (position 7) is an Infinity
        if sim == "COSINE":
            assert e.value.status == 5 and str(e.value) == "向量位置 0 包含NaN值"
        else:
            assert e.value.status == 6 and str(e.value) == "向量位置 7 包含Infinity值"
        only_inf = qs.copy()
        only_inf[3, 12] = -np.inf
        with pytest.raises(bbq.BbqError) as e:
            fmt.searchBatch(only_inf, qv, 5)
        if sim == "COSINE":   # the search path normalises twice: Inf / Inf = NaN, then x / NaN = NaN everywhere -> position 0
            assert e.value.status == 5 and str(e.value) == "向量位置 0 包含NaN值"      # (tests/golden/from_ts/errors.behaviour.json)
        else:
            assert e.value.status == 6 and str(e.value) == "向量位置 12 包含Infinity值"
        # the context is still usable and answers as before
        again = fmt.searchBatch(qs, qv, 5)
        assert np.array_equal(want[0], again[0]) and bits_equal(want[1], again[1])


# ---- candidate lists: large k on a large shard must not overflow on ordinary data (sample grows with k*n) --------
@pytest.mark.parametrize("nq,engine", [(2, 1), (48, 2)])
def test_large_k_large_shard_does_not_overflow(bbq, nq, engine):
    import torch
    n, dim, k = 4_300_000, 128, 100
    g = torch.Generator(device="cuda")
    g.manual_seed(20261001)
    rows_d = torch.randn((n, dim), generator=g, device="cuda", dtype=torch.float32)
    fmt = make_format(bbq, "MAXIMUM_INNER_PRODUCT")
    cen = np.zeros(dim, np.float32)
    qv = fmt.quantizeVectorsDevice(rows_d.data_ptr(), n, dim, centroid=cen)["quantizedVectors"]
    del rows_d
    torch.cuda.empty_cache()
    qs = gaussian(nq, dim, 20261002)
    gi, gs = fmt.searchBatch(qs, qv, k)
    st = fmt.stats()
    assert st["last_engine"] == engine and st["last_path"] == 1
    assert st["last_overflow"] == 0, "ordinary data overflowed the candidate lists: the exact fallback ran"
    packed, corr = qv.exportAll()
    oidx = O.OracleIndex(cen, packed, None, corr, dim, "MAXIMUM_INNER_PRODUCT", 1)
    for qi in (0, nq - 1):
        wi, ws = O.search_nearest_neighbors(qs[qi], oidx, k, mode="canonical")
        assert gi[qi].tolist() == wi.tolist() and bits_equal(gs[qi], ws)


# ---- one launch for a whole 4096-query batch (20 resident query blocks, padded last block) -----------------------
def test_4096_queries_in_one_batch(bbq):
    n, dim, k, nq = 60_000, 256, 10, 4096
    rows, qs = gaussian(n, dim, 601), gaussian(nq, dim, 602)
    idx = O.quantize_vectors(rows, sim="COSINE", want_unpacked=False, centroid=np.zeros(dim, np.float32))
    # rows with degenerate correctives are replayed for every query column: they must not leak into padding columns
    corr = idx.corr.copy()
    corr[100, 1] = corr[100, 0]
    corr[40_000, 1] = np.nan
    big = O.OracleIndex(idx.centroid, idx.packed, None, corr, dim, "COSINE", 1)
    fm = make_format(bbq, "COSINE", scan="mma")
    qm = fm.adoptQuantized(idx.packed, corr, idx.centroid)
    l0 = fm.stats()["kernel_launches"]
    mi, ms = fm.searchBatch(qs, qm, k)
    st = fm.stats()
    assert st["last_engine"] == 2 and st["last_overflow"] == 0 and st["mma_passes"] >= 16
    assert st["kernel_launches"] - l0 < 16, "a 4096-query batch must be one pass through the launch sequence"
    for qi in list(range(0, nq, 257)) + [nq - 1]:
        wi, ws = O.search_nearest_neighbors(qs[qi], big, k, mode="canonical")
        assert mi[qi].tolist() == wi.tolist() and bits_equal(ms[qi], ws), qi
    sub_i, sub_s = fm.searchBatch(qs[1000:1300], qm, k)
    assert np.array_equal(sub_i, mi[1000:1300]) and bits_equal(sub_s, ms[1000:1300])


# ---- the class members beside the search path ---------------------------------------------------------------------
@pytest.mark.parametrize("sim", SIMS)
@pytest.mark.parametrize("qb", [1, 4])
@pytest.mark.parametrize("n,dim", [(60, 128), (257, 100)])
def test_compute_quantization_accuracy_bit_exact(bbq, sim, qb, n, dim):
    """computeQuantizationAccuracy (src/binaryQuantizationFormat.ts:420-475, src/binaryQuantizedScorer.ts:524-617) on
    the device == its oracle restatement, all five statistics as f64 bit patterns."""
    rows, qs = gaussian(n, dim, 701 + n), gaussian(n, dim, 702 + n)
    fmt = make_format(bbq, sim, qb=qb)
    got = fmt.computeQuantizationAccuracy(rows, qs)
    want = O.compute_quantization_accuracy(rows, qs, sim, qb)
    for key in ("meanError", "maxError", "minError", "stdError", "correlation"):
        assert np.float64(got[key]).view(np.uint64) == np.float64(want[key]).view(np.uint64), (key, got[key], want[key])
    got7 = fmt.computeQuantizationAccuracy(rows, qs, targetOrd=7)
    want7 = O.compute_quantization_accuracy(rows, qs, sim, qb, target=7)
    assert got7 == want7 and got7 != got


def test_compute_quantization_accuracy_errors(bbq):
    rows = gaussian(10, 32, 1)
    fmt = make_format(bbq, "COSINE")
    with pytest.raises(bbq.BbqError, match="原始向量集合不能为空"):
        fmt.computeQuantizationAccuracy([], rows)
    with pytest.raises(bbq.BbqError, match="查询向量集合不能为空"):
        fmt.computeQuantizationAccuracy(rows, [])
    with pytest.raises(bbq.BbqError, match="长度不匹配"):
        fmt.computeQuantizationAccuracy(rows, rows[:5])
    with pytest.raises(bbq.BbqError, match="不支持的查询位数: 8"):
        make_format(bbq, "COSINE", qb=8).computeQuantizationAccuracy(rows, rows)
    bad = rows.copy()
    bad[3, 2] = np.nan
    with pytest.raises(bbq.BbqError, match="包含NaN值"):
        fmt.computeQuantizationAccuracy(rows, bad)


@pytest.mark.parametrize("sim", SIMS)
@pytest.mark.parametrize("qb", [1, 4, 8])
def test_quantize_query_vector_with_centroid_is_the_class_member(bbq, sim, qb):
    """format.quantizeQueryVector(query, centroid) — ONE normalisation for COSINE (the search path does two)."""
    dim = 200
    q, cen = gaussian(1, dim, 801)[0], (gaussian(1, dim, 802)[0] * 0.05).astype(np.float32)
    fmt = make_format(bbq, sim, qb=qb)
    got = fmt.quantizeQueryVector(q, cen)
    codes, corr = O.quantize_query_vector_once(q, cen, sim, qb)
    assert np.array_equal(got["quantizedQuery"], codes)
    gc = got["queryCorrections"]
    have = np.array([gc["lowerInterval"], gc["upperInterval"], gc["additionalCorrection"], gc["quantizedComponentSum"]])
    assert bits_equal(have, corr)
    if sim == "COSINE":   # and it differs from the search path's double normalisation in general
        c2, corr2 = O.quantize_query_vector(q, cen, sim, qb)
        assert not bits_equal(corr, corr2) or np.array_equal(codes, c2)


def test_serialize_roundtrip(bbq):
    rows, qs = gaussian(300, 96, 901), gaussian(3, 96, 902)
    fmt = make_format(bbq, "COSINE")
    ser = fmt.serializeVectorData(rows)
    assert ser["metadata"]["vectorCount"] == 300 and ser["metadata"]["dimensions"] == 96
    assert len(ser["vectorData"]) == 300 and ser["vectorData"][0]["binaryValues"].size == 12
    back = fmt.deserializeVectorData(ser["vectorData"], ser["metadata"])
    direct = fmt.quantizeVectors(rows)["quantizedVectors"]
    a, b = fmt.searchBatch(qs, back, 10), fmt.searchBatch(qs, direct, 10)
    assert np.array_equal(a[0], b[0]) and bits_equal(a[1], b[1])


# ---- EXTENSION: indexBits = 2 (BASELINE configs[4]; the reference throws — semantics defined in oracle/bbq_oracle.cpp) --
def make_format_ib(bbq, sim, qb, ib, scan=None):
    import os
    saved = os.environ.pop("BBQ_SCAN", None)
    if scan is not None:
        os.environ["BBQ_SCAN"] = scan
    try:
        return bbq.createBinaryQuantizationFormat(
            {"queryBits": qb, "indexBits": ib, "quantizer": {"similarityFunction": sim, "lambda": 0.1, "iters": 5}})
    finally:
        os.environ.pop("BBQ_SCAN", None)
        if saved is not None:
            os.environ["BBQ_SCAN"] = saved


@pytest.mark.parametrize("sim", SIMS)
@pytest.mark.parametrize("n,dim", [(700, 128), (300, 100), (257, 1536)])
def test_two_bit_index_build_bit_exact(bbq, sim, n, dim):
    """K5 with indexBits = 2: codes (0..3), intervals, additional correction and component sums == the oracle's
    quantizeVectors (the reference's own quantiser, which accepts indexBits = 2), through the plane-interleaved rows."""
    rows = gaussian(n, dim, 1001 + dim)
    want = O.quantize_vectors(rows, sim=sim, index_bits=2, want_unpacked=False)
    fmt = make_format_ib(bbq, sim, 8, 2)
    qv = fmt.quantizeVectors(rows)["quantizedVectors"]
    codes, corr = qv.exportAll()
    assert codes.shape == (n, dim) and codes.max() <= 3
    assert np.array_equal(codes, want.packed) and bits_equal(corr, want.corr)
    assert bits_equal(qv.getCentroid(), want.centroid)
    back = fmt.adoptQuantized(codes, corr, want.centroid)          # and the adopt path re-interleaves the planes
    c2, r2 = back.exportAll()
    assert np.array_equal(c2, codes) and bits_equal(r2, corr)


@pytest.mark.parametrize("sim", SIMS)
@pytest.mark.parametrize("qb", [4, 8])
@pytest.mark.parametrize("n,dim", [(3000, 128), (1100, 100), (900, 1536)])
def test_two_bit_index_dots_and_scores_bit_exact(bbq, sim, qb, n, dim):
    rows, qs = gaussian(n, dim, 1101 + dim), gaussian(2, dim, 1102 + dim)
    idx = O.quantize_vectors(rows, sim=sim, index_bits=2, want_unpacked=False)
    fmt = make_format_ib(bbq, sim, qb, 2)
    qv = fmt.adoptQuantized(idx.packed, idx.corr, idx.centroid)
    for q in qs:
        _, _, alls, alld = O.search_nearest_neighbors(q, idx, 5, query_bits=qb, want_all=True)
        assert np.array_equal(fmt.debugQcDist(q, qv), alld)           # popcount engine over 9 (5) virtual planes
        assert bits_equal(fmt.debugScores(q, qv), alls)
    got = fmt.debugQcDistBatch(gaussian(21, dim, 1103), qv)           # tensor-core engine: nibble columns for 8-bit codes
    for qi, q in enumerate(gaussian(21, dim, 1103)):
        _, _, _, alld = O.search_nearest_neighbors(q, idx, 5, query_bits=qb, want_all=True)
        assert np.array_equal(got[qi], alld), qi


@pytest.mark.parametrize("sim", SIMS)
@pytest.mark.parametrize("n,dim,nq,k,qb", [(30000, 256, 70, 10, 8), (20000, 1536, 40, 100, 8), (25000, 384, 33, 10, 4)])
def test_two_bit_index_search_both_engines(bbq, sim, n, dim, nq, k, qb):
    rows, qs = gaussian(n, dim, 1201 + dim), gaussian(nq, dim, 1202 + dim)
    idx = O.quantize_vectors(rows, sim=sim, index_bits=2, want_unpacked=False, centroid=np.zeros(dim, np.float32))
    fm, fp = make_format_ib(bbq, sim, qb, 2, scan="mma"), make_format_ib(bbq, sim, qb, 2, scan="popc")
    qm = fm.adoptQuantized(idx.packed, idx.corr, idx.centroid)
    qp = fp.adoptQuantized(idx.packed, idx.corr, idx.centroid)
    mi, ms = fm.searchBatch(qs, qm, k)
    assert fm.stats()["last_engine"] == 2 and fm.stats()["last_overflow"] == 0
    pi, ps = fp.searchBatch(qs[:6], qp, k)
    assert fp.stats()["last_engine"] == 1
    assert np.array_equal(mi[:6], pi) and bits_equal(ms[:6], ps)
    for qi in range(0, nq, max(1, nq // 5)):
        wi, ws = O.search_nearest_neighbors(qs[qi], idx, k, query_bits=qb, mode="canonical")
        assert mi[qi].tolist() == wi.tolist() and bits_equal(ms[qi], ws), qi


@pytest.mark.parametrize("qb", [6, 8])
def test_wide_queries_on_the_tensor_core_scan(bbq, qb):
    """6..8-bit queries on a 1-bit index (reference behaviour: the 4-bit path with its 1/15 scale) now also run on
    tcgen05: the code is split into two nibble columns."""
    n, dim, nq, k = 30000, 256, 50, 10
    rows, qs = gaussian(n, dim, 1301), gaussian(nq, dim, 1302)
    idx = O.quantize_vectors(rows, sim="COSINE", want_unpacked=False)
    fm = make_format(bbq, "COSINE", qb=qb, scan="mma")
    qm = fm.adoptQuantized(idx.packed, idx.corr, idx.centroid)
    mi, ms = fm.searchBatch(qs, qm, k)
    assert fm.stats()["last_engine"] == 2
    for qi in range(0, nq, 7):
        wi, ws = O.search_nearest_neighbors(qs[qi], idx, k, query_bits=qb, mode="canonical")
        assert mi[qi].tolist() == wi.tolist() and bits_equal(ms[qi], ws)
    d = fm.debugQcDistBatch(qs[:5], qm)
    _, _, _, alld = O.search_nearest_neighbors(qs[3], idx, 1, query_bits=qb, want_all=True)
    assert np.array_equal(d[3], alld)


def test_pinned_host_buffers_and_limits(bbq):
    """bbq_host_alloc / bbq_host_free (page-locked query / result buffers for a host) and the documented limits."""
    import ctypes as C
    L = bbq._native.load()
    rows, qs = gaussian(5000, 64, 1401), gaussian(6, 64, 1402)
    fmt = make_format(bbq, "COSINE")
    qv = fmt.quantizeVectors(rows)["quantizedVectors"]
    want_i, want_s = fmt.searchBatch(qs, qv, 10)
    nbytes_q, nbytes_o = qs.nbytes, 6 * 10 * 4
    hq, hi, hs = L.bbq_host_alloc(nbytes_q), L.bbq_host_alloc(nbytes_o), L.bbq_host_alloc(nbytes_o)
    assert hq and hi and hs
    try:
        C.memmove(hq, qs.ctypes.data, nbytes_q)
        cnt = C.c_uint32(0)
        assert L.bbq_search(qv._h, hq, 6, 10, hi, hs, C.byref(cnt)) == 0 and cnt.value == 10
        got_i = np.frombuffer((C.c_int32 * 60).from_address(hi), np.int32).reshape(6, 10)
        got_s = np.frombuffer((C.c_float * 60).from_address(hs), np.float32).reshape(6, 10)
        assert np.array_equal(got_i, want_i) and bits_equal(got_s.copy(), want_s)
    finally:
        for p in (hq, hi, hs):
            L.bbq_host_free(p)
    with pytest.raises(bbq.BbqError) as e:       # the device top-k serves k <= 4096
        fmt.searchBatch(qs, qv, 4097)
    assert e.value.status == 9
    with pytest.raises(bbq.BbqError) as e:       # indexBits > 2: neither the reference's search nor this build's
        bbq.createBinaryQuantizationFormat({"indexBits": 3, "quantizer": {"similarityFunction": "COSINE"}}).quantizeVectors(rows)
    assert e.value.status == 9


# ---- narrow batches: the captured launch sequence (CUDA graph) of bbq_search ------------------------------------------
@pytest.mark.parametrize("sim", SIMS)
@pytest.mark.parametrize("n,nq", [(30000, 1), (30000, 3), (30000, 8), (900, 2)])   # popcount stream scan, tensor-core scan, direct path
def test_replayed_launch_sequence_equals_oracle(bbq, sim, n, nq):
    """bbq_search captures the device sequence of a narrow batch the second time the same (index, batch size, k) arrives
    and replays it afterwards: every call — direct, capturing, replayed — must return the oracle's lists for ITS queries,
    report bad queries like the reference, and survive other calls in between (which move scratch buffers)."""
    dim, k = 96, 7
    rows = gaussian(n, dim, 700 + nq)
    idx = O.quantize_vectors(rows, sim=sim, want_unpacked=False)
    fmt = make_format(bbq, sim)
    qv = fmt.adoptQuantized(idx.packed, idx.corr, idx.centroid)

    def check(seed):
        qs = gaussian(nq, dim, seed)
        gi, gs = fmt.searchBatch(qs, qv, k)
        for qi in range(nq):
            wi, ws = O.search_nearest_neighbors(qs[qi], idx, k, mode="canonical")
            assert gi[qi].tolist() == wi.tolist() and bits_equal(gs[qi], ws), (seed, qi)

    for seed in range(800, 806):
        check(seed)
    r0 = fmt.stats()["graph_replays"]
    assert r0 >= 3, "the third and later calls with one key replay the captured sequence"
    launches_before = fmt.stats()["kernel_launches"]
    check(806)
    assert fmt.stats()["kernel_launches"] > launches_before        # the replayed kernels are still counted
    # a bad query through the replayed sequence: the reference's message, and the context answers as before afterwards
    bad = gaussian(nq, dim, 807)
    bad[nq - 1, 5] = np.nan
    with pytest.raises(bbq.BbqError) as e:
        fmt.searchBatch(bad, qv, k)
    assert e.value.status == 5
    check(808)
    # another batch size in between moves scratch buffers: the old sequence must not be replayed on stale addresses
    big = gaussian(300, dim, 809)
    gi, gs = fmt.searchBatch(big, qv, k)
    wi, ws = O.search_nearest_neighbors(big[299], idx, k, mode="canonical")
    assert gi[299].tolist() == wi.tolist() and bits_equal(gs[299], ws)
    for seed in range(810, 814):
        check(seed)
    assert fmt.stats()["graph_replays"] > r0
    # a second index of the same shape on the same context: its own key, its own answers
    rows2 = gaussian(n, dim, 900 + nq)
    idx2 = O.quantize_vectors(rows2, sim=sim, want_unpacked=False)
    qv2 = fmt.adoptQuantized(idx2.packed, idx2.corr, idx2.centroid)
    for seed in range(820, 824):
        qs = gaussian(nq, dim, seed)
        gi, gs = fmt.searchBatch(qs, qv2, k)
        wi, ws = O.search_nearest_neighbors(qs[0], idx2, k, mode="canonical")
        assert gi[0].tolist() == wi.tolist() and bits_equal(gs[0], ws)
        check(seed + 10)                                           # and the first index, alternating


def test_replayed_sequence_overflow_falls_back(bbq):
    """All rows equal: the candidate lists overflow on every call; with the overflow flag read after the call's single
    synchronisation (and the sequence replayed from a graph) the exact path must still take over."""
    n, dim, nq, k = 40000, 64, 2, 10
    row = gaussian(1, dim, 131)
    idx = O.quantize_vectors(np.concatenate([row, gaussian(3, dim, 132)]), sim="COSINE", want_unpacked=False,
                             centroid=np.zeros(dim, np.float32))
    packed, corr = np.repeat(idx.packed[:1], n, 0), np.repeat(idx.corr[:1], n, 0)
    big = O.OracleIndex(idx.centroid, packed, None, corr, dim, "COSINE", 1)
    fmt = make_format(bbq, "COSINE")
    qv = fmt.adoptQuantized(packed, corr, idx.centroid)
    for rep in range(5):
        qs = gaussian(nq, dim, 133 + rep)
        gi, gs = fmt.searchBatch(qs, qv, k)
        st = fmt.stats()
        assert st["last_overflow"] == 1 and st["last_path"] == 2
        wi, ws = O.search_nearest_neighbors(qs[1], big, k, mode="canonical")
        assert gi[1].tolist() == wi.tolist() == list(range(k)) and bits_equal(gs[1], ws)
