"""The exact-arithmetic header the CUDA kernels compile (csrc/bbq_numerics.cuh) is also compiled for the
host here and compared with the oracle BIT-FOR-BIT: quantiser intervals/codes, scores, top-k keys.
No GPU.  (The GPU suite then checks that the device executes the same header identically.)"""
import ctypes as C

import numpy as np
import pytest

from oracle import oracle as O
from tests.fixtures import gaussian, sincos_dataset
from tests.native.build_native import build_host_numerics

f32p, u8p, f64p, i32p = C.POINTER(C.c_float), C.POINTER(C.c_uint8), C.POINTER(C.c_double), C.POINTER(C.c_int32)


@pytest.fixture(scope="module")
def hn():
    L = C.CDLL(build_host_numerics())
    L.hn_osq.argtypes = [f32p, f32p, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, u8p, f64p]
    L.hn_scores.argtypes = [i32p, f64p, C.c_int64, f64p, C.c_int, C.c_double, C.c_int, C.c_int, f32p]
    L.hn_topk_key.argtypes = [C.c_float, C.c_uint32]
    L.hn_topk_key.restype = C.c_uint64
    L.hn_key_score.argtypes = [C.c_uint64]
    L.hn_key_score.restype = C.c_float
    L.hn_key_id.argtypes = [C.c_uint64]
    L.hn_key_id.restype = C.c_uint32
    return L


def _osq(L, v, c, bits, sim, lam, iters):
    v = np.ascontiguousarray(v, np.float32)
    c = np.ascontiguousarray(c, np.float32)
    codes = np.empty(v.size, np.uint8)
    corr = np.empty(4, np.float64)
    L.hn_osq(v.ctypes.data_as(f32p), c.ctypes.data_as(f32p), v.size, bits, O.SIM[sim], lam, iters,
             codes.ctypes.data_as(u8p), corr.ctypes.data_as(f64p))
    return codes, corr


def _same(a, b):
    return np.array_equal(np.asarray(a).view(np.uint64), np.asarray(b).view(np.uint64))


@pytest.mark.parametrize("sim", ["EUCLIDEAN", "COSINE", "MAXIMUM_INNER_PRODUCT"])
@pytest.mark.parametrize("bits", [1, 2, 4, 8])
@pytest.mark.parametrize("dim", [8, 100, 768])
def test_osq_bit_exact(hn, sim, bits, dim):
    rows = gaussian(24, dim, 100 + dim)
    cen = O.compute_centroid(rows)
    for lam, iters in [(0.1, 5), (0.001, 20)]:
        for v in rows:
            c0, r0 = O.scalar_quantize(v, cen, bits, sim, lam, iters)
            c1, r1 = _osq(hn, v, cen, bits, sim, lam, iters)
            assert np.array_equal(c0, c1) and _same(r0, r1)


def test_osq_degenerate_bit_exact(hn):
    cases = [(np.zeros(16, np.float32), np.zeros(16, np.float32)),                # constant: a == b, NaN interval
             (np.full(16, 2.5, np.float32), np.full(16, 2.5, np.float32)),
             (np.array([1, -1, .5, -.5], np.float32), np.zeros(4, np.float32)),
             (np.array([1e30, -1e30, 3, 4], np.float32), np.zeros(4, np.float32)),
             (np.array([1e-30, 2e-30, -1e-30, 0], np.float32), np.zeros(4, np.float32)),
             # signed zeros: Math.min(+0, -0) = -0 and Math.max(-0, +0) = +0 decide the sign of a zero interval
             (np.array([0, 0, 0, -0.0, 0, 0, 0, -0.0], np.float32), np.zeros(8, np.float32)),
             (np.array([-0.0, 0, -0.0, 0], np.float32), np.zeros(4, np.float32)),
             (np.array([-0.0, -0.0, -0.0, -0.0], np.float32), np.zeros(4, np.float32)),
             (np.array([0, -0.0, 1, -1], np.float32), np.array([0, 0, 1, -1], np.float32))]
    for v, c in cases:
        for bits in (1, 4):
            for sim in ("EUCLIDEAN", "COSINE"):
                c0, r0 = O.scalar_quantize(v, c, bits, sim)
                c1, r1 = _osq(hn, v, c, bits, sim, 0.1, 5)
                assert np.array_equal(c0, c1) and _same(r0, r1)


def test_osq_sincos_fixture_bit_exact(hn):
    base, _ = sincos_dataset(128, 40, 1)
    cen = O.compute_centroid(base)
    for v in base:
        for bits in (1, 4):
            c0, r0 = O.scalar_quantize(v, cen, bits, "COSINE", 0.001, 20)
            c1, r1 = _osq(hn, v, cen, bits, "COSINE", 0.001, 20)
            assert np.array_equal(c0, c1) and _same(r0, r1)


@pytest.mark.parametrize("sim", ["EUCLIDEAN", "COSINE", "MAXIMUM_INNER_PRODUCT"])
@pytest.mark.parametrize("qb", [1, 4, 8])
def test_scores_bit_exact(hn, sim, qb):
    rows, qs = gaussian(400, 256, 5), gaussian(4, 256, 6)
    idx = O.quantize_vectors(rows, sim=sim)
    cdp = O.centroid_dp(idx.centroid)
    for q in qs:
        qc, qcorr = O.quantize_query_vector(q, idx.centroid, sim=sim, query_bits=qb)
        dots = (idx.unpacked.astype(np.int32) @ qc.astype(np.int32)).astype(np.int32)
        want = O.batch_scores(dots, idx.corr, qcorr, 256, cdp, sim, qb)
        got = np.empty(len(dots), np.float32)
        hn.hn_scores(dots.ctypes.data_as(i32p), idx.corr.ctypes.data_as(f64p), len(dots), qcorr.ctypes.data_as(f64p),
                     256, cdp, O.SIM[sim], qb, got.ctypes.data_as(f32p))
        assert np.array_equal(want.view(np.uint32), got.view(np.uint32))


def test_topk_key_total_order(hn):
    vals = np.array([np.inf, 3.5, 1.0, 1e-45, 0.0, -0.0, -1e-45, -2.0, -np.inf, np.nan], np.float32)
    keys = [hn.hn_topk_key(float(v), 7) for v in vals]
    # descending scores -> strictly descending keys, except +0 == -0; NaN last
    assert keys[4] == keys[5]
    ks = keys[:5] + keys[6:]
    assert all(a > b for a, b in zip(ks, ks[1:]))
    # ties: lower id wins
    assert hn.hn_topk_key(1.0, 3) > hn.hn_topk_key(1.0, 4)
    for v in vals[:9]:
        k = hn.hn_topk_key(float(v), 123)
        assert hn.hn_key_id(k) == 123 and hn.hn_key_score(k) == np.float32(v) + np.float32(0)
    assert np.isnan(hn.hn_key_score(hn.hn_topk_key(float("nan"), 5)))
    assert hn.hn_topk_key(float("nan"), 0xFFFFFFFE) > 0   # 0 is reserved for "empty slot"


# ---- round 2: the extension's score formula, the single-vector scorer, the accuracy statistics ------------------
@pytest.mark.parametrize("sim", ["EUCLIDEAN", "COSINE", "MAXIMUM_INNER_PRODUCT"])
@pytest.mark.parametrize("qb,ib", [(8, 2), (4, 2), (1, 2), (4, 1), (8, 1), (1, 1)])
def test_scores_any_bits_bit_exact(hn, sim, qb, ib):
    """score_f32 as the scan kernels call it (score mode, lx divisor, query terms) == the oracle: the reference's formulas
    for a 1-bit index, the extension's (oracle/bbq_oracle.cpp:score_ext) for a 2-bit one."""
    hn.hn_scores_bits.argtypes = [i32p, f64p, C.c_int64, f64p, C.c_int, C.c_double, C.c_int, C.c_int, C.c_int, f32p]
    rows, qs = gaussian(300, 192, 15 + ib), gaussian(3, 192, 16 + qb)
    idx = O.quantize_vectors(rows, sim=sim, index_bits=ib)
    cdp = O.centroid_dp(idx.centroid)
    for q in qs:
        _, _, alls, alld = O.search_nearest_neighbors(q, idx, 3, query_bits=qb, want_all=True)
        _, qcorr = O.quantize_query_vector(q, idx.centroid, sim=sim, query_bits=qb)
        got = np.empty(len(alld), np.float32)
        hn.hn_scores_bits(alld.ctypes.data_as(i32p), idx.corr.ctypes.data_as(f64p), len(alld), qcorr.ctypes.data_as(f64p),
                          192, cdp, O.SIM[sim], qb, ib, got.ctypes.data_as(f32p))
        assert np.array_equal(alls.view(np.uint32), got.view(np.uint32))


@pytest.mark.parametrize("sim", ["EUCLIDEAN", "COSINE", "MAXIMUM_INNER_PRODUCT"])
@pytest.mark.parametrize("qb", [1, 4])
def test_single_vector_score_and_accuracy_stats_bit_exact(hn, sim, qb):
    """The single-vector scorer (computeQuantizedScore) and the statistics of computeQuantizationAccuracy, header vs
    oracle, f64 bit patterns."""
    hn.hn_score_single.argtypes = [C.c_double, f64p, f64p, C.c_int, C.c_double, C.c_int, C.c_int]
    hn.hn_score_single.restype = C.c_double
    hn.hn_accuracy_stats.argtypes = [f64p, f64p, C.c_int64, f64p]
    L = O.lib()
    rows, qs = gaussian(40, 96, 21), gaussian(40, 96, 22)
    idx = O.quantize_vectors(rows, sim=sim)
    cdp = O.centroid_dp(idx.centroid) if qb == 1 else 0.0
    xc = np.ascontiguousarray(idx.corr[0])
    for q in qs[:10]:
        codes, qc = O.quantize_query_vector_once(q, idx.centroid, sim, qb)
        dot = float(O.dot_unpacked(codes, idx.unpacked[0]))
        want = L.bbqo_score_single(dot, xc.ctypes.data_as(f64p), qc.ctypes.data_as(f64p), 96, cdp, O.SIM[sim], qb)
        got = hn.hn_score_single(dot, xc.ctypes.data_as(f64p), qc.ctypes.data_as(f64p), 96, cdp, O.SIM[sim], qb)
        assert np.float64(want).view(np.uint64) == np.float64(got).view(np.uint64)
    stats, orig, quant = O.compute_quantization_accuracy(rows, qs, sim, qb, want_scores=True)
    out = np.empty(5, np.float64)
    hn.hn_accuracy_stats(orig.ctypes.data_as(f64p), quant.ctypes.data_as(f64p), len(orig), out.ctypes.data_as(f64p))
    want5 = np.array([stats[k] for k in ("meanError", "maxError", "minError", "stdError", "correlation")])
    assert _same(out, want5)
