"""ctypes wrapper around oracle/libbbq_oracle.so — the CPU ORACLE (TEST INFRASTRUCTURE ONLY).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` legs may
import this module.  Nothing under better-binary-quantization_b200/ does.

PARITY STATUS: pinned bit for bit by fixtures generated from the reference's own source text
(tests/golden/from_ts/*.ts.json, tests/test_golden_from_ts.py) — see the header of bbq_oracle.cpp.

Function names follow the reference (leolee9086/Better-Binary-Quantization, TypeScript):
  normalize_vector      src/vectorOperations.ts:11-34
  compute_centroid      src/vectorOperations.ts:126-163
  scalar_quantize       src/optimizedScalarQuantizer.ts:108-227
  pack_as_binary        src/optimizedScalarQuantizer.ts:420-446
  quantize_vectors      src/binaryQuantizationFormat.ts:165-263
  quantize_query_vector src/binaryQuantizationFormat.ts:337-347,271-299
  qcdist_packed         src/utils/computeBatchFourBitDotProductDirectPacked.ts:10-53
  batch_scores          src/batchDotProduct.ts:478-541,554-617
  search_nearest_neighbors src/binaryQuantizationFormat.ts:308-412
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libbbq_oracle.so")
_LIB_OVERRIDE = os.environ.get("BBQ_ORACLE_LIB")  # tools/sanitize.sh host: an ASan/UBSan build of the same source

SIM = {"EUCLIDEAN": 0, "COSINE": 1, "MAXIMUM_INNER_PRODUCT": 2}


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "bbq_oracle.cpp")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "-B", "libbbq_oracle.so"])
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not _LIB_OVERRIDE:
            build()
        L = C.CDLL(os.path.abspath(_LIB_OVERRIDE) if _LIB_OVERRIDE else _LIB_PATH)
        f32p, u8p, f64p, i32p = (C.POINTER(C.c_float), C.POINTER(C.c_uint8), C.POINTER(C.c_double),
                                 C.POINTER(C.c_int32))
        L.bbqo_normalize.argtypes = [f32p, C.c_int, f32p]
        L.bbqo_centroid.argtypes = [f32p, C.c_int64, C.c_int, f32p]
        L.bbqo_osq.argtypes = [f32p, f32p, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, u8p, f64p]
        L.bbqo_pack_binary.argtypes = [u8p, C.c_int, u8p]
        L.bbqo_build_index.argtypes = [f32p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int,
                                       f32p, f32p, u8p, u8p, f64p]
        L.bbqo_quantize_query.argtypes = [f32p, f32p, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, u8p, f64p]
        L.bbqo_qcdist_packed.argtypes = [u8p, u8p, C.c_int64, C.c_int, i32p]
        L.bbqo_qcdist_packed_planes.argtypes = [u8p, C.c_int, u8p, C.c_int64, C.c_int, i32p]
        L.bbqo_qcdist_1bit.argtypes = [u8p, u8p, C.c_int64, C.c_int, i32p]
        L.bbqo_dot_unpacked.argtypes = [u8p, u8p, C.c_int]
        L.bbqo_dot_unpacked.restype = C.c_int32
        L.bbqo_centroid_dp.argtypes = [f32p, C.c_int]
        L.bbqo_centroid_dp.restype = C.c_double
        L.bbqo_scores.argtypes = [i32p, f64p, C.c_int64, f64p, C.c_int, C.c_double, C.c_int, C.c_int, f32p]
        L.bbqo_scale_mip.argtypes = [C.c_double]
        L.bbqo_scale_mip.restype = C.c_double
        L.bbqo_topk_heap.argtypes = [f32p, C.c_int64, C.c_int64, i32p, f32p]
        L.bbqo_topk_heap.restype = C.c_int64
        L.bbqo_topk_canonical.argtypes = [f32p, C.c_int64, C.c_int64, i32p, f32p]
        L.bbqo_topk_canonical.restype = C.c_int64
        L.bbqo_cosine.argtypes = [f32p, f32p, C.c_int]
        L.bbqo_cosine.restype = C.c_double
        L.bbqo_rerank_heap.argtypes = [f64p, C.c_int64, C.c_int64, i32p]
        L.bbqo_rerank_heap.restype = C.c_int64
        L.bbqo_search.argtypes = [f32p, f32p, u8p, f64p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_double,
                                  C.c_int, C.c_int64, C.c_int, i32p, f32p, f32p, i32p]
        L.bbqo_search.restype = C.c_int64
        L.bbqo_search_ext.argtypes = [f32p, f32p, u8p, f64p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double,
                                      C.c_int, C.c_int64, C.c_int, i32p, f32p, f32p, i32p]
        L.bbqo_search_ext.restype = C.c_int64
        L.bbqo_score_ext.argtypes = [C.c_double, f64p, f64p, C.c_int, C.c_double, C.c_int, C.c_int, C.c_int]
        L.bbqo_score_ext.restype = C.c_double
        L.bbqo_quantize_query_once.argtypes = [f32p, f32p, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, u8p, f64p]
        L.bbqo_score_single.argtypes = [C.c_double, f64p, f64p, C.c_int, C.c_double, C.c_int, C.c_int]
        L.bbqo_score_single.restype = C.c_double
        L.bbqo_similarity.argtypes = [f32p, f32p, C.c_int, C.c_int]
        L.bbqo_similarity.restype = C.c_double
        L.bbqo_accuracy_stats.argtypes = [f64p, f64p, C.c_int64, f64p]
        L.bbqo_quantization_accuracy.argtypes = [f32p, f32p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int,
                                                 C.c_int64, f64p, f64p, f64p]
        L.bbqo_quantization_accuracy.restype = C.c_int
        _lib = L
    return _lib


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t)) if a is not None else None


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def normalize_vector(v):
    v = _f32(v)
    out = np.empty_like(v)
    lib().bbqo_normalize(_p(v, C.c_float), v.size, _p(out, C.c_float))
    return out


def compute_centroid(rows):
    rows = _f32(rows)
    n, d = rows.shape
    c = np.empty(d, np.float32)
    lib().bbqo_centroid(_p(rows, C.c_float), n, d, _p(c, C.c_float))
    return c


def scalar_quantize(v, centroid, bits, sim="EUCLIDEAN", lam=0.1, iters=5):
    """-> (codes u8[d], corr f64[4] = lower, upper, additional, componentSum)"""
    v, centroid = _f32(v), _f32(centroid)
    d = v.size
    codes = np.empty(d, np.uint8)
    corr = np.empty(4, np.float64)
    lib().bbqo_osq(_p(v, C.c_float), _p(centroid, C.c_float), d, bits, SIM[sim], lam, iters,
                   _p(codes, C.c_uint8), _p(corr, C.c_double))
    return codes, corr


def pack_as_binary(codes):
    codes = np.ascontiguousarray(codes, np.uint8)
    out = np.empty((codes.size + 7) // 8, np.uint8)
    lib().bbqo_pack_binary(_p(codes, C.c_uint8), codes.size, _p(out, C.c_uint8))
    return out


@dataclass
class OracleIndex:
    """In-memory index as BinarizedByteVectorValuesImpl holds it (binaryQuantizationFormat.ts:24-126)."""
    centroid: np.ndarray      # f32[d]
    packed: np.ndarray        # u8[n, ceil(d/8)] (index_bits==1) else u8[n, d]
    unpacked: np.ndarray      # u8[n, d]
    corr: np.ndarray          # f64[n, 4]
    dim: int
    sim: str
    index_bits: int

    def size(self):
        return self.packed.shape[0]


def quantize_vectors(rows, sim="COSINE", index_bits=1, lam=0.1, iters=5, centroid=None,
                     want_unpacked=True) -> OracleIndex:
    rows = _f32(rows)
    n, d = rows.shape
    p = (d + 7) // 8 if index_bits == 1 else d
    c_out = np.empty(d, np.float32)
    packed = np.empty((n, p), np.uint8)
    unpacked = np.empty((n, d), np.uint8) if want_unpacked else None
    corr = np.empty((n, 4), np.float64)
    c_in = _f32(centroid) if centroid is not None else None
    lib().bbqo_build_index(_p(rows, C.c_float), n, d, SIM[sim], index_bits, lam, iters, _p(c_in, C.c_float),
                           _p(c_out, C.c_float), _p(packed, C.c_uint8), _p(unpacked, C.c_uint8),
                           _p(corr, C.c_double))
    return OracleIndex(c_out, packed, unpacked, corr, d, sim, index_bits)


def quantize_query_vector(q, centroid, sim="COSINE", query_bits=4, lam=0.1, iters=5):
    q, centroid = _f32(q), _f32(centroid)
    d = q.size
    codes = np.empty(d, np.uint8)
    corr = np.empty(4, np.float64)
    lib().bbqo_quantize_query(_p(q, C.c_float), _p(centroid, C.c_float), d, SIM[sim], query_bits, lam, iters,
                              _p(codes, C.c_uint8), _p(corr, C.c_double))
    return codes, corr


def quantize_query_vector_once(q, centroid, sim="COSINE", query_bits=4, lam=0.1, iters=5):
    """format.quantizeQueryVector called directly (src/binaryQuantizationFormat.ts:271-299): ONE normalisation."""
    q, centroid = _f32(q), _f32(centroid)
    d = q.size
    codes = np.empty(d, np.uint8)
    corr = np.empty(4, np.float64)
    lib().bbqo_quantize_query_once(_p(q, C.c_float), _p(centroid, C.c_float), d, SIM[sim], query_bits, lam, iters,
                                   _p(codes, C.c_uint8), _p(corr, C.c_double))
    return codes, corr


def accuracy_stats(orig, quant):
    """BinaryQuantizedScorer.computeQuantizationAccuracy, src/binaryQuantizedScorer.ts:524-617"""
    orig = np.ascontiguousarray(orig, np.float64)
    quant = np.ascontiguousarray(quant, np.float64)
    out = np.empty(5, np.float64)
    lib().bbqo_accuracy_stats(_p(orig, C.c_double), _p(quant, C.c_double), orig.size, _p(out, C.c_double))
    return dict(zip(("meanError", "maxError", "minError", "stdError", "correlation"), out.tolist()))


def compute_quantization_accuracy(rows, queries, sim="COSINE", query_bits=4, lam=0.1, iters=5, target=0,
                                  want_scores=False):
    """BinaryQuantizationFormat.computeQuantizationAccuracy, src/binaryQuantizationFormat.ts:420-475."""
    rows, queries = _f32(rows), _f32(queries)
    n, d = rows.shape
    assert queries.shape == rows.shape
    out = np.empty(5, np.float64)
    orig, quant = np.empty(n, np.float64), np.empty(n, np.float64)
    rc = lib().bbqo_quantization_accuracy(_p(rows, C.c_float), _p(queries, C.c_float), n, d, SIM[sim], query_bits,
                                          lam, iters, target, _p(out, C.c_double), _p(orig, C.c_double),
                                          _p(quant, C.c_double))
    if rc != 0:
        raise ValueError(f"不支持的查询位数: {query_bits}，只支持1位和4位")
    stats = dict(zip(("meanError", "maxError", "minError", "stdError", "correlation"), out.tolist()))
    return (stats, orig, quant) if want_scores else stats


def qcdist_packed(qcodes, packed, d, planes=None):
    qcodes = np.ascontiguousarray(qcodes, np.uint8)
    packed = np.ascontiguousarray(packed, np.uint8)
    n = packed.shape[0]
    out = np.empty(n, np.int32)
    if planes is None:
        lib().bbqo_qcdist_packed(_p(qcodes, C.c_uint8), _p(packed, C.c_uint8), n, d, _p(out, C.c_int32))
    else:
        lib().bbqo_qcdist_packed_planes(_p(qcodes, C.c_uint8), planes, _p(packed, C.c_uint8), n, d,
                                        _p(out, C.c_int32))
    return out


def qcdist_matrix(qcodes, packed, d, chunk=65536):
    """qcDist for MANY queries at once, in numpy: out[q][v] = sum_d qcodes[q][d] * bit_d(x_v) — the same integer as
    computeBatchFourBitDotProductDirectPacked (src/utils/computeBatchFourBitDotProductDirectPacked.ts:10-53: bit d of
    row v is (byte[d>>3] >> (7-(d&7))) & 1).  A float32 matrix product is exact here: every partial sum is an
    integer below 255 * d <= 2^24 for d <= 65793.  qcodes u8[nq, d], packed u8[n, ceil(d/8)] -> int32[nq, n]."""
    qcodes = np.ascontiguousarray(qcodes, np.uint8)
    packed = np.ascontiguousarray(packed, np.uint8)
    assert 255 * d <= (1 << 24)
    qf = qcodes[:, :d].astype(np.float32).T.copy()            # [d, nq]
    n = packed.shape[0]
    out = np.empty((qcodes.shape[0], n), np.int32)
    for r0 in range(0, n, chunk):
        bits = np.unpackbits(packed[r0:r0 + chunk], axis=1, bitorder="big")[:, :d].astype(np.float32)
        out[:, r0:r0 + chunk] = (bits @ qf).T.astype(np.int32)
    return out


def qcdist_1bit(qpacked, packed, d):
    qpacked = np.ascontiguousarray(qpacked, np.uint8)
    packed = np.ascontiguousarray(packed, np.uint8)
    n = packed.shape[0]
    out = np.empty(n, np.int32)
    lib().bbqo_qcdist_1bit(_p(qpacked, C.c_uint8), _p(packed, C.c_uint8), n, d, _p(out, C.c_int32))
    return out


def dot_unpacked(q, x):
    q = np.ascontiguousarray(q, np.uint8)
    x = np.ascontiguousarray(x, np.uint8)
    return int(lib().bbqo_dot_unpacked(_p(q, C.c_uint8), _p(x, C.c_uint8), q.size))


def centroid_dp(c):
    c = _f32(c)
    return float(lib().bbqo_centroid_dp(_p(c, C.c_float), c.size))


def batch_scores(dots, xcorr, qcorr, d, cdp, sim, query_bits):
    dots = np.ascontiguousarray(dots, np.int32)
    xcorr = np.ascontiguousarray(xcorr, np.float64)
    qcorr = np.ascontiguousarray(qcorr, np.float64)
    n = dots.size
    out = np.empty(n, np.float32)
    lib().bbqo_scores(_p(dots, C.c_int32), _p(xcorr, C.c_double), n, _p(qcorr, C.c_double), d, cdp,
                      SIM[sim], query_bits, _p(out, C.c_float))
    return out


def scale_mip(s):
    return float(lib().bbqo_scale_mip(s))


def topk(scores, k, mode="canonical"):
    scores = _f32(scores)
    n = scores.size
    kk = max(0, min(k, n))
    idx = np.empty(max(kk, 1), np.int32)
    sc = np.empty(max(kk, 1), np.float32)
    fn = lib().bbqo_topk_heap if mode == "heap" else lib().bbqo_topk_canonical
    cnt = fn(_p(scores, C.c_float), n, k, _p(idx, C.c_int32), _p(sc, C.c_float))
    return idx[:cnt].copy(), sc[:cnt].copy()


def search_nearest_neighbors(query, index: OracleIndex, k, query_bits=4, lam=0.1, iters=5,
                             mode="canonical", want_all=False):
    """-> (idx i32[<=k], score f32[<=k]) [, all_scores f32[n], all_dots i32[n]]"""
    q = _f32(query)
    if q.size != index.dim:
        raise ValueError("查询向量维度与目标向量维度不匹配")
    if k < 0:
        raise ValueError("k值不能为负数")
    n = index.size()
    kk = max(0, min(k, n))
    idx = np.empty(max(kk, 1), np.int32)
    sc = np.empty(max(kk, 1), np.float32)
    alls = np.empty(n, np.float32) if want_all else None
    alld = np.empty(n, np.int32) if want_all else None
    if index.index_bits == 1:
        cnt = lib().bbqo_search(_p(q, C.c_float), _p(index.centroid, C.c_float), _p(index.packed, C.c_uint8),
                                _p(index.corr, C.c_double), n, index.dim, SIM[index.sim], query_bits, lam, iters,
                                k, 0 if mode == "heap" else 1, _p(idx, C.c_int32), _p(sc, C.c_float),
                                _p(alls, C.c_float), _p(alld, C.c_int32))
    else:
        # EXTENSION (the reference throws for indexBits >= 2: SURVEY §8 a5/a11): see score_ext in bbq_oracle.cpp;
        # `packed` holds the unpacked codes (n x d bytes) as BinarizedByteVectorValuesImpl would
        cnt = lib().bbqo_search_ext(_p(q, C.c_float), _p(index.centroid, C.c_float), _p(index.packed, C.c_uint8),
                                    _p(index.corr, C.c_double), n, index.dim, SIM[index.sim], query_bits,
                                    index.index_bits, lam, iters, k, 0 if mode == "heap" else 1, _p(idx, C.c_int32),
                                    _p(sc, C.c_float), _p(alls, C.c_float), _p(alld, C.c_int32))
    if want_all:
        return idx[:cnt].copy(), sc[:cnt].copy(), alls, alld
    return idx[:cnt].copy(), sc[:cnt].copy()


def quick_search(query, rows, k, sim="COSINE"):
    """src/index.ts:95-111: rebuilds the index on every call (lambda=0.1, iters=5, 4b x 1b)."""
    index = quantize_vectors(rows, sim=sim, index_bits=1, lam=0.1, iters=5)
    return search_nearest_neighbors(query, index, k, query_bits=4, lam=0.1, iters=5, mode="heap")


def cosine_similarity(a, b):
    """src/vectorSimilarity.ts:75-102"""
    a, b = _f32(a), _f32(b)
    return float(lib().bbqo_cosine(_p(a, C.c_float), _p(b, C.c_float), a.size))


def oversampled_topk(query, rows, index: OracleIndex, k, factor, query_bits=4, lam=0.1, iters=5, mode="sort"):
    """src/topKSelector.ts: getOversampledTopKWithSort (:90-114, mode="sort": stable sort by trueScore desc, the
    canonical contract) / getOversampledTopKWithHeap (:29-78, mode="heap").
    -> (idx i32[<=k], quantizedScore f32, trueScore f64)"""
    rows = _f32(rows)
    ci, cs = search_nearest_neighbors(query, index, k * factor, query_bits=query_bits, lam=lam, iters=iters,
                                      mode="heap" if mode == "heap" else "canonical")
    true = np.array([cosine_similarity(query, rows[i]) for i in ci], np.float64)
    if mode == "heap":
        pos = np.empty(max(min(k, len(ci)), 1), np.int32)
        cnt = lib().bbqo_rerank_heap(_p(true, C.c_double), len(ci), k, _p(pos, C.c_int32))
        pos = pos[:cnt]
    else:
        pos = np.argsort(-true, kind="stable")[:k]
    return ci[pos], cs[pos], true[pos]
