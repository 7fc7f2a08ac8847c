"""The CUDA path against the fixtures generated from the reference's own source text (tests/golden/from_ts/*.ts.json,
see its README) — DIRECTLY, not through the oracle: index build (centroid, codes, f64 correctives), every row's f32
score, and the top-k lists.  The library returns the canonical order (score desc, row id asc); the reference's MinHeap
returns exact ties in heap order, so lists are compared as score sequences (identical whatever the tie order) and as
index sets whenever no tie straddles the k-th place.  (File name: sorted last on purpose — a cheap cross-check after
the parity suites proper.)"""
import glob
import json
import os
import struct

import numpy as np
import pytest

from tests.golden.make_golden import CASES, case_inputs
from tests.test_gpu_parity import make_format

pytestmark = pytest.mark.gpu

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "from_ts")
FIXTURES = sorted(glob.glob(os.path.join(HERE, "*.ts.json")))


@pytest.fixture(scope="module")
def bbq():
    import bbq_b200
    bbq_b200.build_library()
    return bbq_b200


def check_against_fixture(fmt, d, name):
    """fmt: anything with the host interface's quantizeVectors / searchBatch / debugScores (the GPU format here; the CPU
    dry run of this very function in tests/test_golden_from_ts.py hands in an oracle-backed stand-in)."""
    n, dim, sim, qb, k, nq, lam, iters, data = CASES[name]
    base, queries = case_inputs(name)
    qv = fmt.quantizeVectors(base)["quantizedVectors"]
    assert np.ascontiguousarray(qv.getCentroid(), np.float32).view(np.uint32).tolist() == d["centroid_bits"]
    packed, corr = qv.exportAll()
    want_corr = np.array([[struct.unpack(">d", bytes.fromhex(h))[0] for h in row] for row in d["corrections_bits"]], np.float64)
    assert np.array_equal(np.ascontiguousarray(corr, np.float64).view(np.uint64), want_corr.view(np.uint64))
    assert int(np.asarray(packed).astype(np.uint64).sum()) == d["packed_sum"]
    assert np.asarray(packed)[:16].tolist() == d["packed_head"]
    gi, gs = fmt.searchBatch(queries, qv, k)
    for qi, (q, ref) in enumerate(zip(queries, d["queries"])):
        scores = np.ascontiguousarray(fmt.debugScores(q, qv), np.float32)
        assert scores.view(np.uint32).tolist() == ref["all_score_bits"], (name, qi)
        got_s = np.ascontiguousarray(gs[qi], np.float32)
        assert got_s.view(np.uint32).tolist() == ref["top_score_bits"], (name, qi)
        ordered = np.sort(scores)[::-1]
        if len(ordered) > k and ordered[k - 1] != ordered[k]:          # no exact tie across the k-th place: same set
            assert sorted(np.asarray(gi[qi]).tolist()) == sorted(ref["top_index"]), (name, qi)
        if len(set(ref["top_score_bits"])) == len(ref["top_score_bits"]) and (len(ordered) <= k or ordered[k - 1] != ordered[k]):
            assert np.asarray(gi[qi]).tolist() == ref["top_index"], (name, qi)   # no ties at all: same list


@pytest.mark.parametrize("path", FIXTURES, ids=[os.path.basename(p)[:-len(".ts.json")] for p in FIXTURES])
def test_gpu_equals_the_typescript_reference(bbq, path):
    d = json.load(open(path))
    name = os.path.basename(path)[:-len(".ts.json")]
    n, dim, sim, qb, k, nq, lam, iters, data = CASES[name]
    check_against_fixture(make_format(bbq, sim, qb=qb, lam=lam, iters=iters), d, name)


def test_error_behaviour_equals_the_executed_reference(bbq):
    """tests/golden/from_ts/errors.behaviour.json: which Error the reference throws for invalid inputs (message text and,
    for NaN / Infinity, the position it names under each similarity function), observed by executing its source.  The
    host interface over the C ABI must raise the same messages."""
    from tests.fixtures import gaussian
    d = json.load(open(os.path.join(HERE, "errors.behaviour.json")))
    rows, qs = gaussian(d["n"], d["dim"], d["rows_seed"]), gaussian(6, d["dim"], d["queries_seed"])
    formats = {}

    def setup(sim):
        if sim not in formats:
            fmt = make_format(bbq, sim)
            formats[sim] = (fmt, fmt.quantizeVectors(rows)["quantizedVectors"])
        return formats[sim]

    def run_case(c):
        sim, label = c["sim"], c["label"]
        fmt, qv = setup(sim)
        if c["what"] == "search" and "edits" in c:
            q = qs[2].copy()
            for pos, val in c["edits"].items():
                q[int(pos)] = float(val)
            return fmt.searchNearestNeighbors(q, qv, 5)
        if c["what"] == "build" and "row" in c:
            b = rows[:5].copy()
            b[c["row"], c["pos"]] = float(c["value"])
            return fmt.quantizeVectors(b)
        if label == "k = -1":
            return fmt.searchNearestNeighbors(qs[0], qv, -1)
        if label.startswith("k = 0"):
            assert fmt.searchNearestNeighbors(qs[0], qv, 0) == []
            return None
        if label == "dimension mismatch":
            return fmt.searchNearestNeighbors(qs[0][:10], qv, 3)
        if label == "null query":
            return fmt.searchNearestNeighbors(None, qv, 3)
        if label == "null targets":
            return fmt.searchNearestNeighbors(qs[0], None, 3)
        if label == "empty vector set":
            return fmt.quantizeVectors([])
        if label.startswith("ragged rows"):
            return fmt.quantizeVectors([rows[0], rows[1][:10]])
        if label == "queryBits = 9":
            return bbq.createBinaryQuantizationFormat({"queryBits": 9, "quantizer": {"similarityFunction": "COSINE"}})
        if label == "indexBits = 0":
            return bbq.createBinaryQuantizationFormat({"indexBits": 0, "quantizer": {"similarityFunction": "COSINE"}})
        raise AssertionError(f"unknown case {label}")

    for c in d["cases"]:
        if c["message"] is None:
            run_case(c)
            continue
        with pytest.raises(bbq.BbqError) as e:
            run_case(c)
        assert str(e.value) == c["message"], (c["sim"], c["label"])
