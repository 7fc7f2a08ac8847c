"""Pins the CPU oracle against every known-answer test / fixture the reference's own tests hold for the
hot path.  No GPU.  Exact score values are NOT pinned by the reference ("parity unpinned")."""
import numpy as np
import pytest

from oracle import oracle as O
from tests.fixtures import gaussian, sincos_dataset, true_topk_cosine


# ---- Rust KATs (rust-wasm/src/*.rs inline tests) ------------------------------------------------
def test_kat_batch_four_bit_dot_product():
    # batch_dot_product.rs:137-153
    q = np.arange(1, 9, dtype=np.uint8)
    buf = np.array([[0xFF], [0x00]], np.uint8)
    assert O.qcdist_packed(q, buf, 8).tolist() == [36, 0]
    assert O.qcdist_packed(q, buf, 8, planes=4).tolist() == [36, 0]


def test_kat_int4_bit_dot_product():
    # bitwise_dot_product.rs:114-120
    assert O.dot_unpacked([15, 14, 13, 12], [1, 1, 0, 1]) == 41
    assert O.dot_unpacked([1, 2, 3, 4], [5, 6, 7, 8]) == 70


def test_kat_pack_as_binary():
    # optimized_scalar_quantizer.rs:322-327
    assert O.pack_as_binary([1, 0, 1, 0, 1, 0, 1, 0]).tolist() == [0b10101010]
    # tail bits zero, MSB first (optimizedScalarQuantizer.ts:420-446)
    assert O.pack_as_binary([1, 1, 1, 1, 1, 1, 1, 1, 1, 0, 1]).tolist() == [0xFF, 0b10100000]


def test_kat_scalar_quantize_1bit():
    # optimized_scalar_quantizer.rs:307-319
    codes, corr = O.scalar_quantize([1.0, -1.0, 0.5, -0.5], [0, 0, 0, 0], 1, "EUCLIDEAN")
    assert codes.tolist() == [1, 0, 1, 0]
    assert corr[3] == 2.0


def test_kat_scale_mip():
    # binary_quantized_scorer.rs:336-339
    assert O.scale_mip(1.0) == 2.0 and O.scale_mip(-1.0) == 0.5


# ---- closed forms from tests/utils.test.ts / computeCentroid-correctness.test.ts ---------------------
def test_centroid_known():
    c = O.compute_centroid(np.array([[1, 2, 3], [4, 5, 6], [7, 8, 9]], np.float32))
    assert c.tolist() == [4.0, 5.0, 6.0]


def test_normalize_known():
    v = O.normalize_vector([3.0, 4.0])
    assert np.allclose(v, [0.6, 0.8], atol=1e-7)
    assert O.normalize_vector([0.0, 0.0, 0.0]).tolist() == [0, 0, 0]


# ---- integer identity: packed 4-bit dot == bit-plane AND+popcount == unpacked byte dot --------------
@pytest.mark.parametrize("d", [8, 64, 100, 128, 768, 1021])
def test_qcdist_identities(d):
    rng = np.random.default_rng(d)
    q = rng.integers(0, 16, d, dtype=np.uint8)
    x = rng.integers(0, 2, (37, d), dtype=np.uint8)
    packed = np.stack([O.pack_as_binary(r) for r in x])
    a = O.qcdist_packed(q, packed, d)
    b = O.qcdist_packed(q, packed, d, planes=4)
    c = np.array([O.dot_unpacked(q, r) for r in x])
    assert np.array_equal(a, b) and np.array_equal(a, c)
    # 1-bit query: AND+popcount == byte dot (batchDotProduct.ts:22-49)
    q1 = rng.integers(0, 2, d, dtype=np.uint8)
    assert np.array_equal(O.qcdist_1bit(O.pack_as_binary(q1), packed, d), np.array([O.dot_unpacked(q1, r) for r in x]))


# ---- deterministic recall fixtures (tests/recall.test.ts, tests/recall-all-dimensions.test.ts) -------
def _recall(base, queries, k, query_bits, lam, iters):
    idx = O.quantize_vectors(base, sim="COSINE", index_bits=1, lam=lam, iters=iters)
    tot = 0.0
    for q in queries:
        got, sc = O.search_nearest_neighbors(q, idx, k, query_bits=query_bits, lam=lam, iters=iters, mode="heap")
        assert len(got) == k and np.all(np.diff(sc) <= 0)
        truth = true_topk_cosine(q, base, k)
        tot += len(set(got.tolist()) & set(truth.tolist())) / k
    return tot / len(queries)


def test_recall_fixture_128d():
    base, queries = sincos_dataset(128, 100, 10)
    assert _recall(base, queries, 10, 1, 0.001, 20) >= 0.70   # recall.test.ts:91,163
    assert _recall(base, queries, 10, 4, 0.001, 20) >= 0.60   # recall.test.ts:390,506


@pytest.mark.parametrize("dim,t1,t4", [(384, 0.60, 0.75), (768, 0.55, 0.70), (1024, 0.50, 0.65), (1536, 0.45, 0.60)])
def test_recall_fixture_all_dimensions(dim, t1, t4):
    # recall-common.ts:43-107 thresholds; normalised fixture :112-138
    base, queries = sincos_dataset(dim, 1000, 20, normalise=True)
    assert _recall(base, queries, 10, 4, 0.001, 20) >= t4
    assert _recall(base, queries, 10, 1, 0.001, 20) >= t1


# ---- behavioural properties the reference asserts ----------------------------------------------------
def test_k_edge_cases():
    base = gaussian(7, 32, 1)
    idx = O.quantize_vectors(base, sim="COSINE")
    q = gaussian(1, 32, 2)[0]
    assert len(O.search_nearest_neighbors(q, idx, 0)[0]) == 0          # k == 0 -> []
    assert len(O.search_nearest_neighbors(q, idx, 10)[0]) == 7         # k > N -> N
    with pytest.raises(ValueError):
        O.search_nearest_neighbors(q[:16], idx, 3)                     # dim mismatch
    with pytest.raises(ValueError):
        O.search_nearest_neighbors(q, idx, -1)                         # k < 0


def test_heap_vs_canonical_tie_semantics():
    # SURVEY §7 hard-part 2: k=2, scores [5,5,7] -> heap {1,2}, canonical {0,2}
    s = np.array([5, 5, 7], np.float32)
    assert sorted(O.topk(s, 2, "heap")[0].tolist()) == [1, 2]
    assert O.topk(s, 2, "canonical")[0].tolist() == [2, 0]
    # without boundary ties the sets agree
    r = np.random.default_rng(0).standard_normal(1000).astype(np.float32)
    assert set(O.topk(r, 10, "heap")[0].tolist()) == set(O.topk(r, 10, "canonical")[0].tolist())


@pytest.mark.parametrize("sim", ["EUCLIDEAN", "COSINE", "MAXIMUM_INNER_PRODUCT"])
@pytest.mark.parametrize("qb", [1, 4])
def test_search_consistency(sim, qb):
    """search == scores(all) + selection; scores are f32, descending, heap set == canonical set (no ties)."""
    base, qs = gaussian(500, 96, 3), gaussian(3, 96, 4)
    idx = O.quantize_vectors(base, sim=sim)
    for q in qs:
        i1, s1, alls, alld = O.search_nearest_neighbors(q, idx, 10, query_bits=qb, mode="heap", want_all=True)
        i2, s2 = O.search_nearest_neighbors(q, idx, 10, query_bits=qb, mode="canonical")
        assert np.all(np.diff(s1) <= 0) and alls.dtype == np.float32
        if len(np.unique(alls)) == len(alls):
            assert i1.tolist() == i2.tolist()
        assert np.array_equal(alls[i2], s2)
        qc, _ = O.quantize_query_vector(q, idx.centroid, sim=sim, query_bits=qb)
        assert np.array_equal(alld, idx.unpacked.astype(np.int32) @ qc.astype(np.int32))


# ---- oversampled search + exact re-rank (src/topKSelector.ts; tests/recall.test.ts:519,635,693) --------------------
def test_oversampled_rerank_recall_fixture():
    base, queries = sincos_dataset(128, 100, 10)
    idx = O.quantize_vectors(base, sim="COSINE", index_bits=1, lam=0.001, iters=20)
    over = plain = 0.0
    for q in queries:
        i, qs, ts = O.oversampled_topk(q, base, idx, 10, 3, lam=0.001, iters=20)
        ih, _, _ = O.oversampled_topk(q, base, idx, 10, 3, lam=0.001, iters=20, mode="heap")
        assert set(i.tolist()) == set(ih.tolist()) and np.all(np.diff(ts) <= 0) and len(i) == 10
        truth = set(true_topk_cosine(q, base, 10).tolist())
        over += len(set(i.tolist()) & truth) / 10
        plain += len(set(O.search_nearest_neighbors(q, idx, 10, lam=0.001, iters=20)[0].tolist()) & truth) / 10
    assert over / 10 >= 0.75 and over >= plain        # recall.test.ts:519,635 and :693


def test_cosine_similarity_known():
    # tests/utils.test.ts closed forms
    assert O.cosine_similarity([1, 0, 0], [0, 1, 0]) == 0.0
    assert abs(O.cosine_similarity([1, 2, 3], [2, 4, 6]) - 1.0) < 1e-15
    assert O.cosine_similarity([0, 0, 0], [1, 2, 3]) == 0.0
