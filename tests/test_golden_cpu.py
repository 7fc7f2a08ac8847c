"""The oracle must reproduce the committed golden fixtures bit-for-bit (guards against compiler / libm /
numpy drift between the box that made them and the box that runs the GPU suite).  No GPU."""
import numpy as np
import pytest

from oracle import oracle as O
from tests.golden_util import golden_cases, load_golden


@pytest.mark.parametrize("name", golden_cases())
def test_oracle_reproduces_golden(name):
    g = load_golden(name)
    idx = O.quantize_vectors(g["base"], sim=g["sim"], index_bits=1, lam=g["lam"], iters=g["iters"])
    assert np.array_equal(idx.centroid.view(np.uint32), g["centroid"].view(np.uint32))
    assert np.array_equal(idx.packed[:16], g["packed_head"])
    assert np.array_equal(idx.corr[:16].view(np.uint64), g["corr_head"].view(np.uint64))
    assert np.frombuffer(idx.packed.tobytes(), np.uint8).astype(np.uint64).sum() == g["packed_crc"]
    assert np.bitwise_xor.reduce(idx.corr.view(np.uint64).ravel()) == g["corr_bits_xor"]
    for qi, q in enumerate(g["queries"]):
        i, s, alls, alld = O.search_nearest_neighbors(q, idx, g["k"], query_bits=g["query_bits"], lam=g["lam"],
                                                      iters=g["iters"], mode="canonical", want_all=True)
        assert i.tolist() == g["top_idx"][qi].tolist()
        assert np.array_equal(s.view(np.uint32), g["top_score"][qi].view(np.uint32))
        assert np.array_equal(alld[:32], g["dots_head"][qi])
        assert np.bitwise_xor.reduce(alls.view(np.uint32)) == g["score_xor"][qi]
