"""Generates tests/golden/*.npz from the CPU oracle (run here, committed with its outputs).

These are ORACLE outputs on seeded inputs.  They pin (a) the oracle against drift between compilers / boxes and (b) the
CUDA path on the GPU box, where /root/reference does not exist.  The oracle itself is pinned to the reference by
tests/golden/from_ts/*.ts.json — the same seeded inputs run through the reference's own source text
(tests/golden/from_ts/make_golden_with_interp.py, tests/test_golden_from_ts.py).

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402
from tests.fixtures import edge_dataset, gaussian, sincos_dataset  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

CASES = {
    # name: (n, dim, sim, query_bits, k, nq, lam, iters, data)
    "c1_cosine_1000x128": (1000, 128, "COSINE", 4, 10, 4, 0.1, 5, "gauss"),          # BASELINE configs[0]
    "mip_3000x96_k100": (3000, 96, "MAXIMUM_INNER_PRODUCT", 4, 100, 3, 0.1, 5, "gauss"),
    "euclid_2000x100": (2000, 100, "EUCLIDEAN", 4, 10, 3, 0.1, 5, "gauss"),           # dim % 8 != 0
    "cosine_1bit_query_500x64": (500, 64, "COSINE", 1, 10, 3, 0.1, 5, "gauss"),
    "sincos_recall_100x128": (100, 128, "COSINE", 4, 10, 10, 0.001, 20, "sincos"),    # tests/recall.test.ts fixture
    # degenerate rows / exact ties / zero query (tests/fixtures.py:edge_dataset) — fixtures from the reference source only
    # (tests/golden/from_ts/*.ts.json); no .npz is generated for them
    "edge_cosine_60x40": (60, 40, "COSINE", 4, 8, 3, 0.1, 5, "edge"),
    "edge_euclid_60x40": (60, 40, "EUCLIDEAN", 4, 8, 3, 0.1, 5, "edge"),
    "edge_mip_1bit_query_60x40": (60, 40, "MAXIMUM_INNER_PRODUCT", 1, 8, 3, 0.1, 5, "edge"),
    # other query widths on a 1-bit index (the reference's batch path takes its "4-bit" branch for every width != 1)
    "qb8_cosine_60x40": (60, 40, "COSINE", 8, 8, 2, 0.1, 5, "gauss"),
    "qb2_euclid_60x40": (60, 40, "EUCLIDEAN", 2, 8, 2, 0.1, 5, "gauss"),
    "qb7_mip_50x48": (50, 48, "MAXIMUM_INNER_PRODUCT", 7, 8, 2, 0.1, 5, "gauss"),
}
FROM_TS_ONLY = {"edge_cosine_60x40", "edge_euclid_60x40", "edge_mip_1bit_query_60x40", "qb8_cosine_60x40", "qb2_euclid_60x40",
                "qb7_mip_50x48"}      # fixtures from the reference source only: no .npz


def case_inputs(name):
    """-> (base f32[n, dim], queries f32[nq, dim]) of a case; every generator and test goes through here."""
    n, dim, sim, qb, k, nq, lam, iters, data = CASES[name]
    seed = 20260101 + sum(map(ord, name))
    if data == "gauss":
        return gaussian(n, dim, seed), gaussian(nq, dim, seed + 100)
    if data == "edge":
        return edge_dataset(n, dim, nq, seed)
    return sincos_dataset(dim, n, nq)


def make(name):
    n, dim, sim, qb, k, nq, lam, iters, data = CASES[name]
    base, queries = case_inputs(name)
    gen = {"seed": 20260101 + sum(map(ord, name)) if data != "sincos" else -1}
    idx = O.quantize_vectors(base, sim=sim, index_bits=1, lam=lam, iters=iters)
    out = {"n": n, "dim": dim, "sim": sim, "query_bits": qb, "k": k, "lam": lam, "iters": iters, **gen,
           "centroid": idx.centroid, "packed_head": idx.packed[:16], "corr_head": idx.corr[:16],
           "packed_crc": np.frombuffer(idx.packed.tobytes(), np.uint8).astype(np.uint64).sum(),
           "corr_bits_xor": np.bitwise_xor.reduce(idx.corr.view(np.uint64).ravel())}
    top_idx, top_sc, qcodes, qcorr, dots_head, score_xor = [], [], [], [], [], []
    for q in queries:
        i, s, alls, alld = O.search_nearest_neighbors(q, idx, k, query_bits=qb, lam=lam, iters=iters,
                                                      mode="canonical", want_all=True)
        c, r = O.quantize_query_vector(q, idx.centroid, sim=sim, query_bits=qb, lam=lam, iters=iters)
        top_idx.append(i); top_sc.append(s); qcodes.append(c); qcorr.append(r)
        dots_head.append(alld[:32]); score_xor.append(np.bitwise_xor.reduce(alls.view(np.uint32)))
    out.update(top_idx=np.stack(top_idx), top_score=np.stack(top_sc), qcodes=np.stack(qcodes),
               qcorr=np.stack(qcorr), dots_head=np.stack(dots_head), score_xor=np.array(score_xor, np.uint32))
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, "ok", out["top_idx"][0][:5])


if __name__ == "__main__":
    for nm in CASES:
        if nm not in FROM_TS_ONLY:
            make(nm)
