"""The small search that tools/sanitize.sh runs under compute-sanitizer: every kernel family of the path on a
20k-row index — K5 build, K4 query quantiser (both forms), K2 tensor-core scan (sample dump + filtered scan with the
running threshold, the parked-hit ring and the drainer), K1s streaming popcount scan, K1 tile scan, K3 selection
(threshold, keys, dense, pairs merge), the rerank and the accuracy kernels — checked against the oracle so that a
sanitizer-clean run is also a correct run."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bbq_b200  # noqa: E402
from oracle import oracle as O  # noqa: E402

n, dim, k = int(os.environ.get("SAN_ROWS", 20000)), 256, 10
rng = np.random.default_rng(5)
rows = rng.standard_normal((n, dim), dtype=np.float32)
qs = rng.standard_normal((48, dim), dtype=np.float32)
for sim in ("COSINE", "EUCLIDEAN"):
    fmt = bbq_b200.createBinaryQuantizationFormat(
        {"queryBits": 4, "indexBits": 1, "quantizer": {"similarityFunction": sim, "lambda": 0.1, "iters": 5}})
    qv = fmt.quantizeVectors(rows)["quantizedVectors"]                 # K5
    want = O.quantize_vectors(rows, sim=sim, want_unpacked=False)
    bi, bs = fmt.searchBatch(qs, qv, k)                                # K4 warp form, K2 (sample + filter), K3
    assert fmt.stats()["last_engine"] == 2
    si, ss = fmt.searchBatch(qs[:2], qv, k)                            # K1s + K3
    assert fmt.stats()["last_engine"] == 1
    for i in (0, 1, 47):
        wi, ws = O.search_nearest_neighbors(qs[i], want, k, mode="canonical")
        assert bi[i].tolist() == wi.tolist() and bs[i].tolist() == ws.tolist()
    assert np.array_equal(si, bi[:2])
    dots = fmt.debugQcDistBatch(qs[:8], qv)                            # K2 dump with the integer tap
    assert np.array_equal(dots[3], fmt.debugQcDist(qs[3], qv))        # K1s dump
    fmt.attachOriginalVectors(qv, rows)
    fmt.searchOversampledBatch(qs[:4], qv, k, 3)                       # rerank kernels
    fmt.computeQuantizationAccuracy(rows[:64], qs[:48].repeat(2, 0)[:64])
os.environ["BBQ_POPC_FORM"] = "tile"
os.environ["BBQ_SCAN"] = "popc"
os.environ["BBQ_QQUANT"] = "thread"
fmt = bbq_b200.createBinaryQuantizationFormat()
qv = fmt.quantizeVectors(rows[:6000])["quantizedVectors"]
fmt.searchBatch(qs[:40], qv, k)                                        # K1 tile scan, K4 thread form
os.environ["BBQ_FORCE_PATH"] = "2"
fmt = bbq_b200.createBinaryQuantizationFormat()
qv = fmt.quantizeVectors(rows)["quantizedVectors"]
fmt.searchBatch(qs[:3], qv, k)                                         # exact chunked path: dense select + pairs merge
print("sanitize_case ok")
