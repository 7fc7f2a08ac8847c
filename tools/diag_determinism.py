"""Diagnostic: repeated 1M x 1024 batched searches on the tensor-core engine vs the popcount engine."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bbq_b200
n, dim, nq, k, sim = 1_000_000, 1024, 1024, 10, os.environ.get("SIM", "EUCLIDEAN")
def mk(**env):
    old = {kk: os.environ.pop(kk, None) for kk in env}
    os.environ.update({kk: str(v) for kk, v in env.items()})
    f = bbq_b200.createBinaryQuantizationFormat({"quantizer": {"similarityFunction": sim, "lambda": 0.1, "iters": 5}})
    for kk in env:
        os.environ.pop(kk, None)
    return f
g = torch.Generator(device="cuda"); g.manual_seed(20260101)
rows = torch.randn((n, dim), generator=g, device="cuda")
cen = np.zeros(dim, np.float32)
qs = np.random.default_rng(20260201).standard_normal((nq, dim), dtype=np.float32)
ref = mk(BBQ_SCAN="popc")
ir = ref.quantizeVectorsDevice(rows.data_ptr(), n, dim, centroid=cen)["quantizedVectors"]
ri, rs = ref.searchBatch(qs, ir, k)
for name, env in (("mma", {"BBQ_SCAN": "mma"}), ("mma dyntau=0", {"BBQ_SCAN": "mma", "BBQ_DYNTAU": 0})):
    f = mk(**env)
    ix = f.quantizeVectorsDevice(rows.data_ptr(), n, dim, centroid=cen)["quantizedVectors"]
    for rep in range(4):
        mi, ms = f.searchBatch(qs, ix, k)
        bad = np.where((mi != ri).any(1))[0]
        st = f.stats()
        print(name, "rep", rep, "queries differing from popcount engine:", len(bad), "cands", st["last_candidates"], "overflow", st["last_overflow"], "path", st["last_path"])
        for q in bad[:3]:
            print("   q", q, "mma", mi[q].tolist(), ms[q].tolist()); print("        ref", ri[q].tolist(), rs[q].tolist())
