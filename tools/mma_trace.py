"""Dumps the hand-off timeline of the tensor-core scan (CTA 0): BBQ_MMA_DEBUG=32 python tools/mma_trace.py"""
import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["BBQ_MMA_DEBUG"] = os.environ.get("BBQ_MMA_DEBUG", "32")
import torch, bbq_b200
n, dim, nq = 400000, 1024, int(os.environ.get("NQ", 1024))
fmt = bbq_b200.createBinaryQuantizationFormat({"quantizer": {"similarityFunction": os.environ.get("SIM", "COSINE"), "lambda": 0.1, "iters": 5}})
ix = fmt.reserveIndex(n, dim, np.zeros(dim, np.float32))
g = torch.Generator(device="cuda"); g.manual_seed(1)
for c in range(0, n, 65536):
    r = torch.randn((min(65536, n - c), dim), generator=g, device="cuda")
    fmt.appendRows(ix, d_rows_ptr=r.data_ptr(), n=r.shape[0])
qs = np.random.default_rng(2).standard_normal((nq, dim), dtype=np.float32)
for _ in range(3):
    fmt.searchBatch(qs, ix, 10)
buf = np.zeros(4 * 4096, np.int64)
L = bbq_b200._native.load()
assert L.bbq_debug_trace(fmt._ctx, buf.ctypes.data, buf.size) == 0
mma = buf[:4000].reshape(1000, 4)[:, :3]
exp = buf[4096:4096 + 8000].reshape(1000, 8)[:, :7]
mma = mma[mma[:, 0] > 0]; exp = exp[exp[:, 0] > 0]
t0 = min(mma[0, 0], exp[0, 0])
print("MMA warp per chunk: wait_start, wait_done(+), issued+commit(+)  [cycles]")
for i in range(40, 52):
    a, b, c = mma[i] - t0
    print(f"  chunk {i:3d}: t={a:8d}  wait {b - a:5d}  issue {c - b:5d}   next-start gap {mma[i + 1, 0] - mma[i, 2]:5d}")
d = np.diff(mma[:, 0]); print("MMA issuer-0 chunk period (every 2nd chunk): median", np.median(d[8:]), "mean", d[8:].mean())
print("MMA wait median", np.median(mma[8:, 1] - mma[8:, 0]), "issue median", np.median(mma[8:, 2] - mma[8:, 1]))
print("expansion warp 4 per hand-off [cycles]: expand+loads, stage waits, tcgen05.st issue, wait::st, fence+syncwarp+arrive | until wait::st")
for i in range(20, 32):
    x = exp[i]
    print(f"  hand-off {i:3d}: t={x[0]-t0:8d} expand {x[1]:5d} stage-wait {x[2]:5d} st-issue {x[3]:5d} wait_st {x[4]:5d} arrive {x[5]:5d} | {x[6]:5d}")
dd = np.diff(exp[:, 0]); print("expansion hand-off period: median", np.median(dd[4:]), "mean", dd[4:].mean())
for nm, j in (("expand", 1), ("stage-wait", 2), ("st-issue", 3), ("wait_st", 4), ("arrive", 5)):
    print(f"  {nm:10s} median {np.median(exp[4:, j]):7.0f} mean {exp[4:, j].mean():7.0f}")
