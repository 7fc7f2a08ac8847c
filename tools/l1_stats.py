"""First-level screen statistics of the tensor-core scan (BBQ_MMA_DEBUG bit 256): chunks tested / passed, and one
sample of the quantities involved.  python tools/l1_stats.py [sim] [rows] [nq]"""
import ctypes as C, os, struct, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["BBQ_MMA_DEBUG"] = "256"
import torch, bbq_b200
sim = sys.argv[1] if len(sys.argv) > 1 else "COSINE"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
nq = int(sys.argv[3]) if len(sys.argv) > 3 else 1024
dim = 1024
fmt = bbq_b200.createBinaryQuantizationFormat({"queryBits": 4, "indexBits": 1, "quantizer": {"similarityFunction": sim, "lambda": 0.1, "iters": 5}})
g = torch.Generator(device="cuda"); g.manual_seed(1)
rows = torch.randn((n, dim), generator=g, device="cuda")
qv = fmt.quantizeVectorsDevice(rows.data_ptr(), n, dim, centroid=np.zeros(dim, np.float32))["quantizedVectors"]
qs = np.random.default_rng(2).standard_normal((nq, dim), dtype=np.float32)
fmt.searchBatch(qs, qv, 10)
buf = (C.c_longlong * (4 * 4096))()
assert bbq_b200._native.load().bbq_debug_trace(fmt._ctx, buf, 4 * 4096) == 0
t = list(buf)[3 * 4096:3 * 4096 + 40]
f = lambda v: struct.unpack("f", struct.pack("i", int(v) & 0xFFFFFFFF if v >= 0 else int(v)))[0] if -2**31 <= v < 2**31 else float("nan")
fi = lambda v: struct.unpack("f", struct.pack("I", int(v) & 0xFFFFFFFF))[0]
print(sim, "row-chunks tested", t[0], "passed", t[1], "rate", t[1] / max(t[0], 1), "warp-chunks going on", t[2], "rate", t[2] / max(t[0] / 32, 1), "candidates", fmt.stats()["last_candidates"])
print(" sample: T", t[8], "m", t[9], "qoff0", t[10], "D0", t[11])
print(" rho0", [fi(v) for v in t[12:16]], "cbar", [fi(v) for v in t[16:20]], "hdev", [fi(v) for v in t[20:24]], "row", [fi(v) for v in t[24:28]])
