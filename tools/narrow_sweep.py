"""Same-box, same-process sweep of the tensor-core scan's role layouts for narrow / medium query batches.

For each similarity function one 1 M x 1024 index is built per configuration (a configuration = the BBQ_MMA_* knobs a
context reads when it is created); every batch size runs WARM + STEPS searches with the library's own CUDA-event
profiling on, and the lists of every configuration are compared with those of the first one (the shipped wide-batch
roles) — so one run gives both the timing table and a parity check of the new roles.

  python tools/narrow_sweep.py            # ROWS, NQS, SIMS, CONFIGS, STEPS from the environment
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import bbq_b200  # noqa: E402

ROWS = int(os.environ.get("ROWS", 1_000_000))
DIM = int(os.environ.get("DIM", 1024))
NQS = [int(x) for x in os.environ.get("NQS", "8 16 32 48 64 96 128 208").split()]
SIMS = os.environ.get("SIMS", "COSINE EUCLIDEAN").split()
STEPS = int(os.environ.get("STEPS", 5))
WARM = int(os.environ.get("WARM", 2))
K = int(os.environ.get("K", 10))
# name: {env}
CONFIGS = {
    "wide(e8g1,2iss)": {"BBQ_MMA_LAYOUT": "0", "BBQ_MMA_ISSUERS": "2"},
    "narrow(e4g2,2iss)": {"BBQ_MMA_LAYOUT": "1", "BBQ_MMA_ISSUERS": "2"},
    "narrow(e4g2,3iss)": {"BBQ_MMA_LAYOUT": "1", "BBQ_MMA_ISSUERS": "3"},
    "wide(e8g1,3iss)": {"BBQ_MMA_LAYOUT": "0", "BBQ_MMA_ISSUERS": "3"},
    "narrow,single-chunk": {"BBQ_MMA_LAYOUT": "1", "BBQ_MMA_ISSUERS": "3", "BBQ_MMA_DEBUG": "1024"},
}
if os.environ.get("CONFIGS_JSON"):  # e.g. {"narrow3,no-acc-loads": {"BBQ_MMA_LAYOUT": "1", "BBQ_MMA_ISSUERS": "3", "BBQ_MMA_DEBUG": "4"}}
    import json
    CONFIGS = json.loads(os.environ["CONFIGS_JSON"])
elif os.environ.get("CONFIGS"):
    CONFIGS = {k: v for k, v in CONFIGS.items() if k in os.environ["CONFIGS"].split(";")}
KNOBS = ("BBQ_MMA_LAYOUT", "BBQ_MMA_ISSUERS", "BBQ_MMA_DEBUG", "BBQ_MMA_NARROW_MAX")


def build(fmt):
    ix = fmt.reserveIndex(ROWS, DIM, np.zeros(DIM, np.float32))
    g = torch.Generator(device="cuda")
    g.manual_seed(1)
    for c in range(0, ROWS, 65536):
        r = torch.randn((min(65536, ROWS - c), DIM), generator=g, device="cuda")
        fmt.appendRows(ix, d_rows_ptr=r.data_ptr(), n=r.shape[0])
    torch.cuda.synchronize()
    return ix


def main():
    queries = np.random.default_rng(2).standard_normal((max(NQS), DIM), dtype=np.float32)
    for sim in SIMS:
        base = {}
        table = {}
        for name, env in CONFIGS.items():
            for kn in KNOBS:
                os.environ.pop(kn, None)
            os.environ.update(env)
            fmt = bbq_b200.createBinaryQuantizationFormat(
                {"quantizer": {"similarityFunction": sim, "lambda": 0.1, "iters": 5}})
            ix = build(fmt)
            fmt.setProfiling(True)
            for nq in NQS:
                qs = queries[:nq]
                t_start = time.time()
                for _ in range(WARM):
                    idx, sc = fmt.searchBatch(qs, ix, K)
                fmt.resetProfiling()
                for _ in range(STEPS):
                    idx, sc = fmt.searchBatch(qs, ix, K)
                st = fmt.stats()
                per = st["scan_ms"] / max(st["scan_launches"], 1) / max(st["mma_passes"], 1)
                key = (nq,)
                if name == next(iter(CONFIGS)):
                    base[key] = (idx.copy(), sc.copy())
                    same = "base"
                else:
                    same = "same" if (np.array_equal(idx, base[key][0]) and np.array_equal(
                        sc.view(np.uint32), base[key][1].view(np.uint32))) else "DIFFERENT"
                table[(name, nq)] = per
                print(f"{sim:9s} {name:22s} nq={nq:4d} n_tile={st['mma_n_tile']:3d} passes={st['mma_passes']} "
                      f"layout={st['mma_layout']} scan/pass {per:.4f} ms  overflow={st['last_overflow']} lists {same} "
                      f"({time.time() - t_start:.1f}s)", flush=True)
            del ix, fmt
        print(f"--- {sim}: ms per pass over {ROWS} x {DIM}")
        print("nq".rjust(6) + "".join(n.rjust(24) for n in CONFIGS))
        for nq in NQS:
            print(f"{nq:6d}" + "".join(f"{table[(n, nq)]:24.4f}" for n in CONFIGS))


if __name__ == "__main__":
    main()
