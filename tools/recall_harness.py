#!/usr/bin/env python
"""Recall harness on the GPU path: recall@k of the quantised search and of the oversampled +
re-ranked search against exact cosine ground truth, swept over dimensions — the reference's
tests/recall-all-dimensions.test.ts / tests/recall-common.ts:188-289 flow, at sizes the CPU reference cannot reach.

    python tools/recall_harness.py --rows 1000000 --queries 256 --dims 384 768 1024 1536 --k 10 --factors 1 3 5

Ground truth: exact f32 inner products of L2-normalised rows by torch (cuBLAS) in row chunks, top-k by
(score desc, index asc) — test plumbing, not the product path.  Prints one JSON line per (dim, factor)."""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def exact_topk(rows_d, queries_d, k, chunk=262144):
    import torch
    qn = torch.nn.functional.normalize(queries_d.double(), dim=1).float()
    best_s = torch.full((queries_d.shape[0], k), -float("inf"), device=rows_d.device)
    best_i = torch.zeros((queries_d.shape[0], k), dtype=torch.int64, device=rows_d.device)
    for r0 in range(0, rows_d.shape[0], chunk):
        blk = torch.nn.functional.normalize(rows_d[r0:r0 + chunk].double(), dim=1).float()
        s = qn @ blk.T
        cs, ci = torch.topk(s, min(k, s.shape[1]), dim=1)
        ms = torch.cat([best_s, cs], 1)
        mi = torch.cat([best_i, ci + r0], 1)
        o = torch.argsort(ms, dim=1, descending=True, stable=True)[:, :k]
        best_s, best_i = torch.gather(ms, 1, o), torch.gather(mi, 1, o)
    return best_i.cpu().numpy()


def recall(found, truth):
    return float(np.mean([len(set(f.tolist()) & set(t.tolist())) / len(t) for f, t in zip(found, truth)]))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=1000000)
    ap.add_argument("--queries", type=int, default=256)
    ap.add_argument("--dims", type=int, nargs="+", default=[384, 768, 1024, 1536])
    ap.add_argument("--k", type=int, default=10)
    ap.add_argument("--factors", type=int, nargs="+", default=[1, 3, 5])
    ap.add_argument("--query-bits", type=int, default=4)
    ap.add_argument("--clusters", type=int, default=0, help="0: i.i.d. N(0,1); >0: Gaussian mixture with this many centres")
    ap.add_argument("--seed", type=int, default=20260101)
    a = ap.parse_args()
    import torch
    import bbq_b200
    bbq_b200.build_library()
    dev = torch.device("cuda:0")
    for dim in a.dims:
        g = torch.Generator(device=dev).manual_seed(a.seed + dim)
        if a.clusters:
            centres = torch.randn(a.clusters, dim, generator=g, device=dev)
            pick = torch.randint(0, a.clusters, (a.rows,), generator=g, device=dev)
            rows = centres[pick] + 0.5 * torch.randn(a.rows, dim, generator=g, device=dev)
            qpick = torch.randint(0, a.clusters, (a.queries,), generator=g, device=dev)
            queries = centres[qpick] + 0.5 * torch.randn(a.queries, dim, generator=g, device=dev)
        else:
            rows = torch.randn(a.rows, dim, generator=g, device=dev)
            queries = torch.randn(a.queries, dim, generator=g, device=dev)
        truth = exact_topk(rows, queries, a.k)
        fmt = bbq_b200.createBinaryQuantizationFormat({
            "queryBits": a.query_bits, "indexBits": 1,
            "quantizer": {"similarityFunction": bbq_b200.VectorSimilarityFunction.COSINE, "lambda": 0.1, "iters": 5}})
        t0 = time.time()
        qv = fmt.quantizeVectorsDevice(rows.data_ptr(), a.rows, dim)["quantizedVectors"]
        build_s = time.time() - t0
        torch.cuda.synchronize()
        fmt.attachOriginalVectorsDevice(qv, rows.data_ptr())
        qh = queries.cpu().numpy()
        for f in a.factors:
            t0 = time.time()
            if f == 1:
                idx, _ = fmt.searchBatch(qh, qv, a.k)
            else:
                idx, _, _ = fmt.searchOversampledBatch(qh, qv, a.k, f)
            dt = time.time() - t0
            print(json.dumps({"dim": dim, "rows": a.rows, "queries": a.queries, "k": a.k, "query_bits": a.query_bits,
                              "oversample_factor": f, "recall_at_k": round(recall(idx, truth), 4),
                              "search_s": round(dt, 4), "build_s": round(build_s, 2),
                              "data": f"gaussian mixture, {a.clusters} centres" if a.clusters else "iid N(0,1)"}), flush=True)
        del qv, fmt, rows
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
