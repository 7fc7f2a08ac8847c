"""A stand-in for the N-API addon (`bbq_b200.node`: bindings/napi/bbq_napi.c over libbbq_b200.so) with the CPU ORACLE as
its compute engine, for executing the TypeScript drop-in class (bindings/ts/binaryQuantizationFormat.gpu.ts) under
tests/golden/from_ts/tsinterp.py — no Node and no GPU exist where the CPU tests run.  Same function names, argument
order, result shapes and error codes ("BBQ:<status>:<vector>:<position>") as the real addon; the validation rules are
the library's (csrc/bbq_api.cu:validate_rows, csrc/bbq_kernels.cuh:query_verdict).  Test infrastructure only."""
import math
import os
import sys
from array import array

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "golden", "from_ts"))
import tsinterp as T  # noqa: E402
from oracle import oracle as O  # noqa: E402

SIMS = ["EUCLIDEAN", "COSINE", "MAXIMUM_INNER_PRODUCT"]


def throw_status(status, vec=-1, pos=-1, message="bbq error"):
    raise T.JSThrow(T.JSObj(T.ERR["Error"], {"message": message, "name": "Error", "stack": "",
                                             "code": f"BBQ:{status}:{vec}:{pos}"}))


def f32(ta, shape=None):
    a = np.frombuffer(ta.a, dtype=np.float32).copy()
    return a.reshape(shape) if shape is not None else a


def typed(kind, values):
    code = T.KIND_CODE[kind]
    return T.TypedArray(kind, array(code, values))


def first_bad(row, cosine, normalize_times=1):
    """-> None | (status, position): the library's verdict for one vector (status 5 NaN, 6 Infinity).  COSINE vectors are
    normalised before the check: once for rows and for quantizeQueryVector, twice on the search path."""
    nan = np.flatnonzero(np.isnan(row))
    inf = np.flatnonzero(np.isinf(row))
    if nan.size == 0 and inf.size == 0:
        return None
    if cosine:
        return (5, 0) if (nan.size or normalize_times >= 2) else (5, int(inf[0]))
    fn = int(nan[0]) if nan.size else 1 << 62
    fi = int(inf[0]) if inf.size else 1 << 62
    return (5, fn) if fn < fi else (6, fi)


class OracleAddon:
    def __init__(self):
        self.calls = []

    def exports(self):
        names = ("create build info search searchSharded rows attachRows searchRerank saveIndex loadIndex quantizeQuery accuracy "
                 "fromQuantized commUniqueId commInit setBase").split()
        out = {}
        for n in names:
            def wrap(*a, _n=n):
                self.calls.append(_n)
                return getattr(self, "f_" + _n)(*a)
            out[n] = wrap
        return out

    # -- the addon's functions ---------------------------------------------------------------------------------
    def f_create(self, query_bits, index_bits, sim, lam, iters, device):
        qb, ib = int(query_bits), int(index_bits)
        if not 1 <= qb <= 8:
            throw_status(1)
        if not 1 <= ib <= 8:
            throw_status(2)
        if ib > 2:
            throw_status(9, message="indexBits > 2")
        return {"__ctx__": True, "qb": qb, "ib": ib, "sim": SIMS[int(sim)], "lam": float(lam), "iters": int(iters)}

    def f_build(self, ctx, rows, n, dim, centroid):
        n, dim = int(n), int(dim)
        if n == 0:
            throw_status(3)
        m = f32(rows, (n, dim))
        for i in range(n):
            bad = first_bad(m[i], ctx["sim"] == "COSINE")
            if bad is not None:
                throw_status(bad[0], i, bad[1])
        cen = f32(centroid) if isinstance(centroid, T.TypedArray) else None
        idx = O.quantize_vectors(m, sim=ctx["sim"], index_bits=ctx["ib"], lam=ctx["lam"], iters=ctx["iters"], centroid=cen)
        return {"__index__": True, "idx": idx, "ctx": ctx, "rows": None}

    def f_info(self, h):
        idx = h["idx"]
        return {"size": float(idx.packed.shape[0]), "dimension": float(idx.dim), "centroid": typed("Float32Array", idx.centroid.tolist()),
                "centroidDP": float(O.centroid_dp(idx.centroid))}

    def _queries(self, h, queries, nq):
        ctx, idx = h["ctx"], h["idx"]
        q = f32(queries, (int(nq), idx.dim))
        for i in range(int(nq)):
            bad = first_bad(q[i], ctx["sim"] == "COSINE", normalize_times=2)
            if bad is not None:
                throw_status(bad[0], i, bad[1])
        return q

    def f_search(self, h, queries, nq, k):
        ctx, idx = h["ctx"], h["idx"]
        if k < 0:
            throw_status(7)
        q = self._queries(h, queries, nq)
        kk = int(min(k, idx.packed.shape[0]))
        ind, sc = [], []
        for row in q:
            i, s = O.search_nearest_neighbors(row, idx, kk, query_bits=ctx["qb"], lam=ctx["lam"], iters=ctx["iters"], mode="canonical")
            ind.extend(int(x) for x in i)
            sc.extend(float(x) for x in s)
        return {"indices": typed("Int32Array", ind), "scores": typed("Float32Array", sc), "count": float(kk), "stride": float(kk)}

    f_searchSharded = f_search

    def f_rows(self, h, first, count):
        idx, a, b = h["idx"], int(first), int(first) + int(count)
        if a < 0 or b > idx.packed.shape[0]:
            throw_status(10, message="row range")
        body = idx.packed if h["ctx"]["ib"] == 1 else idx.unpacked
        return {"packed": typed("Uint8Array", body[a:b].ravel().tolist()), "corrections": typed("Float64Array", idx.corr[a:b].ravel().tolist())}

    def f_attachRows(self, h, rows):
        h["rows"] = f32(rows, (h["idx"].packed.shape[0], h["idx"].dim))
        return T.UNDEF

    def f_searchRerank(self, h, queries, nq, k, factor):
        if h["rows"] is None:
            throw_status(10, message="no original rows attached")
        ctx, idx = h["ctx"], h["idx"]
        q = self._queries(h, queries, nq)[0]
        i, qs, ts = O.oversampled_topk(q, h["rows"], idx, int(k), int(factor), query_bits=ctx["qb"], lam=ctx["lam"], iters=ctx["iters"], mode="sort")
        return {"indices": typed("Int32Array", [int(x) for x in i]), "quantizedScores": typed("Float32Array", [float(x) for x in qs]),
                "trueScores": typed("Float64Array", [float(x) for x in ts]), "count": float(len(i))}

    def f_quantizeQuery(self, ctx, query, centroid):
        q, c = f32(query), f32(centroid)
        bad = first_bad(q, ctx["sim"] == "COSINE")
        if bad is not None:
            throw_status(bad[0], -1, bad[1])
        codes, corr = O.quantize_query_vector_once(q, c, ctx["sim"], ctx["qb"], ctx["lam"], ctx["iters"])
        return {"codes": typed("Uint8Array", codes.tolist()), "corrections": typed("Float64Array", corr.tolist())}

    def f_accuracy(self, ctx, rows, queries, n, dim, target):
        n, dim = int(n), int(dim)
        if ctx["qb"] not in (1, 4):
            throw_status(9, message="unsupported query bits")
        st = O.compute_quantization_accuracy(f32(rows, (n, dim)), f32(queries, (n, dim)), ctx["sim"], ctx["qb"], ctx["lam"], ctx["iters"],
                                             target=int(target))
        return typed("Float64Array", [st[f] for f in ("meanError", "maxError", "minError", "stdError", "correlation")])

    def f_fromQuantized(self, ctx, packed, corrections, centroid, n, dim):
        n, dim = int(n), int(dim)
        p = (dim + 7) // 8
        pk = np.frombuffer(packed.a, dtype=np.uint8).copy().reshape(n, p)
        cr = np.frombuffer(corrections.a, dtype=np.float64).copy().reshape(n, 4)
        idx = O.OracleIndex(f32(centroid), pk, None, cr, dim, ctx["sim"], 1)
        return {"__index__": True, "idx": idx, "ctx": ctx, "rows": None}

    def _unsupported(self, *a):
        throw_status(9, message="not available in the CPU stand-in of the addon")

    f_saveIndex = f_loadIndex = f_commUniqueId = f_commInit = f_setBase = _unsupported


def dropin_interp_kwargs(reference_root, repo_root, addon):
    """Interp(...) arguments that install the drop-in into the reference tree the way INTEGRATION.md says."""
    src = os.path.join(os.path.realpath(reference_root), "src")
    ts = os.path.join(repo_root, "better-binary-quantization_b200", "bindings", "ts")
    exports = addon.exports()
    return dict(require=lambda spec: exports if spec.endswith("bbq_b200.node") else T.throw_type_error(f"Cannot find module {spec}"),
                path_overrides={os.path.join(src, "binaryQuantizationFormat.ts"): os.path.join(ts, "binaryQuantizationFormat.gpu.ts"),
                                os.path.join(src, "binaryQuantizationFormat.cpu.ts"): os.path.join(src, "binaryQuantizationFormat.ts"),
                                os.path.join(src, "errors.ts"): os.path.join(ts, "errors.ts")})


def load_dropin_package(reference_root, repo_root, console=None):
    """The reference package with the drop-in installed as INTEGRATION.md says: src/binaryQuantizationFormat.ts becomes
    the GPU class, the original is reachable as src/binaryQuantizationFormat.cpu.ts, errors.ts is added;
    src/index.ts stays the reference's own file.  -> (interp, exports of src/index.ts, the addon stand-in)"""
    src = os.path.join(os.path.realpath(reference_root), "src")
    addon = OracleAddon()
    interp = T.Interp(log=(lambda *a: console.append(" ".join(map(str, a)))) if console is not None else None,
                      stub_modules=["/src/wasm/index.ts"], **dropin_interp_kwargs(reference_root, repo_root, addon))
    return interp, interp.load(os.path.join(src, "index.ts")), addon
