"""Writes the seeded inputs of tests/golden/make_golden.py as raw little-endian f32 files a JS runtime can read:
<case>.in.bin = int32 header [n, dim, nq, k, query_bits, sim(0 EUCLIDEAN|1 COSINE|2 MIP), iters] + f64 lambda +
f32 base[n*dim] + f32 queries[nq*dim]."""
import os
import struct
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, ROOT)
from tests.golden.make_golden import CASES, case_inputs  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
SIM = {"EUCLIDEAN": 0, "COSINE": 1, "MAXIMUM_INNER_PRODUCT": 2}

for name, (n, dim, sim, qb, k, nq, lam, iters, data) in CASES.items():
    base, queries = case_inputs(name)
    with open(os.path.join(HERE, name + ".in.bin"), "wb") as f:
        f.write(struct.pack("<7id", n, dim, nq, k, qb, SIM[sim], iters, lam))
        f.write(np.ascontiguousarray(base, "<f4").tobytes())
        f.write(np.ascontiguousarray(queries, "<f4").tobytes())
    print("wrote", name)
