"""bench.py's output contract, checked on the CPU with the reference arm (the one arm that runs without a GPU):
exactly ONE line on stdout, JSON, with the keys the driver reads."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0", "--workload", "c3", "--ref-rows", "3000", "--ref-threads", "2"],
                         capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["higher_is_better"] is True and d["unit"] == "queries/s"
    for key in ("metric", "value", "n_gpus", "steps", "warmup", "ms_per_step", "scaling", "vs_baseline", "dtype", "data",
                "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] == 2 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["value"] > 0 and "workload" in d["config"]


def test_traffic_table_points_at_committed_captures():
    table = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
    for e in table["captures"]:
        assert os.path.exists(os.path.join(ROOT, e["source"])), e["source"]
        assert e["dram_bytes_per_launch"] >= e["algorithmic_streamed_bytes_per_launch"] * 0.99


def test_clustered_corpus_generators_and_reference_arm():
    """--corpus clustered: rows and queries are centre + noise (host and device generators of one family), and the
    reference arm runs on it and says so in the workload name."""
    import importlib.util
    import numpy as np
    import torch
    spec = importlib.util.spec_from_file_location("bench_under_test", os.path.join(ROOT, "bench.py"))
    b = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(b)
    b.CORPUS = "clustered"
    dim = 48
    cen = b.centres_host(dim)
    assert cen.shape == (b.N_CENTRES, dim)

    def nearest(x):
        return np.sqrt(((x[:, None, :] - cen[None, :, :]) ** 2).sum(-1)).min(1).mean()

    host, dev, qs = b.gen_chunk_host(5, dim, 400), b.gen_chunk_device(torch, 5, dim, 400, device="cpu").numpy(), b.gen_queries(8, dim)
    want = b.CLUSTER_NOISE * np.sqrt(dim)
    for x in (host, dev, qs):
        assert x.dtype == np.float32 and abs(nearest(x) - want) < 0.15 * want
    assert np.array_equal(host, b.gen_chunk_host(5, dim, 400)) and not np.array_equal(host, b.gen_chunk_host(6, dim, 400))
    b.CORPUS = "gaussian"
    assert nearest(b.gen_chunk_host(5, dim, 400)) > 1.5 * want      # i.i.d. rows sit nowhere near a centre
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0", "--workload", "c3", "--ref-rows", "2000", "--ref-threads", "2", "--corpus", "clustered"],
                         capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    d = json.loads([l for l in out.stdout.splitlines() if l.strip()][0])
    assert "[corpus: clustered]" in d["config"]["workload"] and d["value"] > 0
