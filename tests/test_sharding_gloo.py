"""Multi-rank host logic on CPU (gloo, world_size 2 and 3): row-wise shard bounds, the all_gather layout
[shards][queries][k] and the deterministic merge rule (score desc, global row id asc, -1 = empty slot).
Each rank searches its shard with the ORACLE (this is a test; the product's shards are searched on the GPU),
gathers, merges with the numpy statement of the merge rule, and must reproduce the un-sharded oracle answer."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import bbq_b200
from oracle import oracle as O
from tests.fixtures import gaussian


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, dim, k, nq, sim, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rows, qs = gaussian(n, dim, 7), gaussian(nq, dim, 8)
        full = O.quantize_vectors(rows, sim=sim, want_unpacked=False, centroid=np.zeros(dim, np.float32))
        r0, r1 = bbq_b200.shard_bounds(n, world, rank)
        shard = O.OracleIndex(full.centroid, full.packed[r0:r1], None, full.corr[r0:r1], dim, sim, 1)
        loc_i = np.full((nq, k), -1, np.int32)
        loc_s = np.full((nq, k), -np.inf, np.float32)
        for qi, q in enumerate(qs):
            if r1 > r0:
                i, s = O.search_nearest_neighbors(q, shard, k, mode="canonical")
                loc_i[qi, :len(i)] = i + r0          # global row id = shard base + local row
                loc_s[qi, :len(i)] = s
        ti, ts = torch.from_numpy(loc_i), torch.from_numpy(loc_s)
        gi = [torch.empty_like(ti) for _ in range(world)]
        gs = [torch.empty_like(ts) for _ in range(world)]
        dist.all_gather(gi, ti)
        dist.all_gather(gs, ts)
        mi, ms = bbq_b200.merge_host([t.numpy() for t in gi], [t.numpy() for t in gs], k)
        ok = True
        for qi, q in enumerate(qs):
            wi, ws = O.search_nearest_neighbors(q, full, k, mode="canonical")
            ok &= mi[qi, :len(wi)].tolist() == wi.tolist() and ms[qi, :len(wi)].tolist() == ws.tolist()
        flag = torch.tensor([1 if ok else 0])
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if rank == 0:
            out.put(int(flag.item()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n,k", [(2, 3000, 10), (3, 1000, 25), (2, 5, 10)])
def test_sharded_merge_equals_unsharded(world, n, k):
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, 64, k, 5, "COSINE", out)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert out.get(timeout=5) == 1


def test_shard_bounds_cover_rows_exactly():
    for n in (1, 7, 1000, 1_000_000, 100_000_000):
        for w in (1, 2, 3, 4, 8):
            b = [bbq_b200.shard_bounds(n, w, r) for r in range(w)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            assert all(0 <= lo <= hi for lo, hi in b)


def test_merge_host_tie_rule():
    i1, s1 = np.array([[4, 9, -1]], np.int32), np.array([[5.0, 3.0, -np.inf]], np.float32)
    i2, s2 = np.array([[2, 7, 8]], np.int32), np.array([[5.0, 5.0, 1.0]], np.float32)
    mi, ms = bbq_b200.merge_host([i1, i2], [s1, s2], 4)
    assert mi.tolist() == [[2, 4, 7, 9]] and ms.tolist() == [[5.0, 5.0, 5.0, 3.0]]


def _id_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        calls = []

        def make_id():          # stands in for bbq_comm_unique_id (needs NCCL + a GPU): rank 0 only may call it
            calls.append(rank)
            return bytes(range(128))

        cid = bbq_b200.broadcast_comm_id(make_id, rank, world)
        out.put((rank, cid == bytes(range(128)), calls))
    finally:
        dist.destroy_process_group()


def test_comm_id_bootstrap_reaches_every_rank():
    """The only thing torch.distributed does for the sharded search: ship rank 0's 128-byte communicator id."""
    world = 3
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_id_worker, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    got = sorted(out.get(timeout=5) for _ in range(world))
    assert got == [(0, True, [0]), (1, True, []), (2, True, [])]
