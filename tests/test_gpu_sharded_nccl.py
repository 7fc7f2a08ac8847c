"""Sharded search INSIDE the library (bbq_comm_* + bbq_search_sharded): one process per GPU, the communicator id
travels over a multiprocessing queue (no torch.distributed anywhere on this path), every rank builds its row shard,
and every rank must return exactly the lists a single-shard search of the whole corpus returns (== the oracle).
Needs >= 2 GPUs (`gpurun --gpus 2`); skipped on a 1-GPU box."""
import numpy as np
import pytest

from tests.fixtures import gaussian

pytestmark = pytest.mark.gpu


def _worker(rank, world, ids, out, n, dim, nq, k, sim):
    import torch
    torch.cuda.set_device(rank)
    import bbq_b200
    from oracle import oracle as O
    cfg = {"queryBits": 4, "indexBits": 1, "quantizer": {"similarityFunction": sim, "lambda": 0.1, "iters": 5}}
    fmt = bbq_b200.createBinaryQuantizationFormat(cfg, device=rank)
    rows, qs = gaussian(n, dim, 900), gaussian(nq, dim, 901)
    cen = np.zeros(dim, np.float32)
    r0, r1 = bbq_b200.shard_bounds(n, world, rank)
    shard = fmt.quantizeVectors(rows[r0:r1], centroid=cen)["quantizedVectors"]
    assert bbq_b200._native.load().bbq_index_set_base(shard._h, r0) == 0
    if rank == 0:
        cid = fmt.commUniqueId()
        for _ in range(world - 1):
            ids.put(cid)
    else:
        cid = ids.get(timeout=120)
    fmt.commInit(cid, rank, world)
    info = fmt.commInfo()
    assert info["world"] == world and info["rank"] == rank and info["nccl_version"] > 20000
    hq = torch.from_numpy(qs).pin_memory()
    hi = torch.empty((nq, k), dtype=torch.int32).pin_memory()
    hs = torch.empty((nq, k), dtype=torch.float32).pin_memory()
    results = []
    for batch in (nq, 3):                       # tensor-core batch, then a 3-query popcount batch
        cnt = fmt.searchShardedHost(hq.data_ptr(), batch, shard, k, hi.data_ptr(), hs.data_ptr())
        results.append((cnt, hi.numpy()[:batch].copy(), hs.numpy()[:batch].copy()))
    # k larger than any single shard but not than the corpus: count = min(k, rows over all ranks).  (The SAME k on
    # every rank: the call is collective.)
    kbig = -(-n // world) + 5
    bi = torch.empty((2, kbig), dtype=torch.int32).pin_memory()
    bs = torch.empty((2, kbig), dtype=torch.float32).pin_memory()
    cnt_big = fmt.searchShardedHost(hq.data_ptr(), 2, shard, kbig, bi.data_ptr(), bs.data_ptr()) if kbig <= 4096 else None
    ok = True
    full = O.quantize_vectors(rows, sim=sim, want_unpacked=False, centroid=cen)
    for cnt, gi, gs in results:
        ok &= cnt == k
        for qi in range(gi.shape[0]):
            wi, ws = O.search_nearest_neighbors(qs[qi], full, k, mode="canonical")
            ok &= gi[qi].tolist() == wi.tolist() and gs[qi].view(np.uint32).tolist() == ws.view(np.uint32).tolist()
    if cnt_big is not None:
        ok &= cnt_big == min(kbig, n)
        wi, ws = O.search_nearest_neighbors(qs[1], full, kbig, mode="canonical")
        ok &= bi.numpy()[1, :cnt_big].tolist() == wi.tolist()
    fmt.commDestroy()
    out.put((rank, bool(ok)))


@pytest.mark.parametrize("world", [2])
@pytest.mark.parametrize("sim,n,dim,nq,k", [("COSINE", 50_000, 256, 96, 10), ("EUCLIDEAN", 3001, 128, 40, 25)])
def test_library_sharded_search_equals_unsharded(world, sim, n, dim, nq, k):
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    ctx = mp.get_context("spawn")
    ids, out = ctx.Queue(), ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, ids, out, n, dim, nq, k, sim)) for r in range(world)]
    for p in procs:
        p.start()
    try:
        for p in procs:
            p.join(240)
            assert p.exitcode == 0, "a rank failed or hung"
    finally:
        for p in procs:      # never leave a rank spinning in a collective on the GPU behind a failed test
            if p.is_alive():
                p.kill()
    got = dict(out.get(timeout=5) for _ in range(world))
    assert got == {r: True for r in range(world)}


def test_sharded_entry_without_communicator_is_the_plain_search():
    import bbq_b200
    bbq_b200.build_library()
    fmt = bbq_b200.createBinaryQuantizationFormat()
    rows, qs = gaussian(2000, 64, 5), gaussian(4, 64, 6)
    qv = fmt.quantizeVectors(rows)["quantizedVectors"]
    a = fmt.searchBatch(qs, qv, 7)
    idx = np.empty((4, 7), np.int32)
    sc = np.empty((4, 7), np.float32)
    q = np.ascontiguousarray(qs)
    cnt = fmt.searchShardedHost(q.ctypes.data, 4, qv, 7, idx.ctypes.data, sc.ctypes.data)
    assert cnt == 7 and np.array_equal(idx, a[0]) and np.array_equal(sc.view(np.uint32), a[1].view(np.uint32))
    assert fmt.commInfo()["world"] == 1
