"""Diagnostic for the sharded path under torchrun (2 ranks): stage markers + a Python stack dump if anything blocks."""
import faulthandler, os, sys, time
faulthandler.dump_traceback_later(45, exit=True)
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import torch.distributed as dist
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
def say(*a):
    print(f"[r{rank} {time.time() % 1000:.2f}]", *a, flush=True)
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
import bbq_b200
say("pg up")
fmt = bbq_b200.createBinaryQuantizationFormat({"queryBits": 4, "indexBits": 1, "quantizer": {"similarityFunction": "COSINE", "lambda": 0.1, "iters": 5}}, device=lr)
n, dim, nq, k = 60000, 256, 96, 10
rows = np.random.default_rng(1).standard_normal((n, dim), dtype=np.float32)
qs = np.random.default_rng(2).standard_normal((nq, dim), dtype=np.float32)
r0, r1 = bbq_b200.shard_bounds(n, world, rank)
shard = fmt.quantizeVectors(rows[r0:r1], centroid=np.zeros(dim, np.float32))["quantizedVectors"]
say("shard built", r0, r1)
s = bbq_b200.ShardedSearcher(fmt, shard, r0, rank, world)
say("comm up", fmt.commInfo())
hq = torch.from_numpy(qs).pin_memory()
dq = hq.cuda()
torch.cuda.synchronize()
for it in range(3):
    oi, os_ = s.search_device(dq, k)
    s.stream.synchronize()
    say("search_device", it, oi[0, :3].tolist())
dist.barrier(); torch.cuda.synchronize()
say("barrier ok")
for it in range(3):
    hi, hs = s.search(hq, k)
    say("search host", it, hi[0, :3].tolist())
dist.barrier()
whole = fmt.quantizeVectors(rows, centroid=np.zeros(dim, np.float32))["quantizedVectors"]
wi, ws = fmt.searchBatch(qs, whole, k)
say("equal to unsharded:", bool(np.array_equal(wi, hi.numpy()) and np.array_equal(ws.view(np.uint32), hs.numpy().view(np.uint32))))
dist.destroy_process_group()
say("done")
