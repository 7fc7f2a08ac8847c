"""Row-wise sharded search over the GPUs of one box: one process per GPU, each rank owns a
contiguous row shard (global id = shard base + local row); a query batch is searched on every shard, the
per-shard top-k lists are exchanged with ONE NCCL all_gather over NVLink (torch.distributed is plumbing
only) and merged by the deterministic device merge (bbq_merge_topk_device), so the result is independent
of the number of shards.  The reference has no counterpart (it is single-process); the merge rule is its
MinHeap contract made canonical (src/binaryQuantizationFormat.ts:383-411)."""
from __future__ import annotations

import numpy as np

from .format import BinarizedByteVectorValues, BinaryQuantizationFormat


def shard_bounds(n_total: int, world: int, rank: int):
    """Contiguous shards: rank r owns [r*ceil(n/world), min(n, (r+1)*ceil(n/world)))."""
    per = -(-n_total // world)
    return min(n_total, rank * per), min(n_total, (rank + 1) * per)


class ShardedSearcher:
    def __init__(self, fmt: BinaryQuantizationFormat, shard: BinarizedByteVectorValues, base: int, rank: int = 0,
                 world: int = 1, group=None):
        import torch
        from .. import _native
        self.torch = torch
        self.fmt, self.shard, self.rank, self.world, self.group = fmt, shard, rank, world, group
        st = _native.load().bbq_index_set_base(shard._h, int(base))
        if st != 0:
            raise RuntimeError(_native.load().bbq_last_error().decode())
        self.stream = torch.cuda.Stream()
        self._bufs = {}

    def _buffers(self, nq, k, dim):
        key = (nq, k, dim)
        if key not in self._bufs:
            t, dev = self.torch, "cuda"
            self._bufs = {key: dict(
                dq=t.empty((nq, dim), dtype=t.float32, device=dev),
                loc_idx=t.empty((nq, k), dtype=t.int32, device=dev),
                loc_sc=t.empty((nq, k), dtype=t.float32, device=dev),
                all_idx=t.empty((self.world, nq, k), dtype=t.int32, device=dev),
                all_sc=t.empty((self.world, nq, k), dtype=t.float32, device=dev),
                out_idx=t.empty((nq, k), dtype=t.int32, device=dev),
                out_sc=t.empty((nq, k), dtype=t.float32, device=dev),
                h_idx=t.empty((nq, k), dtype=t.int32, pin_memory=True),
                h_sc=t.empty((nq, k), dtype=t.float32, pin_memory=True))}
        return self._bufs[key]

    def search_device(self, dq, k):
        """dq: [nq, dim] f32 CUDA tensor (replicated on every rank).  Returns (idx, score) CUDA tensors [nq, k];
        enqueued on self.stream (the caller synchronises)."""
        t = self.torch
        nq, dim = dq.shape
        b = self._buffers(nq, k, dim)
        s = self.stream.cuda_stream
        with t.cuda.stream(self.stream):
            if self.world == 1:
                self.fmt.searchDevice(dq.data_ptr(), nq, self.shard, k, b["out_idx"].data_ptr(), b["out_sc"].data_ptr(), s)
            else:
                self.fmt.searchDevice(dq.data_ptr(), nq, self.shard, k, b["loc_idx"].data_ptr(), b["loc_sc"].data_ptr(), s)
                t.distributed.all_gather_into_tensor(b["all_idx"], b["loc_idx"], group=self.group)
                t.distributed.all_gather_into_tensor(b["all_sc"], b["loc_sc"], group=self.group)
                self.fmt.mergeTopKDevice(b["all_idx"].data_ptr(), b["all_sc"].data_ptr(), self.world, nq, k,
                                         b["out_idx"].data_ptr(), b["out_sc"].data_ptr(), s)
        return b["out_idx"], b["out_sc"]

    def search(self, h_queries, k):
        """End-to-end: pinned host queries -> device, sharded search + merge, results back to pinned host."""
        t = self.torch
        nq, dim = h_queries.shape
        b = self._buffers(nq, k, dim)
        with t.cuda.stream(self.stream):
            b["dq"].copy_(h_queries, non_blocking=True)
        oi, os_ = self.search_device(b["dq"], k)
        with t.cuda.stream(self.stream):
            b["h_idx"].copy_(oi, non_blocking=True)
            b["h_sc"].copy_(os_, non_blocking=True)
        self.stream.synchronize()
        return b["h_idx"], b["h_sc"]


def merge_host(idx_lists, score_lists, k):
    """Pure-numpy statement of the merge rule (score desc, id asc, empty = id -1) for the gloo CPU tests of the
    multi-rank plumbing; NOT used by the product path."""
    idx = np.concatenate(idx_lists, axis=1)
    sc = np.concatenate(score_lists, axis=1)
    out_i = np.full((idx.shape[0], k), -1, np.int32)
    out_s = np.full((idx.shape[0], k), -np.inf, np.float32)
    for q in range(idx.shape[0]):
        valid = idx[q] >= 0
        i, s = idx[q][valid], sc[q][valid]
        nan = np.isnan(s)
        order = np.lexsort((i, -np.where(nan, -np.inf, s + 0.0), nan))
        order = order[:k]
        out_i[q, :len(order)], out_s[q, :len(order)] = i[order], s[order]
    return out_i, out_s
