"""better-binary-quantization_b200 — B200-native (sm_100a) brute-force quantized search path of
leolee9086/Better-Binary-Quantization, behind the reference's own public API (src/index.ts:20-139).

Only the hot path lives here: csrc/ (CUDA kernels + the C ABI of include/bbq_b200.h), host/ (the
reference-facing operator interface) and bindings/ (the N-API shim + TypeScript glue a Node host uses).
"""
from . import _native
from ._native import build_library
from .host.sharded import ShardedSearcher, merge_host, shard_bounds
from .host.format import (BbqError, BinarizedByteVectorValues, BinaryQuantizationFormat,
                          VectorSimilarityFunction)

VERSION = "1.0.0"  # src/index.ts:139

# src/index.ts:47-55
DEFAULT_CONFIG = {"queryBits": 4, "indexBits": 1,
                  "quantizer": {"similarityFunction": VectorSimilarityFunction.COSINE, "lambda": 0.1, "iters": 5}}


def createBinaryQuantizationFormat(config=None, device: int = -1) -> BinaryQuantizationFormat:
    """src/index.ts:62-64"""
    return BinaryQuantizationFormat(config if config is not None else DEFAULT_CONFIG, device=device)


def quickQuantize(vectors, similarityFunction=VectorSimilarityFunction.COSINE):
    """src/index.ts:72-85"""
    fmt = BinaryQuantizationFormat({"quantizer": {"similarityFunction": similarityFunction, "lambda": 0.1, "iters": 5}})
    return fmt.quantizeVectors(vectors)


def quickSearch(queryVector, targetVectors, k, similarityFunction=VectorSimilarityFunction.COSINE):
    """src/index.ts:95-111 — like the reference, re-quantises the corpus on every call."""
    fmt = BinaryQuantizationFormat({"quantizer": {"similarityFunction": similarityFunction, "lambda": 0.1, "iters": 5}})
    qv = fmt.quantizeVectors(targetVectors)["quantizedVectors"]
    return fmt.searchNearestNeighbors(queryVector, qv, k)


__all__ = ["BinaryQuantizationFormat", "BinarizedByteVectorValues", "VectorSimilarityFunction", "BbqError",
           "createBinaryQuantizationFormat", "quickQuantize", "quickSearch", "DEFAULT_CONFIG", "VERSION",
           "build_library", "ShardedSearcher", "shard_bounds", "merge_host"]
