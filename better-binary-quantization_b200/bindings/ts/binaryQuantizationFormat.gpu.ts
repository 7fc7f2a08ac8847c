/**
 * Drop-in replacement for src/binaryQuantizationFormat.ts of leolee9086/Better-Binary-Quantization:
 * same class name, constructor, method names, argument meaning, return shapes and error messages
 * (reference lines cited per member), but quantizeVectors / searchNearestNeighbors run on a B200 through the
 * N-API addon (bindings/napi) and libbbq_b200.so.  There is no JS/WASM fallback once this file is selected.
 *
 * Install: replace the reference's `src/binaryQuantizationFormat.ts` with this file (src/index.ts:20-139 stays
 * byte-for-byte unchanged) and ship `bbq_b200.node` + `libbbq_b200.so` next to dist/.
 */
import type { BinaryQuantizationConfig, BinarizedByteVectorValues, QuantizationResult } from './types';
import { VectorSimilarityFunction } from './types';
import { QUERY_BITS, INDEX_BITS } from './constants';
import { toReferenceError } from './errors';

// eslint-disable-next-line @typescript-eslint/no-var-requires
const addon = require('./bbq_b200.node') as {
  create(queryBits: number, indexBits: number, sim: number, lambda: number, iters: number, device: number): unknown;
  build(ctx: unknown, rows: Float32Array, n: number, dim: number, centroid: Float32Array | null): unknown;
  info(index: unknown): { size: number; dimension: number; centroid: Float32Array; centroidDP: number };
  search(index: unknown, queries: Float32Array, nq: number, k: number):
    { indices: Int32Array; scores: Float32Array; count: number };
  rows(index: unknown, first: number, count: number): { packed: Uint8Array; corrections: Float64Array };
};

const SIM_CODE: Record<string, number> = { EUCLIDEAN: 0, COSINE: 1, MAXIMUM_INNER_PRODUCT: 2 };

/** src/binaryQuantizationFormat.ts:24-126 — the index now lives in HBM; this object is a handle. */
class DeviceBinarizedByteVectorValues implements BinarizedByteVectorValues {
  private readonly meta: { size: number; dimension: number; centroid: Float32Array; centroidDP: number };
  constructor(public readonly handle: unknown) { this.meta = addon.info(handle); }
  dimension(): number { return this.meta.dimension; }                       // :45
  size(): number { return this.meta.size; }                                 // :49
  getCentroid(): Float32Array { return this.meta.centroid; }                // :123
  getCentroidDP(queryVector?: Float32Array): number {                       // :113-121
    if (!queryVector) return this.meta.centroidDP;
    let s = 0;
    for (let i = 0; i < queryVector.length; i++) s += queryVector[i]! * this.meta.centroid[i]!;
    return s;
  }
  vectorValue(ord: number): Uint8Array {                                    // :53 (lazy device->host copy)
    if (ord < 0 || ord >= this.meta.size) throw new Error(`向量索引 ${ord} 不存在`);
    return addon.rows(this.handle, ord, 1).packed;
  }
  getUnpackedVector(ord: number): Uint8Array {                              // :68
    const packed = this.vectorValue(ord);
    const out = new Uint8Array(this.meta.dimension);
    for (let i = 0; i < out.length; i++) out[i] = (packed[i >> 3]! >> (7 - (i & 7))) & 1;
    return out;
  }
  getCorrectiveTerms(ord: number): QuantizationResult {                     // :105
    if (ord < 0 || ord >= this.meta.size) throw new Error(`修正项索引 ${ord} 不存在`);
    const c = addon.rows(this.handle, ord, 1).corrections;
    return { lowerInterval: c[0]!, upperInterval: c[1]!, additionalCorrection: c[2]!, quantizedComponentSum: c[3]! };
  }
  clearUnpackedVectorCache(): void { /* nothing is cached on the host */ }
}

export class BinaryQuantizationFormat {
  private readonly config: BinaryQuantizationConfig;
  private readonly ctx: unknown;

  constructor(config: BinaryQuantizationConfig) {                            // :141-158
    this.config = { queryBits: QUERY_BITS, indexBits: INDEX_BITS, ...config };
    const q = config.quantizer;
    try {
      this.ctx = addon.create(this.config.queryBits!, this.config.indexBits!,
        SIM_CODE[q.similarityFunction ?? VectorSimilarityFunction.EUCLIDEAN]!, q.lambda ?? 0.1, q.iters ?? 5,
        Number(process.env.BBQ_DEVICE ?? -1));
    } catch (e) { throw toReferenceError(e, 'build'); }
  }

  public quantizeVectors(vectors: Float32Array[]) {                          // :165-263
    if (vectors.length === 0) throw new Error('向量集合不能为空');
    const dim = vectors[0]!.length;
    const flat = new Float32Array(vectors.length * dim);
    for (let i = 0; i < vectors.length; i++) {
      const v = vectors[i];
      if (!v) throw new Error(`向量 ${i} 不能为空`);
      if (v.length !== dim) throw new Error(`向量 ${i} 维度 ${v.length} 与第一个向量维度 ${dim} 不匹配`);
      flat.set(v, i * dim);
    }
    try {
      const handle = addon.build(this.ctx, flat, vectors.length, dim, null);
      return { quantizedVectors: new DeviceBinarizedByteVectorValues(handle) as BinarizedByteVectorValues,
               queryQuantizer: this };
    } catch (e) { throw toReferenceError(e, 'build'); }
  }

  public searchNearestNeighbors(queryVector: Float32Array, targetVectors: BinarizedByteVectorValues, k: number):
      Array<{ index: number; score: number; originalScore?: number }> {      // :308-412
    if (!queryVector) throw new Error('查询向量不能为空');
    if (!targetVectors) throw new Error('目标向量集合不能为空');
    if (k < 0) throw new Error('k值不能为负数');
    if (queryVector.length !== targetVectors.dimension()) throw new Error('查询向量维度与目标向量维度不匹配');
    if (k === 0) return [];
    try {
      const r = addon.search((targetVectors as DeviceBinarizedByteVectorValues).handle, queryVector, 1, k);
      const out = new Array(r.count);
      for (let i = 0; i < r.count; i++) out[i] = { index: r.indices[i]!, score: r.scores[i]! };
      return out;
    } catch (e) { throw toReferenceError(e, 'search'); }
  }

  /** Additive: a batch of queries in one call; row i equals searchNearestNeighbors(queries[i], ...). */
  public searchBatch(queries: Float32Array[], targetVectors: BinarizedByteVectorValues, k: number) {
    const dim = targetVectors.dimension();
    const flat = new Float32Array(queries.length * dim);
    queries.forEach((q, i) => {
      if (q.length !== dim) throw new Error('查询向量维度与目标向量维度不匹配');
      flat.set(q, i * dim);
    });
    try {
      const r = addon.search((targetVectors as DeviceBinarizedByteVectorValues).handle, flat, queries.length, k);
      return queries.map((_, i) => Array.from({ length: r.count }, (_u, j) =>
        ({ index: r.indices[i * k + j]!, score: r.scores[i * k + j]! })));
    } catch (e) { throw toReferenceError(e, 'search'); }
  }

  /** getOversampledTopKWithSort / ...WithHeap (src/topKSelector.ts:29-114) on the device: quantised top k*factor,
   *  exact cosine against the attached original rows, best k by true score. */
  public attachOriginalVectors(targetVectors: BinarizedByteVectorValues, vectors: Float32Array[]): void {
    const dim = targetVectors.dimension();
    const flat = new Float32Array(vectors.length * dim);
    vectors.forEach((v, i) => flat.set(v, i * dim));
    try { addon.attachRows((targetVectors as DeviceBinarizedByteVectorValues).handle, flat); }
    catch (e) { throw toReferenceError(e, 'build'); }
  }
  public searchOversampled(queryVector: Float32Array, targetVectors: BinarizedByteVectorValues, k: number,
      oversampleFactor: number): Array<{ index: number; quantizedScore: number; trueScore: number }> {
    try {
      const r = addon.searchRerank((targetVectors as DeviceBinarizedByteVectorValues).handle, queryVector, 1, k,
        oversampleFactor);
      return Array.from({ length: r.count }, (_u, i) =>
        ({ index: r.indices[i]!, quantizedScore: r.quantizedScores[i]!, trueScore: r.trueScores[i]! }));
    } catch (e) { throw toReferenceError(e, 'search'); }
  }

  /** On-disk form of serializeVectorData / deserializeVectorData (:483-560): `${prefix}.veb` + `${prefix}.vemb`
   *  (FILE_EXTENSIONS, src/constants.ts:52-57), streamed between the files and device memory natively. */
  public saveIndex(targetVectors: BinarizedByteVectorValues, prefix: string): void {
    try { addon.saveIndex((targetVectors as DeviceBinarizedByteVectorValues).handle, `${prefix}.veb`, `${prefix}.vemb`); }
    catch (e) { throw toReferenceError(e, 'build'); }
  }
  public loadIndex(prefix: string): BinarizedByteVectorValues {
    try {
      return new DeviceBinarizedByteVectorValues(addon.loadIndex(this.ctx, `${prefix}.veb`, `${prefix}.vemb`)) as
        BinarizedByteVectorValues;
    } catch (e) { throw toReferenceError(e, 'build'); }
  }

  public getConfig(): BinaryQuantizationConfig { return this.config; }        // :583
}
