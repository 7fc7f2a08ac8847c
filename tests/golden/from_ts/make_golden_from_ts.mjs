// Runs the UNMODIFIED reference on the seeded inputs exported by export_inputs.py and writes <case>.ts.json.
//   node --experimental-strip-types --import ./register.mjs make_golden_from_ts.mjs <reference root> <dir with *.in.bin>
// Never executed in this repository's environment (no JS runtime exists: see README.md).
import { readdirSync, readFileSync, writeFileSync } from 'node:fs';
import { join } from 'node:path';
import { pathToFileURL } from 'node:url';

const [refRoot, dir] = process.argv.slice(2);
const ref = await import(pathToFileURL(join(refRoot, 'src', 'index.ts')).href);
const SIMS = ['EUCLIDEAN', 'COSINE', 'MAXIMUM_INNER_PRODUCT'];
const f64bits = (x) => { const b = new BigUint64Array(new Float64Array([x]).buffer)[0]; return b.toString(16).padStart(16, '0'); };
const f32bits = (x) => new Uint32Array(new Float32Array([x]).buffer)[0];

for (const file of readdirSync(dir).filter((f) => f.endsWith('.in.bin'))) {
  const buf = readFileSync(join(dir, file));
  const dv = new DataView(buf.buffer, buf.byteOffset, buf.byteLength);
  const [n, dim, nq, k, queryBits, sim, iters] = [0, 4, 8, 12, 16, 20, 24].map((o) => dv.getInt32(o, true));
  const lambda = dv.getFloat64(28, true);
  const f32 = (off, len) => new Float32Array(buf.buffer.slice(buf.byteOffset + off, buf.byteOffset + off + 4 * len));
  const base = Array.from({ length: n }, (_, i) => f32(36 + 4 * i * dim, dim));
  const queries = Array.from({ length: nq }, (_, i) => f32(36 + 4 * (n + i) * dim, dim));
  const format = ref.createBinaryQuantizationFormat({
    queryBits, indexBits: 1, quantizer: { similarityFunction: SIMS[sim], lambda, iters } });
  const { quantizedVectors } = format.quantizeVectors(base);                       // src/binaryQuantizationFormat.ts:165-263
  const out = { n, dim, nq, k, queryBits, sim: SIMS[sim], lambda, iters,
    centroid_bits: Array.from(quantizedVectors.getCentroid(), f32bits),
    corrections_bits: [], packed_sum: 0, queries: [] };
  for (let i = 0; i < n; i++) {
    const c = quantizedVectors.getCorrectiveTerms(i);
    out.corrections_bits.push([c.lowerInterval, c.upperInterval, c.additionalCorrection, c.quantizedComponentSum].map(f64bits));
    for (const b of quantizedVectors.vectorValue(i)) out.packed_sum += b;
  }
  for (const q of queries) {
    const res = format.searchNearestNeighbors(q, quantizedVectors, k);            // :308-412
    const all = format.searchNearestNeighbors(q, quantizedVectors, n);            // every row's f32 score, via k = n
    const byIndex = new Array(n);
    for (const r of all) byIndex[r.index] = f32bits(r.score);
    out.queries.push({ top_index: res.map((r) => r.index), top_score_bits: res.map((r) => f32bits(r.score)),
      all_score_bits: byIndex });
  }
  writeFileSync(join(dir, file.replace('.in.bin', '.ts.json')), JSON.stringify(out));
  console.log('wrote', file.replace('.in.bin', '.ts.json'));
}
