// bbq_oracle.cpp — CPU ORACLE (TEST INFRASTRUCTURE ONLY, NOT PRODUCT CODE).
//
// A numerical restatement, in IEEE binary64/binary32 with JavaScript semantics,
// of the brute-force quantized search path of leolee9086/Better-Binary-Quantization
// (TypeScript).  It exists so the CUDA path can be checked for parity.  Only
// tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
// legs may load it; nothing under better-binary-quantization_b200/ links or calls it.
//
// PARITY STATUS: PINNED BY FIXTURES GENERATED FROM THE REFERENCE'S OWN SOURCE TEXT (round 2), with one caveat.
// No JavaScript / TypeScript runtime exists in the build container or on the GPU box (profiles/r02_js_runtime_probe.txt)
// and the reference's own tests hold no golden score vectors.  tests/golden/from_ts/tsinterp.py — a TypeScript-subset
// interpreter written for this purpose — executes the UNMODIFIED /root/reference/src/index.ts (every module on the
// path) on seeded inputs; tests/golden/from_ts/*.ts.json are its outputs (each records the SHA-256 of the reference
// files), and tests/test_golden_from_ts.py requires this oracle to reproduce them BIT FOR BIT: centroid, every row's
// code and correctives (f64 bit patterns), every row's f32 score, the heap-ordered top-k lists, quantizeQueryVector,
// the statistics of computeQuantizationAccuracy, getOversampledTopKWithHeap — all three similarity functions, 4-bit and
// 1-bit queries, dim % 8 != 0, iters = 20.  Caveat: the executing engine is that interpreter, not V8; its ECMAScript
// number semantics are tested separately (tests/test_tsinterp.py).  Two independent derivations from the same text — a
// hand restatement in C++ and a mechanical execution — agree on every bit.
// Also pinned by (tests/test_oracle_kat.py):
//   * the four exact known-answer tests in rust-wasm/src/*.rs,
//   * the deterministic sin/cos recall fixtures + thresholds of tests/recall*.ts,
//   * the behavioural properties the reference tests assert (k=0, k>N, ordering...).
// indexBits = 2: the index BUILD is pinned the same way (tests/golden/from_ts/index_bits_2.behaviour.json); the SEARCH
// extension below has no reference behaviour to be pinned to — executed, the reference throws for queryBits = 8 and
// falls back to a per-vector formula for queryBits = 4 that this extension deliberately does not follow.
//
// Build: g++ -O2 -std=c++17 -ffp-contract=off -fno-fast-math (see oracle/Makefile).
// -ffp-contract=off is REQUIRED: JS never fuses a*b+c.
//
// JS semantics used throughout:
//   number == double; Float32Array store == (float) round-to-nearest-even;
//   Math.round(x) == floor(x) + (x - floor(x) >= 0.5); Math.min/max propagate NaN;
//   sums strictly in index order, evaluated left-to-right.
//
// Every function cites the reference file:line (relative to the reference root) it follows.

#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <vector>
#include <algorithm>

namespace {

constexpr int SIM_EUCLIDEAN = 0;
constexpr int SIM_COSINE = 1;
constexpr int SIM_MIP = 2;

// src/constants.ts:38-47 (MINIMUM_MSE_GRID, second column; first is its negation)
const double kGrid[8] = {0.798, 1.493, 2.051, 2.514, 2.916, 3.278, 3.611, 3.922};
// src/constants.ts:20
const double kFourBitScale = 1.0 / 15.0;

inline double js_min(double a, double b) {  // Math.min: NaN-propagating, and -0 < +0
  if (std::isnan(a) || std::isnan(b)) return std::numeric_limits<double>::quiet_NaN();
  if (a == 0 && b == 0) return std::signbit(a) ? a : b;
  return a < b ? a : b;
}
inline double js_max(double a, double b) {  // Math.max: NaN-propagating, and +0 > -0
  if (std::isnan(a) || std::isnan(b)) return std::numeric_limits<double>::quiet_NaN();
  if (a == 0 && b == 0) return std::signbit(a) ? b : a;
  return a > b ? a : b;
}
// src/utils.ts:79-81
inline double js_clamp(double x, double lo, double hi) { return js_min(js_max(x, lo), hi); }
// ECMAScript Math.round: nearest integer, ties toward +infinity.
inline double js_round(double x) {
  if (!(std::fabs(x) < 4503599627370496.0)) return x;  // NaN, Inf, |x| >= 2^52: already integral
  double f = std::floor(x);
  return (x - f >= 0.5) ? f + 1.0 : f;
}

// src/optimizedScalarQuantizer.ts:373-407 (computeLoss)
double osq_loss(const float* w, int d, double a, double b, int points, double nrm, double lambda) {
  const double step = (b - a) / (double)(points - 1);
  const double stepInv = 1.0 / step;
  double xe = 0.0, e = 0.0;
  for (int i = 0; i < d; i++) {
    const double xi = (double)w[i];
    const double clamped = js_clamp(xi, a, b);
    const double k = js_round((clamped - a) * stepInv);
    const double xiq = a + step * k;
    xe += xi * (xi - xiq);
    e += (xi - xiq) * (xi - xiq);
  }
  return (1.0 - lambda) * xe * xe / nrm + lambda * e;
}

// src/optimizedScalarQuantizer.ts:280-353 (optimizeIntervals)
void osq_optimize(const float* w, int d, double& a_io, double& b_io, int points, double nrm,
                  double lambda, int iters) {
  double initialLoss = osq_loss(w, d, a_io, b_io, points, nrm, lambda);
  const double scale = (1.0 - lambda) / nrm;
  if (!std::isfinite(scale)) return;
  for (int iter = 0; iter < iters; iter++) {
    const double a = a_io, b = b_io;
    const double stepInv = (double)(points - 1) / (b - a);
    double daa = 0, dab = 0, dbb = 0, dax = 0, dbx = 0;
    for (int i = 0; i < d; i++) {
      const double xi = (double)w[i];
      const double clamped = js_clamp(xi, a, b);
      const double k = js_round((clamped - a) * stepInv);
      const double s = k / (double)(points - 1);
      daa += (1.0 - s) * (1.0 - s);
      dab += (1.0 - s) * s;
      dbb += s * s;
      dax += xi * (1.0 - s);
      dbx += xi * s;
    }
    const double m0 = scale * dax * dax + lambda * daa;
    const double m1 = scale * dax * dbx + lambda * dab;
    const double m2 = scale * dbx * dbx + lambda * dbb;
    const double det = m0 * m2 - m1 * m1;
    if (std::fabs(det) < 1e-12) return;  // src/utils.ts:150 isNearZero(det, MIN_DETERMINANT)
    const double aOpt = (m2 * dax - m1 * dbx) / det;
    const double bOpt = (m0 * dbx - m1 * dax) / det;
    if (std::fabs(a_io - aOpt) < 1e-8 && std::fabs(b_io - bOpt) < 1e-8) return;  // isNearEqual
    const double newLoss = osq_loss(w, d, aOpt, bOpt, points, nrm, lambda);
    if (newLoss > initialLoss) return;
    a_io = aOpt;
    b_io = bOpt;
    initialLoss = newLoss;
  }
}

// src/optimizedScalarQuantizer.ts:108-227 (scalarQuantize) + :245-265 (getInitialInterval)
// corr = {lowerInterval, upperInterval, additionalCorrection, quantizedComponentSum}
void osq_quantize(const float* v, const float* c, int d, int bits, int sim, double lambda, int iters,
                  uint8_t* codes, double corr[4]) {
  std::vector<float> w(d);
  double centroidDot = 0.0;
  if (sim != SIM_EUCLIDEAN) {
    for (int i = 0; i < d; i++) centroidDot += (double)v[i] * (double)c[i];
  }
  double mn = std::numeric_limits<double>::max();
  double mx = -std::numeric_limits<double>::max();
  for (int i = 0; i < d; i++) {
    const double centered = (double)v[i] - (double)c[i];
    w[i] = (float)centered;  // Float32Array store
    mn = js_min(mn, centered);
    mx = js_max(mx, centered);
  }
  // src/utils.ts:41-68 computeMean / computeStd, :25-34 computeL2Norm (returns the NORM)
  double sum = 0.0;
  for (int i = 0; i < d; i++) sum += (double)w[i];
  const double mean = sum / (double)d;
  double ss = 0.0;
  for (int i = 0; i < d; i++) {
    const double diff = (double)w[i] - mean;
    ss += diff * diff;
  }
  const double sd = std::sqrt(ss / (double)d);
  double n2 = 0.0;
  for (int i = 0; i < d; i++) n2 += (double)w[i] * (double)w[i];
  const double nrm = std::sqrt(n2);

  const double g = kGrid[bits - 1];
  double a = js_clamp(-g * sd + mean, mn, mx);
  double b = js_clamp(g * sd + mean, mn, mx);
  const int points = 1 << bits;
  osq_optimize(w.data(), d, a, b, points, nrm, lambda, iters);

  const int nSteps = points - 1;
  const double step = nSteps > 0 ? (b - a) / (double)nSteps : 0.0;
  const double stepInv = step > 0 ? 1.0 / step : 0.0;
  double qsum = 0.0;
  for (int i = 0; i < d; i++) {
    const double xi = (double)w[i];
    const double clamped = js_clamp(xi, a, b);
    if (bits == 1) {
      const double threshold = (a + b) / 2;
      const int q = clamped >= threshold ? 1 : 0;
      codes[i] = (uint8_t)q;
      qsum += (double)q;
    } else {
      const double assignment = js_round((clamped - a) * stepInv);
      const double stored = js_min(assignment, (double)nSteps);
      // Uint8Array store: ToUint8 (NaN -> 0, modulo 256)
      codes[i] = std::isnan(stored) ? 0 : (uint8_t)(((long long)stored) & 0xFF);
      qsum += assignment;
    }
  }
  corr[0] = a;
  corr[1] = b;
  corr[2] = (sim == SIM_EUCLIDEAN) ? nrm : centroidDot;
  corr[3] = qsum;
}

// src/vectorOperations.ts:11-34 (normalizeVector)
void normalize(const float* v, int d, float* out) {
  double n = 0.0;
  for (int i = 0; i < d; i++) n += (double)v[i] * (double)v[i];
  n = std::sqrt(n);
  if (n == 0) {
    for (int i = 0; i < d; i++) out[i] = 0.0f;
    return;
  }
  for (int i = 0; i < d; i++) out[i] = (float)((double)v[i] / n);
}

// One corrected score, src/batchDotProduct.ts:554-617 (query_bits != 1) and :478-541 (query_bits == 1).
// Returned as the f64 the reference computes; the caller rounds to f32 (binaryQuantizationFormat.ts:353,378).
double score_one(double dot, const double xc[4], const double qc[4], int d, double cdp, int sim,
                 int query_bits) {
  const double x1 = xc[3];
  const double ax = xc[0];
  const double lx = xc[1] - ax;
  const double ay = qc[0];
  const double ly = (query_bits == 1) ? (qc[1] - ay) : (qc[1] - ay) * kFourBitScale;
  const double y1 = qc[3];
  double score = ax * ay * (double)d + ay * lx * x1 + ax * ly * y1 + lx * ly * dot;
  if (sim == SIM_EUCLIDEAN) {
    const double e = qc[2] + xc[2] - 2 * score;
    return js_max(1 / (1 + e), 0);
  }
  if (query_bits == 1) {
    // `score += addq + addi - centroidDP` : right-hand side evaluated first
    score = score + (qc[2] + xc[2] - cdp);
    if (sim == SIM_COSINE) return js_max((1 + score) / 2, 0);
    return score < 0 ? 1 / (1 - score) : score + 1;
  }
  const double adj = score + qc[2] + xc[2] - cdp;
  if (sim == SIM_COSINE) return js_max((1 + adj) / 2, 0);
  return adj < 0 ? 1 / (1 - adj / kFourBitScale) : adj / kFourBitScale + 1;
}

// src/minHeap.ts:9-130 with compareFn (a,b) => a.score - b.score
struct HeapItem {
  double score;
  int32_t index;
};
struct MinHeap {
  std::vector<HeapItem> h;
  static double cmp(const HeapItem& a, const HeapItem& b) { return a.score - b.score; }
  void push(HeapItem it) {
    h.push_back(it);
    size_t index = h.size() - 1;
    while (index > 0) {
      size_t parent = (index - 1) / 2;
      if (cmp(h[index], h[parent]) >= 0) break;
      std::swap(h[index], h[parent]);
      index = parent;
    }
  }
  HeapItem pop() {
    HeapItem mn = h[0];
    HeapItem last = h.back();
    h.pop_back();
    if (!h.empty()) {
      h[0] = last;
      size_t index = 0;
      for (;;) {
        size_t smallest = index, l = 2 * index + 1, r = 2 * index + 2;
        if (l < h.size() && cmp(h[l], h[smallest]) < 0) smallest = l;
        if (r < h.size() && cmp(h[r], h[smallest]) < 0) smallest = r;
        if (smallest == index) break;
        std::swap(h[index], h[smallest]);
        index = smallest;
      }
    }
    return mn;
  }
};

}  // namespace

extern "C" {

// src/vectorOperations.ts:11-34
void bbqo_normalize(const float* v, int d, float* out) { normalize(v, d, out); }

// src/vectorOperations.ts:126-163 (computeCentroid): f32 in-place accumulation over vectors, then /N in f32
void bbqo_centroid(const float* rows, int64_t n, int d, float* c) {
  for (int i = 0; i < d; i++) c[i] = rows[i];
  for (int64_t j = 1; j < n; j++) {
    const float* r = rows + j * (int64_t)d;
    for (int i = 0; i < d; i++) c[i] = (float)((double)c[i] + (double)r[i]);
  }
  for (int i = 0; i < d; i++) c[i] = (float)((double)c[i] / (double)n);
}

// src/optimizedScalarQuantizer.ts:108-227
void bbqo_osq(const float* v, const float* c, int d, int bits, int sim, double lambda, int iters,
              uint8_t* codes, double* corr4) {
  osq_quantize(v, c, d, bits, sim, lambda, iters, codes, corr4);
}

// src/optimizedScalarQuantizer.ts:420-446 (packAsBinary): MSB-first, tail bits zero
void bbqo_pack_binary(const uint8_t* codes, int d, uint8_t* packed) {
  const int p = (d + 7) / 8;
  for (int j = 0; j < p; j++) {
    unsigned r = 0;
    for (int t = 0; t < 8; t++) {
      const int i = 8 * j + t;
      if (i < d) r |= (unsigned)(codes[i] & 1) << (7 - t);
    }
    packed[j] = (uint8_t)r;
  }
}

// src/binaryQuantizationFormat.ts:165-263 (quantizeVectors), index_bits == 1 -> packed rows of ceil(d/8) bytes,
// otherwise rows of d unpacked codes.  `centroid_in` non-null overrides the computed centroid (bench-scale use).
// `unpacked` (n*d) may be null.  corr is n*4 doubles.
void bbqo_build_index(const float* rows, int64_t n, int d, int sim, int index_bits, double lambda,
                      int iters, const float* centroid_in, float* centroid_out, uint8_t* packed,
                      uint8_t* unpacked, double* corr) {
  std::vector<float> proc;
  const float* src = rows;
  if (sim == SIM_COSINE) {
    proc.resize((size_t)n * d);
    for (int64_t j = 0; j < n; j++) normalize(rows + j * (int64_t)d, d, proc.data() + j * (int64_t)d);
    src = proc.data();
  }
  if (centroid_in) std::memcpy(centroid_out, centroid_in, sizeof(float) * d);
  else bbqo_centroid(src, n, d, centroid_out);
  const int p = (index_bits == 1) ? (d + 7) / 8 : d;
  std::vector<uint8_t> codes(d);
  for (int64_t j = 0; j < n; j++) {
    osq_quantize(src + j * (int64_t)d, centroid_out, d, index_bits, sim, lambda, iters, codes.data(),
                 corr + 4 * j);
    if (index_bits == 1) bbqo_pack_binary(codes.data(), d, packed + j * (int64_t)p);
    else std::memcpy(packed + j * (int64_t)p, codes.data(), d);
    if (unpacked) std::memcpy(unpacked + j * (int64_t)d, codes.data(), d);
  }
}

// src/binaryQuantizationFormat.ts:337-347 + :271-299: COSINE normalises TWICE, then scalarQuantize(queryBits)
void bbqo_quantize_query(const float* q, const float* c, int d, int sim, int query_bits, double lambda,
                         int iters, uint8_t* codes, double* corr4) {
  std::vector<float> a(q, q + d), b(d);
  if (sim == SIM_COSINE) {
    normalize(a.data(), d, b.data());
    normalize(b.data(), d, a.data());
  }
  osq_quantize(a.data(), c, d, query_bits, sim, lambda, iters, codes, corr4);
}

// src/utils/computeBatchFourBitDotProductDirectPacked.ts:10-53: sum_d q[d] * bit_d(x), bit_d MSB-first
void bbqo_qcdist_packed(const uint8_t* qcodes, const uint8_t* packed, int64_t n, int d, int32_t* out) {
  const int p = (d + 7) / 8;
  for (int64_t v = 0; v < n; v++) {
    const uint8_t* row = packed + v * (int64_t)p;
    int32_t acc = 0;
    for (int i = 0; i < d; i++) acc += (int32_t)qcodes[i] * ((row[i >> 3] >> (7 - (i & 7))) & 1);
    out[v] = acc;
  }
}

// src/batchDotProduct.ts:22-49 + src/utils/bitcount.ts:7-15: sum over bytes popcount(qbyte & xbyte)
void bbqo_qcdist_1bit(const uint8_t* qpacked, const uint8_t* packed, int64_t n, int d, int32_t* out) {
  const int p = (d + 7) / 8;
  for (int64_t v = 0; v < n; v++) {
    const uint8_t* row = packed + v * (int64_t)p;
    int32_t acc = 0;
    for (int j = 0; j < p; j++) acc += __builtin_popcount((unsigned)(qpacked[j] & row[j]));
    out[v] = acc;
  }
}

// src/bitwiseDotProduct.ts:14-31 (computeQuantizedDotProduct on UNPACKED codes) — defines the integer semantics
int32_t bbqo_dot_unpacked(const uint8_t* q, const uint8_t* x, int d) {
  int32_t acc = 0;
  for (int i = 0; i < d; i++) acc += (int32_t)q[i] * (int32_t)x[i];
  return acc;
}

// src/binaryQuantizationFormat.ts:113-121 getCentroidDP(undefined) = centroid . centroid (f64, sequential)
double bbqo_centroid_dp(const float* c, int d) {
  double s = 0.0;
  for (int i = 0; i < d; i++) s += (double)c[i] * (double)c[i];
  return s;
}

// src/batchDotProduct.ts:478-541,554-617 then Float32Array store (binaryQuantizationFormat.ts:353,378)
void bbqo_scores(const int32_t* dots, const double* xcorr, int64_t n, const double* qcorr4, int d,
                 double cdp, int sim, int query_bits, float* out) {
  for (int64_t v = 0; v < n; v++)
    out[v] = (float)score_one((double)dots[v], xcorr + 4 * v, qcorr4, d, cdp, sim, query_bits);
}

// src/binaryQuantizedScorer.ts:35-40 scaleMaxInnerProductScore (single-vector form; KAT only)
double bbqo_scale_mip(double s) { return s < 0 ? 1 / (1 - s) : s + 1; }

// src/binaryQuantizationFormat.ts:383-411: MinHeap emulation, exactly as the reference selects.
// Returns count = min(k, n); results descending (pop-all then reverse).
int64_t bbqo_topk_heap(const float* scores, int64_t n, int64_t k, int32_t* out_idx, float* out_score) {
  if (k <= 0) return 0;
  const int64_t k2 = k < n ? k : n;
  MinHeap heap;
  for (int64_t i = 0; i < n; i++) {
    const double s = (double)scores[i];
    if ((int64_t)heap.h.size() < k2) heap.push({s, (int32_t)i});
    else if (s > heap.h[0].score) {
      heap.pop();
      heap.push({s, (int32_t)i});
    }
  }
  std::vector<HeapItem> res;
  while (!heap.h.empty()) res.push_back(heap.pop());
  std::reverse(res.begin(), res.end());
  for (size_t i = 0; i < res.size(); i++) {
    out_idx[i] = res[i].index;
    out_score[i] = (float)res[i].score;
  }
  return (int64_t)res.size();
}

// Canonical contract of the new build (BASELINE.json north_star): the min(k,n) best under
// (f32 score descending, index ascending).  Equals the heap's SET whenever no exact f32 tie
// straddles the k-th place.  NaN scores rank last (the reference heap never admits them once full).
int64_t bbqo_topk_canonical(const float* scores, int64_t n, int64_t k, int32_t* out_idx,
                            float* out_score) {
  if (k <= 0) return 0;
  const int64_t k2 = k < n ? k : n;
  std::vector<int32_t> idx(n);
  for (int64_t i = 0; i < n; i++) idx[i] = (int32_t)i;
  auto better = [&](int32_t a, int32_t b) {
    const float sa = scores[a], sb = scores[b];
    const bool na = std::isnan(sa), nb = std::isnan(sb);
    if (na != nb) return nb;  // non-NaN first
    if (!na && sa != sb) return sa > sb;
    return a < b;
  };
  std::partial_sort(idx.begin(), idx.begin() + k2, idx.end(), better);
  for (int64_t i = 0; i < k2; i++) {
    out_idx[i] = idx[i];
    out_score[i] = scores[idx[i]];
  }
  return k2;
}

// src/binaryQuantizationFormat.ts:308-412 (searchNearestNeighbors) over an index_bits==1 index.
// mode 0 = reference heap selection, 1 = canonical selection.  all_scores (n floats) may be null.
int64_t bbqo_search(const float* query, const float* centroid, const uint8_t* packed, const double* xcorr,
                    int64_t n, int d, int sim, int query_bits, double lambda, int iters, int64_t k,
                    int mode, int32_t* out_idx, float* out_score, float* all_scores, int32_t* all_dots) {
  if (k <= 0) return 0;
  std::vector<uint8_t> qcodes(d);
  double qcorr[4];
  bbqo_quantize_query(query, centroid, d, sim, query_bits, lambda, iters, qcodes.data(), qcorr);
  std::vector<int32_t> dots_local;
  int32_t* dots = all_dots;
  if (!dots) {
    dots_local.resize(n);
    dots = dots_local.data();
  }
  if (query_bits == 1) {
    std::vector<uint8_t> qp((d + 7) / 8);
    bbqo_pack_binary(qcodes.data(), d, qp.data());
    bbqo_qcdist_1bit(qp.data(), packed, n, d, dots);
  } else {
    bbqo_qcdist_packed(qcodes.data(), packed, n, d, dots);
  }
  std::vector<float> sc_local;
  float* sc = all_scores;
  if (!sc) {
    sc_local.resize(n);
    sc = sc_local.data();
  }
  bbqo_scores(dots, xcorr, n, qcorr, d, bbqo_centroid_dp(centroid, d), sim, query_bits, sc);
  return mode == 0 ? bbqo_topk_heap(sc, n, k, out_idx, out_score)
                   : bbqo_topk_canonical(sc, n, k, out_idx, out_score);
}

// src/vectorSimilarity.ts:75-102 (computeCosineSimilarity): three interleaved sequential f64 sums
double bbqo_cosine(const float* a, const float* b, int d) {
  double dot = 0, na = 0, nb = 0;
  for (int i = 0; i < d; i++) {
    dot += (double)a[i] * (double)b[i];
    na += (double)a[i] * (double)a[i];
    nb += (double)b[i] * (double)b[i];
  }
  if (na == 0 || nb == 0) return 0;
  return dot / (std::sqrt(na) * std::sqrt(nb));
}

// src/topKSelector.ts:29-78 (getOversampledTopKWithHeap) after the quantised search: heap on trueScore, pop all,
// then sort descending.  cand_idx/true_scores have m entries in quantised-rank order; returns min(k, m) positions.
int64_t bbqo_rerank_heap(const double* true_scores, int64_t m, int64_t k, int32_t* out_pos) {
  MinHeap heap;
  for (int64_t i = 0; i < m; i++) {
    const double s = true_scores[i];
    if ((int64_t)heap.h.size() < k) heap.push({s, (int32_t)i});
    else if (!heap.h.empty() && s > heap.h[0].score) {
      heap.pop();
      heap.push({s, (int32_t)i});
    }
  }
  std::vector<HeapItem> res;
  while (!heap.h.empty()) res.push_back(heap.pop());
  std::stable_sort(res.begin(), res.end(), [](const HeapItem& a, const HeapItem& b) { return a.score > b.score; });
  for (size_t i = 0; i < res.size(); i++) out_pos[i] = res[i].index;
  return (int64_t)res.size();
}

// Faster inner loop for the cpu_baseline leg only: same integers as bbqo_qcdist_packed (it is the
// bit-plane identity sum_i 2^i popc(qplane_i & x)), used when timing larger samples.  Validated
// against bbqo_qcdist_packed in tests/test_oracle_kat.py.
void bbqo_qcdist_packed_planes(const uint8_t* qcodes, int query_bits, const uint8_t* packed, int64_t n,
                               int d, int32_t* out) {
  const int p = (d + 7) / 8;
  std::vector<uint8_t> planes((size_t)query_bits * p, 0);
  for (int i = 0; i < d; i++)
    for (int b = 0; b < query_bits; b++)
      if ((qcodes[i] >> b) & 1) planes[(size_t)b * p + (i >> 3)] |= (uint8_t)(1u << (7 - (i & 7)));
  for (int64_t v = 0; v < n; v++) {
    const uint8_t* row = packed + v * (int64_t)p;
    int32_t acc = 0;
    for (int b = 0; b < query_bits; b++) {
      const uint8_t* pl = planes.data() + (size_t)b * p;
      int32_t c = 0;
      for (int j = 0; j < p; j++) c += __builtin_popcount((unsigned)(pl[j] & row[j]));
      acc += c << b;
    }
    out[v] = acc;
  }
}

// src/binaryQuantizationFormat.ts:271-299 ALONE (quantizeQueryVector called directly, e.g. by
// computeQuantizationAccuracy :448): COSINE normalises ONCE, then scalarQuantize(queryBits).
void bbqo_quantize_query_once(const float* q, const float* c, int d, int sim, int query_bits, double lambda,
                              int iters, uint8_t* codes, double* corr4) {
  std::vector<float> a(q, q + d), b(d);
  if (sim == SIM_COSINE) {
    normalize(a.data(), d, b.data());
    a = b;
  }
  osq_quantize(a.data(), c, d, query_bits, sim, lambda, iters, codes, corr4);
}

// The single-vector scorer: computeQuantizedScore, src/binaryQuantizedScorer.ts:69-98 ->
// computeOneBitSimilarityScore :112-160 (queryBits == 1) / computeFourBitSimilarityScore :174-217 (queryBits == 4).
// Unlike the batch path: f64 result, plain scaleMaxInnerProductScore for MIP, centroidDP supplied by the caller
// (getCentroidDP() = c.c for 1-bit :245; 0 for 4-bit when no original query is passed :290).
double bbqo_score_single(double dot, const double xc[4], const double qc[4], int d, double cdp, int sim,
                         int query_bits) {
  const double x1 = xc[3], ax = xc[0], lx = xc[1] - ax, ay = qc[0], y1 = qc[3];
  if (query_bits == 1) {
    const double ly = qc[1] - ay;
    double score = ax * ay * d + ay * lx * x1 + ax * ly * y1 + lx * ly * dot;
    if (sim == SIM_EUCLIDEAN) {
      score = qc[2] + xc[2] - 2 * score;
      return js_max(1 / (1 + score), 0.0);
    }
    score += qc[2] + xc[2] - cdp;
    return sim == SIM_COSINE ? js_max((1 + score) / 2, 0.0) : bbqo_scale_mip(score);
  }
  const double ly = (qc[1] - ay) * (1.0 / 15.0);
  const double score = ax * ay * d + ay * lx * x1 + ax * ly * y1 + lx * ly * dot;
  if (sim == SIM_EUCLIDEAN) {
    const double e = qc[2] + xc[2] - 2 * score;
    return js_max(1 / (1 + e), 0.0);
  }
  const double adjusted = score + qc[2] + xc[2] - cdp;
  return sim == SIM_MIP ? bbqo_scale_mip(adjusted) : js_max((1 + adjusted) / 2, 0.0);
}

// computeSimilarity, src/vectorSimilarity.ts:15-31 -> :38-67 (1/(1+sqrt(sum (a-b)^2))), :75-102, :110-120
double bbqo_similarity(const float* a, const float* b, int d, int sim) {
  if (sim == SIM_COSINE) return bbqo_cosine(a, b, d);
  double s = 0;
  if (sim == SIM_EUCLIDEAN) {
    for (int i = 0; i < d; i++) {
      const double diff = (double)a[i] - (double)b[i];
      s += diff * diff;
    }
    return 1.0 / (1.0 + std::sqrt(s));
  }
  for (int i = 0; i < d; i++) s += (double)a[i] * (double)b[i];
  return s;
}

// BinaryQuantizedScorer.computeQuantizationAccuracy, src/binaryQuantizedScorer.ts:524-617 (+ computeStandardDeviation,
// computePearsonCorrelation): out5 = meanError, maxError, minError, stdError, correlation.
void bbqo_accuracy_stats(const double* orig, const double* quant, int64_t n, double* out5) {
  std::vector<double> errors;
  double sumError = 0, maxError = 0, minError = std::numeric_limits<double>::infinity();
  for (int64_t i = 0; i < n; i++) {
    const double e = std::fabs(orig[i] - quant[i]);
    errors.push_back(e);
    sumError += e;
    maxError = js_max(maxError, e);
    minError = js_min(minError, e);
  }
  const double mean = sumError / (double)errors.size();
  double ss = 0;
  for (double v : errors) {
    const double diff = v - mean;
    ss += diff * diff;
  }
  const double sd = std::sqrt(ss / (double)errors.size());
  double sx = 0, sy = 0, sxy = 0, sx2 = 0, sy2 = 0;
  for (int64_t i = 0; i < n; i++) {
    sx += orig[i];
    sy += quant[i];
    sxy += orig[i] * quant[i];
    sx2 += orig[i] * orig[i];
    sy2 += quant[i] * quant[i];
  }
  const double dn = (double)n;
  const double num = dn * sxy - sx * sy;
  const double den = std::sqrt((dn * sx2 - sx * sx) * (dn * sy2 - sy * sy));
  out5[0] = mean;
  out5[1] = maxError;
  out5[2] = minError;
  out5[3] = sd;
  out5[4] = den == 0 ? 0.0 : num / den;
}

// BinaryQuantizationFormat.computeQuantizationAccuracy, src/binaryQuantizationFormat.ts:420-475: quantise the rows,
// then score EVERY query against row `target` (the reference: 0) through the single-vector scorer and exactly.
// Returns 0, or -1 for a query width the single-vector scorer rejects (:96).  orig/quant (n each) may be null.
int bbqo_quantization_accuracy(const float* rows, const float* queries, int64_t n, int d, int sim, int query_bits,
                               double lambda, int iters, int64_t target, double* out5, double* orig_out,
                               double* quant_out) {
  if (query_bits != 1 && query_bits != 4) return -1;
  std::vector<float> centroid(d);
  std::vector<uint8_t> packed((size_t)n * ((d + 7) / 8)), unpacked((size_t)n * d);
  std::vector<double> corr((size_t)n * 4);
  bbqo_build_index(rows, n, d, sim, 1, lambda, iters, nullptr, centroid.data(), packed.data(), unpacked.data(),
                   corr.data());
  const double cdp = query_bits == 1 ? bbqo_centroid_dp(centroid.data(), d) : 0.0;
  std::vector<double> orig(n), quant(n);
  std::vector<uint8_t> qcodes(d);
  double qc[4];
  for (int64_t i = 0; i < n; i++) {
    const float* q = queries + i * (int64_t)d;
    bbqo_quantize_query_once(q, centroid.data(), d, sim, query_bits, lambda, iters, qcodes.data(), qc);
    const int32_t dot = bbqo_dot_unpacked(qcodes.data(), unpacked.data() + target * (int64_t)d, d);
    quant[i] = bbqo_score_single((double)dot, corr.data() + 4 * target, qc, d, cdp, sim, query_bits);
    orig[i] = bbqo_similarity(q, rows + target * (int64_t)d, d, sim);
  }
  bbqo_accuracy_stats(orig.data(), quant.data(), n, out5);
  if (orig_out) std::memcpy(orig_out, orig.data(), sizeof(double) * n);
  if (quant_out) std::memcpy(quant_out, quant.data(), sizeof(double) * n);
  return 0;
}

}  // extern "C"

// ---- EXTENSION (NOT reference behaviour; "parity unpinned by construction") -------------------------------------
// indexBits >= 2.  The reference accepts the config (src/binaryQuantizationFormat.ts:143-148) and quantises the
// index with it (rows of D unpacked codes 0 .. 2^indexBits-1, :221-249), but its search throws: the batch path's
// createDirectPackedBuffer rejects unpacked rows (src/batchDotProduct.ts:420-436) and the per-vector fallback only
// knows 1- and 4-bit queries and has no 1/(2^indexBits-1) factor (src/binaryQuantizedScorer.ts:69-98).  SURVEY §8c
// recommends the natural generalisation, which is what this build (oracle AND device) defines:
//   qcDist = sum_d q[d] * x[d]                       (integer, unpacked codes)
//   lx = (upper_i - lower_i) / (2^indexBits - 1),    ly = (upper_q - lower_q) / (2^queryBits - 1)
//   score = ax*ay*D + ay*lx*x1 + ax*ly*y1 + lx*ly*qcDist        (left to right, no FMA)
//   EUCLIDEAN max(1/(1 + addq + addi - 2*score), 0);  adj = score + addq + addi - centroidDP (centroidDP = c.c)
//   COSINE max((1 + adj)/2, 0);  MAXIMUM_INNER_PRODUCT scaleMaxInnerProductScore(adj)  (src/utils.ts:171-176)
// Everything around it is the reference's: quantisers, the double COSINE normalisation of the query, the f32 store
// of the score, the heap / canonical selection.
double score_ext(double dot, const double xc[4], const double qc[4], int d, double cdp, int sim, int query_bits,
                 int index_bits) {
  const double x1 = xc[3], ax = xc[0], ay = qc[0], y1 = qc[3];
  const double lx = (xc[1] - ax) / (double)((1 << index_bits) - 1);
  const double ly = (qc[1] - ay) / (double)((1 << query_bits) - 1);
  const double score = ax * ay * (double)d + ay * lx * x1 + ax * ly * y1 + lx * ly * dot;
  if (sim == SIM_EUCLIDEAN) {
    const double e = qc[2] + xc[2] - 2 * score;
    return js_max(1 / (1 + e), 0);
  }
  const double adj = score + qc[2] + xc[2] - cdp;
  if (sim == SIM_COSINE) return js_max((1 + adj) / 2, 0);
  return adj < 0 ? 1 / (1 - adj) : adj + 1;
}

extern "C" {
// searchNearestNeighbors over an index of UNPACKED codes (n x d bytes, values 0 .. 2^index_bits-1), index_bits >= 2.
int64_t bbqo_search_ext(const float* query, const float* centroid, const uint8_t* codes, const double* xcorr,
                        int64_t n, int d, int sim, int query_bits, int index_bits, double lambda, int iters,
                        int64_t k, int mode, int32_t* out_idx, float* out_score, float* all_scores,
                        int32_t* all_dots) {
  if (k <= 0) return 0;
  std::vector<uint8_t> qcodes(d);
  double qcorr[4];
  bbqo_quantize_query(query, centroid, d, sim, query_bits, lambda, iters, qcodes.data(), qcorr);
  std::vector<int32_t> dots_local;
  int32_t* dots = all_dots;
  if (!dots) {
    dots_local.resize(n);
    dots = dots_local.data();
  }
  for (int64_t v = 0; v < n; v++) dots[v] = bbqo_dot_unpacked(qcodes.data(), codes + v * (int64_t)d, d);
  std::vector<float> sc_local;
  float* sc = all_scores;
  if (!sc) {
    sc_local.resize(n);
    sc = sc_local.data();
  }
  const double cdp = bbqo_centroid_dp(centroid, d);
  for (int64_t v = 0; v < n; v++)
    sc[v] = (float)score_ext((double)dots[v], xcorr + 4 * v, qcorr, d, cdp, sim, query_bits, index_bits);
  return mode == 0 ? bbqo_topk_heap(sc, n, k, out_idx, out_score) : bbqo_topk_canonical(sc, n, k, out_idx, out_score);
}
double bbqo_score_ext(double dot, const double* xc, const double* qc, int d, double cdp, int sim, int query_bits,
                      int index_bits) {
  return score_ext(dot, xc, qc, d, cdp, sim, query_bits, index_bits);
}
}  // extern "C"
