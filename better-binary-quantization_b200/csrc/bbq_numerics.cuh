// bbq_numerics.cuh — the exact-arithmetic core shared by the quantiser kernels (K4/K5) and the score
// epilogue (K1/K2).  Everything here is IEEE binary64/binary32, evaluated in the reference's operation
// order, and MUST be compiled with --fmad=false (nvcc) / -ffp-contract=off (host): JavaScript never
// contracts a*b+c.  The functions are __host__ __device__ only so that tests/ can compile this very
// header for the host and compare it with the oracle bit-for-bit without a GPU; the product never
// runs them on the CPU.
//
// Reference (leolee9086/Better-Binary-Quantization, TypeScript) lines restated here:
//   src/optimizedScalarQuantizer.ts:108-227 scalarQuantize, :245-265 getInitialInterval,
//   :280-353 optimizeIntervals, :373-407 computeLoss; src/utils.ts:25-81 stats + clamp;
//   src/vectorOperations.ts:11-34 normalizeVector; src/batchDotProduct.ts:478-541,554-617 scores;
//   src/constants.ts:20,38-47.
#pragma once
#include <cmath>
#include <cstdint>

#if defined(__CUDACC__)
#define BBQ_HD __host__ __device__ __forceinline__
#else
#define BBQ_HD inline
#endif

namespace bbqn {

enum { SIM_EUCLIDEAN = 0, SIM_COSINE = 1, SIM_MIP = 2 };

// src/constants.ts:38-47, upper column (the lower column is its negation)
BBQ_HD double mse_grid(int bits) {
  switch (bits) {
    case 1: return 0.798;
    case 2: return 1.493;
    case 3: return 2.051;
    case 4: return 2.514;
    case 5: return 2.916;
    case 6: return 3.278;
    case 7: return 3.611;
    default: return 3.922;
  }
}

BBQ_HD double js_nan() { return (double)NAN; }
// Math.min / Math.max: NaN-propagating
BBQ_HD double js_min(double a, double b) { return (a != a || b != b) ? js_nan() : (a < b ? a : b); }
BBQ_HD double js_max(double a, double b) { return (a != a || b != b) ? js_nan() : (a > b ? a : b); }
// ... and with ECMAScript's zero rule (Math.min(+0, -0) = -0, Math.max(-0, +0) = +0): what the QUANTISERS use — the min / max
// over the centred components and clamp() — so that a vector holding signed zeros gets the reference's interval bit for
// bit (found by fuzzing the oracle against the executed reference).  The score formulas keep the plain forms above:
// their only use is Math.max(x, 0), for which the plain form already returns +0.
BBQ_HD double js_minz(double a, double b) {
  if (a != a || b != b) return js_nan();
  if (a == 0.0 && b == 0.0) return copysign(1.0, a) < 0.0 ? a : b;
  return a < b ? a : b;
}
BBQ_HD double js_maxz(double a, double b) {
  if (a != a || b != b) return js_nan();
  if (a == 0.0 && b == 0.0) return copysign(1.0, a) < 0.0 ? b : a;
  return a > b ? a : b;
}
// src/utils.ts:79-81
BBQ_HD double js_clamp(double x, double lo, double hi) { return js_minz(js_maxz(x, lo), hi); }
// ECMAScript Math.round: nearest integer, ties toward +infinity
BBQ_HD double js_round(double x) {
  if (!(fabs(x) < 4503599627370496.0)) return x;
  const double f = floor(x);
  return (x - f >= 0.5) ? f + 1.0 : f;
}
BBQ_HD bool js_isfinite(double x) { return fabs(x) <= 1.7976931348623157e308; }  // false for NaN/Inf

// w_i = Float32Array store of (v_i - c_i)
template <class V, class Cn>
BBQ_HD double centred(const V& v, const Cn& c, int i) {
  return (double)(float)((double)v(i) - (double)c(i));
}

// src/optimizedScalarQuantizer.ts:373-407
template <class V, class Cn>
BBQ_HD double osq_loss(const V& v, const Cn& c, int d, double a, double b, int points, double nrm,
                       double lambda) {
  const double step = (b - a) / (double)(points - 1);
  const double stepInv = 1.0 / step;
  double xe = 0.0, e = 0.0;
  for (int i = 0; i < d; i++) {
    const double xi = centred(v, c, i);
    const double clamped = js_clamp(xi, a, b);
    const double k = js_round((clamped - a) * stepInv);
    const double xiq = a + step * k;
    const double diff = xi - xiq;
    xe += xi * diff;
    e += diff * diff;
  }
  return (1.0 - lambda) * xe * xe / nrm + lambda * e;
}

struct OsqResult {
  double lower, upper, additional, nrm;
};

// Steps 1-5 of scalarQuantize: statistics, initial interval, coordinate descent.
// v(i): the (already normalised, for COSINE) vector; c(i): centroid.  Returns the final interval.
template <class V, class Cn>
BBQ_HD OsqResult osq_interval(const V& v, const Cn& c, int d, int bits, int sim, double lambda, int iters) {
  double centroidDot = 0.0;
  if (sim != SIM_EUCLIDEAN) {
    for (int i = 0; i < d; i++) centroidDot += (double)v(i) * (double)c(i);
  }
  double mn = 1.7976931348623157e308, mx = -1.7976931348623157e308;
  double sum = 0.0, n2 = 0.0;
  for (int i = 0; i < d; i++) {
    const double cv = (double)v(i) - (double)c(i);
    mn = js_minz(mn, cv);
    mx = js_maxz(mx, cv);
    const double w = (double)(float)cv;
    sum += w;      // computeMean  (src/utils.ts:41-50)
    n2 += w * w;   // computeL2Norm (src/utils.ts:25-34) — independent chains, same per-chain order
  }
  const double mean = sum / (double)d;
  double ss = 0.0;
  for (int i = 0; i < d; i++) {
    const double diff = centred(v, c, i) - mean;
    ss += diff * diff;
  }
  const double sd = sqrt(ss / (double)d);
  const double nrm = sqrt(n2);

  const double g = mse_grid(bits);
  double a = js_clamp(-g * sd + mean, mn, mx);
  double b = js_clamp(g * sd + mean, mn, mx);
  const int points = 1 << bits;

  // optimizeIntervals
  double loss0 = osq_loss(v, c, d, a, b, points, nrm, lambda);
  const double scale = (1.0 - lambda) / nrm;
  if (js_isfinite(scale)) {
    const double pm1 = (double)(points - 1);
    for (int iter = 0; iter < iters; iter++) {
      const double stepInv = pm1 / (b - a);
      double daa = 0, dab = 0, dbb = 0, dax = 0, dbx = 0;
      for (int i = 0; i < d; i++) {
        const double xi = centred(v, c, i);
        const double clamped = js_clamp(xi, a, b);
        const double k = js_round((clamped - a) * stepInv);
        const double s = k / pm1;
        const double oms = 1.0 - s;
        daa += oms * oms;
        dab += oms * s;
        dbb += s * s;
        dax += xi * oms;
        dbx += xi * s;
      }
      const double m0 = scale * dax * dax + lambda * daa;
      const double m1 = scale * dax * dbx + lambda * dab;
      const double m2 = scale * dbx * dbx + lambda * dbb;
      const double det = m0 * m2 - m1 * m1;
      if (fabs(det) < 1e-12) break;
      const double aOpt = (m2 * dax - m1 * dbx) / det;
      const double bOpt = (m0 * dbx - m1 * dax) / det;
      if (fabs(a - aOpt) < 1e-8 && fabs(b - bOpt) < 1e-8) break;
      const double loss1 = osq_loss(v, c, d, aOpt, bOpt, points, nrm, lambda);
      if (loss1 > loss0) break;
      a = aOpt;
      b = bOpt;
      loss0 = loss1;
    }
  }
  OsqResult r;
  r.lower = a;
  r.upper = b;
  r.additional = (sim == SIM_EUCLIDEAN) ? nrm : centroidDot;
  r.nrm = nrm;
  return r;
}

// Step 6 of scalarQuantize (src/optimizedScalarQuantizer.ts:192-216): emit(i, code) per component;
// returns quantizedComponentSum.
template <class V, class Cn, class Emit>
BBQ_HD double osq_codes(const V& v, const Cn& c, int d, int bits, double a, double b, Emit&& emit) {
  const int nSteps = (1 << bits) - 1;
  const double step = nSteps > 0 ? (b - a) / (double)nSteps : 0.0;
  const double stepInv = step > 0 ? 1.0 / step : 0.0;
  const double threshold = (a + b) / 2;
  double qsum = 0.0;
  for (int i = 0; i < d; i++) {
    const double xi = centred(v, c, i);
    const double clamped = js_clamp(xi, a, b);
    if (bits == 1) {
      const int q = clamped >= threshold ? 1 : 0;
      emit(i, (uint8_t)q);
      qsum += (double)q;
    } else {
      const double assignment = js_round((clamped - a) * stepInv);
      const double stored = js_min(assignment, (double)nSteps);
      emit(i, (stored != stored) ? (uint8_t)0 : (uint8_t)(((long long)stored) & 0xFF));  // Uint8Array store
      qsum += assignment;
    }
  }
  return qsum;
}

// src/vectorOperations.ts:11-34: returns the norm; caller stores (float)(v_i / norm), or 0 when norm == 0
template <class V>
BBQ_HD double l2norm_seq(const V& v, int d) {
  double n = 0.0;
  for (int i = 0; i < d; i++) n += (double)v(i) * (double)v(i);
  return sqrt(n);
}

// Per-query constants of the score formula, hoisted once per query.
struct QueryTerms {
  double ay;    // lowerInterval
  double ly;    // (upper - lower) [* FOUR_BIT_SCALE when queryBits != 1]
  double y1;    // quantizedComponentSum
  double addq;  // additionalCorrection
};

// Which score formula: the reference's 4-bit path (every queryBits != 1 with a 1-bit index: FOUR_BIT_SCALE = 1/15 and
// the batch-path MIP quirk), its 1-bit path, or this build's EXTENSION for indexBits >= 2 (the reference throws there;
// natural generalisation, SURVEY §8c: ly = (u_q - l_q)/(2^queryBits - 1), lx = (u_i - l_i)/(2^indexBits - 1), plain
// scaleMaxInnerProductScore).
enum { SCORE_REF_MULTIBIT = 0, SCORE_REF_ONEBIT = 1, SCORE_EXT = 2 };
BBQ_HD int score_mode(int query_bits, int index_bits) {
  return index_bits >= 2 ? SCORE_EXT : (query_bits == 1 ? SCORE_REF_ONEBIT : SCORE_REF_MULTIBIT);
}
// divisor of the index interval: lx = (upper_i - lower_i) / lx_div (1 for the reference's 1-bit index: x / 1.0 == x)
BBQ_HD double index_lx_div(int index_bits) { return (double)((1 << index_bits) - 1); }

BBQ_HD QueryTerms make_query_terms(double lower, double upper, double additional, double compsum,
                                   int query_bits, int index_bits = 1) {
  QueryTerms t;
  t.ay = lower;
  if (index_bits >= 2) t.ly = (upper - lower) / (double)((1 << query_bits) - 1);  // EXTENSION
  else
  t.ly = (query_bits == 1) ? (upper - lower) : (upper - lower) * (1.0 / 15.0);  // src/constants.ts:20
  t.y1 = compsum;
  t.addq = additional;
  return t;
}

// One corrected score as the reference's batch path computes it, INCLUDING the Float32Array store
// (src/binaryQuantizationFormat.ts:353,378).  ax = lower_i, lx = upper_i - lower_i, x1 = componentSum_i.
// queryBits != 1: src/batchDotProduct.ts:554-617;  queryBits == 1: :478-541 (different association).
// `mode`: SCORE_REF_MULTIBIT / SCORE_REF_ONEBIT / SCORE_EXT; lx = (upper_i - lower_i) / index_lx_div(indexBits).
BBQ_HD float score_f32(double dot, double ax, double lx, double addx, double x1, const QueryTerms& q,
                       double dim, double cdp, int sim, int mode) {
  double s = ax * q.ay * dim + q.ay * lx * x1 + ax * q.ly * q.y1 + lx * q.ly * dot;
  double r;
  if (sim == SIM_EUCLIDEAN) {
    const double e = q.addq + addx - 2 * s;
    r = js_max(1 / (1 + e), 0.0);
  } else if (mode == SCORE_EXT) {
    const double adj = s + q.addq + addx - cdp;
    if (sim == SIM_COSINE) r = js_max((1 + adj) / 2, 0.0);
    else r = (adj < 0) ? 1 / (1 - adj) : adj + 1;
  } else if (mode == SCORE_REF_ONEBIT) {
    s = s + (q.addq + addx - cdp);
    if (sim == SIM_COSINE) r = js_max((1 + s) / 2, 0.0);
    else r = (s < 0) ? 1 / (1 - s) : s + 1;
  } else {
    const double adj = s + q.addq + addx - cdp;
    if (sim == SIM_COSINE) r = js_max((1 + adj) / 2, 0.0);
    else r = (adj < 0) ? 1 / (1 - adj / (1.0 / 15.0)) : adj / (1.0 / 15.0) + 1;
  }
  return (float)r;
}

// src/utils.ts:171-176 scaleMaxInnerProductScore
BBQ_HD double scale_mip(double s) { return (s < 0) ? 1 / (1 - s) : s + 1; }

// One score as the SINGLE-VECTOR scorer computes it (computeQuantizedScore, src/binaryQuantizedScorer.ts:69-98:
// computeOneBitSimilarityScore :112-160 / computeFourBitSimilarityScore :174-217) — the path
// computeQuantizationAccuracy (src/binaryQuantizationFormat.ts:420-475) takes.  It is NOT the batch path's formula:
// the result stays f64, MIP uses the plain scaleMaxInnerProductScore, and the caller passes centroidDP = c.c for
// 1-bit queries but 0 for 4-bit queries (no originalQueryVector is given, :290).  lower/upper/add/sum as stored.
BBQ_HD double score_single_f64(double dot, double ax, double ux, double addx, double x1, double ay, double uy,
                               double addq, double y1, double dim, double cdp, int sim, bool one_bit_query) {
  const double lx = ux - ax;
  if (one_bit_query) {
    const double ly = uy - ay;
    double score = ax * ay * dim + ay * lx * x1 + ax * ly * y1 + lx * ly * dot;
    if (sim == SIM_EUCLIDEAN) {
      score = addq + addx - 2 * score;
      return js_max(1 / (1 + score), 0.0);
    }
    score += addq + addx - cdp;
    return (sim == SIM_COSINE) ? js_max((1 + score) / 2, 0.0) : scale_mip(score);
  }
  const double ly = (uy - ay) * (1.0 / 15.0);
  const double score = ax * ay * dim + ay * lx * x1 + ax * ly * y1 + lx * ly * dot;
  if (sim == SIM_EUCLIDEAN) {
    const double e = addq + addx - 2 * score;
    return js_max(1 / (1 + e), 0.0);
  }
  const double adjusted = score + addq + addx - cdp;
  return (sim == SIM_MIP) ? scale_mip(adjusted) : js_max((1 + adjusted) / 2, 0.0);
}

// computeQuantizationAccuracy(originalScores, quantizedScores), src/binaryQuantizedScorer.ts:524-617: strictly
// sequential f64 sums.  out5 = meanError, maxError, minError, stdError, correlation.
BBQ_HD void accuracy_stats(const double* orig, const double* quant, int64_t n, double* out5) {
  double sumError = 0, maxError = 0, minError = (double)INFINITY;
  for (int64_t i = 0; i < n; i++) {
    const double e = fabs(orig[i] - quant[i]);
    sumError += e;
    maxError = js_max(maxError, e);
    minError = js_min(minError, e);
  }
  const double mean = sumError / (double)n;
  double ss = 0;
  for (int64_t i = 0; i < n; i++) {
    const double d = fabs(orig[i] - quant[i]) - mean;
    ss += d * d;
  }
  const double sd = sqrt(ss / (double)n);
  double sx = 0, sy = 0, sxy = 0, sx2 = 0, sy2 = 0;
  for (int64_t i = 0; i < n; i++) {
    const double x = orig[i], y = quant[i];
    sx += x;
    sy += y;
    sxy += x * y;
    sx2 += x * x;
    sy2 += y * y;
  }
  const double dn = (double)n;
  const double num = dn * sxy - sx * sy;
  const double den = sqrt((dn * sx2 - sx * sx) * (dn * sy2 - sy * sy));
  out5[0] = mean;
  out5[1] = maxError;
  out5[2] = minError;
  out5[3] = sd;
  out5[4] = (den == 0) ? 0.0 : num / den;
}

// Total order used everywhere a top-k is taken: larger key = better.
// (f32 score descending, row id ascending); NaN ranks below everything; -0 == +0.
BBQ_HD uint64_t topk_key(float score, uint32_t id) {
  uint32_t u;
  if (score != score) {
    u = 0u;
  } else {
    const float s = score + 0.0f;  // -0 -> +0
#if defined(__CUDA_ARCH__)
    u = __float_as_uint(s);
#else
    union { float f; uint32_t i; } cv;
    cv.f = s;
    u = cv.i;
#endif
    u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
    // u >= 0x007FFFFF for every non-NaN (-inf maps to 0x007FFFFF), so 0 is free for NaN
  }
  return ((uint64_t)u << 32) | (uint64_t)(0xFFFFFFFFu - id);
}
BBQ_HD float topk_key_score(uint64_t key) {
  uint32_t u = (uint32_t)(key >> 32);
  if (u == 0u) return NAN;
  u = (u & 0x80000000u) ? (u & 0x7FFFFFFFu) : ~u;
#if defined(__CUDA_ARCH__)
  return __uint_as_float(u);
#else
  union { float f; uint32_t i; } cv;
  cv.i = u;
  return cv.f;
#endif
}
BBQ_HD uint32_t topk_key_id(uint64_t key) { return 0xFFFFFFFFu - (uint32_t)key; }

}  // namespace bbqn
