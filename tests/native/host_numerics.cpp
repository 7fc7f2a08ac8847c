// TEST-ONLY host build of better-binary-quantization_b200/csrc/bbq_numerics.cuh (the exact-arithmetic core
// the CUDA kernels use), so that the CPU suite can compare it with the oracle bit-for-bit without a GPU.
// Never loaded by the product.  Build: g++ -O2 -std=c++17 -ffp-contract=off -shared -fPIC.
#include <cstdint>
#include "../../better-binary-quantization_b200/csrc/bbq_numerics.cuh"

struct Acc {
  const float* p;
  float operator()(int i) const { return p[i]; }
};

extern "C" {
void hn_osq(const float* v, const float* c, int d, int bits, int sim, double lambda, int iters, uint8_t* codes,
            double* corr4) {
  Acc va{v}, ca{c};
  const bbqn::OsqResult r = bbqn::osq_interval(va, ca, d, bits, sim, lambda, iters);
  const double qsum = bbqn::osq_codes(va, ca, d, bits, r.lower, r.upper, [&](int i, uint8_t q) { codes[i] = q; });
  corr4[0] = r.lower;
  corr4[1] = r.upper;
  corr4[2] = r.additional;
  corr4[3] = qsum;
}
double hn_norm(const float* v, int d) {
  Acc va{v};
  return bbqn::l2norm_seq(va, d);
}
void hn_scores(const int32_t* dots, const double* xcorr, int64_t n, const double* qcorr4, int d, double cdp, int sim,
               int query_bits, float* out) {
  const bbqn::QueryTerms q = bbqn::make_query_terms(qcorr4[0], qcorr4[1], qcorr4[2], qcorr4[3], query_bits);
  for (int64_t i = 0; i < n; i++) {
    const double* x = xcorr + 4 * i;
    out[i] = bbqn::score_f32((double)dots[i], x[0], x[1] - x[0], x[2], (double)(uint32_t)x[3], q, (double)d, cdp, sim,
                             query_bits == 1);
  }
}
// any (queryBits, indexBits): the reference's formulas for a 1-bit index, this build's extension for a wider one —
// exactly as the scan kernels call the header (mode, lx divisor, query terms)
void hn_scores_bits(const int32_t* dots, const double* xcorr, int64_t n, const double* qcorr4, int d, double cdp, int sim,
                    int query_bits, int index_bits, float* out) {
  const bbqn::QueryTerms q = bbqn::make_query_terms(qcorr4[0], qcorr4[1], qcorr4[2], qcorr4[3], query_bits, index_bits);
  const int mode = bbqn::score_mode(query_bits, index_bits);
  const double div = bbqn::index_lx_div(index_bits);
  for (int64_t i = 0; i < n; i++) {
    const double* x = xcorr + 4 * i;
    out[i] = bbqn::score_f32((double)dots[i], x[0], (x[1] - x[0]) / div, x[2], (double)(uint32_t)x[3], q, (double)d, cdp,
                             sim, mode);
  }
}
double hn_score_single(double dot, const double* xc, const double* qc, int d, double cdp, int sim, int query_bits) {
  return bbqn::score_single_f64(dot, xc[0], xc[1], xc[2], xc[3], qc[0], qc[1], qc[2], qc[3], (double)d, cdp, sim,
                                query_bits == 1);
}
void hn_accuracy_stats(const double* orig, const double* quant, int64_t n, double* out5) {
  bbqn::accuracy_stats(orig, quant, n, out5);
}
uint64_t hn_topk_key(float score, uint32_t id) { return bbqn::topk_key(score, id); }
float hn_key_score(uint64_t key) { return bbqn::topk_key_score(key); }
uint32_t hn_key_id(uint64_t key) { return bbqn::topk_key_id(key); }
}
