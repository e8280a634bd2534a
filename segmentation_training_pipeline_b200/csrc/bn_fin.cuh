// BatchNorm finalisation shared by the reduction kernels (bn.cu) and the conv epilogue that accumulates the forward
// statistics of its own output (conv_tc2.cu): partial sums -> mean / invstd / scale / shift (+ moving statistics),
// keras.layers.BatchNormalization training-mode semantics.
#pragma once
#include "common.cuh"

namespace stp {

// What the LAST block of a reduction kernel does with the partial sums (mode 0: nothing, a separate finalize kernel runs)
struct FinArgs {
  int mode;  // 0 none | 1 forward statistics -> coef (+ moving stats) | 2 backward -> dgamma, dbeta, bcoef | 3 column sums -> dbeta
  unsigned int* sync;
  double* acc;  // non-null: blocks add their sums here with double atomics (2*C, zero on entry, returned to zero) and the
                // last block finalises from 2*C values; null: deterministic fixed-order reduction of per-block partials
  double inv_count, bessel;
  const float* gamma;
  const float* beta;
  float eps, momentum;
  float* mov_mean;
  float* mov_var;
  float* coef;  // mode 1: output; mode 2: input
  float* dgamma;
  float* dbeta;
  float* bcoef;
};

// keras.layers.BatchNormalization (2.2.x, normalization.py) turns the biased batch variance into the moving-average update with
//   variance *= sample_size / (sample_size - (1.0 + epsilon))
// -- Bessel's correction with the layer's epsilon in the denominator [DEP keras>=2.2.4].
inline double keras_bessel(int64_t count, float eps) {
  return count > 1 ? (double)count / ((double)count - (1.0 + (double)eps)) : 1.0;
}

__device__ __forceinline__ void fin_forward(const FinArgs& f, int C, int c, double s, double ss) {
  double mean = s * f.inv_count;
  double var = ss * f.inv_count - mean * mean;
  if (var < 0.0) var = 0.0;
  double invstd = rsqrt(var + (double)f.eps);
  float g = f.gamma ? f.gamma[c] : 1.f;
  float b = f.beta ? f.beta[c] : 0.f;
  float scale = g * (float)invstd;
  f.coef[c] = (float)mean;
  f.coef[C + c] = (float)invstd;
  f.coef[2 * C + c] = scale;
  f.coef[3 * C + c] = b - (float)mean * scale;
  if (f.mov_mean) {
    f.mov_mean[c] = f.mov_mean[c] * f.momentum + (float)mean * (1.f - f.momentum);
    f.mov_var[c] = f.mov_var[c] * f.momentum + (float)(var * f.bessel) * (1.f - f.momentum);
  }
}
__device__ __forceinline__ void fin_backward(const FinArgs& f, int C, int c, double s, double ss) {
  if (f.dbeta) f.dbeta[c] = (float)s;
  if (f.dgamma) f.dgamma[c] = (float)ss;
  double mean = f.coef[c], invstd = f.coef[C + c], a = f.coef[2 * C + c];
  double b = -a * invstd * ss * f.inv_count;
  double cc = -a * s * f.inv_count - b * mean;
  f.bcoef[c] = (float)a;
  f.bcoef[C + c] = (float)b;
  f.bcoef[2 * C + c] = (float)cc;
}

}  // namespace stp
