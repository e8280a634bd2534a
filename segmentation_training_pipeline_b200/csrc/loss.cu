// K11/K14: fused sigmoid + binary_crossentropy + dice/iou loss + metrics, one pass over (logits, mask).
// Formulas: keras.losses.binary_crossentropy (TF backend, from probabilities) and musket_core.losses
// dice / iou_coef / iot_coef (reference segmentation.py:15-22).  fp32 math, precise
// libm functions (no fast-math) so results track the fp32 oracle; deterministic two-stage reduction.
#include "common.cuh"

namespace stp {

constexpr int kLossBlocks = kNumSMs * 8;
constexpr int kLossSlots = 10;
constexpr float kJaccardSmooth = 100.f, kFocalGamma = 2.f, kFocalAlpha = 0.75f;  // musket_core.losses defaults [DEP]

__device__ __forceinline__ float sigmoidf_precise(float z) { return 1.f / (1.f + expf(-z)); }

__global__ void __launch_bounds__(256) loss_fwd_kernel(const float* __restrict__ logits,
                                                       const uint8_t* __restrict__ mask, int64_t count,
                                                       float* __restrict__ partial) {
  float s[kLossSlots];
#pragma unroll
  for (int i = 0; i < kLossSlots; ++i) s[i] = 0.f;
  const float eps = 1e-7f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x) {
    float z = logits[i];
    float t = (float)mask[i];
    float p = sigmoidf_precise(z);
    float pc = fminf(fmaxf(p, eps), 1.f - eps);
    float x = logf(pc / (1.f - pc));
    float l = fmaxf(x, 0.f) - x * t + log1pf(expf(-fabsf(x)));
    float hard = p > 0.5f ? 1.f : 0.f;
    s[0] += l;
    s[1] += p * t;
    s[2] += p;
    s[3] += t;
    s[4] += (hard == t) ? 1.f : 0.f;
    s[5] += hard * t;
    s[6] += hard;
    // jaccard_loss (per pixel over the 1-channel last axis, smooth = 100): (1 - (tp + S) / (t + p - tp + S)) * S
    s[7] += (1.f - (t * p + kJaccardSmooth) / (t + p - t * p + kJaccardSmooth)) * kJaccardSmooth;
    // focal_loss (gamma 2, alpha .75) on the clipped probability
    s[8] += t != 0.f ? -kFocalAlpha * (1.f - pc) * (1.f - pc) * logf(pc) : -(1.f - kFocalAlpha) * pc * pc * logf(1.f - pc);
  }
  __shared__ float sm[kLossSlots][8];
  int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int i = 0; i < kLossSlots; ++i) {
    float v = warp_sum(s[i]);
    if (lane == 0) sm[i][warp] = v;
  }
  __syncthreads();
  if (threadIdx.x < kLossSlots) {
    float a = 0.f;
    for (int w = 0; w < 8; ++w) a += sm[threadIdx.x][w];
    partial[(int64_t)blockIdx.x * kLossSlots + threadIdx.x] = a;
  }
}

// one warp per slot: lanes stride over the block partials (fixed order -> deterministic), double accumulation
__global__ void __launch_bounds__(32 * kLossSlots) loss_finalize_kernel(const float* __restrict__ partial, int nblk,
                                                                        double count, float w_bce, float w_dice,
                                                                        float w_iou, float w_jac, float w_focal,
                                                                        float* __restrict__ result) {
  __shared__ double tot[kLossSlots];
  {
    const int slot = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double a = 0.0;
    for (int b = lane; b < nblk; b += 32) a += (double)partial[(int64_t)b * kLossSlots + slot];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    if (lane == 0) tot[slot] = a;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double bce = tot[0] / count, I = tot[1], P = tot[2], T = tot[3];
    double dice = (2.0 * I + 1.0) / (P + T + 1.0);
    double iou = (I + 1.0) / (P + T - I + 1.0);
    double Ih = tot[5], Ph = tot[6];
    double iot = (Ih + 1.0) / (Ph + T - Ih + 1.0);
    const double jac = tot[7] / count, focal = tot[8] / count;
    double loss = (double)w_bce * bce + (double)w_dice * (1.0 - dice) + (double)w_iou * (1.0 - iou) +
                  (double)w_jac * jac + (double)w_focal * focal;
    result[STP_L_LOSS] = (float)loss;
    result[STP_L_BCE] = (float)bce;
    result[STP_L_DICE] = (float)dice;
    result[STP_L_IOU] = (float)iou;
    result[STP_L_ACC] = (float)(tot[4] / count);
    result[STP_L_IOT] = (float)iot;
    result[STP_L_SUM_P] = (float)P;
    result[STP_L_SUM_T] = (float)T;
    result[STP_L_SUM_PT] = (float)I;
    result[STP_L_COUNT] = (float)count;
    result[STP_L_LOVASZ] = 0.f;
    result[STP_L_JACCARD] = (float)jac;
    result[STP_L_FOCAL] = (float)focal;
    for (int i = 13; i < 16; ++i) result[i] = 0.f;
  }
}

__global__ void __launch_bounds__(256) loss_bwd_kernel(const float* __restrict__ logits,
                                                       const uint8_t* __restrict__ mask, int64_t count, float w_bce,
                                                       float w_dice, float w_iou, float w_jac, float w_focal,
                                                       const float* __restrict__ result, float* __restrict__ dlogits) {
  const float eps = 1e-7f;
  const float I = result[STP_L_SUM_PT], P = result[STP_L_SUM_P], T = result[STP_L_SUM_T];
  const float inv_count = 1.f / (float)count;
  const float S1 = P + T + 1.f;           // dice denominator
  const float U1 = P + T - I + 1.f;       // iou denominator
  const float dice_a = 2.f / S1, dice_b = (2.f * I + 1.f) / (S1 * S1);
  const float iou_a = 1.f / U1, iou_b = (I + 1.f) / (U1 * U1);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x) {
    float z = logits[i];
    float t = (float)mask[i];
    float p = sigmoidf_precise(z);
    float dp = 0.f;
    if (w_bce != 0.f && p > eps && p < 1.f - eps) dp += w_bce * inv_count * (p - t) / (p * (1.f - p));
    if (w_dice != 0.f) dp -= w_dice * (t * dice_a - dice_b);
    if (w_iou != 0.f) dp -= w_iou * (t * iou_a - (1.f - t) * iou_b);
    if (w_jac != 0.f) {
      const float D = t + p - t * p + kJaccardSmooth;
      dp -= w_jac * inv_count * kJaccardSmooth * (t * D - (t * p + kJaccardSmooth) * (1.f - t)) / (D * D);
    }
    if (w_focal != 0.f && p > eps && p < 1.f - eps) {
      const float q = 1.f - p;
      dp += w_focal * inv_count * (t != 0.f ? kFocalAlpha * (2.f * q * logf(p) - q * q / p)
                                            : -(1.f - kFocalAlpha) * (2.f * p * logf(q) - p * p / q));
    }
    dlogits[i] = dp * p * (1.f - p);
  }
}

// ---- softmax + categorical_crossentropy (schema segmentation.raml:12-21 `categorical_crossentropy`, :62-63 `activation: softmax`)
// keras.losses.categorical_crossentropy on PROBABILITIES with the TF backend [DEP]: p = softmax(z); p /= sum(p); p = clip(p,
// eps, 1-eps); l = -sum_c t_c*log(p_c); mean over pixels.  Also the categorical accuracy (argmax match, first maximum wins).
constexpr int kMaxCls = 4;
template <int C>
__device__ __forceinline__ void softmax_c(const float* __restrict__ z, float* p) {
  float m = z[0];
#pragma unroll
  for (int c = 1; c < C; ++c) m = fmaxf(m, z[c]);
  float s = 0.f;
#pragma unroll
  for (int c = 0; c < C; ++c) {
    p[c] = expf(z[c] - m);
    s += p[c];
  }
#pragma unroll
  for (int c = 0; c < C; ++c) p[c] /= s;
}
template <int C>
__global__ void __launch_bounds__(256) cce_fwd_kernel(const float* __restrict__ logits, const uint8_t* __restrict__ mask, int64_t M,
                                                      float* __restrict__ partial) {
  const float eps = 1e-7f;
  float s0 = 0.f, s1 = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < M; i += (int64_t)gridDim.x * blockDim.x) {
    float p[C];
    softmax_c<C>(logits + i * C, p);
    float sum = 0.f;
#pragma unroll
    for (int c = 0; c < C; ++c) sum += p[c];
    float l = 0.f;
    int ap = 0, at = 0;
#pragma unroll
    for (int c = 0; c < C; ++c) {
      const float t = (float)mask[i * C + c];
      const float q = fminf(fmaxf(p[c] / sum, eps), 1.f - eps);
      l -= t * logf(q);
      if (p[c] > p[ap]) ap = c;
      if (mask[i * C + c] > mask[i * C + at]) at = c;
    }
    s0 += l;
    s1 += ap == at ? 1.f : 0.f;
  }
  __shared__ float sm[2][8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  s0 = warp_sum(s0);
  s1 = warp_sum(s1);
  if (lane == 0) {
    sm[0][warp] = s0;
    sm[1][warp] = s1;
  }
  __syncthreads();
  if (threadIdx.x < 2) {
    float a = 0.f;
    for (int w = 0; w < 8; ++w) a += sm[threadIdx.x][w];
    partial[(int64_t)blockIdx.x * 2 + threadIdx.x] = a;
  }
}
__global__ void __launch_bounds__(64) cce_finalize_kernel(const float* __restrict__ partial, int nblk, double count, float weight,
                                                          int accumulate, float* __restrict__ result) {
  __shared__ double tot[2];
  const int slot = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double a = 0.0;
  for (int b = lane; b < nblk; b += 32) a += (double)partial[(int64_t)b * 2 + slot];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
  if (lane == 0) tot[slot] = a;
  __syncthreads();
  if (threadIdx.x == 0) {
    const float cce = (float)(tot[0] / count);
    result[STP_L_CCE] = cce;
    result[STP_L_CACC] = (float)(tot[1] / count);
    result[STP_L_LOSS] = (accumulate ? result[STP_L_LOSS] : 0.f) + weight * cce;
  }
}
template <int C>
__global__ void __launch_bounds__(256) cce_bwd_kernel(const float* __restrict__ logits, const uint8_t* __restrict__ mask, int64_t M,
                                                      float scale, int accumulate, float* __restrict__ dlogits) {
  const float eps = 1e-7f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < M; i += (int64_t)gridDim.x * blockDim.x) {
    float p[C], dq[C], dp[C];
    softmax_c<C>(logits + i * C, p);
    float sum = 0.f;
#pragma unroll
    for (int c = 0; c < C; ++c) sum += p[c];
    // l = -sum_c t_c log(clip(p_c / sum)):  dl/dq_c = -t_c / q_c inside the clip range, 0 outside
    float dot = 0.f;
#pragma unroll
    for (int c = 0; c < C; ++c) {
      const float q = p[c] / sum;
      dq[c] = (q > eps && q < 1.f - eps) ? -(float)mask[i * C + c] / q : 0.f;
      dot += dq[c] * p[c];
    }
    // q_c = p_c / sum:  dl/dp_k = dq_k / sum - dot / sum^2
    float dot2 = 0.f;
#pragma unroll
    for (int c = 0; c < C; ++c) {
      dp[c] = dq[c] / sum - dot / (sum * sum);
      dot2 += dp[c] * p[c];
    }
    // softmax:  dl/dz_j = p_j (dp_j - sum_k p_k dp_k)
#pragma unroll
    for (int c = 0; c < C; ++c) {
      const float g = scale * p[c] * (dp[c] - dot2);
      dlogits[i * C + c] = accumulate ? dlogits[i * C + c] + g : g;
    }
  }
}

}  // namespace stp

using namespace stp;

extern "C" int stp_softmax_cce_fwd(const float* logits, const uint8_t* mask, int64_t pixels, int32_t classes, float weight,
                                   int32_t accumulate, float* partial, float* result16, stp_stream stream) {
  STP_REQUIRE(logits && mask && partial && result16 && pixels > 0, "softmax_cce_fwd: bad args");
  STP_REQUIRE(classes >= 2 && classes <= kMaxCls, "softmax_cce_fwd: 2 <= classes <= 4");
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t nb = (pixels + 255) / 256;
  const int nblk = (int)(nb < kLossBlocks ? nb : kLossBlocks);
  if (classes == 2) cce_fwd_kernel<2><<<nblk, 256, 0, st>>>(logits, mask, pixels, partial);
  else if (classes == 3) cce_fwd_kernel<3><<<nblk, 256, 0, st>>>(logits, mask, pixels, partial);
  else cce_fwd_kernel<4><<<nblk, 256, 0, st>>>(logits, mask, pixels, partial);
  int rc = check_launch("softmax_cce_fwd");
  if (rc) return rc;
  cce_finalize_kernel<<<1, 64, 0, st>>>(partial, nblk, (double)pixels, weight, accumulate, result16);
  return check_launch("softmax_cce_finalize");
}

extern "C" int stp_softmax_cce_bwd(const float* logits, const uint8_t* mask, int64_t pixels, int32_t classes, float weight,
                                   int32_t accumulate, float* dlogits, stp_stream stream) {
  STP_REQUIRE(logits && mask && dlogits && pixels > 0, "softmax_cce_bwd: bad args");
  STP_REQUIRE(classes >= 2 && classes <= kMaxCls, "softmax_cce_bwd: 2 <= classes <= 4");
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t nb = (pixels + 255) / 256;
  const int nblk = (int)(nb < kLossBlocks ? nb : kLossBlocks);
  const float scale = weight / (float)pixels;
  if (classes == 2) cce_bwd_kernel<2><<<nblk, 256, 0, st>>>(logits, mask, pixels, scale, accumulate, dlogits);
  else if (classes == 3) cce_bwd_kernel<3><<<nblk, 256, 0, st>>>(logits, mask, pixels, scale, accumulate, dlogits);
  else cce_bwd_kernel<4><<<nblk, 256, 0, st>>>(logits, mask, pixels, scale, accumulate, dlogits);
  return check_launch("softmax_cce_bwd");
}

extern "C" size_t stp_loss_partial_floats(void) { return (size_t)kLossBlocks * kLossSlots; }

extern "C" int stp_loss_fwd(const float* logits, const uint8_t* mask, int64_t count, const stp_loss_spec* h_spec,
                            float* partial, float* result16, stp_stream stream) {
  STP_REQUIRE(logits && mask && h_spec && partial && result16 && count > 0, "loss_fwd: bad args");
  cudaStream_t st = (cudaStream_t)stream;
  int64_t nb = (count + 255) / 256;
  int nblk = (int)(nb < kLossBlocks ? nb : kLossBlocks);
  loss_fwd_kernel<<<nblk, 256, 0, st>>>(logits, mask, count, partial);
  int rc = check_launch("loss_fwd");
  if (rc) return rc;
  loss_finalize_kernel<<<1, 32 * kLossSlots, 0, st>>>(partial, nblk, (double)count, h_spec->w_bce, h_spec->w_dice, h_spec->w_iou,
                                                     h_spec->w_jaccard, h_spec->w_focal, result16);
  return check_launch("loss_finalize");
}

extern "C" int stp_loss_bwd(const float* logits, const uint8_t* mask, int64_t count, const stp_loss_spec* h_spec,
                            const float* result16, float* dlogits, stp_stream stream) {
  STP_REQUIRE(logits && mask && h_spec && result16 && dlogits && count > 0, "loss_bwd: bad args");
  int64_t nb = (count + 255) / 256;
  int nblk = (int)(nb < kLossBlocks ? nb : kLossBlocks);
  loss_bwd_kernel<<<nblk, 256, 0, (cudaStream_t)stream>>>(logits, mask, count, h_spec->w_bce, h_spec->w_dice,
                                                          h_spec->w_iou, h_spec->w_jaccard, h_spec->w_focal, result16, dlogits);
  return check_launch("loss_bwd");
}
