// keras Dropout(rate) in training mode (DeepLabV3+ head, impl/deeplab/model.py:486: Dropout(0.1) after concat_projection):
//     y = x * keep / (1 - rate),  keep ~ Bernoulli(1 - rate) per element.
// The mask is never stored: it is a pure function of (seed, salt, step, element index) through Philox4x32-10 (the generator of
// the augmentation kernels, twin oracle/philox.py), so the backward launches the same kernel on the gradient.  TensorFlow's own
// random stream is not reproducible outside TF; parity is on the distribution and, against the oracle, on the identical mask.
//   counter = (step lo, octet index lo, (octet index hi << 1) | half, step hi);  key = (seed lo, seed hi ^ salt)
//   half 0 -> channels 0..3 of the 8-channel octet, half 1 -> channels 4..7;  keep  <=>  word >= rate * 2^32
#include "common.cuh"
#include "f32_path.h"

namespace stp {

__device__ __forceinline__ void philox_do(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t* out) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

__global__ void __launch_bounds__(256) dropout_kernel(const __nv_bfloat16* __restrict__ x, int ldx, __nv_bfloat16* __restrict__ y, int ldy,
                                                      int64_t total, int cv, uint32_t thresh, float scale, uint64_t seed, uint32_t salt,
                                                      const int64_t* __restrict__ d_step) {
  const int64_t step = *d_step;
  const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32) ^ salt;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / cv;
    const int v = (int)(i - r * cv);
    uint32_t w[8];
    philox_do((uint32_t)step, (uint32_t)i, ((uint32_t)((uint64_t)i >> 32) << 1) | 0u, (uint32_t)(step >> 32), k0, k1, w);
    philox_do((uint32_t)step, (uint32_t)i, ((uint32_t)((uint64_t)i >> 32) << 1) | 1u, (uint32_t)(step >> 32), k0, k1, w + 4);
    float f[8];
    unpack8(ld8(x + r * ldx + v * 8), f);
#pragma unroll
    for (int c = 0; c < 8; ++c) f[c] = w[c] >= thresh ? f[c] * scale : 0.f;
    st8(y + r * ldy + v * 8, pack8(f));
  }
}

}  // namespace stp

using namespace stp;

extern "C" int stp_dropout(const stp_tensor* x, float rate, uint64_t seed, uint32_t salt, const int64_t* d_step, const stp_tensor* y,
                           stp_stream stream) {
  STP_REQUIRE(rate >= 0.f && rate < 1.f, "dropout: 0 <= rate < 1");
  if (x && x->dtype == STP_F32) {   // parity mode: the same mask on fp32 tensors
    STP_REQUIRE(f32::f32_ok(x) && f32::f32_ok(y) && x->c == y->c && x->c % 8 == 0 && pixels(x) == pixels(y) && d_step,
                "dropout (fp32): tensors of one shape, c %% 8 == 0, device step");
    const double tf = (double)rate * 4294967296.0;
    return f32::dropout(x, tf >= 4294967295.0 ? 0xFFFFFFFFu : (uint32_t)tf, 1.f / (1.f - rate), seed, salt, d_step, y, (cudaStream_t)stream);
  }
  STP_REQUIRE(vec_ok(x) && vec_ok(y) && x->c == y->c && pixels(x) == pixels(y) && d_step, "dropout: bf16 tensors of one shape, device step");
  const int cv = x->c / 8;
  const int64_t total = pixels(x) * cv;
  const double t = (double)rate * 4294967296.0;
  const uint32_t thresh = t >= 4294967295.0 ? 0xFFFFFFFFu : (uint32_t)t;
  int64_t nb = (total + 255) / 256;
  const int64_t cap = (int64_t)kNumSMs * 16;
  if (nb > cap) nb = cap;
  dropout_kernel<<<(int)(nb < 1 ? 1 : nb), 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)x->ptr, x->ld, (__nv_bfloat16*)y->ptr, y->ld,
                                                                           total, cv, thresh, 1.f / (1.f - rate), seed, salt, d_step);
  return check_launch("dropout");
}
