// Parity mode (fp32 activations / weights, CUDA-core FFMA): launchers the C ABI entry points dispatch to when their tensors are
// STP_F32.  See f32_path.cu.
#pragma once
#include "bn_fin.cuh"
#include "common.cuh"

namespace stp {
namespace f32 {

struct ConvF {
  const float* x;
  int ldx, N, H, W, Cin;
  const float* w;  // forward: KRSC [Cout][R][S][Cin].  dgrad = 1: the FORWARD weights, read tap-flipped / transposed in-kernel
  float* y;
  int ldy, Ho, Wo, Cout;
  const float* res;
  int ldr;
  const float* bias;
  int R, S, stride, pad_h, pad_w, up, relu, dgrad;
  int64_t M;
  int K;
};

inline bool f32_ok(const stp_tensor* t) { return t && t->ptr && t->dtype == STP_F32 && t->ld >= t->c; }

int launch_conv(const ConvF& p, cudaStream_t st);
int launch_wgrad(const ConvF& p, const float* dy, int lddy, int cdy, float* dw, cudaStream_t st);
// mode 0: sums of x, x^2; mode 1: BatchNorm-backward sums; then FinArgs finalisation (fin.acc required)
int launch_reduce(int mode, const stp_tensor* x, const stp_tensor* dy, const float* coef, int relu, int pool, const FinArgs& fin,
                  cudaStream_t st);
int bn_apply(const stp_tensor* x, const float* coef, int relu, int up, const stp_tensor* y, cudaStream_t st);
int bwd_apply(int mode, const stp_tensor* dy, const stp_tensor* x, const float* coef, const float* bcoef, int relu, int pool,
              const stp_tensor* res, const stp_tensor* dx, cudaStream_t st);
int stem_prep(const uint8_t* img, int64_t rows, int cimg, const float* coef, const stp_tensor* y, cudaStream_t st);
int copy_up(const stp_tensor* x, int up, const stp_tensor* y, cudaStream_t st);
int add(const stp_tensor* a, const stp_tensor* b, const stp_tensor* y, cudaStream_t st);
int maxpool_fwd(const stp_tensor* x, int k, int stride, int pad, const stp_tensor* y, uint8_t* argmax, cudaStream_t st);
int maxpool_bwd(const stp_tensor* dy, const uint8_t* argmax, int k, int stride, int pad, const stp_tensor* res, const stp_tensor* dx,
                cudaStream_t st);
int stem_wgrad_post(float* dw8, const float* w, int taps, int cpad, int cimg, float* dbeta, cudaStream_t st);
// DeepLabV3 graph (f32_deeplab.cu): depthwise conv (weights = the fp32 master [k][k][C]), whole-map mean / broadcast, dropout
int dwconv_fwd(const stp_dwconv_desc* d, const stp_tensor* x, const float* w, const stp_tensor* y, cudaStream_t st);
int dwconv_dgrad(const stp_dwconv_desc* d, const stp_tensor* dy, const float* w, const stp_tensor* res, const stp_tensor* dx, cudaStream_t st);
int dwconv_wgrad(const stp_dwconv_desc* d, const stp_tensor* x, const stp_tensor* dy, float* dw, cudaStream_t st);
int spatial_reduce(const stp_tensor* x, double scale, const stp_tensor* y, cudaStream_t st);
int spatial_bcast(const stp_tensor* x, float scale, const stp_tensor* res, const stp_tensor* y, cudaStream_t st);
int avgpool_fwd(const stp_tensor* x, int k, const stp_tensor* y, cudaStream_t st);
int avgpool_bwd(const stp_tensor* dy, int k, const stp_tensor* res, const stp_tensor* dx, cudaStream_t st);
int dropout(const stp_tensor* x, uint32_t thresh, float scale, uint64_t seed, uint32_t salt, const int64_t* d_step, const stp_tensor* y,
            cudaStream_t st);

}  // namespace f32
}  // namespace stp
