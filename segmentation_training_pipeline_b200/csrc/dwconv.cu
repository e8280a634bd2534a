// Depthwise 3x3 (k x k) convolution, NHWC bf16, stride 1 / 2, dilation (atrous rate) >= 1 -- keras DepthwiseConv2D of the in-tree
// DeepLabV3+ / MobileNetV2 model the reference registers as architecture `DeepLabV3` (impl/deeplab/model.py:236-275,
// registration segmentation.py:31-33; every example config of the reference uses it).  2 * k^2 FLOP per output element against
// 4 bytes of traffic: HBM-bound byte work on the CUDA cores -- thread = (pixel, 8-channel vector), 16-byte accesses, weights
// (k*k*C fp32, rounded to bf16 in compute like every other conv operand) through the read-only path.
//   fwd  : y[n,ho,wo,c]  = sum_{r,s} x[n, ho*stride - pad_h + r*dil, wo*stride - pad_w + s*dil, c] * w[r][s][c]
//   dgrad: dx[n,h,w,c]   = sum_{r,s} dy[n, (h + pad_h - r*dil)/stride, (w + pad_w - s*dil)/stride, c] * w[r][s][c]  (divisible taps)
//   wgrad: dw[r][s][c]   = sum_{n,ho,wo} dy[n,ho,wo,c] * x[...]    per-block partials + fixed-order double reduction
// pad_h / pad_w are the padding BEFORE (TF 'same' with stride 2 on even sizes pads 0 before, 1 after: the output size implies
// the padding after).
#include "common.cuh"

namespace stp {

constexpr int kDwMaxTaps = 25;

struct DwP {
  const __nv_bfloat16* x;
  int ldx, N, H, W, C;
  const float* w;   // [k][k][C] fp32 master
  __nv_bfloat16* y;
  int ldy, Ho, Wo;
  int k, stride, dil, pad_h, pad_w;
};

__device__ __forceinline__ float bfr(float v) { return __bfloat162float(__float2bfloat16(v)); }

__global__ void __launch_bounds__(256) dwconv_fwd_kernel(const DwP p) {
  const int cv = p.C / 8;
  const int64_t total = (int64_t)p.N * p.Ho * p.Wo * cv;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int v = (int)(i % cv);
    const int64_t r = i / cv;
    const int64_t n = r / ((int64_t)p.Ho * p.Wo);
    const int rem = (int)(r - n * (int64_t)p.Ho * p.Wo);
    const int ho = rem / p.Wo, wo = rem - ho * p.Wo;
    float acc[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) acc[c] = 0.f;
    for (int a = 0; a < p.k; ++a) {
      const int hi = ho * p.stride - p.pad_h + a * p.dil;
      if (hi < 0 || hi >= p.H) continue;
      for (int b = 0; b < p.k; ++b) {
        const int wi = wo * p.stride - p.pad_w + b * p.dil;
        if (wi < 0 || wi >= p.W) continue;
        float f[8];
        unpack8(ld8(p.x + ((n * p.H + hi) * (int64_t)p.W + wi) * p.ldx + v * 8), f);
        const float* wp = p.w + (a * p.k + b) * p.C + v * 8;
#pragma unroll
        for (int c = 0; c < 8; ++c) acc[c] += f[c] * bfr(__ldg(wp + c));
      }
    }
    st8(p.y + r * p.ldy + v * 8, pack8(acc));
  }
}

// x := dy (Ho x Wo), y := dx (H x W) in DwP terms of the FORWARD conv geometry
__global__ void __launch_bounds__(256) dwconv_dgrad_kernel(const DwP p, const __nv_bfloat16* __restrict__ dy, int lddy,
                                                           const __nv_bfloat16* __restrict__ res, int ldr, __nv_bfloat16* __restrict__ dx,
                                                           int lddx) {
  const int cv = p.C / 8;
  const int64_t total = (int64_t)p.N * p.H * p.W * cv;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int v = (int)(i % cv);
    const int64_t r = i / cv;
    const int64_t n = r / ((int64_t)p.H * p.W);
    const int rem = (int)(r - n * (int64_t)p.H * p.W);
    const int h = rem / p.W, w = rem - h * p.W;
    float acc[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) acc[c] = 0.f;
    for (int a = 0; a < p.k; ++a) {
      const int th = h + p.pad_h - a * p.dil;
      if (th < 0 || th % p.stride != 0) continue;
      const int ho = th / p.stride;
      if (ho >= p.Ho) continue;
      for (int b = 0; b < p.k; ++b) {
        const int tw = w + p.pad_w - b * p.dil;
        if (tw < 0 || tw % p.stride != 0) continue;
        const int wo = tw / p.stride;
        if (wo >= p.Wo) continue;
        float g[8];
        unpack8(ld8(dy + ((n * p.Ho + ho) * (int64_t)p.Wo + wo) * lddy + v * 8), g);
        const float* wp = p.w + (a * p.k + b) * p.C + v * 8;
#pragma unroll
        for (int c = 0; c < 8; ++c) acc[c] += g[c] * bfr(__ldg(wp + c));
      }
    }
    if (res) {
      float rf[8];
      unpack8(ld8(res + r * ldr + v * 8), rf);
#pragma unroll
      for (int c = 0; c < 8; ++c) acc[c] += rf[c];
    }
    st8(dx + r * lddx + v * 8, pack8(acc));
  }
}

// partial[blk][tap][C]: thread = (8-channel vector, pixel lane) walking this block's output pixels
__global__ void __launch_bounds__(256) dwconv_wgrad_kernel(const DwP p, const __nv_bfloat16* __restrict__ dy, int lddy,
                                                           float* __restrict__ partial, int64_t pix_per_blk) {
  extern __shared__ float sm[];   // [lanes][C] per tap pass
  const int cv = p.C / 8;
  const int lanes = blockDim.x / cv;             // launcher: cv divides blockDim
  const int v = threadIdx.x % cv, pl = threadIdx.x / cv;
  const int64_t M = (int64_t)p.N * p.Ho * p.Wo;
  const int64_t m_begin = (int64_t)blockIdx.x * pix_per_blk;
  int64_t m_end = m_begin + pix_per_blk;
  if (m_end > M) m_end = M;
  const int taps = p.k * p.k;
  for (int t = 0; t < taps; ++t) {
    const int a = t / p.k, b = t - a * p.k;
    float acc[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) acc[c] = 0.f;
    for (int64_t m = m_begin + pl; m < m_end; m += lanes) {
      const int64_t n = m / ((int64_t)p.Ho * p.Wo);
      const int rem = (int)(m - n * (int64_t)p.Ho * p.Wo);
      const int ho = rem / p.Wo, wo = rem - ho * p.Wo;
      const int hi = ho * p.stride - p.pad_h + a * p.dil, wi = wo * p.stride - p.pad_w + b * p.dil;
      if (hi < 0 || hi >= p.H || wi < 0 || wi >= p.W) continue;
      float g[8], f[8];
      unpack8(ld8(dy + m * lddy + v * 8), g);
      unpack8(ld8(p.x + ((n * p.H + hi) * (int64_t)p.W + wi) * p.ldx + v * 8), f);
#pragma unroll
      for (int c = 0; c < 8; ++c) acc[c] += g[c] * f[c];
    }
    __syncthreads();
#pragma unroll
    for (int c = 0; c < 8; ++c) sm[pl * p.C + v * 8 + c] = acc[c];
    __syncthreads();
    for (int c = threadIdx.x; c < p.C; c += blockDim.x) {
      float s = 0.f;
      for (int l = 0; l < lanes; ++l) s += sm[l * p.C + c];
      partial[((int64_t)blockIdx.x * taps + t) * p.C + c] = s;
    }
  }
}
__global__ void dwconv_wgrad_final_kernel(const float* __restrict__ partial, int nblk, int n_out, float* __restrict__ dw) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_out) return;
  double s = 0.0;
  for (int b = 0; b < nblk; ++b) s += (double)partial[(int64_t)b * n_out + i];
  dw[i] = (float)s;
}

static int dw_grid(int64_t total) {
  int64_t b = (total + 255) / 256;
  const int64_t cap = (int64_t)kNumSMs * 16;
  return (int)(b < 1 ? 1 : (b < cap ? b : cap));
}
static int dw_blocks(int64_t M) {
  int64_t nb = (M + 2047) / 2048;
  if (nb > kNumSMs * 4) nb = kNumSMs * 4;
  return (int)(nb < 1 ? 1 : nb);
}

}  // namespace stp

using namespace stp;

static int dw_check(const stp_dwconv_desc* d, const stp_tensor* x, const stp_tensor* y, const char* who) {
  STP_REQUIRE(d && x && y, "%s: null argument", who);
  STP_REQUIRE(vec_ok(x) && vec_ok(y) && x->c == y->c && x->n == y->n, "%s: tensors must be bf16 NHWC with equal channel counts (c%%8==0)", who);
  STP_REQUIRE(d->k >= 1 && d->k * d->k <= kDwMaxTaps && d->stride >= 1 && d->dilation >= 1 && d->pad_h >= 0 && d->pad_w >= 0,
              "%s: square filter up to 5x5, stride >= 1, dilation >= 1", who);
  STP_REQUIRE(x->c <= 2048, "%s: at most 2048 channels", who);
  return STP_OK;
}
static DwP make_dw(const stp_dwconv_desc* d, const stp_tensor* x, const float* w, const stp_tensor* y) {
  DwP p;
  p.x = (const __nv_bfloat16*)x->ptr; p.ldx = x->ld; p.N = x->n; p.H = x->h; p.W = x->w; p.C = x->c;
  p.w = w;
  p.y = (__nv_bfloat16*)y->ptr; p.ldy = y->ld; p.Ho = y->h; p.Wo = y->w;
  p.k = d->k; p.stride = d->stride; p.dil = d->dilation; p.pad_h = d->pad_h; p.pad_w = d->pad_w;
  return p;
}

extern "C" int stp_dwconv_fwd(const stp_dwconv_desc* d, const stp_tensor* x, const float* w_rsc, const stp_tensor* y, stp_stream stream) {
  int rc = dw_check(d, x, y, "dwconv_fwd");
  if (rc) return rc;
  STP_REQUIRE(w_rsc, "dwconv_fwd: null weights");
  DwP p = make_dw(d, x, w_rsc, y);
  dwconv_fwd_kernel<<<dw_grid(pixels(y) * (x->c / 8)), 256, 0, (cudaStream_t)stream>>>(p);
  return check_launch("dwconv_fwd");
}

extern "C" int stp_dwconv_dgrad(const stp_dwconv_desc* d, const stp_tensor* dy, const float* w_rsc, const stp_tensor* residual,
                                const stp_tensor* dx, stp_stream stream) {
  int rc = dw_check(d, dx, dy, "dwconv_dgrad");
  if (rc) return rc;
  STP_REQUIRE(w_rsc, "dwconv_dgrad: null weights");
  if (residual) STP_REQUIRE(vec_ok(residual) && residual->c == dx->c && pixels(residual) == pixels(dx), "dwconv_dgrad: bad residual");
  DwP p = make_dw(d, dx, w_rsc, dy);   // forward geometry: input = dx's tensor, output = dy's tensor
  dwconv_dgrad_kernel<<<dw_grid(pixels(dx) * (dx->c / 8)), 256, 0, (cudaStream_t)stream>>>(
      p, (const __nv_bfloat16*)dy->ptr, dy->ld, residual ? (const __nv_bfloat16*)residual->ptr : nullptr, residual ? residual->ld : 0,
      (__nv_bfloat16*)dx->ptr, dx->ld);
  return check_launch("dwconv_dgrad");
}

extern "C" size_t stp_dwconv_wgrad_workspace(const stp_dwconv_desc* d, const stp_tensor* x, const stp_tensor* dy) {
  if (!d || !x || !dy) return 0;
  return (size_t)dw_blocks(pixels(dy)) * d->k * d->k * x->c * sizeof(float);
}

extern "C" int stp_dwconv_wgrad(const stp_dwconv_desc* d, const stp_tensor* x, const stp_tensor* dy, float* dw_rsc, void* workspace,
                                size_t workspace_bytes, stp_stream stream) {
  int rc = dw_check(d, x, dy, "dwconv_wgrad");
  if (rc) return rc;
  STP_REQUIRE(dw_rsc && workspace, "dwconv_wgrad: null output / workspace");
  if (workspace_bytes < stp_dwconv_wgrad_workspace(d, x, dy)) {
    set_error("dwconv_wgrad: workspace too small");
    return STP_E_WORKSPACE;
  }
  DwP p = make_dw(d, x, nullptr, dy);
  const int cv = x->c / 8;
  const int lanes = 256 / cv > 0 ? 256 / cv : 1;   // cv <= 256 (C <= 2048)
  const int threads = cv * lanes;
  const int64_t M = pixels(dy);
  const int nblk = dw_blocks(M);
  const int64_t ppb = (M + nblk - 1) / nblk;
  cudaStream_t st = (cudaStream_t)stream;
  dwconv_wgrad_kernel<<<nblk, threads, (size_t)lanes * x->c * sizeof(float), st>>>(p, (const __nv_bfloat16*)dy->ptr, dy->ld,
                                                                                    (float*)workspace, ppb);
  rc = check_launch("dwconv_wgrad");
  if (rc) return rc;
  const int n_out = d->k * d->k * x->c;
  dwconv_wgrad_final_kernel<<<(n_out + 255) / 256, 256, 0, st>>>((const float*)workspace, nblk, n_out, dw_rsc);
  return check_launch("dwconv_wgrad_final");
}
