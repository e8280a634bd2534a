// Depthwise 3x3 (k x k) convolution, NHWC bf16, stride 1 / 2, dilation (atrous rate) >= 1 -- keras DepthwiseConv2D of the in-tree
// DeepLabV3+ / MobileNetV2 model the reference registers as architecture `DeepLabV3` (impl/deeplab/model.py:236-275,
// registration segmentation.py:31-33; every example config of the reference uses it).  2 * k^2 FLOP per output element against
// 4 bytes of traffic: HBM-bound byte work on the CUDA cores -- thread = (pixel pair, 8-channel vector), 16-byte accesses, weights
// bf16 [k][k][C] (the bf16 copy stp_weight_prep makes of the fp32 master with Cout = 1, like every other conv operand).
//   fwd  : y[n,ho,wo,c]  = sum_{r,s} x[n, ho*stride - pad_h + r*dil, wo*stride - pad_w + s*dil, c] * w[r][s][c]
//   dgrad: dx[n,h,w,c]   = sum_{r,s} dy[n, (h + pad_h - r*dil)/stride, (w + pad_w - s*dil)/stride, c] * w[r][s][c]  (divisible taps)
//   wgrad: dw[r][s][c]   = sum_{n,ho,wo} dy[n,ho,wo,c] * x[...]    per-block partials + fixed-order double reduction
// pad_h / pad_w are the padding BEFORE (TF 'same' with stride 2 on even sizes pads 0 before, 1 after: the output size implies
// the padding after).
#include "common.cuh"
#include "f32_path.h"

namespace stp {

constexpr int kDwMaxTaps = 25;

struct DwP {
  const __nv_bfloat16* x;
  int ldx, N, H, W, C;
  __nv_bfloat16* y;
  int ldy, Ho, Wo;
  int k, stride, dil, pad_h, pad_w;
};

__device__ __forceinline__ bf16x8 ld8_or_zero(const __nv_bfloat16* p, bool ok) {
  uint4 v = make_uint4(0u, 0u, 0u, 0u);
  if (ok) v = *reinterpret_cast<const uint4*>(p);   // predicated load, no branch
  return *reinterpret_cast<bf16x8*>(&v);
}

// thread = (8-channel octet, kDwP adjacent output pixels along W).  Every tap is a predicated 16-byte load (no branches, so
// all loads of a thread are in flight together); the bf16 weights of a tap are one 16-byte load shared by the kDwP pixels.
// (First version: fp32 weights re-rounded per pixel behind branchy bounds checks, 630 instructions per output octet, issue- and
// latency-bound at 0.9 TB/s -- profiles/r2_dwconv_probe.txt.)
constexpr int kDwP = 2;

template <int K>
__global__ void __launch_bounds__(256) dwconv_fwd_kernel(const DwP p, const __nv_bfloat16* __restrict__ wq, int total, int strips) {
  const int cv = p.C / 8;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int v = i % cv;
    int r = i / cv;
    const int sw = r % strips;
    r /= strips;
    const int ho = r % p.Ho;
    const int n = r / p.Ho;
    const int wo0 = sw * kDwP;
    const __nv_bfloat16* xb = p.x + (int64_t)n * p.H * p.W * p.ldx + v * 8;
    float acc[kDwP][8];
#pragma unroll
    for (int q = 0; q < kDwP; ++q)
#pragma unroll
      for (int c = 0; c < 8; ++c) acc[q][c] = 0.f;
#pragma unroll
    for (int a = 0; a < K; ++a) {
      const int hi = ho * p.stride - p.pad_h + a * p.dil;
      const bool okh = (unsigned)hi < (unsigned)p.H;
      const __nv_bfloat16* row = xb + (int64_t)(okh ? hi : 0) * p.W * p.ldx;
#pragma unroll
      for (int b = 0; b < K; ++b) {
        float w[8];
        unpack8(ld8(wq + (a * K + b) * p.C + v * 8), w);
#pragma unroll
        for (int q = 0; q < kDwP; ++q) {
          const int wi = (wo0 + q) * p.stride - p.pad_w + b * p.dil;
          const bool ok = okh && (unsigned)wi < (unsigned)p.W;
          float f[8];
          unpack8(ld8_or_zero(row + (int64_t)(ok ? wi : 0) * p.ldx, ok), f);
#pragma unroll
          for (int c = 0; c < 8; ++c) acc[q][c] += f[c] * w[c];
        }
      }
    }
    __nv_bfloat16* yb = p.y + ((int64_t)(n * p.Ho + ho) * p.Wo) * p.ldy + v * 8;
#pragma unroll
    for (int q = 0; q < kDwP; ++q)
      if (wo0 + q < p.Wo) st8(yb + (int64_t)(wo0 + q) * p.ldy, pack8(acc[q]));
  }
}

// p describes the FORWARD conv geometry (x: H x W input, y: Ho x Wo output); the pixel pairs run along W of dx
template <int K>
__global__ void __launch_bounds__(256) dwconv_dgrad_kernel(const DwP p, const __nv_bfloat16* __restrict__ wq,
                                                           const __nv_bfloat16* __restrict__ dy, int lddy,
                                                           const __nv_bfloat16* __restrict__ res, int ldr, __nv_bfloat16* __restrict__ dx,
                                                           int lddx, int total, int strips) {
  const int cv = p.C / 8;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int v = i % cv;
    int r = i / cv;
    const int sw = r % strips;
    r /= strips;
    const int h = r % p.H;
    const int n = r / p.H;
    const int w0 = sw * kDwP;
    const __nv_bfloat16* gb = dy + (int64_t)n * p.Ho * p.Wo * lddy + v * 8;
    float acc[kDwP][8];
#pragma unroll
    for (int q = 0; q < kDwP; ++q)
#pragma unroll
      for (int c = 0; c < 8; ++c) acc[q][c] = 0.f;
#pragma unroll
    for (int a = 0; a < K; ++a) {
      const int th = h + p.pad_h - a * p.dil;
      const int ho = th / p.stride;
      const bool okh = th >= 0 && ho * p.stride == th && ho < p.Ho;
      const __nv_bfloat16* row = gb + (int64_t)(okh ? ho : 0) * p.Wo * lddy;
#pragma unroll
      for (int b = 0; b < K; ++b) {
        float wt[8];
        unpack8(ld8(wq + (a * K + b) * p.C + v * 8), wt);
#pragma unroll
        for (int q = 0; q < kDwP; ++q) {
          const int tw = w0 + q + p.pad_w - b * p.dil;
          const int wo = tw / p.stride;
          const bool ok = okh && tw >= 0 && wo * p.stride == tw && wo < p.Wo;
          float g[8];
          unpack8(ld8_or_zero(row + (int64_t)(ok ? wo : 0) * lddy, ok), g);
#pragma unroll
          for (int c = 0; c < 8; ++c) acc[q][c] += g[c] * wt[c];
        }
      }
    }
    const int64_t rowpix = (int64_t)(n * p.H + h) * p.W;
#pragma unroll
    for (int q = 0; q < kDwP; ++q) {
      if (w0 + q >= p.W) continue;
      if (res) {
        float rf[8];
        unpack8(ld8(res + (rowpix + w0 + q) * ldr + v * 8), rf);
#pragma unroll
        for (int c = 0; c < 8; ++c) acc[q][c] += rf[c];
      }
      st8(dx + (rowpix + w0 + q) * lddx + v * 8, pack8(acc[q]));
    }
  }
}

// Stride-1 correlation with tap reuse: out[h][w] = sum_{a,b} in[h + oh + a*dil][w + ow + b*dil] * wsel(a, b).  A thread owns kDwS
// output pixels of one row spaced by the atrous rate (w_q = wbase + q*dil), so tap b of pixel q reads column q + b of a window
// of kDwS + K - 1 columns: K * (kDwS + K - 1) predicated loads for kDwS outputs (18 for 4 with K = 3, against 36), each unpacked
// once.  FLIP selects w[K-1-a][K-1-b]: the dgrad of a stride-1 depthwise conv is the same correlation over dy with the flipped
// filter and offsets (pad - (K-1)*dil).  The L1 data path, not DRAM, bounds these kernels (profiles/r2_dwconv_probe.txt).
constexpr int kDwS = 4;

template <int K, bool FLIP>
__global__ void __launch_bounds__(256, 2) dwconv_s1_kernel(const __nv_bfloat16* __restrict__ in, int ldi, int Hi, int Wi,
                                                        const __nv_bfloat16* __restrict__ wq, const __nv_bfloat16* __restrict__ res,
                                                        int ldr, __nv_bfloat16* __restrict__ out, int ldo, int Ho, int Wo, int C,
                                                        int dil, int oh, int ow, int total, int groups) {
  const int cv = C / 8;
  constexpr int NC = kDwS + K - 1;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int v = i % cv;
    int r = i / cv;
    const int gidx = r % groups;          // (residue class j, group g): wbase = j + g * kDwS * dil
    r /= groups;
    const int h = r % Ho;
    const int n = r / Ho;
    const int j = gidx % dil, g = gidx / dil;
    const int wbase = j + g * kDwS * dil;
    if (wbase >= Wo) continue;
    const __nv_bfloat16* ib = in + (int64_t)n * Hi * Wi * ldi + v * 8;
    float acc[kDwS][8];
#pragma unroll
    for (int q = 0; q < kDwS; ++q)
#pragma unroll
      for (int c = 0; c < 8; ++c) acc[q][c] = 0.f;
#pragma unroll
    for (int a = 0; a < K; ++a) {
      const int hi = h + oh + a * dil;
      const bool okh = (unsigned)hi < (unsigned)Hi;
      const __nv_bfloat16* row = ib + (int64_t)(okh ? hi : 0) * Wi * ldi;
      float col[NC][8];
#pragma unroll
      for (int m = 0; m < NC; ++m) {
        const int wi = wbase + ow + m * dil;
        const bool ok = okh && (unsigned)wi < (unsigned)Wi;
        unpack8(ld8_or_zero(row + (int64_t)(ok ? wi : 0) * ldi, ok), col[m]);
      }
#pragma unroll
      for (int b = 0; b < K; ++b) {
        const int t = FLIP ? (K - 1 - a) * K + (K - 1 - b) : a * K + b;
        float w[8];
        unpack8(ld8(wq + t * C + v * 8), w);
#pragma unroll
        for (int q = 0; q < kDwS; ++q)
#pragma unroll
          for (int c = 0; c < 8; ++c) acc[q][c] += col[q + b][c] * w[c];
      }
    }
    const int64_t rowpix = (int64_t)(n * Ho + h) * Wo;
#pragma unroll
    for (int q = 0; q < kDwS; ++q) {
      const int w = wbase + q * dil;
      if (w >= Wo) continue;
      if (res) {
        float rf[8];
        unpack8(ld8(res + (rowpix + w) * ldr + v * 8), rf);
#pragma unroll
        for (int c = 0; c < 8; ++c) acc[q][c] += rf[c];
      }
      st8(out + (rowpix + w) * ldo + v * 8, pack8(acc[q]));
    }
  }
}

// partial[chunk][tap][C]: block = (pixel chunk, group of up to 32 channel octets); thread = (octet, pixel lane).  Every thread
// keeps all K*K tap sums of its 8 channels in registers, so dy is read once and x K*K times (L1/L2 hits: neighbouring taps).
template <int K>
__global__ void __launch_bounds__(256) dwconv_wgrad_kernel(const DwP p, const __nv_bfloat16* __restrict__ dy, int lddy,
                                                           float* __restrict__ partial, int ppb, int cvb, int lanes) {
  extern __shared__ float sm[];   // [lanes][cvb*8]
  constexpr int T = K * K;
  const int cv = p.C / 8;
  const int vl = threadIdx.x % cvb, pl = threadIdx.x / cvb;
  const int v = blockIdx.y * cvb + vl;
  const bool active = v < cv && pl < lanes;
  const int HoWo = p.Ho * p.Wo;
  const int M = p.N * HoWo;
  const int m_begin = blockIdx.x * ppb;
  const int m_end = min(m_begin + ppb, M);
  float acc[T][8];
#pragma unroll
  for (int t = 0; t < T; ++t)
#pragma unroll
    for (int c = 0; c < 8; ++c) acc[t][c] = 0.f;
  if (active)
    for (int m = m_begin + pl; m < m_end; m += lanes) {
      const int n = m / HoWo;
      const int rem = m - n * HoWo;
      const int ho = rem / p.Wo, wo = rem - ho * p.Wo;
      float g[8];
      unpack8(ld8(dy + (int64_t)m * lddy + v * 8), g);
      const __nv_bfloat16* xb = p.x + (int64_t)n * p.H * p.W * p.ldx + v * 8;
#pragma unroll
      for (int a = 0; a < K; ++a) {
        const int hi = ho * p.stride - p.pad_h + a * p.dil;
        const bool okh = (unsigned)hi < (unsigned)p.H;
        const __nv_bfloat16* row = xb + (int64_t)(okh ? hi : 0) * p.W * p.ldx;
#pragma unroll
        for (int b = 0; b < K; ++b) {
          const int wi = wo * p.stride - p.pad_w + b * p.dil;
          const bool ok = okh && (unsigned)wi < (unsigned)p.W;
          float f[8];
          unpack8(ld8_or_zero(row + (int64_t)(ok ? wi : 0) * p.ldx, ok), f);
#pragma unroll
          for (int c = 0; c < 8; ++c) acc[a * K + b][c] += g[c] * f[c];
        }
      }
    }
  const int row = cvb * 8;
#pragma unroll
  for (int t = 0; t < T; ++t) {
    __syncthreads();
    if (pl < lanes) {
#pragma unroll
      for (int c = 0; c < 8; ++c) sm[pl * row + vl * 8 + c] = acc[t][c];
    }
    __syncthreads();
    for (int c = threadIdx.x; c < row; c += blockDim.x) {
      const int ch = blockIdx.y * row + c;
      if (ch < p.C) {
        float s = 0.f;
        for (int l = 0; l < lanes; ++l) s += sm[l * row + c];
        partial[((int64_t)blockIdx.x * T + t) * p.C + ch] = s;
      }
    }
  }
}
// dw[i] = sum over the pixel chunks of partial[chunk][i], fixed order, double.  Block = 32 outputs x 8 chunk lanes (coalesced
// 128-byte reads of the partial rows); the first version had ONE thread walk all chunks of an output: 20 us per layer, 3 % of the
// DeepLabV3 step (profiles/r2_people_launches.summary.txt).
__global__ void __launch_bounds__(256) dwconv_wgrad_final_kernel(const float* __restrict__ partial, int nblk, int n_out, float* __restrict__ dw) {
  __shared__ double sm[8][33];
  const int ol = threadIdx.x & 31, lane = threadIdx.x >> 5;
  const int i = blockIdx.x * 32 + ol;
  double s = 0.0;
  if (i < n_out)
    for (int b = lane; b < nblk; b += 8) s += (double)partial[(int64_t)b * n_out + i];
  sm[lane][ol] = s;
  __syncthreads();
  if (lane == 0 && i < n_out) {
    double t = 0.0;
#pragma unroll
    for (int l = 0; l < 8; ++l) t += sm[l][ol];
    dw[i] = (float)t;
  }
}

static int dw_grid(int64_t total) {
  int64_t b = (total + 255) / 256;
  const int64_t cap = (int64_t)kNumSMs * 16;
  return (int)(b < 1 ? 1 : (b < cap ? b : cap));
}
// pixel chunks of the wgrad: 128 pixels per block, at most 512 chunks (workspace = chunks * K*K * C floats)
static int dw_ppb(int64_t M) {
  int64_t ppb = 128;
  while ((M + ppb - 1) / ppb > 512) ppb *= 2;
  return (int)ppb;
}
static int dw_blocks(int64_t M) {
  const int ppb = dw_ppb(M);
  const int64_t nb = (M + ppb - 1) / ppb;
  return (int)(nb < 1 ? 1 : nb);
}

}  // namespace stp

using namespace stp;

// parity mode: fp32 tensors, fp32 master weights
static bool dw_f32(const stp_dwconv_desc* d, const stp_tensor* x, const stp_tensor* y) {
  return d && x && y && x->dtype == STP_F32 && f32::f32_ok(x) && f32::f32_ok(y) && x->c == y->c && x->n == y->n && d->k >= 1 && d->k <= 5 &&
         d->stride >= 1 && d->dilation >= 1 && d->pad_h >= 0 && d->pad_w >= 0;
}

static int dw_check(const stp_dwconv_desc* d, const stp_tensor* x, const stp_tensor* y, const char* who) {
  STP_REQUIRE(d && x && y, "%s: null argument", who);
  STP_REQUIRE(vec_ok(x) && vec_ok(y) && x->c == y->c && x->n == y->n, "%s: tensors must be bf16 NHWC with equal channel counts (c%%8==0)", who);
  STP_REQUIRE((d->k == 1 || d->k == 3 || d->k == 5) && d->stride >= 1 && d->dilation >= 1 && d->pad_h >= 0 && d->pad_w >= 0,
              "%s: filter 1x1, 3x3 or 5x5, stride >= 1, dilation >= 1", who);
  STP_REQUIRE(x->c <= 2048, "%s: at most 2048 channels", who);
  return STP_OK;
}
static DwP make_dw(const stp_dwconv_desc* d, const stp_tensor* x, const stp_tensor* y) {
  DwP p;
  p.x = (const __nv_bfloat16*)x->ptr; p.ldx = x->ld; p.N = x->n; p.H = x->h; p.W = x->w; p.C = x->c;
  p.y = (__nv_bfloat16*)y->ptr; p.ldy = y->ld; p.Ho = y->h; p.Wo = y->w;
  p.k = d->k; p.stride = d->stride; p.dil = d->dilation; p.pad_h = d->pad_h; p.pad_w = d->pad_w;
  return p;
}

extern "C" int stp_dwconv_fwd(const stp_dwconv_desc* d, const stp_tensor* x, const void* w_kkc, const stp_tensor* y, stp_stream stream) {
  if (x && x->dtype == STP_F32) {
    STP_REQUIRE(dw_f32(d, x, y) && w_kkc, "dwconv_fwd (fp32): bad arguments");
    return f32::dwconv_fwd(d, x, (const float*)w_kkc, y, (cudaStream_t)stream);
  }
  int rc = dw_check(d, x, y, "dwconv_fwd");
  if (rc) return rc;
  STP_REQUIRE(w_kkc && aligned16(w_kkc), "dwconv_fwd: weights (bf16 [k][k][c]) must be 16-byte aligned");
  DwP p = make_dw(d, x, y);
  const __nv_bfloat16* wq = (const __nv_bfloat16*)w_kkc;
  const int strips = (y->w + kDwP - 1) / kDwP;
  const int64_t total = (int64_t)y->n * y->h * strips * (x->c / 8);
  STP_REQUIRE(total < (int64_t)1 << 31 && pixels(x) < (int64_t)1 << 31, "dwconv_fwd: tensor too large");
  cudaStream_t st = (cudaStream_t)stream;
  if (d->k == 3 && d->stride == 1) {
    const int groups = d->dilation * (((y->w + d->dilation - 1) / d->dilation + kDwS - 1) / kDwS);
    const int64_t tot = (int64_t)y->n * y->h * groups * (x->c / 8);
    STP_REQUIRE(tot < (int64_t)1 << 31, "dwconv_fwd: tensor too large");
    dwconv_s1_kernel<3, false><<<dw_grid(tot), 256, 0, st>>>(p.x, p.ldx, p.H, p.W, wq, nullptr, 0, p.y, p.ldy, p.Ho, p.Wo, p.C, p.dil,
                                                             -p.pad_h, -p.pad_w, (int)tot, groups);
    return check_launch("dwconv_fwd");
  }
  if (d->k == 3) dwconv_fwd_kernel<3><<<dw_grid(total), 256, 0, st>>>(p, wq, (int)total, strips);
  else if (d->k == 5) dwconv_fwd_kernel<5><<<dw_grid(total), 256, 0, st>>>(p, wq, (int)total, strips);
  else dwconv_fwd_kernel<1><<<dw_grid(total), 256, 0, st>>>(p, wq, (int)total, strips);
  return check_launch("dwconv_fwd");
}

extern "C" int stp_dwconv_dgrad(const stp_dwconv_desc* d, const stp_tensor* dy, const void* w_kkc, const stp_tensor* residual,
                                const stp_tensor* dx, stp_stream stream) {
  if (dx && dx->dtype == STP_F32) {
    STP_REQUIRE(dw_f32(d, dx, dy) && w_kkc && (!residual || (f32::f32_ok(residual) && residual->c == dx->c && pixels(residual) == pixels(dx))),
                "dwconv_dgrad (fp32): bad arguments");
    return f32::dwconv_dgrad(d, dy, (const float*)w_kkc, residual, dx, (cudaStream_t)stream);
  }
  int rc = dw_check(d, dx, dy, "dwconv_dgrad");
  if (rc) return rc;
  STP_REQUIRE(w_kkc && aligned16(w_kkc), "dwconv_dgrad: weights (bf16 [k][k][c]) must be 16-byte aligned");
  if (residual) STP_REQUIRE(vec_ok(residual) && residual->c == dx->c && pixels(residual) == pixels(dx), "dwconv_dgrad: bad residual");
  DwP p = make_dw(d, dx, dy);   // forward geometry: input = dx's tensor, output = dy's tensor
  const __nv_bfloat16* wq = (const __nv_bfloat16*)w_kkc;
  const int strips = (dx->w + kDwP - 1) / kDwP;
  const int64_t total = (int64_t)dx->n * dx->h * strips * (dx->c / 8);
  STP_REQUIRE(total < (int64_t)1 << 31 && pixels(dx) < (int64_t)1 << 31, "dwconv_dgrad: tensor too large");
  cudaStream_t st = (cudaStream_t)stream;
  const __nv_bfloat16* dyp = (const __nv_bfloat16*)dy->ptr;
  const __nv_bfloat16* rp = residual ? (const __nv_bfloat16*)residual->ptr : nullptr;
  const int ldr = residual ? residual->ld : 0;
  __nv_bfloat16* dxp = (__nv_bfloat16*)dx->ptr;
  if (d->k == 3 && d->stride == 1) {
    // dx[h][w] = sum_{a,b} dy[h + pad_h - a*dil][w + pad_w - b*dil] w[a][b]: correlation over dy with the flipped filter
    const int groups = d->dilation * (((dx->w + d->dilation - 1) / d->dilation + kDwS - 1) / kDwS);
    const int64_t tot = (int64_t)dx->n * dx->h * groups * (dx->c / 8);
    STP_REQUIRE(tot < (int64_t)1 << 31, "dwconv_dgrad: tensor too large");
    dwconv_s1_kernel<3, true><<<dw_grid(tot), 256, 0, st>>>(dyp, dy->ld, p.Ho, p.Wo, wq, rp, ldr, dxp, dx->ld, p.H, p.W, p.C, p.dil,
                                                            p.pad_h - 2 * p.dil, p.pad_w - 2 * p.dil, (int)tot, groups);
    return check_launch("dwconv_dgrad");
  }
  if (d->k == 3) dwconv_dgrad_kernel<3><<<dw_grid(total), 256, 0, st>>>(p, wq, dyp, dy->ld, rp, ldr, dxp, dx->ld, (int)total, strips);
  else if (d->k == 5) dwconv_dgrad_kernel<5><<<dw_grid(total), 256, 0, st>>>(p, wq, dyp, dy->ld, rp, ldr, dxp, dx->ld, (int)total, strips);
  else dwconv_dgrad_kernel<1><<<dw_grid(total), 256, 0, st>>>(p, wq, dyp, dy->ld, rp, ldr, dxp, dx->ld, (int)total, strips);
  return check_launch("dwconv_dgrad");
}

extern "C" size_t stp_dwconv_wgrad_workspace(const stp_dwconv_desc* d, const stp_tensor* x, const stp_tensor* dy) {
  if (!d || !x || !dy) return 0;
  return (size_t)dw_blocks(pixels(dy)) * d->k * d->k * x->c * sizeof(float);
}

extern "C" int stp_dwconv_wgrad(const stp_dwconv_desc* d, const stp_tensor* x, const stp_tensor* dy, float* dw_rsc, void* workspace,
                                size_t workspace_bytes, stp_stream stream) {
  if (x && x->dtype == STP_F32) {
    STP_REQUIRE(dw_f32(d, x, dy) && dw_rsc, "dwconv_wgrad (fp32): bad arguments");
    return f32::dwconv_wgrad(d, x, dy, dw_rsc, (cudaStream_t)stream);
  }
  int rc = dw_check(d, x, dy, "dwconv_wgrad");
  if (rc) return rc;
  STP_REQUIRE(dw_rsc && workspace, "dwconv_wgrad: null output / workspace");
  if (workspace_bytes < stp_dwconv_wgrad_workspace(d, x, dy)) {
    set_error("dwconv_wgrad: workspace too small");
    return STP_E_WORKSPACE;
  }
  DwP p = make_dw(d, x, dy);
  const int cv = x->c / 8;
  const int64_t M = pixels(dy);
  STP_REQUIRE(M < (int64_t)1 << 31, "dwconv_wgrad: fewer than 2^31 output pixels");
  const int cvb = cv < 32 ? cv : 32;
  const int lanes = 256 / cvb;
  const int nblk = dw_blocks(M);
  const int ppb = dw_ppb(M);
  cudaStream_t st = (cudaStream_t)stream;
  dim3 grid(nblk, (cv + cvb - 1) / cvb);
  const size_t smem = (size_t)lanes * cvb * 8 * sizeof(float);
  const __nv_bfloat16* dyp = (const __nv_bfloat16*)dy->ptr;
  if (d->k == 3) dwconv_wgrad_kernel<3><<<grid, 256, smem, st>>>(p, dyp, dy->ld, (float*)workspace, ppb, cvb, lanes);
  else if (d->k == 5) dwconv_wgrad_kernel<5><<<grid, 256, smem, st>>>(p, dyp, dy->ld, (float*)workspace, ppb, cvb, lanes);
  else if (d->k == 1) dwconv_wgrad_kernel<1><<<grid, 256, smem, st>>>(p, dyp, dy->ld, (float*)workspace, ppb, cvb, lanes);
  else {
    set_error("dwconv_wgrad: filter sizes 1, 3 and 5 are built");
    return STP_E_UNSUPPORTED;
  }
  rc = check_launch("dwconv_wgrad");
  if (rc) return rc;
  const int n_out = d->k * d->k * x->c;
  dwconv_wgrad_final_kernel<<<(n_out + 31) / 32, 256, 0, st>>>((const float*)workspace, nblk, n_out, dw_rsc);
  return check_launch("dwconv_wgrad_final");
}
