// Parity mode (fp32 activations / gradients / weights) of the kernels the DeepLabV3+ / MobileNetV2 graph adds: depthwise
// convolution, whole-map mean / broadcast, dropout.  Plain CUDA-core kernels, one thread per output element, double accumulators,
// fixed summation order (deterministic) -- NOT a performance path; it exists so that this graph, too, can be checked against the
// fp64 oracle at the 1e-3 of north_star (bf16 storage alone is 30-40 % from fp32 at random init, profiles/r2_deeplab_depth_profile.txt).
#include "f32_path.h"

namespace stp {
namespace f32 {
namespace {

int grid1(int64_t total) {
  int64_t b = (total + 255) / 256;
  const int64_t cap = (int64_t)kNumSMs * 32;
  return (int)(b < 1 ? 1 : (b < cap ? b : cap));
}

struct DwF {
  int N, H, W, C, Ho, Wo, k, stride, dil, pad_h, pad_w;
};

__global__ void dw_fwd_kernel(const DwF p, const float* __restrict__ x, int ldx, const float* __restrict__ w, float* __restrict__ y, int ldy) {
  const int64_t total = (int64_t)p.N * p.Ho * p.Wo * p.C;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % p.C);
    int64_t r = i / p.C;
    const int wo = (int)(r % p.Wo);
    const int ho = (int)((r / p.Wo) % p.Ho);
    const int64_t n = r / ((int64_t)p.Wo * p.Ho);
    double acc = 0.0;
    for (int a = 0; a < p.k; ++a) {
      const int hi = ho * p.stride - p.pad_h + a * p.dil;
      if (hi < 0 || hi >= p.H) continue;
      for (int b = 0; b < p.k; ++b) {
        const int wi = wo * p.stride - p.pad_w + b * p.dil;
        if (wi < 0 || wi >= p.W) continue;
        acc += (double)x[((n * p.H + hi) * (int64_t)p.W + wi) * ldx + c] * (double)w[(a * p.k + b) * p.C + c];
      }
    }
    y[r * ldy + c] = (float)acc;
  }
}

__global__ void dw_dgrad_kernel(const DwF p, const float* __restrict__ dy, int lddy, const float* __restrict__ w,
                                const float* __restrict__ res, int ldr, float* __restrict__ dx, int lddx) {
  const int64_t total = (int64_t)p.N * p.H * p.W * p.C;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % p.C);
    int64_t r = i / p.C;
    const int wq = (int)(r % p.W);
    const int h = (int)((r / p.W) % p.H);
    const int64_t n = r / ((int64_t)p.W * p.H);
    double acc = 0.0;
    for (int a = 0; a < p.k; ++a) {
      const int th = h + p.pad_h - a * p.dil;
      if (th < 0 || th % p.stride != 0 || th / p.stride >= p.Ho) continue;
      for (int b = 0; b < p.k; ++b) {
        const int tw = wq + p.pad_w - b * p.dil;
        if (tw < 0 || tw % p.stride != 0 || tw / p.stride >= p.Wo) continue;
        acc += (double)dy[((n * p.Ho + th / p.stride) * (int64_t)p.Wo + tw / p.stride) * lddy + c] * (double)w[(a * p.k + b) * p.C + c];
      }
    }
    if (res) acc += (double)res[r * ldr + c];
    dx[r * lddx + c] = (float)acc;
  }
}

// block = one (tap, channel) pair; 256 threads stride the output pixels, fixed-order tree in shared memory
__global__ void __launch_bounds__(256) dw_wgrad_kernel(const DwF p, const float* __restrict__ x, int ldx, const float* __restrict__ dy,
                                                       int lddy, float* __restrict__ dw) {
  __shared__ double red[256];
  const int c = blockIdx.x % p.C, t = blockIdx.x / p.C;
  const int a = t / p.k, b = t - a * p.k;
  const int64_t M = (int64_t)p.N * p.Ho * p.Wo;
  double acc = 0.0;
  for (int64_t m = threadIdx.x; m < M; m += 256) {
    const int wo = (int)(m % p.Wo);
    const int ho = (int)((m / p.Wo) % p.Ho);
    const int64_t n = m / ((int64_t)p.Wo * p.Ho);
    const int hi = ho * p.stride - p.pad_h + a * p.dil, wi = wo * p.stride - p.pad_w + b * p.dil;
    if (hi < 0 || hi >= p.H || wi < 0 || wi >= p.W) continue;
    acc += (double)dy[m * lddy + c] * (double)x[((n * p.H + hi) * (int64_t)p.W + wi) * ldx + c];
  }
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) dw[t * p.C + c] = (float)red[0];
}

// y[n][c] = scale * sum over the map;  block = (n, 256-channel slab), thread = channel, fixed order over pixels
__global__ void spatial_reduce_kernel(const float* __restrict__ x, int ldx, int HW, int C, double scale, float* __restrict__ y, int ldy) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x, n = blockIdx.y;
  if (c >= C) return;
  double acc = 0.0;
  for (int p = 0; p < HW; ++p) acc += (double)x[((int64_t)n * HW + p) * ldx + c];
  y[(int64_t)n * ldy + c] = (float)(acc * scale);
}
__global__ void spatial_bcast_kernel(const float* __restrict__ x, int ldx, int HW, int C, float scale, const float* __restrict__ res,
                                     int ldr, float* __restrict__ y, int ldy, int64_t total) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const int64_t r = i / C;
    float v = x[(r / HW) * ldx + c] * scale;
    if (res) v += res[r * ldr + c];
    y[r * ldy + c] = v;
  }
}

__device__ __forceinline__ uint32_t philox_word(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, int word) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  return word == 0 ? c0 : word == 1 ? c1 : word == 2 ? c2 : c3;
}
// the mask of csrc/dropout.cu: element (row, ch) -> octet i = row*(C/8) + ch/8, half = (ch%8)/4, word = ch%4
__global__ void dropout_kernel(const float* __restrict__ x, int ldx, float* __restrict__ y, int ldy, int64_t total, int C, uint32_t thresh,
                               float scale, uint64_t seed, uint32_t salt, const int64_t* __restrict__ d_step) {
  const int64_t step = *d_step;
  const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32) ^ salt;
  const int cv = C / 8;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int ch = (int)(e % C);
    const int64_t r = e / C;
    const uint64_t i = (uint64_t)r * cv + ch / 8;
    const uint32_t wv = philox_word((uint32_t)step, (uint32_t)i, ((uint32_t)(i >> 32) << 1) | (uint32_t)((ch % 8) / 4),
                                    (uint32_t)(step >> 32), k0, k1, ch % 4);
    y[r * ldy + ch] = wv >= thresh ? x[r * ldx + ch] * scale : 0.f;
  }
}

// AveragePooling2D(k, k) over exact windows (PSPNet pyramid) and its gradient
__global__ void avgpool_fwd_kernel(const float* __restrict__ x, int ldx, int H, int W, int k, float* __restrict__ y, int ldy, int Ho, int Wo,
                                   int C, int64_t total) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const int64_t r = i / C;
    const int wo = (int)(r % Wo), ho = (int)((r / Wo) % Ho);
    const int64_t n = r / ((int64_t)Wo * Ho);
    double acc = 0.0;
    for (int a = 0; a < k; ++a)
      for (int b = 0; b < k; ++b) acc += (double)x[((n * H + ho * k + a) * (int64_t)W + wo * k + b) * ldx + c];
    y[r * ldy + c] = (float)(acc / (double)(k * k));
  }
}
__global__ void avgpool_bwd_kernel(const float* __restrict__ dy, int lddy, int Ho, int Wo, int k, const float* __restrict__ res, int ldr,
                                   float* __restrict__ dx, int lddx, int H, int W, int C, int64_t total) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const int64_t r = i / C;
    const int w = (int)(r % W), h = (int)((r / W) % H);
    const int64_t n = r / ((int64_t)W * H);
    float g = dy[((n * Ho + h / k) * (int64_t)Wo + w / k) * lddy + c] / (float)(k * k);
    if (res) g += res[r * ldr + c];
    dx[r * lddx + c] = g;
  }
}

DwF make(const stp_dwconv_desc* d, const stp_tensor* x, const stp_tensor* y) {
  DwF p;
  p.N = x->n; p.H = x->h; p.W = x->w; p.C = x->c; p.Ho = y->h; p.Wo = y->w;
  p.k = d->k; p.stride = d->stride; p.dil = d->dilation; p.pad_h = d->pad_h; p.pad_w = d->pad_w;
  return p;
}

}  // namespace

int dwconv_fwd(const stp_dwconv_desc* d, const stp_tensor* x, const float* w, const stp_tensor* y, cudaStream_t st) {
  dw_fwd_kernel<<<grid1(pixels(y) * x->c), 256, 0, st>>>(make(d, x, y), (const float*)x->ptr, x->ld, w, (float*)y->ptr, y->ld);
  return check_launch("dwconv_fwd (fp32)");
}
int dwconv_dgrad(const stp_dwconv_desc* d, const stp_tensor* dy, const float* w, const stp_tensor* res, const stp_tensor* dx, cudaStream_t st) {
  dw_dgrad_kernel<<<grid1(pixels(dx) * dx->c), 256, 0, st>>>(make(d, dx, dy), (const float*)dy->ptr, dy->ld, w,
                                                           res ? (const float*)res->ptr : nullptr, res ? res->ld : 0, (float*)dx->ptr, dx->ld);
  return check_launch("dwconv_dgrad (fp32)");
}
int dwconv_wgrad(const stp_dwconv_desc* d, const stp_tensor* x, const stp_tensor* dy, float* dw, cudaStream_t st) {
  dw_wgrad_kernel<<<d->k * d->k * x->c, 256, 0, st>>>(make(d, x, dy), (const float*)x->ptr, x->ld, (const float*)dy->ptr, dy->ld, dw);
  return check_launch("dwconv_wgrad (fp32)");
}
int spatial_reduce(const stp_tensor* x, double scale, const stp_tensor* y, cudaStream_t st) {
  dim3 grid((x->c + 255) / 256, x->n);
  spatial_reduce_kernel<<<grid, 256, 0, st>>>((const float*)x->ptr, x->ld, x->h * x->w, x->c, scale, (float*)y->ptr, y->ld);
  return check_launch("spatial_reduce (fp32)");
}
int spatial_bcast(const stp_tensor* x, float scale, const stp_tensor* res, const stp_tensor* y, cudaStream_t st) {
  const int64_t total = pixels(y) * y->c;
  spatial_bcast_kernel<<<grid1(total), 256, 0, st>>>((const float*)x->ptr, x->ld, y->h * y->w, y->c, scale,
                                                    res ? (const float*)res->ptr : nullptr, res ? res->ld : 0, (float*)y->ptr, y->ld, total);
  return check_launch("spatial_bcast (fp32)");
}
int avgpool_fwd(const stp_tensor* x, int k, const stp_tensor* y, cudaStream_t st) {
  const int64_t total = pixels(y) * y->c;
  avgpool_fwd_kernel<<<grid1(total), 256, 0, st>>>((const float*)x->ptr, x->ld, x->h, x->w, k, (float*)y->ptr, y->ld, y->h, y->w, y->c, total);
  return check_launch("avgpool_fwd (fp32)");
}
int avgpool_bwd(const stp_tensor* dy, int k, const stp_tensor* res, const stp_tensor* dx, cudaStream_t st) {
  const int64_t total = pixels(dx) * dx->c;
  avgpool_bwd_kernel<<<grid1(total), 256, 0, st>>>((const float*)dy->ptr, dy->ld, dy->h, dy->w, k, res ? (const float*)res->ptr : nullptr,
                                                  res ? res->ld : 0, (float*)dx->ptr, dx->ld, dx->h, dx->w, dx->c, total);
  return check_launch("avgpool_bwd (fp32)");
}
int dropout(const stp_tensor* x, uint32_t thresh, float scale, uint64_t seed, uint32_t salt, const int64_t* d_step, const stp_tensor* y,
            cudaStream_t st) {
  const int64_t total = pixels(x) * x->c;
  dropout_kernel<<<grid1(total), 256, 0, st>>>((const float*)x->ptr, x->ld, (float*)y->ptr, y->ld, total, x->c, thresh, scale, seed, salt, d_step);
  return check_launch("dropout (fp32)");
}

}  // namespace f32
}  // namespace stp
