// K8b: bilinear resize (TF1 legacy tf.image.resize_bilinear, align_corners=False, NO half-pixel offset) forward and
// backward, and the gradient of UpSampling2D(2, nearest) for tensors that are not post-ReLU.
//
// Replaces keras UpSampling2D(interpolation='bilinear') -> tf.image.resize_bilinear in the FPN decoder that
// segmentation_models.FPN builds for the reference (segmentation.py:109-113; schema segmentation.raml:179-204: the
// x8/x4/x2 upsampling of the segmentation branches and the x4 `last_upsample` of the logits) [DEP tensorflow==1.15]:
//     scale = in/out (fp32);  src = dst*scale;  lo = floor(src);  hi = min(lo+1, in-1);  t = src - lo
//     top = tl + (tr - tl)*tx;  bot = bl + (br - bl)*tx;  out = top + (bot - top)*ty
// HBM-bound streaming kernels: bf16 NHWC with 16-byte (8-channel) vectors, thread = (pixel, channel octet), so a warp
// covers consecutive channel octets of consecutive pixels (coalesced on both the strided concat-slice destination and
// the source).  The backward is a deterministic GATHER (thread = source pixel; loops over the destination pixels whose
// lo/hi hit it, recomputing exactly the forward's fp32 index arithmetic) -- no atomics.
#include "common.cuh"

namespace stp {
namespace {

struct Axis {
  int lo, hi;
  float t;
};
__device__ __forceinline__ Axis src_of(int d, float scale, int in) {
  Axis a;
  const float s = (float)d * scale;
  const float f = floorf(s);
  a.lo = (int)f;
  if (a.lo > in - 1) a.lo = in - 1;
  a.hi = a.lo + 1 < in ? a.lo + 1 : in - 1;
  a.t = s - f;
  return a;
}
// weight with which destination index d reads source index i along one axis
__device__ __forceinline__ float axis_weight(int d, int i, float scale, int in) {
  const Axis a = src_of(d, scale, in);
  float w = 0.f;
  if (a.lo == i) w += 1.f - a.t;
  if (a.hi == i) w += a.t;
  return w;
}
// destination indices that can touch source index i: lo(d) in {i-1, i}  (one extra on each side for fp32 rounding)
__device__ __forceinline__ void cand_range(int i, float inv_scale, int out, int& d0, int& d1) {
  d0 = (int)floorf((float)(i - 1) * inv_scale) - 1;
  d1 = (int)ceilf((float)(i + 1) * inv_scale) + 1;
  if (d0 < 0) d0 = 0;
  if (d1 > out - 1) d1 = out - 1;
}

__global__ void __launch_bounds__(256) resize_fwd_bf16_kernel(const __nv_bfloat16* __restrict__ x, int ldx, int h, int w,
                                                               __nv_bfloat16* __restrict__ y, int ldy, int H, int W,
                                                               int64_t total, int cv, float sy, float sx) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / cv;
    const int v = (int)(i - r * cv);
    const int64_t n = r / ((int64_t)H * W);
    const int rem = (int)(r - n * (int64_t)H * W);
    const int oy = rem / W, ox = rem - oy * W;
    const Axis ay = src_of(oy, sy, h), ax = src_of(ox, sx, w);
    const __nv_bfloat16* base = x + n * (int64_t)h * w * ldx + v * 8;
    float tl[8], tr[8], bl[8], br[8], o[8];
    unpack8(ld8(base + ((int64_t)ay.lo * w + ax.lo) * ldx), tl);
    unpack8(ld8(base + ((int64_t)ay.lo * w + ax.hi) * ldx), tr);
    unpack8(ld8(base + ((int64_t)ay.hi * w + ax.lo) * ldx), bl);
    unpack8(ld8(base + ((int64_t)ay.hi * w + ax.hi) * ldx), br);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float top = tl[k] + (tr[k] - tl[k]) * ax.t;
      const float bot = bl[k] + (br[k] - bl[k]) * ax.t;
      o[k] = top + (bot - top) * ay.t;
    }
    st8(y + r * ldy + v * 8, pack8(o));
  }
}

// f32 [n,h,w,(ldx)] -> f32 [n,H,W,cy] (first cy channels; the padded logits of the FPN head -> dense [pixels][classes])
__global__ void __launch_bounds__(256) resize_fwd_f32_kernel(const float* __restrict__ x, int ldx, int h, int w,
                                                              float* __restrict__ y, int ldy, int H, int W, int64_t total,
                                                              int cy, float sy, float sx) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / cy;
    const int c = (int)(i - r * cy);
    const int64_t n = r / ((int64_t)H * W);
    const int rem = (int)(r - n * (int64_t)H * W);
    const int oy = rem / W, ox = rem - oy * W;
    const Axis ay = src_of(oy, sy, h), ax = src_of(ox, sx, w);
    const float* base = x + n * (int64_t)h * w * ldx + c;
    const float tl = base[((int64_t)ay.lo * w + ax.lo) * ldx], tr = base[((int64_t)ay.lo * w + ax.hi) * ldx];
    const float bl = base[((int64_t)ay.hi * w + ax.lo) * ldx], br = base[((int64_t)ay.hi * w + ax.hi) * ldx];
    const float top = tl + (tr - tl) * ax.t;
    const float bot = bl + (br - bl) * ax.t;
    y[r * ldy + c] = top + (bot - top) * ay.t;
  }
}

__global__ void __launch_bounds__(256) resize_bwd_bf16_kernel(const __nv_bfloat16* __restrict__ dy, int lddy, int H, int W,
                                                               const __nv_bfloat16* __restrict__ res, int ldr,
                                                               __nv_bfloat16* __restrict__ dx, int lddx, int h, int w,
                                                               int64_t total, int cv, float sy, float sx, float isy,
                                                               float isx) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / cv;
    const int v = (int)(i - r * cv);
    const int64_t n = r / ((int64_t)h * w);
    const int rem = (int)(r - n * (int64_t)h * w);
    const int iy = rem / w, ix = rem - iy * w;
    int y0, y1, x0, x1;
    cand_range(iy, isy, H, y0, y1);
    cand_range(ix, isx, W, x0, x1);
    float acc[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] = 0.f;
    const __nv_bfloat16* base = dy + n * (int64_t)H * W * lddy + v * 8;
    for (int oy = y0; oy <= y1; ++oy) {
      const float wy = axis_weight(oy, iy, sy, h);
      if (wy == 0.f) continue;
      for (int ox = x0; ox <= x1; ++ox) {
        const float wx = axis_weight(ox, ix, sx, w);
        if (wx == 0.f) continue;
        float g[8];
        unpack8(ld8(base + ((int64_t)oy * W + ox) * lddy), g);
        const float ww = wy * wx;
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[k] += ww * g[k];
      }
    }
    if (res) {
      float f[8];
      unpack8(ld8(res + r * ldr + v * 8), f);
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[k] += f[k];
    }
    st8(dx + r * lddx + v * 8, pack8(acc));
  }
}

// dy f32 [n,H,W,cy] dense -> dx bf16 [n,h,w,cx] (channels >= cy zero: the padded logit channels of the FPN head)
__global__ void __launch_bounds__(256) resize_bwd_f32_kernel(const float* __restrict__ dy, int lddy, int H, int W, int cy,
                                                              __nv_bfloat16* __restrict__ dx, int lddx, int h, int w,
                                                              int64_t total, int cx, float sy, float sx, float isy,
                                                              float isx) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / cx;
    const int c = (int)(i - r * cx);
    float acc = 0.f;
    if (c < cy) {
      const int64_t n = r / ((int64_t)h * w);
      const int rem = (int)(r - n * (int64_t)h * w);
      const int iy = rem / w, ix = rem - iy * w;
      int y0, y1, x0, x1;
      cand_range(iy, isy, H, y0, y1);
      cand_range(ix, isx, W, x0, x1);
      const float* base = dy + n * (int64_t)H * W * lddy + c;
      for (int oy = y0; oy <= y1; ++oy) {
        const float wy = axis_weight(oy, iy, sy, h);
        if (wy == 0.f) continue;
        for (int ox = x0; ox <= x1; ++ox) {
          const float wx = axis_weight(ox, ix, sx, w);
          if (wx == 0.f) continue;
          acc += wy * wx * base[((int64_t)oy * W + ox) * lddy];
        }
      }
    }
    dx[r * lddx + c] = __float2bfloat16_rn(acc);
  }
}

// parity mode: dy f32 [n,H,W,c] -> dx f32 [n,h,w,c] (+ residual), double accumulation in the forward's index arithmetic
__global__ void __launch_bounds__(256) resize_bwd_f32f32_kernel(const float* __restrict__ dy, int lddy, int H, int W, const float* __restrict__ res,
                                                                 int ldr, float* __restrict__ dx, int lddx, int h, int w, int64_t total, int c,
                                                                 int cy, float sy, float sx, float isy, float isx) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / c;
    const int ch = (int)(i - r * c);
    if (ch >= cy) {   // padded channels of the head conv output carry no gradient
      dx[r * lddx + ch] = res ? res[r * ldr + ch] : 0.f;
      continue;
    }
    const int64_t n = r / ((int64_t)h * w);
    const int rem = (int)(r - n * (int64_t)h * w);
    const int iy = rem / w, ix = rem - iy * w;
    int y0, y1, x0, x1;
    cand_range(iy, isy, H, y0, y1);
    cand_range(ix, isx, W, x0, x1);
    const float* base = dy + n * (int64_t)H * W * lddy + ch;
    double acc = 0.0;
    for (int oy = y0; oy <= y1; ++oy) {
      const float wy = axis_weight(oy, iy, sy, h);
      if (wy == 0.f) continue;
      for (int ox = x0; ox <= x1; ++ox) {
        const float wx = axis_weight(ox, ix, sx, w);
        if (wx == 0.f) continue;
        acc += (double)wy * (double)wx * (double)base[((int64_t)oy * W + ox) * lddy];
      }
    }
    if (res) acc += (double)res[r * ldr + ch];
    dx[r * lddx + ch] = (float)acc;
  }
}

// gradient of UpSampling2D(2): dx = 2x2 sum of dy (+ residual), no activation mask
__global__ void __launch_bounds__(256) upsample2x_bwd_kernel(const __nv_bfloat16* __restrict__ dy, int lddy,
                                                              const __nv_bfloat16* __restrict__ res, int ldr,
                                                              __nv_bfloat16* __restrict__ dx, int lddx, int h, int w,
                                                              int64_t total, int cv) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / cv;
    const int v = (int)(i - r * cv);
    const int64_t n = r / ((int64_t)h * w);
    const int rem = (int)(r - n * (int64_t)h * w);
    const int iy = rem / w, ix = rem - iy * w;
    const __nv_bfloat16* base = dy + ((n * 2 * h + 2 * iy) * (int64_t)(2 * w) + 2 * ix) * lddy + v * 8;
    const bf16x8 a0 = ld8(base), a1 = ld8(base + lddy), a2 = ld8(base + (int64_t)2 * w * lddy),
                 a3 = ld8(base + (int64_t)2 * w * lddy + lddy);
    float g[8], t[8];
    unpack8(a0, g);
    unpack8(a1, t);
#pragma unroll
    for (int k = 0; k < 8; ++k) g[k] += t[k];
    unpack8(a2, t);
#pragma unroll
    for (int k = 0; k < 8; ++k) g[k] += t[k];
    unpack8(a3, t);
#pragma unroll
    for (int k = 0; k < 8; ++k) g[k] += t[k];
    if (res) {
      unpack8(ld8(res + r * ldr + v * 8), t);
#pragma unroll
      for (int k = 0; k < 8; ++k) g[k] += t[k];
    }
    st8(dx + r * lddx + v * 8, pack8(g));
  }
}

// ---- probability head of the reference's DeepLabV3+ (impl/deeplab/model.py:494-500): Conv2D(classes, 1x1, activation)
// at 1/8 resolution, THEN BilinearUpsampling(output_size = input size) with align_corners=True (:81-83) -- the network's
// output is the bilinear blend of low-resolution PROBABILITIES.  The loss / predict kernels of this library consume logits
// and apply the activation themselves, so the forward emits  l = log(p/(1-p))  (sigmoid; p clipped to [1e-7, 1-1e-7] exactly
// where keras binary_crossentropy clips before turning its probability input back into logits)  or  l = log(p)  (softmax:
// softmax(log p) = p because the blended probability vectors still sum to 1).  The backward undoes that map
// (dp = dl / (p(1-p)) resp. dl / p), gathers through the resize like resize_bwd_f32_kernel and applies the activation
// derivative at low resolution; the result is the bf16 gradient of the padded 1x1 conv output.
constexpr int kProbMaxClasses = 16;
constexpr float kProbEps = 1e-7f;
__device__ __forceinline__ float sigmoidf_(float z) { return 1.f / (1.f + __expf(-z)); }

__global__ void __launch_bounds__(256) prob_head_fwd_kernel(const float* __restrict__ z, int ldz, int h, int w, int classes, int act,
                                                             float* __restrict__ logits, int H, int W, int64_t total, float sy, float sx) {
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < total; r += (int64_t)gridDim.x * blockDim.x) {
    const int64_t n = r / ((int64_t)H * W);
    const int rem = (int)(r - n * (int64_t)H * W);
    const int oy = rem / W, ox = rem - oy * W;
    const Axis ay = src_of(oy, sy, h), ax = src_of(ox, sx, w);
    const float* base = z + n * (int64_t)h * w * ldz;
    const float* c4[4] = {base + ((int64_t)ay.lo * w + ax.lo) * ldz, base + ((int64_t)ay.lo * w + ax.hi) * ldz,
                          base + ((int64_t)ay.hi * w + ax.lo) * ldz, base + ((int64_t)ay.hi * w + ax.hi) * ldz};
    float mx[4], inv[4];
    if (act == 2) {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        float m = -INFINITY, s = 0.f;
        for (int c = 0; c < classes; ++c) m = fmaxf(m, c4[q][c]);
        for (int c = 0; c < classes; ++c) s += __expf(c4[q][c] - m);
        mx[q] = m; inv[q] = 1.f / s;
      }
    }
    for (int c = 0; c < classes; ++c) {
      float pq[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) pq[q] = act == 2 ? __expf(c4[q][c] - mx[q]) * inv[q] : sigmoidf_(c4[q][c]);
      const float top = pq[0] + (pq[1] - pq[0]) * ax.t;
      const float bot = pq[2] + (pq[3] - pq[2]) * ax.t;
      float p = top + (bot - top) * ay.t;
      if (act == 2) {
        logits[r * classes + c] = __logf(fmaxf(p, kProbEps));
      } else {
        p = fminf(fmaxf(p, kProbEps), 1.f - kProbEps);
        logits[r * classes + c] = __logf(p / (1.f - p));
      }
    }
  }
}

// thread = (low-resolution pixel); dz bf16 [n,h,w,cx] (channels >= classes zero)
__device__ __forceinline__ void store_dz(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }
__device__ __forceinline__ void store_dz(float* p, float v) { *p = v; }

template <typename TZ>
__global__ void __launch_bounds__(128) prob_head_bwd_kernel(const float* __restrict__ dlogits, const float* __restrict__ logits, int H, int W,
                                                             int classes, int act, const float* __restrict__ z, int ldz,
                                                             TZ* __restrict__ dz, int lddz, int cx, int h, int w, int64_t total,
                                                             float sy, float sx, float isy, float isx) {
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < total; r += (int64_t)gridDim.x * blockDim.x) {
    const int64_t n = r / ((int64_t)h * w);
    const int rem = (int)(r - n * (int64_t)h * w);
    const int iy = rem / w, ix = rem - iy * w;
    int y0, y1, x0, x1;
    cand_range(iy, isy, H, y0, y1);
    cand_range(ix, isx, W, x0, x1);
    float acc[kProbMaxClasses];
#pragma unroll
    for (int c = 0; c < kProbMaxClasses; ++c) acc[c] = 0.f;
    for (int oy = y0; oy <= y1; ++oy) {
      const float wy = axis_weight(oy, iy, sy, h);
      if (wy == 0.f) continue;
      for (int ox = x0; ox <= x1; ++ox) {
        const float wx = axis_weight(ox, ix, sx, w);
        if (wx == 0.f) continue;
        const int64_t o = ((n * H + oy) * (int64_t)W + ox) * classes;
        const float ww = wy * wx;
#pragma unroll
        for (int c = 0; c < kProbMaxClasses; ++c) {
          if (c >= classes) break;
          const float l = logits[o + c], g = dlogits[o + c];
          float dp;
          if (act == 2) {
            const float p = __expf(l);
            dp = p > kProbEps ? g / p : 0.f;       // clipped probabilities carry no gradient
          } else {
            const float p = sigmoidf_(l);
            const float q = p * (1.f - p);
            dp = (p > kProbEps * 1.0001f && p < 1.f - kProbEps * 1.0001f) ? g / q : 0.f;
          }
          acc[c] += ww * dp;
        }
      }
    }
    const float* zr = z + r * ldz;
    TZ* out = dz + r * lddz;
    if (act == 2) {
      float m = -INFINITY, s = 0.f, dot = 0.f;
      for (int c = 0; c < classes; ++c) m = fmaxf(m, zr[c]);
      for (int c = 0; c < classes; ++c) s += __expf(zr[c] - m);
      const float inv = 1.f / s;
#pragma unroll
      for (int c = 0; c < kProbMaxClasses; ++c)
        if (c < classes) dot += acc[c] * __expf(zr[c] - m) * inv;
#pragma unroll
      for (int c = 0; c < kProbMaxClasses; ++c)
        if (c < classes) store_dz(out + c, __expf(zr[c] - m) * inv * (acc[c] - dot));
    } else {
#pragma unroll
      for (int c = 0; c < kProbMaxClasses; ++c)
        if (c < classes) {
          const float sg = sigmoidf_(zr[c]);
          store_dz(out + c, acc[c] * sg * (1.f - sg));
        }
    }
    for (int c = classes; c < cx; ++c) store_dz(out + c, 0.f);
  }
}

// parity mode: f32 2x2 sum (+ residual)
__global__ void __launch_bounds__(256) upsample2x_bwd_f32_kernel(const float* __restrict__ dy, int lddy, const float* __restrict__ res, int ldr,
                                                                  float* __restrict__ dx, int lddx, int h, int w, int64_t total, int c) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / c;
    const int ch = (int)(i - r * c);
    const int64_t n = r / ((int64_t)h * w);
    const int rem = (int)(r - n * (int64_t)h * w);
    const int iy = rem / w, ix = rem - iy * w;
    const float* base = dy + ((n * 2 * h + 2 * iy) * (int64_t)(2 * w) + 2 * ix) * lddy + ch;
    double g = (double)base[0] + (double)base[lddy] + (double)base[(int64_t)2 * w * lddy] + (double)base[(int64_t)2 * w * lddy + lddy];
    if (res) g += (double)res[r * ldr + ch];
    dx[r * lddx + ch] = (float)g;
  }
}

int grid_for(int64_t total) {
  int64_t b = (total + 255) / 256;
  const int64_t cap = (int64_t)kNumSMs * 16;
  return (int)(b < cap ? (b < 1 ? 1 : b) : cap);
}

bool f32_ok(const stp_tensor* t) { return t && t->ptr && t->dtype == STP_F32 && t->ld >= t->c && t->c >= 1; }

}  // namespace
}  // namespace stp

using namespace stp;

static void align_scales(int in_h, int in_w, int out_h, int out_w, float& sy, float& sx, float& isy, float& isx) {
  // tf.image.resize_bilinear(align_corners=True): scale = (in-1)/(out-1) when out > 1 [DEP tensorflow==1.15 CalculateResizeScale]
  sy = out_h > 1 ? (float)(in_h - 1) / (float)(out_h - 1) : (float)in_h / (float)out_h;
  sx = out_w > 1 ? (float)(in_w - 1) / (float)(out_w - 1) : (float)in_w / (float)out_w;
  isy = in_h > 1 ? (float)(out_h - 1) / (float)(in_h - 1) : (float)out_h;
  isx = in_w > 1 ? (float)(out_w - 1) / (float)(in_w - 1) : (float)out_w;
}

static int resize_fwd_impl(const stp_tensor* x, const stp_tensor* y, bool align, stp_stream stream) {
  STP_REQUIRE(x && y && x->n == y->n && x->h >= 1 && x->w >= 1 && y->h >= 1 && y->w >= 1, "resize_bilinear_fwd: bad shapes");
  float sy = (float)((double)x->h / (double)y->h), sx = (float)((double)x->w / (double)y->w), isy_, isx_;
  if (align) align_scales(x->h, x->w, y->h, y->w, sy, sx, isy_, isx_);
  cudaStream_t st = (cudaStream_t)stream;
  if (x->dtype == STP_BF16) {
    STP_REQUIRE(vec_ok(x) && vec_ok(y) && x->c == y->c, "resize_bilinear_fwd: bf16 tensors must have equal channel counts");
    const int cv = x->c / 8;
    const int64_t total = pixels(y) * cv;
    resize_fwd_bf16_kernel<<<grid_for(total), 256, 0, st>>>((const __nv_bfloat16*)x->ptr, x->ld, x->h, x->w,
                                                            (__nv_bfloat16*)y->ptr, y->ld, y->h, y->w, total, cv, sy, sx);
  } else {
    STP_REQUIRE(f32_ok(x) && f32_ok(y) && y->c <= x->c, "resize_bilinear_fwd: f32 tensors need y.c <= x.c");
    const int64_t total = pixels(y) * y->c;
    resize_fwd_f32_kernel<<<grid_for(total), 256, 0, st>>>((const float*)x->ptr, x->ld, x->h, x->w, (float*)y->ptr, y->ld,
                                                           y->h, y->w, total, y->c, sy, sx);
  }
  return check_launch("resize_bilinear_fwd");
}
extern "C" int stp_resize_bilinear_fwd(const stp_tensor* x, const stp_tensor* y, stp_stream stream) {
  return resize_fwd_impl(x, y, false, stream);
}
extern "C" int stp_resize_bilinear_ac_fwd(const stp_tensor* x, const stp_tensor* y, stp_stream stream) {
  return resize_fwd_impl(x, y, true, stream);
}

static int resize_bwd_impl(const stp_tensor* dy, const stp_tensor* residual, const stp_tensor* dx, bool align, stp_stream stream) {
  STP_REQUIRE(dy && dx && dy->n == dx->n && (vec_ok(dx) || (dx->dtype == STP_F32 && f32_ok(dx))), "resize_bilinear_bwd: bad tensors");
  STP_REQUIRE(dy->h >= dx->h && dy->w >= dx->w, "resize_bilinear_bwd: only up-scaling resizes have a gather backward here");
  float sy = (float)((double)dx->h / (double)dy->h), sx = (float)((double)dx->w / (double)dy->w);
  float isy = (float)((double)dy->h / (double)dx->h), isx = (float)((double)dy->w / (double)dx->w);
  if (align) align_scales(dx->h, dx->w, dy->h, dy->w, sy, sx, isy, isx);
  cudaStream_t st = (cudaStream_t)stream;
  if (dy->dtype == STP_BF16) {
    STP_REQUIRE(vec_ok(dy) && dy->c == dx->c, "resize_bilinear_bwd: bf16 tensors must have equal channel counts");
    if (residual) STP_REQUIRE(vec_ok(residual) && residual->c == dx->c && pixels(residual) == pixels(dx), "resize_bilinear_bwd: bad residual");
    const int cv = dx->c / 8;
    const int64_t total = pixels(dx) * cv;
    resize_bwd_bf16_kernel<<<grid_for(total), 256, 0, st>>>(
        (const __nv_bfloat16*)dy->ptr, dy->ld, dy->h, dy->w, residual ? (const __nv_bfloat16*)residual->ptr : nullptr,
        residual ? residual->ld : 0, (__nv_bfloat16*)dx->ptr, dx->ld, dx->h, dx->w, total, cv, sy, sx, isy, isx);
  } else if (dx->dtype == STP_F32) {   // parity mode
    STP_REQUIRE(f32_ok(dy) && f32_ok(dx) && dy->c <= dx->c, "resize_bilinear_bwd (fp32): dy.c <= dx.c");
    if (residual) STP_REQUIRE(f32_ok(residual) && residual->c == dx->c && pixels(residual) == pixels(dx), "resize_bilinear_bwd (fp32): bad residual");
    const int64_t total = pixels(dx) * dx->c;
    resize_bwd_f32f32_kernel<<<grid_for(total), 256, 0, st>>>((const float*)dy->ptr, dy->ld, dy->h, dy->w,
                                                              residual ? (const float*)residual->ptr : nullptr, residual ? residual->ld : 0,
                                                              (float*)dx->ptr, dx->ld, dx->h, dx->w, total, dx->c, dy->c, sy, sx, isy, isx);
  } else {
    STP_REQUIRE(f32_ok(dy) && dy->c <= dx->c && !residual, "resize_bilinear_bwd: f32 dy needs dy.c <= dx.c and no residual");
    const int64_t total = pixels(dx) * dx->c;
    resize_bwd_f32_kernel<<<grid_for(total), 256, 0, st>>>((const float*)dy->ptr, dy->ld, dy->h, dy->w, dy->c,
                                                           (__nv_bfloat16*)dx->ptr, dx->ld, dx->h, dx->w, total, dx->c, sy,
                                                           sx, isy, isx);
  }
  return check_launch("resize_bilinear_bwd");
}
extern "C" int stp_resize_bilinear_bwd(const stp_tensor* dy, const stp_tensor* residual, const stp_tensor* dx, stp_stream stream) {
  return resize_bwd_impl(dy, residual, dx, false, stream);
}
extern "C" int stp_resize_bilinear_ac_bwd(const stp_tensor* dy, const stp_tensor* residual, const stp_tensor* dx, stp_stream stream) {
  return resize_bwd_impl(dy, residual, dx, true, stream);
}

extern "C" int stp_upsample2x_bwd(const stp_tensor* dy, const stp_tensor* residual, const stp_tensor* dx,
                                  stp_stream stream) {
  if (dx && dx->dtype == STP_F32) {   // parity mode
    STP_REQUIRE(f32_ok(dy) && f32_ok(dx) && dy->c == dx->c && dy->n == dx->n && dy->h == 2 * dx->h && dy->w == 2 * dx->w &&
                    (!residual || (f32_ok(residual) && residual->c == dx->c && pixels(residual) == pixels(dx))),
                "upsample2x_bwd (fp32): shape mismatch");
    const int64_t total = pixels(dx) * dx->c;
    upsample2x_bwd_f32_kernel<<<grid_for(total), 256, 0, (cudaStream_t)stream>>>(
        (const float*)dy->ptr, dy->ld, residual ? (const float*)residual->ptr : nullptr, residual ? residual->ld : 0, (float*)dx->ptr,
        dx->ld, dx->h, dx->w, total, dx->c);
    return check_launch("upsample2x_bwd (fp32)");
  }
  STP_REQUIRE(vec_ok(dy) && vec_ok(dx) && dy->c == dx->c && dy->n == dx->n && dy->h == 2 * dx->h && dy->w == 2 * dx->w,
              "upsample2x_bwd: shape mismatch");
  if (residual) STP_REQUIRE(vec_ok(residual) && residual->c == dx->c && pixels(residual) == pixels(dx), "upsample2x_bwd: bad residual");
  const int cv = dx->c / 8;
  const int64_t total = pixels(dx) * cv;
  upsample2x_bwd_kernel<<<grid_for(total), 256, 0, (cudaStream_t)stream>>>(
      (const __nv_bfloat16*)dy->ptr, dy->ld, residual ? (const __nv_bfloat16*)residual->ptr : nullptr,
      residual ? residual->ld : 0, (__nv_bfloat16*)dx->ptr, dx->ld, dx->h, dx->w, total, cv);
  return check_launch("upsample2x_bwd");
}

extern "C" int stp_prob_head_fwd(const stp_tensor* z, int32_t classes, int32_t activation, const stp_tensor* logits, stp_stream stream) {
  STP_REQUIRE(z && logits && f32_ok(z) && f32_ok(logits) && z->n == logits->n, "prob_head_fwd: f32 tensors of one batch");
  STP_REQUIRE(classes >= 1 && classes <= kProbMaxClasses && classes <= z->c && logits->c == classes && logits->ld == classes,
              "prob_head_fwd: 1 <= classes <= 16 <= z.c; logits dense [n,H,W,classes]");
  STP_REQUIRE(activation == 1 || (activation == 2 && classes >= 2), "prob_head_fwd: activation 1 (sigmoid) or 2 (softmax, classes >= 2)");
  float sy, sx, isy, isx;
  align_scales(z->h, z->w, logits->h, logits->w, sy, sx, isy, isx);
  const int64_t total = pixels(logits);
  prob_head_fwd_kernel<<<grid_for(total), 256, 0, (cudaStream_t)stream>>>((const float*)z->ptr, z->ld, z->h, z->w, classes, activation,
                                                                          (float*)logits->ptr, logits->h, logits->w, total, sy, sx);
  return check_launch("prob_head_fwd");
}

extern "C" int stp_prob_head_bwd(const stp_tensor* dlogits, const stp_tensor* logits, const stp_tensor* z, int32_t classes,
                                 int32_t activation, const stp_tensor* dz, stp_stream stream) {
  STP_REQUIRE(dlogits && logits && z && dz && f32_ok(dlogits) && f32_ok(logits) && f32_ok(z), "prob_head_bwd: f32 dlogits / logits / z");
  STP_REQUIRE(dz->ptr && (dz->dtype == STP_BF16 || dz->dtype == STP_F32) && dz->ld >= dz->c && dz->c >= classes && pixels(dz) == pixels(z) && dz->n == z->n &&
                  dz->h == z->h && dz->w == z->w, "prob_head_bwd: dz bf16 (or f32 in parity mode) with z's geometry");
  STP_REQUIRE(classes >= 1 && classes <= kProbMaxClasses && classes <= z->c && logits->c == classes && logits->ld == classes &&
                  dlogits->c == classes && dlogits->ld == classes && pixels(dlogits) == pixels(logits) && logits->n == z->n,
              "prob_head_bwd: dense [n,H,W,classes] logits and gradient");
  STP_REQUIRE(logits->h >= z->h && logits->w >= z->w, "prob_head_bwd: up-scaling only");
  STP_REQUIRE(activation == 1 || (activation == 2 && classes >= 2), "prob_head_bwd: activation 1 (sigmoid) or 2 (softmax)");
  float sy, sx, isy, isx;
  align_scales(z->h, z->w, logits->h, logits->w, sy, sx, isy, isx);
  const int64_t total = pixels(z);
  const int64_t nb = (total + 127) / 128;
  if (dz->dtype == STP_F32)
    prob_head_bwd_kernel<float><<<(int)(nb < 1 ? 1 : nb), 128, 0, (cudaStream_t)stream>>>(
        (const float*)dlogits->ptr, (const float*)logits->ptr, logits->h, logits->w, classes, activation, (const float*)z->ptr, z->ld,
        (float*)dz->ptr, dz->ld, dz->c, z->h, z->w, total, sy, sx, isy, isx);
  else
    prob_head_bwd_kernel<__nv_bfloat16><<<(int)(nb < 1 ? 1 : nb), 128, 0, (cudaStream_t)stream>>>(
        (const float*)dlogits->ptr, (const float*)logits->ptr, logits->h, logits->w, classes, activation, (const float*)z->ptr, z->ld,
        (__nv_bfloat16*)dz->ptr, dz->ld, dz->c, z->h, z->w, total, sy, sx, isy, isx);
  return check_launch("prob_head_bwd");
}
