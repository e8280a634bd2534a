// K7: MaxPooling2D forward (+argmax) / backward.  HBM-bound, 16-byte vectors over channels.
// Semantics: ZeroPadding2D(pad) + MaxPooling2D(k, stride, 'valid') on post-ReLU data == max_pool(k, stride, pad)
// with the FIRST maximum in (kh, kw) scan order taking the gradient.
#include "common.cuh"
#include "conv.h"
#include "f32_path.h"

namespace stp {

__global__ void __launch_bounds__(256) maxpool_fwd_kernel(const __nv_bfloat16* __restrict__ x, int ldx, int H, int W,
                                                          int k, int stride, int pad, __nv_bfloat16* __restrict__ y,
                                                          int ldy, int Ho, int Wo, uint8_t* __restrict__ argmax,
                                                          int64_t rows_out, int C, int cv) {
  int64_t total = rows_out * cv;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t r = i / cv;
    int v = (int)(i - r * cv);
    int64_t n = r / ((int64_t)Ho * Wo);
    int rem = (int)(r - n * (int64_t)Ho * Wo);
    int ho = rem / Wo, wo = rem - ho * Wo;
    float best[8];
    int bi[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      best[c] = -INFINITY;
      bi[c] = 0;
    }
    for (int a = 0; a < k; ++a) {
      int hi = ho * stride - pad + a;
      for (int b = 0; b < k; ++b) {
        int wi = wo * stride - pad + b;
        float f[8];
        if (hi >= 0 && hi < H && wi >= 0 && wi < W) {
          unpack8(ld8(x + ((n * H + hi) * (int64_t)W + wi) * ldx + v * 8), f);
        } else {
#pragma unroll
          for (int c = 0; c < 8; ++c) f[c] = pad > 0 ? 0.f : -INFINITY;  // explicit ZeroPadding2D contributes zeros
        }
#pragma unroll
        for (int c = 0; c < 8; ++c)
          if (f[c] > best[c]) {
            best[c] = f[c];
            bi[c] = a * k + b;
          }
      }
    }
    st8(y + r * ldy + v * 8, pack8(best));
    if (argmax) {
      uint2 pk;
      pk.x = (uint32_t)bi[0] | ((uint32_t)bi[1] << 8) | ((uint32_t)bi[2] << 16) | ((uint32_t)bi[3] << 24);
      pk.y = (uint32_t)bi[4] | ((uint32_t)bi[5] << 8) | ((uint32_t)bi[6] << 16) | ((uint32_t)bi[7] << 24);
      *reinterpret_cast<uint2*>(argmax + r * C + v * 8) = pk;
    }
  }
}

// 3x3 / stride 2 / pad 1 (the ResNet stem pool) with packed bf16x2 compares: per tap and channel PAIR one HMNMX2, one compare
// mask and one LOP3 for the argmax index (the generic kernel above spends ~107 instructions per output channel on unpack /
// compare / select and is issue-bound: 56 M warp instructions, 69 % issue active, 85 us for 134 MB in --
// profiles/r2_ncu_full_tail_mpf.metrics.txt).  Same semantics: zero padding takes part in the max (explicit ZeroPadding2D),
// the first maximum in tap order wins (strict >).  The output is the selected bf16 value itself, bit exact.
__global__ void __launch_bounds__(256) maxpool_fwd_k3s2_kernel(const __nv_bfloat16* __restrict__ x, int ldx, int H, int W,
                                                               __nv_bfloat16* __restrict__ y, int ldy, int Ho, int Wo,
                                                               uint8_t* __restrict__ argmax, int total, int C, int cv) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int v = i % cv;
    int q = i / cv;
    const int wo = q % Wo;
    q /= Wo;
    const int ho = q % Ho;
    const int n = q / Ho;
    const __nv_bfloat16* xb = x + (int64_t)n * H * W * ldx + v * 8;
    uint32_t best[4], idx[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      best[c] = 0xFF80FF80u;   // (-inf, -inf)
      idx[c] = 0u;
    }
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const int hi = 2 * ho - 1 + a;
      const bool okh = (unsigned)hi < (unsigned)H;
#pragma unroll
      for (int b = 0; b < 3; ++b) {
        const int wi = 2 * wo - 1 + b;
        const bool ok = okh && (unsigned)wi < (unsigned)W;
        uint4 f = make_uint4(0u, 0u, 0u, 0u);   // padded taps are zeros
        if (ok) f = *reinterpret_cast<const uint4*>(xb + ((int64_t)hi * W + wi) * ldx);
        const uint32_t fw[4] = {f.x, f.y, f.z, f.w};
        const uint32_t tap = (uint32_t)(a * 3 + b) * 0x00010001u;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const __nv_bfloat162 fv = *reinterpret_cast<const __nv_bfloat162*>(&fw[c]);
          const __nv_bfloat162 bv = *reinterpret_cast<const __nv_bfloat162*>(&best[c]);
          const uint32_t m = __hgt2_mask(fv, bv);            // 0xFFFF per half where f > best
          best[c] = (fw[c] & m) | (best[c] & ~m);
          idx[c] = (tap & m) | (idx[c] & ~m);
        }
      }
    }
    const int64_t r = ((int64_t)n * Ho + ho) * Wo + wo;
    *reinterpret_cast<uint4*>(y + r * ldy + v * 8) = make_uint4(best[0], best[1], best[2], best[3]);
    if (argmax) {
      uint2 pk;
      pk.x = __byte_perm(idx[0], idx[1], 0x6420);
      pk.y = __byte_perm(idx[2], idx[3], 0x6420);
      *reinterpret_cast<uint2*>(argmax + r * C + v * 8) = pk;
    }
  }
}

// gather form: each input pixel looks at the <= ceil(k/stride)^2 windows covering it
__global__ void __launch_bounds__(256) maxpool_bwd_kernel(const __nv_bfloat16* __restrict__ dy, int lddy, int Ho,
                                                          int Wo, const uint8_t* __restrict__ argmax, int k,
                                                          int stride, int pad, const __nv_bfloat16* __restrict__ res,
                                                          int ldr, __nv_bfloat16* __restrict__ dx, int lddx, int H,
                                                          int W, int64_t rows_in, int C, int cv) {
  int64_t total = rows_in * cv;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t r = i / cv;
    int v = (int)(i - r * cv);
    int64_t n = r / ((int64_t)H * W);
    int rem = (int)(r - n * (int64_t)H * W);
    int h = rem / W, w = rem - h * W;
    float acc[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) acc[c] = 0.f;
    // windows ho with ho*stride - pad <= h <= ho*stride - pad + k - 1
    int ho_hi = (h + pad) / stride;
    int ho_lo = (h + pad - k + stride) / stride;  // ceil((h+pad-k+1)/stride) for non-negative numerators
    if (h + pad - k + 1 <= 0) ho_lo = 0;
    int wo_hi = (w + pad) / stride;
    int wo_lo = (w + pad - k + stride) / stride;
    if (w + pad - k + 1 <= 0) wo_lo = 0;
    for (int ho = ho_lo; ho <= ho_hi && ho < Ho; ++ho) {
      int a = h - (ho * stride - pad);
      for (int wo = wo_lo; wo <= wo_hi && wo < Wo; ++wo) {
        int b = w - (wo * stride - pad);
        int idx = a * k + b;
        int64_t ro = (n * Ho + ho) * (int64_t)Wo + wo;
        uint2 pk = *reinterpret_cast<const uint2*>(argmax + ro * C + v * 8);
        float g[8];
        unpack8(ld8(dy + ro * lddy + v * 8), g);
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          uint32_t word = c < 4 ? pk.x : pk.y;
          int am = (word >> (8 * (c & 3))) & 0xff;
          if (am == idx) acc[c] += g[c];
        }
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

// 3x3 / stride 2 / pad 1 (the ResNet stem pool): an input row is covered by one window (even rows, offset 1) or two
// (odd rows, offsets 2 and 0), likewise columns -> at most 4 windows, all loads issued before use; thread = fixed
// 8-channel vector walking pixels (32-bit index math).
__global__ void __launch_bounds__(256) maxpool_bwd_k3s2_kernel(const __nv_bfloat16* __restrict__ dy, int lddy, int Ho, int Wo,
                                                               const uint8_t* __restrict__ argmax,
                                                               const __nv_bfloat16* __restrict__ res, int ldr,
                                                               __nv_bfloat16* __restrict__ dx, int lddx, int H, int W,
                                                               int rows_in, int C, int cv, int ppi, int pix_per_blk) {
  const int v = threadIdx.x % cv, pl = threadIdx.x / cv;
  const int m_begin = blockIdx.x * pix_per_blk;
  int m_end = m_begin + pix_per_blk;
  if (m_end > rows_in) m_end = rows_in;
  // U pixels per iteration, all indices compile-time (registers, no local-memory arrays): up to U*4 argmax words + dy vectors
  // (+ U residual vectors) are in flight per thread before the first use
  constexpr int U = 2;
  for (int m0 = m_begin + pl; m0 < m_end; m0 += U * ppi) {
    uint2 pk[U][4];
    bf16x8 gq[U][4], rq[U];
    int idx[U][4];
    bool ok[U][4];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int m = m0 + u * ppi;
      const bool live = m < m_end;
      const unsigned n = (unsigned)m / (unsigned)(H * W);
      const unsigned rem = (unsigned)m - n * (unsigned)(H * W);
      const int h = (int)(rem / (unsigned)W), w = (int)(rem - (unsigned)h * (unsigned)W);
      // window candidates: an even row is covered by one window (offset 1), an odd row by two (offsets 2 and 0)
      const int ho0 = h >> 1, a0 = (h & 1) ? 2 : 1, ho1 = (h >> 1) + 1;
      const bool vh1 = (h & 1) && ho1 < Ho;
      const int wo0 = w >> 1, b0 = (w & 1) ? 2 : 1, wo1 = (w >> 1) + 1;
      const bool vw1 = (w & 1) && wo1 < Wo;
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const int ho = i ? ho1 : ho0, wo = j ? wo1 : wo0;
          const bool valid = live && (i ? vh1 : ho0 < Ho) && (j ? vw1 : wo0 < Wo);
          ok[u][i * 2 + j] = valid;
          idx[u][i * 2 + j] = (i ? 0 : a0) * 3 + (j ? 0 : b0);
          if (valid) {
            const int64_t ro = ((int64_t)n * Ho + ho) * Wo + wo;
            pk[u][i * 2 + j] = *reinterpret_cast<const uint2*>(argmax + ro * C + v * 8);
            gq[u][i * 2 + j] = ld8(dy + ro * lddy + v * 8);
          }
        }
      if (res && live) rq[u] = ld8(res + (int64_t)m * ldr + v * 8);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int m = m0 + u * ppi;
      if (m >= m_end) break;
      float acc[8];
#pragma unroll
      for (int c = 0; c < 8; ++c) acc[c] = 0.f;
#pragma unroll
      for (int q = 0; q < 4; ++q)
        if (ok[u][q]) {
          float g[8];
          unpack8(gq[u][q], g);
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const uint32_t word = c < 4 ? pk[u][q].x : pk[u][q].y;
            if ((int)((word >> (8 * (c & 3))) & 0xff) == idx[u][q]) acc[c] += g[c];
          }
        }
      if (res) {
        float rf[8];
        unpack8(rq[u], rf);
#pragma unroll
        for (int c = 0; c < 8; ++c) acc[c] += rf[c];
      }
      st8(dx + (int64_t)m * lddx + v * 8, pack8(acc));
    }
  }
}
// 3x3 / stride 2 / pad 1, even H and W: thread = (8-channel vector, 2x2 block of input pixels).  The block (2i..2i+1, 2j..2j+1)
// is covered by the windows (i, j) [all four pixels], (i+1, j) [odd row], (i, j+1) [odd column] and (i+1, j+1) [odd, odd]:
// four (argmax word, dy vector) loads serve four pixels (the per-pixel kernel above issues nine for them), no index divisions
// per pixel.  ncu on the per-pixel kernel: 108 registers, 24 % occupancy, 135 us for 134 MB in + 134 MB out + 50 MB of dy /
// argmax (profiles/r2_ncu_full_tail_mpb.metrics.txt).
__global__ void __launch_bounds__(256) maxpool_bwd_k3s2_block_kernel(const __nv_bfloat16* __restrict__ dy, int lddy, int Ho, int Wo,
                                                                     const uint8_t* __restrict__ argmax,
                                                                     const __nv_bfloat16* __restrict__ res, int ldr,
                                                                     __nv_bfloat16* __restrict__ dx, int lddx, int H, int W, int C, int cv,
                                                                     int total) {
  const int Wb = W / 2, Hb = H / 2;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x) {
    const int v = t % cv;
    int q = t / cv;
    const int j = q % Wb;
    q /= Wb;
    const int i = q % Hb;
    const int n = q / Hb;
    float g[4][8];
    uint2 pk[4];
    bool ok[4];
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int b = 0; b < 2; ++b) {
        const int ho = i + a, wo = j + b;
        const bool valid = ho < Ho && wo < Wo;
        ok[a * 2 + b] = valid;
        const int64_t ro = ((int64_t)n * Ho + (valid ? ho : 0)) * Wo + (valid ? wo : 0);
        uint2 w2 = make_uint2(0xffffffffu, 0xffffffffu);
        uint4 d4 = make_uint4(0u, 0u, 0u, 0u);
        if (valid) {
          w2 = *reinterpret_cast<const uint2*>(argmax + ro * C + v * 8);
          d4 = *reinterpret_cast<const uint4*>(dy + ro * lddy + v * 8);
        }
        pk[a * 2 + b] = w2;
        unpack8(*reinterpret_cast<bf16x8*>(&d4), g[a * 2 + b]);
      }
#pragma unroll
    for (int pa = 0; pa < 2; ++pa)
#pragma unroll
      for (int pb = 0; pb < 2; ++pb) {
        float acc[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) acc[c] = 0.f;
        // pixel (2i+pa, 2j+pb): window rows {i} (offset 1) for pa = 0, {i (offset 2), i+1 (offset 0)} for pa = 1; columns alike
#pragma unroll
        for (int a = 0; a <= pa; ++a)
#pragma unroll
          for (int b = 0; b <= pb; ++b) {
            const int roff = pa == 0 ? 1 : (a == 0 ? 2 : 0), coff = pb == 0 ? 1 : (b == 0 ? 2 : 0);
            const uint32_t idx = (uint32_t)(roff * 3 + coff);
            const uint2 w2 = pk[a * 2 + b];
#pragma unroll
            for (int c = 0; c < 8; ++c) {
              const uint32_t word = c < 4 ? w2.x : w2.y;
              if (((word >> (8 * (c & 3))) & 0xffu) == idx) acc[c] += g[a * 2 + b][c];
            }
          }
        const int64_t m = ((int64_t)n * H + 2 * i + pa) * W + 2 * j + pb;
        if (res) {
          float rf[8];
          unpack8(ld8(res + m * ldr + v * 8), rf);
#pragma unroll
          for (int c = 0; c < 8; ++c) acc[c] += rf[c];
        }
        st8(dx + m * lddx + v * 8, pack8(acc));
      }
  }
}

// AveragePooling2D(pool_size k, strides k) over exact windows (PSPNet pyramid pooling, schema segmentation.raml:226-248 ->
// segmentation_models PSPNet InterpBlock [DEP]): y = mean of the k x k window, fp32 accumulation.  Backward: every input
// pixel receives dy / k^2 of its window (+ residual: the feature map feeds four pyramid levels and the concat).
__global__ void __launch_bounds__(256) avgpool_fwd_kernel(const __nv_bfloat16* __restrict__ x, int ldx, int H, int W, int k,
                                                          __nv_bfloat16* __restrict__ y, int ldy, int Ho, int Wo, int64_t rows_out, int cv) {
  const int64_t total = rows_out * cv;
  const float inv = 1.f / (float)(k * k);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / cv;
    const int v = (int)(i - r * cv);
    const int64_t n = r / ((int64_t)Ho * Wo);
    const int rem = (int)(r - n * (int64_t)Ho * Wo);
    const int ho = rem / Wo, wo = rem - ho * Wo;
    float acc[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) acc[c] = 0.f;
    for (int a = 0; a < k; ++a)
      for (int b = 0; b < k; ++b) {
        float f[8];
        unpack8(ld8(x + ((n * H + ho * k + a) * (int64_t)W + wo * k + b) * ldx + v * 8), f);
#pragma unroll
        for (int c = 0; c < 8; ++c) acc[c] += f[c];
      }
#pragma unroll
    for (int c = 0; c < 8; ++c) acc[c] *= inv;
    st8(y + r * ldy + v * 8, pack8(acc));
  }
}
__global__ void __launch_bounds__(256) avgpool_bwd_kernel(const __nv_bfloat16* __restrict__ dy, int lddy, int Ho, int Wo, int k,
                                                          const __nv_bfloat16* __restrict__ res, int ldr, __nv_bfloat16* __restrict__ dx,
                                                          int lddx, int H, int W, int64_t rows_in, int cv) {
  const int64_t total = rows_in * cv;
  const float inv = 1.f / (float)(k * k);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / cv;
    const int v = (int)(i - r * cv);
    const int64_t n = r / ((int64_t)H * W);
    const int rem = (int)(r - n * (int64_t)H * W);
    const int h = rem / W, w = rem - h * W;
    float g[8];
    unpack8(ld8(dy + ((n * Ho + h / k) * (int64_t)Wo + w / k) * lddy + v * 8), g);
#pragma unroll
    for (int c = 0; c < 8; ++c) g[c] *= inv;
    if (res) {
      float rf[8];
      unpack8(ld8(res + r * ldr + v * 8), rf);
#pragma unroll
      for (int c = 0; c < 8; ++c) g[c] += rf[c];
    }
    st8(dx + r * lddx + v * 8, pack8(g));
  }
}

// ---- whole-map mean and its inverse (DeepLabV3+ image-pooling branch, impl/deeplab/model.py:462-469: AveragePooling2D over
// the whole 1/8 map -> 1x1 conv -> BN -> ReLU -> BilinearUpsampling back to the map, which from a single pixel is a broadcast).
// spatial_reduce: y[n,c] = scale * sum_{h,w} x[n,h,w,c];  block = (image, 8 channel octets) x 32 pixel lanes.
// spatial_bcast : y[n,h,w,c] = scale * x[n,c] (+ residual).
__global__ void __launch_bounds__(256) spatial_reduce_kernel(const __nv_bfloat16* __restrict__ x, int ldx, int HW, int cv, float scale,
                                                             __nv_bfloat16* __restrict__ y, int ldy) {
  __shared__ float sm[32][8][9];
  const int n = blockIdx.y;
  const int vl = threadIdx.x & 7, pl = threadIdx.x >> 3;
  const int v = blockIdx.x * 8 + vl;
  float acc[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) acc[c] = 0.f;
  if (v < cv)
    for (int p = pl; p < HW; p += 32) {
      float f[8];
      unpack8(ld8(x + ((int64_t)n * HW + p) * ldx + v * 8), f);
#pragma unroll
      for (int c = 0; c < 8; ++c) acc[c] += f[c];
    }
#pragma unroll
  for (int c = 0; c < 8; ++c) sm[pl][vl][c] = acc[c];
  __syncthreads();
  if (threadIdx.x < 64) {
    const int vv = threadIdx.x >> 3, c = threadIdx.x & 7;
    float s = 0.f;
    for (int l = 0; l < 32; ++l) s += sm[l][vv][c];
    if (blockIdx.x * 8 + vv < cv) y[(int64_t)n * ldy + (blockIdx.x * 8 + vv) * 8 + c] = __float2bfloat16_rn(s * scale);
  }
}
__global__ void __launch_bounds__(256) spatial_bcast_kernel(const __nv_bfloat16* __restrict__ x, int ldx, int HW, int cv, float scale,
                                                            const __nv_bfloat16* __restrict__ res, int ldr, __nv_bfloat16* __restrict__ y,
                                                            int ldy, int64_t total) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / cv;
    const int v = (int)(i - r * cv);
    const int64_t n = r / HW;
    float f[8];
    unpack8(ld8(x + n * ldx + v * 8), f);
#pragma unroll
    for (int c = 0; c < 8; ++c) f[c] *= scale;
    if (res) {
      float rf[8];
      unpack8(ld8(res + r * ldr + v * 8), rf);
#pragma unroll
      for (int c = 0; c < 8; ++c) f[c] += rf[c];
    }
    st8(y + r * ldy + v * 8, pack8(f));
  }
}

static int ew_grid2(int64_t total) {
  int64_t b = (total + 255) / 256;
  int64_t cap = (int64_t)kNumSMs * 16;
  return (int)(b < cap ? (b < 1 ? 1 : b) : cap);
}

}  // namespace stp

using namespace stp;

extern "C" int stp_maxpool_fwd(const stp_tensor* x, int32_t k, int32_t stride, int32_t pad, const stp_tensor* y,
                               uint8_t* argmax, stp_stream stream) {
  if (x && x->dtype == STP_F32) {   // parity mode
    STP_REQUIRE(f32::f32_ok(x) && f32::f32_ok(y) && k >= 1 && k <= 15 && stride >= 1 && y->c == x->c && y->n == x->n &&
                    y->h == (x->h + 2 * pad - k) / stride + 1 && y->w == (x->w + 2 * pad - k) / stride + 1, "maxpool_fwd (fp32): bad args");
    return f32::maxpool_fwd(x, k, stride, pad, y, argmax, (cudaStream_t)stream);
  }
  STP_REQUIRE(vec_ok(x) && vec_ok(y), "maxpool_fwd: bad tensors");
  STP_REQUIRE(k >= 1 && k <= 15 && stride >= 1 && y->c == x->c && y->n == x->n, "maxpool_fwd: bad args");
  STP_REQUIRE(y->h == (x->h + 2 * pad - k) / stride + 1 && y->w == (x->w + 2 * pad - k) / stride + 1,
              "maxpool_fwd: output size mismatch");
  int64_t rows = pixels(y);
  int cv = x->c / 8;
  if (k == 3 && stride == 2 && pad == 1 && rows * cv < 0x7fffffff && pixels(x) < 0x7fffffff && get_option(OPT_HEAD_STRIP) != 1) {
    maxpool_fwd_k3s2_kernel<<<ew_grid2(rows * cv), 256, 0, (cudaStream_t)stream>>>(
        (const __nv_bfloat16*)x->ptr, x->ld, x->h, x->w, (__nv_bfloat16*)y->ptr, y->ld, y->h, y->w, argmax, (int)(rows * cv), x->c, cv);
    return check_launch("maxpool_fwd");
  }
  maxpool_fwd_kernel<<<ew_grid2(rows * cv), 256, 0, (cudaStream_t)stream>>>(
      (const __nv_bfloat16*)x->ptr, x->ld, x->h, x->w, k, stride, pad, (__nv_bfloat16*)y->ptr, y->ld, y->h, y->w,
      argmax, rows, x->c, cv);
  return check_launch("maxpool_fwd");
}

extern "C" int stp_maxpool_bwd(const stp_tensor* dy, const uint8_t* argmax, int32_t k, int32_t stride, int32_t pad,
                               const stp_tensor* residual, const stp_tensor* dx, stp_stream stream) {
  if (dy && dy->dtype == STP_F32) {   // parity mode
    STP_REQUIRE(f32::f32_ok(dy) && f32::f32_ok(dx) && argmax && dy->c == dx->c && dy->n == dx->n, "maxpool_bwd (fp32): bad tensors");
    if (residual) STP_REQUIRE(f32::f32_ok(residual) && residual->c == dx->c && pixels(residual) == pixels(dx), "maxpool_bwd (fp32): bad residual");
    return f32::maxpool_bwd(dy, argmax, k, stride, pad, residual, dx, (cudaStream_t)stream);
  }
  STP_REQUIRE(vec_ok(dy) && vec_ok(dx) && argmax, "maxpool_bwd: bad tensors");
  STP_REQUIRE(dy->c == dx->c && dy->n == dx->n, "maxpool_bwd: shape mismatch");
  if (residual) STP_REQUIRE(vec_ok(residual) && residual->c == dx->c && pixels(residual) == pixels(dx), "maxpool_bwd: bad residual");
  int64_t rows = pixels(dx);
  int cv = dx->c / 8;
  if (k == 3 && stride == 2 && pad == 1 && rows < 0x7fffffff && dx->h % 2 == 0 && dx->w % 2 == 0 && dy->h == dx->h / 2 &&
      dy->w == dx->w / 2 && get_option(OPT_HEAD_STRIP) != 1) {
    const int64_t total = rows / 4 * cv;
    maxpool_bwd_k3s2_block_kernel<<<ew_grid2(total), 256, 0, (cudaStream_t)stream>>>(
        (const __nv_bfloat16*)dy->ptr, dy->ld, dy->h, dy->w, argmax, residual ? (const __nv_bfloat16*)residual->ptr : nullptr,
        residual ? residual->ld : 0, (__nv_bfloat16*)dx->ptr, dx->ld, dx->h, dx->w, dx->c, cv, (int)total);
    return check_launch("maxpool_bwd");
  }
  if (k == 3 && stride == 2 && pad == 1 && rows < 0x7fffffff && cv <= 256 && 256 % cv == 0) {
    const int ppi = 256 / cv;
    int64_t nb = (rows + (int64_t)ppi * 4 - 1) / ((int64_t)ppi * 4);
    if (nb > kNumSMs * 16) nb = kNumSMs * 16;
    int64_t ppb = (rows + nb - 1) / nb;
    ppb = (ppb + ppi - 1) / ppi * ppi;
    nb = (rows + ppb - 1) / ppb;
    maxpool_bwd_k3s2_kernel<<<(int)nb, 256, 0, (cudaStream_t)stream>>>(
        (const __nv_bfloat16*)dy->ptr, dy->ld, dy->h, dy->w, argmax, residual ? (const __nv_bfloat16*)residual->ptr : nullptr,
        residual ? residual->ld : 0, (__nv_bfloat16*)dx->ptr, dx->ld, dx->h, dx->w, (int)rows, dx->c, cv, ppi, (int)ppb);
    return check_launch("maxpool_bwd");
  }
  maxpool_bwd_kernel<<<ew_grid2(rows * cv), 256, 0, (cudaStream_t)stream>>>(
      (const __nv_bfloat16*)dy->ptr, dy->ld, dy->h, dy->w, argmax, k, stride, pad,
      residual ? (const __nv_bfloat16*)residual->ptr : nullptr, residual ? residual->ld : 0, (__nv_bfloat16*)dx->ptr,
      dx->ld, dx->h, dx->w, rows, dx->c, cv);
  return check_launch("maxpool_bwd");
}

extern "C" int stp_avgpool_fwd(const stp_tensor* x, int32_t k, const stp_tensor* y, stp_stream stream) {
  if (x && x->dtype == STP_F32) {   // parity mode
    STP_REQUIRE(f32::f32_ok(x) && f32::f32_ok(y) && k >= 1 && y->c == x->c && y->n == x->n && x->h == y->h * k && x->w == y->w * k,
                "avgpool_fwd (fp32): windows must tile the input exactly");
    return f32::avgpool_fwd(x, k, y, (cudaStream_t)stream);
  }
  STP_REQUIRE(vec_ok(x) && vec_ok(y), "avgpool_fwd: bad tensors");
  STP_REQUIRE(k >= 1 && y->c == x->c && y->n == x->n && x->h == y->h * k && x->w == y->w * k, "avgpool_fwd: windows must tile the input exactly");
  const int64_t rows = pixels(y);
  const int cv = x->c / 8;
  avgpool_fwd_kernel<<<ew_grid2(rows * cv), 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)x->ptr, x->ld, x->h, x->w, k,
                                                                            (__nv_bfloat16*)y->ptr, y->ld, y->h, y->w, rows, cv);
  return check_launch("avgpool_fwd");
}
extern "C" int stp_avgpool_bwd(const stp_tensor* dy, int32_t k, const stp_tensor* residual, const stp_tensor* dx, stp_stream stream) {
  if (dx && dx->dtype == STP_F32) {   // parity mode
    STP_REQUIRE(f32::f32_ok(dy) && f32::f32_ok(dx) && k >= 1 && dy->c == dx->c && dy->n == dx->n && dx->h == dy->h * k && dx->w == dy->w * k &&
                    (!residual || (f32::f32_ok(residual) && residual->c == dx->c && pixels(residual) == pixels(dx))),
                "avgpool_bwd (fp32): shape mismatch");
    return f32::avgpool_bwd(dy, k, residual, dx, (cudaStream_t)stream);
  }
  STP_REQUIRE(vec_ok(dy) && vec_ok(dx), "avgpool_bwd: bad tensors");
  STP_REQUIRE(k >= 1 && dy->c == dx->c && dy->n == dx->n && dx->h == dy->h * k && dx->w == dy->w * k, "avgpool_bwd: shape mismatch");
  if (residual) STP_REQUIRE(vec_ok(residual) && residual->c == dx->c && pixels(residual) == pixels(dx), "avgpool_bwd: bad residual");
  const int64_t rows = pixels(dx);
  const int cv = dx->c / 8;
  avgpool_bwd_kernel<<<ew_grid2(rows * cv), 256, 0, (cudaStream_t)stream>>>(
      (const __nv_bfloat16*)dy->ptr, dy->ld, dy->h, dy->w, k, residual ? (const __nv_bfloat16*)residual->ptr : nullptr,
      residual ? residual->ld : 0, (__nv_bfloat16*)dx->ptr, dx->ld, dx->h, dx->w, rows, cv);
  return check_launch("avgpool_bwd");
}

static bool spatial_f32(const stp_tensor* small, const stp_tensor* big) {
  return small && big && big->dtype == STP_F32 && f32::f32_ok(small) && f32::f32_ok(big) && small->c == big->c && small->n == big->n &&
         small->h == 1 && small->w == 1;
}
static int spatial_check(const stp_tensor* small, const stp_tensor* big, const char* who) {
  STP_REQUIRE(vec_ok(small) && vec_ok(big) && small->c == big->c && small->n == big->n && small->h == 1 && small->w == 1,
              "%s: [n,1,1,c] against [n,h,w,c], bf16, c %% 8 == 0", who);
  return STP_OK;
}
static int spatial_reduce(const stp_tensor* x, float scale, const stp_tensor* y, stp_stream stream, const char* who) {
  const int cv = x->c / 8;
  dim3 grid((cv + 7) / 8, x->n);
  spatial_reduce_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)x->ptr, x->ld, x->h * x->w, cv, scale,
                                                                (__nv_bfloat16*)y->ptr, y->ld);
  return check_launch(who);
}
static int spatial_bcast(const stp_tensor* x, float scale, const stp_tensor* residual, const stp_tensor* y, stp_stream stream, const char* who) {
  if (residual) STP_REQUIRE(vec_ok(residual) && residual->c == y->c && pixels(residual) == pixels(y), "%s: bad residual", who);
  const int cv = y->c / 8;
  const int64_t total = pixels(y) * cv;
  spatial_bcast_kernel<<<ew_grid2(total), 256, 0, (cudaStream_t)stream>>>(
      (const __nv_bfloat16*)x->ptr, x->ld, y->h * y->w, cv, scale, residual ? (const __nv_bfloat16*)residual->ptr : nullptr,
      residual ? residual->ld : 0, (__nv_bfloat16*)y->ptr, y->ld, total);
  return check_launch(who);
}

extern "C" int stp_global_avgpool_fwd(const stp_tensor* x, const stp_tensor* y, stp_stream stream) {
  if (x && x->dtype == STP_F32) {
    STP_REQUIRE(spatial_f32(y, x), "global_avgpool_fwd (fp32): bad tensors");
    return f32::spatial_reduce(x, 1.0 / (double)(x->h * x->w), y, (cudaStream_t)stream);
  }
  int rc = spatial_check(y, x, "global_avgpool_fwd");
  return rc ? rc : spatial_reduce(x, 1.f / (float)(x->h * x->w), y, stream, "global_avgpool_fwd");
}
extern "C" int stp_global_avgpool_bwd(const stp_tensor* dy, const stp_tensor* residual, const stp_tensor* dx, stp_stream stream) {
  if (dx && dx->dtype == STP_F32) {
    STP_REQUIRE(spatial_f32(dy, dx) && (!residual || (f32::f32_ok(residual) && residual->c == dx->c && pixels(residual) == pixels(dx))),
                "global_avgpool_bwd (fp32): bad tensors");
    return f32::spatial_bcast(dy, 1.f / (float)(dx->h * dx->w), residual, dx, (cudaStream_t)stream);
  }
  int rc = spatial_check(dy, dx, "global_avgpool_bwd");
  return rc ? rc : spatial_bcast(dy, 1.f / (float)(dx->h * dx->w), residual, dx, stream, "global_avgpool_bwd");
}
extern "C" int stp_broadcast_fwd(const stp_tensor* x, const stp_tensor* y, stp_stream stream) {
  if (y && y->dtype == STP_F32) {
    STP_REQUIRE(spatial_f32(x, y), "broadcast_fwd (fp32): bad tensors");
    return f32::spatial_bcast(x, 1.f, nullptr, y, (cudaStream_t)stream);
  }
  int rc = spatial_check(x, y, "broadcast_fwd");
  return rc ? rc : spatial_bcast(x, 1.f, nullptr, y, stream, "broadcast_fwd");
}
extern "C" int stp_broadcast_bwd(const stp_tensor* dy, const stp_tensor* dx, stp_stream stream) {
  if (dy && dy->dtype == STP_F32) {
    STP_REQUIRE(spatial_f32(dx, dy), "broadcast_bwd (fp32): bad tensors");
    return f32::spatial_reduce(dy, 1.0, dx, (cudaStream_t)stream);
  }
  int rc = spatial_check(dx, dy, "broadcast_bwd");
  return rc ? rc : spatial_reduce(dy, 1.f, dx, stream, "broadcast_bwd");
}
