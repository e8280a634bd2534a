// Segmentation head: final 3x3 'same' conv with classes <= 4 outputs (+bias).  AI ~ 8 FLOP/B -> HBM bound,
// so it runs on CUDA cores with fp32 weights: one thread per output pixel, 16-byte channel vectors.
// Replaces keras Conv2D(classes,(3,3),padding='same',name='final_conv') built by segmentation_models
// (reference segmentation.py:155).  Backward computes dX, dW and dbias in two passes.
#include "common.cuh"
#include "conv.h"
#include "f32_path.h"

namespace stp {

constexpr int kHeadMaxCin = 64;
constexpr int kHeadMaxCls = 4;

__global__ void __launch_bounds__(256) head_fwd_kernel(const __nv_bfloat16* __restrict__ x, int ldx, int H, int W,
                                                       int Cin, const float* __restrict__ w,
                                                       const float* __restrict__ bias, int classes,
                                                       float* __restrict__ logits, int64_t M) {
  __shared__ float ws[kHeadMaxCls * 9 * kHeadMaxCin];
  for (int i = threadIdx.x; i < classes * 9 * Cin; i += blockDim.x) ws[i] = __bfloat162float(__float2bfloat16(w[i]));  // weights are bf16 in compute
  __syncthreads();
  for (int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; m < M; m += (int64_t)gridDim.x * blockDim.x) {
    int64_t n = m / ((int64_t)H * W);
    int rem = (int)(m - n * (int64_t)H * W);
    int h = rem / W, wq = rem - h * W;
    float acc[kHeadMaxCls];
#pragma unroll
    for (int c = 0; c < kHeadMaxCls; ++c) acc[c] = (bias && c < classes) ? bias[c] : 0.f;
    for (int r = 0; r < 3; ++r) {
      int hi = h + r - 1;
      if (hi < 0 || hi >= H) continue;
      for (int s = 0; s < 3; ++s) {
        int wi = wq + s - 1;
        if (wi < 0 || wi >= W) continue;
        const __nv_bfloat16* px = x + ((n * H + hi) * (int64_t)W + wi) * ldx;
        for (int v = 0; v < Cin / 8; ++v) {
          float f[8];
          unpack8(ld8(px + v * 8), f);
#pragma unroll
          for (int c = 0; c < kHeadMaxCls; ++c) {
            if (c < classes) {
              const float* wp = ws + ((c * 3 + r) * 3 + s) * Cin + v * 8;
#pragma unroll
              for (int k = 0; k < 8; ++k) acc[c] += f[k] * wp[k];
            }
          }
        }
      }
    }
    for (int c = 0; c < classes; ++c) logits[m * classes + c] = acc[c];
  }
}

// dx[m][ci] = sum_{r,s,c} dl[m - off(r,s)][c] * w[c][r][s][ci]   (off = (r-1, s-1))
__global__ void __launch_bounds__(256) head_dgrad_kernel(const float* __restrict__ dl, int H, int W, int Cin,
                                                         const float* __restrict__ w, int classes,
                                                         __nv_bfloat16* __restrict__ dx, int lddx, int64_t M) {
  __shared__ float ws[kHeadMaxCls * 9 * kHeadMaxCin];
  for (int i = threadIdx.x; i < classes * 9 * Cin; i += blockDim.x) ws[i] = __bfloat162float(__float2bfloat16(w[i]));  // weights are bf16 in compute
  __syncthreads();
  const int cv = Cin / 8;
  const int64_t total = M * cv;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t m = i / cv;
    int v = (int)(i - m * cv);
    int64_t n = m / ((int64_t)H * W);
    int rem = (int)(m - n * (int64_t)H * W);
    int h = rem / W, wq = rem - h * W;
    float acc[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] = 0.f;
    for (int r = 0; r < 3; ++r) {
      int ho = h - (r - 1);
      if (ho < 0 || ho >= H) continue;
      for (int s = 0; s < 3; ++s) {
        int wo = wq - (s - 1);
        if (wo < 0 || wo >= W) continue;
        const float* g = dl + ((n * H + ho) * (int64_t)W + wo) * classes;
        for (int c = 0; c < classes; ++c) {
          float gc = g[c];
          const float* wp = ws + ((c * 3 + r) * 3 + s) * Cin + v * 8;
#pragma unroll
          for (int k = 0; k < 8; ++k) acc[k] += gc * wp[k];
        }
      }
    }
    st8(dx + m * lddx + v * 8, pack8(acc));
  }
}

// classes == 1 fast path: a thread owns one 8-channel vector (its 9x8 weights live in registers) and walks pixels, so a
// pixel costs 9 L1-resident dlogit loads + 72 FMAs + one 16-byte store.
__global__ void __launch_bounds__(256) head_dgrad1_kernel(const float* __restrict__ dl, int H, int W, int Cin,
                                                          const float* __restrict__ w, __nv_bfloat16* __restrict__ dx,
                                                          int lddx, int M, int cv, int ppi, int pix_per_blk) {
  const int v = threadIdx.x % cv, pl = threadIdx.x / cv;
  float wr[9][8];
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int k = 0; k < 8; ++k) wr[t][k] = __bfloat162float(__float2bfloat16(w[t * Cin + v * 8 + k]));
  const int m_begin = blockIdx.x * pix_per_blk;
  int m_end = m_begin + pix_per_blk;
  if (m_end > M) m_end = M;
  // U pixels per iteration: their 9 dlogit loads each are independent, so U*9 loads are in flight per thread instead of 9
  constexpr int U = 4;
  for (int m0 = m_begin + pl; m0 < m_end; m0 += U * ppi) {
    float g[U][9];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int m = m0 + u * ppi;
      const unsigned n = (unsigned)m / (unsigned)(H * W);
      const unsigned rem = (unsigned)m - n * (unsigned)(H * W);
      const int h = (int)(rem / (unsigned)W), wq = (int)(rem - (unsigned)h * (unsigned)W);
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        const int ho = h - (r - 1);
#pragma unroll
        for (int s = 0; s < 3; ++s) {
          const int wo = wq - (s - 1);
          g[u][r * 3 + s] = (m < m_end && ho >= 0 && ho < H && wo >= 0 && wo < W) ? __ldg(dl + ((int64_t)n * H + ho) * W + wo) : 0.f;
        }
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int m = m0 + u * ppi;
      if (m >= m_end) break;
      float acc[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[k] = 0.f;
#pragma unroll
      for (int t = 0; t < 9; ++t)
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[k] += g[u][t] * wr[t][k];
      st8(dx + (int64_t)m * lddx + v * 8, pack8(acc));
    }
  }
}

// f32 [classes][9][Cin] master weights -> bf16 [16][9][Cin], rows >= classes zero (operand of the tcgen05 head conv)
__global__ void head_weight_pad_kernel(const float* __restrict__ w, int classes, int k, __nv_bfloat16* __restrict__ out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 16 * k) return;
  out[i] = __float2bfloat16(i < classes * k ? w[i] : 0.f);
}

// Block reduction shared by the head wgrad kernels: xor shuffles over the lanes of a warp that hold the same 8-channel vector
// (threadIdx.x % cv), then across the 8 warps through shared memory; one partial row [9][Cin] (+ the bias sum) per block.
__device__ __forceinline__ void head_wgrad_block_reduce(float (&acc)[9][8], float bsum, int cv, int Cin, float* __restrict__ partial) {
  // lanes l, l^cv, l^2cv ... of a warp hold the same vector
  for (int off = 16; off >= cv; off >>= 1) {
#pragma unroll
    for (int t = 0; t < 9; ++t)
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[t][k] += __shfl_xor_sync(0xffffffffu, acc[t][k], off);
    bsum += __shfl_xor_sync(0xffffffffu, bsum, off);
  }
  __shared__ float sm[8][8 * 72 + 1];  // [warp][vector * 72 + tap * 8 + k], +1 for the bias sum
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane < cv) {
#pragma unroll
    for (int t = 0; t < 9; ++t)
#pragma unroll
      for (int k = 0; k < 8; ++k) sm[warp][lane * 72 + t * 8 + k] = acc[t][k];
    if (lane == 0) sm[warp][8 * 72] = bsum;
  }
  __syncthreads();
  float* out = partial + (int64_t)blockIdx.x * (9 * Cin + 1);
  for (int i = threadIdx.x; i < 9 * Cin + 1; i += blockDim.x) {
    float a = 0.f;
    if (i == 9 * Cin) {
      for (int w = 0; w < 8; ++w) a += sm[w][8 * 72];
    } else {
      const int t = i / Cin, ci = i - t * Cin;
      const int vv = ci >> 3, k = ci & 7;
      for (int w = 0; w < 8; ++w) a += sm[w][vv * 72 + t * 8 + k];
    }
    out[i] = a;
  }
}

// dw[c][r][s][ci] = sum_m dl[m][c] * x[m + off(r,s)][ci]  ==  sum_p x[p][ci] * dl[p - off][c]
// Thread = one 8-channel vector (fixed) walking pixels, 9 taps x 8 accumulators in registers.  Block reduction: xor
// shuffles over the lanes that share the vector, then across the 8 warps through shared memory; one partial row per
// block; a second kernel (one block per output element) sums the block partials in a fixed order.
__global__ void __launch_bounds__(256) head_wgrad_kernel(const __nv_bfloat16* __restrict__ x, int ldx, int H, int W,
                                                         int Cin, const float* __restrict__ dl, int classes, int cls,
                                                         float* __restrict__ partial, int M, int pix_per_blk) {
  // partial[blk][9][Cin] for class `cls`, and slot 9*Cin for dbias
  const int cv = Cin / 8;                 // power of two <= 8 (checked by the launcher)
  const int v = threadIdx.x % cv;
  const int lanes = blockDim.x / cv;
  const int pl = threadIdx.x / cv;
  float acc[9][8];
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[t][k] = 0.f;
  float bsum = 0.f;
  const int m_begin = blockIdx.x * pix_per_blk;
  int m_end = m_begin + pix_per_blk;
  if (m_end > M) m_end = M;
  // U pixels per iteration: U independent 16-byte activation loads (+ 9 L1-resident dlogit loads each) in flight per thread;
  // the accumulation order over a thread's pixels is unchanged (u = 0..U-1 in sequence) -> same partial sums as before
  constexpr int U = 4;
  for (int p0 = m_begin + pl; p0 < m_end; p0 += U * lanes) {
    bf16x8 q[U];
    float g[U][9];
    float gb[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int p = p0 + u * lanes;
      const bool ok = p < m_end;
      if (ok) q[u] = ld8(x + (int64_t)p * ldx + v * 8);
      const unsigned n = (unsigned)p / (unsigned)(H * W);
      const unsigned rem = (unsigned)p - n * (unsigned)(H * W);
      const int h = (int)(rem / (unsigned)W), wq = (int)(rem - (unsigned)h * (unsigned)W);
      gb[u] = (ok && v == 0) ? dl[(int64_t)p * classes + cls] : 0.f;
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        const int ho = h - (r - 1);
#pragma unroll
        for (int s = 0; s < 3; ++s) {
          const int wo = wq - (s - 1);
          g[u][r * 3 + s] = (ok && ho >= 0 && ho < H && wo >= 0 && wo < W)
                                ? __ldg(dl + (((int64_t)n * H + ho) * W + wo) * classes + cls) : 0.f;
        }
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (p0 + u * lanes >= m_end) break;
      float f[8];
      unpack8(q[u], f);
      bsum += gb[u];
#pragma unroll
      for (int t = 0; t < 9; ++t)
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[t][k] += f[k] * g[u][t];
    }
  }
  head_wgrad_block_reduce(acc, bsum, cv, Cin, partial);
}

// Column-strip variants (H % strip == 0): a thread owns one 8-channel vector of ONE image column and walks `strip` rows down,
// with a 3x3 sliding window over the dlogits -- 3 new dlogit loads per pixel instead of 9, no per-pixel index divisions, and
// the lanes of a warp cover consecutive (pixel, vector) pairs of a row, so every load / store instruction is one contiguous
// 512-byte segment.  History (profiles/r2_ncu_full_tail_head*.metrics.txt): the per-pixel kernels ran 365 instructions per
// (pixel, vector) at 12 % occupancy (167 / 107 us for 134 MB of activations); ROW strips cut the instructions 2.2x but every
// warp load then touched 16 different 128-byte lines and the L1 pipe saturated (82 %, 143 / 114 us).
__global__ void __launch_bounds__(256, 2) head_wgrad_strip_kernel(const __nv_bfloat16* __restrict__ x, int ldx, int H, int W, int Cin,
                                                                  const float* __restrict__ dl, int classes, int cls,
                                                                  float* __restrict__ partial, int items, int strip) {
  const int cv = Cin / 8;
  const int item = blockIdx.x * blockDim.x + threadIdx.x;
  const int v = threadIdx.x % cv;           // == item % cv (256 % cv == 0)
  float acc[9][8];
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[t][k] = 0.f;
  float bsum = 0.f;
  if (item < items) {
    int q = item / cv;
    const int w = q % W;
    q /= W;                                  // n * (H / strip) + hb
    const int hbs = H / strip;
    const int n = q / hbs;
    const int h0 = (q - n * hbs) * strip;
    // d[r][s] = dl[h - (r-1)][w - (s-1)]  (tap (r, s) of the forward conv pairs x[p] with dl at p - off(r, s))
    const float* col[3];
    bool okc[3];
#pragma unroll
    for (int sx = 0; sx < 3; ++sx) {
      const int wo = w - (sx - 1);
      okc[sx] = wo >= 0 && wo < W;
      col[sx] = dl + ((int64_t)n * H * W + (okc[sx] ? wo : 0)) * classes + cls;
    }
    float d[3][3];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const int ho = h0 - (r - 1);
#pragma unroll
      for (int sx = 0; sx < 3; ++sx)
        d[r][sx] = (okc[sx] && ho >= 0 && ho < H) ? __ldg(col[sx] + (int64_t)ho * W * classes) : 0.f;
    }
    const __nv_bfloat16* xp = x + (((int64_t)n * H + h0) * W + w) * ldx + v * 8;
    const int64_t xstep = (int64_t)W * ldx;
#pragma unroll 2
    for (int i = 0; i < strip; ++i) {
      float f[8];
      unpack8(ld8(xp), f);
      xp += xstep;
      if (v == 0) bsum += d[1][1];
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int sx = 0; sx < 3; ++sx)
#pragma unroll
          for (int k = 0; k < 8; ++k) acc[r * 3 + sx][k] += f[k] * d[r][sx];
      const int hn = h0 + i + 2;             // the row that enters the window as r = 0 of the next pixel
#pragma unroll
      for (int sx = 0; sx < 3; ++sx) {
        d[2][sx] = d[1][sx];
        d[1][sx] = d[0][sx];
        d[0][sx] = (okc[sx] && hn < H) ? __ldg(col[sx] + (int64_t)hn * W * classes) : 0.f;
      }
    }
  }
  head_wgrad_block_reduce(acc, bsum, cv, Cin, partial);
}

// classes == 1: dx[p][ci] = sum_{r,s} dl[h-(r-1)][w-(s-1)] * w[r][s][ci]
__global__ void __launch_bounds__(256, 2) head_dgrad1_strip_kernel(const float* __restrict__ dl, int H, int W, int Cin,
                                                                   const float* __restrict__ w, __nv_bfloat16* __restrict__ dx, int lddx,
                                                                   int items, int strip) {
  const int cv = Cin / 8;
  const int item = blockIdx.x * blockDim.x + threadIdx.x;
  if (item >= items) return;
  const int v = item % cv;
  int q = item / cv;
  const int wq = q % W;
  q /= W;
  const int hbs = H / strip;
  const int n = q / hbs;
  const int h0 = (q - n * hbs) * strip;
  float wr[9][8];
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int k = 0; k < 8; ++k) wr[t][k] = __bfloat162float(__float2bfloat16(w[t * Cin + v * 8 + k]));
  const float* col[3];
  bool okc[3];
#pragma unroll
  for (int sx = 0; sx < 3; ++sx) {
    const int wo = wq - (sx - 1);
    okc[sx] = wo >= 0 && wo < W;
    col[sx] = dl + (int64_t)n * H * W + (okc[sx] ? wo : 0);
  }
  float d[3][3];
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    const int ho = h0 - (r - 1);
#pragma unroll
    for (int sx = 0; sx < 3; ++sx) d[r][sx] = (okc[sx] && ho >= 0 && ho < H) ? __ldg(col[sx] + (int64_t)ho * W) : 0.f;
  }
  __nv_bfloat16* op = dx + (((int64_t)n * H + h0) * W + wq) * lddx + v * 8;
  const int64_t ostep = (int64_t)W * lddx;
#pragma unroll 2
  for (int i = 0; i < strip; ++i) {
    float acc[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] = 0.f;
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int sx = 0; sx < 3; ++sx)
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[k] += d[r][sx] * wr[r * 3 + sx][k];
    st8(op, pack8(acc));
    op += ostep;
    const int hn = h0 + i + 2;
#pragma unroll
    for (int sx = 0; sx < 3; ++sx) {
      d[2][sx] = d[1][sx];
      d[1][sx] = d[0][sx];
      d[0][sx] = (okc[sx] && hn < H) ? __ldg(col[sx] + (int64_t)hn * W) : 0.f;
    }
  }
}

// one block per output element: strided partial sums in double, fixed tree -> deterministic
__global__ void __launch_bounds__(128) head_wgrad_final_kernel(const float* __restrict__ partial, int nblk, int Cin, int cls,
                                                               float* __restrict__ dw, float* __restrict__ dbias) {
  const int i = blockIdx.x;
  const int n = 9 * Cin + 1;
  double a = 0.0;
  for (int b = threadIdx.x; b < nblk; b += 128) a += (double)partial[(int64_t)b * n + i];
  __shared__ double red[128];
  red[threadIdx.x] = a;
  __syncthreads();
  for (int half = 64; half >= 1; half >>= 1) {
    if (threadIdx.x < half) red[threadIdx.x] += red[threadIdx.x + half];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    if (i == 9 * Cin) {
      if (dbias) dbias[cls] = (float)red[0];
    } else {
      dw[(int64_t)cls * 9 * Cin + i] = (float)red[0];
    }
  }
}

constexpr int kHeadWgradBlocks = kNumSMs * 8;

}  // namespace stp

using namespace stp;

extern "C" size_t stp_head_fwd_workspace(const stp_tensor* x, int32_t classes) {
  (void)classes;
  return x ? (size_t)16 * 9 * (size_t)x->c * sizeof(__nv_bfloat16) : 0;
}

extern "C" int stp_head_fwd(const stp_tensor* x, const float* w_krsc_f32, const float* bias, int32_t classes,
                            float* logits, void* workspace, size_t workspace_bytes, stp_stream stream) {
  if (x && x->dtype == STP_F32) {   // parity mode: the same 3x3 'same' convolution + bias on the CUDA-core fp32 kernel
    STP_REQUIRE(f32::f32_ok(x) && w_krsc_f32 && logits && classes >= 1, "head_fwd (fp32): bad args");
    f32::ConvF p;
    p.x = (const float*)x->ptr; p.ldx = x->ld; p.N = x->n; p.H = x->h; p.W = x->w; p.Cin = x->c;
    p.w = w_krsc_f32; p.y = logits; p.ldy = classes; p.Ho = x->h; p.Wo = x->w; p.Cout = classes;
    p.res = nullptr; p.ldr = 0; p.bias = bias;
    p.R = 3; p.S = 3; p.stride = 1; p.pad_h = 1; p.pad_w = 1; p.up = 1; p.relu = 0; p.dgrad = 0;
    p.M = pixels(x); p.K = 9 * x->c;
    return f32::launch_conv(p, (cudaStream_t)stream);
  }
  STP_REQUIRE(vec_ok(x) && w_krsc_f32 && logits, "head_fwd: bad args");
  STP_REQUIRE(classes >= 1 && classes <= kHeadMaxCls && x->c <= kHeadMaxCin, "head_fwd: classes<=4, Cin<=64");
  int64_t M = pixels(x);
  cudaStream_t st = (cudaStream_t)stream;
  // tcgen05 path: implicit GEMM with Cout padded to 16 (N = 16 MMA), only `classes` columns stored
  ConvP p;
  p.x = (const __nv_bfloat16*)x->ptr; p.ldx = x->ld; p.N = x->n; p.H = x->h; p.W = x->w; p.Cin = x->c;
  p.w = (const __nv_bfloat16*)workspace;
  p.y = logits; p.ldy = classes; p.Ho = x->h; p.Wo = x->w; p.Cout = 16; p.y_f32 = 1;
  p.res = nullptr; p.ldr = 0; p.bias = bias;
  p.R = 3; p.S = 3; p.stride = 1; p.pad_h = 1; p.pad_w = 1; p.up = 1; p.relu = 0;
  p.M = M; p.K = 9 * x->c; p.ncls = classes;
  if (stp_tc_enabled() && workspace && workspace_bytes >= stp_head_fwd_workspace(x, classes) && aligned16(workspace) &&
      tc2_conv_supported(p)) {
    const int k = 9 * x->c;
    head_weight_pad_kernel<<<(16 * k + 255) / 256, 256, 0, st>>>(w_krsc_f32, classes, k, (__nv_bfloat16*)workspace);
    int rc = check_launch("head_weight_pad");
    if (rc) return rc;
    return launch_tc2_conv(p, st);
  }
  int64_t nb = (M + 255) / 256;
  head_fwd_kernel<<<(int)(nb < kNumSMs * 16 ? nb : kNumSMs * 16), 256, 0, st>>>(
      (const __nv_bfloat16*)x->ptr, x->ld, x->h, x->w, x->c, w_krsc_f32, bias, classes, logits, M);
  return check_launch("head_fwd");
}

// strip height of the column-strip backward kernels: the smallest of 16, 32, 64, ... that divides H and keeps the block count
// (one partial row per block) within `max_blocks`; 0 = not applicable (the per-pixel kernels run)
static int head_strip(const stp_tensor* x, int64_t max_blocks) {
  const int cv = x->c / 8;
  for (int L = 16; L <= x->h; L *= 2) {
    if (x->h % L != 0) return 0;
    const int64_t items = pixels(x) / L * cv;
    if ((items + 255) / 256 <= max_blocks) return L;
  }
  return 0;
}

extern "C" size_t stp_head_bwd_workspace(const stp_tensor* x, int32_t classes) {
  (void)classes;
  return (size_t)kHeadWgradBlocks * (9 * (size_t)x->c + 1) * sizeof(float);
}

extern "C" int stp_head_bwd(const stp_tensor* x, const float* w_krsc_f32, const float* dlogits, int32_t classes,
                            const stp_tensor* dx, float* dw, float* dbias, void* workspace, size_t workspace_bytes,
                            stp_stream stream) {
  if (x && x->dtype == STP_F32) {   // parity mode: dgrad + wgrad of the same convolution; db = column sums of dlogits
    STP_REQUIRE(f32::f32_ok(x) && w_krsc_f32 && dlogits && dw && classes >= 1 && workspace && workspace_bytes >= 2 * sizeof(double) * classes,
                "head_bwd (fp32): bad args / workspace");
    cudaStream_t st = (cudaStream_t)stream;
    f32::ConvF p;
    p.R = 3; p.S = 3; p.bias = nullptr; p.res = nullptr; p.ldr = 0; p.relu = 0;
    if (dx) {
      STP_REQUIRE(f32::f32_ok(dx) && dx->c == x->c && pixels(dx) == pixels(x), "head_bwd (fp32): bad dx");
      p.x = dlogits; p.ldx = classes; p.N = x->n; p.H = x->h; p.W = x->w; p.Cin = classes;
      p.w = w_krsc_f32; p.y = (float*)dx->ptr; p.ldy = dx->ld; p.Ho = x->h; p.Wo = x->w; p.Cout = x->c;
      p.stride = 1; p.up = 1; p.pad_h = 1; p.pad_w = 1; p.dgrad = 1; p.M = pixels(x); p.K = 9 * classes;
      int rc = f32::launch_conv(p, st);
      if (rc) return rc;
    }
    p.x = (const float*)x->ptr; p.ldx = x->ld; p.N = x->n; p.H = x->h; p.W = x->w; p.Cin = x->c;
    p.w = nullptr; p.y = nullptr; p.ldy = classes; p.Ho = x->h; p.Wo = x->w; p.Cout = classes;
    p.stride = 1; p.up = 1; p.pad_h = 1; p.pad_w = 1; p.dgrad = 0; p.M = pixels(x); p.K = 9 * x->c;
    int rc = f32::launch_wgrad(p, dlogits, classes, classes, dw, st);
    if (rc || !dbias) return rc;
    stp_tensor dl = {const_cast<float*>(dlogits), x->n, x->h, x->w, classes, classes, STP_F32};
    cudaMemsetAsync(workspace, 0, 2 * sizeof(double) * classes, st);
    FinArgs fin = {};
    fin.mode = 3; fin.acc = (double*)workspace; fin.dbeta = dbias;
    return f32::launch_reduce(0, &dl, nullptr, nullptr, 0, 1, fin, st);
  }
  STP_REQUIRE(vec_ok(x) && w_krsc_f32 && dlogits && dw, "head_bwd: bad args");
  STP_REQUIRE(classes >= 1 && classes <= kHeadMaxCls && x->c <= kHeadMaxCin, "head_bwd: classes<=4, Cin<=64");
  STP_REQUIRE(x->c == 8 || x->c == 16 || x->c == 32 || x->c == 64, "head_bwd: Cin must be 8, 16, 32 or 64");
  if (workspace_bytes < stp_head_bwd_workspace(x, classes) || !workspace) {
    set_error("head_bwd: workspace too small");
    return STP_E_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  int64_t M = pixels(x);
  if (dx) {
    STP_REQUIRE(vec_ok(dx) && dx->c == x->c && pixels(dx) == M, "head_bwd: bad dx");
    const int cv = x->c / 8;
    const int dstrip = head_strip(x, 0x7fffffff);
    if (classes == 1 && M < 0x7fffffff && dstrip && get_option(OPT_HEAD_STRIP) != 1) {
      const int items = (int)(M / dstrip) * cv;
      head_dgrad1_strip_kernel<<<(items + 255) / 256, 256, 0, st>>>(dlogits, x->h, x->w, x->c, w_krsc_f32, (__nv_bfloat16*)dx->ptr,
                                                                    dx->ld, items, dstrip);
    } else if (classes == 1 && M < 0x7fffffff) {
      const int ppi = 256 / cv;
      int64_t nb = (M + (int64_t)ppi * 8 - 1) / ((int64_t)ppi * 8);
      if (nb > kNumSMs * 8) nb = kNumSMs * 8;
      int64_t ppb = (M + nb - 1) / nb;
      ppb = (ppb + ppi - 1) / ppi * ppi;
      nb = (M + ppb - 1) / ppb;
      head_dgrad1_kernel<<<(int)nb, ppi * cv, 0, st>>>(dlogits, x->h, x->w, x->c, w_krsc_f32,
                                                       (__nv_bfloat16*)dx->ptr, dx->ld, (int)M, cv, ppi, (int)ppb);
    } else {
      int64_t nb = (M * cv + 255) / 256;
      head_dgrad_kernel<<<(int)(nb < kNumSMs * 16 ? nb : kNumSMs * 16), 256, 0, st>>>(
          dlogits, x->h, x->w, x->c, w_krsc_f32, classes, (__nv_bfloat16*)dx->ptr, dx->ld, M);
    }
    int rc = check_launch("head_dgrad");
    if (rc) return rc;
  }
  STP_REQUIRE(M < 0x7fffffff, "head_bwd: too many pixels");
  const int wstrip = head_strip(x, kHeadWgradBlocks);
  if (wstrip && get_option(OPT_HEAD_STRIP) != 1) {
    const int items = (int)(M / wstrip) * (x->c / 8);
    const int nbs = (items + 255) / 256;
    for (int cls = 0; cls < classes; ++cls) {
      head_wgrad_strip_kernel<<<nbs, 256, 0, st>>>((const __nv_bfloat16*)x->ptr, x->ld, x->h, x->w, x->c, dlogits, classes, cls,
                                                   (float*)workspace, items, wstrip);
      int rc = check_launch("head_wgrad_strip");
      if (rc) return rc;
      head_wgrad_final_kernel<<<9 * x->c + 1, 128, 0, st>>>((const float*)workspace, nbs, x->c, cls, dw, dbias);
      rc = check_launch("head_wgrad_final");
      if (rc) return rc;
    }
    return STP_OK;
  }
  const int lanes = 256 / (x->c / 8);
  int64_t nb = (M + (int64_t)lanes * 8 - 1) / ((int64_t)lanes * 8);
  if (nb > kHeadWgradBlocks) nb = kHeadWgradBlocks;
  int64_t ppb = (M + nb - 1) / nb;
  ppb = (ppb + lanes - 1) / lanes * lanes;
  nb = (M + ppb - 1) / ppb;
  for (int cls = 0; cls < classes; ++cls) {
    head_wgrad_kernel<<<(int)nb, 256, 0, st>>>((const __nv_bfloat16*)x->ptr, x->ld, x->h, x->w, x->c, dlogits, classes, cls,
                                               (float*)workspace, (int)M, (int)ppb);
    int rc = check_launch("head_wgrad");
    if (rc) return rc;
    head_wgrad_final_kernel<<<9 * x->c + 1, 128, 0, st>>>((const float*)workspace, (int)nb, x->c, cls, dw, dbias);
    rc = check_launch("head_wgrad_final");
    if (rc) return rc;
  }
  return STP_OK;
}
