// PARITY MODE: the whole step with fp32 activations, fp32 weights and fp32 FFMA accumulation on the CUDA cores
// (SURVEY.md section 7 hard-part 7).  Not the fast path -- the product path is bf16 storage on the tcgen05
// kernels -- but the same graph / the same C ABI entry points (they dispatch here when their tensors are STP_F32), so that
// the north_star criterion "loss curve within 1e-3 of the reference over 100 steps" can be checked without the bf16 rounding
// noise that training dynamics amplify to ~1e-2 (tests/test_gpu_model.py::test_loss_curve_100_steps_fp32_parity_mode).
//
// Semantics are those of the bf16 kernels they shadow (same formulas, keras.layers semantics per SURVEY.md Appendix B);
// reductions and convolution accumulators run in double; each conv output is summed by ONE thread in a fixed k order
// (deterministic).
#include "bn_fin.cuh"
#include "common.cuh"
#include "f32_path.h"

namespace stp {
namespace f32 {

static int grid_for(int64_t total, int threads = 256) {
  int64_t b = (total + threads - 1) / threads;
  const int64_t cap = (int64_t)kNumSMs * 32;
  return (int)(b < 1 ? 1 : (b < cap ? b : cap));
}

// ------------------------------------------------------------------------------------------------------------------
// convolution: implicit GEMM, 64 (pixels) x 64 (channels) tile per block, BK = 16, thread = 4 x 4 outputs
// ------------------------------------------------------------------------------------------------------------------
constexpr int TM = 64, TN = 64, TK = 16;

__device__ __forceinline__ float gather_x(const ConvF& p, bool ok, int64_t img_base, int hi0, int wi0, int k) {
  if (!ok || k >= p.K) return 0.f;
  const int tap = k / p.Cin, ci = k - tap * p.Cin;
  const int r = tap / p.S, s = tap - r * p.S;
  int hu = hi0 + r, wu = wi0 + s;
  if (p.up > 1) {
    if (hu < 0 || wu < 0 || (hu % p.up) != 0 || (wu % p.up) != 0) return 0.f;
    hu /= p.up;
    wu /= p.up;
  }
  if (hu < 0 || hu >= p.H || wu < 0 || wu >= p.W) return 0.f;
  return p.x[(img_base + (int64_t)hu * p.W + wu) * p.ldx + ci];
}
// weight element B[n][k]: forward = KRSC [Cout][K]; dgrad = the FORWARD weights [Cf_out = p.Cin][R][S][Cf_in = p.Cout] read
// tap-flipped and transposed (no separate dgrad copy in parity mode)
__device__ __forceinline__ float weight_at(const ConvF& p, int n, int k) {
  if (n >= p.Cout || k >= p.K) return 0.f;
  if (!p.dgrad) return p.w[(int64_t)n * p.K + k];
  const int tap = k / p.Cin, co = k - tap * p.Cin;
  const int r = tap / p.S, s = tap - r * p.S;
  return p.w[(((int64_t)co * p.R + (p.R - 1 - r)) * p.S + (p.S - 1 - s)) * p.Cout + n];
}

__global__ void __launch_bounds__(256) conv_kernel(const ConvF p) {
  __shared__ float As[TK][TM + 4];
  __shared__ float Bs[TK][TN + 4];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;  // tx: channel group, ty: pixel group
  const int64_t m0 = (int64_t)blockIdx.x * TM;
  const int n0 = blockIdx.y * TN;
  // loader mapping: element e = tid + j*256 of the 64 x 16 tile -> row e / 16, k e % 16
  bool ok[4];
  int64_t img_base[4];
  int hi0[4], wi0[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int row = (tid + j * 256) >> 4;
    const int64_t m = m0 + row;
    ok[j] = m < p.M;
    const int64_t mm = ok[j] ? m : 0;
    const int64_t n = mm / ((int64_t)p.Ho * p.Wo);
    const int rem = (int)(mm - n * (int64_t)p.Ho * p.Wo);
    const int ho = rem / p.Wo, wo = rem - ho * p.Wo;
    img_base[j] = n * (int64_t)p.H * p.W;
    hi0[j] = ho * p.stride - p.pad_h;
    wi0[j] = wo * p.stride - p.pad_w;
  }
  // fp32 operands, products and sums in DOUBLE (DFMA): a K = 4608 dot product then carries one final rounding instead of
  // ~sqrt(K) fp32 ones -- the parity mode's job is to sit as close to exact arithmetic as the fp32 storage allows
  double acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;
  for (int k0 = 0; k0 < p.K; k0 += TK) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int e = tid + j * 256, row = e >> 4, kk = e & 15;
      As[kk][row] = gather_x(p, ok[j], img_base[j], hi0[j], wi0[j], k0 + kk);
      Bs[kk][row] = weight_at(p, n0 + row, k0 + kk);
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < TK; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fma((double)a[i], (double)b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t m = m0 + ty * 4 + i;
    if (m >= p.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= p.Cout) continue;
      float v = (float)acc[i][j];
      if (p.bias) v += p.bias[n];
      if (p.res) v += p.res[m * p.ldr + n];
      if (p.relu) v = fmaxf(v, 0.f);
      p.y[m * p.ldy + n] = v;
    }
  }
}

// wgrad: dW[co][k] = sum_m dY[m][co] * A[m][k]; one block = 64 (co) x 64 (k) outputs, loops over ALL pixels (no split:
// one thread sums each output in pixel order -> deterministic)
__global__ void __launch_bounds__(256) wgrad_kernel(const ConvF p, const float* __restrict__ dy, int lddy, int Cdy,
                                                    float* __restrict__ dw) {
  __shared__ float Gs[TK][TM + 4];   // dY[m][co]
  __shared__ float Xs[TK][TN + 4];   // A[m][k]
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int co0 = blockIdx.x * TM, k0 = blockIdx.y * TN;
  double acc[4][4];   // sums over up to millions of pixels: double accumulation (see conv_kernel)
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;
  for (int64_t mb = 0; mb < p.M; mb += TK) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int e = tid + j * 256, mm = e >> 6, c = e & 63;   // 16 pixels x 64 columns
      const int64_t m = mb + mm;
      const bool okm = m < p.M;
      Gs[mm][c] = (okm && co0 + c < Cdy) ? dy[m * lddy + co0 + c] : 0.f;
      float xv = 0.f;
      if (okm && k0 + c < p.K) {
        const int64_t n = m / ((int64_t)p.Ho * p.Wo);
        const int rem = (int)(m - n * (int64_t)p.Ho * p.Wo);
        const int ho = rem / p.Wo, wo = rem - ho * p.Wo;
        xv = gather_x(p, true, n * (int64_t)p.H * p.W, ho * p.stride - p.pad_h, wo * p.stride - p.pad_w, k0 + c);
      }
      Xs[mm][c] = xv;
    }
    __syncthreads();
#pragma unroll
    for (int mm = 0; mm < TK; ++mm) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = Gs[mm][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Xs[mm][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fma((double)a[i], (double)b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int co = co0 + ty * 4 + i;
    if (co >= Cdy) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k = k0 + tx * 4 + j;
      if (k < p.K) dw[(int64_t)co * p.K + k] = (float)acc[i][j];
    }
  }
}

int launch_conv(const ConvF& p, cudaStream_t st) {
  dim3 grid((unsigned)((p.M + TM - 1) / TM), (unsigned)((p.Cout + TN - 1) / TN));
  conv_kernel<<<grid, 256, 0, st>>>(p);
  return check_launch("f32 conv");
}
int launch_wgrad(const ConvF& p, const float* dy, int lddy, int cdy, float* dw, cudaStream_t st) {
  dim3 grid((unsigned)((cdy + TM - 1) / TM), (unsigned)((p.K + TN - 1) / TN));
  wgrad_kernel<<<grid, 256, 0, st>>>(p, dy, lddy, cdy, dw);
  return check_launch("f32 wgrad");
}

// ------------------------------------------------------------------------------------------------------------------
// per-channel row reductions -> double atomics into acc[2][C] -> finalize kernel (FinArgs modes of bn_fin.cuh)
// MODE 0: (sum x, sum x^2).  MODE 1: BatchNorm backward (sum g, sum g*xhat), g = dy [2x2-pooled] masked by the ReLU.
// ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float pooled(const float* __restrict__ dy, int lddy, int pool, int H, int W, int64_t r, int c) {
  if (pool == 1) return dy[r * lddy + c];
  const int64_t n = r / ((int64_t)H * W);
  const int rem = (int)(r - n * (int64_t)H * W);
  const int h = rem / W, w = rem - h * W;
  const float* b = dy + ((n * 2 * H + 2 * h) * (int64_t)(2 * W) + 2 * w) * lddy + c;
  return ((b[0] + b[lddy]) + b[(int64_t)2 * W * lddy]) + b[(int64_t)2 * W * lddy + lddy];
}

template <int MODE>
__global__ void __launch_bounds__(256) reduce_kernel(const float* __restrict__ x, int ldx, const float* __restrict__ dy, int lddy,
                                                     const float* __restrict__ coef, int relu, int pool, int H, int W,
                                                     int64_t rows, int C, double* __restrict__ acc) {
  // thread = (channel c = threadIdx.x % CC + chunk, row lane); blockDim = 256
  const int CC = C < 256 ? C : 256;
  const int lanes = 256 / CC;
  const int cl = threadIdx.x % CC, rl = threadIdx.x / CC;
  __shared__ double sm[2][256];
  for (int c0 = 0; c0 < C; c0 += CC) {
    const int c = c0 + cl;
    double s0 = 0.0, s1 = 0.0;
    if (c < C && rl < lanes) {
      float mean = 0.f, invstd = 0.f, scale = 0.f, shift = 0.f;
      if (MODE == 1) {
        mean = coef[c]; invstd = coef[C + c]; scale = coef[2 * C + c]; shift = coef[3 * C + c];
      }
      for (int64_t r = (int64_t)blockIdx.x * lanes + rl; r < rows; r += (int64_t)gridDim.x * lanes) {
        const float xv = x[r * ldx + c];
        if (MODE == 0) {
          s0 += (double)xv;
          s1 += (double)xv * (double)xv;
        } else {
          float g = pooled(dy, lddy, pool, H, W, r, c);
          if (relu && !relu_pass(xv * scale + shift, relu)) g = 0.f;
          s0 += (double)g;
          s1 += (double)g * (double)((xv - mean) * invstd);
        }
      }
    }
    sm[0][threadIdx.x] = s0;
    sm[1][threadIdx.x] = s1;
    __syncthreads();
    if (rl == 0 && c < C) {
      for (int j = 1; j < lanes; ++j) {
        s0 += sm[0][j * CC + cl];
        s1 += sm[1][j * CC + cl];
      }
      atomicAdd(acc + c, s0);
      atomicAdd(acc + C + c, s1);
    }
    __syncthreads();
  }
}
__global__ void finalize_kernel(const FinArgs f, int C) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const double s = f.acc[c], ss = f.acc[C + c];
  if (f.mode == 1) fin_forward(f, C, c, s, ss); else if (f.mode == 2) fin_backward(f, C, c, s, ss); else f.dbeta[c] = (float)s;
  f.acc[c] = 0.0;
  f.acc[C + c] = 0.0;
}

int launch_reduce(int mode, const stp_tensor* x, const stp_tensor* dy, const float* coef, int relu, int pool, const FinArgs& fin,
                  cudaStream_t st) {
  STP_REQUIRE(fin.acc, "f32 reduce: the double accumulator (acc) is required in parity mode");
  const int64_t rows = pixels(x);
  const int C = x->c;
  const int CC = C < 256 ? C : 256;
  const int lanes = 256 / CC;
  int64_t nb = (rows + lanes * 8 - 1) / (lanes * 8);
  if (nb > kNumSMs * 4) nb = kNumSMs * 4;
  if (nb < 1) nb = 1;
  if (mode == 0)
    reduce_kernel<0><<<(int)nb, 256, 0, st>>>((const float*)x->ptr, x->ld, nullptr, 0, nullptr, 0, 1, x->h, x->w, rows, C, fin.acc);
  else
    reduce_kernel<1><<<(int)nb, 256, 0, st>>>((const float*)x->ptr, x->ld, (const float*)dy->ptr, dy->ld, coef, relu, pool, x->h,
                                              x->w, rows, C, fin.acc);
  int rc = check_launch("f32 reduce");
  if (rc) return rc;
  finalize_kernel<<<(C + 127) / 128, 128, 0, st>>>(fin, C);
  return check_launch("f32 finalize");
}

// ------------------------------------------------------------------------------------------------------------------
// elementwise
// ------------------------------------------------------------------------------------------------------------------
__global__ void bn_apply_kernel(const float* __restrict__ x, int ldx, const float* __restrict__ coef, int relu, int up,
                                float* __restrict__ y, int ldy, int H, int W, int64_t rows, int C) {
  const int64_t total = rows * C;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / C;
    const int c = (int)(i - r * C);
    float t = x[r * ldx + c] * coef[2 * C + c] + coef[3 * C + c];
    t = relu_act(t, relu);
    if (up == 1) {
      y[r * ldy + c] = t;
    } else {
      const int64_t n = r / ((int64_t)H * W);
      const int rem = (int)(r - n * (int64_t)H * W);
      const int h = rem / W, w = rem - h * W;
      float* b = y + ((n * 2 * H + 2 * h) * (int64_t)(2 * W) + 2 * w) * ldy + c;
      b[0] = t; b[ldy] = t; b[(int64_t)2 * W * ldy] = t; b[(int64_t)2 * W * ldy + ldy] = t;
    }
  }
}
// MODE 0: dx = a*g + b*x + cc (+res), g = dy masked by relu(bn(x)).  MODE 1: ReLU backward, mask from x > 0.
template <int MODE>
__global__ void bwd_apply_kernel(const float* __restrict__ dy, int lddy, const float* __restrict__ x, int ldx,
                                 const float* __restrict__ coef, const float* __restrict__ bcoef, int relu, int pool,
                                 const float* __restrict__ res, int ldr, float* __restrict__ dx, int lddx, int H, int W,
                                 int64_t rows, int C) {
  const int64_t total = rows * C;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / C;
    const int c = (int)(i - r * C);
    const float xv = x[r * ldx + c];
    float g = pooled(dy, lddy, pool, H, W, r, c);
    float o;
    if (MODE == 0) {
      if (relu && !relu_pass(xv * coef[2 * C + c] + coef[3 * C + c], relu)) g = 0.f;
      o = bcoef[c] * g + bcoef[C + c] * xv + bcoef[2 * C + c];
    } else {
      o = xv > 0.f ? g : 0.f;
    }
    if (res) o += res[r * ldr + c];
    dx[r * lddx + c] = o;
  }
}
__global__ void stem_prep_kernel(const uint8_t* __restrict__ img, int64_t rows, int cimg, const float* __restrict__ coef,
                                 float* __restrict__ y, int ldy, int C) {
  const int64_t total = rows * C;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / C;
    const int c = (int)(i - r * C);
    float t = 0.f;
    if (c < cimg) t = (float)img[r * cimg + c] * coef[2 * cimg + c] + coef[3 * cimg + c];
    else if (c == cimg) t = 1.f;
    y[r * ldy + c] = t;
  }
}
__global__ void copy_up_kernel(const float* __restrict__ x, int ldx, int up, float* __restrict__ y, int ldy, int H, int W,
                               int64_t rows, int C) {
  const int64_t total = rows * C;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / C;
    const int c = (int)(i - r * C);
    const float v = x[r * ldx + c];
    const int64_t n = r / ((int64_t)H * W);
    const int rem = (int)(r - n * (int64_t)H * W);
    const int h = rem / W, w = rem - h * W;
    for (int a = 0; a < up; ++a)
      for (int b = 0; b < up; ++b) y[((n * up * H + up * h + a) * (int64_t)(up * W) + up * w + b) * ldy + c] = v;
  }
}
__global__ void add_kernel(const float* __restrict__ a, int lda, const float* __restrict__ b, int ldb, float* __restrict__ y,
                           int ldy, int64_t rows, int C) {
  const int64_t total = rows * C;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / C;
    const int c = (int)(i - r * C);
    y[r * ldy + c] = a[r * lda + c] + b[r * ldb + c];
  }
}
__global__ void maxpool_fwd_kernel(const float* __restrict__ x, int ldx, int H, int W, int k, int stride, int pad,
                                   float* __restrict__ y, int ldy, int Ho, int Wo, uint8_t* __restrict__ argmax, int64_t rows_out,
                                   int C) {
  const int64_t total = rows_out * C;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / C;
    const int c = (int)(i - r * C);
    const int64_t n = r / ((int64_t)Ho * Wo);
    const int rem = (int)(r - n * (int64_t)Ho * Wo);
    const int ho = rem / Wo, wo = rem - ho * Wo;
    float best = -INFINITY;
    int bi = 0;
    for (int a = 0; a < k; ++a) {
      const int hi = ho * stride - pad + a;
      for (int b = 0; b < k; ++b) {
        const int wi = wo * stride - pad + b;
        float f;
        if (hi >= 0 && hi < H && wi >= 0 && wi < W) f = x[((n * H + hi) * (int64_t)W + wi) * ldx + c];
        else f = pad > 0 ? 0.f : -INFINITY;   // explicit ZeroPadding2D contributes zeros (pool.cu)
        if (f > best) {
          best = f;
          bi = a * k + b;
        }
      }
    }
    y[r * ldy + c] = best;
    if (argmax) argmax[r * C + c] = (uint8_t)bi;
  }
}
__global__ void maxpool_bwd_kernel(const float* __restrict__ dy, int lddy, int Ho, int Wo, const uint8_t* __restrict__ argmax, int k,
                                   int stride, int pad, const float* __restrict__ res, int ldr, float* __restrict__ dx, int lddx,
                                   int H, int W, int64_t rows_in, int C) {
  const int64_t total = rows_in * C;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / C;
    const int c = (int)(i - r * C);
    const int64_t n = r / ((int64_t)H * W);
    const int rem = (int)(r - n * (int64_t)H * W);
    const int h = rem / W, w = rem - h * W;
    float acc = 0.f;
    int ho_hi = (h + pad) / stride;
    int ho_lo = (h + pad - k + stride) / stride;
    if (h + pad - k + 1 <= 0) ho_lo = 0;
    int wo_hi = (w + pad) / stride;
    int wo_lo = (w + pad - k + stride) / stride;
    if (w + pad - k + 1 <= 0) wo_lo = 0;
    for (int ho = ho_lo; ho <= ho_hi && ho < Ho; ++ho) {
      const int a = h - (ho * stride - pad);
      for (int wo = wo_lo; wo <= wo_hi && wo < Wo; ++wo) {
        const int b = w - (wo * stride - pad);
        const int64_t ro = (n * Ho + ho) * (int64_t)Wo + wo;
        if ((int)argmax[ro * C + c] == a * k + b) acc += dy[ro * lddy + c];
      }
    }
    if (res) acc += res[r * ldr + c];
    dx[r * lddx + c] = acc;
  }
}
// stem wgrad post-processing on the fp32 weights (bn.cu stem_wgrad_post_kernel multiplies by the bf16 weight copy)
__global__ void stem_wgrad_post_kernel(float* __restrict__ dw8, const float* __restrict__ w, int taps, int cpad, int cimg,
                                       float* __restrict__ dbeta) {
  __shared__ double sm[4][256];
  double acc[4] = {0.0, 0.0, 0.0, 0.0};
  for (int t = threadIdx.x; t < taps; t += blockDim.x) {
    const double g = dw8[(int64_t)t * cpad + cimg];
    for (int c = 0; c < cimg; ++c) acc[c] += (double)w[(int64_t)t * cpad + c] * g;
  }
  for (int c = 0; c < 4; ++c) sm[c][threadIdx.x] = acc[c];
  __syncthreads();
  if (threadIdx.x < cimg) {
    double a = 0.0;
    for (int i = 0; i < (int)blockDim.x; ++i) a += sm[threadIdx.x][i];
    if (dbeta) dbeta[threadIdx.x] = (float)a;
  }
  __syncthreads();
  for (int t = threadIdx.x; t < taps; t += blockDim.x)
    for (int c = cimg; c < cpad; ++c) dw8[(int64_t)t * cpad + c] = 0.f;
}

// ---- launchers used by the C ABI dispatch --------------------------------------------------------------------------
int bn_apply(const stp_tensor* x, const float* coef, int relu, int up, const stp_tensor* y, cudaStream_t st) {
  const int64_t rows = pixels(x);
  bn_apply_kernel<<<grid_for(rows * x->c), 256, 0, st>>>((const float*)x->ptr, x->ld, coef, relu, up, (float*)y->ptr, y->ld, x->h,
                                                         x->w, rows, x->c);
  return check_launch("f32 bn_apply");
}
int bwd_apply(int mode, const stp_tensor* dy, const stp_tensor* x, const float* coef, const float* bcoef, int relu, int pool,
              const stp_tensor* res, const stp_tensor* dx, cudaStream_t st) {
  const int64_t rows = pixels(x);
  const float* rp = res ? (const float*)res->ptr : nullptr;
  const int ldr = res ? res->ld : 0;
  if (mode == 0)
    bwd_apply_kernel<0><<<grid_for(rows * x->c), 256, 0, st>>>((const float*)dy->ptr, dy->ld, (const float*)x->ptr, x->ld, coef, bcoef,
                                                               relu, pool, rp, ldr, (float*)dx->ptr, dx->ld, x->h, x->w, rows, x->c);
  else
    bwd_apply_kernel<1><<<grid_for(rows * x->c), 256, 0, st>>>((const float*)dy->ptr, dy->ld, (const float*)x->ptr, x->ld, coef, bcoef,
                                                               relu, pool, rp, ldr, (float*)dx->ptr, dx->ld, x->h, x->w, rows, x->c);
  return check_launch("f32 bwd_apply");
}
int stem_prep(const uint8_t* img, int64_t rows, int cimg, const float* coef, const stp_tensor* y, cudaStream_t st) {
  stem_prep_kernel<<<grid_for(rows * y->c), 256, 0, st>>>(img, rows, cimg, coef, (float*)y->ptr, y->ld, y->c);
  return check_launch("f32 stem_prep");
}
int copy_up(const stp_tensor* x, int up, const stp_tensor* y, cudaStream_t st) {
  const int64_t rows = pixels(x);
  copy_up_kernel<<<grid_for(rows * x->c), 256, 0, st>>>((const float*)x->ptr, x->ld, up, (float*)y->ptr, y->ld, x->h, x->w, rows, x->c);
  return check_launch("f32 copy_up");
}
int add(const stp_tensor* a, const stp_tensor* b, const stp_tensor* y, cudaStream_t st) {
  const int64_t rows = pixels(a);
  add_kernel<<<grid_for(rows * a->c), 256, 0, st>>>((const float*)a->ptr, a->ld, (const float*)b->ptr, b->ld, (float*)y->ptr, y->ld, rows,
                                                    a->c);
  return check_launch("f32 add");
}
int maxpool_fwd(const stp_tensor* x, int k, int stride, int pad, const stp_tensor* y, uint8_t* argmax, cudaStream_t st) {
  const int64_t rows = pixels(y);
  maxpool_fwd_kernel<<<grid_for(rows * x->c), 256, 0, st>>>((const float*)x->ptr, x->ld, x->h, x->w, k, stride, pad, (float*)y->ptr,
                                                            y->ld, y->h, y->w, argmax, rows, x->c);
  return check_launch("f32 maxpool_fwd");
}
int maxpool_bwd(const stp_tensor* dy, const uint8_t* argmax, int k, int stride, int pad, const stp_tensor* res, const stp_tensor* dx,
                cudaStream_t st) {
  const int64_t rows = pixels(dx);
  maxpool_bwd_kernel<<<grid_for(rows * dx->c), 256, 0, st>>>((const float*)dy->ptr, dy->ld, dy->h, dy->w, argmax, k, stride, pad,
                                                             res ? (const float*)res->ptr : nullptr, res ? res->ld : 0,
                                                             (float*)dx->ptr, dx->ld, dx->h, dx->w, rows, dx->c);
  return check_launch("f32 maxpool_bwd");
}
int stem_wgrad_post(float* dw8, const float* w, int taps, int cpad, int cimg, float* dbeta, cudaStream_t st) {
  stem_wgrad_post_kernel<<<1, 256, 0, st>>>(dw8, w, taps, cpad, cimg, dbeta);
  return check_launch("f32 stem_wgrad_post");
}

}  // namespace f32
}  // namespace stp
