// K3/K6: BatchNormalization (+ReLU, + fused 2x nearest upsample) forward / backward, stem prep.
// HBM-bound elementwise / reduction kernels: 16-byte vector access, deterministic two-stage reductions.
// Semantics: keras.layers.BatchNormalization training mode -- biased batch
// variance, moving stats with Bessel-corrected variance.
#include "common.cuh"

namespace stp {

// thread geometry for a [rows, C] tensor: CV = C/8 vectors per row, RPI rows per block iteration
struct RowGeom {
  int cv, rpi, threads, nblk;
  int64_t rows_per_blk;
};
static RowGeom geom(int64_t rows, int c) {
  RowGeom g;
  g.cv = c / 8;
  g.rpi = g.cv >= 256 ? 1 : 256 / g.cv;
  g.threads = g.rpi * (g.cv > 256 ? 256 : g.cv);
  int64_t nb = (rows + (int64_t)g.rpi * 8 - 1) / ((int64_t)g.rpi * 8);
  if (nb < 1) nb = 1;
  if (nb > STP_BN_MAX_PARTIALS) nb = STP_BN_MAX_PARTIALS;
  g.nblk = (int)nb;
  g.rows_per_blk = (rows + nb - 1) / nb;
  return g;
}

// ---------------------------------------------------------------------------------------------
// stats: partial[0][blk][c] = sum x, partial[1][blk][c] = sum x^2
// MODE 0: plain stats of x.   MODE 1: bn backward reduce (sum g, sum g*xhat), dy optional 2x2 pooled.
// ---------------------------------------------------------------------------------------------
template <int MODE>
__global__ void __launch_bounds__(256) reduce_rows_kernel(const __nv_bfloat16* __restrict__ x, int ldx,
                                                          const __nv_bfloat16* __restrict__ dy, int lddy,
                                                          const float* __restrict__ coef, int relu, int pool,
                                                          int H, int W, int64_t rows, int C, int cv, int rpi,
                                                          int64_t rows_per_blk, float* __restrict__ partial) {
  extern __shared__ float sm[];  // [2][rpi][C]
  const int nv = cv > 256 ? 256 : cv;
  const int v0 = threadIdx.x % nv;
  const int rl = threadIdx.x / nv;
  const int64_t r_begin = (int64_t)blockIdx.x * rows_per_blk;
  int64_t r_end = r_begin + rows_per_blk;
  if (r_end > rows) r_end = rows;
  for (int v = v0; v < cv; v += nv) {
    float s0[8], s1[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) s0[i] = s1[i] = 0.f;
    float mean[8], invstd[8], scale[8], shift[8];
    if (MODE == 1) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        int c = v * 8 + i;
        mean[i] = coef[c];
        invstd[i] = coef[C + c];
        scale[i] = coef[2 * C + c];
        shift[i] = coef[3 * C + c];
      }
    }
    for (int64_t r = r_begin + rl; r < r_end; r += rpi) {
      float xf[8];
      unpack8(ld8(x + r * ldx + v * 8), xf);
      if (MODE == 0) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          s0[i] += xf[i];
          s1[i] += xf[i] * xf[i];
        }
      } else {
        float g[8];
        if (pool == 2) {
          int64_t n = r / ((int64_t)H * W);
          int rem = (int)(r - n * (int64_t)H * W);
          int h = rem / W, w = rem - h * W;
          const __nv_bfloat16* base = dy + ((n * 2 * H + 2 * h) * (int64_t)(2 * W) + 2 * w) * lddy + v * 8;
          float t[8];
          unpack8(ld8(base), g);
          unpack8(ld8(base + lddy), t);
#pragma unroll
          for (int i = 0; i < 8; ++i) g[i] += t[i];
          unpack8(ld8(base + (int64_t)2 * W * lddy), t);
#pragma unroll
          for (int i = 0; i < 8; ++i) g[i] += t[i];
          unpack8(ld8(base + (int64_t)2 * W * lddy + lddy), t);
#pragma unroll
          for (int i = 0; i < 8; ++i) g[i] += t[i];
        } else {
          unpack8(ld8(dy + r * lddy + v * 8), g);
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float gi = g[i];
          if (relu && !(xf[i] * scale[i] + shift[i] > 0.f)) gi = 0.f;
          s0[i] += gi;
          s1[i] += gi * ((xf[i] - mean[i]) * invstd[i]);
        }
      }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      sm[(0 * rpi + rl) * C + v * 8 + i] = s0[i];
      sm[(1 * rpi + rl) * C + v * 8 + i] = s1[i];
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < 2 * C; c += blockDim.x) {
    int which = c / C, ch = c - which * C;
    float a = 0.f;
    for (int r = 0; r < rpi; ++r) a += sm[(which * rpi + r) * C + ch];
    partial[((int64_t)which * gridDim.x + blockIdx.x) * C + ch] = a;
  }
}

__global__ void stats_u8_kernel(const uint8_t* __restrict__ x, int64_t rows, int C, int64_t rows_per_blk,
                                float* __restrict__ partial) {
  // C <= 4; one thread per pixel, block reduce
  __shared__ float sm[2][4][8];
  const int64_t r_begin = (int64_t)blockIdx.x * rows_per_blk;
  int64_t r_end = r_begin + rows_per_blk;
  if (r_end > rows) r_end = rows;
  float s0[4] = {0, 0, 0, 0}, s1[4] = {0, 0, 0, 0};
  for (int64_t r = r_begin + threadIdx.x; r < r_end; r += blockDim.x) {
    for (int c = 0; c < C; ++c) {
      float v = (float)x[r * C + c];
      s0[c] += v;
      s1[c] += v * v;
    }
  }
  for (int c = 0; c < 4; ++c) {
    s0[c] = warp_sum(s0[c]);
    s1[c] = warp_sum(s1[c]);
  }
  int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0)
    for (int c = 0; c < 4; ++c) {
      sm[0][c][warp] = s0[c];
      sm[1][c][warp] = s1[c];
    }
  __syncthreads();
  if (threadIdx.x < 2 * C) {
    int which = threadIdx.x / C, ch = threadIdx.x % C;
    float a = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) a += sm[which][ch][w];
    partial[((int64_t)which * gridDim.x + blockIdx.x) * C + ch] = a;
  }
}

// Sum the nblk block partials of 8 channels with 128 threads per channel (fixed order -> deterministic), in double.
// blockDim = (8 channels, 128 partial groups); returns the two sums to the threads with ty == 0.
constexpr int kFinCh = 8, kFinGroups = 128;
__device__ __forceinline__ void reduce_partials(const float* __restrict__ partial, int nblk, int C, int c, double& s,
                                                double& ss) {
  __shared__ double red[2][kFinGroups][kFinCh + 1];
  const int tx = threadIdx.x, ty = threadIdx.y;
  double a = 0.0, b = 0.0;
  if (c < C) {
#pragma unroll 4
    for (int k = ty; k < nblk; k += kFinGroups) {
      a += (double)partial[(int64_t)k * C + c];
      b += (double)partial[((int64_t)nblk + k) * C + c];
    }
  }
  red[0][ty][tx] = a;
  red[1][ty][tx] = b;
  __syncthreads();
  // tree over the 128 groups (fixed shape -> deterministic)
  for (int half = kFinGroups / 2; half >= 1; half >>= 1) {
    if (ty < half) {
      red[0][ty][tx] += red[0][ty + half][tx];
      red[1][ty][tx] += red[1][ty + half][tx];
    }
    __syncthreads();
  }
  s = red[0][0][tx];
  ss = red[1][0][tx];
}

__global__ void __launch_bounds__(1024) bn_finalize_kernel(const float* __restrict__ partial, int nblk, int C,
                                                           double inv_count, double bessel,
                                                           const float* __restrict__ gamma,
                                                           const float* __restrict__ beta, float eps, float momentum,
                                                           float* __restrict__ mov_mean, float* __restrict__ mov_var,
                                                           float* __restrict__ coef) {
  const int c = blockIdx.x * kFinCh + threadIdx.x;
  double s, ss;
  reduce_partials(partial, nblk, C, c, s, ss);
  if (threadIdx.y != 0 || c >= C) return;
  double mean = s * inv_count;
  double var = ss * inv_count - mean * mean;
  if (var < 0.0) var = 0.0;
  double invstd = rsqrt(var + (double)eps);
  float g = gamma ? gamma[c] : 1.f;
  float b = beta ? beta[c] : 0.f;
  float scale = g * (float)invstd;
  coef[c] = (float)mean;
  coef[C + c] = (float)invstd;
  coef[2 * C + c] = scale;
  coef[3 * C + c] = b - (float)mean * scale;
  if (mov_mean) {
    mov_mean[c] = mov_mean[c] * momentum + (float)mean * (1.f - momentum);
    mov_var[c] = mov_var[c] * momentum + (float)(var * bessel) * (1.f - momentum);
  }
}

__global__ void bn_coef_infer_kernel(const float* gamma, const float* beta, const float* mm, const float* mv,
                                     float eps, int C, float* coef) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float invstd = rsqrtf(mv[c] + eps);
  float scale = (gamma ? gamma[c] : 1.f) * invstd;
  coef[c] = mm[c];
  coef[C + c] = invstd;
  coef[2 * C + c] = scale;
  coef[3 * C + c] = (beta ? beta[c] : 0.f) - mm[c] * scale;
}

__global__ void __launch_bounds__(1024) bn_bwd_finalize_kernel(const float* __restrict__ partial, int nblk, int C,
                                                               double inv_count, const float* __restrict__ coef,
                                                               float* __restrict__ dgamma, float* __restrict__ dbeta,
                                                               float* __restrict__ bcoef) {
  const int c = blockIdx.x * kFinCh + threadIdx.x;
  double s, ss;
  reduce_partials(partial, nblk, C, c, s, ss);
  if (threadIdx.y != 0 || c >= C) return;
  if (dbeta) dbeta[c] = (float)s;
  if (dgamma) dgamma[c] = (float)ss;
  double mean = coef[c], invstd = coef[C + c], a = coef[2 * C + c];
  double b = -a * invstd * ss * inv_count;
  double cc = -a * s * inv_count - b * mean;
  bcoef[c] = (float)a;
  bcoef[C + c] = (float)b;
  bcoef[2 * C + c] = (float)cc;
}

// y = [relu](x*scale+shift), optional 2x nearest upsample on write
__global__ void __launch_bounds__(256) bn_apply_kernel(const __nv_bfloat16* __restrict__ x, int ldx,
                                                       const float* __restrict__ coef, int relu, int up,
                                                       __nv_bfloat16* __restrict__ y, int ldy, int H, int W,
                                                       int64_t rows, int C, int cv) {
  int64_t total = rows * cv;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t r = i / cv;
    int v = (int)(i - r * cv);
    float f[8];
    unpack8(ld8(x + r * ldx + v * 8), f);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      int c = v * 8 + k;
      float t = f[k] * __ldg(coef + 2 * C + c) + __ldg(coef + 3 * C + c);
      f[k] = (relu && !(t > 0.f)) ? 0.f : t;
    }
    bf16x8 o = pack8(f);
    if (up == 1) {
      st8(y + r * ldy + v * 8, o);
    } else {
      int64_t n = r / ((int64_t)H * W);
      int rem = (int)(r - n * (int64_t)H * W);
      int h = rem / W, w = rem - h * W;
      __nv_bfloat16* base = y + ((n * 2 * H + 2 * h) * (int64_t)(2 * W) + 2 * w) * ldy + v * 8;
      st8(base, o);
      st8(base + ldy, o);
      st8(base + (int64_t)2 * W * ldy, o);
      st8(base + (int64_t)2 * W * ldy + ldy, o);
    }
  }
}

// MODE 0: dx = a*g + b*x + cc (+res) with g = dy masked by relu(bn(x)).  MODE 1: relu bwd, mask from y=x>0.
template <int MODE>
__global__ void __launch_bounds__(256) bwd_apply_kernel(const __nv_bfloat16* __restrict__ dy, int lddy,
                                                        const __nv_bfloat16* __restrict__ x, int ldx,
                                                        const float* __restrict__ coef,
                                                        const float* __restrict__ bcoef, int relu, int pool,
                                                        const __nv_bfloat16* __restrict__ res, int ldr,
                                                        __nv_bfloat16* __restrict__ dx, int lddx, int H, int W,
                                                        int64_t rows, int C, int cv) {
  int64_t total = rows * cv;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t r = i / cv;
    int v = (int)(i - r * cv);
    float xf[8], g[8];
    unpack8(ld8(x + r * ldx + v * 8), xf);
    if (pool == 2) {
      int64_t n = r / ((int64_t)H * W);
      int rem = (int)(r - n * (int64_t)H * W);
      int h = rem / W, w = rem - h * W;
      const __nv_bfloat16* base = dy + ((n * 2 * H + 2 * h) * (int64_t)(2 * W) + 2 * w) * lddy + v * 8;
      float t[8];
      unpack8(ld8(base), g);
      unpack8(ld8(base + lddy), t);
#pragma unroll
      for (int k = 0; k < 8; ++k) g[k] += t[k];
      unpack8(ld8(base + (int64_t)2 * W * lddy), t);
#pragma unroll
      for (int k = 0; k < 8; ++k) g[k] += t[k];
      unpack8(ld8(base + (int64_t)2 * W * lddy + lddy), t);
#pragma unroll
      for (int k = 0; k < 8; ++k) g[k] += t[k];
    } else {
      unpack8(ld8(dy + r * lddy + v * 8), g);
    }
    float rf[8];
    if (res) unpack8(ld8(res + r * ldr + v * 8), rf);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      int c = v * 8 + k;
      float o;
      if (MODE == 0) {
        float gi = g[k];
        if (relu && !(xf[k] * __ldg(coef + 2 * C + c) + __ldg(coef + 3 * C + c) > 0.f)) gi = 0.f;
        o = __ldg(bcoef + c) * gi + __ldg(bcoef + C + c) * xf[k] + __ldg(bcoef + 2 * C + c);
      } else {
        o = xf[k] > 0.f ? g[k] : 0.f;
      }
      if (res) o += rf[k];
      g[k] = o;
    }
    st8(dx + r * lddx + v * 8, pack8(g));
  }
}

__global__ void stem_prep_kernel(const uint8_t* __restrict__ img, int64_t rows, int cimg,
                                 const float* __restrict__ coef, __nv_bfloat16* __restrict__ y, int ldy, int C) {
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += (int64_t)gridDim.x * blockDim.x) {
    for (int v = 0; v < C / 8; ++v) {
      float f[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        int c = v * 8 + k;
        float t = 0.f;
        if (c < cimg) t = (float)img[r * cimg + c] * __ldg(coef + 2 * cimg + c) + __ldg(coef + 3 * cimg + c);
        else if (c == cimg) t = 1.f;
        f[k] = t;
      }
      st8(y + r * ldy + v * 8, pack8(f));
    }
  }
}

__global__ void copy_up_kernel(const __nv_bfloat16* __restrict__ x, int ldx, int up, __nv_bfloat16* __restrict__ y,
                               int ldy, int H, int W, int64_t rows, int cv) {
  int64_t total = rows * cv;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t r = i / cv;
    int v = (int)(i - r * cv);
    bf16x8 o = ld8(x + r * ldx + v * 8);
    if (up == 1) {
      st8(y + r * ldy + v * 8, o);
    } else {
      int64_t n = r / ((int64_t)H * W);
      int rem = (int)(r - n * (int64_t)H * W);
      int h = rem / W, w = rem - h * W;
      for (int a = 0; a < up; ++a)
        for (int b = 0; b < up; ++b)
          st8(y + ((n * up * H + up * h + a) * (int64_t)(up * W) + up * w + b) * ldy + v * 8, o);
    }
  }
}

__global__ void add_kernel(const __nv_bfloat16* __restrict__ a, int lda, const __nv_bfloat16* __restrict__ b, int ldb,
                           __nv_bfloat16* __restrict__ y, int ldy, int64_t rows, int cv) {
  int64_t total = rows * cv;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t r = i / cv;
    int v = (int)(i - r * cv);
    float fa[8], fb[8];
    unpack8(ld8(a + r * lda + v * 8), fa);
    unpack8(ld8(b + r * ldb + v * 8), fb);
#pragma unroll
    for (int k = 0; k < 8; ++k) fa[k] += fb[k];
    st8(y + r * ldy + v * 8, pack8(fa));
  }
}

static int ew_grid(int64_t total) {
  int64_t b = (total + 255) / 256;
  int64_t cap = (int64_t)kNumSMs * 16;
  return (int)(b < cap ? (b < 1 ? 1 : b) : cap);
}

}  // namespace stp

using namespace stp;

extern "C" int32_t stp_bn_nblk(int64_t rows, int32_t c) { return geom(rows, c < 8 ? 8 : c).nblk; }

extern "C" int stp_bn_stats(const stp_tensor* x, float* partial, stp_stream stream) {
  STP_REQUIRE(x && partial, "bn_stats: null");
  cudaStream_t st = (cudaStream_t)stream;
  int64_t rows = pixels(x);
  if (x->dtype == STP_U8) {
    STP_REQUIRE(x->c <= 4 && x->ld == x->c, "bn_stats u8: c<=4 dense only");
    int nblk = geom(rows, 8).nblk;  // same rule as stp_bn_nblk(rows, c<8)
    int64_t rpb = (rows + nblk - 1) / nblk;
    stats_u8_kernel<<<nblk, 256, 0, st>>>((const uint8_t*)x->ptr, rows, x->c, rpb, partial);
    return check_launch("stats_u8");
  }
  STP_REQUIRE(vec_ok(x), "bn_stats: tensor must be bf16, c%%8==0, ld%%8==0, 16B aligned");
  STP_REQUIRE(x->c <= 2048, "bn_stats: c too large");
  RowGeom g = geom(rows, x->c);
  size_t smem = (size_t)2 * g.rpi * x->c * sizeof(float);
  reduce_rows_kernel<0><<<g.nblk, g.threads, smem, st>>>((const __nv_bfloat16*)x->ptr, x->ld, nullptr, 0, nullptr, 0,
                                                          1, x->h, x->w, rows, x->c, g.cv, g.rpi, g.rows_per_blk,
                                                          partial);
  return check_launch("bn_stats");
}

extern "C" int stp_bn_finalize(const float* partial, int32_t nblk, int32_t c, int64_t count, const float* gamma,
                               const float* beta, float eps, float momentum, float* moving_mean, float* moving_var,
                               float* coef, stp_stream stream) {
  STP_REQUIRE(partial && coef && nblk > 0 && c > 0 && count > 0, "bn_finalize: bad args");
  double bessel = count > 1 ? (double)count / (double)(count - 1) : 1.0;
  bn_finalize_kernel<<<(c + kFinCh - 1) / kFinCh, dim3(kFinCh, kFinGroups), 0, (cudaStream_t)stream>>>(partial, nblk, c, 1.0 / (double)count, bessel,
                                                                         gamma, beta, eps, momentum, moving_mean,
                                                                         moving_var, coef);
  return check_launch("bn_finalize");
}

extern "C" int stp_bn_coef_infer(const float* gamma, const float* beta, const float* moving_mean,
                                 const float* moving_var, float eps, int32_t c, float* coef, stp_stream stream) {
  STP_REQUIRE(moving_mean && moving_var && coef, "bn_coef_infer: null");
  bn_coef_infer_kernel<<<(c + 127) / 128, 128, 0, (cudaStream_t)stream>>>(gamma, beta, moving_mean, moving_var, eps, c,
                                                                           coef);
  return check_launch("bn_coef_infer");
}

extern "C" int stp_bn_apply(const stp_tensor* x, const float* coef, int32_t relu, int32_t up, const stp_tensor* y,
                            stp_stream stream) {
  STP_REQUIRE(vec_ok(x) && vec_ok(y) && coef, "bn_apply: bad tensors");
  STP_REQUIRE(up == 1 || up == 2, "bn_apply: up must be 1 or 2");
  STP_REQUIRE(y->c == x->c && y->n == x->n && y->h == x->h * up && y->w == x->w * up, "bn_apply: shape mismatch");
  int64_t rows = pixels(x);
  int cv = x->c / 8;
  bn_apply_kernel<<<ew_grid(rows * cv), 256, 0, (cudaStream_t)stream>>>(
      (const __nv_bfloat16*)x->ptr, x->ld, coef, relu, up, (__nv_bfloat16*)y->ptr, y->ld, x->h, x->w, rows, x->c, cv);
  return check_launch("bn_apply");
}

extern "C" int stp_bn_bwd_reduce(const stp_tensor* dy, const stp_tensor* x, const float* coef, int32_t relu,
                                 int32_t pool, float* partial, stp_stream stream) {
  STP_REQUIRE(vec_ok(dy) && vec_ok(x) && coef && partial, "bn_bwd_reduce: bad tensors");
  STP_REQUIRE(pool == 1 || pool == 2, "bn_bwd_reduce: pool must be 1 or 2");
  STP_REQUIRE(dy->c == x->c && dy->h == x->h * pool && dy->w == x->w * pool && dy->n == x->n,
              "bn_bwd_reduce: shape mismatch");
  STP_REQUIRE(x->c <= 2048, "bn_bwd_reduce: c too large");
  int64_t rows = pixels(x);
  RowGeom g = geom(rows, x->c);
  size_t smem = (size_t)2 * g.rpi * x->c * sizeof(float);
  reduce_rows_kernel<1><<<g.nblk, g.threads, smem, (cudaStream_t)stream>>>(
      (const __nv_bfloat16*)x->ptr, x->ld, (const __nv_bfloat16*)dy->ptr, dy->ld, coef, relu, pool, x->h, x->w, rows,
      x->c, g.cv, g.rpi, g.rows_per_blk, partial);
  return check_launch("bn_bwd_reduce");
}

extern "C" int stp_bn_bwd_finalize(const float* partial, int32_t nblk, int32_t c, int64_t count, const float* coef,
                                   float* dgamma, float* dbeta, float* bcoef, stp_stream stream) {
  STP_REQUIRE(partial && coef && bcoef && nblk > 0, "bn_bwd_finalize: bad args");
  bn_bwd_finalize_kernel<<<(c + kFinCh - 1) / kFinCh, dim3(kFinCh, kFinGroups), 0, (cudaStream_t)stream>>>(partial, nblk, c, 1.0 / (double)count,
                                                                             coef, dgamma, dbeta, bcoef);
  return check_launch("bn_bwd_finalize");
}

static int bwd_apply_common(int mode, const stp_tensor* dy, const stp_tensor* x, const float* coef,
                            const float* bcoef, int relu, int pool, const stp_tensor* residual, const stp_tensor* dx,
                            stp_stream stream) {
  STP_REQUIRE(vec_ok(dy) && vec_ok(x) && vec_ok(dx), "bwd_apply: bad tensors");
  STP_REQUIRE(pool == 1 || pool == 2, "bwd_apply: pool must be 1 or 2");
  STP_REQUIRE(dy->c == x->c && dx->c == x->c && dy->h == x->h * pool && dy->w == x->w * pool && dx->h == x->h &&
                  dx->w == x->w && dx->n == x->n && dy->n == x->n,
              "bwd_apply: shape mismatch");
  if (residual) STP_REQUIRE(vec_ok(residual) && residual->c == x->c && pixels(residual) == pixels(x), "bwd_apply: bad residual");
  int64_t rows = pixels(x);
  int cv = x->c / 8;
  const __nv_bfloat16* rp = residual ? (const __nv_bfloat16*)residual->ptr : nullptr;
  int ldr = residual ? residual->ld : 0;
  if (mode == 0)
    bwd_apply_kernel<0><<<ew_grid(rows * cv), 256, 0, (cudaStream_t)stream>>>(
        (const __nv_bfloat16*)dy->ptr, dy->ld, (const __nv_bfloat16*)x->ptr, x->ld, coef, bcoef, relu, pool, rp, ldr,
        (__nv_bfloat16*)dx->ptr, dx->ld, x->h, x->w, rows, x->c, cv);
  else
    bwd_apply_kernel<1><<<ew_grid(rows * cv), 256, 0, (cudaStream_t)stream>>>(
        (const __nv_bfloat16*)dy->ptr, dy->ld, (const __nv_bfloat16*)x->ptr, x->ld, coef, bcoef, relu, pool, rp, ldr,
        (__nv_bfloat16*)dx->ptr, dx->ld, x->h, x->w, rows, x->c, cv);
  return check_launch("bwd_apply");
}

extern "C" int stp_bn_bwd_apply(const stp_tensor* dy, const stp_tensor* x, const float* coef, const float* bcoef,
                                int32_t relu, int32_t pool, const stp_tensor* residual, const stp_tensor* dx,
                                stp_stream stream) {
  STP_REQUIRE(coef && bcoef, "bn_bwd_apply: null coef");
  return bwd_apply_common(0, dy, x, coef, bcoef, relu, pool, residual, dx, stream);
}

extern "C" int stp_relu_bwd(const stp_tensor* dy, const stp_tensor* y, int32_t pool, const stp_tensor* residual,
                            const stp_tensor* dx, stp_stream stream) {
  return bwd_apply_common(1, dy, y, nullptr, nullptr, 1, pool, residual, dx, stream);
}

extern "C" int stp_stem_prep(const uint8_t* img, int32_t n, int32_t h, int32_t w, int32_t c_img, const float* coef,
                             const stp_tensor* y, stp_stream stream) {
  STP_REQUIRE(img && coef && vec_ok(y), "stem_prep: bad args");
  STP_REQUIRE(c_img < y->c && y->n == n && y->h == h && y->w == w, "stem_prep: shape mismatch");
  int64_t rows = (int64_t)n * h * w;
  stem_prep_kernel<<<ew_grid(rows), 256, 0, (cudaStream_t)stream>>>(img, rows, c_img, coef, (__nv_bfloat16*)y->ptr,
                                                                    y->ld, y->c);
  return check_launch("stem_prep");
}

extern "C" int stp_copy_up(const stp_tensor* x, int32_t up, const stp_tensor* y, stp_stream stream) {
  STP_REQUIRE(vec_ok(x) && vec_ok(y) && up >= 1, "copy_up: bad tensors");
  STP_REQUIRE(y->c == x->c && y->n == x->n && y->h == x->h * up && y->w == x->w * up, "copy_up: shape mismatch");
  int64_t rows = pixels(x);
  int cv = x->c / 8;
  copy_up_kernel<<<ew_grid(rows * cv), 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)x->ptr, x->ld, up,
                                                                       (__nv_bfloat16*)y->ptr, y->ld, x->h, x->w, rows,
                                                                       cv);
  return check_launch("copy_up");
}

extern "C" int stp_add(const stp_tensor* a, const stp_tensor* b, const stp_tensor* y, stp_stream stream) {
  STP_REQUIRE(vec_ok(a) && vec_ok(b) && vec_ok(y), "add: bad tensors");
  STP_REQUIRE(a->c == b->c && a->c == y->c && pixels(a) == pixels(b) && pixels(a) == pixels(y), "add: shape mismatch");
  int64_t rows = pixels(a);
  int cv = a->c / 8;
  add_kernel<<<ew_grid(rows * cv), 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)a->ptr, a->ld,
                                                                   (const __nv_bfloat16*)b->ptr, b->ld,
                                                                   (__nv_bfloat16*)y->ptr, y->ld, rows, cv);
  return check_launch("add");
}

// ---- stem wgrad post-processing ---------------------------------------------------------------
// conv0 runs on an input padded to 8 channels whose channel `c_img` is constant 1 inside the image, so the
// wgrad column dw8[co][r][s][c_img] = G[co][r][s] = sum of dY over the positions where tap (r,s) is inside
// the image.  d(beta of bn_data)[c] = sum_{co,r,s} W[co][r][s][c] * G[co][r][s]  (== sum of conv0's dgrad
// over valid pixels, without ever running that dgrad).  Padded columns of dw8 are then zeroed.
namespace stp {
__global__ void stem_wgrad_post_kernel(float* __restrict__ dw8, const float* __restrict__ w, int taps, int cpad,
                                       int cimg, float* __restrict__ dbeta) {
  __shared__ float sm[4][256];
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int t = threadIdx.x; t < taps; t += blockDim.x) {
    float g = dw8[(int64_t)t * cpad + cimg];
    for (int c = 0; c < cimg; ++c)
      acc[c] += __bfloat162float(__float2bfloat16(w[(int64_t)t * cpad + c])) * g;
  }
  for (int c = 0; c < 4; ++c) sm[c][threadIdx.x] = acc[c];
  __syncthreads();
  if (threadIdx.x < cimg) {
    double a = 0.0;
    for (int i = 0; i < (int)blockDim.x; ++i) a += (double)sm[threadIdx.x][i];
    if (dbeta) dbeta[threadIdx.x] = (float)a;
  }
  __syncthreads();
  for (int t = threadIdx.x; t < taps; t += blockDim.x)
    for (int c = cimg; c < cpad; ++c) dw8[(int64_t)t * cpad + c] = 0.f;
}
}  // namespace stp

extern "C" int stp_stem_wgrad_post(float* dw8, const float* w_master, int32_t cout, int32_t r, int32_t s,
                                   int32_t cin_pad, int32_t c_img, float* dbeta, stp_stream stream) {
  STP_REQUIRE(dw8 && w_master && c_img >= 1 && c_img <= 4 && c_img < cin_pad, "stem_wgrad_post: bad args");
  stp::stem_wgrad_post_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(dw8, w_master, cout * r * s, cin_pad, c_img, dbeta);
  return check_launch("stem_wgrad_post");
}
