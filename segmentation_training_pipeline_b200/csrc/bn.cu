// K3/K6: BatchNormalization (+ReLU, + fused 2x nearest upsample) forward / backward, stem prep.
// HBM-bound elementwise / reduction kernels: 16-byte vector access, deterministic two-stage reductions.
// Semantics: keras.layers.BatchNormalization training mode -- biased batch
// variance, moving stats with Bessel-corrected variance.
#include "common.cuh"
#include "bn_fin.cuh"
#include "conv.h"
#include "f32_path.h"

namespace stp {

// Thread geometry for a [rows, C] tensor streamed row-wise: a thread owns ONE 8-channel vector (its per-channel
// coefficients live in registers) and walks rows; a block covers `rpi` consecutive rows per step, so a block step
// touches rpi*C*2 contiguous bytes (when ld == C).  Blocks own contiguous row ranges (deterministic partials).
struct RowGeom {
  int cv, nv, rpi, threads, nblk;
  int rows_per_blk;
};
static RowGeom geom(int64_t rows, int c, int unroll, int max_blk) {
  RowGeom g;
  g.cv = c / 8;
  g.nv = g.cv > 256 ? 256 : g.cv;
  g.rpi = 256 / g.nv;
  g.threads = g.rpi * g.nv;
  const int64_t step = (int64_t)g.rpi * unroll;
  int64_t nb = (rows + step - 1) / step;
  if (nb < 1) nb = 1;
  if (nb > max_blk) nb = max_blk;
  int64_t rpb = (rows + nb - 1) / nb;
  rpb = (rpb + g.rpi - 1) / g.rpi * g.rpi;
  g.rows_per_blk = (int)rpb;
  g.nblk = (int)((rows + rpb - 1) / rpb);
  return g;
}
// partial blocks of the reduction kernels: enough CTAs to fill the GPU, few enough that the last CTA can finalise
// nblk x 2C partial sums in a couple of microseconds
static int reduce_max_blk(int c) {
  int m = 24576 / (c < 8 ? 8 : c);
  if (m > 2 * kNumSMs) m = 2 * kNumSMs;
  if (m < 32) m = 32;
  return m;
}
static RowGeom reduce_geom(int64_t rows, int c, bool atomic_acc = false) {
  if (!atomic_acc) return geom(rows, c, 4, reduce_max_blk(c));
  // atomic accumulation: every block issues 2*c double atomics + one ticket; keep the total (and the same-address
  // contention, = the block count) bounded
  const int budget = get_option(OPT_BN_BLOCKS) > 0 ? get_option(OPT_BN_BLOCKS) * 1024 : 16384;
  int m = budget / (c < 8 ? 8 : c);
  const int cap = get_option(OPT_BN_BLOCKS) > 0 ? 4 * kNumSMs : 2 * kNumSMs;
  if (m > cap) m = cap;
  if (get_option(OPT_BN_BLOCKS) > 0 && m < kNumSMs) m = kNumSMs;
  if (m < 32) m = 32;
  // wide AND large tensors (MobileNetV2 expansions: 960 channels x 25600 pixels = 49 MB): the atomic budget above would leave
  // 32 CTAs for the whole GPU; one CTA per SM measured 12.1 -> 10.4 ms on the DeepLabV3 step (profiles/r2_deeplab_bench.txt)
  if (m < kNumSMs && rows * (int64_t)c * 2 >= ((int64_t)16 << 20)) m = kNumSMs;
  return geom(rows, c, 4, m);
}

// Executed by every thread of a reduction block after its partials are written: elects the last block of the grid
// (threadfence + atomic ticket) which then sums the nblk partials per channel in a fixed order (deterministic), in
// double, and finalises.  The ticket counter is returned to zero for the next launch / graph replay.
__device__ void last_block_finalize(const FinArgs& f, const float* __restrict__ partial, int nblk, int C, float* smf) {
  __shared__ int is_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned int t = atomicAdd(f.sync, 1u);
    is_last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  double* red = reinterpret_cast<double*>(smf);  // [2][blockDim]
  const int nt = blockDim.x;
  const int CC = C < nt ? C : nt;
  const int G = nt / CC;
  const int cl = threadIdx.x % CC, g = threadIdx.x / CC;
  for (int c0 = 0; c0 < C; c0 += CC) {
    const int c = c0 + cl;
    double a = 0.0, b = 0.0;
    if (g < G && c < C) {
      // L2-coherent loads (other blocks wrote these), 16 independent loads in flight per thread
      const float* p0 = partial + c;
      const float* p1 = partial + (int64_t)nblk * C + c;
      int k = g;
      for (; k + 7 * G < nblk; k += 8 * G) {
        float t0[8], t1[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          t0[j] = __ldcg(p0 + (int64_t)(k + j * G) * C);
          t1[j] = __ldcg(p1 + (int64_t)(k + j * G) * C);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          a += (double)t0[j];
          b += (double)t1[j];
        }
      }
      for (; k < nblk; k += G) {
        a += (double)__ldcg(p0 + (int64_t)k * C);
        b += (double)__ldcg(p1 + (int64_t)k * C);
      }
    }
    __syncthreads();
    red[threadIdx.x] = a;
    red[nt + threadIdx.x] = b;
    __syncthreads();
    if (g == 0 && c < C) {
      for (int j = 1; j < G; ++j) {
        a += red[j * CC + cl];
        b += red[nt + j * CC + cl];
      }
      if (f.mode == 1) fin_forward(f, C, c, a, b); else if (f.mode == 2) fin_backward(f, C, c, a, b); else f.dbeta[c] = (float)a;
    }
  }
  if (threadIdx.x == 0) *f.sync = 0u;
}

// ---------------------------------------------------------------------------------------------
// row reductions: partial[0][blk][c], partial[1][blk][c]
// MODE 0: sum x, sum x^2.   MODE 1: bn backward (sum g, sum g*xhat), g = dy masked by relu; dy optionally the
// gradient of the 2x-upsampled tensor (2x2 summed on the fly).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void load_g(const __nv_bfloat16* __restrict__ dy, int lddy, int pool, int H, int W, int r,
                                       int v, float* g) {
  if (pool == 2) {
    unsigned n = (unsigned)r / (unsigned)(H * W);
    unsigned rem = (unsigned)r - n * (unsigned)(H * W);
    unsigned h = rem / (unsigned)W, w = rem - h * (unsigned)W;
    const __nv_bfloat16* base = dy + (((int64_t)n * 2 * H + 2 * h) * (int64_t)(2 * W) + 2 * w) * lddy + v * 8;
    bf16x8 a0 = ld8(base), a1 = ld8(base + lddy), a2 = ld8(base + (int64_t)2 * W * lddy),
           a3 = ld8(base + (int64_t)2 * W * lddy + lddy);
    float t[8];
    unpack8(a0, g);
    unpack8(a1, t);
#pragma unroll
    for (int i = 0; i < 8; ++i) g[i] += t[i];
    unpack8(a2, t);
#pragma unroll
    for (int i = 0; i < 8; ++i) g[i] += t[i];
    unpack8(a3, t);
#pragma unroll
    for (int i = 0; i < 8; ++i) g[i] += t[i];
  } else {
    unpack8(ld8(dy + (int64_t)r * lddy + v * 8), g);
  }
}

template <int MODE, int POOL>
__global__ void __launch_bounds__(256, 2) reduce_rows_kernel(const __nv_bfloat16* __restrict__ x, int ldx,
                                                          const __nv_bfloat16* __restrict__ dy, int lddy,
                                                          const float* __restrict__ coef, int relu, int pool,
                                                          int H, int W, int rows, int C, int cv, int nv, int rpi,
                                                          int rows_per_blk, float* __restrict__ partial,
                                                          const FinArgs fin) {
  extern __shared__ float sm[];  // [2][groups][C], groups <= 8 (>= 2*blockDim doubles for the finalize)
  pdl_launch_dependents();
  pdl_wait();
  const int v0 = threadIdx.x % nv;
  const int rl = threadIdx.x / nv;
  const int r_begin = blockIdx.x * rows_per_blk;
  int r_end = r_begin + rows_per_blk;
  if (r_end > rows) r_end = rows;
  // rows of one warp can be pre-reduced with shuffles when a warp spans whole rows (nv a power of two <= 32)
  const bool shfl = nv <= 32 && (nv & (nv - 1)) == 0 && (blockDim.x & 31) == 0;
  const int groups = shfl ? (blockDim.x >> 5) : rpi;
  const int grp = shfl ? (threadIdx.x >> 5) : rl;
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
    if (MODE == 0) {
      for (int r = r_begin + rl; r < r_end; r += 8 * rpi) {
        bf16x8 q[8];
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (r + j * rpi < r_end) q[j] = ld8(x + (int64_t)(r + j * rpi) * ldx + v * 8);
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (r + j * rpi < r_end) {
            float xf[8];
            unpack8(q[j], xf);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              s0[i] += xf[i];
              s1[i] += xf[i] * xf[i];
            }
          }
      }
    } else {
      constexpr int U = POOL == 2 ? 2 : 4;
      for (int r = r_begin + rl; r < r_end; r += U * rpi) {
        bf16x8 q[U];
        bf16x8 qg[U];     // POOL == 1: raw dy vectors
        float g2[POOL == 2 ? U : 1][8];  // POOL == 2: 2x2-summed dy
#pragma unroll
        for (int j = 0; j < U; ++j)
          if (r + j * rpi < r_end) {
            q[j] = ld8(x + (int64_t)(r + j * rpi) * ldx + v * 8);
            if (POOL == 2) load_g(dy, lddy, 2, H, W, r + j * rpi, v, g2[POOL == 2 ? j : 0]);
            else qg[j] = ld8(dy + (int64_t)(r + j * rpi) * lddy + v * 8);
          }
#pragma unroll
        for (int j = 0; j < U; ++j)
          if (r + j * rpi < r_end) {
            float xf[8], g[8];
            unpack8(q[j], xf);
            if (POOL == 2) {
#pragma unroll
              for (int i = 0; i < 8; ++i) g[i] = g2[POOL == 2 ? j : 0][i];
            } else {
              unpack8(qg[j], g);
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              float gi = g[i];
              if (relu && !relu_pass(xf[i] * scale[i] + shift[i], relu)) gi = 0.f;
              s0[i] += gi;
              s1[i] += gi * ((xf[i] - mean[i]) * invstd[i]);
            }
          }
      }
    }
    if (shfl) {
      for (int off = 16; off >= nv; off >>= 1) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          s0[i] += __shfl_xor_sync(0xffffffffu, s0[i], off);
          s1[i] += __shfl_xor_sync(0xffffffffu, s1[i], off);
        }
      }
    }
    if (!shfl || (threadIdx.x & 31) < nv) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        sm[(0 * groups + grp) * C + v * 8 + i] = s0[i];
        sm[(1 * groups + grp) * C + v * 8 + i] = s1[i];
      }
    }
  }
  __syncthreads();
  const bool atomic_acc = fin.mode != 0 && fin.acc != nullptr;
  for (int c = threadIdx.x; c < 2 * C; c += blockDim.x) {
    int which = c / C, ch = c - which * C;
    float a = 0.f;
    for (int r = 0; r < groups; ++r) a += sm[(which * groups + r) * C + ch];
    if (atomic_acc) atomicAdd(fin.acc + c, (double)a);
    else partial[((int64_t)which * gridDim.x + blockIdx.x) * C + ch] = a;
  }
  if (atomic_acc) {
    // last block (ticket) finalises straight from the 2*C accumulated doubles and returns them to zero
    __shared__ int is_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) is_last = (atomicAdd(fin.sync, 1u) == gridDim.x - 1) ? 1 : 0;
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      const double s = __ldcg(fin.acc + c), ss = __ldcg(fin.acc + C + c);
      if (fin.mode == 1) fin_forward(fin, C, c, s, ss); else if (fin.mode == 2) fin_backward(fin, C, c, s, ss); else fin.dbeta[c] = (float)s;
      fin.acc[c] = 0.0;
      fin.acc[C + c] = 0.0;
    }
    if (threadIdx.x == 0) *fin.sync = 0u;
  } else if (fin.mode != 0) {
    last_block_finalize(fin, partial, gridDim.x, C, sm);
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

// y = [relu](x*scale+shift), optional 2x nearest upsample on write.  Row-streamed (see RowGeom), 4 loads in flight.
__global__ void __launch_bounds__(256) bn_apply_kernel(const __nv_bfloat16* __restrict__ x, int ldx,
                                                       const float* __restrict__ coef, int relu, int up,
                                                       __nv_bfloat16* __restrict__ y, int ldy, int H, int W,
                                                       int rows, int C, int cv, int nv, int rpi, int rows_per_blk) {
  pdl_launch_dependents();
  pdl_wait();
  const int v0 = threadIdx.x % nv;
  const int rl = threadIdx.x / nv;
  const int r_begin = blockIdx.x * rows_per_blk;
  int r_end = r_begin + rows_per_blk;
  if (r_end > rows) r_end = rows;
  for (int v = v0; v < cv; v += nv) {
    float scale[8], shift[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      scale[i] = coef[2 * C + v * 8 + i];
      shift[i] = coef[3 * C + v * 8 + i];
    }
    for (int r = r_begin + rl; r < r_end; r += 4 * rpi) {
      bf16x8 q[4];
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (r + j * rpi < r_end) q[j] = ld8(x + (int64_t)(r + j * rpi) * ldx + v * 8);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int rr = r + j * rpi;
        if (rr < r_end) {
          float f[8];
          unpack8(q[j], f);
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            float t = f[k] * scale[k] + shift[k];
            f[k] = relu_act(t, relu);
          }
          bf16x8 o = pack8(f);
          if (up == 1) {
            st8(y + (int64_t)rr * ldy + v * 8, o);
          } else {
            unsigned n = (unsigned)rr / (unsigned)(H * W);
            unsigned rem = (unsigned)rr - n * (unsigned)(H * W);
            unsigned h = rem / (unsigned)W, w = rem - h * (unsigned)W;
            __nv_bfloat16* base = y + (((int64_t)n * 2 * H + 2 * h) * (int64_t)(2 * W) + 2 * w) * ldy + v * 8;
            st8(base, o);
            st8(base + ldy, o);
            st8(base + (int64_t)2 * W * ldy, o);
            st8(base + (int64_t)2 * W * ldy + ldy, o);
          }
        }
      }
    }
  }
}

// MODE 0: dx = a*g + b*x + cc (+res) with g = dy masked by relu(bn(x)).  MODE 1: relu bwd, mask from y=x>0.
template <int MODE, int POOL>
__global__ void __launch_bounds__(256, 2) bwd_apply_kernel(const __nv_bfloat16* __restrict__ dy, int lddy,
                                                        const __nv_bfloat16* __restrict__ x, int ldx,
                                                        const float* __restrict__ coef,
                                                        const float* __restrict__ bcoef, int relu, int pool,
                                                        const __nv_bfloat16* __restrict__ res, int ldr,
                                                        __nv_bfloat16* __restrict__ dx, int lddx, int H, int W,
                                                        int rows, int C, int cv, int nv, int rpi, int rows_per_blk) {
  pdl_launch_dependents();
  pdl_wait();
  const int v0 = threadIdx.x % nv;
  const int rl = threadIdx.x / nv;
  const int r_begin = blockIdx.x * rows_per_blk;
  int r_end = r_begin + rows_per_blk;
  if (r_end > rows) r_end = rows;
  for (int v = v0; v < cv; v += nv) {
    float scale[8], shift[8], ca[8], cb[8], cc[8];
    if (MODE == 0) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int c = v * 8 + i;
        scale[i] = coef[2 * C + c];
        shift[i] = coef[3 * C + c];
        ca[i] = bcoef[c];
        cb[i] = bcoef[C + c];
        cc[i] = bcoef[2 * C + c];
      }
    }
    constexpr int U = POOL == 2 ? 2 : 4;
    for (int r = r_begin + rl; r < r_end; r += U * rpi) {
      bf16x8 qx[U], qr[U], qg[U];
      float g2[POOL == 2 ? U : 1][8];
#pragma unroll
      for (int j = 0; j < U; ++j) {
        const int rr = r + j * rpi;
        if (rr < r_end) {
          qx[j] = ld8(x + (int64_t)rr * ldx + v * 8);
          if (res) qr[j] = ld8(res + (int64_t)rr * ldr + v * 8);
          if (POOL == 2) load_g(dy, lddy, 2, H, W, rr, v, g2[POOL == 2 ? j : 0]);
          else qg[j] = ld8(dy + (int64_t)rr * lddy + v * 8);
        }
      }
#pragma unroll
      for (int j = 0; j < U; ++j) {
        const int rr = r + j * rpi;
        if (rr < r_end) {
          float xf[8], rf[8], g[8];
          unpack8(qx[j], xf);
          if (res) unpack8(qr[j], rf);
          if (POOL == 2) {
#pragma unroll
            for (int k = 0; k < 8; ++k) g[k] = g2[POOL == 2 ? j : 0][k];
          } else {
            unpack8(qg[j], g);
          }
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            float o;
            if (MODE == 0) {
              float gi = g[k];
              if (relu && !relu_pass(xf[k] * scale[k] + shift[k], relu)) gi = 0.f;
              o = ca[k] * gi + cb[k] * xf[k] + cc[k];
            } else {
              o = xf[k] > 0.f ? g[k] : 0.f;
            }
            if (res) o += rf[k];
            g[k] = o;
          }
          st8(dx + (int64_t)rr * lddx + v * 8, pack8(g));
        }
      }
    }
  }
}

// s2d == 0: y[n,h,w,C].  s2d == 1: space-to-depth output y[n,h/2,w/2,4*C8] with channel ((h&1)*2 + (w&1))*C8 + c, the
// layout in which the 7x7/2 stem convolution is a stride-1 4x4 convolution (DESIGN.md "stem").
__global__ void stem_prep_kernel(const uint8_t* __restrict__ img, int64_t rows, int cimg,
                                 const float* __restrict__ coef, __nv_bfloat16* __restrict__ y, int ldy, int C, int s2d,
                                 int H, int W) {
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += (int64_t)gridDim.x * blockDim.x) {
    __nv_bfloat16* dst;
    if (s2d) {
      int64_t n = r / ((int64_t)H * W);
      int rem = (int)(r - n * (int64_t)H * W);
      int h = rem / W, w = rem - h * W;
      dst = y + ((n * (H / 2) + (h >> 1)) * (int64_t)(W / 2) + (w >> 1)) * ldy + ((h & 1) * 2 + (w & 1)) * C;
    } else {
      dst = y + r * ldy;
    }
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
      st8(dst + v * 8, pack8(f));
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

// Sums that a producing kernel left in bn.acc ([2][C] doubles, fire-and-forget reductions) -> BatchNorm coefficients (mode 1) or
// dgamma / dbeta / bcoef (mode 2); the accumulators return to zero.  One small launch instead of a fence + ticket in every CTA of
// a kernel with thousands of short-lived CTAs (gemm1x1.cu, dwconv.cu).
__global__ void bn_acc_finalize_kernel(const BnFuse bn, int C, int mode) {
  pdl_launch_dependents();
  pdl_wait();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const double t1 = __ldcg(bn.acc + c), t2 = __ldcg(bn.acc + C + c);
  if (mode == 2) {  // sum g*xhat = invstd * (sum g*x - mean * sum g)
    const double mean = bn.fin.coef[c], invstd = bn.fin.coef[C + c];
    fin_backward(bn.fin, C, c, t1, invstd * (t2 - mean * t1));
  } else {
    fin_forward(bn.fin, C, c, t1, t2);
  }
  bn.acc[c] = 0.0;
  bn.acc[C + c] = 0.0;
}

int launch_bn_acc_finalize(const BnFuse& bn, int C, int mode, cudaStream_t st) {
  if (launch_pdl(bn_acc_finalize_kernel, dim3((C + 127) / 128), dim3(128), 0, st, bn, C, mode) != cudaSuccess) {
    set_error("bn_acc_finalize: launch failed");
    return STP_E_CUDA;
  }
  return check_launch("bn_acc_finalize");
}

}  // namespace stp

using namespace stp;

extern "C" int32_t stp_bn_nblk(int64_t rows, int32_t c) { return reduce_geom(rows, c < 8 ? 8 : c).nblk; }

static size_t reduce_smem(const RowGeom& g, int c) {
  const bool shfl = g.nv <= 32 && (g.nv & (g.nv - 1)) == 0 && (g.threads & 31) == 0;
  const int groups = shfl ? g.threads / 32 : g.rpi;
  size_t a = (size_t)2 * groups * c * sizeof(float);
  size_t b = (size_t)2 * g.threads * sizeof(double);
  return a > b ? a : b;
}

static int launch_stats(const stp_tensor* x, float* partial, const FinArgs& fin, cudaStream_t st) {
  int64_t rows = pixels(x);
  if (x->dtype == STP_F32) {   // parity mode
    STP_REQUIRE(f32::f32_ok(x) && fin.mode != 0, "bn_stats (fp32 parity mode): fused finalize only");
    return f32::launch_reduce(0, x, nullptr, nullptr, 0, 1, fin, st);
  }
  if (x->dtype == STP_U8) {
    STP_REQUIRE(fin.mode == 0, "bn_stats u8: fused finalize unsupported");
    STP_REQUIRE(x->c <= 4 && x->ld == x->c, "bn_stats u8: c<=4 dense only");
    int nblk = reduce_geom(rows, 8).nblk;  // same rule as stp_bn_nblk(rows, c<8)
    int64_t rpb = (rows + nblk - 1) / nblk;
    stats_u8_kernel<<<nblk, 256, 0, st>>>((const uint8_t*)x->ptr, rows, x->c, rpb, partial);
    return check_launch("stats_u8");
  }
  STP_REQUIRE(vec_ok(x), "bn_stats: tensor must be bf16, c%%8==0, ld%%8==0, 16B aligned");
  STP_REQUIRE(x->c <= 2048 && rows < 0x7fffffff, "bn_stats: c or rows too large");
  RowGeom g = reduce_geom(rows, x->c, fin.mode != 0 && fin.acc != nullptr);
  launch_pdl(reduce_rows_kernel<0, 1>, dim3(g.nblk), dim3(g.threads), reduce_smem(g, x->c), st,
             (const __nv_bfloat16*)x->ptr, x->ld, (const __nv_bfloat16*)nullptr, 0, (const float*)nullptr, 0, 1, x->h, x->w,
             (int)rows, x->c, g.cv, g.nv, g.rpi, g.rows_per_blk, partial, fin);
  return check_launch("bn_stats");
}

static int launch_bwd_reduce(const stp_tensor* dy, const stp_tensor* x, const float* coef, int relu, int pool,
                             float* partial, const FinArgs& fin, cudaStream_t st) {
  if (x && x->dtype == STP_F32) {   // parity mode
    STP_REQUIRE(f32::f32_ok(x) && f32::f32_ok(dy) && coef && fin.mode == 2 && (pool == 1 || pool == 2), "bn_bwd_reduce (fp32): bad args");
    STP_REQUIRE(dy->c == x->c && dy->h == x->h * pool && dy->w == x->w * pool && dy->n == x->n, "bn_bwd_reduce (fp32): shape mismatch");
    return f32::launch_reduce(1, x, dy, coef, relu, pool, fin, st);
  }
  STP_REQUIRE(vec_ok(dy) && vec_ok(x) && coef && partial, "bn_bwd_reduce: bad tensors");
  STP_REQUIRE(pool == 1 || pool == 2, "bn_bwd_reduce: pool must be 1 or 2");
  STP_REQUIRE(dy->c == x->c && dy->h == x->h * pool && dy->w == x->w * pool && dy->n == x->n,
              "bn_bwd_reduce: shape mismatch");
  int64_t rows = pixels(x);
  STP_REQUIRE(x->c <= 2048 && rows < 0x7fffffff, "bn_bwd_reduce: c or rows too large");
  RowGeom g = reduce_geom(rows, x->c, fin.mode != 0 && fin.acc != nullptr);
  if (pool == 2)
    launch_pdl(reduce_rows_kernel<1, 2>, dim3(g.nblk), dim3(g.threads), reduce_smem(g, x->c), st,
               (const __nv_bfloat16*)x->ptr, x->ld, (const __nv_bfloat16*)dy->ptr, dy->ld, coef, relu, pool, x->h, x->w,
               (int)rows, x->c, g.cv, g.nv, g.rpi, g.rows_per_blk, partial, fin);
  else
    launch_pdl(reduce_rows_kernel<1, 1>, dim3(g.nblk), dim3(g.threads), reduce_smem(g, x->c), st,
               (const __nv_bfloat16*)x->ptr, x->ld, (const __nv_bfloat16*)dy->ptr, dy->ld, coef, relu, pool, x->h, x->w,
               (int)rows, x->c, g.cv, g.nv, g.rpi, g.rows_per_blk, partial, fin);
  return check_launch("bn_bwd_reduce");
}

extern "C" int stp_bn_stats(const stp_tensor* x, float* partial, stp_stream stream) {
  STP_REQUIRE(x && partial, "bn_stats: null");
  FinArgs fin = {};
  return launch_stats(x, partial, fin, (cudaStream_t)stream);
}

extern "C" int stp_bn_stats_fused(const stp_tensor* x, float* partial, uint32_t* sync, double* acc, const float* gamma,
                                  const float* beta, float eps, float momentum, float* moving_mean, float* moving_var,
                                  float* coef, stp_stream stream) {
  STP_REQUIRE(x && partial && sync && coef, "bn_stats_fused: null");
  const int64_t count = pixels(x);
  STP_REQUIRE(count > 0, "bn_stats_fused: empty tensor");
  FinArgs fin = {};
  fin.mode = 1; fin.sync = sync; fin.acc = acc; fin.inv_count = 1.0 / (double)count;
  fin.bessel = keras_bessel(count, eps);
  fin.gamma = gamma; fin.beta = beta; fin.eps = eps; fin.momentum = momentum;
  fin.mov_mean = moving_mean; fin.mov_var = moving_var; fin.coef = coef;
  return launch_stats(x, partial, fin, (cudaStream_t)stream);
}

extern "C" int stp_bias_grad(const stp_tensor* dz, float* partial, uint32_t* sync, double* acc, float* dbias,
                             stp_stream stream) {
  STP_REQUIRE(dz && partial && sync && dbias, "bias_grad: null");
  FinArgs fin = {};
  fin.mode = 3; fin.sync = sync; fin.acc = acc; fin.dbeta = dbias;
  return launch_stats(dz, partial, fin, (cudaStream_t)stream);
}

extern "C" int stp_bn_finalize(const float* partial, int32_t nblk, int32_t c, int64_t count, const float* gamma,
                               const float* beta, float eps, float momentum, float* moving_mean, float* moving_var,
                               float* coef, stp_stream stream) {
  STP_REQUIRE(partial && coef && nblk > 0 && c > 0 && count > 0, "bn_finalize: bad args");
  double bessel = keras_bessel(count, eps);
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
  if (x && x->dtype == STP_F32) {   // parity mode
    STP_REQUIRE(f32::f32_ok(x) && f32::f32_ok(y) && coef && (up == 1 || up == 2), "bn_apply (fp32): bad tensors");
    STP_REQUIRE(y->c == x->c && y->n == x->n && y->h == x->h * up && y->w == x->w * up, "bn_apply (fp32): shape mismatch");
    return f32::bn_apply(x, coef, relu, up, y, (cudaStream_t)stream);
  }
  STP_REQUIRE(vec_ok(x) && vec_ok(y) && coef, "bn_apply: bad tensors");
  STP_REQUIRE(up == 1 || up == 2, "bn_apply: up must be 1 or 2");
  STP_REQUIRE(y->c == x->c && y->n == x->n && y->h == x->h * up && y->w == x->w * up, "bn_apply: shape mismatch");
  int64_t rows = pixels(x);
  STP_REQUIRE(rows < 0x7fffffff, "bn_apply: too many rows");
  RowGeom g = geom(rows, x->c, 4, kNumSMs * 8);
  launch_pdl(bn_apply_kernel, dim3(g.nblk), dim3(g.threads), 0, (cudaStream_t)stream, (const __nv_bfloat16*)x->ptr,
             x->ld, coef, relu, up, (__nv_bfloat16*)y->ptr, y->ld, x->h, x->w, (int)rows, x->c, g.cv, g.nv, g.rpi,
             g.rows_per_blk);
  return check_launch("bn_apply");
}

extern "C" int stp_bn_bwd_reduce(const stp_tensor* dy, const stp_tensor* x, const float* coef, int32_t relu,
                                 int32_t pool, float* partial, stp_stream stream) {
  FinArgs fin = {};
  return launch_bwd_reduce(dy, x, coef, relu, pool, partial, fin, (cudaStream_t)stream);
}

extern "C" int stp_bn_bwd_reduce_fused(const stp_tensor* dy, const stp_tensor* x, const float* coef, int32_t relu,
                                       int32_t pool, float* partial, uint32_t* sync, double* acc, float* dgamma,
                                       float* dbeta, float* bcoef, stp_stream stream) {
  STP_REQUIRE(x && sync && bcoef, "bn_bwd_reduce_fused: null");
  const int64_t count = pixels(x);
  STP_REQUIRE(count > 0, "bn_bwd_reduce_fused: empty tensor");
  FinArgs fin = {};
  fin.mode = 2; fin.sync = sync; fin.acc = acc; fin.inv_count = 1.0 / (double)count;
  fin.coef = const_cast<float*>(coef); fin.dgamma = dgamma; fin.dbeta = dbeta; fin.bcoef = bcoef;
  return launch_bwd_reduce(dy, x, coef, relu, pool, partial, fin, (cudaStream_t)stream);
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
  if (x && x->dtype == STP_F32) {   // parity mode
    STP_REQUIRE(f32::f32_ok(x) && f32::f32_ok(dy) && f32::f32_ok(dx) && (pool == 1 || pool == 2), "bwd_apply (fp32): bad tensors");
    STP_REQUIRE(dy->c == x->c && dx->c == x->c && dy->h == x->h * pool && dy->w == x->w * pool && dx->h == x->h && dx->w == x->w &&
                    dx->n == x->n && dy->n == x->n, "bwd_apply (fp32): shape mismatch");
    if (residual) STP_REQUIRE(f32::f32_ok(residual) && residual->c == x->c && pixels(residual) == pixels(x), "bwd_apply (fp32): bad residual");
    return f32::bwd_apply(mode, dy, x, coef, bcoef, relu, pool, residual, dx, (cudaStream_t)stream);
  }
  STP_REQUIRE(vec_ok(dy) && vec_ok(x) && vec_ok(dx), "bwd_apply: bad tensors");
  STP_REQUIRE(pool == 1 || pool == 2, "bwd_apply: pool must be 1 or 2");
  STP_REQUIRE(dy->c == x->c && dx->c == x->c && dy->h == x->h * pool && dy->w == x->w * pool && dx->h == x->h &&
                  dx->w == x->w && dx->n == x->n && dy->n == x->n,
              "bwd_apply: shape mismatch");
  if (residual) STP_REQUIRE(vec_ok(residual) && residual->c == x->c && pixels(residual) == pixels(x), "bwd_apply: bad residual");
  int64_t rows = pixels(x);
  STP_REQUIRE(rows < 0x7fffffff, "bwd_apply: too many rows");
  RowGeom g = geom(rows, x->c, pool == 2 ? 2 : 4, kNumSMs * 8);
  const __nv_bfloat16* rp = residual ? (const __nv_bfloat16*)residual->ptr : nullptr;
  int ldr = residual ? residual->ld : 0;
#define STP_BWD_APPLY(MODE, POOL)                                                                                      \
  launch_pdl(bwd_apply_kernel<MODE, POOL>, dim3(g.nblk), dim3(g.threads), 0, (cudaStream_t)stream,                     \
             (const __nv_bfloat16*)dy->ptr, dy->ld, (const __nv_bfloat16*)x->ptr, x->ld, coef, bcoef, relu, pool, rp,  \
             ldr, (__nv_bfloat16*)dx->ptr, dx->ld, x->h, x->w, (int)rows, x->c, g.cv, g.nv, g.rpi, g.rows_per_blk)
  if (mode == 0) {
    if (pool == 2) STP_BWD_APPLY(0, 2); else STP_BWD_APPLY(0, 1);
  } else {
    if (pool == 2) STP_BWD_APPLY(1, 2); else STP_BWD_APPLY(1, 1);
  }
#undef STP_BWD_APPLY
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
  if (y && y->dtype == STP_F32) {   // parity mode: plain [n, h, w, C] output (no space-to-depth form)
    STP_REQUIRE(img && coef && f32::f32_ok(y) && y->h == h && y->w == w && y->n == n && c_img < y->c, "stem_prep (fp32): bad args");
    return f32::stem_prep(img, (int64_t)n * h * w, c_img, coef, y, (cudaStream_t)stream);
  }
  STP_REQUIRE(img && coef && vec_ok(y), "stem_prep: bad args");
  int64_t rows = (int64_t)n * h * w;
  if (y->h == h && y->w == w) {
    STP_REQUIRE(c_img < y->c && y->n == n, "stem_prep: shape mismatch");
    stem_prep_kernel<<<ew_grid(rows), 256, 0, (cudaStream_t)stream>>>(img, rows, c_img, coef, (__nv_bfloat16*)y->ptr,
                                                                      y->ld, y->c, 0, h, w);
  } else {
    // space-to-depth: y [n, h/2, w/2, 4*8]
    STP_REQUIRE(h % 2 == 0 && w % 2 == 0 && y->h == h / 2 && y->w == w / 2 && y->c == 32 && y->n == n && c_img < 8,
                "stem_prep: space-to-depth output must be [n, h/2, w/2, 32]");
    stem_prep_kernel<<<ew_grid(rows), 256, 0, (cudaStream_t)stream>>>(img, rows, c_img, coef, (__nv_bfloat16*)y->ptr,
                                                                      y->ld, 8, 1, h, w);
  }
  return check_launch("stem_prep");
}

extern "C" int stp_copy_up(const stp_tensor* x, int32_t up, const stp_tensor* y, stp_stream stream) {
  if (x && x->dtype == STP_F32) {   // parity mode
    STP_REQUIRE(f32::f32_ok(x) && f32::f32_ok(y) && up >= 1 && y->c == x->c && y->n == x->n && y->h == x->h * up && y->w == x->w * up,
                "copy_up (fp32): bad tensors");
    return f32::copy_up(x, up, y, (cudaStream_t)stream);
  }
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
  if (a && a->dtype == STP_F32) {   // parity mode
    STP_REQUIRE(f32::f32_ok(a) && f32::f32_ok(b) && f32::f32_ok(y) && a->c == b->c && a->c == y->c && pixels(a) == pixels(b) &&
                    pixels(a) == pixels(y), "add (fp32): bad tensors");
    return f32::add(a, b, y, (cudaStream_t)stream);
  }
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

// ---- space-to-depth stem: 7x7/2 over [H,W,8]  ==  4x4/1 (pad 2 before) over [H/2,W/2,32] ------------------------------
// w2[co][r'][s'][(dy*2+dx)*8 + c] = w[co][2r'+dy-1][2s'+dx-1][c]  (zero where the 7x7 index falls outside)
namespace stp {
__global__ void stem_weight_s2d_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ w2, int cout) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= cout * 16 * 32) return;
  int c = i & 7, q = (i >> 3) & 3, sp = (i >> 5) & 3, rp = (i >> 7) & 3, co = i >> 9;
  int r = 2 * rp + (q >> 1) - 1, s = 2 * sp + (q & 1) - 1;
  float v = 0.f;
  if (r >= 0 && r < 7 && s >= 0 && s < 7) v = w[((co * 7 + r) * 7 + s) * 8 + c];
  w2[i] = __float2bfloat16(v);
}
// inverse gather of the gradient: dw[co][r][s][c] = dw2[co][(r+1)/2][(s+1)/2][(((r+1)&1)*2 + ((s+1)&1))*8 + c]
__global__ void stem_wgrad_s2d_gather_kernel(const float* __restrict__ dw2, float* __restrict__ dw, int cout) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= cout * 49 * 8) return;
  int c = i & 7, t = i >> 3;
  int s = t % 7;
  t /= 7;
  int r = t % 7, co = t / 7;
  int rp = (r + 1) >> 1, sp = (s + 1) >> 1, q = ((r + 1) & 1) * 2 + ((s + 1) & 1);
  dw[i] = dw2[(((co * 4 + rp) * 4 + sp) * 4 + q) * 8 + c];
}
}  // namespace stp

extern "C" int stp_stem_weight_s2d(const float* w_master, void* w2, int32_t cout, stp_stream stream) {
  STP_REQUIRE(w_master && w2 && cout > 0, "stem_weight_s2d: bad args");
  stp::stem_weight_s2d_kernel<<<(cout * 512 + 255) / 256, 256, 0, (cudaStream_t)stream>>>(w_master, (__nv_bfloat16*)w2, cout);
  return check_launch("stem_weight_s2d");
}

extern "C" int stp_stem_wgrad_s2d_gather(const float* dw2, float* dw, int32_t cout, stp_stream stream) {
  STP_REQUIRE(dw2 && dw && cout > 0, "stem_wgrad_s2d_gather: bad args");
  stp::stem_wgrad_s2d_gather_kernel<<<(cout * 392 + 255) / 256, 256, 0, (cudaStream_t)stream>>>(dw2, dw, cout);
  return check_launch("stem_wgrad_s2d_gather");
}

extern "C" int stp_stem_wgrad_post(float* dw8, const float* w_master, int32_t cout, int32_t r, int32_t s,
                                   int32_t cin_pad, int32_t c_img, float* dbeta, stp_stream stream) {
  STP_REQUIRE(dw8 && w_master && c_img >= 1 && c_img <= 4 && c_img < cin_pad, "stem_wgrad_post: bad args");
  if (r < 0) {   // parity mode (engine passes -r): the conv ran on the fp32 weights themselves, not on a bf16 copy
    return stp::f32::stem_wgrad_post(dw8, w_master, cout * (-r) * s, cin_pad, c_img, dbeta, (cudaStream_t)stream);
  }
  stp::stem_wgrad_post_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(dw8, w_master, cout * r * s, cin_pad, c_img, dbeta);
  return check_launch("stem_wgrad_post");
}
