// Generic implicit-GEMM convolution on the legacy tensor path (mma.sync m16n8k16 bf16, fp32 accumulate).
// Covers EVERY shape (any R,S, stride, zero-insertion `up`, strided channel views, Cout not a tile
// multiple).  It is the correctness scaffold and the kernel for the layers that are HBM bound or
// irregular (stem 7x7/2, stride-2 convs and their dgrads, 1x1 shortcuts); the FLOP-heavy 3x3/s1 layers
// are routed to the tcgen05/TMA kernels in conv_tc.cu when stp_tc_enabled().
//
// fwd  : Y[m, co]  = sum_k A[m, k] * Wt[co, k]      m=(n,ho,wo)  k=(r,s,ci)   A gathered from X
// dgrad: same kernel, X:=dY, W:=tap-flipped [Cin][R][S][Cout], stride:=1, up:=stride, pad:=R-1-pad
// wgrad: dW[co, k] = sum_m dY[m, co] * A[m, k]      (split over m, deterministic reduction)
#include "common.cuh"
#include "conv.h"

namespace stp {

__device__ __forceinline__ void ldmatrix_x4(uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3, const void* p) {
  uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];\n"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(a));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3,
                                                  const void* p) {
  uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];\n"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(a));
}
__device__ __forceinline__ void mma_bf16(float* c, const uint32_t* a, const uint32_t* b) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

constexpr int BK = 32;
constexpr int PITCH = BK + 8;  // 80 B rows: conflict-free ldmatrix

// gather one 16-byte chunk (8 channels) of the implicit A matrix
__device__ __forceinline__ uint4 gather_a(const ConvP& p, bool row_ok, int64_t img_base, int hi0, int wi0, int k) {
  uint4 z = make_uint4(0, 0, 0, 0);
  if (!row_ok || k >= p.K) return z;
  int tap = k / p.Cin;
  int ci = k - tap * p.Cin;
  int r = tap / p.S;
  int s = tap - r * p.S;
  int hu = hi0 + r, wu = wi0 + s;
  if (p.up > 1) {
    if (hu < 0 || wu < 0 || (hu % p.up) != 0 || (wu % p.up) != 0) return z;
    hu /= p.up;
    wu /= p.up;
  }
  if (hu < 0 || hu >= p.H || wu < 0 || wu >= p.W) return z;
  return *reinterpret_cast<const uint4*>(p.x + (img_base + (int64_t)hu * p.W + wu) * p.ldx + ci);
}

template <int BM, int BN, int WM, int WN>
__global__ void __launch_bounds__(256) conv_igemm_kernel(const ConvP p) {
  constexpr int WTM = BM / WM, WTN = BN / WN;  // warp tile
  constexpr int MI = WTM / 16, NI = WTN / 8;
  static_assert(WM * WN == 8 && MI >= 1 && NI >= 2 && NI % 2 == 0, "bad tile");
  constexpr int A_CHUNKS = BM * (BK / 8) / 256;                  // per thread
  constexpr int B_CHUNKS = (BN * (BK / 8) + 255) / 256;          // per thread (some idle)
  __shared__ __align__(16) __nv_bfloat16 As[2][BM][PITCH];
  __shared__ __align__(16) __nv_bfloat16 Bs[2][BN][PITCH];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = warp % WM, wn = warp / WM;
  const int64_t m0 = (int64_t)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;

  // per-thread A rows
  bool row_ok[A_CHUNKS];
  int64_t img_base[A_CHUNKS];
  int hi0[A_CHUNKS], wi0[A_CHUNKS];
  const int a_chunk = tid & 3;
#pragma unroll
  for (int j = 0; j < A_CHUNKS; ++j) {
    int row = (tid >> 2) + j * 64;
    int64_t m = m0 + row;
    row_ok[j] = m < p.M;
    int64_t mm = row_ok[j] ? m : 0;
    int64_t n = mm / ((int64_t)p.Ho * p.Wo);
    int rem = (int)(mm - n * (int64_t)p.Ho * p.Wo);
    int ho = rem / p.Wo, wo = rem - ho * p.Wo;
    img_base[j] = n * (int64_t)p.H * p.W;
    hi0[j] = ho * p.stride - p.pad_h;
    wi0[j] = wo * p.stride - p.pad_w;
  }
  const int b_chunk = tid & 3;
  uint4 ra[A_CHUNKS], rb[B_CHUNKS];

  auto load_global = [&](int kt) {
    const int kbase = kt * BK;
#pragma unroll
    for (int j = 0; j < A_CHUNKS; ++j) ra[j] = gather_a(p, row_ok[j], img_base[j], hi0[j], wi0[j], kbase + a_chunk * 8);
#pragma unroll
    for (int j = 0; j < B_CHUNKS; ++j) {
      int row = (tid >> 2) + j * 64;
      int co = n0 + row;
      int k = kbase + b_chunk * 8;
      rb[j] = make_uint4(0, 0, 0, 0);
      if (row < BN && co < p.Cout && k < p.K) rb[j] = *reinterpret_cast<const uint4*>(p.w + (int64_t)co * p.K + k);
    }
  };
  auto store_smem = [&](int buf) {
#pragma unroll
    for (int j = 0; j < A_CHUNKS; ++j)
      *reinterpret_cast<uint4*>(&As[buf][(tid >> 2) + j * 64][a_chunk * 8]) = ra[j];
#pragma unroll
    for (int j = 0; j < B_CHUNKS; ++j) {
      int row = (tid >> 2) + j * 64;
      if (row < BN) *reinterpret_cast<uint4*>(&Bs[buf][row][b_chunk * 8]) = rb[j];
    }
  };

  float acc[MI][NI][4];
#pragma unroll
  for (int i = 0; i < MI; ++i)
#pragma unroll
    for (int j = 0; j < NI; ++j)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[i][j][e] = 0.f;

  const int KT = (p.K + BK - 1) / BK;
  load_global(0);
  store_smem(0);
  __syncthreads();
  for (int kt = 0; kt < KT; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < KT) load_global(kt + 1);
#pragma unroll
    for (int kk = 0; kk < BK; kk += 16) {
      uint32_t af[MI][4], bf[NI][2];
#pragma unroll
      for (int i = 0; i < MI; ++i)
        ldmatrix_x4(af[i][0], af[i][1], af[i][2], af[i][3],
                    &As[buf][wm * WTM + i * 16 + (lane & 15)][kk + (lane >> 4) * 8]);
#pragma unroll
      for (int j = 0; j < NI; j += 2) {
        const int mj = lane >> 3;
        ldmatrix_x4(bf[j][0], bf[j][1], bf[j + 1][0], bf[j + 1][1],
                    &Bs[buf][wn * WTN + j * 8 + (mj >> 1) * 8 + (lane & 7)][kk + (mj & 1) * 8]);
      }
#pragma unroll
      for (int i = 0; i < MI; ++i)
#pragma unroll
        for (int j = 0; j < NI; ++j) mma_bf16(acc[i][j], af[i], bf[j]);
    }
    if (kt + 1 < KT) store_smem(buf ^ 1);
    __syncthreads();
  }

  // epilogue
#pragma unroll
  for (int i = 0; i < MI; ++i) {
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int64_t m = m0 + wm * WTM + i * 16 + (lane >> 2) + half * 8;
      if (m >= p.M) continue;
#pragma unroll
      for (int j = 0; j < NI; ++j) {
        const int n = n0 + wn * WTN + j * 8 + (lane & 3) * 2;
        if (n >= p.Cout) continue;
        float v0 = acc[i][j][half * 2 + 0], v1 = acc[i][j][half * 2 + 1];
        const bool has1 = (n + 1) < p.Cout;
        if (p.bias) {
          v0 += p.bias[n];
          if (has1) v1 += p.bias[n + 1];
        }
        if (p.res) {
          v0 += __bfloat162float(p.res[m * p.ldr + n]);
          if (has1) v1 += __bfloat162float(p.res[m * p.ldr + n + 1]);
        }
        if (p.relu) {
          v0 = fmaxf(v0, 0.f);
          v1 = fmaxf(v1, 0.f);
        }
        if (p.y_f32) {
          float* y = reinterpret_cast<float*>(p.y) + m * p.ldy + n;
          y[0] = v0;
          if (has1) y[1] = v1;
        } else {
          __nv_bfloat16* y = reinterpret_cast<__nv_bfloat16*>(p.y) + m * p.ldy + n;
          if (has1 && ((p.ldy & 1) == 0)) {
            *reinterpret_cast<__nv_bfloat162*>(y) = __floats2bfloat162_rn(v0, v1);
          } else {
            y[0] = __float2bfloat16(v0);
            if (has1) y[1] = __float2bfloat16(v1);
          }
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// wgrad: out[split][co][k] (f32) = sum over this split's pixels of dY[m][co] * A[m][k]
// tile 64 (co) x 64 (k), 32 pixels per step, 8 warps as 2 (co) x 4 (k)
// ------------------------------------------------------------------------------------------------
constexpr int WP = 64 + 8;  // 144 B rows

__global__ void __launch_bounds__(256) conv_wgrad_kernel(const WgradP p) {
  __shared__ __align__(16) __nv_bfloat16 Ds[2][32][WP];
  __shared__ __align__(16) __nv_bfloat16 Xs[2][32][WP];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wco = warp & 1, wk = warp >> 1;
  const int co0 = blockIdx.y * 64;
  const int k0 = blockIdx.x * 64;
  const int split = blockIdx.z;
  const int64_t chunk_begin = (int64_t)split * p.chunks_per_split;
  const int64_t total_chunks = (p.M + 31) / 32;
  int64_t chunk_end = chunk_begin + p.chunks_per_split;
  if (chunk_end > total_chunks) chunk_end = total_chunks;

  const int pix = tid >> 3, chunk = tid & 7;
  // fixed k decomposition for this thread's X chunk
  const int k = k0 + chunk * 8;
  const bool k_ok = k < p.K;
  int tap = k_ok ? k / p.Cin : 0;
  const int ci = k_ok ? k - tap * p.Cin : 0;
  const int fr = tap / p.S, fs = tap - fr * p.S;
  const int co = co0 + chunk * 8;
  const bool co_ok = co < p.Cout;

  uint4 rd, rx;
  auto load_global = [&](int64_t ch) {
    int64_t m = ch * 32 + pix;
    rd = make_uint4(0, 0, 0, 0);
    rx = make_uint4(0, 0, 0, 0);
    if (m < p.M) {
      if (co_ok) rd = *reinterpret_cast<const uint4*>(p.dy + m * p.lddy + co);
      if (k_ok) {
        int64_t n = m / ((int64_t)p.Ho * p.Wo);
        int rem = (int)(m - n * (int64_t)p.Ho * p.Wo);
        int ho = rem / p.Wo, wo = rem - ho * p.Wo;
        int hu = ho * p.stride - p.pad_h + fr, wu = wo * p.stride - p.pad_w + fs;
        bool ok = true;
        if (p.up > 1) {
          ok = hu >= 0 && wu >= 0 && (hu % p.up) == 0 && (wu % p.up) == 0;
          hu /= p.up;
          wu /= p.up;
        }
        if (ok && hu >= 0 && hu < p.H && wu >= 0 && wu < p.W)
          rx = *reinterpret_cast<const uint4*>(p.x + ((n * p.H + hu) * (int64_t)p.W + wu) * p.ldx + ci);
      }
    }
  };
  auto store_smem = [&](int buf) {
    *reinterpret_cast<uint4*>(&Ds[buf][pix][chunk * 8]) = rd;
    *reinterpret_cast<uint4*>(&Xs[buf][pix][chunk * 8]) = rx;
  };

  float acc[2][2][4];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[i][j][e] = 0.f;

  if (chunk_begin < chunk_end) {
    load_global(chunk_begin);
    store_smem(0);
  }
  __syncthreads();
  for (int64_t ch = chunk_begin; ch < chunk_end; ++ch) {
    const int buf = (int)((ch - chunk_begin) & 1);
    if (ch + 1 < chunk_end) load_global(ch + 1);
#pragma unroll
    for (int kk = 0; kk < 32; kk += 16) {
      uint32_t af[2][4], bf[2][2];
      const int mj = lane >> 3;
#pragma unroll
      for (int i = 0; i < 2; ++i)
        ldmatrix_x4_trans(af[i][0], af[i][1], af[i][2], af[i][3],
                          &Ds[buf][kk + (mj >> 1) * 8 + (lane & 7)][wco * 32 + i * 16 + (mj & 1) * 8]);
      ldmatrix_x4_trans(bf[0][0], bf[0][1], bf[1][0], bf[1][1],
                        &Xs[buf][kk + (mj & 1) * 8 + (lane & 7)][wk * 16 + (mj >> 1) * 8]);
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j) mma_bf16(acc[i][j], af[i], bf[j]);
    }
    if (ch + 1 < chunk_end) store_smem(buf ^ 1);
    __syncthreads();
  }
  float* out = p.out + (int64_t)split * p.Cout * p.K;
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int c = co0 + wco * 32 + i * 16 + (lane >> 2) + half * 8;
      if (c >= p.Cout) continue;
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int kc = k0 + wk * 16 + j * 8 + (lane & 3) * 2;
        if (kc >= p.K) continue;
        *reinterpret_cast<float2*>(out + (int64_t)c * p.K + kc) = make_float2(acc[i][j][half * 2], acc[i][j][half * 2 + 1]);
      }
    }
}

__global__ void __launch_bounds__(256) split_reduce_kernel(const float* __restrict__ ws, int splits, int64_t n,
                                                           float* __restrict__ out) {
  pdl_launch_dependents();
  pdl_wait();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float a = 0.f;
    int s = 0;
    for (; s + 8 <= splits; s += 8) {  // 8 independent loads in flight, summed in split order (deterministic)
      float t[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) t[j] = ws[(int64_t)(s + j) * n + i];
#pragma unroll
      for (int j = 0; j < 8; ++j) a += t[j];
    }
    for (; s < splits; ++s) a += ws[(int64_t)s * n + i];
    out[i] = a;
  }
}

__global__ void __launch_bounds__(256) weight_prep_kernel(const float* __restrict__ wm, __nv_bfloat16* __restrict__ wf,
                                                          __nv_bfloat16* __restrict__ wd, int Cout, int R, int S,
                                                          int Cin) {
  const int64_t total = (int64_t)Cout * R * S * Cin;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    __nv_bfloat16 v = __float2bfloat16(wm[i]);
    if (wf) wf[i] = v;
    if (wd) {
      int ci = (int)(i % Cin);
      int64_t t = i / Cin;
      int s = (int)(t % S);
      t /= S;
      int r = (int)(t % R);
      int co = (int)(t / R);
      wd[(((int64_t)ci * R + (R - 1 - r)) * S + (S - 1 - s)) * Cout + co] = v;
    }
  }
}

// All conv layers of a network in ONE launch.  items: int64 [n][8] = {offset (elements, into the three flat buffers),
// Cout, R, S, Cin, has_dgrad, first_tile, unused}; one block = one 32(co) x 32(ci) tile, looping over the layer's R*S
// filter taps, each transposed through shared memory so that the KRSC read, the bf16 KRSC write and the tap-flipped
// [Cin][R][S][Cout] write are all row-contiguous.
__global__ void __launch_bounds__(256) weight_prep_batched_kernel(const float* __restrict__ wm,
                                                                  __nv_bfloat16* __restrict__ wf,
                                                                  __nv_bfloat16* __restrict__ wd,
                                                                  const int64_t* __restrict__ items, int n_items) {
  __shared__ float tile[32][33];
  int lo = 0, hi = n_items - 1;
  while (lo < hi) {
    int mid = (lo + hi + 1) >> 1;
    if (items[mid * 8 + 6] <= (int64_t)blockIdx.x) lo = mid; else hi = mid - 1;
  }
  const int64_t* it = items + lo * 8;
  const int64_t off = it[0];
  const int Cout = (int)it[1], R = (int)it[2], S = (int)it[3], Cin = (int)it[4];
  const bool dg = it[5] != 0;
  int t = (int)((int64_t)blockIdx.x - it[6]);
  const int tci = (Cin + 31) / 32;
  const int ci0 = (t % tci) * 32;
  const int co0 = (t / tci) * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int64_t K = (int64_t)R * S * Cin;
  for (int tap = 0; tap < R * S; ++tap) {
    const int r = tap / S, sx = tap - r * S;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int co = co0 + ty + j * 8, ci = ci0 + tx;
      if (co < Cout && ci < Cin) {
        const int64_t i = off + (int64_t)co * K + (int64_t)tap * Cin + ci;
        const float v = wm[i];
        wf[i] = __float2bfloat16(v);
        tile[ty + j * 8][tx] = v;
      }
    }
    if (!dg) continue;
    __syncthreads();
    const int tapf = (R - 1 - r) * S + (S - 1 - sx);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int ci = ci0 + ty + j * 8, co = co0 + tx;
      if (co < Cout && ci < Cin)
        wd[off + ((int64_t)ci * R * S + tapf) * Cout + co] = __float2bfloat16(tile[tx][ty + j * 8]);
    }
    __syncthreads();
  }
}

// ---- host side ---------------------------------------------------------------------------------
int launch_generic_conv(const ConvP& p, cudaStream_t st) {
  dim3 block(256);
  if (p.Cout <= 16) {
    dim3 grid((unsigned)((p.M + 127) / 128), (p.Cout + 15) / 16);
    conv_igemm_kernel<128, 16, 8, 1><<<grid, block, 0, st>>>(p);
  } else if (p.Cout <= 32) {
    dim3 grid((unsigned)((p.M + 127) / 128), (p.Cout + 31) / 32);
    conv_igemm_kernel<128, 32, 8, 1><<<grid, block, 0, st>>>(p);
  } else if (p.Cout <= 64 || p.Cout % 128 != 0) {
    dim3 grid((unsigned)((p.M + 127) / 128), (p.Cout + 63) / 64);
    conv_igemm_kernel<128, 64, 4, 2><<<grid, block, 0, st>>>(p);
  } else {
    dim3 grid((unsigned)((p.M + 127) / 128), (p.Cout + 127) / 128);
    conv_igemm_kernel<128, 128, 2, 4><<<grid, block, 0, st>>>(p);
  }
  return check_launch("conv_igemm");
}

static int wgrad_splits(int64_t M, int Cout, int K) {
  int64_t tiles = (int64_t)((Cout + 63) / 64) * ((K + 63) / 64);
  int64_t target = (int64_t)kNumSMs * 4;
  int64_t s = (target + tiles - 1) / tiles;
  int64_t max_s = (M + 255) / 256;  // at least 8 chunks per split
  if (s > max_s) s = max_s;
  if (s < 1) s = 1;
  if (s > 512) s = 512;
  return (int)s;
}

size_t generic_wgrad_workspace(int64_t M, int Cout, int K) {
  int s = wgrad_splits(M, Cout, K);
  return s > 1 ? (size_t)s * Cout * K * sizeof(float) : 0;
}

int launch_generic_wgrad(WgradP p, float* dw, void* ws, size_t ws_bytes, cudaStream_t st) {
  int splits = wgrad_splits(p.M, p.Cout, p.K);
  size_t need = splits > 1 ? (size_t)splits * p.Cout * p.K * sizeof(float) : 0;
  if (need > ws_bytes || (need && !ws)) {
    set_error("conv_wgrad: workspace too small (%zu < %zu)", ws_bytes, need);
    return STP_E_WORKSPACE;
  }
  int64_t total_chunks = (p.M + 31) / 32;
  p.chunks_per_split = (total_chunks + splits - 1) / splits;
  p.out = splits > 1 ? (float*)ws : dw;
  dim3 grid((p.K + 63) / 64, (p.Cout + 63) / 64, splits);
  conv_wgrad_kernel<<<grid, 256, 0, st>>>(p);
  int rc = check_launch("conv_wgrad");
  if (rc || splits == 1) return rc;
  return launch_split_reduce((const float*)ws, splits, (int64_t)p.Cout * p.K, dw, st);
}

int launch_split_reduce(const float* ws, int splits, int64_t n, float* out, cudaStream_t st) {
  int64_t nb = (n + 255) / 256;
  launch_pdl(split_reduce_kernel, dim3((unsigned)(nb < 2048 ? nb : 2048)), dim3(256), 0, st, ws, splits, n, out);
  return check_launch("split_reduce");
}

}  // namespace stp

using namespace stp;

extern "C" int stp_weight_prep_batched(const float* flat_master, void* flat_fwd, void* flat_dgrad,
                                       const int64_t* d_items, int32_t n_items, int64_t total_tiles, stp_stream stream) {
  STP_REQUIRE(flat_master && flat_fwd && flat_dgrad && d_items && n_items > 0 && total_tiles > 0 &&
                  total_tiles < 0x7fffffff,
              "weight_prep_batched: bad args");
  weight_prep_batched_kernel<<<(unsigned)total_tiles, 256, 0, (cudaStream_t)stream>>>(
      flat_master, (__nv_bfloat16*)flat_fwd, (__nv_bfloat16*)flat_dgrad, d_items, n_items);
  return check_launch("weight_prep_batched");
}

extern "C" int stp_weight_prep(const float* w_master, void* w_fwd, void* w_dgrad, int32_t cout, int32_t r, int32_t s,
                               int32_t cin, stp_stream stream) {
  STP_REQUIRE(w_master && (w_fwd || w_dgrad) && cout > 0 && r > 0 && s > 0 && cin > 0, "weight_prep: bad args");
  int64_t total = (int64_t)cout * r * s * cin;
  int64_t nb = (total + 255) / 256;
  weight_prep_kernel<<<(int)(nb < 1184 ? nb : 1184), 256, 0, (cudaStream_t)stream>>>(
      w_master, (__nv_bfloat16*)w_fwd, (__nv_bfloat16*)w_dgrad, cout, r, s, cin);
  return check_launch("weight_prep");
}
