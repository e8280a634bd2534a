// 1x1 stride-1 convolution = plain GEMM  Y[M][N] = X[M][K] * W[N][K]^T (+ bias, + residual, ReLU)  for the channel counts the
// tcgen05 halo kernel does not tile (MobileNetV2 / Xception widths: 24, 96, 144, 160, 728 ...).  Where that kernel does tile the
// shape it is ~2x faster and takes the layer (csrc/api.cu dispatch_conv, profiles/r2_s10_g1_bench.txt): this kernel runs on the
// legacy tensor path, whose ceiling on sm_100a is ~290 TF/s (K-heavy shapes measure 219-257 TF/s here).
// Design: 16-byte cp.async into a 4-stage XOR-swizzled shared-memory ring, ldmatrix + mma.sync.m16n8k16 (bf16 in, fp32
// accumulate), three short-lived CTAs per SM (a persistent variant with one continuous cp.async stream measured 20 % slower),
// the residual tile prefetched by the first cp.async group, the epilogue staged through shared memory so that every global
// store is a row-contiguous 16-byte vector.  The dgrad of such a layer is the same GEMM over dY with the [Cin][Cout] weight copy.
// BatchNorm epilogues are compile-time variants (BNM): forward statistics of the stored values, or -- for a dgrad -- masking by
// the producing BatchNorm's activation (ReLU or ReLU6) + the (sum g, sum g*x) reduction, finalised by bn_acc_finalize_kernel (bn.cu).
// It replaced the implicit-GEMM generic kernel (conv_generic.cu), which spends its time on im2col index arithmetic that a 1x1
// filter does not need: 160 -> 960 @40x40 bs16 took 101 us there (profiles/r2_s7_deeplab_bench.txt).
#include "conv.h"

namespace stp {
namespace {

constexpr int kBM = 128, kBK = 32, kStages = 4, kThreads = 256;

struct G1Args {
  const __nv_bfloat16* A;
  int lda;
  const __nv_bfloat16* B;   // [N][K]
  __nv_bfloat16* Y;
  int ldy;
  const __nv_bfloat16* res;
  int ldr;
  const float* bias;
  int relu, M, N, K, tiles_n;
  int bn_on;  // 0 | 1 forward BatchNorm statistics of Y | 2 fused BatchNorm backward: res = the BatchNorm input x (mask + reduce)
  BnFuse bn;
  const float* bnb_coef;
  int bnb_relu;
};

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, int bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(dst), "l"(src), "r"(bytes));
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }
__device__ __forceinline__ void ldsm4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];\n" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void mma16816(float* c, const uint32_t* a, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// byte offset of (row, 16-byte chunk c in 0..3) inside a [rows][32] bf16 tile: chunk index XORed with (row >> 1) & 3
__device__ __forceinline__ uint32_t swz(int row, int c) { return (uint32_t)(row * 64 + ((c ^ ((row >> 1) & 3)) << 4)); }

// BNM: 0 plain | 1 forward BatchNorm statistics | 2 fused BatchNorm backward -- compile-time, so the plain GEMM carries none of the
// epilogue's registers / instructions (with a run-time switch the plain launches of the DeepLabV3 step were 3 % slower)
template <int BN, int BNM>
__global__ void __launch_bounds__(kThreads) gemm1x1_kernel(const G1Args a) {
  extern __shared__ __align__(128) uint8_t smem[];
  constexpr int kABytes = kBM * kBK * 2, kBBytes = BN * kBK * 2, kStage = kABytes + kBBytes;
  constexpr int WM = BN == 128 ? 2 : 4, WN = 8 / WM;           // warp grid
  constexpr int TM = kBM / WM / 16, TN = BN / WN / 8;           // mma tiles per warp: (4 x 4) for BN = 128, (2 x 4) for BN = 64
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int tn = blockIdx.x % a.tiles_n, tm = blockIdx.x / a.tiles_n;
  const int m0 = tm * kBM, n0 = tn * BN;
  const int wm = warp % WM, wn = warp / WM;
  const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(smem);
  const int KT = (a.K + kBK - 1) / kBK;
  // residual tile [128][BN + 8] bf16 behind the ring: fetched as coalesced 16-byte row chunks by the first cp.async group (it
  // lands under the whole K loop) and read back in fragment layout by the epilogue.  Per-fragment 4-byte global loads (8 rows
  // x 16 B per warp instruction, half of every sector wasted) made the accumulate-in-place dgrads of the ResNet-50 bottlenecks
  // the slowest launches of the FPN step (256 ch @128^2: 134 us for a 46 us HBM floor, profiles/r2_c3_launches.csv).
  constexpr int LDR = BN + 8;
  __nv_bfloat16* sres = reinterpret_cast<__nv_bfloat16*>(smem + kStages * kStage);
  if (a.res) {
    constexpr int CVR = BN / 8;
    const uint32_t sr = sbase + kStages * kStage;
    for (int q = tid; q < kBM * CVR; q += kThreads) {
      const int row = q / CVR, c = q - row * CVR;
      const int m = m0 + row, n = n0 + c * 8;
      const bool ok = m < a.M && n < a.N;
      cp_async16(sr + (uint32_t)(row * LDR + c * 8) * 2, ok ? (const void*)(a.res + (int64_t)m * a.ldr + n) : (const void*)a.res, ok ? 16 : 0);
    }
    cp_commit();
  }

  auto load_stage = [&](int kt, int stage) {
    const uint32_t sa = sbase + stage * kStage, sb = sa + kABytes;
    const int k0 = kt * kBK;
#pragma unroll
    for (int i = 0; i < kBM * 4 / kThreads; ++i) {
      const int q = tid + i * kThreads, row = q >> 2, c = q & 3;
      const int m = m0 + row, k = k0 + c * 8;
      const bool ok = m < a.M && k < a.K;
      cp_async16(sa + swz(row, c), ok ? (const void*)(a.A + (int64_t)m * a.lda + k) : (const void*)a.A, ok ? 16 : 0);
    }
#pragma unroll
    for (int i = 0; i < (BN * 4 + kThreads - 1) / kThreads; ++i) {
      const int q = tid + i * kThreads, row = q >> 2, c = q & 3;
      if (row < BN) {
        const int n = n0 + row, k = k0 + c * 8;
        const bool ok = n < a.N && k < a.K;
        cp_async16(sb + swz(row, c), ok ? (const void*)(a.B + (int64_t)n * a.K + k) : (const void*)a.B, ok ? 16 : 0);
      }
    }
  };

  float acc[TM][TN][4];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j)
#pragma unroll
      for (int q = 0; q < 4; ++q) acc[i][j][q] = 0.f;

#pragma unroll
  for (int s = 0; s < kStages - 1; ++s) {
    if (s < KT) load_stage(s, s);
    cp_commit();
  }
  for (int kt = 0; kt < KT; ++kt) {
    cp_wait<kStages - 2>();
    __syncthreads();
    if (kt + kStages - 1 < KT) load_stage(kt + kStages - 1, (kt + kStages - 1) % kStages);
    cp_commit();
    const uint32_t sa = sbase + (kt % kStages) * kStage, sb = sa + kABytes;
#pragma unroll
    for (int ks = 0; ks < kBK / 16; ++ks) {
      uint32_t af[TM][4], bf[TN][2];
#pragma unroll
      for (int i = 0; i < TM; ++i) {
        const int row = wm * (kBM / WM) + i * 16 + (lane & 15);
        ldsm4(sa + swz(row, ks * 2 + (lane >> 4)), af[i][0], af[i][1], af[i][2], af[i][3]);
      }
#pragma unroll
      for (int j = 0; j < TN; j += 2) {
        const int row = wn * (BN / WN) + j * 8 + (lane & 7) + ((lane >> 4) << 3);
        ldsm4(sb + swz(row, ks * 2 + ((lane >> 3) & 1)), bf[j][0], bf[j][1], bf[j + 1][0], bf[j + 1][1]);
      }
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) mma16816(acc[i][j], af[i], bf[j][0], bf[j][1]);
    }
  }
  cp_wait<0>();
  __syncthreads();

  // epilogue: bias / residual / ReLU on the fragments (one rounding), staged as bf16 [128][BN + 8], then 16-byte row stores.
  // bn_on == 1: per-channel (sum y, sum y^2) of exactly the stored bf16 values (BatchNorm forward statistics); bn_on == 2 (a dgrad
  // launch): the residual slot carries the BatchNorm INPUT x of the same pixels, the stored value is g = y*[act'(x*scale+shift)]
  // and the sums are (sum g, sum g*x) -- conv_tc2's epilogue contract.  The sums leave as fire-and-forget double reductions and
  // bn_acc_finalize_kernel (bn.cu) turns them into coefficients: with ~1600 short-lived CTAs per launch a per-CTA fence + ticket (the
  // last-CTA protocol of the persistent tcgen05 kernels) added ~2 us to every wave and made the fused form a net loss.
  constexpr int LDS = BN + 8;
  __nv_bfloat16* st = reinterpret_cast<__nv_bfloat16*>(smem);
  float* sRed = reinterpret_cast<float*>(smem + 40 * 1024);  // [WM][2][BN] behind the staging tile, inside the (now idle) ring
  static_assert(kBM * LDS * 2 <= 40 * 1024 && 40 * 1024 + WM * 2 * BN * 4 <= kStages * kStage, "epilogue scratch must fit the ring");
  float s1[TN][2], s2[TN][2];
#pragma unroll
  for (int j = 0; j < TN; ++j) s1[j][0] = s1[j][1] = s2[j][0] = s2[j][1] = 0.f;
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j)
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        const int row = wm * (kBM / WM) + i * 16 + (lane >> 2) + hh * 8;
        const int col = wn * (BN / WN) + j * 8 + (lane & 3) * 2;
        float v0 = acc[i][j][hh * 2], v1 = acc[i][j][hh * 2 + 1];
        const int m = m0 + row, n = n0 + col;
        const bool pv = m < a.M && n < a.N;
        if (a.bias && n < a.N) { v0 += a.bias[n]; v1 += a.bias[n + 1]; }
        float2 rf = make_float2(0.f, 0.f);
        if (a.res) rf = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(sres + row * LDR + col));
        if (BNM != 2) { v0 += rf.x; v1 += rf.y; }
        if (a.relu) { v0 = fmaxf(v0, 0.f); v1 = fmaxf(v1, 0.f); }
        __nv_bfloat162 o = __floats2bfloat162_rn(v0, v1);
        if constexpr (BNM == 1) {
          if (pv) {
            const float2 of = __bfloat1622float2(o);
            s1[j][0] += of.x; s1[j][1] += of.y;
            s2[j][0] += of.x * of.x; s2[j][1] += of.y * of.y;
          }
        } else if constexpr (BNM == 2) {
          float2 of = __bfloat1622float2(o);
          bool k0 = pv, k1 = pv;
          if (pv && a.bnb_relu) {
            k0 = relu_pass(rf.x * __ldg(a.bnb_coef + 2 * a.N + n) + __ldg(a.bnb_coef + 3 * a.N + n), a.bnb_relu);
            k1 = relu_pass(rf.y * __ldg(a.bnb_coef + 2 * a.N + n + 1) + __ldg(a.bnb_coef + 3 * a.N + n + 1), a.bnb_relu);
          }
          of.x = k0 ? of.x : 0.f;
          of.y = k1 ? of.y : 0.f;
          o = __floats2bfloat162_rn(of.x, of.y);
          s1[j][0] += of.x; s1[j][1] += of.y;
          s2[j][0] += of.x * rf.x; s2[j][1] += of.y * rf.y;
        }
        *reinterpret_cast<__nv_bfloat162*>(st + row * LDS + col) = o;
      }
  if constexpr (BNM != 0) {  // lanes sharing (lane & 3) hold the same channel pair: xor tree over lane bits 2..4, then the WM warps of a column
#pragma unroll
    for (int j = 0; j < TN; ++j)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        float t1 = s1[j][e], t2 = s2[j][e];
#pragma unroll
        for (int off = 4; off < 32; off <<= 1) {
          t1 += __shfl_xor_sync(0xffffffffu, t1, off);
          t2 += __shfl_xor_sync(0xffffffffu, t2, off);
        }
        if (lane < 4) {
          const int col = wn * (BN / WN) + j * 8 + lane * 2 + e;
          sRed[(wm * 2 + 0) * BN + col] = t1;
          sRed[(wm * 2 + 1) * BN + col] = t2;
        }
      }
  }
  __syncthreads();
  if (BNM != 0 && tid < 2 * BN) {
    const int which = tid / BN, c = tid - which * BN;
    if (n0 + c < a.N) {
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < WM; ++w) t += sRed[(w * 2 + which) * BN + c];
      atomicAdd(a.bn.acc + which * a.N + n0 + c, (double)t);
    }
  }
  constexpr int CV = BN / 8;
  for (int q = tid; q < kBM * CV; q += kThreads) {
    const int row = q / CV, c = q - row * CV;
    const int m = m0 + row, n = n0 + c * 8;
    if (m < a.M && n < a.N)
      *reinterpret_cast<uint4*>(a.Y + (int64_t)m * a.ldy + n) = *reinterpret_cast<const uint4*>(st + row * LDS + c * 8);
  }
}

template <int BN, int BNM>
int launch_bn(const G1Args& a, cudaStream_t stv) {
  constexpr int ring = kStages * (kBM * kBK * 2 + BN * kBK * 2), rtile = kBM * (BN + 8) * 2;
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(gemm1x1_kernel<BN, BNM>, cudaFuncAttributeMaxDynamicSharedMemorySize, ring + rtile);
    attr = true;
  }
  const int tiles_m = (a.M + kBM - 1) / kBM;
  gemm1x1_kernel<BN, BNM><<<tiles_m * a.tiles_n, kThreads, ring + (a.res ? rtile : 0), stv>>>(a);
  const int rc = check_launch("gemm1x1");
  if (rc || !a.bn_on) return rc;
  return launch_bn_acc_finalize(a.bn, a.N, a.bn_on, stv);
}

}  // namespace

bool gemm1x1_supported(const ConvP& p) {
  if (p.R != 1 || p.S != 1 || p.stride != 1 || p.up != 1 || p.pad_h != 0 || p.pad_w != 0) return false;
  if (p.y_f32 || p.ncls != 0) return false;
  if (p.bn) {  // BatchNorm epilogues: forward statistics (mode 1), fused backward reduction (mode 2: x travels in the residual slot)
    if (p.bn->fin.mode == 2) {
      if (p.res || !p.bnb_x || !p.bnb_coef || p.bnb_ldx % 8 != 0 || !aligned16(p.bnb_x)) return false;
    } else if (p.bn->fin.mode != 1) {
      return false;
    }
  }
  if (p.Cin % 8 != 0 || p.Cout % 8 != 0 || p.ldx % 8 != 0 || p.ldy % 8 != 0) return false;
  if (!aligned16(p.x) || !aligned16(p.w) || !aligned16(p.y)) return false;
  if (p.res && (p.ldr % 8 != 0 || !aligned16(p.res))) return false;
  if (p.M >= ((int64_t)1 << 31) - kBM) return false;
  return true;
}

int launch_gemm1x1(const ConvP& p, cudaStream_t st) {
  G1Args a;
  a.A = p.x; a.lda = p.ldx; a.B = p.w; a.Y = (__nv_bfloat16*)p.y; a.ldy = p.ldy; a.res = p.res; a.ldr = p.ldr; a.bias = p.bias;
  a.relu = p.relu; a.M = (int)p.M; a.N = p.Cout; a.K = p.Cin;
  a.bn_on = p.bn ? (p.bn->fin.mode == 2 ? 2 : 1) : 0;
  if (a.bn_on) a.bn = *p.bn; else a.bn = BnFuse{};
  a.bnb_coef = p.bnb_coef; a.bnb_relu = p.bnb_relu;
  if (a.bn_on == 2) { a.res = p.bnb_x; a.ldr = p.bnb_ldx; }
  if (p.Cout > 64) {
    a.tiles_n = (p.Cout + 127) / 128;
    return a.bn_on == 0 ? launch_bn<128, 0>(a, st) : a.bn_on == 1 ? launch_bn<128, 1>(a, st) : launch_bn<128, 2>(a, st);
  }
  a.tiles_n = 1;
  return a.bn_on == 0 ? launch_bn<64, 0>(a, st) : a.bn_on == 1 ? launch_bn<64, 1>(a, st) : launch_bn<64, 2>(a, st);
}

}  // namespace stp
