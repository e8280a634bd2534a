// Weight gradient of a 1x1 stride-1 convolution for the channel widths the tcgen05 wgrad kernel does not tile (MobileNetV2 / Xception:
// 24, 96, 144, 160, 728 ...):   dW[co][ci] = sum over pixels m of dY[m][co] * X[m][ci]   -- a GEMM whose reduction dimension is the
// pixel index and whose output (Cout x Cin <= 960 x 320) is tiny, so the job is to stream dY and X once.  The generic wgrad kernel
// (conv_generic.cu: im2col index arithmetic per element) ran these layers at 20 launches = 1.07 ms of the 7.6 ms DeepLabV3 step
// (profiles/r2_s10_launches_people.summary.txt).
//
// CTA = (128-channel slice of Cout) x (64-channel slice of Cin) x (one pixel split).  Per stage, 32 pixels of both tensors arrive
// as coalesced 16-byte cp.async copies into a 4-stage ring of [pixel][channel] rows (padded by 16 B: the transposing ldmatrix
// reads 8 pixel rows at one channel offset); both mma operands come out of ldmatrix.trans -- A = dY^T (rows = output channels,
// k = pixels), B = X (k = pixels, columns = input channels) -- into mma.sync.m16n8k16 with fp32 accumulation.  Partial tiles
// go to workspace [split][Cout][Cin] and are summed in a fixed order by split_reduce_kernel (deterministic, like wgrad_tc).
#include "conv.h"

namespace stp {
namespace {

constexpr int kWN = 128, kWK = 64, kWP = 32, kWStages = 4, kWThreads = 256;
constexpr int kLdA = kWN + 8, kLdB = kWK + 8;   // padded row lengths (elements)
constexpr int kStageA = kWP * kLdA * 2, kStageB = kWP * kLdB * 2, kWStage = kStageA + kStageB;

struct W1Args {
  const __nv_bfloat16* dy;
  int lddy;
  const __nv_bfloat16* x;
  int ldx;
  float* out;   // [splits][Cout][Cin]
  int64_t M, px_per_split;
  int Cout, Cin, tiles_n, tiles_k;
};

__device__ __forceinline__ void w1_cp_async16(uint32_t dst, const void* src, int bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(dst), "l"(src), "r"(bytes));
}
__device__ __forceinline__ void w1_cp_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void w1_cp_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }
__device__ __forceinline__ void w1_ldsm4t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];\n" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void w1_mma(float* c, const uint32_t* a, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__global__ void __launch_bounds__(kWThreads) wgrad1x1_kernel(const W1Args a) {
  extern __shared__ __align__(128) uint8_t smem[];
  const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(smem);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int tile = blockIdx.x % (a.tiles_n * a.tiles_k), split = blockIdx.x / (a.tiles_n * a.tiles_k);
  const int tn = tile / a.tiles_k, tk = tile - tn * a.tiles_k;
  const int n0 = tn * kWN, k0 = tk * kWK;
  const int64_t m_begin = (int64_t)split * a.px_per_split;
  int64_t m_end = m_begin + a.px_per_split;
  if (m_end > a.M) m_end = a.M;
  const int nst = m_end > m_begin ? (int)((m_end - m_begin + kWP - 1) / kWP) : 0;
  pdl_launch_dependents();
  pdl_wait();

  auto load_stage = [&](int it, int stage) {
    const uint32_t sa = sbase + stage * kWStage, sb = sa + kStageA;
    const int64_t m0 = m_begin + (int64_t)it * kWP;
#pragma unroll
    for (int i = 0; i < kWP * (kWN / 8) / kWThreads; ++i) {   // dY: 32 pixels x 16 chunks
      const int q = tid + i * kWThreads, row = q >> 4, c = q & 15;
      const int64_t m = m0 + row;
      const int n = n0 + c * 8;
      const bool ok = m < m_end && n < a.Cout;
      w1_cp_async16(sa + (uint32_t)(row * kLdA + c * 8) * 2, ok ? (const void*)(a.dy + m * a.lddy + n) : (const void*)a.dy, ok ? 16 : 0);
    }
    {   // X: 32 pixels x 8 chunks
      const int row = tid >> 3, c = tid & 7;
      const int64_t m = m0 + row;
      const int k = k0 + c * 8;
      const bool ok = m < m_end && k < a.Cin;
      w1_cp_async16(sb + (uint32_t)(row * kLdB + c * 8) * 2, ok ? (const void*)(a.x + m * a.ldx + k) : (const void*)a.x, ok ? 16 : 0);
    }
  };

  // warp grid 4 (output-channel rows) x 2 (input-channel columns): 32 x 32 per warp = 2 m16 tiles x 4 n8 tiles
  const int wn = warp & 3, wk = warp >> 2;
  float acc[2][4][4];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int q = 0; q < 4; ++q) acc[i][j][q] = 0.f;

#pragma unroll
  for (int s = 0; s < kWStages - 1; ++s) {
    if (s < nst) load_stage(s, s);
    w1_cp_commit();
  }
  const int lq = lane >> 3, lr = lane & 7;
  for (int it = 0; it < nst; ++it) {
    w1_cp_wait<kWStages - 2>();
    __syncthreads();
    if (it + kWStages - 1 < nst) load_stage(it + kWStages - 1, (it + kWStages - 1) % kWStages);
    w1_cp_commit();
    const uint32_t sa = sbase + (it % kWStages) * kWStage, sb = sa + kStageA;
#pragma unroll
    for (int ks = 0; ks < kWP / 16; ++ks) {
      uint32_t af[2][4], bf[4][2];
      // A(n, m) = dY[m][n]: matrix q of the x4 load = (n half q & 1, m half q >> 1) -> a0..a3 of the row-major A fragment
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int prow = ks * 16 + (lq >> 1) * 8 + lr;
        const int ch = wn * 32 + i * 16 + (lq & 1) * 8;
        w1_ldsm4t(sa + (uint32_t)(prow * kLdA + ch) * 2, af[i][0], af[i][1], af[i][2], af[i][3]);
      }
      // B(k = pixel, n = cin) = X[m][ci]: matrix q = (k half q & 1, n8 tile q >> 1) -> b0, b1 of two n8 tiles
#pragma unroll
      for (int jp = 0; jp < 2; ++jp) {
        const int prow = ks * 16 + (lq & 1) * 8 + lr;
        const int ch = wk * 32 + jp * 16 + (lq >> 1) * 8;
        w1_ldsm4t(sb + (uint32_t)(prow * kLdB + ch) * 2, bf[2 * jp][0], bf[2 * jp][1], bf[2 * jp + 1][0], bf[2 * jp + 1][1]);
      }
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) w1_mma(acc[i][j], af[i], bf[j][0], bf[j][1]);
    }
  }
  w1_cp_wait<0>();

  // fragment (row g / g + 8, columns 2t, 2t + 1) -> out[split][co][ci]
  float* o = a.out + (int64_t)split * a.Cout * a.Cin;
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        const int co = n0 + wn * 32 + i * 16 + (lane >> 2) + hh * 8;
        const int ci = k0 + wk * 32 + j * 8 + (lane & 3) * 2;
        if (co < a.Cout && ci < a.Cin) {   // (scalar stores: with one split `o` is the caller's dw, which need not be 8-byte aligned)
          o[(int64_t)co * a.Cin + ci] = acc[i][j][hh * 2];
          o[(int64_t)co * a.Cin + ci + 1] = acc[i][j][hh * 2 + 1];
        }
      }
}

struct W1Plan {
  int tiles_n, tiles_k, splits;
  int64_t px_per_split;
};

bool w1_plan(const WgradP& p, W1Plan* pl) {
  if (p.R != 1 || p.S != 1 || p.stride != 1 || p.up != 1 || p.pad_h != 0 || p.pad_w != 0) return false;
  if (p.Cin % 8 != 0 || p.Cout % 8 != 0 || p.ldx % 8 != 0 || p.lddy % 8 != 0 || !aligned16(p.x) || !aligned16(p.dy)) return false;
  if (p.M <= 0) return false;
  pl->tiles_n = (p.Cout + kWN - 1) / kWN;
  pl->tiles_k = (p.Cin + kWK - 1) / kWK;
  const int tiles = pl->tiles_n * pl->tiles_k;
  // ~4 CTAs per SM in flight, at least 8 stages (256 pixels) of work per CTA
  int64_t splits = (4 * kNumSMs + tiles - 1) / tiles;
  const int64_t max_splits = (p.M + 255) / 256;
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  int64_t pps = (p.M + splits - 1) / splits;
  pps = (pps + kWP - 1) / kWP * kWP;
  pl->px_per_split = pps;
  pl->splits = (int)((p.M + pps - 1) / pps);
  return true;
}

}  // namespace

bool wgrad1x1_supported(const WgradP& p) {
  W1Plan pl;
  return w1_plan(p, &pl);
}

size_t wgrad1x1_workspace(const WgradP& p) {
  W1Plan pl;
  if (!w1_plan(p, &pl)) return 0;
  return pl.splits > 1 ? (size_t)pl.splits * p.Cout * p.Cin * sizeof(float) : 0;
}

int launch_wgrad1x1(const WgradP& p, float* dw, void* ws, size_t ws_bytes, cudaStream_t st) {
  W1Plan pl;
  if (!w1_plan(p, &pl)) {
    set_error("wgrad1x1: unsupported");
    return STP_E_UNSUPPORTED;
  }
  const size_t need = pl.splits > 1 ? (size_t)pl.splits * p.Cout * p.Cin * sizeof(float) : 0;
  if (need > ws_bytes || (need && !ws)) {
    set_error("wgrad1x1: workspace too small (%zu < %zu)", ws_bytes, need);
    return STP_E_INVALID;
  }
  W1Args a;
  a.dy = p.dy; a.lddy = p.lddy; a.x = p.x; a.ldx = p.ldx;
  a.out = pl.splits > 1 ? (float*)ws : dw;
  a.M = p.M; a.px_per_split = pl.px_per_split; a.Cout = p.Cout; a.Cin = p.Cin; a.tiles_n = pl.tiles_n; a.tiles_k = pl.tiles_k;
  constexpr int smem = kWStages * kWStage;
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(wgrad1x1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    attr = true;
  }
  cudaError_t e = launch_pdl(wgrad1x1_kernel, dim3((unsigned)(pl.tiles_n * pl.tiles_k * pl.splits)), dim3(kWThreads), (size_t)smem, st, a);
  if (e != cudaSuccess) {
    set_error("wgrad1x1: launch: %s", cudaGetErrorString(e));
    return STP_E_CUDA;
  }
  int rc = check_launch("wgrad1x1");
  if (rc || pl.splits == 1) return rc;
  return launch_split_reduce((const float*)ws, pl.splits, (int64_t)p.Cout * p.Cin, dw, st);
}

}  // namespace stp
