// OPT-IN EXPERIMENT (option nconv = 1), MEASURED SLOWER than the tcgen05 halo kernel -- kept as a recorded negative result.
//
// 3x3 stride-1 "same" convolution (forward and dgrad) for the narrow decoder layers: Cin, Cout in {16, 32} at 256^2 / 512^2.
// These layers run at 0.31-0.46 of the copy bandwidth on the halo kernel (conv_tc2.cu), which feeds its UMMA operands with
// 32/64-byte TMA box rows and is paced by the TMA request rate (~3 clk per 32-byte row, profiles/README.md s15/s16).  The
// hypothesis tested here: build the operand path for the byte stream instead and take the math on the legacy tensor path -- a
// persistent CTA walks 8 x 32-pixel output tiles; the haloed 10 x 34-pixel input tile arrives as coalesced 16-byte cp.async
// copies (zero-filled outside the image) into a two-stage XOR-swizzled ring, one tile ahead of the math; each warp owns one
// output row (two m16 tiles) and takes its nine taps as shifted ldmatrix views of the same tile (mma.sync.m16n8k16 bf16 -> fp32);
// weights stay resident in shared memory (registers for 16 -> 16); the result leaves through a padded staging tile as 16-byte
// row-contiguous stores; epilogues follow conv_tc2's contract (bias / ReLU, forward BatchNorm statistics, fused BatchNorm
// backward, last-CTA finalize).  Result (profiles/r2_s9_nconv_bench.txt, bs16 back to back): 16 -> 16 @512^2 105.8 us against
// 108.4 us, every other shape slower (32 -> 32 @256^2 74.6 vs 39.4 us), the step 8.70 vs 8.53 ms.  Reason: mma.sync.m16n8k16
// issues at ~16 clk per SM sub-partition on sm_100a (~290 TF/s for the whole chip, one fifth of tcgen05), and these layers need
// 4.7 M of them per launch -- 65 us of issue time before any load or store.  The remedy for the narrow tail has to stay on
// tcgen05 (wider TMA rows through a pixel-merged view).
#include "conv.h"

namespace stp {
namespace {

constexpr int kTH = 8, kTW = 32, kHH = kTH + 2, kHW = kTW + 2, kNcThreads = 256;

struct NcArgs {
  const __nv_bfloat16* x;
  int ldx;
  const __nv_bfloat16* w;  // [Cout][3][3][Cin]
  __nv_bfloat16* y;
  int ldy;
  const float* bias;
  int relu;
  int N, H, W, tilesH, tilesW, num_tiles;
  int bn_on;  // 0 | 1 forward statistics of y | 2 fused BatchNorm backward (mask + reduce)
  BnFuse bn;
  const __nv_bfloat16* bnb_x;
  int bnb_ldx;
  const float* bnb_coef;
  int bnb_relu;
};

__device__ __forceinline__ void nc_cp_async16(uint32_t dst, const void* src, int bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(dst), "l"(src), "r"(bytes));
}
__device__ __forceinline__ void nc_cp_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void nc_cp_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }
__device__ __forceinline__ void nc_ldsm4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];\n" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void nc_mma(float* c, const uint32_t* a, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// byte offset of 16-byte chunk c of row `row` in a [rows][C] bf16 tile.  Rows are 32 (C = 16) or 64 (C = 32) bytes, so eight
// consecutive rows -- one ldmatrix phase, whatever the tap shift -- would fold onto 4 / 2 bank groups: the chunk index is XORed
// with row bits so that they cover all eight 16-byte bank groups of a 128-byte line.
template <int C>
__device__ __forceinline__ uint32_t nc_swz(int row, int c) {
  if (C == 16) return (uint32_t)(row * 32 + ((c ^ ((row >> 2) & 1)) << 4));
  return (uint32_t)(row * 64 + ((c ^ ((row >> 1) & 3)) << 4));
}

template <int CIN, int COUT>
struct NcCfg {
  static constexpr int kHalo = kHH * kHW * CIN * 2;
  static constexpr int kWB = 9 * COUT * CIN * 2;
  static constexpr int LDS = COUT + 8;  // staging pixel stride (elements): fragment stores and 16-byte row reads conflict-free
  static constexpr int kOut = kTH * kTW * LDS * 2;
  static constexpr int kRed = 8 * 2 * COUT * 4;
  static constexpr int kSmem = 2 * kHalo + kWB + kOut + kRed;
};

template <int CIN, int COUT>
__global__ void __launch_bounds__(kNcThreads, 2) conv_narrow_kernel(const NcArgs a) {
  using Cfg = NcCfg<CIN, COUT>;
  constexpr int CH = CIN / 8, NT = COUT / 8, KS = CIN / 16, LDS = Cfg::LDS;
  constexpr int kChunks = kHH * kHW * CH;
  constexpr int NLD = (kChunks + kNcThreads - 1) / kNcThreads;
  constexpr bool BREG = (CIN == 16 && COUT == 16);  // 36 registers of weight fragments
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ int s_last;
  uint8_t* sW = smem + 2 * Cfg::kHalo;
  __nv_bfloat16* sOut = reinterpret_cast<__nv_bfloat16*>(smem + 2 * Cfg::kHalo + Cfg::kWB);
  float* sRed = reinterpret_cast<float*>(smem + 2 * Cfg::kHalo + Cfg::kWB + Cfg::kOut);
  const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(smem);
  const uint32_t swb = sbase + 2 * Cfg::kHalo;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  pdl_launch_dependents();
  // per-thread halo chunks (fixed across tiles): halo row / column, channel chunk, swizzled shared-memory offset
  int ld_hr[NLD], ld_hc[NLD], ld_c[NLD];
  uint32_t ld_so[NLD];
#pragma unroll
  for (int i = 0; i < NLD; ++i) {
    const int q = tid + i * kNcThreads;
    const int pix = q / CH, c = q - pix * CH;
    const int hr = pix / kHW;
    ld_hr[i] = q < kChunks ? hr : -(1 << 20);
    ld_hc[i] = pix - hr * kHW;
    ld_c[i] = c * 8;
    ld_so[i] = nc_swz<CIN>(pix, c);
  }
  pdl_wait();

  // weights [Cout][9][Cin] -> shared [tap][Cout][Cin] rows, swizzled like the halo tile
  for (int q = tid; q < 9 * COUT * CH; q += kNcThreads) {
    const int n = q / (9 * CH), rem = q - n * (9 * CH);
    const int tap = rem / CH, c = rem - tap * CH;
    *reinterpret_cast<uint4*>(sW + nc_swz<CIN>(tap * COUT + n, c)) = __ldg(reinterpret_cast<const uint4*>(a.w) + q);
  }

  auto decode = [&](int tile, int& n, int& h0, int& w0) {
    const int tw = tile % a.tilesW, t2 = tile / a.tilesW;
    const int th = t2 % a.tilesH;
    n = t2 / a.tilesH;
    h0 = th * kTH;
    w0 = tw * kTW;
  };
  auto load_tile = [&](int tile, int stage) {
    int n, h0, w0;
    decode(tile, n, h0, w0);
    const uint32_t hb = sbase + stage * Cfg::kHalo;
#pragma unroll
    for (int i = 0; i < NLD; ++i) {
      if (i < NLD - 1 || tid + i * kNcThreads < kChunks) {
        const int gh = h0 - 1 + ld_hr[i], gw = w0 - 1 + ld_hc[i];
        const bool ok = (unsigned)gh < (unsigned)a.H && (unsigned)gw < (unsigned)a.W;
        const __nv_bfloat16* src = ok ? a.x + (((int64_t)n * a.H + gh) * a.W + gw) * a.ldx + ld_c[i] : a.x;
        nc_cp_async16(hb + ld_so[i], src, ok ? 16 : 0);
      }
    }
  };

  int tile = blockIdx.x;
  if (tile < a.num_tiles) load_tile(tile, 0);
  nc_cp_commit();
  __syncthreads();  // weights visible

  uint32_t breg[9][2][2];  // 16 -> 16 only: the nine 16 x 16 weight taps as mma B fragments (dead otherwise)
  if constexpr (BREG) {
#pragma unroll
    for (int tap = 0; tap < 9; ++tap)
      nc_ldsm4(swb + nc_swz<CIN>(tap * COUT + (lane & 7) + ((lane >> 4) << 3), (lane >> 3) & 1), breg[tap][0][0], breg[tap][0][1],
               breg[tap][1][0], breg[tap][1][1]);
  }

  // per-thread epilogue constants: this thread's channel pairs (fragment columns) nt*8 + (lane & 3)*2 + {0, 1}
  float bz[NT][2], sc[NT][2], sh[NT][2];
  float s1[NT][2], s2[NT][2];
#pragma unroll
  for (int nt = 0; nt < NT; ++nt)
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int ch = nt * 8 + (lane & 3) * 2 + e;
      bz[nt][e] = a.bias ? __ldg(a.bias + ch) : 0.f;
      sc[nt][e] = a.bn_on == 2 ? __ldg(a.bnb_coef + 2 * COUT + ch) : 0.f;
      sh[nt][e] = a.bn_on == 2 ? __ldg(a.bnb_coef + 3 * COUT + ch) : 0.f;
      s1[nt][e] = s2[nt][e] = 0.f;
    }

  for (int it = 0; tile < a.num_tiles; tile += gridDim.x, ++it) {
    const int stage = it & 1;
    if (tile + (int)gridDim.x < a.num_tiles) load_tile(tile + gridDim.x, stage ^ 1);
    nc_cp_commit();
    nc_cp_wait<1>();
    __syncthreads();  // (A) this tile's halo has landed; everyone has left the previous tile's staging reads

    float acc[2][NT][4];
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
      for (int nt = 0; nt < NT; ++nt)
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[j][nt][q] = 0.f;
    const uint32_t hb = sbase + stage * Cfg::kHalo;
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int s = 0; s < 3; ++s) {
        const int tap = r * 3 + s;
#pragma unroll
        for (int kk = 0; kk < KS; ++kk) {
          uint32_t af[2][4], bf[NT][2];
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            const int pix = (warp + r) * kHW + j * 16 + s + (lane & 15);
            nc_ldsm4(hb + nc_swz<CIN>(pix, kk * 2 + (lane >> 4)), af[j][0], af[j][1], af[j][2], af[j][3]);
          }
          if constexpr (BREG) {
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) {
              bf[nt][0] = breg[tap][nt][0];
              bf[nt][1] = breg[tap][nt][1];
            }
          } else {
#pragma unroll
            for (int np = 0; np < NT / 2; ++np)
              nc_ldsm4(swb + nc_swz<CIN>(tap * COUT + np * 16 + (lane & 7) + ((lane >> 4) << 3), kk * 2 + ((lane >> 3) & 1)),
                       bf[2 * np][0], bf[2 * np][1], bf[2 * np + 1][0], bf[2 * np + 1][1]);
          }
#pragma unroll
          for (int j = 0; j < 2; ++j)
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) nc_mma(acc[j][nt], af[j], bf[nt][0], bf[nt][1]);
        }
      }

    // epilogue on the fragments: bias / ReLU, one rounding to bf16, BatchNorm work on exactly the stored values
    int n, h0, w0;
    decode(tile, n, h0, w0);
    const int gh = h0 + warp;
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        const int col = j * 16 + (lane >> 2) + hh * 8;
        const int gw = w0 + col;
        const bool pv = gh < a.H && gw < a.W;
        const int pl = warp * kTW + col;
        const __nv_bfloat16* xp = nullptr;
        if (a.bn_on == 2 && pv) xp = a.bnb_x + (((int64_t)n * a.H + gh) * a.W + gw) * a.bnb_ldx + (lane & 3) * 2;
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
          float v0 = acc[j][nt][hh * 2] + bz[nt][0], v1 = acc[j][nt][hh * 2 + 1] + bz[nt][1];
          if (a.relu) { v0 = fmaxf(v0, 0.f); v1 = fmaxf(v1, 0.f); }
          __nv_bfloat162 o = __floats2bfloat162_rn(v0, v1);
          float2 of = __bfloat1622float2(o);
          if (a.bn_on == 1) {
            if (pv) {
              s1[nt][0] += of.x; s1[nt][1] += of.y;
              s2[nt][0] += of.x * of.x; s2[nt][1] += of.y * of.y;
            }
          } else if (a.bn_on == 2) {
            float2 xf = make_float2(0.f, 0.f);
            if (pv) xf = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(xp + nt * 8));
            const bool k0 = pv && (!a.bnb_relu || xf.x * sc[nt][0] + sh[nt][0] > 0.f);
            const bool k1 = pv && (!a.bnb_relu || xf.y * sc[nt][1] + sh[nt][1] > 0.f);
            of.x = k0 ? of.x : 0.f;
            of.y = k1 ? of.y : 0.f;
            o = __floats2bfloat162_rn(of.x, of.y);
            s1[nt][0] += of.x; s1[nt][1] += of.y;
            s2[nt][0] += of.x * xf.x; s2[nt][1] += of.y * xf.y;
          }
          *reinterpret_cast<__nv_bfloat162*>(sOut + pl * LDS + nt * 8 + (lane & 3) * 2) = o;
        }
      }
    __syncthreads();  // (B) staging tile complete; everyone is done with this stage's halo
#pragma unroll
    for (int i = 0; i < kTH * kTW * NT / kNcThreads; ++i) {
      const int q = tid + i * kNcThreads;
      const int pl = q / NT, c = q - pl * NT;
      const int oh = h0 + (pl >> 5), ow = w0 + (pl & 31);
      if (oh < a.H && ow < a.W)
        *reinterpret_cast<uint4*>(a.y + (((int64_t)n * a.H + oh) * a.W + ow) * a.ldy + c * 8) =
            *reinterpret_cast<const uint4*>(sOut + pl * LDS + c * 8);
    }
  }
  nc_cp_wait<0>();

  if (a.bn_on) {
    // lanes sharing (lane & 3) hold the same channel pair: fixed xor tree over lane bits 2..4, then the 8 warps through sRed
#pragma unroll
    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        float v1 = s1[nt][e], v2 = s2[nt][e];
#pragma unroll
        for (int off = 4; off < 32; off <<= 1) {
          v1 += __shfl_xor_sync(0xffffffffu, v1, off);
          v2 += __shfl_xor_sync(0xffffffffu, v2, off);
        }
        if (lane < 4) {
          sRed[(warp * 2 + 0) * COUT + nt * 8 + lane * 2 + e] = v1;
          sRed[(warp * 2 + 1) * COUT + nt * 8 + lane * 2 + e] = v2;
        }
      }
    __syncthreads();
    if (tid < 2 * COUT) {
      const int which = tid / COUT, c = tid - which * COUT;
      float t = 0.f;
#pragma unroll
      for (int wq = 0; wq < 8; ++wq) t += sRed[(wq * 2 + which) * COUT + c];
      atomicAdd(a.bn.acc + which * COUT + c, (double)t);
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) s_last = (atomicAdd(a.bn.fin.sync, 1u) == gridDim.x - 1) ? 1 : 0;
    __syncthreads();
    if (s_last) {  // every other CTA's sums have landed: finalise, return the accumulators to zero (conv_tc2's protocol)
      __threadfence();
      for (int c = tid; c < COUT; c += kNcThreads) {
        const double t1 = __ldcg(a.bn.acc + c), t2 = __ldcg(a.bn.acc + COUT + c);
        if (a.bn_on == 2) {  // sum g*xhat = invstd * (sum g*x - mean * sum g)
          const double mean = a.bn.fin.coef[c], invstd = a.bn.fin.coef[COUT + c];
          fin_backward(a.bn.fin, COUT, c, t1, invstd * (t2 - mean * t1));
        } else {
          fin_forward(a.bn.fin, COUT, c, t1, t2);
        }
        a.bn.acc[c] = 0.0;
        a.bn.acc[COUT + c] = 0.0;
      }
      if (tid == 0) *a.bn.fin.sync = 0u;
    }
  }
}

template <int CIN, int COUT>
int launch_nc(const NcArgs& a, cudaStream_t st) {
  using Cfg = NcCfg<CIN, COUT>;
  auto kernel = conv_narrow_kernel<CIN, COUT>;
  static bool attr = false;
  static int per_sm = 2;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmem);
    if (e != cudaSuccess) {
      set_error("conv_narrow: cudaFuncSetAttribute(%d B): %s", Cfg::kSmem, cudaGetErrorString(e));
      return STP_E_CUDA;
    }
    int n = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kernel, kNcThreads, Cfg::kSmem) == cudaSuccess && n > 0) per_sm = n;
    (void)cudaGetLastError();
    attr = true;
  }
  int grid = kNumSMs * per_sm;
  if (grid > a.num_tiles) grid = a.num_tiles;
  cudaError_t e = launch_pdl(kernel, dim3(grid), dim3(kNcThreads), (size_t)Cfg::kSmem, st, a);
  if (e != cudaSuccess) {
    set_error("conv_narrow: launch: %s", cudaGetErrorString(e));
    return STP_E_CUDA;
  }
  return check_launch("conv_narrow");
}

}  // namespace

bool narrow_conv_supported(const ConvP& p) {
  if (p.R != 3 || p.S != 3 || p.stride != 1 || p.up != 1 || p.pad_h != 1 || p.pad_w != 1) return false;
  if (!(p.Cin == 16 || p.Cin == 32) || !(p.Cout == 16 || p.Cout == 32)) return false;
  if (p.Ho != p.H || p.Wo != p.W || p.y_f32 || p.ncls != 0 || p.res) return false;
  if (p.ldx % 8 != 0 || p.ldy % 8 != 0 || !aligned16(p.x) || !aligned16(p.w) || !aligned16(p.y)) return false;
  if (p.bn && p.bn->fin.mode == 2 && (!p.bnb_x || !p.bnb_coef || p.bnb_ldx % 2 != 0)) return false;
  if (p.bn && p.bn->fin.mode != 1 && p.bn->fin.mode != 2) return false;
  // the HBM-bound regime this kernel is built for; small maps stay on the halo kernel
  if ((int64_t)p.N * ((p.H + kTH - 1) / kTH) * ((p.W + kTW - 1) / kTW) > 0x7fffffff) return false;
  return true;
}

int launch_narrow_conv(const ConvP& p, cudaStream_t st) {
  NcArgs a;
  a.x = p.x; a.ldx = p.ldx; a.w = p.w; a.y = (__nv_bfloat16*)p.y; a.ldy = p.ldy; a.bias = p.bias; a.relu = p.relu;
  a.N = p.N; a.H = p.H; a.W = p.W;
  a.tilesH = (p.H + kTH - 1) / kTH;
  a.tilesW = (p.W + kTW - 1) / kTW;
  a.num_tiles = p.N * a.tilesH * a.tilesW;
  a.bn_on = p.bn ? (p.bn->fin.mode == 2 ? 2 : 1) : 0;
  if (a.bn_on) a.bn = *p.bn; else a.bn = BnFuse{};
  a.bnb_x = p.bnb_x; a.bnb_ldx = p.bnb_ldx; a.bnb_coef = p.bnb_coef; a.bnb_relu = p.bnb_relu;
  if (p.Cin == 16 && p.Cout == 16) return launch_nc<16, 16>(a, st);
  if (p.Cin == 32 && p.Cout == 16) return launch_nc<32, 16>(a, st);
  if (p.Cin == 16 && p.Cout == 32) return launch_nc<16, 32>(a, st);
  return launch_nc<32, 32>(a, st);
}

}  // namespace stp
