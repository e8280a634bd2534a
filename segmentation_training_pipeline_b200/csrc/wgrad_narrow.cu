// wgrad for the narrow decoder-tail convolutions (3x3, stride 1, Cin and Cout in {16, 32}; U-Net dec3_c2 / dec4_c1 /
// dec4_c2 at 256^2..512^2).  These layers reduce over 1-4 M pixels into a tiny [Cout][3][3][Cin] result and sit on
// the HBM roofline (AI 72-144 FLOP/B): tcgen05 cannot help -- its minimum 128 x N>=64-ish MMA would burn 4-8x the
// tensor time on zero padding -- so this kernel uses the warp-level mma.sync m16n8k16 path whose 16x8 tiles fit
// the channel counts exactly, and concentrates on reading every byte ONCE:
//   * a producer warp TMA-loads (per 12x32-pixel tile) the dY rectangle and the input rectangle with a 1-pixel halo
//     (zero filled at the image border by the TMA unit) into a 3-stage shared-memory ring;
//   * 9 consumer warps = 3 filter rows x 3 pixel-row phases; the 3 filter COLUMNS are just +-1 pixel offsets of the
//     ldmatrix row addresses into the same haloed tile, so the input is fetched once for all 9 taps;
//   * accumulators stay in registers over the CTA's whole (persistent) pixel range; one deterministic cross-warp
//     reduction and one [blocks][Cout][3][3][Cin] partial write at the end, then split_reduce_kernel.
// dW[co][r][s][ci] = sum_p dY[p][co] * X[p + (r-1, s-1)][ci]     (TF Conv2DBackpropFilter of a 'same' 3x3 conv)
#include "conv.h"
#include "tc_common.cuh"

namespace stp {
namespace {

using namespace tc;

constexpr int TH = 12, TW = 32, XH = TH + 2, XW = TW + 2;
constexpr int kWarps = 9;
constexpr int kNThreads = (kWarps + 1) * 32;
constexpr int kStagesN = 3;

__device__ __forceinline__ void ldsm_x4_trans(uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3, const void* p) {
  uint32_t a = smem_u32(p);
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];\n"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(a));
}
__device__ __forceinline__ void mma16816(float* c, const uint32_t* a, const uint32_t* b) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

struct NarrowArgs {
  float* out;  // [gridDim.x][COUT][3][3][CIN]
  int tilesW, tilesH, num_tiles;
};

template <int CIN, int COUT>
struct NarrowCfg {
  static constexpr int kXBytes = XH * XW * CIN * 2;
  static constexpr int kYBytes = TH * TW * COUT * 2;
  static constexpr int kXPad = (kXBytes + 127) / 128 * 128;
  static constexpr int kStage = kXPad + (kYBytes + 127) / 128 * 128;
  static constexpr int kRed = COUT * 9 * CIN * 4;
  static constexpr int kSmem = kStagesN * kStage + 256 + 128;
  static_assert(kRed <= kStagesN * kStage, "reduction scratch reuses the pipeline buffers");
};

template <int CIN, int COUT>
__global__ void __launch_bounds__(kNThreads, 1)
wgrad_narrow_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmDY, const NarrowArgs a) {
  using Cfg = NarrowCfg<CIN, COUT>;
  constexpr int MT = COUT / 16, NT = CIN / 8;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((128u - (raw & 127u)) & 127u);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + kStagesN * Cfg::kStage);
  uint64_t* empty = full + kStagesN;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmX);
    prefetch_tmap(&tmDY);
    for (int i = 0; i < kStagesN; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], kWarps);
    }
    fence_barrier_init();
  }
  pdl_launch_dependents();
  __syncthreads();
  pdl_wait();

  float acc[3][MT][NT][4];
#pragma unroll
  for (int s = 0; s < 3; ++s)
#pragma unroll
    for (int i = 0; i < MT; ++i)
#pragma unroll
      for (int j = 0; j < NT; ++j)
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[s][i][j][e] = 0.f;

  if (warp == kWarps) {
    // ===== producer =====
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x) {
        const int tw = tile % a.tilesW;
        const int t = tile / a.tilesW;
        const int th = t % a.tilesH;
        const int img = t / a.tilesH;
        mbar_wait(&empty[stage], phase ^ 1);
        mbar_expect_tx(&full[stage], (uint32_t)(Cfg::kXBytes + Cfg::kYBytes));
        uint8_t* sx = smem + stage * Cfg::kStage;
        tma_load_4d(sx, &tmX, &full[stage], 0, tw * TW - 1, th * TH - 1, img);
        tma_load_4d(sx + Cfg::kXPad, &tmDY, &full[stage], 0, tw * TW, th * TH, img);
        if (++stage == kStagesN) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else {
    // ===== consumers: warp = (filter row r, pixel-row phase ps) =====
    const int r = warp / 3, ps = warp % 3;
    const int mj = lane >> 3, l7 = lane & 7;
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x) {
      mbar_wait(&full[stage], phase);
      const __nv_bfloat16* sx = reinterpret_cast<const __nv_bfloat16*>(smem + stage * Cfg::kStage);
      const __nv_bfloat16* sy = reinterpret_cast<const __nv_bfloat16*>(smem + stage * Cfg::kStage + Cfg::kXPad);
#pragma unroll 1
      for (int h = ps; h < TH; h += 3) {
#pragma unroll
        for (int wk = 0; wk < TW; wk += 16) {
          uint32_t af[MT][4];
#pragma unroll
          for (int i = 0; i < MT; ++i)
            ldsm_x4_trans(af[i][0], af[i][1], af[i][2], af[i][3],
                          sy + (h * TW + wk + (mj >> 1) * 8 + l7) * COUT + i * 16 + (mj & 1) * 8);
#pragma unroll
          for (int s = 0; s < 3; ++s) {
            const __nv_bfloat16* xrow = sx + ((h + r) * XW + wk + s + (mj & 1) * 8 + l7) * CIN + (mj >> 1) * 8;
#pragma unroll
            for (int j = 0; j < NT; j += 2) {
              uint32_t bf[2][2];
              ldsm_x4_trans(bf[0][0], bf[0][1], bf[1][0], bf[1][1], xrow + j * 8);
#pragma unroll
              for (int i = 0; i < MT; ++i) {
                mma16816(acc[s][i][j], af[i], bf[0]);
                mma16816(acc[s][i][j + 1], af[i], bf[1]);
              }
            }
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[stage]);
      if (++stage == kStagesN) {
        stage = 0;
        phase ^= 1;
      }
    }
  }

  // ---- deterministic cross-warp reduction over the 3 pixel-row phases, through the (now idle) pipeline smem ----
  __syncthreads();
  float* red = reinterpret_cast<float*>(smem);  // [COUT][3][3][CIN]
  for (int turn = 0; turn < 3; ++turn) {
    if (warp < kWarps && (warp % 3) == turn) {
      const int r = warp / 3;
#pragma unroll
      for (int s = 0; s < 3; ++s)
#pragma unroll
        for (int i = 0; i < MT; ++i)
#pragma unroll
          for (int j = 0; j < NT; ++j)
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int co = i * 16 + (lane >> 2) + (e >> 1) * 8;
              const int ci = j * 8 + (lane & 3) * 2 + (e & 1);
              float* p = red + ((co * 3 + r) * 3 + s) * CIN + ci;
              *p = (turn == 0) ? acc[s][i][j][e] : (*p + acc[s][i][j][e]);
            }
    }
    __syncthreads();
  }
  float* out = a.out + (int64_t)blockIdx.x * (COUT * 9 * CIN);
  for (int i = threadIdx.x; i < COUT * 9 * CIN; i += blockDim.x) out[i] = red[i];
}

template <int CIN, int COUT>
int launch_narrow(const WgradP& p, float* dw, void* ws, size_t ws_bytes, cudaStream_t st) {
  using Cfg = NarrowCfg<CIN, COUT>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(wgrad_narrow_kernel<CIN, COUT>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmem);
    if (e != cudaSuccess) {
      set_error("wgrad_narrow: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      return STP_E_CUDA;
    }
    attr_set = true;
  }
  NarrowArgs a;
  a.tilesW = (p.Wo + TW - 1) / TW;
  a.tilesH = (p.Ho + TH - 1) / TH;
  a.num_tiles = p.N * a.tilesW * a.tilesH;
  const int grid = a.num_tiles < kNumSMs ? a.num_tiles : kNumSMs;
  const size_t need = (size_t)grid * COUT * 9 * CIN * sizeof(float);
  if (!ws || ws_bytes < need) {
    set_error("conv_wgrad: workspace too small (%zu < %zu)", ws_bytes, need);
    return STP_E_WORKSPACE;
  }
  a.out = (float*)ws;
  CUtensorMap tmX, tmDY;
  {
    uint64_t dims[4] = {(uint64_t)CIN, (uint64_t)p.W, (uint64_t)p.H, (uint64_t)p.N};
    uint64_t strides[3] = {(uint64_t)p.ldx * 2, (uint64_t)p.W * p.ldx * 2, (uint64_t)p.H * p.W * p.ldx * 2};
    uint32_t box[4] = {(uint32_t)CIN, XW, XH, 1};
    if (!make_tmap_bf16(&tmX, p.x, 4, dims, strides, box, 0)) return STP_E_CUDA;
  }
  {
    uint64_t dims[4] = {(uint64_t)COUT, (uint64_t)p.Wo, (uint64_t)p.Ho, (uint64_t)p.N};
    uint64_t strides[3] = {(uint64_t)p.lddy * 2, (uint64_t)p.Wo * p.lddy * 2, (uint64_t)p.Ho * p.Wo * p.lddy * 2};
    uint32_t box[4] = {(uint32_t)COUT, TW, TH, 1};
    if (!make_tmap_bf16(&tmDY, p.dy, 4, dims, strides, box, 0)) return STP_E_CUDA;
  }
  launch_pdl(wgrad_narrow_kernel<CIN, COUT>, dim3(grid), dim3(kNThreads), (size_t)Cfg::kSmem, st, tmX, tmDY, a);
  g_tc_launches.fetch_add(1, std::memory_order_relaxed);  // counted with the TMA-fed kernels
  int rc = check_launch("wgrad_narrow");
  if (rc) return rc;
  return launch_split_reduce((const float*)ws, grid, (int64_t)COUT * 9 * CIN, dw, st);
}

}  // namespace

bool narrow_wgrad_supported(const WgradP& p) {
  if (p.stride != 1 || p.up != 1 || p.R != 3 || p.S != 3 || p.pad_h != 1 || p.pad_w != 1) return false;
  if (!(p.Cin == 16 || p.Cin == 32) || !(p.Cout == 16 || p.Cout == 32)) return false;
  if (p.H != p.Ho || p.W != p.Wo) return false;
  if (p.ldx % 8 != 0 || p.lddy % 8 != 0 || !aligned16(p.x) || !aligned16(p.dy)) return false;
  if ((int64_t)p.N * ((p.Wo + TW - 1) / TW) * ((p.Ho + TH - 1) / TH) > 0x7fffffff) return false;
  return get_encode_tiled() != nullptr;
}

size_t narrow_wgrad_workspace(const WgradP& p) {
  if (!narrow_wgrad_supported(p)) return 0;
  return (size_t)kNumSMs * p.Cout * 9 * p.Cin * sizeof(float);
}

int launch_narrow_wgrad(const WgradP& p, float* dw, void* ws, size_t ws_bytes, cudaStream_t st) {
  if (p.Cin == 16 && p.Cout == 16) return launch_narrow<16, 16>(p, dw, ws, ws_bytes, st);
  if (p.Cin == 32 && p.Cout == 16) return launch_narrow<32, 16>(p, dw, ws, ws_bytes, st);
  if (p.Cin == 16 && p.Cout == 32) return launch_narrow<16, 32>(p, dw, ws, ws_bytes, st);
  if (p.Cin == 32 && p.Cout == 32) return launch_narrow<32, 32>(p, dw, ws, ws_bytes, st);
  set_error("wgrad_narrow: unsupported channels");
  return STP_E_UNSUPPORTED;
}

}  // namespace stp
