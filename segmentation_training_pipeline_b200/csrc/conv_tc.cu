// tcgen05 / TMA implicit-GEMM convolution for sm_100a (stride 1, any RxS, NHWC bf16, fp32 accumulate in TMEM).
//
//   Y[(n,ho,wo), co] = sum_{r,s,ci} X[n, ho+r-pad_h, wo+s-pad_w, ci] * W[co, r, s, ci]
//
// GEMM view: M = output pixels, N = Cout, K = R*S*Cin.  One CTA computes a 128 x BN tile whose 128 rows are a
// BH x BW RECTANGLE of output pixels of one image (BH*BW = 128).  For filter tap (r,s) and channel block c0 the A
// operand is the input rectangle shifted by (r-pad_h, s-pad_w): ONE 4-D TMA box {BK ch, BW, BH, 1} whose out-of-
// range rows/columns are zero filled by the TMA unit == the convolution's zero padding, landing in shared memory
// directly in the K-major 128B/64B/32B-swizzled layout tcgen05.mma reads.  The B operand is a 2-D TMA box
// {BK, BN} of the [Cout][R*S*Cin] weight matrix.  No im2col buffer, no index arithmetic in the main loop.
//
// Warp roles (192 threads, persistent over tiles, 1 CTA / SM):
//   warp 0      TMA producer (one elected lane) : ring of NS stages, full/empty mbarriers
//   warp 1      TMEM allocator + MMA issuer (one elected lane): tcgen05.mma M=128 N=BN K=16, commit -> empty / accum-full
//   warps 2..5  epilogue: tcgen05.ld 32 lanes x 32 columns -> (+bias, +residual, ReLU) -> bf16/f32 NHWC stores;
//               the accumulator is double buffered in TMEM so the epilogue of tile i overlaps the MMAs of tile i+1.
//
// dgrad of a stride-1 conv is the same kernel on dY with tap-flipped [Cin][R][S][Cout] weights (stp_weight_prep).
// Replaces TF Conv2D / Conv2DBackpropInput reached from keras Conv2D in the graph built at reference
// segmentation.py:109-113,155.
#include "conv.h"
#include "tc_common.cuh"

namespace stp {

// ---- host: tensor maps ------------------------------------------------------------------------------
PFN_encodeTiled get_encode_tiled() {
  static PFN_encodeTiled fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (PFN_encodeTiled)p;
  }
  return fn;
}

static bool make_tmap_any(CUtensorMap* tm, CUtensorMapDataType dt, const void* base, int rank, const uint64_t* dims,
                          const uint64_t* strides_bytes, const uint32_t* box, uint32_t swizzle_bytes,
                          const uint32_t* elem_strides);

bool make_tmap_bf16(CUtensorMap* tm, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                    const uint32_t* box, uint32_t swizzle_bytes, const uint32_t* elem_strides) {
  return make_tmap_any(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, base, rank, dims, strides_bytes, box, swizzle_bytes, elem_strides);
}
bool make_tmap_f32(CUtensorMap* tm, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                   const uint32_t* box, uint32_t swizzle_bytes) {
  return make_tmap_any(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, base, rank, dims, strides_bytes, box, swizzle_bytes, nullptr);
}

static bool make_tmap_any(CUtensorMap* tm, CUtensorMapDataType dt, const void* base, int rank, const uint64_t* dims,
                          const uint64_t* strides_bytes, const uint32_t* box, uint32_t swizzle_bytes,
                          const uint32_t* elem_strides) {
  PFN_encodeTiled enc = get_encode_tiled();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled entry point unavailable");
    return false;
  }
  cuuint64_t gd[5], gs[4];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) {
    gd[i] = dims[i];
    bx[i] = box[i];
    es[i] = elem_strides ? elem_strides[i] : 1;
  }
  for (int i = 0; i + 1 < rank; ++i) gs[i] = strides_bytes[i];
  CUtensorMapSwizzle sw = swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                          : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                          : swizzle_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B
                                                : CU_TENSOR_MAP_SWIZZLE_NONE;
  CUresult r = enc(tm, dt, (cuuint32_t)rank, const_cast<void*>(base), gd, gs, bx, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d) rank %d dims %llu,%llu box %u,%u swz %u", (int)r, rank,
              (unsigned long long)dims[0], (unsigned long long)dims[1], box[0], box[1], swizzle_bytes);
    return false;
  }
  return true;
}

namespace {

using namespace tc;

constexpr int kThreads = 192;
constexpr int kSmemBudget = 227 * 1024;

// One output view + tap list.  A normal convolution has one; the dgrad of a stride-2 convolution has four (one per
// output parity class, each a stride-1 convolution over dY with a subset of the taps), all served by ONE launch.
struct TcClass {
  int64_t y_off, r_off;  // element offset of the view's pixel (0,0) in the output / residual tensors
  int Ho, Wo, tilesW, tilesH;
  int tile_begin;        // first tile index of this class
  int ntaps;             // 0: no tap reaches this class -> zeros (+ residual)
  int tap_dh[9], tap_dw[9], tap_k[9];  // A box offset (rows, pixels) and weight K block of each tap
};

struct TcConvArgs {
  void* y;
  const __nv_bfloat16* res;
  const float* bias;
  int y_f32, relu;
  int64_t y_sn, y_sh, y_sw;  // element strides of the output views (image, row, pixel)
  int64_t r_sn, r_sh, r_sw;  // ... and of the residual views
  int Cout, Cin;
  int stride;          // output stride (1 or 2): the A box is fetched with TMA element strides {1, stride, stride, 1}
  int BW, BH, log2BW;  // pixel rectangle of a tile, BW*BH = 128
  int tilesN;
  int num_tiles;
  int ncls;
  TcClass cls[4];
};

template <int BN, int BK>
struct TcCfg {
  static constexpr int kABytes = 128 * BK * 2;
  static constexpr int kBBytes = BN * BK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStagesRaw = (kSmemBudget - 2048) / kStageBytes;
  static constexpr int kStages = kStagesRaw > 8 ? 8 : kStagesRaw;
  static constexpr int kTmemCols = 2 * BN < 32 ? 32 : 2 * BN;
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align slack*/ + 256 /*barriers*/;
};

template <int BN, int BK>
__global__ void __launch_bounds__(kThreads, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const TcConvArgs a) {
  using Cfg = TcCfg<BN, BK>;
  constexpr int NS = Cfg::kStages;
  extern __shared__ uint8_t smem_raw[];
  // 1024-B aligned stage buffers (swizzle atoms), barriers after them
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw & 1023u)) & 1023u);
  uint8_t* sA = smem;
  uint8_t* sB = smem + NS * Cfg::kABytes;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + NS * Cfg::kStageBytes);
  uint64_t* empty = full + NS;
  uint64_t* acc_full = empty + NS;   // [2]
  uint64_t* acc_empty = acc_full + 2;  // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kcb = a.Cin / BK;            // channel blocks per tap
  __shared__ TcClass s_cls[4];           // dynamic indexing of kernel parameters would go through local memory
  if (threadIdx.x < 4) s_cls[threadIdx.x] = a.cls[threadIdx.x];

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
    for (int i = 0; i < NS; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&acc_full[i], 1);
      mbar_init(&acc_empty[i], 4);  // one arrive per epilogue warp
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, Cfg::kTmemCols);
    tmem_relinquish();
  }
  pdl_launch_dependents();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();

  auto decode = [&](int tile, int& c, int& img, int& h0, int& w0, int& n0) {
    c = 0;
    while (c + 1 < a.ncls && tile >= s_cls[c + 1].tile_begin) ++c;
    int t = tile - s_cls[c].tile_begin;
    int tn = t % a.tilesN;
    t /= a.tilesN;
    int tw = t % s_cls[c].tilesW;
    t /= s_cls[c].tilesW;
    int th = t % s_cls[c].tilesH;
    img = t / s_cls[c].tilesH;
    h0 = th * a.BH;
    w0 = tw * a.BW;
    n0 = tn * BN;
  };

  if (warp == 0) {
    // ================= TMA producer (warp-uniform loop, one elected lane issues: tc_common.cuh elect_one()) ==========
    {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x) {
        int c, img, h0, w0, n0;
        decode(tile, c, img, h0, w0, n0);
        const TcClass& cl = s_cls[c];
        for (int tap = 0; tap < cl.ntaps; ++tap) {
          for (int cb = 0; cb < kcb; ++cb) {
            mbar_wait(&empty[stage], phase ^ 1);
            if (elect_one()) {
              mbar_expect_tx(&full[stage], Cfg::kStageBytes);
              tma_load_4d(sA + stage * Cfg::kABytes, &tmA, &full[stage], cb * BK, w0 * a.stride + cl.tap_dw[tap],
                          h0 * a.stride + cl.tap_dh[tap], img);
              tma_load_2d(sB + stage * Cfg::kBBytes, &tmB, &full[stage], cl.tap_k[tap] * a.Cin + cb * BK, n0);
            }
            __syncwarp();
            if (++stage == NS) {
              stage = 0;
              phase ^= 1;
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer (warp-uniform loop, one elected lane issues) =================
    {
      constexpr uint32_t idesc = idesc_bf16(128, BN, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      int local = 0;
      for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x, ++local) {
        const int as = local & 1;
        const uint32_t aphase = (local >> 1) & 1;
        int c, img_, h0_, w0_, n0_;
        decode(tile, c, img_, h0_, w0_, n0_);
        const int num_kb = s_cls[c].ntaps * kcb;
        mbar_wait(&acc_empty[as], aphase ^ 1);  // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(as * BN);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint64_t ad0 = desc_kmajor(smem_u32(sA + stage * Cfg::kABytes), BK * 2);
          const uint64_t bd0 = desc_kmajor(smem_u32(sB + stage * Cfg::kBBytes), BK * 2);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < BK / 16; ++k) umma_bf16(d_tmem, ad0 + (uint64_t)(k * 2), bd0 + (uint64_t)(k * 2), idesc, (kb | k) != 0);
            umma_commit(&empty[stage]);  // smem slot reusable once these MMAs have read it
          }
          __syncwarp();
          if (++stage == NS) {
            stage = 0;
            phase ^= 1;
          }
        }
        if (elect_one()) umma_commit(&acc_full[as]);  // accumulator complete
        __syncwarp();
      }
    }
  } else {
    // ================= epilogue warps =================
    const int q = warp & 3;               // TMEM lane quarter this warp may access
    const int m = q * 32 + lane;          // tile row == TMEM lane
    const int hl = m >> a.log2BW, wl = m & (a.BW - 1);
    int local = 0;
    for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x, ++local) {
      const int as = local & 1;
      const uint32_t aphase = (local >> 1) & 1;
      int c, img, h0, w0, n0;
      decode(tile, c, img, h0, w0, n0);
      const int num_kb = s_cls[c].ntaps * kcb;
      const int ho = h0 + hl, wo = w0 + wl;
      const bool valid = ho < s_cls[c].Ho && wo < s_cls[c].Wo;
      const int64_t yoff = s_cls[c].y_off + (int64_t)img * a.y_sn + (int64_t)ho * a.y_sh + (int64_t)wo * a.y_sw;
      const int64_t roff = s_cls[c].r_off + (int64_t)img * a.r_sn + (int64_t)ho * a.r_sh + (int64_t)wo * a.r_sw;
      mbar_wait(&acc_full[as], aphase);
      tc_fence_after();
      const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * BN);
      constexpr int CH = BN >= 32 ? 32 : 16;
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += CH) {
        uint32_t rr[CH];
        if (num_kb > 0) {
          if constexpr (CH == 32) tmem_ld32(t_addr + c0, rr); else tmem_ld16(t_addr + c0, rr);
          tmem_ld_wait();
        } else {
#pragma unroll
          for (int i = 0; i < CH; ++i) rr[i] = 0u;
        }
        if (valid) {
          float v[CH];
#pragma unroll
          for (int i = 0; i < CH; ++i) v[i] = __uint_as_float(rr[i]);
          const int n = n0 + c0;
          if (a.bias) {
#pragma unroll
            for (int i = 0; i < CH; ++i) v[i] += __ldg(a.bias + n + i);
          }
          if (a.res) {
            const __nv_bfloat16* rp = a.res + roff + n;
#pragma unroll
            for (int i = 0; i < CH; i += 8) {
              float f[8];
              unpack8(ld8(rp + i), f);
#pragma unroll
              for (int j = 0; j < 8; ++j) v[i + j] += f[j];
            }
          }
          if (a.relu) {
#pragma unroll
            for (int i = 0; i < CH; ++i) v[i] = fmaxf(v[i], 0.f);
          }
          if (a.y_f32) {
            float* yp = reinterpret_cast<float*>(a.y) + yoff + n;
#pragma unroll
            for (int i = 0; i < CH; i += 4) *reinterpret_cast<float4*>(yp + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
          } else {
            __nv_bfloat16* yp = reinterpret_cast<__nv_bfloat16*>(a.y) + yoff + n;
#pragma unroll
            for (int i = 0; i < CH; i += 8) st8(yp + i, pack8(v + i));
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[as]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

// widest N tile that divides Cout and still leaves at least ~one tile per SM (m_tiles = 128-pixel tiles of the launch)
int pick_bn(int cout, int64_t m_tiles = 1 << 30) {
  int fallback = 0;
  for (int bn : {256, 128, 64, 32, 16})
    if (cout % bn == 0) {
      if (m_tiles * (cout / bn) >= kNumSMs - 20 || bn <= 64) return bn;
      fallback = bn;
    }
  return fallback;
}
int pick_bk(int cin) {
  if (cin % 64 == 0) return 64;
  if (cin == 32) return 32;
  if (cin == 16) return 16;
  return 0;
}

template <int BN, int BK>
int launch_cfg(const CUtensorMap& tmA, const CUtensorMap& tmB, const TcConvArgs& a, cudaStream_t st) {
  using Cfg = TcCfg<BN, BK>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(conv_tc_kernel<BN, BK>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes);
    if (e != cudaSuccess) {
      set_error("conv_tc: cudaFuncSetAttribute(%d B): %s", Cfg::kSmemBytes, cudaGetErrorString(e));
      return STP_E_CUDA;
    }
    attr_set = true;
  }
  int grid = a.num_tiles < kNumSMs ? a.num_tiles : kNumSMs;
  launch_pdl(conv_tc_kernel<BN, BK>, dim3(grid), dim3(kThreads), (size_t)Cfg::kSmemBytes, st, tmA, tmB, a);
  g_tc_launches.fetch_add(1, std::memory_order_relaxed);
  return check_launch("conv_tc");
}

}  // namespace

static bool tc_common_ok(const ConvP& p) {
  if (pick_bn(p.Cout) == 0 || pick_bk(p.Cin) == 0) return false;
  if (p.ldx % 8 != 0 || !aligned16(p.x) || !aligned16(p.w)) return false;
  if (p.y_f32 ? (p.ldy % 4 != 0) : (p.ldy % 8 != 0)) return false;
  if (!aligned16(p.y)) return false;
  if (p.res && (p.ldr % 8 != 0 || !aligned16(p.res))) return false;
  if (p.K % 8 != 0) return false;
  if (p.H > 65535 || p.W > 65535) return false;
  return get_encode_tiled() != nullptr;
}

// stride 1 / 2 forward-type problems (one launch), and the zero-insertion (up == 2) problems that the dgrad of a
// stride-2 convolution turns into (four launches, one per output parity class)
bool tc_conv_supported(const ConvP& p) {
  // 1x1 convolutions with a tiny reduction dimension and many output channels (MobileNetV2 "expand" layers and the dgrads of
  // its "project" layers: 16->96, 32->192, 64->384): one 16/32/64-deep MMA per 128 x BN tile, so the per-tile TMA -> MMA ->
  // TMEM -> store latency chain is all there is; the mma.sync kernel measured 0.090 vs 0.345 ms (16->96 @160^2 bs16), 0.021 vs
  // 0.046 (32->192 @40^2), 0.037 vs 0.053 (64->384 @40^2) -- profiles/r2_deeplab_bench.txt
  if (p.R * p.S == 1 && p.up == 1 && (p.Cin <= 32 || (p.Cin == 64 && p.Cout > 256)) && p.Cout >= 3 * p.Cin) return false;
  if (p.up == 1) {
    if (p.stride != 1 && p.stride != 2) return false;
    if (p.R * p.S > 9) return false;
    if (p.Wo < 8 || p.Ho < 1) return false;
  } else {
    if (p.up != 2 || p.stride != 1 || p.R > 3 || p.S > 3) return false;
    if (p.Wo < 16 || p.Ho < 2) return false;
  }
  return tc_common_ok(p);
}

struct ViewSpec {  // host description of one class
  int ho, wo;
  int64_t y_off, r_off;
  int ntaps, dh[9], dw[9], tk[9];
};

// One launch over `ncls` output views (element strides shared), each with an explicit tap list.
static int launch_tc_views(const ConvP& p, int ncls, const ViewSpec* v, int64_t y_sn, int64_t y_sh, int64_t y_sw,
                           int64_t r_sn, int64_t r_sh, int64_t r_sw, int in_stride, cudaStream_t st) {
  const int BK = pick_bk(p.Cin);
  TcConvArgs a;
  a.y = p.y; a.res = p.res; a.bias = p.bias; a.y_f32 = p.y_f32; a.relu = p.relu;
  a.y_sn = y_sn; a.y_sh = y_sh; a.y_sw = y_sw; a.r_sn = r_sn; a.r_sh = r_sh; a.r_sw = r_sw;
  a.Cout = p.Cout; a.Cin = p.Cin;
  a.stride = in_stride;
  int wmax = 0;
  for (int c = 0; c < ncls; ++c) wmax = v[c].wo > wmax ? v[c].wo : wmax;
  int bw = 128;
  while (bw > 8 && bw / 2 >= wmax) bw /= 2;  // smallest power of two >= Wo, clamped to [8,128]
  a.BW = bw; a.BH = 128 / bw;
  a.log2BW = 0;
  while ((1 << a.log2BW) < bw) ++a.log2BW;
  int64_t m_tiles = 0;
  for (int c = 0; c < ncls; ++c)
    m_tiles += (int64_t)p.N * ((v[c].wo + a.BW - 1) / a.BW) * ((v[c].ho + a.BH - 1) / a.BH);
  const int BN = pick_bn(p.Cout, m_tiles);
  a.tilesN = p.Cout / BN;
  a.ncls = ncls;
  int64_t nt = 0;
  for (int c = 0; c < 4; ++c) {
    TcClass& k = a.cls[c];
    const ViewSpec& s = v[c < ncls ? c : 0];
    k.y_off = s.y_off; k.r_off = s.r_off; k.Ho = s.ho; k.Wo = s.wo;
    k.tilesW = (s.wo + a.BW - 1) / a.BW;
    k.tilesH = (s.ho + a.BH - 1) / a.BH;
    k.tile_begin = (int)nt;
    k.ntaps = s.ntaps;
    for (int t = 0; t < 9; ++t) {
      k.tap_dh[t] = t < s.ntaps ? s.dh[t] : 0;
      k.tap_dw[t] = t < s.ntaps ? s.dw[t] : 0;
      k.tap_k[t] = t < s.ntaps ? s.tk[t] : 0;
    }
    if (c < ncls) nt += (int64_t)p.N * k.tilesH * k.tilesW * a.tilesN;
    if (nt > 0x7fffffff) {
      set_error("conv_tc: too many tiles");
      return STP_E_UNSUPPORTED;
    }
  }
  a.num_tiles = (int)nt;

  CUtensorMap tmA, tmB;
  {
    uint64_t dims[4] = {(uint64_t)p.Cin, (uint64_t)p.W, (uint64_t)p.H, (uint64_t)p.N};
    uint64_t strides[3] = {(uint64_t)p.ldx * 2, (uint64_t)p.W * p.ldx * 2, (uint64_t)p.H * p.W * p.ldx * 2};
    // stride 2: the box spans stride*BW x stride*BH input pixels and the TMA unit keeps every stride-th one
    uint32_t box[4] = {(uint32_t)BK, (uint32_t)(a.BW * in_stride), (uint32_t)(a.BH * in_stride), 1};
    uint32_t es[4] = {1, (uint32_t)in_stride, (uint32_t)in_stride, 1};
    if (!make_tmap_bf16(&tmA, p.x, 4, dims, strides, box, BK * 2, es)) return STP_E_CUDA;
  }
  {
    uint64_t dims[2] = {(uint64_t)p.K, (uint64_t)p.Cout};
    uint64_t strides[1] = {(uint64_t)p.K * 2};
    uint32_t box[2] = {(uint32_t)BK, (uint32_t)BN};
    if (!make_tmap_bf16(&tmB, p.w, 2, dims, strides, box, BK * 2)) return STP_E_CUDA;
  }
#define STP_TC_CASE(bn, bk) \
  if (BN == bn && BK == bk) return launch_cfg<bn, bk>(tmA, tmB, a, st);
  STP_TC_CASE(256, 64) STP_TC_CASE(128, 64) STP_TC_CASE(64, 64) STP_TC_CASE(32, 64) STP_TC_CASE(16, 64)
  STP_TC_CASE(256, 32) STP_TC_CASE(128, 32) STP_TC_CASE(64, 32) STP_TC_CASE(32, 32) STP_TC_CASE(16, 32)
  STP_TC_CASE(256, 16) STP_TC_CASE(128, 16) STP_TC_CASE(64, 16) STP_TC_CASE(32, 16) STP_TC_CASE(16, 16)
#undef STP_TC_CASE
  set_error("conv_tc: no specialisation for BN=%d BK=%d", BN, BK);
  return STP_E_UNSUPPORTED;
}

int launch_tc_conv(const ConvP& p, cudaStream_t st) {
  if (p.up == 1) {
    if (p.R * p.S > 9) {
      set_error("conv_tc: more than 9 taps");
      return STP_E_UNSUPPORTED;
    }
    ViewSpec v;
    v.ho = p.Ho; v.wo = p.Wo; v.y_off = 0; v.r_off = 0; v.ntaps = p.R * p.S;
    for (int r = 0; r < p.R; ++r)
      for (int s = 0; s < p.S; ++s) {
        v.dh[r * p.S + s] = r - p.pad_h;
        v.dw[r * p.S + s] = s - p.pad_w;
        v.tk[r * p.S + s] = r * p.S + s;
      }
    return launch_tc_views(p, 1, &v, (int64_t)p.Ho * p.Wo * p.ldy, (int64_t)p.Wo * p.ldy, p.ldy,
                           (int64_t)p.Ho * p.Wo * p.ldr, (int64_t)p.Wo * p.ldr, p.ldr, p.stride, st);
  }
  // up == 2:  y[i] = sum_r' xup[i - pad + r'] w[r'],  xup[2o] = x[o].  For the output parity class i = 2a + ph only the
  // taps with (ph - pad + r') even contribute, reading x[a + (ph - pad + r')/2]: a stride-1 conv per class.
  // A class no tap reaches is all zeros (+ residual): skipped when the call accumulates in place (residual == output).
  const bool in_place = p.res && (const void*)p.res == (const void*)p.y && p.ldr == p.ldy && !p.y_f32;
  ViewSpec v[4];
  int ncls = 0;
  for (int ph = 0; ph < 2; ++ph)
    for (int pw = 0; pw < 2; ++pw) {
      ViewSpec& k = v[ncls];
      k.ntaps = 0;
      for (int r = 0; r < p.R; ++r) {
        if ((ph - p.pad_h + r) & 1) continue;
        for (int s = 0; s < p.S; ++s) {
          if ((pw - p.pad_w + s) & 1) continue;
          k.dh[k.ntaps] = (ph - p.pad_h + r) / 2;  // numerator is even: exact for negatives too
          k.dw[k.ntaps] = (pw - p.pad_w + s) / 2;
          k.tk[k.ntaps] = r * p.S + s;
          ++k.ntaps;
        }
      }
      k.ho = (p.Ho - ph + 1) / 2;
      k.wo = (p.Wo - pw + 1) / 2;
      if (k.ho <= 0 || k.wo <= 0) continue;
      if (k.ntaps == 0 && in_place && !p.relu && !p.bias) continue;
      k.y_off = ((int64_t)ph * p.Wo + pw) * p.ldy;
      k.r_off = ((int64_t)ph * p.Wo + pw) * p.ldr;
      ++ncls;
    }
  if (ncls == 0) return STP_OK;
  return launch_tc_views(p, ncls, v, (int64_t)p.Ho * p.Wo * p.ldy, (int64_t)2 * p.Wo * p.ldy, (int64_t)2 * p.ldy,
                         (int64_t)p.Ho * p.Wo * p.ldr, (int64_t)2 * p.Wo * p.ldr, (int64_t)2 * p.ldr, 1, st);
}

// ====================================================================================================
// wgrad on tcgen05:  dW[co][r][s][ci] = sum over output pixels p of dY[p][co] * X[p + (r,s) - pad][ci]
//
// GEMM view: M = co (128 rows / MMA), N = ci tile (64 or 128), K = pixels.  Both operands are "MN-major"
// (channels contiguous, pixels = K rows of 128 B) which is exactly what a 128B-swizzled TMA box
// {64 ch, BW, BH(+2), 1} of the NHWC tensors produces -- no transposition anywhere.
// One CTA = (co tile, ci tile, filter column s, pixel split).  Per 128-pixel rectangle it loads dY once and the
// input rectangle shifted by s WITH a +-1 row halo once; the R filter rows are then R shifted views of the same
// shared-memory tile (shift by r rows = r*BW*128 B, a whole number of swizzle atoms), each accumulating into its
// own TMEM accumulator [128 x BN].  Split partials are reduced deterministically by split_reduce_kernel.
// ====================================================================================================
struct TcWgradArgs {
  float* out;  // [splits][Cout][R][S][Cin]
  int Cout, Cin, R, S, pad_h, pad_w;
  int BW, BH, tilesW, tilesH;
  int num_pb, pb_per_split, splits, co_tiles, ci_tiles;
  int a_boxes;    // TMA boxes of dY per stage: 2 when Cout >= 128, else 1
  int a_row;      // bytes per pixel row of a dY box: min(Cout,64)*2  (== its swizzle mode)
  int a_lbo;      // byte distance between consecutive MN atoms of A (0: rows beyond the first atom alias it, discarded)
  int m_valid;    // accumulator rows that hold real output channels
  int b_row;      // bytes per pixel row of an X box: min(Cin,64)*2
  int xrows;      // (BH + R - 1) * BW rows of the haloed input tile (stride 2: 128, one box per filter row)
  int stride;     // 1: the R filter rows are shifted views of ONE haloed box; 2: one element-strided TMA box per row
  int dbg;        // timing experiments only (results invalid): 1 skip dY loads, 2 skip X loads, 4 skip stores, 8 skip MMAs
};

// ANARROW: Cout <= 32 (dY rows of 32/64 B) -> small A stage; otherwise two 128-B-row boxes.
template <int BN, int ANARROW, int STR>
struct TcWgradCfg {
  static constexpr int kABytes = ANARROW ? 128 * 64 : 2 * 16384;
  static constexpr int kBRow = (BN < 64 ? BN : 64) * 2;
  // >= (BH+R-1)*BW rows: 160 for R <= 3; the 32-channel variant also serves the 4x4 space-to-depth stem (176 rows)
  static constexpr int kXBoxBytes = (BN == 32 ? 192 : 160) * kBRow;
  static constexpr int kRBytes = (BN < 64 ? 1 : BN / 64) * kXBoxBytes;  // input box(es) of one filter row
  static constexpr int kBBytes = (STR ? 3 : 1) * kRBytes;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStagesRaw = (kSmemBudget - 2048) / kStageBytes;
  static constexpr int kStages = kStagesRaw > 12 ? 12 : kStagesRaw;
  static_assert(kStages >= 2, "wgrad needs two pipeline stages");
  static constexpr int kTmemCols = 4 * BN <= 64 ? 64 : 4 * BN <= 128 ? 128 : 4 * BN <= 256 ? 256 : 512;  // up to 4 filter rows
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 + 256;
  static_assert(kXBoxBytes % 1024 == 0, "stage buffers must stay swizzle-atom aligned");
};

template <int BN, int ANARROW, int STR>
__global__ void __launch_bounds__(kThreads, 1)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap tmDY, const __grid_constant__ CUtensorMap tmX,
                const __grid_constant__ CUtensorMap tmOut, const TcWgradArgs a) {
  using Cfg = TcWgradCfg<BN, ANARROW, STR>;
  constexpr int NS = Cfg::kStages;
  constexpr int BBOX = BN < 64 ? 1 : BN / 64;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw & 1023u)) & 1023u);
  uint8_t* sA = smem;
  uint8_t* sB = smem + NS * Cfg::kABytes;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + NS * Cfg::kStageBytes);
  uint64_t* empty = full + NS;
  uint64_t* acc_full = empty + NS;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // work item
  int item = blockIdx.x;
  const int split = item % a.splits; item /= a.splits;
  const int s = item % a.S; item /= a.S;
  const int cit = item % a.ci_tiles;
  const int cot = item / a.ci_tiles;
  const int co0 = cot * 128, ci0 = cit * BN;
  const int pb_begin = split * a.pb_per_split;
  int pb_end = pb_begin + a.pb_per_split;
  if (pb_end > a.num_pb) pb_end = a.num_pb;
  const int npb = pb_end > pb_begin ? pb_end - pb_begin : 0;
  const uint32_t a_box_bytes = 128u * (uint32_t)a.a_row;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmDY);
    prefetch_tmap(&tmX);
    for (int i = 0; i < NS; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    mbar_init(acc_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, Cfg::kTmemCols);
    tmem_relinquish();
  }
  pdl_launch_dependents();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();

  if (warp == 0) {
    {  // warp-uniform loop, one elected lane issues (tc_common.cuh elect_one())
      const uint32_t tx_bytes = (uint32_t)a.a_boxes * a_box_bytes +
                                (uint32_t)(STR ? a.R : 1) * (uint32_t)BBOX * (uint32_t)a.xrows * (uint32_t)a.b_row;
      int stage = 0;
      uint32_t phase = 0;
      for (int pb = pb_begin; pb < pb_end; ++pb) {
        int tw = pb % a.tilesW;
        int t = pb / a.tilesW;
        int th = t % a.tilesH;
        int img = t / a.tilesH;
        const int h0 = th * a.BH, w0 = tw * a.BW;
        mbar_wait(&empty[stage], phase ^ 1);
        uint8_t* pa = sA + stage * Cfg::kABytes;
        uint8_t* pbuf = sB + stage * Cfg::kBBytes;
        if (elect_one()) {
        if (a.dbg & 3) {
          const uint32_t txa = (a.dbg & 1) ? 0u : (uint32_t)a.a_boxes * a_box_bytes;
          const uint32_t txb = (a.dbg & 2) ? 0u : tx_bytes - (uint32_t)a.a_boxes * a_box_bytes;
          if (txa + txb) mbar_expect_tx(&full[stage], txa + txb); else mbar_arrive(&full[stage]);
        } else {
          mbar_expect_tx(&full[stage], tx_bytes);
        }
        if (!(a.dbg & 1))
        for (int j = 0; j < a.a_boxes; ++j) tma_load_4d(pa + j * a_box_bytes, &tmDY, &full[stage], co0 + 64 * j, w0, h0, img);
        if (a.dbg & 2) {
        } else if (STR) {
          for (int r = 0; r < a.R; ++r)
#pragma unroll
            for (int j = 0; j < BBOX; ++j)
              tma_load_4d(pbuf + r * Cfg::kRBytes + j * Cfg::kXBoxBytes, &tmX, &full[stage], ci0 + 64 * j,
                          w0 * a.stride + s - a.pad_w, h0 * a.stride + r - a.pad_h, img);
        } else {
#pragma unroll
          for (int j = 0; j < BBOX; ++j)
            tma_load_4d(pbuf + j * Cfg::kXBoxBytes, &tmX, &full[stage], ci0 + 64 * j, w0 + s - a.pad_w, h0 - a.pad_h, img);
        }
        }
        __syncwarp();
        if (++stage == NS) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    {
      constexpr uint32_t idesc = idesc_bf16(128, BN, 1, 1);
      const uint32_t a_step = 16u * (uint32_t)a.a_row, b_step = 16u * (uint32_t)a.b_row;  // 16 pixels (one MMA K) further
      int stage = 0;
      uint32_t phase = 0;
      for (int i = 0; i < npb; ++i) {
        mbar_wait(&full[stage], phase);
        tc_fence_after();
        // descriptors of the stage's first K step; later steps / filter rows only move the 14-bit start-address field
        const uint64_t ad0 = desc_mnmajor(smem_u32(sA + stage * Cfg::kABytes), (uint32_t)a.a_lbo, (uint32_t)a.a_row);
        const uint64_t bd0 = desc_mnmajor(smem_u32(sB + stage * Cfg::kBBytes), Cfg::kXBoxBytes, (uint32_t)a.b_row);
        if (elect_one()) {
          if (!(a.dbg & 8))
          for (int r = 0; r < a.R; ++r) {
            const uint64_t b_r = bd0 + (uint64_t)((STR ? (uint32_t)r * (uint32_t)Cfg::kRBytes
                                                       : (uint32_t)(r * a.BW) * (uint32_t)a.b_row) >> 4);
#pragma unroll
            for (int kk = 0; kk < 8; ++kk)
              umma_bf16(tmem_base + (uint32_t)(r * BN), ad0 + (uint64_t)((kk * a_step) >> 4), b_r + (uint64_t)((kk * b_step) >> 4),
                        idesc, (i | kk) != 0);
          }
          umma_commit(&empty[stage]);
        }
        __syncwarp();
        if (++stage == NS) {
          stage = 0;
          phase ^= 1;
        }
      }
      if (elect_one()) umma_commit(acc_full);
      __syncwarp();
    }
  } else {
    const int q = warp & 3;
    const int co = co0 + q * 32 + lane;
    mbar_wait(acc_full, 0);
    tc_fence_after();
    // The accumulators leave through a swizzled per-warp staging tile [32 rows (co)][CH fp32] and 3-D TMA stores into
    // out[split][co][K] (coalesced 128-byte rows; rows beyond Cout are clipped by the tensor map).  Direct per-thread
    // stores would scatter 16-byte pieces over 32 rows K*4 bytes apart -- measured as ~half of this kernel's time.
    // The pipeline stage buffers are idle by now (every MMA has completed) and serve as staging, double buffered.
    constexpr int CH = BN >= 32 ? 32 : 16;
    constexpr int RBO = CH * 4;                       // bytes per staged row: 128 or 64 (== TMA swizzle span)
    uint8_t* stg = sA + q * (2 * 32 * RBO);
    const bool rows_live = co0 + q * 32 < a.Cout && q * 32 < a.m_valid;  // warp-uniform: any real output channel here?
    const uint32_t swz = RBO == 128 ? (uint32_t)(lane & 7) : (uint32_t)((lane >> 1) & 3);
    int buf = 0;
    (void)co;
    for (int r = 0; r < a.R; ++r) {
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += CH) {
        uint32_t rr[CH];
        if (npb > 0) {
          const uint32_t ta = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(r * BN + c0);
          if constexpr (CH == 32) tmem_ld32(ta, rr); else tmem_ld16(ta, rr);
          tmem_ld_wait();
        } else {
#pragma unroll
          for (int i = 0; i < CH; ++i) rr[i] = 0u;
        }
        if (!rows_live) continue;
        if (lane == 0) tma_store_wait_read1();  // the store that used this buffer two rounds ago has drained it
        __syncwarp();
        uint8_t* row = stg + buf * (32 * RBO) + lane * RBO;
#pragma unroll
        for (int i = 0; i < CH; i += 4)
          *reinterpret_cast<uint4*>(row + ((((uint32_t)i >> 2) ^ swz) << 4)) = make_uint4(rr[i], rr[i + 1], rr[i + 2], rr[i + 3]);
        fence_proxy_async();
        __syncwarp();
        if (lane == 0 && !(a.dbg & 4)) {
          tma_store_3d(&tmOut, stg + buf * (32 * RBO), ((r * a.S + s) * a.Cin) + ci0 + c0, co0 + q * 32, split);
          tma_store_commit();
        }
        buf ^= 1;
      }
    }
    if (lane == 0) tma_store_wait_all();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

struct WgradPlan {
  int BN, BW, BH, tilesW, tilesH, num_pb, splits, pb_per_split, co_tiles, ci_tiles, items;
};

static bool wgrad_plan(int64_t M, int N_img, int Ho, int Wo, int Cout, int Cin, int R, int S, int stride, WgradPlan* pl) {
  const bool co_ok = Cout % 64 == 0 || Cout == 32 || Cout == 16;
  const bool ci_ok = Cin % 64 == 0 || Cin == 32 || Cin == 16;
  const int rmax = (Cin == 32 && stride == 1) ? 4 : 3;
  if (!co_ok || !ci_ok || R > rmax || S > rmax || Wo < 8) return false;
  pl->BN = (Cin % 128 == 0 && stride == 1) ? 128 : (Cin % 64 == 0 ? 64 : Cin);
  pl->BW = Wo >= 16 ? 16 : 8;
  pl->BH = 128 / pl->BW;
  if ((pl->BH + R - 1) * pl->BW > (pl->BN == 32 ? 192 : 160)) return false;
  pl->tilesW = (Wo + pl->BW - 1) / pl->BW;
  pl->tilesH = (Ho + pl->BH - 1) / pl->BH;
  int64_t npb = (int64_t)N_img * pl->tilesW * pl->tilesH;
  if (npb > 0x7fffffff) return false;
  pl->num_pb = (int)npb;
  pl->co_tiles = (Cout + 127) / 128;
  pl->ci_tiles = Cin / pl->BN;
  int base = pl->co_tiles * pl->ci_tiles * S;
  // One CTA per SM is resident (shared memory), so the grid should be a whole number of waves of (nearly) equal
  // items: pick the split count whose item total fills k waves best (k = 1, 2), preferring fewer splits (less
  // partial traffic, longer main loops) unless that leaves SMs idle.
  int max_s = (pl->num_pb + 3) / 4;  // at least 4 pixel blocks per split
  if (max_s < 1) max_s = 1;
  // keep the fp32 split partials L2 resident (126 MB L2): <= 48 MB in flight
  int64_t per = (int64_t)Cout * R * S * Cin * 4;
  int cap = (int)((48ll << 20) / per);
  if (cap < 1) cap = 1;
  if (max_s > cap) max_s = cap;
  int splits = 1;
  double best = -1.0;
  for (int k = 1; k <= 2; ++k) {
    int sp = (k * kNumSMs) / base;
    if (sp < 1) sp = 1;
    if (sp > max_s) sp = max_s;
    const int items = sp * base;
    const int waves = (items + kNumSMs - 1) / kNumSMs;
    // efficiency of the wave fill, discounted by the fixed per-CTA cost (prologue + accumulator write-out ~ 3 blocks)
    const double pb = (double)pl->num_pb / sp;
    const double eff = ((double)items / (waves * kNumSMs)) * (pb / (pb + 3.0));
    if (eff > best) {
      best = eff;
      splits = sp;
    }
  }
  pl->pb_per_split = (pl->num_pb + splits - 1) / splits;
  pl->splits = (pl->num_pb + pl->pb_per_split - 1) / pl->pb_per_split;
  pl->items = base * pl->splits;
  (void)M;
  return true;
}

bool tc_wgrad_supported(const WgradP& p) {
  if ((p.stride != 1 && p.stride != 2) || p.up != 1) return false;
  WgradPlan pl;
  if (!wgrad_plan(p.M, p.N, p.Ho, p.Wo, p.Cout, p.Cin, p.R, p.S, p.stride, &pl)) return false;
  if (p.ldx % 8 != 0 || p.lddy % 8 != 0 || !aligned16(p.x) || !aligned16(p.dy)) return false;
  return get_encode_tiled() != nullptr;
}

size_t tc_wgrad_workspace(const WgradP& p) {
  if ((p.stride != 1 && p.stride != 2) || p.up != 1) return 0;
  WgradPlan pl;
  if (!wgrad_plan(p.M, p.N, p.Ho, p.Wo, p.Cout, p.Cin, p.R, p.S, p.stride, &pl)) return 0;
  return pl.splits > 1 ? (size_t)pl.splits * p.Cout * p.K * sizeof(float) : 0;
}

template <int BN, int ANARROW, int STR>
static int launch_wgrad_cfg(const CUtensorMap& tmDY, const CUtensorMap& tmX, const CUtensorMap& tmOut,
                            const TcWgradArgs& a, int items, cudaStream_t st) {
  using Cfg = TcWgradCfg<BN, ANARROW, STR>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(wgrad_tc_kernel<BN, ANARROW, STR>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes);
    if (e != cudaSuccess) {
      set_error("wgrad_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      return STP_E_CUDA;
    }
    attr_set = true;
  }
  launch_pdl(wgrad_tc_kernel<BN, ANARROW, STR>, dim3(items), dim3(kThreads), (size_t)Cfg::kSmemBytes, st, tmDY, tmX, tmOut, a);
  g_tc_launches.fetch_add(1, std::memory_order_relaxed);
  return check_launch("wgrad_tc");
}

int launch_tc_wgrad(const WgradP& p, float* dw, void* ws, size_t ws_bytes, cudaStream_t st) {
  WgradPlan pl;
  if (!wgrad_plan(p.M, p.N, p.Ho, p.Wo, p.Cout, p.Cin, p.R, p.S, p.stride, &pl)) {
    set_error("wgrad_tc: unsupported shape");
    return STP_E_UNSUPPORTED;
  }
  const size_t per = (size_t)p.Cout * p.K * sizeof(float);
  if (pl.splits > 1 && (ws == nullptr || ws_bytes < per * pl.splits)) {
    set_error("conv_wgrad: workspace too small (%zu < %zu)", ws_bytes, per * pl.splits);
    return STP_E_WORKSPACE;
  }
  TcWgradArgs a;
  a.out = pl.splits > 1 ? (float*)ws : dw;
  a.Cout = p.Cout; a.Cin = p.Cin; a.R = p.R; a.S = p.S; a.pad_h = p.pad_h; a.pad_w = p.pad_w;
  a.BW = pl.BW; a.BH = pl.BH; a.tilesW = pl.tilesW; a.tilesH = pl.tilesH;
  a.num_pb = pl.num_pb; a.pb_per_split = pl.pb_per_split; a.splits = pl.splits;
  a.co_tiles = pl.co_tiles; a.ci_tiles = pl.ci_tiles;
  a.a_boxes = p.Cout >= 128 ? 2 : 1;
  const int co_atom = p.Cout < 64 ? p.Cout : 64;  // channels per dY box / MN atom
  a.a_row = co_atom * 2;
  a.a_lbo = p.Cout >= 128 ? 128 * a.a_row : 0;
  a.m_valid = p.Cout >= 128 ? 128 : p.Cout;
  const int ci_atom = p.Cin < 64 ? p.Cin : 64;
  a.b_row = ci_atom * 2;
  a.xrows = p.stride == 2 ? 128 : (pl.BH + p.R - 1) * pl.BW;
  a.stride = p.stride;
  a.dbg = get_option(OPT_TC2_DEBUG);
  CUtensorMap tmDY, tmX;
  {
    uint64_t dims[4] = {(uint64_t)p.Cout, (uint64_t)p.Wo, (uint64_t)p.Ho, (uint64_t)p.N};
    uint64_t strides[3] = {(uint64_t)p.lddy * 2, (uint64_t)p.Wo * p.lddy * 2, (uint64_t)p.Ho * p.Wo * p.lddy * 2};
    uint32_t box[4] = {(uint32_t)co_atom, (uint32_t)pl.BW, (uint32_t)pl.BH, 1};
    if (!make_tmap_bf16(&tmDY, p.dy, 4, dims, strides, box, a.a_row)) return STP_E_CUDA;
  }
  {
    uint64_t dims[4] = {(uint64_t)p.Cin, (uint64_t)p.W, (uint64_t)p.H, (uint64_t)p.N};
    uint64_t strides[3] = {(uint64_t)p.ldx * 2, (uint64_t)p.W * p.ldx * 2, (uint64_t)p.H * p.W * p.ldx * 2};
    if (p.stride == 2) {
      uint32_t box[4] = {(uint32_t)ci_atom, (uint32_t)(2 * pl.BW), (uint32_t)(2 * pl.BH), 1};
      uint32_t es[4] = {1, 2, 2, 1};
      if (!make_tmap_bf16(&tmX, p.x, 4, dims, strides, box, a.b_row, es)) return STP_E_CUDA;
    } else {
      uint32_t box[4] = {(uint32_t)ci_atom, (uint32_t)pl.BW, (uint32_t)(pl.BH + p.R - 1), 1};
      if (!make_tmap_bf16(&tmX, p.x, 4, dims, strides, box, a.b_row)) return STP_E_CUDA;
    }
  }
  CUtensorMap tmOut;
  {
    // out[split][co][K] fp32; box = one staged tile [32 rows][CH columns]
    const uint32_t ch = pl.BN >= 32 ? 32 : 16;
    uint64_t dims[3] = {(uint64_t)p.K, (uint64_t)p.Cout, (uint64_t)pl.splits};
    uint64_t strides[2] = {(uint64_t)p.K * 4, (uint64_t)p.Cout * p.K * 4};
    uint32_t box[3] = {ch, 32, 1};
    if (!make_tmap_f32(&tmOut, a.out, 3, dims, strides, box, ch * 4)) return STP_E_CUDA;
  }
  const bool narrow = p.Cout <= 32;
  int rc = STP_E_UNSUPPORTED;
  if (p.stride == 2) {
#define STP_WG_CASE(bn) \
  if (pl.BN == bn) rc = narrow ? launch_wgrad_cfg<bn, 1, 1>(tmDY, tmX, tmOut, a, pl.items, st) : launch_wgrad_cfg<bn, 0, 1>(tmDY, tmX, tmOut, a, pl.items, st);
    STP_WG_CASE(64) STP_WG_CASE(32) STP_WG_CASE(16)
#undef STP_WG_CASE
  } else {
#define STP_WG_CASE(bn) \
  if (pl.BN == bn) rc = narrow ? launch_wgrad_cfg<bn, 1, 0>(tmDY, tmX, tmOut, a, pl.items, st) : launch_wgrad_cfg<bn, 0, 0>(tmDY, tmX, tmOut, a, pl.items, st);
    STP_WG_CASE(128) STP_WG_CASE(64) STP_WG_CASE(32) STP_WG_CASE(16)
#undef STP_WG_CASE
  }
  if (rc || pl.splits == 1) return rc;
  return launch_split_reduce((const float*)ws, pl.splits, (int64_t)p.Cout * p.K, dw, st);
}

}  // namespace stp
