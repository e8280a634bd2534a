// tcgen05 / TMA implicit-GEMM convolution, second generation ("halo" kernel) for RxS > 1x1, stride 1.
//
// The first-generation kernel (conv_tc.cu) issues one TMA box per filter tap, i.e. it pulls every input pixel R*S
// times through L2 -- measured L2->SM bound on B200 (about 4.5-5 TB/s) long before the tensor pipe saturates.
// Here one CTA owns a tall strip of MT sub-tiles (each BH x BW = 128 output pixels, stacked vertically) and, per
// (filter column s, 64/32/16-channel block), loads
//     ONE input box {BK ch, BW, MT*BH + R - 1 rows}  shifted horizontally by (s - pad_w)
//     R weight boxes {BK, BN}                        (taps (0..R-1, s))
// The R filter ROWS are then R shifted *views* of the same shared-memory tile: a shift by r image rows is r*BW
// swizzled 8-row atoms, so only the UMMA descriptor start address changes.  Input traffic drops from R*S to
// S*(1 + (R-1)/(MT*BH)) box loads per pixel and each weight box is reused by MT accumulators (MT*128 pixels).
// Accumulators: MT tiles of [128 x BN] fp32 in TMEM, double buffered (2*MT*BN <= 512 columns) so the epilogue of
// one strip overlaps the MMAs of the next.  Warp roles / barriers as in conv_tc.cu.
#include "conv.h"
#include "tc_common.cuh"

namespace stp {
namespace {

using namespace tc;

__device__ __forceinline__ unsigned long long gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define STP_TRACE(slot)                                                                                   \
  do {                                                                                                    \
    if (a.trace && (blockIdx.x == 0 || blockIdx.x == gridDim.x - 1))                                      \
      a.trace[(blockIdx.x == 0 ? 0 : 16) + (slot)] = gtime();                                             \
  } while (0)

// warp 0: TMA producer, warp 1: MMA issuer, warps 2..9: EIGHT epilogue warps.  A warp may only read the TMEM lane quarter
// (warp % 4), so the two warps of a quarter share its 32 pixel rows and split the work units (32-column chunks / sub-tiles)
// by parity: the epilogue -- one dependent chain per warp, exposed after the last tile of every CTA and the pacing
// resource of the HBM-bound narrow layers -- runs twice as wide as with one warp per quarter.
constexpr int kEpiThreads2 = 256;
constexpr int kThreads2 = 64 + kEpiThreads2;
constexpr int kSmemBudget2 = 227 * 1024;

struct Tc2Args {
  void* y;
  const __nv_bfloat16* res;
  const float* bias;
  int ldy, ldr, y_f32, relu;
  int Ho, Wo, Cout, Cin;
  int R, S, pad_h, pad_w;
  int wr0, wrs, ws0, wss, wS;  // weight tap remap (ConvP::w_*): tap (r, s) reads tap (wr0 + r*wrs, ws0 + s*wss) of a [..][wS] filter
  int BW, BH, log2BW;
  int tilesW, tilesH, tilesN;
  int num_tiles;             // work items: tiles (CL == 1) or cluster items = groups of CL pixel tiles x N tiles
  int numPT, nimg;           // pixel tiles per N tile, images (cluster variants: ranks beyond numPT idle on an OOB image)
  int a_bytes, b_tap_bytes;  // runtime sizes of the A box and of one weight box
  int bn_on;                 // 1: accumulate the forward BatchNorm statistics of the stored output (bf16 path)
                             // 2: this launch is a dgrad whose output is d(BatchNorm+ReLU output): tmR maps the BatchNorm INPUT x;
                             //    store g = dy*[x*scale+shift > 0] and accumulate the BatchNorm-backward sums (sum g, sum g*x)
  BnFuse bn;
  const float* bnb_coef;     // bn_on == 2: f32 [4][Cout] mean, invstd, scale, shift of that BatchNorm
  int bnb_relu;
  int ncls;                  // > 0: "head" epilogue -- only output channels [0, ncls) exist, fp32 [pixels][ncls] dense (+bias)
  unsigned long long* trace; // profiling aid (get_trace_buffer)
  int dbg;                   // timing experiments only (results invalid): 1 skip A loads, 2 skip B loads, 4 skip epilogue memory ops, 8 skip MMAs
};

constexpr int align1k(int x) { return (x + 1023) / 1024 * 1024; }

// RMAX: largest filter height/width served (3; 4 for the space-to-depth stem).  Sizes the A halo and the weight boxes.
template <int BN, int BK, int MT, int RMAX = 3>
struct Tc2Cfg {
  static constexpr int kRowBytes = BK * 2;
  static constexpr int kMaxRows = (MT * 8 + RMAX - 1) * 16 > (MT * 16 + RMAX - 1) * 8 ? (MT * 8 + RMAX - 1) * 16 : (MT * 16 + RMAX - 1) * 8;
  static constexpr int kABytes = align1k(kMaxRows * kRowBytes);   // BW=16/BH=8 or BW=8/BH=16, R<=RMAX
  static constexpr int kBTap = BN * BK * 2;
  static constexpr int kBBytes = align1k(RMAX * kBTap);
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kOutRowBytes = (BN < 64 ? BN : 64) * 2;       // one 64-channel (or narrower) slab of a pixel
  static constexpr int kOutSlabs = BN < 64 ? 1 : BN / 64;
  // Epilogue rounds: EP sub-tiles are converted into the staging buffer per barrier round.  Narrow-N kernels
  // (BN <= 32: the HBM-bound decoder tail, short main loops) stage the whole strip so the ~1000-cycle round latency
  // (store drain, two named barriers, proxy fence) is paid once per strip instead of once per 128 pixels.
  static constexpr int EP = BN <= 32 ? MT : 1;
  static constexpr int kSubBytes = align1k(128 * BN * 2);            // bf16 staging tile of one sub-tile
  static constexpr int kOutBytes = EP * kSubBytes;
  static constexpr int kTailBytes = 256 /*barriers*/ + 2 * 128 * 4 /*sStat*/ + 8 * 2 * 128 * 4 /*sRed*/ + 2 * 128 * 4 /*sCoef*/;
  static constexpr int kStagesRaw = (kSmemBudget2 - 1024 - kTailBytes - 64 - kOutBytes) / kStageBytes;
  static constexpr int kStages = kStagesRaw > 8 ? 8 : kStagesRaw;
  static constexpr int kTmemRaw = 2 * MT * BN;
  static constexpr int kTmemCols = kTmemRaw <= 32 ? 32 : kTmemRaw <= 64 ? 64 : kTmemRaw <= 128 ? 128 : kTmemRaw <= 256 ? 256 : 512;
  static constexpr int kSmemBytes = kStages * kStageBytes + kOutBytes + 1024 + kTailBytes;
  static_assert(kStages >= 2, "need at least two pipeline stages");
  static_assert(kTmemRaw <= 512, "accumulators exceed TMEM");
};

// CL > 1: thread-block clusters of CL CTAs work on CL pixel tiles of the SAME N tile in lock step; every weight box is
// fetched once per cluster -- each CTA loads 1/CL of its rows and TMA-multicasts them into all CL shared memories -- which
// divides the dominant L2->SM stream of the deep encoder layers (weights re-read per pixel tile) by CL.
template <int BN, int BK, int MT, int RMAX = 3, int CL = 1>
__global__ void __launch_bounds__(kThreads2, 1)
conv_tc2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                const __grid_constant__ CUtensorMap tmY, const __grid_constant__ CUtensorMap tmR, const Tc2Args a) {
  using Cfg = Tc2Cfg<BN, BK, MT, RMAX>;
  constexpr int NS = Cfg::kStages;
  constexpr uint16_t kMask = (uint16_t)((1u << CL) - 1u);
  const int crank = CL > 1 ? (int)cluster_ctarank() : 0;
  const int w_first = CL > 1 ? (int)cluster_id_x() : (int)blockIdx.x;
  const int w_step = CL > 1 ? (int)cluster_count_x() : (int)gridDim.x;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw & 1023u)) & 1023u);
  uint8_t* sA = smem;
  uint8_t* sB = smem + NS * Cfg::kABytes;
  uint8_t* sOut = smem + NS * Cfg::kStageBytes;  // epilogue staging tile (TMA store source / residual landing zone)
  uint64_t* full = reinterpret_cast<uint64_t*>(sOut + Cfg::kOutBytes);
  uint64_t* empty = full + NS;
  uint64_t* acc_full = empty + NS;
  uint64_t* acc_empty = acc_full + 2;
  uint64_t* res_bar = acc_empty + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(res_bar + 1);
  int* s_last = reinterpret_cast<int*>(tmem_slot + 1);
  float* sStat = reinterpret_cast<float*>(sOut + Cfg::kOutBytes + 256);  // [2][128] per-CTA sums of the current N tile
  float* sRed = sStat + 256;                                             // [8 warps][2][128] cross-warp combine
  float* sCoef = sRed + 2048;                                            // [2][128] scale, shift of the current N tile (bn_on == 2)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kcb = a.Cin / BK;
  const int num_st = a.S * kcb;  // pipeline stages' worth of work per tile
  const int TH = MT * a.BH;
  if (threadIdx.x == 0) STP_TRACE(0);

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
    for (int i = 0; i < NS; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], CL);  // the MMA warp of every CTA of the cluster releases the stage everywhere
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&acc_full[i], 1);
      mbar_init(&acc_empty[i], kEpiThreads2 / 32);
    }
    mbar_init(res_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, Cfg::kTmemCols);
    tmem_relinquish();
  }
  pdl_launch_dependents();
  tc_fence_before();
  __syncthreads();
  if (CL > 1) cluster_sync_all();  // peers' barriers are initialised before any multicast / remote arrive reaches them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (threadIdx.x == 0) STP_TRACE(1);
  pdl_wait();  // prologue above touched no global data of earlier kernels; everything below may
  if (threadIdx.x == 0) STP_TRACE(2);

  auto decode = [&](int tile, int& img, int& h0, int& w0, int& n0) {
    int tn = tile % a.tilesN;
    int t = tile / a.tilesN;
    if (CL > 1) {
      t = t * CL + crank;
      if (t >= a.numPT) {  // no pixel tile left for this rank: run the item on an out-of-range image (zero fill, clipped stores)
        img = a.nimg;
        h0 = 0;
        w0 = 0;
        n0 = tn * BN;
        return;
      }
    }
    int tw = t % a.tilesW;
    t /= a.tilesW;
    int th = t % a.tilesH;
    img = t / a.tilesH;
    h0 = th * TH;
    w0 = tw * a.BW;
    n0 = tn * BN;
  };

  if (warp == 0) {
    // TMA producer: warp-uniform loop, one elected lane issues (tc_common.cuh elect_one())
    {
      const uint32_t tx = (uint32_t)a.a_bytes + (uint32_t)a.R * (uint32_t)a.b_tap_bytes;
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = w_first; tile < a.num_tiles; tile += w_step) {
        int img, h0, w0, n0;
        decode(tile, img, h0, w0, n0);
        for (int s = 0; s < a.S; ++s) {
          for (int cb = 0; cb < kcb; ++cb) {
            mbar_wait(&empty[stage], phase ^ 1);
            if (elect_one()) {
              if (a.dbg & 3) {
                const uint32_t txd = ((a.dbg & 1) ? 0u : (uint32_t)a.a_bytes) + ((a.dbg & 2) ? 0u : (uint32_t)a.R * (uint32_t)a.b_tap_bytes);
                if (txd) mbar_expect_tx(&full[stage], txd); else mbar_arrive(&full[stage]);
              } else {
                mbar_expect_tx(&full[stage], tx);
              }
              if (!(a.dbg & 1))
              tma_load_4d(sA + stage * Cfg::kABytes, &tmA, &full[stage], cb * BK, w0 + s - a.pad_w, h0 - a.pad_h, img);
              if (CL > 1) {
                // this CTA's 1/CL slice of every tap's weight rows, multicast into all CTAs of the cluster
                constexpr int kSliceRows = BN / CL;
                for (int r = 0; r < a.R; ++r)
                  tma_load_2d_mc(sB + stage * Cfg::kBBytes + r * Cfg::kBTap + crank * (kSliceRows * BK * 2), &tmB, &full[stage],
                                 ((a.wr0 + r * a.wrs) * a.wS + a.ws0 + s * a.wss) * a.Cin + cb * BK, n0 + crank * kSliceRows, kMask);
              } else if (!(a.dbg & 2))
              for (int r = 0; r < a.R; ++r)
                tma_load_2d(sB + stage * Cfg::kBBytes + r * Cfg::kBTap, &tmB, &full[stage],
                            ((a.wr0 + r * a.wrs) * a.wS + a.ws0 + s * a.wss) * a.Cin + cb * BK, n0);
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
    // MMA issue: the whole warp runs the (warp-uniform) loop, one elected lane issues -- see tc_common.cuh elect_one()
    {
      constexpr uint32_t idesc = idesc_bf16(128, BN, 0, 0);
      const uint32_t sub16 = ((uint32_t)(a.BH * a.BW) * Cfg::kRowBytes) >> 4;  // one sub-tile's rows, in descriptor units
      const uint32_t row16 = ((uint32_t)a.BW * Cfg::kRowBytes) >> 4;           // one image row of the box
      int stage = 0;
      uint32_t phase = 0;
      int local = 0;
      for (int tile = w_first; tile < a.num_tiles; tile += w_step, ++local) {
        const int as = local & 1;
        const uint32_t aphase = (local >> 1) & 1;
        mbar_wait(&acc_empty[as], aphase ^ 1);
        tc_fence_after();
        const uint32_t d0 = tmem_base + (uint32_t)(as * MT * BN);
        for (int st = 0; st < num_st; ++st) {
          mbar_wait(&full[stage], phase);
          if (local == 0 && st == 0 && lane == 0) STP_TRACE(3);
          tc_fence_after();
          const uint64_t ad0 = desc_kmajor(smem_u32(sA + stage * Cfg::kABytes), BK * 2);
          const uint64_t bd0 = desc_kmajor(smem_u32(sB + stage * Cfg::kBBytes), BK * 2);
          if (elect_one()) {
            if (!(a.dbg & 8))
#pragma unroll
            for (int j = 0; j < MT; ++j) {
              for (int r = 0; r < a.R; ++r) {
                const uint64_t aj = ad0 + (uint64_t)(j * sub16 + r * row16);
                const uint64_t br = bd0 + (uint64_t)(r * (Cfg::kBTap >> 4));
#pragma unroll
                for (int k = 0; k < BK / 16; ++k)
                  umma_bf16(d0 + (uint32_t)(j * BN), aj + (uint64_t)(k * 2), br + (uint64_t)(k * 2), idesc, (st | r | k) != 0);
              }
            }
            if (CL > 1) umma_commit_mc(&empty[stage], kMask); else umma_commit(&empty[stage]);
          }
          __syncwarp();
          if (++stage == NS) {
            stage = 0;
            phase ^= 1;
          }
        }
        if (elect_one()) umma_commit(&acc_full[as]);
        __syncwarp();
      }
      if (lane == 0) STP_TRACE(4);
    }
  } else {
    const int q = warp & 3;               // TMEM lane quarter this warp may read
    const int m = q * 32 + lane;          // pixel row of the sub-tile == TMEM lane
    const int ew = warp - 2;              // epilogue warp 0..7
    const int half = ew >> 2;             // which of the two warps of the quarter: owns the work units of this parity
    const int t = half * 128 + m;         // epilogue thread id 0..255 (statistics mapping, per-channel ownership)
    const int hl = m >> a.log2BW, wl = m & (a.BW - 1);
    const bool leader = (warp == 2 && lane == 0);
    // 16-byte chunk swizzle of the staging tile == the TMA swizzle mode of its row width (128 / 64 / 32 B)
    constexpr int RB = Cfg::kOutRowBytes;
    const uint32_t swz = RB == 128 ? (uint32_t)(m & 7) : RB == 64 ? (uint32_t)((m >> 1) & 3) : (uint32_t)((m >> 2) & 1);
    uint32_t res_phase = 0;
    int local = 0;
    int cur_n0 = -1;  // N tile whose statistics sStat currently holds
    auto bn_flush = [&]() {
      if (cur_n0 >= 0 && t < BN) {
        atomicAdd(a.bn.acc + cur_n0 + t, (double)sStat[t]);
        atomicAdd(a.bn.acc + a.Cout + cur_n0 + t, (double)sStat[128 + t]);
      }
    };
    for (int tile = w_first; tile < a.num_tiles; tile += w_step, ++local) {
      const int as = local & 1;
      const uint32_t aphase = (local >> 1) & 1;
      int img, h0, w0, n0;
      decode(tile, img, h0, w0, n0);
      if (a.bn_on && n0 != cur_n0) {  // thread t < BN owns sStat[t], sStat[128+t]: no synchronisation needed
        bn_flush();
        cur_n0 = n0;
        if (t < BN) sStat[t] = sStat[128 + t] = 0.f;
        if (a.bn_on == 2 && t < BN) {  // read by every thread after the named barrier that opens each epilogue round
          sCoef[t] = __ldg(a.bnb_coef + 2 * a.Cout + n0 + t);
          sCoef[128 + t] = __ldg(a.bnb_coef + 3 * a.Cout + n0 + t);
        }
      }
      const int wo = w0 + wl;
      mbar_wait(&acc_full[as], aphase);
      if (leader && local == 0) STP_TRACE(5);
      tc_fence_after();
      constexpr int CH = BN >= 32 ? 32 : 16;
      if (a.ncls > 0 || a.y_f32) {
#pragma unroll 1
        for (int j = half; j < MT; j += 2) {  // the two warps of a lane quarter take alternate sub-tiles
          const int hs = h0 + j * a.BH;  // first output row of this sub-tile
          const int ho = hs + hl;
          const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)((as * MT + j) * BN);
          const bool valid = ho < a.Ho && wo < a.Wo && !(a.dbg & 4);
          const int64_t pix = ((int64_t)img * a.Ho + ho) * a.Wo + wo;
          if (a.ncls > 0) {
            // segmentation head: Cout padded to BN, only the first ncls columns are real -> dense fp32 [pixels][ncls]
            uint32_t rr[CH];
            if constexpr (CH == 32) tmem_ld32(t_addr, rr); else tmem_ld16(t_addr, rr);
            tmem_ld_wait();
            if (valid && n0 == 0) {
              float* yp = reinterpret_cast<float*>(a.y) + pix * a.ncls;
#pragma unroll
              for (int i = 0; i < 4; ++i)
                if (i < a.ncls) yp[i] = __uint_as_float(rr[i]) + (a.bias ? __ldg(a.bias + i) : 0.f);
            }
            continue;
          }
          // fp32 output (tests): direct stores
#pragma unroll 1
          for (int c0 = 0; c0 < BN; c0 += CH) {
            uint32_t rr[CH];
            if constexpr (CH == 32) tmem_ld32(t_addr + c0, rr); else tmem_ld16(t_addr + c0, rr);
            tmem_ld_wait();
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
                const __nv_bfloat16* rp = a.res + pix * a.ldr + n;
#pragma unroll
                for (int i = 0; i < CH; i += 8) {
                  float f[8];
                  unpack8(ld8(rp + i), f);
#pragma unroll
                  for (int jj = 0; jj < 8; ++jj) v[i + jj] += f[jj];
                }
              }
              if (a.relu) {
#pragma unroll
                for (int i = 0; i < CH; ++i) v[i] = fmaxf(v[i], 0.f);
              }
              float* yp = reinterpret_cast<float*>(a.y) + pix * a.ldy + n;
#pragma unroll
              for (int i = 0; i < CH; i += 4) *reinterpret_cast<float4*>(yp + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
            }
          }
        }
      } else {
        // ---- bf16 output through the swizzled staging tiles and TMA (coalesced, clipped at the image border) ----
        constexpr int EP = Cfg::EP;
#pragma unroll 1
        for (int j0 = 0; j0 < MT; j0 += EP) {
          if (h0 + j0 * a.BH >= a.Ho) break;  // rest of the strip is below the image (uniform across the CTA)
          int nj = 0;                          // sub-tiles of this round that intersect the image
#pragma unroll
          for (int e = 0; e < EP; ++e) nj += (j0 + e < MT && h0 + (j0 + e) * a.BH < a.Ho) ? 1 : 0;
          if (leader) tma_store_wait_read();   // previous round's stores no longer read the staging tiles
          named_bar_sync(1, kEpiThreads2);
          if (a.res || a.bn_on == 2) {  // residual tile, or (fused BatchNorm backward) the BatchNorm input x of the same pixels
            if (leader) {
              mbar_expect_tx(res_bar, (uint32_t)(nj * 128 * BN * 2));
              for (int e = 0; e < nj; ++e)
#pragma unroll
                for (int sl = 0; sl < Cfg::kOutSlabs; ++sl)
                  tma_load_4d(sOut + e * Cfg::kSubBytes + sl * 128 * RB, &tmR, res_bar, n0 + sl * 64, w0,
                              h0 + (j0 + e) * a.BH, img);
            }
            mbar_wait(res_bar, res_phase);
            res_phase ^= 1;
          }
          // fused BatchNorm backward, narrow tiles (BN <= 32): per-thread sums over the sub-tiles of the round, one
          // warp transpose-reduction per round; wide tiles reduce every 32-channel chunk right away
          float gacc[BN <= 32 ? BN : 1], xacc[BN <= 32 ? BN : 1];
#pragma unroll
          for (int i = 0; i < (BN <= 32 ? BN : 1); ++i) gacc[i] = xacc[i] = 0.f;
#pragma unroll 1
          for (int e = 0; e < nj; ++e) {
            const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)((as * MT + j0 + e) * BN);
            uint8_t* sub = sOut + e * Cfg::kSubBytes;
            const bool pv = h0 + (j0 + e) * a.BH + hl < a.Ho && wo < a.Wo && img < a.nimg;  // this thread's pixel is real
#pragma unroll 1
            for (int c0 = 0; c0 < BN; c0 += CH) {
              if (((e * (BN / CH) + c0 / CH) & 1) != half) continue;  // the other warp of this lane quarter owns the unit
              uint32_t rr[CH];
              if constexpr (CH == 32) tmem_ld32(t_addr + c0, rr); else tmem_ld16(t_addr + c0, rr);
              tmem_ld_wait();
              float v[CH];
#pragma unroll
              for (int i = 0; i < CH; ++i) v[i] = __uint_as_float(rr[i]);
              if (a.bias) {
#pragma unroll
                for (int i = 0; i < CH; ++i) v[i] += __ldg(a.bias + n0 + c0 + i);
              }
              uint8_t* slab = sub + (c0 >> 6) * (128 * RB) + m * RB;
              const uint32_t chunk0 = (uint32_t)((c0 & 63) >> 3);
              if (a.bn_on == 2) {
                float gx[CH];
#pragma unroll
                for (int i = 0; i < CH; i += 8) {
                  bf16x8* sp = reinterpret_cast<bf16x8*>(slab + (((chunk0 + (i >> 3)) ^ swz) << 4));
                  float xf[8];
                  unpack8(*sp, xf);
#pragma unroll
                  for (int jj = 0; jj < 8; ++jj) {
                    float val = __bfloat162float(__float2bfloat16_rn(v[i + jj]));  // the value that is stored
                    const bool keep = pv && (!a.bnb_relu || xf[jj] * sCoef[c0 + i + jj] + sCoef[128 + c0 + i + jj] > 0.f);
                    val = keep ? val : 0.f;
                    v[i + jj] = val;
                    gx[i + jj] = val * xf[jj];
                  }
                  *sp = pack8(v + i);
                }
                if constexpr (BN <= 32) {
#pragma unroll
                  for (int i = 0; i < CH; ++i) {
                    gacc[c0 + i] += v[i];
                    xacc[c0 + i] += gx[i];
                  }
                } else {
                  const float sg = warp_transpose_sum<CH>(v, lane), sx = warp_transpose_sum<CH>(gx, lane);
                  if (lane < CH) {
                    sRed[(q * 2 + 0) * 128 + c0 + lane] = sg;
                    sRed[(q * 2 + 1) * 128 + c0 + lane] = sx;
                  }
                }
                continue;
              }
#pragma unroll
              for (int i = 0; i < CH; i += 8) {
                bf16x8* sp = reinterpret_cast<bf16x8*>(slab + (((chunk0 + (i >> 3)) ^ swz) << 4));
                if (a.res) {
                  float f[8];
                  unpack8(*sp, f);
#pragma unroll
                  for (int jj = 0; jj < 8; ++jj) v[i + jj] += f[jj];
                }
                if (a.relu) {
#pragma unroll
                  for (int jj = 0; jj < 8; ++jj) v[i + jj] = fmaxf(v[i + jj], 0.f);
                }
                *sp = pack8(v + i);
              }
            }
          }
          if constexpr (BN <= 32) {
            if (a.bn_on == 2) {
#pragma unroll
              for (int c0 = 0; c0 < BN; c0 += CH) {
                float tg[CH], tx[CH];
#pragma unroll
                for (int i = 0; i < CH; ++i) {
                  tg[i] = gacc[c0 + i];
                  tx[i] = xacc[c0 + i];
                }
                const float sg = warp_transpose_sum<CH>(tg, lane), sx = warp_transpose_sum<CH>(tx, lane);
                if (lane < CH) {  // narrow tiles: every one of the 8 warps holds partial sums of its own sub-tiles
                  sRed[(ew * 2 + 0) * 128 + c0 + lane] = sg;
                  sRed[(ew * 2 + 1) * 128 + c0 + lane] = sx;
                }
              }
            }
          }
          fence_proxy_async();
          named_bar_sync(1, kEpiThreads2);
          if (leader && !(a.dbg & 4)) {
            for (int e = 0; e < nj; ++e)
#pragma unroll
              for (int sl = 0; sl < Cfg::kOutSlabs; ++sl)
                tma_store_4d(&tmY, sOut + e * Cfg::kSubBytes + sl * 128 * RB, n0 + sl * 64, w0, h0 + (j0 + e) * a.BH, img);
            tma_store_commit();
          }
          if (a.bn_on == 2) {  // cross-warp combine of the round's (sum g, sum g*x); out-of-image pixels contributed zeros
            named_bar_sync(2, kEpiThreads2);
            if (t < BN) {
              // wide tiles: one warp per (lane quarter, chunk) wrote slot q; narrow tiles: all 8 warps wrote slot ew
              constexpr int NW = BN <= 32 ? 8 : 4;
              float t1 = 0.f, t2 = 0.f;
#pragma unroll
              for (int wq = 0; wq < NW; ++wq) {
                t1 += sRed[(wq * 2 + 0) * 128 + t];
                t2 += sRed[(wq * 2 + 1) * 128 + t];
              }
              sStat[t] += t1;
              sStat[128 + t] += t2;
            }
          }
          if (a.bn_on == 1 && img < a.nimg) {
            // BatchNorm statistics of exactly the bf16 values just staged.  Thread = (channel octet o, pixel group g):
            // 16-byte conflict-free loads of the swizzled staging rows g, g+GP, ..., pixels outside the image masked;
            // groups are combined by a fixed xor-shuffle tree inside the warp, then across the 4 warps through sRed.
            constexpr int OCT = BN / 8;        // channel octets per pixel row
            constexpr int GP = kEpiThreads2 / OCT;  // pixel groups (threads per octet)
            const int o = t % OCT, g = t / OCT;
            float s1[8], s2[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) s1[i] = s2[i] = 0.f;
            for (int e = 0; e < nj; ++e) {
              const int hs = h0 + (j0 + e) * a.BH;
              const uint8_t* sub = sOut + e * Cfg::kSubBytes + (o >> 3) * (128 * RB);
#pragma unroll
              for (int i = 0; i < 128 / GP; ++i) {
                const int mm = g + i * GP;
                const int hh = hs + (mm >> a.log2BW), ww = w0 + (mm & (a.BW - 1));
                if (hh < a.Ho && ww < a.Wo) {
                  const uint32_t sw = RB == 128 ? (uint32_t)(mm & 7) : RB == 64 ? (uint32_t)((mm >> 1) & 3) : (uint32_t)((mm >> 2) & 1);
                  float f[8];
                  unpack8(*reinterpret_cast<const bf16x8*>(sub + mm * RB + ((((uint32_t)o & 7u) ^ sw) << 4)), f);
#pragma unroll
                  for (int k = 0; k < 8; ++k) {
                    s1[k] += f[k];
                    s2[k] += f[k] * f[k];
                  }
                }
              }
            }
#pragma unroll
            for (int off = 16; off >= OCT; off >>= 1) {
#pragma unroll
              for (int k = 0; k < 8; ++k) {
                s1[k] += __shfl_xor_sync(0xffffffffu, s1[k], off);
                s2[k] += __shfl_xor_sync(0xffffffffu, s2[k], off);
              }
            }
            if (lane < OCT) {  // lane == o for OCT <= 32
#pragma unroll
              for (int k = 0; k < 8; ++k) {
                sRed[(ew * 2 + 0) * 128 + lane * 8 + k] = s1[k];
                sRed[(ew * 2 + 1) * 128 + lane * 8 + k] = s2[k];
              }
            }
            named_bar_sync(2, kEpiThreads2);
            if (t < BN) {
              float t1 = 0.f, t2 = 0.f;
#pragma unroll
              for (int wq = 0; wq < 8; ++wq) {
                t1 += sRed[(wq * 2 + 0) * 128 + t];
                t2 += sRed[(wq * 2 + 1) * 128 + t];
              }
              sStat[t] += t1;
              sStat[128 + t] += t2;
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[as]);
    }
    if (leader) STP_TRACE(6);
    if (leader) tma_store_wait_all();
    if (leader) STP_TRACE(7);
    if (a.bn_on) {
      bn_flush();
      __threadfence();
      named_bar_sync(1, kEpiThreads2);
      if (leader) *s_last = (atomicAdd(a.bn.fin.sync, 1u) == gridDim.x - 1) ? 1 : 0;
      named_bar_sync(1, kEpiThreads2);
      if (*s_last) {  // every other CTA's sums have landed: finalise all channels, return the accumulators to zero
        __threadfence();
        for (int c = t; c < a.Cout; c += kEpiThreads2) {
          const double s1 = __ldcg(a.bn.acc + c), s2 = __ldcg(a.bn.acc + a.Cout + c);
          if (a.bn_on == 2) {  // sum g*xhat = invstd * (sum g*x - mean * sum g)
            const double mean = a.bn.fin.coef[c], invstd = a.bn.fin.coef[a.Cout + c];
            fin_backward(a.bn.fin, a.Cout, c, s1, invstd * (s2 - mean * s1));
          } else {
            fin_forward(a.bn.fin, a.Cout, c, s1, s2);
          }
          a.bn.acc[c] = 0.0;
          a.bn.acc[a.Cout + c] = 0.0;
        }
        if (leader) *a.bn.fin.sync = 0u;
      }
    }
  }

  if (warp == 2 && lane == 0) STP_TRACE(8);
  tc_fence_before();
  __syncthreads();
  if (CL > 1) cluster_sync_all();  // no peer still multicasts into / arrives on this CTA's shared memory
  if (threadIdx.x == 0) STP_TRACE(9);
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

struct Tc2Plan {
  int BN, BK, MT, CL;
};

// Cycle estimate of one launch for a candidate tiling (B200: 148 SMs, ~40 B/clk/SM from L2, tcgen05 M=128 floor of
// BN/2 cycles per K=16 MMA, 128 B/clk of shared-memory operand bandwidth).  Used only to RANK candidates: it captures
// the three effects measured with ncu -- wave quantisation of the persistent grid, L2->SM bound stages when a weight
// box is reused by too few pixels, and the per-strip epilogue latency.
static double tc2_cost(const ConvP& p, int bn, int bk, int mt, int cl = 1) {
  const int bw = p.Wo >= 16 ? 16 : 8, bh = 128 / bw;
  const int th = mt * bh;
  const int64_t pt = (int64_t)p.N * ((p.Ho + th - 1) / th) * ((p.Wo + bw - 1) / bw);
  const double items = (double)((pt + cl - 1) / cl) * (p.Cout / bn);  // cluster work items (cl pixel tiles each)
  const int slots = kNumSMs / cl;
  const double waves = (double)(int64_t)((items + slots - 1) / slots);
  const int num_st = p.S * (p.Cin / bk);
  // weight boxes are fetched once per cluster (multicast): 1/cl of them per CTA
  const double bytes = (double)(th + p.R - 1) * bw * bk * 2 + (double)p.R * bn * bk * 2 / cl;
  const double load = bytes / 40.0;
  const double mma_each = bn / 2.0 > 32.0 + bn / 4.0 ? bn / 2.0 : 32.0 + bn / 4.0;
  const double mma = (double)mt * p.R * (bk / 16) * mma_each;
  const double stage = (load > mma ? load : mma) + 150.0;
  const double epi = 600.0 + mt * (bn >= 64 ? 500.0 : 350.0);
  const double main = num_st * stage;
  const double tile = main > epi ? main : epi;
  return waves * tile + epi + 4000.0;
}

bool tc2_plan(const ConvP& p, Tc2Plan* pl) {
  if (p.stride != 1 || p.up != 1) return false;
  // 1x1 stride-1 convolutions run here as plain GEMMs over pixel strips (no halo, one weight box per stage): the 8-warp TMA-store
  // epilogue and the MT-stacked accumulators make this kernel ~2x faster than the streaming mma.sync GEMM (gemm1x1.cu) and 2-4x
  // faster than the first-generation kernel on the ResNet-50 bottleneck shapes (profiles/r2_s10_g1_bench.txt).  tc2_1x1 = 1: off.
  if (p.R < 2 && p.S < 2 && (get_option(OPT_TC2_1X1) == 1 || (get_option(OPT_TC2_1X1) == 2 && p.Cin % 64 != 0))) return false;
  int bk = (p.Cin % 64 == 0) ? 64 : (p.Cin == 32 ? 32 : (p.Cin == 16 ? 16 : 0));
  if (!bk) return false;
  if (bk == 64 && get_option(OPT_TC2_BK) == 32) bk = 32;  // experiment: smaller stages -> deeper pipeline
  const bool r4 = p.R == 4 || p.S == 4;  // 4x4: only the stem shape (Cin 32 = 4 sub-pixels x 8 channels, Cout 64) is instantiated
  if (p.R > 4 || p.S > 4 || (r4 && !(bk == 32 && p.Cout % 64 == 0))) return false;
  const int force = get_option(OPT_TC2_FORCE_MT);
  const int clopt = get_option(OPT_TC2_CLUSTER);
  double best = 0.0;
  int best_bn = 0, best_mt = 0, best_cl = 1;
  for (int bn : {128, 64, 32, 16}) {
    if (p.Cout % bn != 0) continue;
    if (r4 && bn != 64) continue;
    if (best_bn != 0 && bn < 64) break;  // narrow N tiles only when Cout demands them
    // strip-height caps from the measured sweep (scripts/mt_sweep.py, profiles/r1_s38_mt_sweep.log): 64->64 @128^2 29.7 us at
    // MT = 2 against 33.8 at MT = 4 (192->64: 68.6 vs 74.8); 32->32 @256^2 44.0 at MT = 4 against 48.2 at MT = 8
    const int mt_max = bn == 128 ? 2 : (bn == 64 && bk == 64) ? 2 : (bn == 64 || bk == 64) ? 4 : (bn == 32 && bk == 32) ? 4 : 8;
    for (int mt = 1; mt <= mt_max; mt *= 2) {
      if (force > 0 && mt != (force < mt_max ? force : mt_max)) continue;
      for (int cl : {1, 2, 4}) {
        // cluster (weight multicast) variants exist for the 128 x 64 tiles, bf16 TMA-store path.  Measured (profiles/
        // README.md, s19): with only two pipeline stages in flight these layers are latency- rather than byte-bound, so
        // multicast pays only where weights dominate outright (Cin, Cout >= 512: -17 %); elsewhere it is neutral and
        // the automatic choice leaves it off.
        if (cl > 1 && !(bn == 128 && bk == 64 && !r4 && !p.y_f32 && p.ncls == 0 && clopt != 1)) continue;
        if (cl > 1 && clopt == 0) continue;  // automatic choice: off (neutral within noise in the full step, s19/s21)
        if (clopt >= 2 && bn == 128 && bk == 64 && !r4 && !p.y_f32 && p.ncls == 0 && cl != clopt) continue;
        const double c = tc2_cost(p, bn, bk, mt, cl);
        if (best_bn == 0 || c < best) {
          best = c;
          best_bn = bn;
          best_mt = mt;
          best_cl = cl;
        }
      }
    }
  }
  if (!best_bn) return false;
  pl->BN = best_bn; pl->BK = bk; pl->MT = best_mt; pl->CL = best_cl;
  return true;
}

template <int BN, int BK, int MT, int RMAX = 3, int CL = 1>
int launch2(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmY, const CUtensorMap& tmR, const Tc2Args& a,
            cudaStream_t st) {
  using Cfg = Tc2Cfg<BN, BK, MT, RMAX>;
  auto kernel = conv_tc2_kernel<BN, BK, MT, RMAX, CL>;
  static bool attr_set = false;
  static int max_clusters = kNumSMs / CL;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes);
    if (e != cudaSuccess) {
      set_error("conv_tc2: cudaFuncSetAttribute(%d B): %s", Cfg::kSmemBytes, cudaGetErrorString(e));
      return STP_E_CUDA;
    }
    if (CL > 1) {  // how many clusters can be co-resident (GPC granularity): the persistent grid must not exceed it
      cudaLaunchConfig_t q = {};
      q.gridDim = dim3(kNumSMs / CL * CL);
      q.blockDim = dim3(kThreads2);
      q.dynamicSmemBytes = Cfg::kSmemBytes;
      cudaLaunchAttribute qa[1];
      qa[0].id = cudaLaunchAttributeClusterDimension;
      qa[0].val.clusterDim.x = CL; qa[0].val.clusterDim.y = 1; qa[0].val.clusterDim.z = 1;
      q.attrs = qa; q.numAttrs = 1;
      int n = 0;
      if (cudaOccupancyMaxActiveClusters(&n, kernel, &q) == cudaSuccess && n > 0 && n < max_clusters) max_clusters = n;
      (void)cudaGetLastError();
    }
    attr_set = true;
  }
  cudaLaunchConfig_t cfg = {};
  if (CL > 1) {
    const int clusters = a.num_tiles < max_clusters ? a.num_tiles : max_clusters;
    cfg.gridDim = dim3(clusters * CL);
  } else {
    cfg.gridDim = dim3(a.num_tiles < kNumSMs ? a.num_tiles : kNumSMs);
  }
  cfg.blockDim = dim3(kThreads2);
  cfg.dynamicSmemBytes = Cfg::kSmemBytes;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = g_pdl_enabled.load(std::memory_order_relaxed) ? 1 : 0;
  attr[1].id = cudaLaunchAttributeClusterDimension;
  attr[1].val.clusterDim.x = CL; attr[1].val.clusterDim.y = 1; attr[1].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = CL > 1 ? 2 : 1;
  cudaLaunchKernelEx(&cfg, kernel, tmA, tmB, tmY, tmR, a);
  g_tc_launches.fetch_add(1, std::memory_order_relaxed);
  return check_launch("conv_tc2");
}

}  // namespace

bool tc2_conv_supported(const ConvP& p) {
  Tc2Plan pl;
  if (!tc2_plan(p, &pl)) return false;
  if (p.Wo < 8 || p.Ho < 1) return false;
  if (p.ldx % 8 != 0 || !aligned16(p.x) || !aligned16(p.w)) return false;
  if (p.ncls > 0) {
    if (p.ncls > 4 || !p.y_f32 || p.res || p.relu) return false;
  } else {
    if (p.y_f32 ? (p.ldy % 4 != 0) : (p.ldy % 8 != 0)) return false;
    if (!aligned16(p.y)) return false;
  }
  if (p.res && (p.ldr % 8 != 0 || !aligned16(p.res))) return false;
  if (p.bn && p.bn->fin.mode == 2) {  // fused BatchNorm backward: bf16 output, no residual, x tensor TMA-mappable
    if (p.res || p.y_f32 || p.ncls > 0 || !p.bnb_x || !p.bnb_coef || p.bnb_ldx % 8 != 0 || !aligned16(p.bnb_x)) return false;
  }
  return get_encode_tiled() != nullptr;
}

int launch_tc2_conv(const ConvP& p, cudaStream_t st) {
  Tc2Plan pl;
  if (!tc2_plan(p, &pl)) {
    set_error("conv_tc2: unsupported");
    return STP_E_UNSUPPORTED;
  }
  Tc2Args a;
  a.y = p.y; a.res = p.res; a.bias = p.bias; a.ldy = p.ldy; a.ldr = p.ldr; a.y_f32 = p.y_f32; a.relu = p.relu;
  a.Ho = p.Ho; a.Wo = p.Wo; a.Cout = p.Cout; a.Cin = p.Cin; a.R = p.R; a.S = p.S; a.pad_h = p.pad_h; a.pad_w = p.pad_w;
  a.wr0 = p.w_r0; a.wrs = p.w_rs; a.ws0 = p.w_s0; a.wss = p.w_ss; a.wS = p.w_S ? p.w_S : p.S;
  a.BW = p.Wo >= 16 ? 16 : 8;
  a.BH = 128 / a.BW;
  a.log2BW = a.BW == 16 ? 4 : 3;
  const int TH = pl.MT * a.BH;
  a.tilesW = (p.Wo + a.BW - 1) / a.BW;
  a.tilesH = (p.Ho + TH - 1) / TH;
  a.tilesN = p.Cout / pl.BN;
  int64_t nt = (int64_t)p.N * a.tilesH * a.tilesW * a.tilesN;
  if (nt > 0x7fffffff) {
    set_error("conv_tc2: too many tiles");
    return STP_E_UNSUPPORTED;
  }
  a.num_tiles = (int)nt;
  a.nimg = p.N;
  a.numPT = p.N * a.tilesH * a.tilesW;
  if (pl.CL > 1) a.num_tiles = ((a.numPT + pl.CL - 1) / pl.CL) * a.tilesN;
  const int box_rows = TH + p.R - 1;
  a.a_bytes = box_rows * a.BW * pl.BK * 2;
  a.b_tap_bytes = pl.BN * pl.BK * 2;
  a.dbg = get_option(OPT_TC2_DEBUG);
  a.trace = get_trace_buffer();
  a.ncls = p.ncls;
  a.bn_on = (p.bn != nullptr && !p.y_f32 && p.ncls == 0) ? (p.bn->fin.mode == 2 ? 2 : 1) : 0;
  if (a.bn_on) a.bn = *p.bn; else a.bn = BnFuse{};
  a.bnb_coef = p.bnb_coef;
  a.bnb_relu = p.bnb_relu;
  CUtensorMap tmA, tmB;
  {
    uint64_t dims[4] = {(uint64_t)p.Cin, (uint64_t)p.W, (uint64_t)p.H, (uint64_t)p.N};
    uint64_t strides[3] = {(uint64_t)p.ldx * 2, (uint64_t)p.W * p.ldx * 2, (uint64_t)p.H * p.W * p.ldx * 2};
    uint32_t box[4] = {(uint32_t)pl.BK, (uint32_t)a.BW, (uint32_t)box_rows, 1};
    if (!make_tmap_bf16(&tmA, p.x, 4, dims, strides, box, pl.BK * 2)) return STP_E_CUDA;
  }
  {
    const uint64_t wk = (uint64_t)(p.w_K ? p.w_K : p.K);
    uint64_t dims[2] = {wk, (uint64_t)p.Cout};
    uint64_t strides[1] = {wk * 2};
    uint32_t box[2] = {(uint32_t)pl.BK, (uint32_t)(pl.BN / pl.CL)};  // cluster variants: each CTA loads 1/CL of the rows
    if (!make_tmap_bf16(&tmB, p.w, 2, dims, strides, box, pl.BK * 2)) return STP_E_CUDA;
  }
  CUtensorMap tmY = tmA, tmR = tmA;  // placeholders when unused (fp32 output / no residual)
  const uint32_t oc = pl.BN < 64 ? pl.BN : 64;
  if (!p.y_f32) {
    uint64_t dims[4] = {(uint64_t)p.Cout, (uint64_t)p.Wo, (uint64_t)p.Ho, (uint64_t)p.N};
    uint64_t strides[3] = {(uint64_t)p.ldy * 2, (uint64_t)p.Wo * p.ldy * 2, (uint64_t)p.Ho * p.Wo * p.ldy * 2};
    if (p.y_sw) { strides[0] = (uint64_t)p.y_sw * 2; strides[1] = (uint64_t)p.y_sh * 2; strides[2] = (uint64_t)p.y_sn * 2; }
    uint32_t box[4] = {oc, (uint32_t)a.BW, (uint32_t)a.BH, 1};
    if (!make_tmap_bf16(&tmY, p.y, 4, dims, strides, box, oc * 2)) return STP_E_CUDA;
    if (p.res) {
      uint64_t rstrides[3] = {(uint64_t)p.ldr * 2, (uint64_t)p.Wo * p.ldr * 2, (uint64_t)p.Ho * p.Wo * p.ldr * 2};
      if (p.r_sw) { rstrides[0] = (uint64_t)p.r_sw * 2; rstrides[1] = (uint64_t)p.r_sh * 2; rstrides[2] = (uint64_t)p.r_sn * 2; }
      if (!make_tmap_bf16(&tmR, p.res, 4, dims, rstrides, box, oc * 2)) return STP_E_CUDA;
    } else if (a.bn_on == 2) {  // the BatchNorm input of the same pixels travels through the residual slot
      uint64_t rstrides[3] = {(uint64_t)p.bnb_ldx * 2, (uint64_t)p.Wo * p.bnb_ldx * 2, (uint64_t)p.Ho * p.Wo * p.bnb_ldx * 2};
      if (!make_tmap_bf16(&tmR, p.bnb_x, 4, dims, rstrides, box, oc * 2)) return STP_E_CUDA;
    }
  }
  if (p.R == 4 || p.S == 4) {
    if (pl.BN == 64 && pl.BK == 32 && pl.MT == 4) return launch2<64, 32, 4, 4>(tmA, tmB, tmY, tmR, a, st);
    if (pl.BN == 64 && pl.BK == 32 && pl.MT == 2) return launch2<64, 32, 2, 4>(tmA, tmB, tmY, tmR, a, st);
    if (pl.BN == 64 && pl.BK == 32 && pl.MT == 1) return launch2<64, 32, 1, 4>(tmA, tmB, tmY, tmR, a, st);
  }
  if (pl.CL == 2) {
    if (pl.BN == 128 && pl.BK == 64 && pl.MT == 2) return launch2<128, 64, 2, 3, 2>(tmA, tmB, tmY, tmR, a, st);
    if (pl.BN == 128 && pl.BK == 64 && pl.MT == 1) return launch2<128, 64, 1, 3, 2>(tmA, tmB, tmY, tmR, a, st);
  } else if (pl.CL == 4) {
    if (pl.BN == 128 && pl.BK == 64 && pl.MT == 2) return launch2<128, 64, 2, 3, 4>(tmA, tmB, tmY, tmR, a, st);
    if (pl.BN == 128 && pl.BK == 64 && pl.MT == 1) return launch2<128, 64, 1, 3, 4>(tmA, tmB, tmY, tmR, a, st);
  }
#define STP_TC2_CASE(bn, bk, mt) \
  if (pl.BN == bn && pl.BK == bk && pl.MT == mt) return launch2<bn, bk, mt>(tmA, tmB, tmY, tmR, a, st);
  STP_TC2_CASE(128, 64, 2) STP_TC2_CASE(128, 64, 1)
  STP_TC2_CASE(64, 64, 4) STP_TC2_CASE(64, 64, 2) STP_TC2_CASE(64, 64, 1)
  STP_TC2_CASE(32, 64, 4) STP_TC2_CASE(32, 64, 2) STP_TC2_CASE(32, 64, 1)
  STP_TC2_CASE(16, 64, 4) STP_TC2_CASE(16, 64, 2) STP_TC2_CASE(16, 64, 1)
  STP_TC2_CASE(128, 32, 2) STP_TC2_CASE(128, 32, 1)
  STP_TC2_CASE(64, 32, 4) STP_TC2_CASE(64, 32, 2) STP_TC2_CASE(64, 32, 1)
  STP_TC2_CASE(32, 32, 8) STP_TC2_CASE(32, 32, 4) STP_TC2_CASE(32, 32, 2) STP_TC2_CASE(32, 32, 1)
  STP_TC2_CASE(16, 32, 8) STP_TC2_CASE(16, 32, 4) STP_TC2_CASE(16, 32, 2) STP_TC2_CASE(16, 32, 1)
  STP_TC2_CASE(128, 16, 2) STP_TC2_CASE(128, 16, 1)
  STP_TC2_CASE(64, 16, 4) STP_TC2_CASE(64, 16, 2) STP_TC2_CASE(64, 16, 1)
  STP_TC2_CASE(32, 16, 8) STP_TC2_CASE(32, 16, 4) STP_TC2_CASE(32, 16, 2) STP_TC2_CASE(32, 16, 1)
  STP_TC2_CASE(16, 16, 8) STP_TC2_CASE(16, 16, 4) STP_TC2_CASE(16, 16, 2) STP_TC2_CASE(16, 16, 1)
#undef STP_TC2_CASE
  set_error("conv_tc2: no specialisation BN=%d BK=%d MT=%d", pl.BN, pl.BK, pl.MT);
  return STP_E_UNSUPPORTED;
}

// ---- zero-insertion (up == 2) problems as four stride-1 launches --------------------------------------------------------
// y[i] = sum_r' xup[i - pad + r'] w[r'],  xup[2o] = x[o]: for the output parity class i = 2a + ph only the taps with
// (ph - pad + r') even contribute, reading x[a + (ph - pad + r')/2] -- consecutive input rows for consecutive kept taps, i.e. a
// stride-1 convolution with every second filter tap, written to every second output pixel (a strided TMA-store view).  The
// first-generation kernel serves all four classes in one launch with per-thread strided stores (conv_tc.cu); here each class
// gets the halo kernel's operand reuse and TMA-store epilogue: dgrad of 3x3/2 64 -> 128 @128^2 bs16 107 us -> see profiles/.
static bool tc2_up2_class(const ConvP& p, int ph, int pw, ConvP* q) {
  *q = p;
  const int r0 = ((p.pad_h - ph) % 2 + 2) % 2, s0 = ((p.pad_w - pw) % 2 + 2) % 2;
  const int rc = r0 < p.R ? (p.R - r0 + 1) / 2 : 0, sc = s0 < p.S ? (p.S - s0 + 1) / 2 : 0;
  q->Ho = (p.Ho - ph + 1) / 2;
  q->Wo = (p.Wo - pw + 1) / 2;
  if (rc == 0 || sc == 0 || q->Ho <= 0 || q->Wo <= 0) return false;
  q->up = 1; q->stride = 1;
  q->R = rc; q->S = sc;
  q->pad_h = -((ph - p.pad_h + r0) / 2);  // numerator even: exact for negatives
  q->pad_w = -((pw - p.pad_w + s0) / 2);
  q->w_r0 = r0; q->w_rs = 2; q->w_s0 = s0; q->w_ss = 2; q->w_S = p.S; q->w_K = p.K;
  q->K = rc * sc * p.Cin;
  q->M = (int64_t)p.N * q->Ho * q->Wo;
  const int64_t yo = ((int64_t)ph * p.Wo + pw) * p.ldy;
  q->y = (__nv_bfloat16*)p.y + yo;
  q->y_sw = (int64_t)2 * p.ldy; q->y_sh = (int64_t)2 * p.Wo * p.ldy; q->y_sn = (int64_t)p.Ho * p.Wo * p.ldy;
  if (p.res) {
    q->res = p.res + ((int64_t)ph * p.Wo + pw) * p.ldr;
    q->r_sw = (int64_t)2 * p.ldr; q->r_sh = (int64_t)2 * p.Wo * p.ldr; q->r_sn = (int64_t)p.Ho * p.Wo * p.ldr;
  }
  return true;
}

bool tc2_up2_supported(const ConvP& p) {
  if (p.up != 2 || p.stride != 1 || p.y_f32 || p.ncls != 0 || p.bn || p.R > 4 || p.S > 4 || p.R < 2 || p.S < 2) return false;
  for (int ph = 0; ph < 2; ++ph)
    for (int pw = 0; pw < 2; ++pw) {
      ConvP q;
      if (!tc2_up2_class(p, ph, pw, &q)) return false;  // a class no tap reaches (zeros + residual): first-generation kernel
      if (!tc2_conv_supported(q)) return false;
    }
  return true;
}

int launch_tc2_up2(const ConvP& p, cudaStream_t st) {
  for (int ph = 0; ph < 2; ++ph)
    for (int pw = 0; pw < 2; ++pw) {
      ConvP q;
      if (!tc2_up2_class(p, ph, pw, &q)) {
        set_error("conv_tc2 (up 2): empty class");
        return STP_E_UNSUPPORTED;
      }
      const int rc = launch_tc2_conv(q, st);
      if (rc) return rc;
    }
  return STP_OK;
}

}  // namespace stp
