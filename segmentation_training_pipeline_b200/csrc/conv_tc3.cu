// tcgen05 / TMA implicit-GEMM convolution, third generation ("pair" kernel): 3x3 stride-1 layers with Cin % 64 == 0 and
// Cout % 128 == 0 -- the FLOP-heavy encoder body / wide decoder layers of the U-Net.
//
// What bounded conv_tc2 on these layers (profiles/README.md, s19 trace): every CTA pulls the full BN x K weight panel
// through L2->SM (~41 B/clk/SM cap) for each 128/256-pixel strip, and its 85 KB (A box + 3 weight taps) stages leave
// room for only two loads in flight, so the main loop ran latency- and L2-bound at 38-46 % tensor-pipe activity.
// Here:
//   * two CTAs of a cluster (one TPC) form a tcgen05 CTA PAIR: one `tcgen05.mma.cta_group::2` of M = 256 covers the
//     128-pixel sub-tile of each CTA against an N = BN weight tile of which each CTA holds only HALF the rows
//     (B is split across the pair by the hardware) -> weight bytes per CTA per FLOP halve, with no multicast;
//   * the A halo boxes and the weight taps travel through SEPARATE rings (NA boxes of 20-36 KB, NB taps of 8-16 KB), so
//     a weight slot is released after MT*4 MMAs and 6-12 loads stay in flight instead of 2;
//   * only the leader CTA issues MMAs; both CTAs' TMA loads complete transaction bytes on the LEADER's mbarriers
//     (cp.async.bulk.tensor ... .cta_group::2), ring slots / accumulators are released in both CTAs by multicast
//     tcgen05.commit, and the peer's epilogue warps release the accumulator with a remote mbarrier arrive.
// The halo scheme (filter rows = shifted views of one input box), the TMEM double buffering, the swizzled staging /
// TMA-store epilogue with residual and the fused BatchNorm statistics are those of conv_tc2.cu.
#include "conv.h"
#include "tc_common.cuh"

namespace stp {
namespace {

using namespace tc;

__device__ __forceinline__ unsigned long long gtime3() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// profiling aid (stp_set_trace_buffer): %globaltimer stamps of the phases of the first and the last CTA
#define STP_TRACE3(slot)                                                                                  \
  do {                                                                                                    \
    if (a.trace && (blockIdx.x == 0 || blockIdx.x == gridDim.x - 1))                                      \
      a.trace[(blockIdx.x == 0 ? 0 : 16) + (slot)] = gtime3();                                            \
  } while (0)

constexpr int kThreads3 = 192;
constexpr int kSmemBudget3 = 227 * 1024;
constexpr int kBK3 = 64;

struct Tc3Args {
  const __nv_bfloat16* res;
  const float* bias;
  int relu;
  int Ho, Wo, Cout, Cin;
  int R, S, pad_h, pad_w;
  int BW, BH, log2BW;
  int BWX;         // width (pixels) of the A halo box in shared memory: BW (one box per filter column) or BW + S - 1 (halo mode)
  int halo;        // 1: ONE haloed box per channel block serves all R*S taps (filter columns = 1-pixel shifted descriptor views)
  int tilesW, tilesH, tilesN;
  int num_items;   // pair work items = ceil(numPT / 2) pixel-tile pairs x N tiles
  int numPT, nimg; // pixel tiles per N tile; images (a rank without a pixel tile runs on image index nimg: zero fill, clipped stores)
  int a_bytes;     // runtime size of one A box
  int bn_on;
  BnFuse bn;
  unsigned long long* trace;
  int dbg;  // timing experiments only (results invalid): 1 skip A loads, 2 skip B loads, 8 skip MMAs
};

constexpr int align1k3(int x) { return (x + 1023) / 1024 * 1024; }

template <int BN, int MT>
struct Tc3Cfg {
  static constexpr int kRowBytes = kBK3 * 2;  // 128-byte swizzled rows
  // largest A box in pixels: 16 x 8-row strips with a 2-row halo (one box per filter column), 8-wide x 16-row strips with a
  // 2-row halo, or -- halo mode -- 8-wide x 16-row strips with a 2-row AND 2-column halo
  static constexpr int kMaxRows = (MT * 16 + 2) * 10 > (MT * 8 + 2) * 16 ? (MT * 16 + 2) * 10 : (MT * 8 + 2) * 16;
  static constexpr int kABytes = align1k3(kMaxRows * kRowBytes);
  static constexpr int kBTap = (BN / 2) * kRowBytes;  // this CTA's half of one tap: BN/2 weight rows x 64 channels
  static constexpr int kOutSlabs = BN / 64;
  static constexpr int kOutBytes = 128 * BN * 2;      // bf16 staging tile of one 128-pixel sub-tile
  static constexpr int kBarBytes = 512;
  static constexpr int kTailBytes = kBarBytes + 2 * BN * 4 /*sStat*/ + 8 * BN * 4 /*sRed*/;
  static constexpr int kRing = kSmemBudget3 - 1024 - kTailBytes - kOutBytes;
  static constexpr int NA = MT == 2 ? 3 : (BN == 256 ? 2 : 4);
  static constexpr int kNBRaw = (kRing - NA * kABytes) / kBTap;
  static constexpr int NB = kNBRaw > 12 ? 12 : kNBRaw;
  static constexpr int kTmemCols = 2 * MT * BN <= 256 ? 256 : 512;
  static constexpr int kSmemBytes = NA * kABytes + NB * kBTap + kOutBytes + 1024 + kTailBytes;
  static_assert(NB >= 4, "weight ring too small");
  static_assert(2 * MT * BN <= 512, "accumulators exceed TMEM");
  static_assert((2 * NA + 2 * NB + 5) * 8 + 16 <= kBarBytes, "barrier area too small");
};

template <int BN, int MT>
__global__ void __launch_bounds__(kThreads3, 1)
conv_tc3_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                const __grid_constant__ CUtensorMap tmY, const __grid_constant__ CUtensorMap tmR, const Tc3Args a) {
  using Cfg = Tc3Cfg<BN, MT>;
  constexpr int NA = Cfg::NA, NB = Cfg::NB;
  const uint32_t crank = cluster_ctarank();
  const bool leader = crank == 0;
  const int w_first = (int)cluster_id_x(), w_step = (int)cluster_count_x();
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw & 1023u)) & 1023u);
  uint8_t* sA = smem;
  uint8_t* sB = smem + NA * Cfg::kABytes;
  uint8_t* sOut = sB + NB * Cfg::kBTap;
  uint64_t* fullA = reinterpret_cast<uint64_t*>(sOut + Cfg::kOutBytes);
  uint64_t* emptyA = fullA + NA;
  uint64_t* fullB = emptyA + NA;
  uint64_t* emptyB = fullB + NB;
  uint64_t* acc_full = emptyB + NB;
  uint64_t* acc_empty = acc_full + 2;
  uint64_t* res_bar = acc_empty + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(res_bar + 1);
  int* s_last = reinterpret_cast<int*>(tmem_slot + 1);
  float* sStat = reinterpret_cast<float*>(sOut + Cfg::kOutBytes + Cfg::kBarBytes);  // [2][BN]
  float* sRed = sStat + 2 * BN;                                                      // [4 warps][2][BN]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kcb = a.Cin / kBK3;
  const int TH = MT * a.BH;
  if (threadIdx.x == 0) STP_TRACE3(0);

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
    for (int i = 0; i < NA; ++i) {
      mbar_init(&fullA[i], 1);   // leader: its producer's arrive.expect_tx (bytes of BOTH CTAs)
      mbar_init(&emptyA[i], 1);  // multicast commit of the leader's MMA thread
    }
    for (int i = 0; i < NB; ++i) {
      mbar_init(&fullB[i], 1);
      mbar_init(&emptyB[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&acc_full[i], 1);
      mbar_init(&acc_empty[i], 8);  // leader: 4 epilogue warps of each CTA of the pair
    }
    mbar_init(res_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc_pair(tmem_slot, Cfg::kTmemCols);
    tmem_relinquish_pair();
  }
  pdl_launch_dependents();
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // the peer's barriers are initialised before any remote complete_tx / arrive reaches them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (threadIdx.x == 0) STP_TRACE3(1);
  pdl_wait();  // nothing above touched global data of earlier kernels
  if (threadIdx.x == 0) STP_TRACE3(2);

  auto decode = [&](int item, int& img, int& h0, int& w0, int& n0) {
    const int tn = item % a.tilesN;
    int t = (item / a.tilesN) * 2 + (int)crank;
    n0 = tn * BN;
    if (t >= a.numPT) {
      img = a.nimg;
      h0 = 0;
      w0 = 0;
      return;
    }
    const int tw = t % a.tilesW;
    t /= a.tilesW;
    const int th = t % a.tilesH;
    img = t / a.tilesH;
    h0 = th * TH;
    w0 = tw * a.BW;
  };

  if (warp == 0) {
    {  // TMA producer: warp-uniform loop, one elected lane issues (tc_common.cuh elect_one())
      int sa = 0, sb = 0;
      uint32_t pha = 0, phb = 0;
      for (int item = w_first; item < a.num_items; item += w_step) {
        int img, h0, w0, n0;
        decode(item, img, h0, w0, n0);
        // outer step = one A box: (filter column s, channel block cb) -- or, halo mode, one channel block whose haloed box
        // serves all S columns; inner step = one weight tap (r [, s])
        const int n_outer = a.halo ? kcb : a.S * kcb;
        const int n_inner = a.halo ? a.R * a.S : a.R;
        for (int o = 0; o < n_outer; ++o) {
          const int s = a.halo ? 0 : o / kcb;
          const int cb = a.halo ? o : o % kcb;
          {
            mbar_wait(&emptyA[sa], pha ^ 1);
            if (elect_one()) {
            if (a.dbg & 1) {
              if (leader) mbar_arrive(&fullA[sa]);
            } else {
            if (leader) mbar_expect_tx(&fullA[sa], 2u * (uint32_t)a.a_bytes);
            tma_load_4d_pair(sA + sa * Cfg::kABytes, &tmA, mapa_u32(smem_u32(&fullA[sa]), 0), cb * kBK3, w0 + s - a.pad_w,
                             h0 - a.pad_h, img);
            }
            }
            __syncwarp();
            if (++sa == NA) {
              sa = 0;
              pha ^= 1;
            }
            for (int t = 0; t < n_inner; ++t) {
              const int tap = a.halo ? t : t * a.S + s;   // r * S + s
              mbar_wait(&emptyB[sb], phb ^ 1);
              if (elect_one()) {
              if (a.dbg & 2) {
                if (leader) mbar_arrive(&fullB[sb]);
              } else {
              if (leader) mbar_expect_tx(&fullB[sb], 2u * (uint32_t)Cfg::kBTap);
              tma_load_2d_pair(sB + sb * Cfg::kBTap, &tmB, mapa_u32(smem_u32(&fullB[sb]), 0),
                               tap * a.Cin + cb * kBK3, n0 + (int)crank * (BN / 2));
              }
              }
              __syncwarp();
              if (++sb == NB) {
                sb = 0;
                phb ^= 1;
              }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (leader) {  // MMA issue by the leader CTA: warp-uniform loop, one elected lane issues
      constexpr uint32_t idesc = idesc_bf16(256, BN, 0, 0);
      // A operand of sub-tile j, filter row r, filter column s (halo mode; 0 otherwise): the 128 pixel rows [BH][BW] of the box
      // [rows][BWX][64 channels] starting at pixel ((j*BH + r)*BWX + s): 8-pixel groups are BWX*128 B apart (SBO).  The
      // hardware applies the 128-byte swizzle to ABSOLUTE shared-memory address bits (scripts/umma_probe.cu, profiles/
      // r2_umma_probe.txt), so a start address that is only 128-byte aligned and a group stride that is not a multiple of
      // 1024 B address exactly the bytes TMA wrote -- base_offset stays 0.
      const uint32_t px16 = (uint32_t)Cfg::kRowBytes >> 4;            // one pixel (128 B) in descriptor units
      const uint32_t row16 = (uint32_t)a.BWX * px16;                  // one box row
      const uint32_t sub16 = (uint32_t)a.BH * row16;                  // one 128-pixel sub-tile
      const uint64_t sbo_field = (uint64_t)(((uint32_t)a.BW == 8u ? row16 : 8u * px16) & 0x3FFFu) << 32;
      const int n_outer = a.halo ? kcb : a.S * kcb;
      const int n_inner = a.halo ? a.R * a.S : a.R;
      int sa = 0, sb = 0;
      uint32_t pha = 0, phb = 0;
      int local = 0;
      for (int item = w_first; item < a.num_items; item += w_step, ++local) {
        const int as = local & 1;
        const uint32_t aphase = (local >> 1) & 1;
        mbar_wait(&acc_empty[as], aphase ^ 1);
        tc_fence_after();
        const uint32_t d0 = tmem_base + (uint32_t)(as * MT * BN);
        for (int st = 0; st < n_outer; ++st) {
          mbar_wait(&fullA[sa], pha);
          const uint64_t ad0 = (desc_kmajor(smem_u32(sA + sa * Cfg::kABytes), 128) & ~((uint64_t)0x3FFF << 32)) | sbo_field;
          for (int t = 0; t < n_inner; ++t) {
            const uint32_t r = a.halo ? (uint32_t)(t / a.S) : (uint32_t)t;
            const uint32_t sx = a.halo ? (uint32_t)(t % a.S) : 0u;
            mbar_wait(&fullB[sb], phb);
            if (local == 0 && st == 0 && t == 0 && lane == 0) STP_TRACE3(3);
            tc_fence_after();
            const uint64_t bd0 = desc_kmajor(smem_u32(sB + sb * Cfg::kBTap), 128);
            if (elect_one()) {
              if (!(a.dbg & 8))
#pragma unroll
              for (int j = 0; j < MT; ++j) {
                const uint64_t aj = ad0 + (uint64_t)(j * sub16 + r * row16 + sx * px16);
#pragma unroll
                for (int k = 0; k < kBK3 / 16; ++k)
                  umma_bf16_pair(d0 + (uint32_t)(j * BN), aj + (uint64_t)(k * 2), bd0 + (uint64_t)(k * 2), idesc, (st | t | k) != 0);
              }
              umma_commit_pair(&emptyB[sb], 3);
            }
            __syncwarp();
            if (++sb == NB) {
              sb = 0;
              phb ^= 1;
            }
          }
          if (elect_one()) umma_commit_pair(&emptyA[sa], 3);
          __syncwarp();
          if (++sa == NA) {
            sa = 0;
            pha ^= 1;
          }
        }
        if (elect_one()) umma_commit_pair(&acc_full[as], 3);
        __syncwarp();
        if (local == 0 && lane == 0) STP_TRACE3(10);
      }
      if (lane == 0) STP_TRACE3(4);
    }
  } else {
    const int q = warp & 3;
    const int m = q * 32 + lane;
    const bool ep_leader = (warp == 2 && lane == 0);
    constexpr int RB = 128;
    const uint32_t swz = (uint32_t)(m & 7);
    uint32_t res_phase = 0;
    int local = 0;
    int cur_n0 = -1;
    auto bn_flush = [&]() {
      if (cur_n0 >= 0) {
        for (int c = m; c < BN; c += 128) {
          atomicAdd(a.bn.acc + cur_n0 + c, (double)sStat[c]);
          atomicAdd(a.bn.acc + a.Cout + cur_n0 + c, (double)sStat[BN + c]);
        }
      }
    };
    for (int item = w_first; item < a.num_items; item += w_step, ++local) {
      const int as = local & 1;
      const uint32_t aphase = (local >> 1) & 1;
      int img, h0, w0, n0;
      decode(item, img, h0, w0, n0);
      if (a.bn_on && n0 != cur_n0) {  // thread m owns sStat[c], c = m, m+128: no synchronisation needed
        bn_flush();
        cur_n0 = n0;
        for (int c = m; c < BN; c += 128) sStat[c] = sStat[BN + c] = 0.f;
      }
      mbar_wait(&acc_full[as], aphase);
      if (ep_leader && local == 0) STP_TRACE3(5);
      tc_fence_after();
#pragma unroll 1
      for (int j = 0; j < MT; ++j) {
        const int hs = h0 + j * a.BH;
        if (hs >= a.Ho) break;  // rest of the strip is below the image (uniform across the CTA)
        if (ep_leader) tma_store_wait_read();  // the previous round's stores no longer read the staging tile
        named_bar_sync(1, 128);
        if (a.res) {
          if (ep_leader) {
            mbar_expect_tx(res_bar, (uint32_t)(128 * BN * 2));
#pragma unroll
            for (int sl = 0; sl < Cfg::kOutSlabs; ++sl) tma_load_4d(sOut + sl * 128 * RB, &tmR, res_bar, n0 + sl * 64, w0, hs, img);
          }
          mbar_wait(res_bar, res_phase);
          res_phase ^= 1;
        }
        const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)((as * MT + j) * BN);
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 32) {
          uint32_t rr[32];
          tmem_ld32(t_addr + c0, rr);
          tmem_ld_wait();
          float v[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(rr[i]);
          if (a.bias) {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] += __ldg(a.bias + n0 + c0 + i);
          }
          uint8_t* slab = sOut + (c0 >> 6) * (128 * RB) + m * RB;
          const uint32_t chunk0 = (uint32_t)((c0 & 63) >> 3);
#pragma unroll
          for (int i = 0; i < 32; i += 8) {
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
        fence_proxy_async();
        named_bar_sync(1, 128);
        if (ep_leader) {
#pragma unroll
          for (int sl = 0; sl < Cfg::kOutSlabs; ++sl) tma_store_4d(&tmY, sOut + sl * 128 * RB, n0 + sl * 64, w0, hs, img);
          tma_store_commit();
        }
        if (a.bn_on && img < a.nimg) {
          // BatchNorm statistics of exactly the bf16 values just staged (scheme of conv_tc2.cu): thread = (channel octet o,
          // pixel group g), 16-byte conflict-free loads of the swizzled rows, xor-shuffle tree, cross-warp combine in sRed
          constexpr int OCT = BN / 8;
          constexpr int GP = 128 / OCT;
          const int o = m % OCT, g = m / OCT;
          float s1[8], s2[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) s1[i] = s2[i] = 0.f;
          const uint8_t* sub = sOut + (o >> 3) * (128 * RB);
#pragma unroll 4
          for (int i = 0; i < OCT; ++i) {
            const int mm = g + i * GP;
            const int hh = hs + (mm >> a.log2BW), ww = w0 + (mm & (a.BW - 1));
            if (hh < a.Ho && ww < a.Wo) {
              float f[8];
              unpack8(*reinterpret_cast<const bf16x8*>(sub + mm * RB + ((((uint32_t)o & 7u) ^ (uint32_t)(mm & 7)) << 4)), f);
#pragma unroll
              for (int k = 0; k < 8; ++k) {
                s1[k] += f[k];
                s2[k] += f[k] * f[k];
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
          if (lane < OCT) {  // lane == o
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              sRed[(q * 2 + 0) * BN + lane * 8 + k] = s1[k];
              sRed[(q * 2 + 1) * BN + lane * 8 + k] = s2[k];
            }
          }
          named_bar_sync(2, 128);
          for (int c = m; c < BN; c += 128) {
            float t1 = 0.f, t2 = 0.f;
#pragma unroll
            for (int wq = 0; wq < 4; ++wq) {
              t1 += sRed[(wq * 2 + 0) * BN + c];
              t2 += sRed[(wq * 2 + 1) * BN + c];
            }
            sStat[c] += t1;
            sStat[BN + c] += t2;
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(mapa_u32(smem_u32(&acc_empty[as]), 0));  // the LEADER's MMA thread waits for both CTAs
      if (ep_leader && local == 0) STP_TRACE3(11);
    }
    if (ep_leader) STP_TRACE3(6);
    if (ep_leader) tma_store_wait_all();
    if (ep_leader) STP_TRACE3(7);
    if (a.bn_on) {
      bn_flush();
      __threadfence();
      named_bar_sync(1, 128);
      if (ep_leader) *s_last = (atomicAdd(a.bn.fin.sync, 1u) == gridDim.x - 1) ? 1 : 0;
      named_bar_sync(1, 128);
      if (*s_last) {  // every other CTA's sums have landed: finalise all channels, return the accumulators to zero
        __threadfence();
        for (int c = m; c < a.Cout; c += 128) {
          const double s1 = __ldcg(a.bn.acc + c), s2 = __ldcg(a.bn.acc + a.Cout + c);
          fin_forward(a.bn.fin, a.Cout, c, s1, s2);
          a.bn.acc[c] = 0.0;
          a.bn.acc[a.Cout + c] = 0.0;
        }
        if (ep_leader) *a.bn.fin.sync = 0u;
      }
    }
  }

  if (warp == 2 && lane == 0) STP_TRACE3(8);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // no peer still completes / arrives on this CTA's shared memory or issues MMAs into its TMEM
  if (threadIdx.x == 0) STP_TRACE3(9);
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, Cfg::kTmemCols);
  }
}

struct Tc3Plan {
  int BN, MT;
};

constexpr int kPairs = kNumSMs / 2;

// cycle estimate used to rank (BN, MT): wave quantisation over the 74 CTA pairs, L2->SM bytes per CTA (~40 B/clk) against
// the M=256 MMA floor (BN/2 clk per K=16 step), per-strip epilogue
// halo mode (option "tc3_halo" = 1; OFF by default): 8-wide x 16-row sub-tiles, ONE [TH+R-1][8+S-1] haloed A box per channel
// block instead of one [TH+R-1][BW] box per filter column.  L2->SM A bytes drop ~2.5x and parity is green, but the MMAs
// themselves run ~1.45x SLOWER when the 8-row groups of the A descriptor are not 1024-byte aligned / SBO is not a multiple of
// 1024 B (pure-MMA loop, all loads skipped: 56.1 vs 38.8 us on 384->128 @64^2; profiles/r2_tc3_dbg_bench.txt), so the layer
// loses (26.6 vs 21.9 us on 128->128 @64^2).  Kept as a measured negative result (profiles/README.md r2).
static bool tc3_halo() { return get_option(OPT_TC3_HALO) == 1; }
static int tc3_bw(const ConvP& p) { return tc3_halo() ? 8 : (p.Wo >= 16 ? 16 : 8); }

static double tc3_cost(const ConvP& p, int bn, int mt) {
  const int bw = tc3_bw(p), bh = 128 / bw;
  const int th = mt * bh;
  const int64_t pt = (int64_t)p.N * ((p.Ho + th - 1) / th) * ((p.Wo + bw - 1) / bw);
  const double items = (double)((pt + 1) / 2) * (p.Cout / bn);
  const double waves = (double)(int64_t)((items + kPairs - 1) / kPairs);
  const bool halo = tc3_halo();
  const int num_st = (halo ? 1 : p.S) * (p.Cin / kBK3);
  const int taps = halo ? p.R * p.S : p.R;
  const double bytes = (double)(th + p.R - 1) * (halo ? bw + p.S - 1 : bw) * 128 + (double)taps * (bn / 2) * 128;
  const double load = bytes / 40.0;
  const double mma = (double)mt * taps * 4 * (bn / 2.0);
  const double stage = (load > mma ? load : mma) + 100.0;
  const double epi = 600.0 + mt * (bn / 64) * 250.0;
  return waves * (num_st * stage + epi) + 4000.0;
}

bool tc3_plan(const ConvP& p, Tc3Plan* pl) {
  if (p.stride != 1 || p.up != 1) return false;
  if (p.R > 3 || p.S > 3 || (p.R < 2 && p.S < 2)) return false;
  if (p.Cin % kBK3 != 0 || p.Cout % 64 != 0) return false;
  if (p.y_f32 || p.ncls > 0) return false;
  // N = 64 pair tiles (Cout = 64 / 192 layers) only on request (option "tc3_bn64" = 1): parity green, but measured SLOWER than
  // the single-CTA kernel on every such layer (64->64 @128^2 29.9 vs 24.7 us back to back, 192->64 69.1 vs 57.7 us,
  // 64->192 73.9 vs 60.3 us; profiles/README.md r2): a weight tap then feeds only 8 MMAs of 32 clk between two commits
  if (p.Cout % 128 != 0 && get_option(OPT_TC3_BN64) != 1) return false;
  double best = 0.0;
  int bbn = 0, bmt = 0;
  // BN = 256 (no TMEM double buffering left for MT = 2, 64 KB staging) never won: only when forced.  BN = 64 serves the
  // Cout = 64 / 192 layers (stage 1, dec2, dgrad of dec2_conv1): an M = 256 x N = 64 pair MMA reads 5 KB of shared memory
  // per CTA per 32-clk MMA where the single-CTA M = 128 x N = 64 MMA of conv_tc2 reads 6 KB (operand-bandwidth bound).
  for (int bn : {128, 64}) {
    if (p.Cout % bn != 0) continue;
    if (bn == 64 && p.Cout % 128 == 0) continue;
    for (int mt = 1; mt <= (bn == 128 ? 2 : 1); ++mt) {
      const double c = tc3_cost(p, bn, mt);
      if (!bbn || c < best) {
        best = c;
        bbn = bn;
        bmt = mt;
      }
    }
  }
  const int fbn = get_option(OPT_TC3_FORCE_BN), fmt = get_option(OPT_TC3_FORCE_MT);
  if ((fbn == 128 && p.Cout % 128 == 0) || (fbn == 256 && p.Cout % 256 == 0) || fbn == 64) {
    bbn = fbn;
    if (bbn == 256) bmt = 1;
  }
  if ((fmt == 1 || fmt == 2) && !(bbn == 256 && fmt == 2)) bmt = fmt;
  pl->BN = bbn;
  pl->MT = bmt;
  return bbn != 0;
}

template <int BN, int MT>
int launch3(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmY, const CUtensorMap& tmR, const Tc3Args& a,
            cudaStream_t st) {
  using Cfg = Tc3Cfg<BN, MT>;
  auto kernel = conv_tc3_kernel<BN, MT>;
  static bool attr_set = false;
  static int max_clusters = kPairs;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes);
    if (e != cudaSuccess) {
      set_error("conv_tc3: cudaFuncSetAttribute(%d B): %s", Cfg::kSmemBytes, cudaGetErrorString(e));
      return STP_E_CUDA;
    }
    cudaLaunchConfig_t q = {};
    q.gridDim = dim3(kPairs * 2);
    q.blockDim = dim3(kThreads3);
    q.dynamicSmemBytes = Cfg::kSmemBytes;
    cudaLaunchAttribute qa[1];
    qa[0].id = cudaLaunchAttributeClusterDimension;
    qa[0].val.clusterDim.x = 2; qa[0].val.clusterDim.y = 1; qa[0].val.clusterDim.z = 1;
    q.attrs = qa; q.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, kernel, &q) == cudaSuccess && n > 0 && n < max_clusters) max_clusters = n;
    (void)cudaGetLastError();
    attr_set = true;
  }
  cudaLaunchConfig_t cfg = {};
  const int clusters = a.num_items < max_clusters ? a.num_items : max_clusters;
  cfg.gridDim = dim3(clusters * 2);
  cfg.blockDim = dim3(kThreads3);
  cfg.dynamicSmemBytes = Cfg::kSmemBytes;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = g_pdl_enabled.load(std::memory_order_relaxed) ? 1 : 0;
  attr[1].id = cudaLaunchAttributeClusterDimension;
  attr[1].val.clusterDim.x = 2; attr[1].val.clusterDim.y = 1; attr[1].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 2;
  cudaLaunchKernelEx(&cfg, kernel, tmA, tmB, tmY, tmR, a);
  g_tc_launches.fetch_add(1, std::memory_order_relaxed);
  g_tc3_launches.fetch_add(1, std::memory_order_relaxed);
  return check_launch("conv_tc3");
}

}  // namespace

bool tc3_conv_supported(const ConvP& p) {
  // option "tc3": 0 automatic | 1 off | 2 on wherever the shape is served.  Automatic: the pair kernel wins once every CTA
  // pair has work (measured on B200, profiles/README.md s27: 128->128 @64^2 25.6 vs 27.6 us, 768->256 @32^2 48.1 vs 54.2 us,
  // 384->128 @64^2 51.2 vs 58.4 us; 256->256 @32^2 equal; 512->512 @16^2 -- 32 work items -- 31.7 vs 29.7 us)
  const int opt = get_option(OPT_TC3);
  if (opt == 1) return false;
  Tc3Plan pl;
  if (!tc3_plan(p, &pl)) return false;
  if (opt == 0) {
    const int bw = tc3_bw(p), th = pl.MT * (128 / bw);
    const int64_t pt = (int64_t)p.N * ((p.Ho + th - 1) / th) * ((p.Wo + bw - 1) / bw);
    if (((pt + 1) / 2) * (p.Cout / pl.BN) < kPairs) return false;
  }
  if (p.Wo < 8 || p.Ho < 1) return false;
  if (p.ldx % 8 != 0 || !aligned16(p.x) || !aligned16(p.w)) return false;
  if (p.ldy % 8 != 0 || !aligned16(p.y)) return false;
  if (p.res && (p.ldr % 8 != 0 || !aligned16(p.res))) return false;
  return get_encode_tiled() != nullptr;
}

int launch_tc3_conv(const ConvP& p, cudaStream_t st) {
  Tc3Plan pl;
  if (!tc3_plan(p, &pl)) {
    set_error("conv_tc3: unsupported");
    return STP_E_UNSUPPORTED;
  }
  Tc3Args a;
  a.res = p.res; a.bias = p.bias; a.relu = p.relu;
  a.Ho = p.Ho; a.Wo = p.Wo; a.Cout = p.Cout; a.Cin = p.Cin; a.R = p.R; a.S = p.S; a.pad_h = p.pad_h; a.pad_w = p.pad_w;
  a.halo = tc3_halo() ? 1 : 0;
  a.BW = tc3_bw(p);
  a.BH = 128 / a.BW;
  a.log2BW = a.BW == 16 ? 4 : 3;
  a.BWX = a.halo ? a.BW + p.S - 1 : a.BW;
  const int TH = pl.MT * a.BH;
  a.tilesW = (p.Wo + a.BW - 1) / a.BW;
  a.tilesH = (p.Ho + TH - 1) / TH;
  a.tilesN = p.Cout / pl.BN;
  const int64_t npt = (int64_t)p.N * a.tilesH * a.tilesW;
  const int64_t items = ((npt + 1) / 2) * a.tilesN;
  if (items > 0x7fffffff) {
    set_error("conv_tc3: too many tiles");
    return STP_E_UNSUPPORTED;
  }
  a.numPT = (int)npt;
  a.num_items = (int)items;
  a.nimg = p.N;
  const int box_rows = TH + p.R - 1;
  a.a_bytes = box_rows * a.BWX * kBK3 * 2;
  a.trace = get_trace_buffer();
  a.dbg = get_option(OPT_TC2_DEBUG);
  a.bn_on = p.bn != nullptr ? 1 : 0;
  if (a.bn_on) a.bn = *p.bn; else a.bn = BnFuse{};
  CUtensorMap tmA, tmB, tmY, tmR;
  {
    uint64_t dims[4] = {(uint64_t)p.Cin, (uint64_t)p.W, (uint64_t)p.H, (uint64_t)p.N};
    uint64_t strides[3] = {(uint64_t)p.ldx * 2, (uint64_t)p.W * p.ldx * 2, (uint64_t)p.H * p.W * p.ldx * 2};
    uint32_t box[4] = {(uint32_t)kBK3, (uint32_t)a.BWX, (uint32_t)box_rows, 1};
    if (!make_tmap_bf16(&tmA, p.x, 4, dims, strides, box, 128)) return STP_E_CUDA;
  }
  {
    uint64_t dims[2] = {(uint64_t)p.K, (uint64_t)p.Cout};
    uint64_t strides[1] = {(uint64_t)p.K * 2};
    uint32_t box[2] = {(uint32_t)kBK3, (uint32_t)(pl.BN / 2)};  // each CTA of the pair loads half of the weight rows
    if (!make_tmap_bf16(&tmB, p.w, 2, dims, strides, box, 128)) return STP_E_CUDA;
  }
  {
    uint64_t dims[4] = {(uint64_t)p.Cout, (uint64_t)p.Wo, (uint64_t)p.Ho, (uint64_t)p.N};
    uint64_t strides[3] = {(uint64_t)p.ldy * 2, (uint64_t)p.Wo * p.ldy * 2, (uint64_t)p.Ho * p.Wo * p.ldy * 2};
    uint32_t box[4] = {64, (uint32_t)a.BW, (uint32_t)a.BH, 1};
    if (!make_tmap_bf16(&tmY, p.y, 4, dims, strides, box, 128)) return STP_E_CUDA;
    tmR = tmY;
    if (p.res) {
      uint64_t rstrides[3] = {(uint64_t)p.ldr * 2, (uint64_t)p.Wo * p.ldr * 2, (uint64_t)p.Ho * p.Wo * p.ldr * 2};
      if (!make_tmap_bf16(&tmR, p.res, 4, dims, rstrides, box, 128)) return STP_E_CUDA;
    }
  }
  if (pl.BN == 64) return pl.MT == 2 ? launch3<64, 2>(tmA, tmB, tmY, tmR, a, st) : launch3<64, 1>(tmA, tmB, tmY, tmR, a, st);
  if (pl.BN == 256) return launch3<256, 1>(tmA, tmB, tmY, tmR, a, st);
  if (pl.MT == 2) return launch3<128, 2>(tmA, tmB, tmY, tmR, a, st);
  return launch3<128, 1>(tmA, tmB, tmY, tmR, a, st);
}

}  // namespace stp
