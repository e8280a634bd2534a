// Shared parameter blocks + launchers of the convolution kernels (generic mma.sync and tcgen05 paths).
#pragma once
#include "bn_fin.cuh"
#include "common.cuh"

namespace stp {

// Forward BatchNorm statistics of a convolution's OUTPUT, accumulated by the conv epilogue itself (conv_tc2.cu): per-CTA
// fp32 sums of the bf16 values it stores -> double atomics into acc[2][Cout] -> the last CTA (ticket fin.sync)
// finalises into fin.coef / moving statistics and returns acc to zero.
struct BnFuse {
  double* acc;
  FinArgs fin;  // mode 1
};

// bn.cu: bn.acc ([2][C] sums left by a producing kernel) -> coefficients (mode 1) / backward terms (mode 2), accumulators re-zeroed
int launch_bn_acc_finalize(const BnFuse& bn, int C, int mode, cudaStream_t st);

struct ConvP {
  const __nv_bfloat16* x;
  int ldx, N, H, W, Cin;
  const __nv_bfloat16* w;
  void* y;
  int ldy, Ho, Wo, Cout, y_f32;
  const __nv_bfloat16* res;
  int ldr;
  const float* bias;
  int R, S, stride, pad_h, pad_w, up, relu;
  int64_t M;
  int K;
  const BnFuse* bn = nullptr;  // non-null: also produce the BatchNorm statistics of y (halo kernel, bf16 output only)
  // bn->fin.mode == 2 (dgrad launches): y is d(BatchNorm+ReLU output); bnb_x is that BatchNorm's INPUT, bnb_coef its
  // [4][Cout] coefficients: the epilogue stores g = y*[x*scale+shift > 0] and reduces (sum g, sum g*x) -> dgamma, dbeta, bcoef
  const __nv_bfloat16* bnb_x = nullptr;
  int bnb_ldx = 0;
  const float* bnb_coef = nullptr;
  int bnb_relu = 0;
  int ncls = 0;  // > 0: segmentation-head epilogue (tcgen05 halo kernel only): y = dense fp32 [M][ncls], Cout is padding
  // conv_tc2 only -- one output parity class of a zero-insertion (up == 2) problem run as a stride-1 convolution (launch_tc2_up2):
  // element strides of the output / residual VIEW (0: dense, ldy / Wo*ldy / Ho*Wo*ldy) and the tap subset of the weight tensor
  // [Cout][R0][w_S][Cin] it uses: filter tap (r, s) of this launch is tap (w_r0 + r*w_rs, w_s0 + s*w_ss) of the original
  int64_t y_sw = 0, y_sh = 0, y_sn = 0, r_sw = 0, r_sh = 0, r_sn = 0;
  int w_r0 = 0, w_rs = 1, w_s0 = 0, w_ss = 1, w_S = 0, w_K = 0;  // w_S / w_K == 0: S / K
};

struct WgradP {
  const __nv_bfloat16* x;
  int ldx, N, H, W, Cin;
  const __nv_bfloat16* dy;
  int lddy, Ho, Wo, Cout;
  float* out;  // [splits][Cout][K]
  int R, S, stride, pad_h, pad_w, up;
  int64_t M;
  int K;
  int64_t chunks_per_split;  // in units of 32 pixels
};


// conv_generic.cu
int launch_generic_conv(const ConvP& p, cudaStream_t st);
int launch_generic_wgrad(WgradP p, float* dw, void* ws, size_t ws_bytes, cudaStream_t st);
size_t generic_wgrad_workspace(int64_t M, int Cout, int K);
int launch_split_reduce(const float* ws, int splits, int64_t n, float* out, cudaStream_t st);

// conv_tc2.cu (tcgen05 + TMA, halo kernel for RxS > 1)
bool tc2_conv_supported(const ConvP& p);
int launch_tc2_conv(const ConvP& p, cudaStream_t st);
// zero-insertion problems (p.up == 2: the dgrad of a stride-2 convolution, Conv2DTranspose): four stride-1 launches, one per
// output parity class, each with its tap subset and a strided output view
bool tc2_up2_supported(const ConvP& p);
int launch_tc2_up2(const ConvP& p, cudaStream_t st);
int get_option(int key);
enum { OPT_TC2_FORCE_MT = 0, OPT_TC_CONV_VERSION = 1, OPT_TC2_DEBUG = 2, OPT_TC2_CLUSTER = 3, OPT_TC2_BK = 4, OPT_TC3 = 5,
       OPT_TC3_FORCE_BN = 6, OPT_TC3_FORCE_MT = 7, OPT_BN_BLOCKS = 8, OPT_BNB_FUSE = 9, OPT_TC3_HALO = 10, OPT_TC3_BN64 = 11, OPT_HEAD_STRIP = 12, OPT_GEMM1X1 = 13, OPT_NCONV = 14, OPT_TC2_1X1 = 15, OPT_TC2_UP2 = 16, OPT_G1_BN = 17, OPT_WGRAD1X1 = 18, OPT_COUNT = 24 };
// device buffer (>= 64 uint64) that CTA 0 and the last CTA of conv_tc2_kernel fill with %globaltimer stamps of their
// phases (scripts/trace_conv.py); nullptr = off.  Profiling aid only.
unsigned long long* get_trace_buffer();

// conv_tc3.cu (tcgen05 cta_group::2 CTA-pair halo kernel: 3x3 stride 1, Cin % 64 == 0, Cout % 128 == 0, bf16 output)
bool tc3_conv_supported(const ConvP& p);
int launch_tc3_conv(const ConvP& p, cudaStream_t st);

// wgrad_narrow.cu (mma.sync + TMA, Cin/Cout in {16,32})
bool narrow_wgrad_supported(const WgradP& p);
size_t narrow_wgrad_workspace(const WgradP& p);
int launch_narrow_wgrad(const WgradP& p, float* dw, void* ws, size_t ws_bytes, cudaStream_t st);

// wgrad1x1.cu (cp.async + ldmatrix.trans + mma.sync: weight gradient of 1x1 stride-1 convs the tcgen05 wgrad kernel does not tile)
bool wgrad1x1_supported(const WgradP& p);
size_t wgrad1x1_workspace(const WgradP& p);
int launch_wgrad1x1(const WgradP& p, float* dw, void* ws, size_t ws_bytes, cudaStream_t st);

// conv_narrow.cu (cp.async halo tiles + ldmatrix + mma.sync: 3x3 stride-1 "same" conv, Cin / Cout in {16, 32}, HBM-bound decoder tail)
bool narrow_conv_supported(const ConvP& p);
int launch_narrow_conv(const ConvP& p, cudaStream_t st);

// gemm1x1.cu (cp.async + ldmatrix + mma.sync streaming GEMM for 1x1 stride-1 convolutions the tcgen05 kernels do not tile)
bool gemm1x1_supported(const ConvP& p);
int launch_gemm1x1(const ConvP& p, cudaStream_t st);

// conv_tc.cu (tcgen05 + TMA)
bool tc_conv_supported(const ConvP& p);
int launch_tc_conv(const ConvP& p, cudaStream_t st);
bool tc_wgrad_supported(const WgradP& p);
int launch_tc_wgrad(const WgradP& p, float* dw, void* ws, size_t ws_bytes, cudaStream_t st);
size_t tc_wgrad_workspace(const WgradP& p);

}  // namespace stp
