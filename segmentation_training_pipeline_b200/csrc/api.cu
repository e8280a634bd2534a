// C-ABI plumbing: error state, version, launch counter and the convolution entry points, which pick a
// kernel SPECIALISATION by shape (tcgen05/TMA implicit GEMM for the FLOP-heavy layers, mma.sync generic
// implicit GEMM for the irregular / HBM-bound ones).  No CPU fallback anywhere.
#include <stdarg.h>
#include <string.h>

#include "common.cuh"
#include "conv.h"
#include "f32_path.h"

namespace stp {

static thread_local char g_err[512] = "";
std::atomic<int64_t> g_launches{0};
std::atomic<int64_t> g_tc_launches{0};
std::atomic<int64_t> g_tc3_launches{0};
static std::atomic<int> g_tc_enabled{1};
std::atomic<int> g_pdl_enabled{1};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

}  // namespace stp

using namespace stp;

namespace stp {
static std::atomic<int> g_options[OPT_COUNT];
int get_option(int key) { return (key >= 0 && key < OPT_COUNT) ? g_options[key].load() : 0; }
}  // namespace stp

namespace stp {
static std::atomic<unsigned long long*> g_trace{nullptr};
unsigned long long* get_trace_buffer() { return g_trace.load(); }
}  // namespace stp
extern "C" void stp_set_trace_buffer(void* dev_ptr) { stp::g_trace.store((unsigned long long*)dev_ptr); }

extern "C" int stp_set_option(const char* name, int32_t value) {
  STP_REQUIRE(name, "set_option: null name");
  int key = -1;
  if (!strcmp(name, "tc2_force_mt")) key = OPT_TC2_FORCE_MT;        /* 0 = heuristic; 1,2,4,8 = force strip height */
  else if (!strcmp(name, "tc_conv_version")) key = OPT_TC_CONV_VERSION; /* 0 = auto, 1 = first-generation kernel only */
  else if (!strcmp(name, "tc2_debug")) key = OPT_TC2_DEBUG;             /* timing experiments, see conv_tc2.cu */
  else if (!strcmp(name, "tc2_cluster")) key = OPT_TC2_CLUSTER;         /* 0 auto | 1 no clusters | 2, 4 force that cluster size */
  else if (!strcmp(name, "tc2_bk")) key = OPT_TC2_BK;                   /* 0 auto | 32: 32-channel K blocks even when Cin % 64 == 0 */
  else if (!strcmp(name, "tc3")) key = OPT_TC3;                         /* 0 auto | 1 off | 2 CTA-pair kernel wherever it serves the shape */
  else if (!strcmp(name, "tc3_force_bn")) key = OPT_TC3_FORCE_BN;       /* 0 heuristic | 128, 256 */
  else if (!strcmp(name, "tc3_force_mt")) key = OPT_TC3_FORCE_MT;       /* 0 heuristic | 1, 2 */
  else if (!strcmp(name, "tc3_halo")) key = OPT_TC3_HALO;               /* 0 off | 1 on: ONE haloed A box per channel block (measured slower, see conv_tc3.cu) */
  else if (!strcmp(name, "gemm1x1")) key = OPT_GEMM1X1;                 /* 0 auto: 1x1 stride-1 convs the tcgen05 halo kernel does not tile (MobileNetV2 / Xception widths) on the streaming mma.sync GEMM | 1 off | 2 every eligible 1x1 conv */
  else if (!strcmp(name, "nconv")) key = OPT_NCONV;                     /* 0 off (measured slower than the tcgen05 halo kernel, conv_narrow.cu) | 1: 3x3 convs with Cin, Cout in {16, 32} on the mma.sync narrow-channel kernel */
  else if (!strcmp(name, "tc2_1x1")) key = OPT_TC2_1X1;                 /* 0 auto: 1x1 stride-1 convs take the halo kernel conv_tc2 (a plain GEMM over pixel strips) where it tiles the shape | 1 off | 2 only Cin % 64 == 0 */
  else if (!strcmp(name, "tc2_up2")) key = OPT_TC2_UP2;                 /* 0 auto: zero-insertion convs (stride-2 dgrads) as four halo-kernel launches, one per output parity class | 1 off */
  else if (!strcmp(name, "g1_bn")) key = OPT_G1_BN;                     /* 0 auto | 1: no BatchNorm epilogues in the streaming 1x1 GEMM (conv + separate reduction pass) */
  else if (!strcmp(name, "wgrad1x1")) key = OPT_WGRAD1X1;               /* 0 auto | 1 off: 1x1 weight gradients the tcgen05 kernel does not tile stay on the generic kernel */
  else if (!strcmp(name, "head_strip")) key = OPT_HEAD_STRIP;           /* 0 on | 1 off: column-strip head backward kernels (sliding dlogit window) */
  else if (!strcmp(name, "tc3_bn64")) key = OPT_TC3_BN64;               /* 0 off | 1 on: N = 64 CTA-pair tiles for Cout = 64 / 192 layers (measured slower) */
  else if (!strcmp(name, "bnb_fuse")) key = OPT_BNB_FUSE;               /* 0 auto | 1: never fuse the BatchNorm-backward reduction into the dgrad epilogue */
  else if (!strcmp(name, "bn_blocks")) key = OPT_BN_BLOCKS;             /* 0 default | n: atomic-mode BN reductions use up to n*1024/C blocks */
  else if (!strcmp(name, "pdl")) {                                      /* programmatic dependent launch on/off */
    g_pdl_enabled.store(value ? 1 : 0);
    return STP_OK;
  }
  STP_REQUIRE(key >= 0, "set_option: unknown option %s", name);
  g_options[key].store(value);
  return STP_OK;
}

extern "C" int stp_version(void) { return STP_VERSION; }
extern "C" const char* stp_last_error(void) { return g_err; }
extern "C" int64_t stp_launch_count(void) { return g_launches.load(); }
extern "C" int64_t stp_tc_launch_count(void) { return g_tc_launches.load(); }
extern "C" int64_t stp_tc3_launch_count(void) { return g_tc3_launches.load(); }
extern "C" int stp_tc_enabled(void) { return g_tc_enabled.load(); }
extern "C" void stp_set_tc_enabled(int on) { g_tc_enabled.store(on ? 1 : 0); }

static int dispatch_conv(const ConvP& p, cudaStream_t st) {
  // 1x1 stride-1 convolutions (profiles/r2_s10_g1_bench.txt, bs16 back to back): the tcgen05 halo kernel conv_tc2 run as a plain
  // GEMM wins wherever it tiles the shape (64 -> 256 @128^2: 42.6 us against 72.5 us streaming mma.sync GEMM and 193 us first-
  // generation tcgen05 kernel); the streaming GEMM (legacy tensor path, ~290 TF/s ceiling) serves the widths it does not tile
  // (24, 96, 144, 160, 728 ...); the first-generation kernel keeps the few-pixel long-K corner (2048 -> 512 @16^2: 16.3 vs 23.0 us).
  const int g1 = get_option(OPT_GEMM1X1);
  const bool tc = stp_tc_enabled() != 0;
  bool long_k_few_pixels = false;
  if (p.R == 1 && p.S == 1 && p.stride == 1 && p.up == 1) {
    if (g1 == 2 && gemm1x1_supported(p)) return launch_gemm1x1(p, st);
    long_k_few_pixels = p.M <= 8192 && p.Cin >= 1024 && tc && tc_conv_supported(p);
    if (tc && get_option(OPT_TC_CONV_VERSION) != 1 && !long_k_few_pixels && tc2_conv_supported(p)) return launch_tc2_conv(p, st);
    if (g1 == 0 && !long_k_few_pixels && gemm1x1_supported(p)) return launch_gemm1x1(p, st);
  }
  if (get_option(OPT_NCONV) == 1 && narrow_conv_supported(p)) return launch_narrow_conv(p, st);
  if (tc) {
    if (get_option(OPT_TC_CONV_VERSION) != 1 && get_option(OPT_TC2_UP2) != 1 && tc2_up2_supported(p)) return launch_tc2_up2(p, st);
    if (get_option(OPT_TC_CONV_VERSION) != 1 && tc3_conv_supported(p)) return launch_tc3_conv(p, st);
    if (get_option(OPT_TC_CONV_VERSION) != 1 && !long_k_few_pixels && tc2_conv_supported(p)) return launch_tc2_conv(p, st);
    if (tc_conv_supported(p)) return launch_tc_conv(p, st);
  }
  return launch_generic_conv(p, st);
}

static int check_conv_common(const stp_conv_desc* d, const stp_tensor* in, const stp_tensor* out, const char* who) {
  STP_REQUIRE(d && in && out, "%s: null argument", who);
  STP_REQUIRE(d->r >= 1 && d->s >= 1 && d->stride >= 1 && d->up >= 1 && d->pad_h >= 0 && d->pad_w >= 0,
              "%s: bad descriptor", who);
  STP_REQUIRE(vec_ok(in), "%s: input must be bf16 NHWC, c%%8==0, ld%%8==0, 16B aligned", who);
  STP_REQUIRE(in->n == out->n, "%s: batch mismatch", who);
  return STP_OK;
}

// ---- parity mode (fp32 tensors): CUDA-core kernels of f32_path.cu -------------------------------------------------
static f32::ConvF make_conv_f(const stp_conv_desc* d, const stp_tensor* x, const void* w, const float* bias,
                              const stp_tensor* residual, const stp_tensor* y) {
  f32::ConvF p;
  p.x = (const float*)x->ptr; p.ldx = x->ld; p.N = x->n; p.H = x->h; p.W = x->w; p.Cin = x->c;
  p.w = (const float*)w;
  p.y = (float*)y->ptr; p.ldy = y->ld; p.Ho = y->h; p.Wo = y->w; p.Cout = y->c;
  p.res = residual ? (const float*)residual->ptr : nullptr; p.ldr = residual ? residual->ld : 0;
  p.bias = bias;
  p.R = d->r; p.S = d->s; p.stride = d->stride; p.pad_h = d->pad_h; p.pad_w = d->pad_w; p.up = d->up;
  p.relu = (d->flags & STP_CONV_RELU) ? 1 : 0; p.dgrad = 0;
  p.M = pixels(y); p.K = d->r * d->s * x->c;
  return p;
}
static int conv_fwd_f32(const stp_conv_desc* d, const stp_tensor* x, const void* w, const float* bias, const stp_tensor* residual,
                        const stp_tensor* y, const stp_bn_fwd* h_bn, stp_stream stream) {
  STP_REQUIRE(d && f32::f32_ok(x) && f32::f32_ok(y) && w && x->n == y->n, "conv_fwd (fp32 parity mode): bad tensors");
  if (residual) STP_REQUIRE(f32::f32_ok(residual) && residual->c == y->c && pixels(residual) == pixels(y), "conv_fwd (fp32): bad residual");
  int rc = f32::launch_conv(make_conv_f(d, x, w, bias, residual, y), (cudaStream_t)stream);
  if (rc || !h_bn) return rc;
  return stp_bn_stats_fused(y, h_bn->partial, h_bn->sync, h_bn->acc, h_bn->gamma, h_bn->beta, h_bn->eps, h_bn->momentum,
                            h_bn->moving_mean, h_bn->moving_var, h_bn->coef, stream);
}

static int conv_fwd_common(const stp_conv_desc* d, const stp_tensor* x, const void* w_krsc, const float* bias,
                           const stp_tensor* residual, const stp_tensor* y, const stp_bn_fwd* h_bn, stp_stream stream) {
  if (x && x->dtype == STP_F32) return conv_fwd_f32(d, x, w_krsc, bias, residual, y, h_bn, stream);
  int rc = check_conv_common(d, x, y, "conv_fwd");
  if (rc) return rc;
  STP_REQUIRE(w_krsc && y->ptr, "conv_fwd: null weights/output");
  STP_REQUIRE(y->dtype == STP_BF16 || y->dtype == STP_F32, "conv_fwd: y must be bf16 or f32");
  STP_REQUIRE(y->ld >= y->c, "conv_fwd: bad output ld");
  if (residual)
    STP_REQUIRE(residual->dtype == STP_BF16 && residual->c == y->c && pixels(residual) == pixels(y),
                "conv_fwd: bad residual");
  ConvP p;
  p.x = (const __nv_bfloat16*)x->ptr; p.ldx = x->ld; p.N = x->n; p.H = x->h; p.W = x->w; p.Cin = x->c;
  p.w = (const __nv_bfloat16*)w_krsc;
  p.y = y->ptr; p.ldy = y->ld; p.Ho = y->h; p.Wo = y->w; p.Cout = y->c; p.y_f32 = y->dtype == STP_F32;
  p.res = residual ? (const __nv_bfloat16*)residual->ptr : nullptr; p.ldr = residual ? residual->ld : 0;
  p.bias = bias;
  p.R = d->r; p.S = d->s; p.stride = d->stride; p.pad_h = d->pad_h; p.pad_w = d->pad_w; p.up = d->up;
  p.relu = (d->flags & STP_CONV_RELU) ? 1 : 0;
  p.M = pixels(y); p.K = d->r * d->s * x->c;
  if (!h_bn) return dispatch_conv(p, (cudaStream_t)stream);
  // BatchNorm statistics of y: inside the conv epilogue when the halo kernel serves this shape, else one extra pass
  STP_REQUIRE(h_bn->partial && h_bn->sync && h_bn->acc && h_bn->coef, "conv_fwd_bn: null statistics buffers");
  STP_REQUIRE(y->dtype == STP_BF16 && pixels(y) > 0, "conv_fwd_bn: y must be a non-empty bf16 tensor");
  {
    const int64_t count = pixels(y);
    BnFuse bn;
    bn.acc = h_bn->acc;
    bn.fin = FinArgs{};
    bn.fin.mode = 1; bn.fin.sync = h_bn->sync; bn.fin.inv_count = 1.0 / (double)count;
    bn.fin.bessel = keras_bessel(count, h_bn->eps);
    bn.fin.gamma = h_bn->gamma; bn.fin.beta = h_bn->beta; bn.fin.eps = h_bn->eps; bn.fin.momentum = h_bn->momentum;
    bn.fin.mov_mean = h_bn->moving_mean; bn.fin.mov_var = h_bn->moving_var; bn.fin.coef = h_bn->coef;
    p.bn = &bn;
    if (get_option(OPT_NCONV) == 1 && narrow_conv_supported(p)) return launch_narrow_conv(p, (cudaStream_t)stream);
    if (stp_tc_enabled() && get_option(OPT_TC_CONV_VERSION) != 1) {
      if (tc3_conv_supported(p)) return launch_tc3_conv(p, (cudaStream_t)stream);
      if (tc2_conv_supported(p)) return launch_tc2_conv(p, (cudaStream_t)stream);
    }
    // 1x1 layers the halo kernel does not tile (MobileNetV2 / Xception widths): statistics in the streaming GEMM's epilogue
    if (get_option(OPT_GEMM1X1) != 1 && get_option(OPT_G1_BN) != 1 && gemm1x1_supported(p)) return launch_gemm1x1(p, (cudaStream_t)stream);
    p.bn = nullptr;
  }
  rc = dispatch_conv(p, (cudaStream_t)stream);
  if (rc) return rc;
  return stp_bn_stats_fused(y, h_bn->partial, h_bn->sync, h_bn->acc, h_bn->gamma, h_bn->beta, h_bn->eps, h_bn->momentum,
                            h_bn->moving_mean, h_bn->moving_var, h_bn->coef, stream);
}

extern "C" int stp_conv_fwd(const stp_conv_desc* d, const stp_tensor* x, const void* w_krsc, const float* bias,
                            const stp_tensor* residual, const stp_tensor* y, void* workspace, size_t workspace_bytes,
                            stp_stream stream) {
  (void)workspace;
  (void)workspace_bytes;
  return conv_fwd_common(d, x, w_krsc, bias, residual, y, nullptr, stream);
}

extern "C" int stp_conv_fwd_bn(const stp_conv_desc* d, const stp_tensor* x, const void* w_krsc, const float* bias,
                               const stp_tensor* residual, const stp_tensor* y, const stp_bn_fwd* h_bn, void* workspace,
                               size_t workspace_bytes, stp_stream stream) {
  (void)workspace;
  (void)workspace_bytes;
  STP_REQUIRE(h_bn, "conv_fwd_bn: null h_bn");
  return conv_fwd_common(d, x, w_krsc, bias, residual, y, h_bn, stream);
}

static int conv_dgrad_common(const stp_conv_desc* d, const stp_tensor* dy, const void* w_dgrad, const stp_tensor* residual,
                             const stp_tensor* dx, const stp_bn_bwd* h_bnb, stp_stream stream) {
  if (dy && dy->dtype == STP_F32) {
    // parity mode: w_dgrad is the FORWARD fp32 KRSC tensor (flipped / transposed inside the kernel)
    STP_REQUIRE(d && f32::f32_ok(dy) && f32::f32_ok(dx) && w_dgrad && dx->n == dy->n && !h_bnb, "conv_dgrad (fp32 parity mode): bad args");
    STP_REQUIRE(d->up == 1 || d->stride == 1, "conv_dgrad: stride and up cannot both be > 1");
    STP_REQUIRE(d->r - 1 - d->pad_h >= 0 && d->s - 1 - d->pad_w >= 0, "conv_dgrad: pad > filter-1 unsupported");
    if (residual) STP_REQUIRE(f32::f32_ok(residual) && residual->c == dx->c && pixels(residual) == pixels(dx), "conv_dgrad (fp32): bad residual");
    f32::ConvF p = make_conv_f(d, dy, w_dgrad, nullptr, residual, dx);
    p.stride = d->up; p.up = d->stride; p.pad_h = d->r - 1 - d->pad_h; p.pad_w = d->s - 1 - d->pad_w; p.relu = 0; p.dgrad = 1;
    return f32::launch_conv(p, (cudaStream_t)stream);
  }
  int rc = check_conv_common(d, dy, dx, "conv_dgrad");
  if (rc) return rc;
  STP_REQUIRE(w_dgrad && dx->ptr && dx->dtype == STP_BF16 && dx->ld >= dx->c, "conv_dgrad: bad dx / weights");
  STP_REQUIRE(d->up == 1 || d->stride == 1, "conv_dgrad: stride and up cannot both be > 1");
  STP_REQUIRE(d->r - 1 - d->pad_h >= 0 && d->s - 1 - d->pad_w >= 0, "conv_dgrad: pad > filter-1 unsupported");
  if (residual)
    STP_REQUIRE(residual->dtype == STP_BF16 && residual->c == dx->c && pixels(residual) == pixels(dx),
                "conv_dgrad: bad residual");
  ConvP p;
  p.x = (const __nv_bfloat16*)dy->ptr; p.ldx = dy->ld; p.N = dy->n; p.H = dy->h; p.W = dy->w; p.Cin = dy->c;
  p.w = (const __nv_bfloat16*)w_dgrad;
  p.y = dx->ptr; p.ldy = dx->ld; p.Ho = dx->h; p.Wo = dx->w; p.Cout = dx->c; p.y_f32 = 0;
  p.res = residual ? (const __nv_bfloat16*)residual->ptr : nullptr; p.ldr = residual ? residual->ld : 0;
  p.bias = nullptr;
  p.R = d->r; p.S = d->s;
  // forward  y[o] = sum_r x[(o*stride - pad + r)/up]      (up==1 or stride==1)
  // backward dx[i] = sum_r' dy[(i*up - (R-1-pad) + r')/stride]  with r' = R-1-r
  p.stride = d->up; p.up = d->stride; p.pad_h = d->r - 1 - d->pad_h; p.pad_w = d->s - 1 - d->pad_w;
  p.relu = 0;
  p.M = pixels(dx); p.K = d->r * d->s * dy->c;
  if (!h_bnb) return dispatch_conv(p, (cudaStream_t)stream);
  // BatchNorm-backward reduction of the layer that PRODUCED this conv's input: inside the dgrad epilogue when the halo
  // kernel serves the shape (dx is then stored already masked by the ReLU), else one extra pass over (dx, x)
  STP_REQUIRE(h_bnb->x && vec_ok(h_bnb->x) && h_bnb->x->c == dx->c && pixels(h_bnb->x) == pixels(dx) && h_bnb->coef &&
                  h_bnb->sync && h_bnb->acc && h_bnb->bcoef && h_bnb->partial,
              "conv_dgrad_bn: bad BatchNorm arguments");
  STP_REQUIRE(!residual, "conv_dgrad_bn: the fused reduction needs the COMPLETE gradient (no residual accumulation)");
  BnFuse bn;
  bn.acc = h_bnb->acc;
  bn.fin = FinArgs{};
  bn.fin.mode = 2; bn.fin.sync = h_bnb->sync; bn.fin.acc = h_bnb->acc; bn.fin.inv_count = 1.0 / (double)pixels(dx);
  bn.fin.coef = const_cast<float*>(h_bnb->coef); bn.fin.dgamma = h_bnb->dgamma; bn.fin.dbeta = h_bnb->dbeta;
  bn.fin.bcoef = h_bnb->bcoef;
  p.bn = &bn;
  p.bnb_x = (const __nv_bfloat16*)h_bnb->x->ptr; p.bnb_ldx = h_bnb->x->ld; p.bnb_coef = h_bnb->coef; p.bnb_relu = h_bnb->relu;
  // (the fused epilogue masks ReLU only: a ReLU6 layer, relu == 2, takes the two-pass path below)
  if (get_option(OPT_BNB_FUSE) != 1 && h_bnb->relu != 2 && get_option(OPT_NCONV) == 1 && narrow_conv_supported(p))
    return launch_narrow_conv(p, (cudaStream_t)stream);
  if (stp_tc_enabled() && get_option(OPT_TC_CONV_VERSION) != 1 && get_option(OPT_BNB_FUSE) != 1 && h_bnb->relu != 2 && tc2_conv_supported(p))
    return launch_tc2_conv(p, (cudaStream_t)stream);
  // 1x1 dgrads the halo kernel does not serve (MobileNetV2 / Xception widths, ReLU6 masks): the streaming GEMM's epilogue
  if (get_option(OPT_BNB_FUSE) != 1 && get_option(OPT_GEMM1X1) != 1 && get_option(OPT_G1_BN) != 1 && gemm1x1_supported(p))
    return launch_gemm1x1(p, (cudaStream_t)stream);
  p.bn = nullptr;
  rc = dispatch_conv(p, (cudaStream_t)stream);
  if (rc) return rc;
  return stp_bn_bwd_reduce_fused(dx, h_bnb->x, h_bnb->coef, h_bnb->relu, 1, h_bnb->partial, h_bnb->sync, h_bnb->acc,
                                 h_bnb->dgamma, h_bnb->dbeta, h_bnb->bcoef, stream);
}

extern "C" int stp_conv_dgrad(const stp_conv_desc* d, const stp_tensor* dy, const void* w_dgrad,
                              const stp_tensor* residual, const stp_tensor* dx, void* workspace,
                              size_t workspace_bytes, stp_stream stream) {
  (void)workspace;
  (void)workspace_bytes;
  return conv_dgrad_common(d, dy, w_dgrad, residual, dx, nullptr, stream);
}

extern "C" int stp_conv_dgrad_bn(const stp_conv_desc* d, const stp_tensor* dy, const void* w_dgrad, const stp_tensor* dx,
                                 const stp_bn_bwd* h_bnb, void* workspace, size_t workspace_bytes, stp_stream stream) {
  (void)workspace;
  (void)workspace_bytes;
  STP_REQUIRE(h_bnb, "conv_dgrad_bn: null h_bnb");
  return conv_dgrad_common(d, dy, w_dgrad, nullptr, dx, h_bnb, stream);
}

static WgradP make_wgrad_p(const stp_conv_desc* d, const stp_tensor* x, const stp_tensor* dy) {
  WgradP p;
  p.x = (const __nv_bfloat16*)x->ptr; p.ldx = x->ld; p.N = x->n; p.H = x->h; p.W = x->w; p.Cin = x->c;
  p.dy = (const __nv_bfloat16*)dy->ptr; p.lddy = dy->ld; p.Ho = dy->h; p.Wo = dy->w; p.Cout = dy->c;
  p.out = nullptr;
  p.R = d->r; p.S = d->s; p.stride = d->stride; p.pad_h = d->pad_h; p.pad_w = d->pad_w; p.up = d->up;
  p.M = pixels(dy); p.K = d->r * d->s * x->c;
  p.chunks_per_split = 0;
  return p;
}

extern "C" size_t stp_conv_wgrad_workspace(const stp_conv_desc* d, const stp_tensor* x, const stp_tensor* dy) {
  if (!d || !x || !dy) return 0;
  WgradP p = make_wgrad_p(d, x, dy);
  size_t a = generic_wgrad_workspace(p.M, p.Cout, p.K);
  size_t b = tc_wgrad_workspace(p);
  size_t c = narrow_wgrad_workspace(p);
  size_t e = wgrad1x1_workspace(p);
  if (b > a) a = b;
  if (e > a) a = e;
  return a > c ? a : c;
}

extern "C" int stp_conv_wgrad(const stp_conv_desc* d, const stp_tensor* x, const stp_tensor* dy, float* dw,
                              void* workspace, size_t workspace_bytes, stp_stream stream) {
  if (x && x->dtype == STP_F32) {
    STP_REQUIRE(d && f32::f32_ok(x) && f32::f32_ok(dy) && dw && x->n == dy->n, "conv_wgrad (fp32 parity mode): bad tensors");
    f32::ConvF p = make_conv_f(d, x, nullptr, nullptr, nullptr, dy);   // geometry: x gathered at the positions of dy's pixels
    return f32::launch_wgrad(p, (const float*)dy->ptr, dy->ld, dy->c, dw, (cudaStream_t)stream);
  }
  int rc = check_conv_common(d, x, dy, "conv_wgrad");
  if (rc) return rc;
  STP_REQUIRE(vec_ok(dy) && dw, "conv_wgrad: dy must be bf16 NHWC c%%8==0; dw non-null");
  WgradP p = make_wgrad_p(d, x, dy);
  if (stp_tc_enabled() && narrow_wgrad_supported(p)) return launch_narrow_wgrad(p, dw, workspace, workspace_bytes, (cudaStream_t)stream);
  if (stp_tc_enabled() && tc_wgrad_supported(p)) return launch_tc_wgrad(p, dw, workspace, workspace_bytes, (cudaStream_t)stream);
  // 1x1 layers of MobileNetV2 / Xception widths: a pixel-reduction GEMM instead of the generic im2col kernel
  if (get_option(OPT_WGRAD1X1) != 1 && wgrad1x1_supported(p)) return launch_wgrad1x1(p, dw, workspace, workspace_bytes, (cudaStream_t)stream);
  return launch_generic_wgrad(p, dw, workspace, workspace_bytes, (cudaStream_t)stream);
}
