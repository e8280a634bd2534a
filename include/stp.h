/*
 * libstp -- C ABI of the B200-native segmentation training hot path.
 *
 * The reference (musket-ml/segmentation_training_pipeline) has no native code and no FFI: its hot path is
 * Keras `train_on_batch` on a graph built at segmentation_pipeline/segmentation.py:96-155 (createNet1) fed by
 * the imgaug generator selected at segmentation.py:54.  Each entry point below names the reference
 * call site / dependency op it replaces.  INTEGRATION.md shows the ctypes
 * binding a maintainer of the reference would add.
 *
 * Conventions
 *   - plain pointers and sizes only; no torch types.  All `void*`/typed pointers are DEVICE pointers
 *     unless the name starts with `h_`.
 *   - the caller owns every buffer; the library never allocates device memory, never synchronises,
 *     never changes the current device.  Work is enqueued on `stream` (a cudaStream_t).
 *   - returns 0 on success, a negative STP_E_* otherwise; message via stp_last_error() (thread local).
 *   - activations: NHWC bf16, channel count a multiple of 8, 16-byte aligned, `ld` = elements between
 *     consecutive pixels (>= c; lets a tensor be a channel slice of a concat buffer -> zero-copy concat).
 *   - conv weights: bf16 [Cout][R][S][Cin] ("KRSC", Cin == x.c); master weights / gradients fp32 same order.
 */
#ifndef STP_H_
#define STP_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define STP_VERSION 100

enum {
  STP_OK = 0,
  STP_E_INVALID = -1,     /* bad argument / shape */
  STP_E_UNSUPPORTED = -2, /* shape not covered by any kernel specialisation */
  STP_E_CUDA = -3,        /* CUDA runtime / driver error (text in stp_last_error) */
  STP_E_WORKSPACE = -4    /* workspace too small */
};

enum { STP_BF16 = 0, STP_F32 = 1, STP_U8 = 2 };

typedef void* stp_stream; /* cudaStream_t */

typedef struct stp_tensor {
  void* ptr;  /* element (0,0,0,0) */
  int32_t n, h, w, c;
  int32_t ld;    /* pixel stride in elements */
  int32_t dtype; /* STP_BF16 | STP_F32 | STP_U8 */
} stp_tensor;

int stp_version(void);
const char* stp_last_error(void);
/* number of kernels this library has enqueued from the calling process since load (bench `gpu_launches`) */
int64_t stp_launch_count(void);
/* 1 if the tcgen05/TMA conv path is compiled in and enabled, 0 if only the mma.sync path is used */
int stp_tc_enabled(void);
/* number of tcgen05/TMA kernels enqueued since load (evidence that the tensor-core path, not the mma.sync one, ran) */
int64_t stp_tc_launch_count(void);
/* ... of which launches of the tcgen05 cta_group::2 CTA-pair conv kernel (conv_tc3.cu) */
int64_t stp_tc3_launch_count(void);
void stp_set_tc_enabled(int on);
/* debugging / A-B knobs: "tc2_force_mt" (0 heuristic | 1,2,4,8), "tc_conv_version" (0 auto | 1 first-generation only),
 * "tc3" (0 auto | 1 off | 2 CTA-pair kernel wherever it serves the shape), "tc3_force_bn" (128 | 256), "tc3_force_mt" (1 | 2);
 * kernel selection of the 1x1 / narrow / strided layers (each with its measurement in csrc/api.cu dispatch_conv):
 * "tc2_1x1" (0 auto: 1x1 stride-1 convs on the tcgen05 halo kernel as plain GEMMs | 1 off | 2 only Cin % 64 == 0),
 * "gemm1x1" (0 auto: streaming mma.sync GEMM where the halo kernel does not tile the widths | 1 off | 2 every eligible 1x1),
 * "g1_bn" (1: no BatchNorm epilogues in that GEMM), "tc2_up2" (1: stride-2 dgrads stay on the first-generation kernel),
 * "wgrad1x1" (1: 1x1 weight gradients of untiled widths stay on the generic kernel), "nconv" (1: narrow-channel mma.sync conv,
 * an opt-in negative result), "bnb_fuse" (1: never fuse the BatchNorm-backward reduction into a dgrad epilogue) */
int stp_set_option(const char* name, int32_t value);
/* profiling aid: device buffer of >= 64 uint64 that the halo conv kernel's first and last thread blocks fill with
 * %globaltimer stamps of their phases (scripts/trace_conv.py); NULL turns it off */
void stp_set_trace_buffer(void* dev_ptr);

/* ------------------------------------------------------------------------------------------------
 * K1  augmentation  -- replaces imgaug.augmenters.{Fliplr,Flipud,Affine,Multiply,Add,Invert} and musket's Rotate90 run by
 *     musket_core.datasets.ImageKFoldedDataSet (segmentation.py:54; schemas/augmenters.raml:43-133).
 *     Sampling arithmetic == cv2.warpAffine fixed point, bit exact.
 * ---------------------------------------------------------------------------------------------- */
typedef struct stp_aug_spec {
  double fliplr_p, flipud_p;
  int32_t affine; /* 0/1 */
  double scale_lo, scale_hi;
  double tx_lo, tx_hi, ty_lo, ty_hi; /* translate_percent */
  double rot_lo, rot_hi;             /* degrees */
  double shear_lo, shear_hi;         /* degrees */
  int32_t has_mul;
  double mul_lo, mul_hi;
  int32_t has_add;
  int32_t add_lo, add_hi;
  int32_t mul_rint; /* 0: imgaug-0.3.0 truncating LUT, 1: round-half-even */
  int32_t rot90;    /* 0/1: musket `Rotate90: true` (reference ds_1.yaml:6): np.rot90 by a uniform k in {0,1,2,3}, applied FIRST;
                       square images only */
  double invert_p;  /* imgaug Invert(p): v -> 255 - v on the whole image with probability p */
  int32_t color_order[3]; /* order in which the colour stage applies 0 = Multiply, 1 = Add, 2 = Invert (imgaug Sequential
                             applies augmenters in YAML order; saturating uint8 ops do not commute) */
  int32_t flip_before_rot90; /* bit 0: Fliplr precedes Rotate90 in the YAML block, bit 1: Flipud does (reference
                                examples/people/ds_1.yaml:3-6 lists Fliplr, Flipud, Rotate90).  All three are index permutations:
                                a flip applied BEFORE an odd quarter turn equals the OTHER flip applied after it, so the kernel
                                keeps its rot90 -> flips gather and the draw swaps the flags. */
} stp_aug_spec;

typedef struct stp_aug_sample { /* per-sample drawn parameters, device resident, 128 bytes */
  double m[6];                 /* forward 2x3 affine (src->dst), row major */
  double inv[6];               /* cv2.warpAffine's fp64 inverse of m (dst->src) */
  int32_t fliplr, flipud;
  int32_t has_affine, has_mul;
  float mul;
  int32_t add;
  int32_t src_index;           /* which pool sample this output is drawn from */
  int32_t flags2;              /* bits 0-1: rot90 k; bit 2: invert; bits 4-9: colour order (three 2-bit op ids, first op lowest) */
} stp_aug_sample;

/* ---- uint8 resize with OpenCV's arithmetic (csrc/resize_u8.cu) -- replaces the `Resize -> shape` tail of the reference's
 * input pipeline (imgaug Resize = cv2.resize: INTER_CUBIC for images, INTER_NEAREST for masks; musket_core.datasets via
 * segmentation.py:54) and serves the crop / pad augmenters (schemas/augmenters.raml:72-87, 113-116).  Each of the n items
 * describes a stored image [sh][sw][c] at byte offset src_off of `d_arena` and a VIRTUAL window of it (origin vy0/vx0 may be
 * negative, size vh x vw may exceed the stored image: pixels outside read 0 = constant padding); the window is resized to
 * [h][w][c] and written to image i of d_dst [n][h][w][c].  A window that already is h x w is copied. */
enum { STP_RESIZE_NEAREST = 0, STP_RESIZE_CUBIC = 1 };
typedef struct stp_resize_item {
  int64_t src_off;
  int32_t sh, sw;
  int32_t vy0, vx0;
  int32_t vh, vw;
} stp_resize_item;
int stp_resize_u8(const uint8_t* d_arena, const stp_resize_item* d_items, int32_t n, int32_t c, uint8_t* d_dst, int32_t h,
                  int32_t w, int32_t mode, stp_stream stream);

/* Crop / pad augmenters (schemas/augmenters.raml:72-87, 113-116 -> imgaug 0.3.0 Pad / PadToFixedSize / CropToFixedSize /
 * CropAndPad [DEP]): each is a window of the sample, possibly extending beyond it (constant zero padding), that the pipeline's
 * final Resize brings back to `shape`.  The ops of the YAML block are composed into ONE window per sample (exact: none of them
 * resamples before the final Resize when `keep_size` ops come last), drawn on the device from Philox calls 6-7 of the
 * sample's stream, and written as stp_resize_item tables for stp_resize_u8 (cubic for the image, nearest for the mask).
 *   STP_CP_PAD            a,b,c,d = px (top, right, bottom, left)
 *   STP_CP_PAD_TO_FIXED   a = width, b = height; pads only a smaller side; left = floor((1 - u) * total)  [position uniform]
 *   STP_CP_CROP_TO_FIXED  a = width, b = height; crops only a larger side; left = floor(u * total)
 *   STP_CP_CROP_AND_PAD   percent: a,b,c,d = (top, right, bottom, left), or ranged = 1: a,b = range, one draw per side;
 *                         pixels = round(percent * size): positive pads, negative crops */
enum { STP_CP_PAD = 1, STP_CP_PAD_TO_FIXED = 2, STP_CP_CROP_TO_FIXED = 3, STP_CP_CROP_AND_PAD = 4, STP_CP_MAX_OPS = 4 };
typedef struct stp_croppad_op {
  int32_t kind, ranged;
  float a, b, c, d;
} stp_croppad_op;
typedef struct stp_croppad_spec {
  int32_t n_ops;
  stp_croppad_op ops[4];
} stp_croppad_spec;
int stp_croppad_draw(const stp_croppad_spec* h_spec, uint64_t seed, const int64_t* d_step, int32_t n, int32_t pool, int32_t h,
                     int32_t w, int32_t c_img, int32_t c_mask, stp_resize_item* d_img_items, stp_resize_item* d_mask_items,
                     stp_stream stream);

/* draw parameters: Philox4x32-10(key=seed, ctr=(step, sample_id, call, step>>32)).  `d_step` is a device
 * int64 so the call is CUDA-graph replayable; sample_id = (step*n + i) % pool (src_index likewise). */
int stp_augment_draw(const stp_aug_spec* h_spec, uint64_t seed, const int64_t* d_step, int32_t n,
                     int32_t pool, int32_t h, int32_t w, stp_aug_sample* d_out, stp_stream stream);
/* apply: img_pool u8 [pool,h,w,c_img], mask_pool u8 [pool,h,w,c_mask] -> img_out/mask_out u8 [n,h,w,*] */
int stp_augment_apply(const uint8_t* img_pool, const uint8_t* mask_pool, const stp_aug_sample* d_params,
                      uint8_t* img_out, uint8_t* mask_out, int32_t n, int32_t h, int32_t w,
                      int32_t c_img, int32_t c_mask, int32_t mul_rint, stp_stream stream);

/* Pixel-wise augmenters in YAML order, in place on the augmented batch (run AFTER stp_augment_apply called with mul_rint | 2,
 * which leaves the colour stage to this kernel): Multiply / Add / Invert use the per-sample draws of stp_augment_draw;
 * AddElementwise (a..b integers), MultiplyElementwise (a..b), Dropout (p ~ U(a, b) per image), AdditiveGaussianNoise
 * (scale ~ U(a, b) per image) draw per pixel -- per channel with probability `per_channel` per image; Grayscale blends
 * 0.299 R + 0.587 G + 0.114 B with alpha ~ U(a, b) per image.  group_size > 0: the op is member `group_member` of OneOf
 * group `group_id` and runs only for the samples that drew it.  imgaug 0.3.0 semantics as recalled [DEP]. */
enum { STP_PIX_MULTIPLY = 0, STP_PIX_ADD = 1, STP_PIX_INVERT = 2, STP_PIX_ADD_ELEMENTWISE = 3, STP_PIX_MULTIPLY_ELEMENTWISE = 4,
       STP_PIX_DROPOUT = 5, STP_PIX_GAUSSIAN_NOISE = 6, STP_PIX_GRAYSCALE = 7, STP_PIX_MAX_OPS = 8 };
typedef struct stp_aug_pix_op {
  int32_t kind;
  float per_channel;
  float a, b;
  int32_t group_id, group_size, group_member;
} stp_aug_pix_op;
typedef struct stp_aug_pix_spec {
  int32_t n_ops, mul_rint;
  stp_aug_pix_op ops[8];
  int32_t k_base;   /* position of ops[0] in the whole colour block (Philox calls 32 + k_base + i): a block interleaved with
                       neighbourhood augmenters runs as several launches */
} stp_aug_pix_spec;
int stp_augment_pixel_ops(uint8_t* d_img, const stp_aug_sample* d_params, const stp_aug_pix_spec* h_spec, uint64_t seed,
                          const int64_t* d_step, int32_t n, int32_t h, int32_t w, int32_t c_img, stp_stream stream);

/* Neighbourhood augmenters (schemas/augmenters.raml:97-112, 117-119): imgaug GaussianBlur (sigma ~ U(a, b)), AverageBlur
 * (k ~ integers a..b), MedianBlur (k ~ odd integers a..b <= 7), Sharpen (alpha ~ U(a, b), lightness ~ U(c, d)), Emboss (alpha,
 * strength), EdgeDetect (alpha), DirectedEdgeDetect (alpha, direction) with the arithmetic of the cv2 calls they make (cv2.GaussianBlur's 8.8 fixed point, cv2.blur,
 * cv2.medianBlur, cv2.filter2D; bit exact, see csrc/augment_nb.cu).  One op per call, d_src -> d_dst (different buffers),
 * images only.  k_index = position of the op in the colour block (Philox call 32 + k_index); group_* as in stp_aug_pix_op.
 * d_work: stp_augment_neighbourhood_workspace(n) bytes of device memory. */
enum { STP_NB_GAUSSIAN_BLUR = 0, STP_NB_AVERAGE_BLUR = 1, STP_NB_MEDIAN_BLUR = 2, STP_NB_SHARPEN = 3, STP_NB_EMBOSS = 4,
       STP_NB_EDGE_DETECT = 5, STP_NB_DIRECTED_EDGE_DETECT = 6 };
typedef struct stp_aug_nb_op {
  int32_t kind;
  float a, b, c, d;
  int32_t k_index;
  int32_t group_id, group_size, group_member;
  const void* d_table;   /* DirectedEdgeDetect (alpha ~ U(a, b), direction ~ U(c, d)): device float32 [360][9], the effect matrix per
                            integer degree (imgaug computes it with double-precision trigonometry; tabulated by the host) */
} stp_aug_nb_op;
size_t stp_augment_neighbourhood_workspace(int32_t n);
int stp_augment_neighbourhood(const uint8_t* d_src, uint8_t* d_dst, const stp_aug_sample* d_params, const stp_aug_nb_op* h_op,
                              uint64_t seed, const int64_t* d_step, int32_t n, int32_t h, int32_t w, int32_t c_img, void* d_work,
                              size_t work_bytes, stp_stream stream);

/* ------------------------------------------------------------------------------------------------
 * K2/K4/K5/K9  convolution -- replaces keras.layers.Conv2D / Conv2DTranspose -> TF Conv2D,
 *     Conv2DBackpropInput, Conv2DBackpropFilter (graph built by segmentation_models.Unet etc. at
 *     segmentation.py:109-113,155).
 * ---------------------------------------------------------------------------------------------- */
enum { STP_CONV_RELU = 1, STP_CONV_STATS = 2 };

typedef struct stp_conv_desc {
  int32_t r, s;         /* filter height, width */
  int32_t stride;       /* output stride */
  int32_t pad_h, pad_w; /* zero padding BEFORE (top/left); bottom/right implied by the output size */
  int32_t up;           /* zero-insertion factor on the input (1; >1 = transposed conv / strided dgrad) */
  int32_t flags;
} stp_conv_desc;

/* y = conv(x, w) [+ bias] [+ residual];  y bf16 or f32.  STP_CONV_STATS: also write per-channel partial
 * sums for BatchNorm into `stats_partial` (see stp_bn_stats_finalize). */
int stp_conv_fwd(const stp_conv_desc* d, const stp_tensor* x, const void* w_krsc, const float* bias,
                 const stp_tensor* residual, const stp_tensor* y, void* workspace, size_t workspace_bytes,
                 stp_stream stream);
/* stp_conv_fwd that ALSO produces the training-mode BatchNorm statistics of its output y (the layer that follows a
 * conv in every encoder/decoder block): sums are accumulated in the conv epilogue from the bf16 values it stores, the
 * last thread block finalises into coef / moving statistics exactly like stp_bn_stats_fused (which is what runs as a
 * second pass when no fused kernel serves the shape).  acc: 2*c zero-initialised device doubles, returned to zero. */
typedef struct stp_bn_fwd {
  float* partial;          /* as stp_bn_stats: 2*stp_bn_nblk(rows,c)*c floats (fallback pass) */
  uint32_t* sync;          /* zero-initialised ticket */
  double* acc;
  const float* gamma;      /* may be NULL */
  const float* beta;       /* may be NULL */
  float eps, momentum;
  float* moving_mean;      /* may be NULL */
  float* moving_var;
  float* coef;             /* out: f32 [4][c] mean, invstd, scale, shift */
} stp_bn_fwd;
int stp_conv_fwd_bn(const stp_conv_desc* d, const stp_tensor* x, const void* w_krsc, const float* bias,
                    const stp_tensor* residual, const stp_tensor* y, const stp_bn_fwd* h_bn, void* workspace,
                    size_t workspace_bytes, stp_stream stream);
/* dx = conv_transpose(dy, w) [+ residual]  for the FORWARD descriptor `d`; `w_dgrad` is the
 * [Cin][R][S][Cout] tap-flipped copy made by stp_weight_prep. */
int stp_conv_dgrad(const stp_conv_desc* d, const stp_tensor* dy, const void* w_dgrad,
                   const stp_tensor* residual, const stp_tensor* dx, void* workspace, size_t workspace_bytes,
                   stp_stream stream);
/* stp_conv_dgrad that ALSO performs the BatchNorm-backward reduction of the BatchNorm(+ReLU) layer whose OUTPUT is this
 * conv's input (keras BatchNormalization -> TF FusedBatchNormGrad reductions): with x = that layer's input and coef its
 * forward coefficients, dx is stored as g = dx*[x*scale+shift > 0] (relu != 0; masking is idempotent, stp_bn_bwd_apply may
 * mask again) and (sum g, sum g*xhat) -> dgamma, dbeta, bcoef exactly as stp_bn_bwd_reduce_fused -- one pass over (dx, x)
 * removed.  The dgrad must deliver the COMPLETE gradient of that tensor (no residual).  Shapes the halo kernel does not
 * serve run dgrad + stp_bn_bwd_reduce_fused. */
typedef struct stp_bn_bwd {
  const stp_tensor* x;  /* BatchNorm input (bf16, same pixels / channels as dx) */
  const float* coef;    /* f32 [4][c] mean, invstd, scale, shift (stp_bn_finalize / stp_conv_fwd_bn) */
  int32_t relu;
  float* partial;       /* as stp_bn_bwd_reduce_fused (fallback pass) */
  uint32_t* sync;       /* zero-initialised ticket */
  double* acc;          /* 2*c zero-initialised doubles, returned to zero */
  float* dgamma;        /* may be NULL */
  float* dbeta;         /* may be NULL */
  float* bcoef;         /* out: f32 [3][c] */
} stp_bn_bwd;
int stp_conv_dgrad_bn(const stp_conv_desc* d, const stp_tensor* dy, const void* w_dgrad, const stp_tensor* dx,
                      const stp_bn_bwd* h_bnb, void* workspace, size_t workspace_bytes, stp_stream stream);
/* dw[Cout][R][S][Cin] (f32) = sum_pixels dy (x) x ; deterministic split reduction through workspace */
int stp_conv_wgrad(const stp_conv_desc* d, const stp_tensor* x, const stp_tensor* dy, float* dw,
                   void* workspace, size_t workspace_bytes, stp_stream stream);
size_t stp_conv_wgrad_workspace(const stp_conv_desc* d, const stp_tensor* x, const stp_tensor* dy);
/* master f32 [Cout][R][S][Cin] -> bf16 same order (w_fwd) and bf16 [Cin][R][S][Cout] flipped (w_dgrad, may be NULL) */
int stp_weight_prep(const float* w_master, void* w_fwd, void* w_dgrad, int32_t cout, int32_t r, int32_t s,
                    int32_t cin, stp_stream stream);

/* stp_weight_prep for every conv layer of a network in one launch over the flat buffers.  d_items: device int64 [n][8] =
 * {element offset into the flat buffers, cout, r, s, cin, has_dgrad, first_tile, 0}; a layer owns
 * ceil(cout/32)*ceil(cin/32) consecutive tiles (each covering all r*s taps) starting at first_tile; total_tiles = their sum. */
int stp_weight_prep_batched(const float* flat_master, void* flat_fwd, void* flat_dgrad, const int64_t* d_items,
                            int32_t n_items, int64_t total_tiles, stp_stream stream);

/* small-Cout head (final_conv + sigmoid, classes<=4): CUDA-core, HBM bound.  logits/probabilities f32 [M,classes] */
/* workspace (stp_head_fwd_workspace bytes, 16-byte aligned) holds the bf16 zero-padded [16][3][3][Cin] weight copy of the
 * tcgen05 path; with workspace == NULL (or an unsupported shape) the CUDA-core kernel runs instead */
int stp_head_fwd(const stp_tensor* x, const float* w_krsc_f32, const float* bias, int32_t classes,
                 float* logits, void* workspace, size_t workspace_bytes, stp_stream stream);
size_t stp_head_fwd_workspace(const stp_tensor* x, int32_t classes);
/* dlogits f32 [M,classes] -> dx bf16, dw f32 [classes][3][3][Cin], dbias f32[classes] (workspace reduction) */
int stp_head_bwd(const stp_tensor* x, const float* w_krsc_f32, const float* dlogits, int32_t classes,
                 const stp_tensor* dx, float* dw, float* dbias, void* workspace, size_t workspace_bytes,
                 stp_stream stream);
size_t stp_head_bwd_workspace(const stp_tensor* x, int32_t classes);

/* ------------------------------------------------------------------------------------------------
 * K3/K6  BatchNormalization(+ReLU) -- replaces keras BatchNormalization -> TF FusedBatchNorm(+Grad)
 * ---------------------------------------------------------------------------------------------- */
#define STP_BN_MAX_PARTIALS 1024
/* number of partial blocks the stats / bwd-reduce kernels use for a [rows, c] tensor (pure host function) */
int32_t stp_bn_nblk(int64_t rows, int32_t c);
/* per-channel sum / sum of squares of x (bf16 or u8) -> partial[2][nblk][c] f32 */
int stp_bn_stats(const stp_tensor* x, float* partial, stp_stream stream);
/* partials -> mean, invstd, scale=gamma*invstd, shift=beta-mean*scale; updates moving stats
 * (momentum, unbiased variance) when moving_mean != NULL.  coef f32 [4][c] = mean, invstd, scale, shift */
int stp_bn_finalize(const float* partial, int32_t nblk, int32_t c, int64_t count, const float* gamma,
                    const float* beta, float eps, float momentum, float* moving_mean, float* moving_var,
                    float* coef, stp_stream stream);
/* stp_bn_stats + stp_bn_finalize in ONE launch: the last thread block to finish (device ticket in *sync, a
 * zero-initialised uint32 the kernel returns to zero) finalises.  acc == NULL: it reduces the per-block partials in a
 * fixed order (bitwise reproducible).  acc != NULL (2*c zero-initialised doubles, returned to zero): blocks add their
 * sums with double-precision atomics instead -- no serial reduction tail, many more blocks in flight; the summation
 * order then varies at the 1e-16 relative level of the double accumulators. */
int stp_bn_stats_fused(const stp_tensor* x, float* partial, uint32_t* sync, double* acc, const float* gamma,
                       const float* beta, float eps, float momentum, float* moving_mean, float* moving_var, float* coef,
                       stp_stream stream);
/* dbias[c] = sum over pixels of dz[.,c] (gradient of a conv bias; the conv+bias+ReLU layers of VGG-style encoders and of
 * decoders without BatchNorm).  Same one-launch reduction as stp_bn_stats_fused (partial / sync / acc as there). */
int stp_bias_grad(const stp_tensor* dz, float* partial, uint32_t* sync, double* acc, float* dbias, stp_stream stream);
/* y = [relu](x*scale+shift); up=2 writes each value to the 2x2 block of y (UpSampling2D fused) */
int stp_bn_apply(const stp_tensor* x, const float* coef, int32_t relu, int32_t up, const stp_tensor* y,
                 stp_stream stream);
/* inference-mode coefficients from moving stats */
int stp_bn_coef_infer(const float* gamma, const float* beta, const float* moving_mean, const float* moving_var,
                      float eps, int32_t c, float* coef, stp_stream stream);
/* backward: g = dy*(x*scale+shift>0 if relu); partial[2][nblk][c] = sum g, sum g*xhat.  pool=2: dy is the
 * gradient of the 2x-upsampled tensor (2x2 summed on the fly). */
int stp_bn_bwd_reduce(const stp_tensor* dy, const stp_tensor* x, const float* coef, int32_t relu,
                      int32_t pool, float* partial, stp_stream stream);
/* partials -> dgamma, dbeta (f32, written; may be NULL for scale=False) and bcoef f32 [3][c] (a,b,cc) with dx = a*g + b*x + cc */
int stp_bn_bwd_finalize(const float* partial, int32_t nblk, int32_t c, int64_t count, const float* coef,
                        float* dgamma, float* dbeta, float* bcoef, stp_stream stream);
/* stp_bn_bwd_reduce + stp_bn_bwd_finalize in ONE launch (same last-block scheme as stp_bn_stats_fused) */
int stp_bn_bwd_reduce_fused(const stp_tensor* dy, const stp_tensor* x, const float* coef, int32_t relu, int32_t pool,
                            float* partial, uint32_t* sync, double* acc, float* dgamma, float* dbeta, float* bcoef,
                            stp_stream stream);
/* dx = a*g + b*x + cc [+ residual] */
int stp_bn_bwd_apply(const stp_tensor* dy, const stp_tensor* x, const float* coef, const float* bcoef,
                     int32_t relu, int32_t pool, const stp_tensor* residual, const stp_tensor* dx,
                     stp_stream stream);
/* relu-only backward (VGG path / decoder without BN): dx = dy*(y>0) [+residual] */
int stp_relu_bwd(const stp_tensor* dy, const stp_tensor* y, int32_t pool, const stp_tensor* residual,
                 const stp_tensor* dx, stp_stream stream);

/* stem: u8 image -> bn_data (scale=False) normalised bf16 with C padded to 8; channel `c_img` is set to
 * 1.0 (the "ones" channel whose wgrad column yields d(beta of bn_data)) */
int stp_stem_prep(const uint8_t* img, int32_t n, int32_t h, int32_t w, int32_t c_img, const float* coef,
                  const stp_tensor* y, stp_stream stream);

/* Space-to-depth stem (DESIGN.md "stem"): with stp_stem_prep writing y as [n, h/2, w/2, 32] (channel ((h&1)*2+(w&1))*8+c)
 * the 7x7 stride-2 pad-3 stem convolution equals a 4x4 stride-1 convolution (pad 2 before) with the weights
 * w2[co][r'][s'][q*8+c] = w[co][2r'+dy-1][2s'+dx-1][c], q = dy*2+dx.  stp_stem_weight_s2d builds the bf16 w2 from the
 * f32 [cout][7][7][8] master; stp_stem_wgrad_s2d_gather maps the f32 gradient of w2 back to [cout][7][7][8]. */
int stp_stem_weight_s2d(const float* w_master, void* w2, int32_t cout, stp_stream stream);
int stp_stem_wgrad_s2d_gather(const float* dw2, float* dw, int32_t cout, stp_stream stream);

/* after the stem wgrad (dw8 f32 [cout][r][s][cin_pad]): dbeta(bn_data)[c<c_img] from the ones-channel column,
 * then zero the padded columns c>=c_img in place (DESIGN.md "stem") */
int stp_stem_wgrad_post(float* dw8, const float* w_master, int32_t cout, int32_t r, int32_t s, int32_t cin_pad,
                        int32_t c_img, float* dbeta, stp_stream stream);

/* ------------------------------------------------------------------------------------------------
 * K7  MaxPooling2D -- replaces keras MaxPooling2D -> TF MaxPool / MaxPoolGrad
 * ---------------------------------------------------------------------------------------------- */
int stp_maxpool_fwd(const stp_tensor* x, int32_t k, int32_t stride, int32_t pad, const stp_tensor* y,
                    uint8_t* argmax, stp_stream stream);
int stp_maxpool_bwd(const stp_tensor* dy, const uint8_t* argmax, int32_t k, int32_t stride, int32_t pad,
                    const stp_tensor* residual, const stp_tensor* dx, stp_stream stream);
/* AveragePooling2D(k, strides k) over windows that tile the input exactly (PSPNet pyramid pooling, schema
 * segmentation.raml:226-248 -> segmentation_models PSPNet [DEP]); backward: dx = dy / k^2 per window (+ residual). */
int stp_avgpool_fwd(const stp_tensor* x, int32_t k, const stp_tensor* y, stp_stream stream);
int stp_avgpool_bwd(const stp_tensor* dy, int32_t k, const stp_tensor* residual, const stp_tensor* dx, stp_stream stream);

/* K8 copy / nearest-upsample a tensor into a (strided) destination: UpSampling2D + Concatenate when the
 * producer could not write in place */
int stp_copy_up(const stp_tensor* x, int32_t up, const stp_tensor* y, stp_stream stream);
/* y = a + b (Add layer; Linknet / FPN) */
int stp_add(const stp_tensor* a, const stp_tensor* b, const stp_tensor* y, stp_stream stream);
/* gradient of UpSampling2D(2, nearest) for a tensor that is NOT post-ReLU (the FPN top-down pathway, where the upsampled
 * tensor is a linear 1x1-conv output): dx = 2x2 sum of dy [+ residual] */
int stp_upsample2x_bwd(const stp_tensor* dy, const stp_tensor* residual, const stp_tensor* dx, stp_stream stream);

/* K8b bilinear resize -- replaces keras UpSampling2D(interpolation='bilinear') -> tf.image.resize_bilinear (TF1 legacy:
 * src = dst*in/out, no half-pixel offset, align_corners=False) in the FPN decoder built by segmentation_models.FPN
 * (reference segmentation.py:109-113; schema segmentation.raml:179-204).  fwd: bf16 -> bf16 (equal channel counts; y may
 * be a channel slice of a concat buffer) or f32 -> f32 (the first y.c channels of x: padded head logits -> dense
 * [pixels][classes]).  bwd (up-scaling only, deterministic gather): bf16 dy -> bf16 dx [+ residual], or f32 dy ->
 * bf16 dx whose channels >= dy.c are written as zero. */
int stp_resize_bilinear_fwd(const stp_tensor* x, const stp_tensor* y, stp_stream stream);
int stp_resize_bilinear_bwd(const stp_tensor* dy, const stp_tensor* residual, const stp_tensor* dx, stp_stream stream);

/* ------------------------------------------------------------------------------------------------
 * K11/K14  loss + metrics -- replaces keras binary_crossentropy, musket_core.losses.{dice,iou,...}
 *     (registered at segmentation.py:15-22).  One pass over (logits f32 [M], mask u8 [M]).
 * ---------------------------------------------------------------------------------------------- */
typedef struct stp_loss_spec {
  float w_bce, w_dice, w_iou; /* composite loss weights: w_bce*binary_crossentropy + w_dice*dice_loss + w_iou*iou_loss */
  float w_jaccard, w_focal;   /* + w_jaccard*jaccard_loss (smooth 100) + w_focal*focal_loss (gamma 2, alpha .75) */
} stp_loss_spec;
enum { /* indices into the f32 result vector (16 floats) */
  STP_L_LOSS = 0, STP_L_BCE = 1, STP_L_DICE = 2, STP_L_IOU = 3, STP_L_ACC = 4, STP_L_IOT = 5,
  STP_L_SUM_P = 6, STP_L_SUM_T = 7, STP_L_SUM_PT = 8, STP_L_COUNT = 9, STP_L_LOVASZ = 10, STP_L_JACCARD = 11, STP_L_FOCAL = 12,
  STP_L_CCE = 13, STP_L_CACC = 14 /* categorical_crossentropy / categorical accuracy (stp_softmax_cce_fwd) */
};
int stp_loss_fwd(const float* logits, const uint8_t* mask, int64_t count, const stp_loss_spec* h_spec,
                 float* partial, float* result16, stp_stream stream);
size_t stp_loss_partial_floats(void);
/* dlogits f32 [count] = dL/dlogit using the sums in result16 */
int stp_loss_bwd(const float* logits, const uint8_t* mask, int64_t count, const stp_loss_spec* h_spec,
                 const float* result16, float* dlogits, stp_stream stream);

/* `activation: softmax` + `loss: categorical_crossentropy` (schema segmentation.raml:12-21, 62-63): keras categorical_crossentropy
 * on probabilities [DEP]: p = softmax(logits); p /= sum(p); clip(p, 1e-7, 1-1e-7); -sum_c t_c log p_c; mean over pixels.
 * logits f32 [pixels][classes], mask u8 [pixels][classes] (one-hot or multi-hot), 2 <= classes <= 4.  fwd writes
 * result16[STP_L_CCE], result16[STP_L_CACC] (argmax agreement) and result16[STP_L_LOSS] (+)= weight * cce (run it AFTER
 * stp_loss_fwd, which fills the other slots); partial: stp_loss_partial_floats() floats.  bwd writes (accumulate: adds)
 * weight * dL/dlogit. */
int stp_softmax_cce_fwd(const float* logits, const uint8_t* mask, int64_t pixels, int32_t classes, float weight,
                        int32_t accumulate, float* partial, float* result16, stp_stream stream);
int stp_softmax_cce_bwd(const float* logits, const uint8_t* mask, int64_t pixels, int32_t classes, float weight,
                        int32_t accumulate, float* dlogits, stp_stream stream);

/* Lovasz hinge (binary, per image, on LOGITS; musket_core.losses.lovasz_loss -- the reference strips the trailing
 * Activation when this loss is compiled).  act_elu: 1 = elu(e)+1 (Kaggle-TGS variant), 0 = relu (Berman).  fwd sorts the
 * batch's hinge errors (one stable radix sort, images kept contiguous), writes result16[STP_L_LOVASZ] = mean over images
 * and result16[STP_L_LOSS] (+)= weight * it, and leaves the per-element gradient in the workspace; bwd scatters
 * weight * dL/dlogit into dlogits (accumulate: += ).  logits f32 [images*pixels_per_image], mask u8 same order. */
size_t stp_lovasz_workspace(int32_t images, int64_t pixels_per_image);
int stp_lovasz_fwd(const float* logits, const uint8_t* mask, int32_t images, int64_t pixels_per_image, int32_t act_elu,
                   float weight, int32_t accumulate, void* workspace, size_t workspace_bytes, float* result16,
                   stp_stream stream);
/* `classes` > 1: logits / mask are [image][pixel][class]; one hinge per (image, class), mean over all of them (the
 * reference's K.squeeze(...,-1) is undefined for classes > 1 -- SURVEY.md 8 a-6 -- this is the definition the oracle
 * uses).  Size the workspace with stp_lovasz_workspace(images*classes, pixels_per_image) and call stp_lovasz_bwd with
 * images*classes. */
int stp_lovasz_fwd_mc(const float* logits, const uint8_t* mask, int32_t images, int64_t pixels_per_image, int32_t classes,
                      int32_t act_elu, float weight, int32_t accumulate, void* workspace, size_t workspace_bytes,
                      float* result16, stp_stream stream);
int stp_lovasz_bwd(const void* workspace, size_t workspace_bytes, int32_t images, int64_t pixels_per_image, float weight,
                   int32_t accumulate, float* dlogits, stp_stream stream);

/* ------------------------------------------------------------------------------------------------
 * K12  optimizer -- replaces keras.optimizers.Adam/SGD/RMSprop update ops (segmentation.raml:77-89)
 *     Keras formulation (eps outside the bias correction).  `d_step` device int64, incremented by
 *     stp_step_advance once per iteration (graph replayable).
 * ---------------------------------------------------------------------------------------------- */
typedef struct stp_grad_xform { /* g' = clipvalue(clipnorm(g * scale)) -- keras clipnorm / clipvalue, DDP mean */
  float scale;           /* e.g. 1/world_size */
  float clipnorm;        /* <=0: off.  needs d_sumsq = sum(g^2) of the UNSCALED flat gradient */
  float clipvalue;       /* <=0: off */
  const float* d_sumsq;  /* device scalar or NULL */
  const float* d_lr_scale; /* device scalar multiplying lr, or NULL: ReduceLROnPlateau / CyclicLR (callbacks.raml:22-48)
                              change the rate between replays of the captured step without re-capturing it */
} stp_grad_xform;
int stp_adam(float* p, const float* g, float* m, float* v, int64_t count, float lr, float beta1, float beta2,
             float eps, const stp_grad_xform* h_gx, const int64_t* d_step, stp_stream stream);
int stp_sgd(float* p, const float* g, float* v, int64_t count, float lr, float momentum, int32_t nesterov,
            const stp_grad_xform* h_gx, stp_stream stream);
int stp_rmsprop(float* p, const float* g, float* a, int64_t count, float lr, float rho, float eps,
                const stp_grad_xform* h_gx, stp_stream stream);
/* keras Nadam (schedule_decay 0.004): sched5 = device float[5], sched5[0] initialised to 1.0 (running momentum-schedule
 * product), the rest scratch */
int stp_nadam(float* p, const float* g, float* m, float* v, float* sched5, int64_t count, float lr, float beta1,
              float beta2, float eps, float schedule_decay, const stp_grad_xform* h_gx, const int64_t* d_step,
              stp_stream stream);
int stp_step_advance(int64_t* d_step, stp_stream stream);
/* sum of squares of g into out[0] (f32, deterministic two-pass through `partial`), for clipnorm */
int stp_sumsq(const float* g, int64_t count, float* partial, float* out, stp_stream stream);

/* ----------------------------------------------------------------------------------------------
 * K14  depthwise convolution (keras DepthwiseConv2D, depth_multiplier 1) of the reference's in-tree DeepLabV3+ /
 *      MobileNetV2 (impl/deeplab/model.py:236-275 `_inverted_res_block`, :104-142 `SepConv_BN`): k x k filter per channel,
 *      stride, atrous rate.  x / y / dy / dx bf16 NHWC (c % 8 == 0); weights bf16 [k][k][c] (= the keras (k,k,c,1) kernel;
 *      stp_weight_prep(master, w, NULL, 1, k, k, c) makes the copy); dw f32 [k][k][c].  pad_* = padding BEFORE; the output size implies the rest
 *      (TF 'same' with stride 2 on an even size pads 0 before / 1 after).  wgrad: per-block partials + fixed-order reduction.
 * ---------------------------------------------------------------------------------------------- */
typedef struct stp_dwconv_desc {
  int32_t k, stride, dilation, pad_h, pad_w;
} stp_dwconv_desc;
int stp_dwconv_fwd(const stp_dwconv_desc* d, const stp_tensor* x, const void* w_kkc, const stp_tensor* y, stp_stream stream);
int stp_dwconv_dgrad(const stp_dwconv_desc* d, const stp_tensor* dy, const void* w_kkc, const stp_tensor* residual,
                     const stp_tensor* dx, stp_stream stream);
int stp_dwconv_wgrad(const stp_dwconv_desc* d, const stp_tensor* x, const stp_tensor* dy, float* dw_kkc, void* workspace,
                     size_t workspace_bytes, stp_stream stream);
size_t stp_dwconv_wgrad_workspace(const stp_dwconv_desc* d, const stp_tensor* x, const stp_tensor* dy);

/* ----------------------------------------------------------------------------------------------
 * K15  the other pieces of the DeepLabV3+ graph (impl/deeplab/model.py): whole-map mean + broadcast (image-pooling branch,
 *      :462-469), training-mode Dropout (:486) and the probability head (:494-500: activation at 1/8 resolution, THEN
 *      align_corners bilinear resize to the input size).
 *      stp_prob_head_fwd: z f32 [n,h,w,>=classes] (1x1 conv output) -> logits f32 dense [n,H,W,classes] such that
 *      activation(logits) == resize(activation(z)): the loss / predict kernels consume it unchanged.  activation 1 sigmoid,
 *      2 softmax.  stp_prob_head_bwd: gradient wrt those logits -> dz bf16 [n,h,w,dz.c] (channels >= classes zero).
 *      stp_dropout: y = x * keep / (1-rate); the mask is a function of (seed, salt, *d_step, element), so the backward is the
 *      same call on the gradient.  x == y allowed.
 * ---------------------------------------------------------------------------------------------- */
int stp_global_avgpool_fwd(const stp_tensor* x, const stp_tensor* y, stp_stream stream);
int stp_global_avgpool_bwd(const stp_tensor* dy, const stp_tensor* residual, const stp_tensor* dx, stp_stream stream);
int stp_broadcast_fwd(const stp_tensor* x, const stp_tensor* y, stp_stream stream);
int stp_broadcast_bwd(const stp_tensor* dy, const stp_tensor* dx, stp_stream stream);
int stp_dropout(const stp_tensor* x, float rate, uint64_t seed, uint32_t salt, const int64_t* d_step, const stp_tensor* y,
                stp_stream stream);
/* stp_resize_bilinear_fwd/bwd with tf.image.resize_bilinear(align_corners=True) index arithmetic: the BilinearUpsampling layer of
 * impl/deeplab/model.py:81-100 on feature maps (xception decoder, :488-490) */
int stp_resize_bilinear_ac_fwd(const stp_tensor* x, const stp_tensor* y, stp_stream stream);
int stp_resize_bilinear_ac_bwd(const stp_tensor* dy, const stp_tensor* residual, const stp_tensor* dx, stp_stream stream);
int stp_prob_head_fwd(const stp_tensor* z, int32_t classes, int32_t activation, const stp_tensor* logits, stp_stream stream);
int stp_prob_head_bwd(const stp_tensor* dlogits, const stp_tensor* logits, const stp_tensor* z, int32_t classes,
                      int32_t activation, const stp_tensor* dz, stp_stream stream);

/* ------------------------------------------------------------------------------------------------
 * K13  data-parallel exchange -- the one collective of the path: SUM all-reduce of the flat fp32 gradient over the GPUs of a
 *      box (replaces keras.utils.multi_gpu_model's CPU-side merge, reference FAQ.md:108-112).  Thin wrappers over NCCL, bound
 *      at run time (dlopen): one process per GPU; rank 0 creates the 128-byte id with stp_comm_unique_id and hands it to the
 *      other ranks out of band; every rank calls stp_comm_init AFTER selecting its device; stp_allreduce is asynchronous on
 *      `stream` (in place, count floats).  STP_E_UNSUPPORTED when libnccl.so.2 cannot be loaded.
 * ---------------------------------------------------------------------------------------------- */
typedef struct stp_comm stp_comm;
int stp_comm_unique_id(void* out128);
int stp_comm_init(int32_t world, int32_t rank, const void* id128, stp_comm** out);
int stp_allreduce(stp_comm* comm, float* d_buf, int64_t count, stp_stream stream);
int stp_comm_destroy(stp_comm* comm);

#ifdef __cplusplus
}
#endif
#endif /* STP_H_ */
