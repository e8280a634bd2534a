// uint8 image / mask resize on the device with OpenCV's arithmetic: the `Resize -> shape` step that the reference's input
// pipeline ends with (imgaug Resize = cv2.resize, INTER_CUBIC for images, INTER_NEAREST for segmentation maps; SURVEY.md 8f
// row N3) and the crop / pad family of augmenters (schemas/augmenters.raml:72-87, 113-116: Pad, PadToFixedSize,
// CropToFixedSize, CropAndPad), all of which are "take a (possibly zero-padded) window of the sample and resize it to `shape`".
//
// Arithmetic = cv::resize's reference (scalar) path, imgproc/src/resize.cpp:
//   cubic  : fx = (float)((dx + 0.5) * scale - 0.5); sx = floor(fx); fx -= sx; interpolateCubic (A = -0.75, fp32, unfused);
//            weights -> short by cvRound(w * 2048); horizontal pass exact in int32; vertical pass int32;
//            result (v + 2^21) >> 22, saturated; taps outside the image replicate the border pixel.
//   nearest: sx = min(floor(dx * (1 / (dst / src))), src - 1)          (double arithmetic)
// Integer / byte work, HBM bound: one thread per output pixel, all channels.  Pinned against the real cv2 (4.13, IPP off)
// in tests/test_gpu_resize.py: identical except at rounding ties where OpenCV's own SIMD vertical pass (fp32 FMA chain,
// round-half-even) disagrees with its scalar path (integer, round-half-up) -- < 0.05 % of pixels, |d| = 1.
#include "common.cuh"

namespace stp {

struct ResizeItem {      // == stp_resize_item (include/stp.h)
  int64_t src_off;       // byte offset of the stored image in the source arena
  int32_t sh, sw;        // stored image size (rows, columns)
  int32_t vy0, vx0;      // origin of the VIRTUAL image inside the stored one (negative: zero padding before it)
  int32_t vh, vw;        // virtual image size: pixels outside the stored image read as 0 (constant padding)
};
static_assert(sizeof(ResizeItem) == 32, "stp_resize_item layout");

__device__ __forceinline__ void cubic_weights(float x, int* w) {
  const float A = -0.75f;
  const float x1 = __fadd_rn(x, 1.f), xm = __fsub_rn(1.f, x);
  const float c0 = __fsub_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fsub_rn(__fmul_rn(A, x1), 5.f * A), x1), 8.f * A), x1), 4.f * A);
  const float c1 = __fadd_rn(__fmul_rn(__fmul_rn(__fsub_rn(__fmul_rn(A + 2.f, x), A + 3.f), x), x), 1.f);
  const float c2 = __fadd_rn(__fmul_rn(__fmul_rn(__fsub_rn(__fmul_rn(A + 2.f, xm), A + 3.f), xm), xm), 1.f);
  const float c3 = __fsub_rn(__fsub_rn(__fsub_rn(1.f, c0), c1), c2);
  const float c[4] = {c0, c1, c2, c3};
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    int v = __float2int_rn(__fmul_rn(c[k], 2048.f));   // saturate_cast<short>(cvRound(.))
    w[k] = v < -32768 ? -32768 : (v > 32767 ? 32767 : v);
  }
}
__device__ __forceinline__ void cubic_axis(int d, int src, int dst, int* s0, int* w) {
  const double scale = __ddiv_rn((double)src, (double)dst);
  float f = (float)__dsub_rn(__dmul_rn(__dadd_rn((double)d, 0.5), scale), 0.5);
  const int s = (int)floorf(f);
  f = __fsub_rn(f, (float)s);
  *s0 = s;
  cubic_weights(f, w);
}

template <int C>
__global__ void __launch_bounds__(256) resize_u8_kernel(const uint8_t* __restrict__ arena, const ResizeItem* __restrict__ items,
                                                        uint8_t* __restrict__ dst, int H, int W, int mode) {
  const ResizeItem it = items[blockIdx.z];
  const uint8_t* src = arena + it.src_off;
  uint8_t* out = dst + (int64_t)blockIdx.z * H * W * C;
  const int64_t total = (int64_t)H * W;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int y = (int)(i / W), x = (int)(i - (int64_t)y * W);
    uint8_t o[C];
    auto fetch = [&](int vy, int vx, int c) -> int {   // virtual image -> stored image, zero outside
      const int sy = vy + it.vy0, sx = vx + it.vx0;
      if (sy < 0 || sy >= it.sh || sx < 0 || sx >= it.sw) return 0;
      return src[((int64_t)sy * it.sw + sx) * C + c];
    };
    if (it.vh == H && it.vw == W) {                    // same size: cv::resize copies
#pragma unroll
      for (int c = 0; c < C; ++c) o[c] = (uint8_t)fetch(y, x, c);
    } else if (mode == 0) {                            // INTER_NEAREST
      const double ify = __ddiv_rn(1.0, __ddiv_rn((double)H, (double)it.vh)), ifx = __ddiv_rn(1.0, __ddiv_rn((double)W, (double)it.vw));
      int sy = (int)floor(__dmul_rn((double)y, ify)), sx = (int)floor(__dmul_rn((double)x, ifx));
      sy = sy < it.vh - 1 ? sy : it.vh - 1;
      sx = sx < it.vw - 1 ? sx : it.vw - 1;
#pragma unroll
      for (int c = 0; c < C; ++c) o[c] = (uint8_t)fetch(sy, sx, c);
    } else {                                           // INTER_CUBIC
      int sx0, sy0, wx[4], wy[4];
      cubic_axis(x, it.vw, W, &sx0, wx);
      cubic_axis(y, it.vh, H, &sy0, wy);
      int acc[C];
#pragma unroll
      for (int c = 0; c < C; ++c) acc[c] = 0;
#pragma unroll
      for (int ky = 0; ky < 4; ++ky) {
        int vy = sy0 - 1 + ky;
        vy = vy < 0 ? 0 : (vy >= it.vh ? it.vh - 1 : vy);
        int row[C];
#pragma unroll
        for (int c = 0; c < C; ++c) row[c] = 0;
#pragma unroll
        for (int kx = 0; kx < 4; ++kx) {
          int vx = sx0 - 1 + kx;
          vx = vx < 0 ? 0 : (vx >= it.vw ? it.vw - 1 : vx);
#pragma unroll
          for (int c = 0; c < C; ++c) row[c] += fetch(vy, vx, c) * wx[kx];
        }
#pragma unroll
        for (int c = 0; c < C; ++c) acc[c] += row[c] * wy[ky];
      }
#pragma unroll
      for (int c = 0; c < C; ++c) {
        const int v = (acc[c] + (1 << 21)) >> 22;
        o[c] = (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
      }
    }
#pragma unroll
    for (int c = 0; c < C; ++c) out[i * C + c] = o[c];
  }
}

// ---- crop / pad augmenters: per-sample window of the pool sample, drawn on the device (CUDA-graph replayable) ----------
// Philox4x32-10 twin of augment.cu / oracle/philox.py (counter = (step, sample id, call, step >> 32), key = seed); the crop /
// pad family uses calls 6 and 7 of a sample's stream (calls 0-5 belong to stp_augment_draw).
__device__ __forceinline__ void philox_cp(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t* out) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
    const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
__device__ __forceinline__ double u53_cp(uint32_t hi, uint32_t lo) {
  return (double)((((uint64_t)hi << 32) | lo) >> 11) * 1.1102230246251565e-16;
}

__global__ void croppad_draw_kernel(stp_croppad_spec spec, uint64_t seed, const int64_t* __restrict__ d_step, int n, int pool, int H,
                                    int W, int c_img, int c_mask, ResizeItem* __restrict__ img_items, ResizeItem* __restrict__ mask_items) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int64_t step = *d_step;
  const uint32_t sid = (uint32_t)((step * n + i) % pool);
  uint32_t r6[4], r7[4];
  philox_cp((uint32_t)step, sid, 6u, (uint32_t)(step >> 32), (uint32_t)seed, (uint32_t)(seed >> 32), r6);
  philox_cp((uint32_t)step, sid, 7u, (uint32_t)(step >> 32), (uint32_t)seed, (uint32_t)(seed >> 32), r7);
  const double u[4] = {u53_cp(r6[0], r6[1]), u53_cp(r6[2], r6[3]), u53_cp(r7[0], r7[1]), u53_cp(r7[2], r7[3])};
  int vy0 = 0, vx0 = 0, vh = H, vw = W;
  for (int k = 0; k < spec.n_ops; ++k) {
    const stp_croppad_op op = spec.ops[k];
    if (op.kind == STP_CP_PAD) {                       // px = (top, right, bottom, left)
      vy0 -= (int)op.a; vh += (int)op.a + (int)op.c;
      vx0 -= (int)op.d; vw += (int)op.d + (int)op.b;
    } else if (op.kind == STP_CP_PAD_TO_FIXED) {       // a = width, b = height; position uniform: u[0] (x), u[1] (y)
      if (vw < (int)op.a) { const int tot = (int)op.a - vw; vx0 -= (int)floor(__dmul_rn(__dsub_rn(1.0, u[0]), (double)tot)); vw = (int)op.a; }
      if (vh < (int)op.b) { const int tot = (int)op.b - vh; vy0 -= (int)floor(__dmul_rn(__dsub_rn(1.0, u[1]), (double)tot)); vh = (int)op.b; }
    } else if (op.kind == STP_CP_CROP_TO_FIXED) {      // a = width, b = height; position uniform: u[2] (x), u[3] (y)
      if (vw > (int)op.a) { const int tot = vw - (int)op.a; vx0 += (int)floor(__dmul_rn(u[2], (double)tot)); vw = (int)op.a; }
      if (vh > (int)op.b) { const int tot = vh - (int)op.b; vy0 += (int)floor(__dmul_rn(u[3], (double)tot)); vh = (int)op.b; }
    } else if (op.kind == STP_CP_CROP_AND_PAD) {       // percent per side (top, right, bottom, left); ranged: one draw per side
      double pt = op.a, pr = op.b, pb = op.c, pl = op.d;
      if (op.ranged) {
        pt = __dadd_rn((double)op.a, __dmul_rn(u[0], __dsub_rn((double)op.b, (double)op.a)));
        pr = __dadd_rn((double)op.a, __dmul_rn(u[1], __dsub_rn((double)op.b, (double)op.a)));
        pb = __dadd_rn((double)op.a, __dmul_rn(u[2], __dsub_rn((double)op.b, (double)op.a)));
        pl = __dadd_rn((double)op.a, __dmul_rn(u[3], __dsub_rn((double)op.b, (double)op.a)));
      }
      const int t = __double2int_rn(__dmul_rn(pt, (double)vh)), b = __double2int_rn(__dmul_rn(pb, (double)vh));
      const int l = __double2int_rn(__dmul_rn(pl, (double)vw)), rr = __double2int_rn(__dmul_rn(pr, (double)vw));
      int nh = vh + t + b, nw = vw + l + rr;          // positive = pad, negative = crop
      if (nh >= 1 && nw >= 1) { vy0 -= t; vx0 -= l; vh = nh; vw = nw; }
    }
  }
  ResizeItem it;
  it.sh = H; it.sw = W; it.vy0 = vy0; it.vx0 = vx0; it.vh = vh; it.vw = vw;
  it.src_off = (int64_t)sid * H * W * c_img;
  img_items[i] = it;
  if (mask_items) {
    it.src_off = (int64_t)sid * H * W * c_mask;
    mask_items[i] = it;
  }
}

}  // namespace stp

using namespace stp;

extern "C" int stp_croppad_draw(const stp_croppad_spec* h_spec, uint64_t seed, const int64_t* d_step, int32_t n, int32_t pool,
                                int32_t h, int32_t w, int32_t c_img, int32_t c_mask, stp_resize_item* d_img_items,
                                stp_resize_item* d_mask_items, stp_stream stream) {
  STP_REQUIRE(h_spec && d_step && d_img_items && n > 0 && pool > 0 && h > 0 && w > 0, "croppad_draw: bad args");
  STP_REQUIRE(h_spec->n_ops >= 0 && h_spec->n_ops <= STP_CP_MAX_OPS, "croppad_draw: at most %d ops", STP_CP_MAX_OPS);
  for (int k = 0; k < h_spec->n_ops; ++k)
    STP_REQUIRE(h_spec->ops[k].kind >= STP_CP_PAD && h_spec->ops[k].kind <= STP_CP_CROP_AND_PAD, "croppad_draw: unknown op kind");
  croppad_draw_kernel<<<(n + 63) / 64, 64, 0, (cudaStream_t)stream>>>(*h_spec, seed, d_step, n, pool, h, w, c_img, c_mask,
                                                                      (ResizeItem*)d_img_items, (ResizeItem*)d_mask_items);
  return check_launch("croppad_draw");
}

extern "C" int stp_resize_u8(const uint8_t* d_arena, const stp_resize_item* d_items, int32_t n, int32_t c, uint8_t* d_dst,
                             int32_t h, int32_t w, int32_t mode, stp_stream stream) {
  STP_REQUIRE(d_arena && d_items && d_dst && n > 0 && h > 0 && w > 0, "resize_u8: bad args");
  STP_REQUIRE(c >= 1 && c <= 4, "resize_u8: 1 <= channels <= 4");
  STP_REQUIRE(mode == STP_RESIZE_NEAREST || mode == STP_RESIZE_CUBIC, "resize_u8: mode must be STP_RESIZE_NEAREST or STP_RESIZE_CUBIC");
  STP_REQUIRE(n <= 65535, "resize_u8: at most 65535 images per call");
  const int64_t total = (int64_t)h * w;
  int64_t nb = (total + 255) / 256;
  if (nb > kNumSMs * 8) nb = kNumSMs * 8;
  dim3 grid((unsigned)nb, 1, (unsigned)n);
  cudaStream_t st = (cudaStream_t)stream;
  const ResizeItem* it = (const ResizeItem*)d_items;
  if (c == 1) resize_u8_kernel<1><<<grid, 256, 0, st>>>(d_arena, it, d_dst, h, w, mode);
  else if (c == 2) resize_u8_kernel<2><<<grid, 256, 0, st>>>(d_arena, it, d_dst, h, w, mode);
  else if (c == 3) resize_u8_kernel<3><<<grid, 256, 0, st>>>(d_arena, it, d_dst, h, w, mode);
  else resize_u8_kernel<4><<<grid, 256, 0, st>>>(d_arena, it, d_dst, h, w, mode);
  return check_launch("resize_u8");
}
