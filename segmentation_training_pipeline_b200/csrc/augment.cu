// K1: on-device augmentation of uint8 image + mask pairs: Fliplr, Flipud, Affine, Multiply, Add fused in ONE
// gather pass (imgaug 0.3.0 Sequential semantics, reference schemas/augmenters.raml:43-133).
// The affine sampling is cv2.warpAffine's fixed-point rule: fp64 rint of the
// per-row / per-column terms (unfused mul/add, round-half-even), then pure integer math per pixel:
// 1/32-pixel coordinates, 15-bit bilinear weights, (sum + 16384) >> 15.  Bit exact against cv2 4.13.
// Integer/byte work, HBM/L2 bound: 4 output pixels per thread, 32-bit packed stores.
#include "common.cuh"

namespace stp {

// ---- Philox4x32-10 (twin of oracle/philox.py) ------------------------------------------------
__device__ __forceinline__ void philox4x32(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                           uint32_t k1, uint32_t* out) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
__device__ __forceinline__ double u53(uint32_t hi, uint32_t lo) {
  return (double)((((uint64_t)hi << 32) | lo) >> 11) * 1.1102230246251565e-16;  // 2^-53, exact
}
__device__ __forceinline__ double lerp_rn(double lo, double hi, double u) {
  return __dadd_rn(lo, __dmul_rn(u, __dsub_rn(hi, lo)));
}

struct DevSample {  // == stp_aug_sample (include/stp.h)
  double m[6];
  double inv[6];
  int32_t fliplr, flipud, has_affine, has_mul;
  float mul;
  int32_t add, src_index, flags2;  // flags2: bits 0-1 rot90 k, bit 2 invert, bits 4-9 colour order
};
static_assert(sizeof(DevSample) == sizeof(stp_aug_sample), "stp_aug_sample layout");

__global__ void augment_draw_kernel(stp_aug_spec spec, uint64_t seed, const int64_t* __restrict__ d_step, int n,
                                    int pool, int H, int W, DevSample* __restrict__ out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int64_t step = *d_step;
  const uint32_t sid = (uint32_t)((step * n + i) % pool);
  const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
  uint32_t r[6][4];
#pragma unroll
  for (int call = 0; call < 6; ++call) philox4x32((uint32_t)step, sid, call, (uint32_t)(step >> 32), k0, k1, r[call]);
  const double u_lr = u53(r[0][0], r[0][1]), u_ud = u53(r[0][2], r[0][3]);
  const double u_sc = u53(r[1][0], r[1][1]), u_rot = u53(r[1][2], r[1][3]);
  const double u_sh = u53(r[2][0], r[2][1]), u_tx = u53(r[2][2], r[2][3]);
  const double u_ty = u53(r[3][0], r[3][1]), u_mul = u53(r[3][2], r[3][3]);
  const double u_add = u53(r[4][0], r[4][1]);
  DevSample s;
  s.fliplr = u_lr < (double)spec.fliplr_p;
  s.flipud = u_ud < (double)spec.flipud_p;
  s.has_affine = spec.affine;
  s.m[0] = 1.0; s.m[1] = 0.0; s.m[2] = 0.0; s.m[3] = 0.0; s.m[4] = 1.0; s.m[5] = 0.0;
  if (spec.affine) {
    const double scale = lerp_rn(spec.scale_lo, spec.scale_hi, u_sc);
    const double rotd = lerp_rn(spec.rot_lo, spec.rot_hi, u_rot);
    const double shd = lerp_rn(spec.shear_lo, spec.shear_hi, u_sh);
    const double tx = lerp_rn(spec.tx_lo, spec.tx_hi, u_tx);
    const double ty = lerp_rn(spec.ty_lo, spec.ty_hi, u_ty);
    const double tx_px = (double)__double2int_rn(__dmul_rn(tx, (double)W));
    const double ty_px = (double)__double2int_rn(__dmul_rn(ty, (double)H));
    const double d2r = 0.017453292519943295;  // math.pi / 180.0
    const double rot = __dmul_rn(rotd, d2r), sh = __dmul_rn(shd, d2r);
    const double rs = __dadd_rn(rot, sh);
    const double a00 = __dmul_rn(scale, cos(rot));
    const double a01 = -__dmul_rn(scale, sin(rs));
    const double a10 = __dmul_rn(scale, sin(rot));
    const double a11 = __dmul_rn(scale, cos(rs));
    const double cx = __dsub_rn(__ddiv_rn((double)W, 2.0), 0.5), cy = __dsub_rn(__ddiv_rn((double)H, 2.0), 0.5);
    const double b0 = __dadd_rn(__dadd_rn(__dmul_rn(a00, -cx), __dmul_rn(a01, -cy)), tx_px);
    const double b1 = __dadd_rn(__dadd_rn(__dmul_rn(a10, -cx), __dmul_rn(a11, -cy)), ty_px);
    s.m[0] = a00; s.m[1] = a01; s.m[2] = __dadd_rn(b0, cx);
    s.m[3] = a10; s.m[4] = a11; s.m[5] = __dadd_rn(b1, cy);
  }
  // cv2.warpAffine's inversion (imgwarp.cpp), unfused fp64
  {
    double D = __dsub_rn(__dmul_rn(s.m[0], s.m[4]), __dmul_rn(s.m[1], s.m[3]));
    D = D != 0.0 ? __ddiv_rn(1.0, D) : 0.0;
    const double A11 = __dmul_rn(s.m[4], D), A22 = __dmul_rn(s.m[0], D);
    const double i0 = A11, i1 = __dmul_rn(s.m[1], -D), i3 = __dmul_rn(s.m[3], -D), i4 = A22;
    s.inv[0] = i0; s.inv[1] = i1; s.inv[3] = i3; s.inv[4] = i4;
    s.inv[2] = __dsub_rn(__dmul_rn(-i0, s.m[2]), __dmul_rn(i1, s.m[5]));
    s.inv[5] = __dsub_rn(__dmul_rn(-i3, s.m[2]), __dmul_rn(i4, s.m[5]));
  }
  s.has_mul = spec.has_mul;
  s.mul = spec.has_mul ? (float)lerp_rn(spec.mul_lo, spec.mul_hi, u_mul) : 1.f;
  s.add = 0;
  if (spec.has_add) s.add = spec.add_lo + (int)floor(__dmul_rn(u_add, (double)(spec.add_hi - spec.add_lo + 1)));
  s.src_index = (int32_t)sid;
  const double u_r90 = u53(r[4][2], r[4][3]), u_inv = u53(r[5][0], r[5][1]);
  int k90 = spec.rot90 ? (int)floor(__dmul_rn(u_r90, 4.0)) : 0;
  if (k90 > 3) k90 = 3;
  if (k90 & 1) {  // a flip listed before Rotate90 acts, after an odd quarter turn, as the other flip (stp.h flip_before_rot90)
    const int a = s.fliplr, b = s.flipud, lr_pre = spec.flip_before_rot90 & 1, ud_pre = (spec.flip_before_rot90 >> 1) & 1;
    s.fliplr = (lr_pre ? 0 : a) ^ (ud_pre ? b : 0);
    s.flipud = (ud_pre ? 0 : b) ^ (lr_pre ? a : 0);
  }
  const int inv = u_inv < spec.invert_p ? 1 : 0;
  s.flags2 = k90 | (inv << 2) | ((spec.color_order[0] & 3) << 4) | ((spec.color_order[1] & 3) << 6) | ((spec.color_order[2] & 3) << 8);
  out[i] = s;
}

// colour stage: Multiply / Add / Invert in the order of the YAML block (flags2 bits 4-9); each op saturates to uint8
// SKIP = true: the colour stage runs in augment_pixel_ops_kernel instead (extended pixel-wise augmenters present)
template <bool SKIP>
__device__ __forceinline__ int colour(int v, const DevSample& s, int mul_rint) {
  if (SKIP) return v;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const int op = (s.flags2 >> (4 + 2 * k)) & 3;
    if (op == 0) {
      if (s.has_mul) {
        float f = __fmul_rn((float)v, s.mul);
        f = fminf(fmaxf(f, 0.f), 255.f);
        v = mul_rint ? __float2int_rn(f) : (int)f;
      }
    } else if (op == 1) {
      v += s.add;
      v = v < 0 ? 0 : (v > 255 ? 255 : v);
    } else if (op == 2) {
      if (s.flags2 & 4) v = 255 - v;
    }
  }
  return v;
}
// pool coordinates of pixel (py, px) of the np.rot90(k)-rotated sample (square images): k = 1 is counter-clockwise
__device__ __forceinline__ int64_t rot_off(int py, int px, int k, int H, int W) {
  int sy = py, sx = px;
  if (k == 1) { sy = px; sx = W - 1 - py; }
  else if (k == 2) { sy = H - 1 - py; sx = W - 1 - px; }
  else if (k == 3) { sy = H - 1 - px; sx = py; }
  return (int64_t)sy * W + sx;
}

// one thread = PX consecutive output pixels of one row
template <int PX, int CI, bool SKIPC = false>
__global__ void __launch_bounds__(256) augment_apply_kernel(const uint8_t* __restrict__ img_pool,
                                                            const uint8_t* __restrict__ mask_pool,
                                                            const DevSample* __restrict__ params,
                                                            uint8_t* __restrict__ img_out,
                                                            uint8_t* __restrict__ mask_out, int n, int H, int W,
                                                            int cm, int mul_rint) {
  const int wq = W / PX;
  const int64_t total = (int64_t)n * H * wq;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int xq = (int)(idx % wq);
    const int y = (int)((idx / wq) % H);
    const int b = (int)(idx / ((int64_t)wq * H));
    const DevSample s = params[b];
    const int k90 = s.flags2 & 3;
    const uint8_t* simg = img_pool + (int64_t)s.src_index * H * W * CI;
    const uint8_t* smsk = mask_pool ? mask_pool + (int64_t)s.src_index * H * W * cm : nullptr;
    uint8_t oi[PX * CI];
    uint8_t om[PX * 4];
    long long X0n = 0, Y0n = 0, X0l = 0, Y0l = 0;
    if (s.has_affine) {
      const double yd = (double)y;
      const long long xr = __double2ll_rn(__dmul_rn(__dadd_rn(__dmul_rn(s.inv[1], yd), s.inv[2]), 1024.0));
      const long long yr = __double2ll_rn(__dmul_rn(__dadd_rn(__dmul_rn(s.inv[4], yd), s.inv[5]), 1024.0));
      X0n = xr + 512; Y0n = yr + 512;
      X0l = xr + 16;  Y0l = yr + 16;
    }
#pragma unroll
    for (int p = 0; p < PX; ++p) {
      const int x = xq * PX + p;
      if (!s.has_affine) {
        const int sx = s.fliplr ? W - 1 - x : x, sy = s.flipud ? H - 1 - y : y;
#pragma unroll
        const int64_t so = rot_off(sy, sx, k90, H, W);
#pragma unroll
        for (int c = 0; c < CI; ++c) oi[p * CI + c] = (uint8_t)colour<SKIPC>(simg[so * CI + c], s, mul_rint);
        for (int c = 0; c < cm; ++c) om[p * cm + c] = smsk ? smsk[so * cm + c] : 0;
        continue;
      }
      const double xd = (double)x;
      const long long ad = __double2ll_rn(__dmul_rn(__dmul_rn(s.inv[0], xd), 1024.0));
      const long long bd = __double2ll_rn(__dmul_rn(__dmul_rn(s.inv[3], xd), 1024.0));
      // mask: nearest
      {
        const long long sx = (X0n + ad) >> 10, sy = (Y0n + bd) >> 10;
        const bool ok = sx >= 0 && sx < W && sy >= 0 && sy < H;
        const int px = s.fliplr ? W - 1 - (int)sx : (int)sx, py = s.flipud ? H - 1 - (int)sy : (int)sy;
        const int64_t so = ok ? rot_off(py, px, k90, H, W) : 0;
        for (int c = 0; c < cm; ++c) om[p * cm + c] = (ok && smsk) ? smsk[so * cm + c] : 0;
      }
      // image: bilinear, 5 fractional bits
      {
        const long long X = (X0l + ad) >> 5, Y = (Y0l + bd) >> 5;
        const long long sx = X >> 5, sy = Y >> 5;
        const int ax = (int)(X & 31), ay = (int)(Y & 31);
        const int w00 = (32 - ay) * (32 - ax) * 32, w01 = (32 - ay) * ax * 32, w10 = ay * (32 - ax) * 32,
                  w11 = ay * ax * 32;
        const bool x0ok = sx >= 0 && sx < W, x1ok = sx + 1 >= 0 && sx + 1 < W;
        const bool y0ok = sy >= 0 && sy < H, y1ok = sy + 1 >= 0 && sy + 1 < H;
        const int px0 = s.fliplr ? W - 1 - (int)sx : (int)sx, px1 = s.fliplr ? px0 - 1 : px0 + 1;
        const int py0 = s.flipud ? H - 1 - (int)sy : (int)sy, py1 = s.flipud ? py0 - 1 : py0 + 1;
        const int64_t o00 = (x0ok && y0ok) ? rot_off(py0, px0, k90, H, W) : 0, o01 = (x1ok && y0ok) ? rot_off(py0, px1, k90, H, W) : 0;
        const int64_t o10 = (x0ok && y1ok) ? rot_off(py1, px0, k90, H, W) : 0, o11 = (x1ok && y1ok) ? rot_off(py1, px1, k90, H, W) : 0;
#pragma unroll
        for (int c = 0; c < CI; ++c) {
          const int v00 = (x0ok && y0ok) ? simg[o00 * CI + c] : 0;
          const int v01 = (x1ok && y0ok) ? simg[o01 * CI + c] : 0;
          const int v10 = (x0ok && y1ok) ? simg[o10 * CI + c] : 0;
          const int v11 = (x1ok && y1ok) ? simg[o11 * CI + c] : 0;
          const int v = (w00 * v00 + w01 * v01 + w10 * v10 + w11 * v11 + 16384) >> 15;
          oi[p * CI + c] = (uint8_t)colour<SKIPC>(v, s, mul_rint);
        }
      }
    }
    uint8_t* dst = img_out + (((int64_t)b * H + y) * W + (int64_t)xq * PX) * CI;
    if ((PX * CI) % 4 == 0) {
#pragma unroll
      for (int k = 0; k < PX * CI / 4; ++k)
        reinterpret_cast<uint32_t*>(dst)[k] =
            (uint32_t)oi[4 * k] | ((uint32_t)oi[4 * k + 1] << 8) | ((uint32_t)oi[4 * k + 2] << 16) | ((uint32_t)oi[4 * k + 3] << 24);
    } else {
      for (int k = 0; k < PX * CI; ++k) dst[k] = oi[k];
    }
    if (mask_out) {
      uint8_t* md = mask_out + (((int64_t)b * H + y) * W + (int64_t)xq * PX) * cm;
      if (PX == 4 && cm == 1) {
        *reinterpret_cast<uint32_t*>(md) =
            (uint32_t)om[0] | ((uint32_t)om[1] << 8) | ((uint32_t)om[2] << 16) | ((uint32_t)om[3] << 24);
      } else {
        for (int k = 0; k < PX * cm; ++k) md[k] = om[k];
      }
    }
  }
}

// ---- pixel-wise augmenters in YAML order (schemas/augmenters.raml:43-60, 88-96, 120-122 -> imgaug 0.3.0 [DEP, recalled]) --------
// Multiply / Add / Invert (per-sample draws of stp_augment_draw) and AddElementwise, MultiplyElementwise, Dropout,
// AdditiveGaussianNoise (per-pixel draws), Grayscale (per-sample alpha), optionally gated by OneOf groups.  One thread per
// pixel, in place on the augmented batch.  Randomness: Philox4x32-10, key = seed, counter = (step, sample id, call, step >> 32)
// with call = 32 + op for per-image draws and call = ((pixel << 8) | (64 + op)) [| 128 for the second Gaussian word set] for
// per-pixel draws -- the CPU oracle (oracle/augment.py apply_pixel_ops) draws the same numbers.
__device__ __forceinline__ float u24(uint32_t w) { return (float)(w >> 8) * 5.9604644775390625e-08f; }   // [0, 1), exact

template <int CI>
__global__ void __launch_bounds__(256) augment_pixel_ops_kernel(uint8_t* __restrict__ img, const DevSample* __restrict__ params,
                                                                const stp_aug_pix_spec spec, uint64_t seed,
                                                                const int64_t* __restrict__ d_step, int n, int H, int W) {
  const int64_t step = *d_step;
  const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32), s_lo = (uint32_t)step, s_hi = (uint32_t)(step >> 32);
  const int64_t total = (int64_t)n * H * W;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int b = (int)(idx / ((int64_t)H * W));
    const uint32_t pix = (uint32_t)(idx - (int64_t)b * H * W);
    const DevSample s = params[b];
    const uint32_t sid = (uint32_t)s.src_index;
    int v[CI];
#pragma unroll
    for (int c = 0; c < CI; ++c) v[c] = img[idx * CI + c];
    for (int k = 0; k < spec.n_ops; ++k) {
      const stp_aug_pix_op op = spec.ops[k];
      uint32_t ri[4];
      const uint32_t kk = (uint32_t)(spec.k_base + k);                             // position in the whole colour block
      philox4x32(s_lo, sid, 32u + kk, s_hi, k0, k1, ri);                           // per-image draws of this op
      if (op.group_size > 0) {                                                     // OneOf: one member per sample
        uint32_t rg[4];
        philox4x32(s_lo, sid, 32u + 16u + (uint32_t)op.group_id, s_hi, k0, k1, rg);
        int pick = (int)floor(__dmul_rn(u53(rg[0], rg[1]), (double)op.group_size));
        if (pick >= op.group_size) pick = op.group_size - 1;
        if (pick != op.group_member) continue;
      }
      const bool pc = u53(ri[0], ri[1]) < (double)op.per_channel;
      const float par = __fadd_rn(op.a, __fmul_rn((float)u53(ri[2], ri[3]), __fsub_rn(op.b, op.a)));   // per-image parameter
      if (op.kind == STP_PIX_MULTIPLY) {
        if (s.has_mul) {
#pragma unroll
          for (int c = 0; c < CI; ++c) {
            float f = fminf(fmaxf(__fmul_rn((float)v[c], s.mul), 0.f), 255.f);
            v[c] = spec.mul_rint ? __float2int_rn(f) : (int)f;
          }
        }
      } else if (op.kind == STP_PIX_ADD) {
#pragma unroll
        for (int c = 0; c < CI; ++c) { int t = v[c] + s.add; v[c] = t < 0 ? 0 : (t > 255 ? 255 : t); }
      } else if (op.kind == STP_PIX_INVERT) {
        if (s.flags2 & 4) {
#pragma unroll
          for (int c = 0; c < CI; ++c) v[c] = 255 - v[c];
        }
      } else if (op.kind == STP_PIX_GRAYSCALE) {
        if (CI >= 3) {
          const float g = __fadd_rn(__fadd_rn(__fmul_rn(0.299f, (float)v[0]), __fmul_rn(0.587f, (float)v[1])), __fmul_rn(0.114f, (float)v[2]));
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            const float f = __fadd_rn(__fmul_rn(__fsub_rn(1.f, par), (float)v[c]), __fmul_rn(par, g));
            const int t = __float2int_rn(f);
            v[c] = t < 0 ? 0 : (t > 255 ? 255 : t);
          }
        }
      } else {
        uint32_t rp[4], rq[4];
        philox4x32(s_lo, sid, (pix << 8) | (64u + kk), s_hi, k0, k1, rp);
        if (op.kind == STP_PIX_GAUSSIAN_NOISE) philox4x32(s_lo, sid, (pix << 8) | (64u + 128u + kk), s_hi, k0, k1, rq);
#pragma unroll
        for (int c = 0; c < CI; ++c) {
          const int wi = pc ? (c & 3) : 0;
          const float u = u24(rp[wi]);
          if (op.kind == STP_PIX_ADD_ELEMENTWISE) {
            const int lo = (int)op.a, hi = (int)op.b;
            const int t = v[c] + lo + (int)floorf(__fmul_rn(u, (float)(hi - lo + 1)));
            v[c] = t < 0 ? 0 : (t > 255 ? 255 : t);
          } else if (op.kind == STP_PIX_MULTIPLY_ELEMENTWISE) {
            const float m = __fadd_rn(op.a, __fmul_rn(u, __fsub_rn(op.b, op.a)));
            const float f = fminf(fmaxf(__fmul_rn((float)v[c], m), 0.f), 255.f);
            v[c] = spec.mul_rint ? __float2int_rn(f) : (int)f;
          } else if (op.kind == STP_PIX_DROPOUT) {
            if (u < par) v[c] = 0;
          } else if (op.kind == STP_PIX_GAUSSIAN_NOISE) {
            const float u2 = u24(rq[wi]);
            const float z = __fmul_rn(sqrtf(__fmul_rn(-2.f, logf(__fsub_rn(1.f, u)))), cosf(__fmul_rn(6.283185307179586f, u2)));
            const int t = __float2int_rn(__fadd_rn((float)v[c], __fmul_rn(z, par)));
            v[c] = t < 0 ? 0 : (t > 255 ? 255 : t);
          }
        }
      }
    }
#pragma unroll
    for (int c = 0; c < CI; ++c) img[idx * CI + c] = (uint8_t)v[c];
  }
}

}  // namespace stp

using namespace stp;

extern "C" int stp_augment_pixel_ops(uint8_t* d_img, const stp_aug_sample* d_params, const stp_aug_pix_spec* h_spec, uint64_t seed,
                                     const int64_t* d_step, int32_t n, int32_t h, int32_t w, int32_t c_img, stp_stream stream) {
  STP_REQUIRE(d_img && d_params && h_spec && d_step && n > 0 && h > 0 && w > 0, "augment_pixel_ops: bad args");
  STP_REQUIRE(c_img == 1 || c_img == 3 || c_img == 4, "augment_pixel_ops: c_img must be 1, 3 or 4");
  STP_REQUIRE(h_spec->n_ops >= 0 && h_spec->n_ops <= STP_PIX_MAX_OPS, "augment_pixel_ops: at most %d ops", STP_PIX_MAX_OPS);
  STP_REQUIRE(h_spec->k_base >= 0 && h_spec->k_base + h_spec->n_ops <= 16, "augment_pixel_ops: k_base + n_ops <= 16");
  STP_REQUIRE((int64_t)h * w <= (1 << 24), "augment_pixel_ops: at most 2^24 pixels per image");
  for (int k = 0; k < h_spec->n_ops; ++k)
    STP_REQUIRE(h_spec->ops[k].kind >= STP_PIX_MULTIPLY && h_spec->ops[k].kind <= STP_PIX_GRAYSCALE, "augment_pixel_ops: unknown op kind");
  if (h_spec->n_ops == 0) return STP_OK;
  const int64_t total = (int64_t)n * h * w;
  int64_t nb = (total + 255) / 256;
  const int grid = (int)(nb < (int64_t)kNumSMs * 16 ? nb : (int64_t)kNumSMs * 16);
  cudaStream_t st = (cudaStream_t)stream;
  const DevSample* P = (const DevSample*)d_params;
  if (c_img == 3) augment_pixel_ops_kernel<3><<<grid, 256, 0, st>>>(d_img, P, *h_spec, seed, d_step, n, h, w);
  else if (c_img == 1) augment_pixel_ops_kernel<1><<<grid, 256, 0, st>>>(d_img, P, *h_spec, seed, d_step, n, h, w);
  else augment_pixel_ops_kernel<4><<<grid, 256, 0, st>>>(d_img, P, *h_spec, seed, d_step, n, h, w);
  return check_launch("augment_pixel_ops");
}

extern "C" int stp_augment_draw(const stp_aug_spec* h_spec, uint64_t seed, const int64_t* d_step, int32_t n,
                                int32_t pool, int32_t h, int32_t w, stp_aug_sample* d_out, stp_stream stream) {
  STP_REQUIRE(h_spec && d_step && d_out && n > 0 && pool > 0, "augment_draw: bad args");
  STP_REQUIRE(!h_spec->rot90 || h == w, "augment_draw: Rotate90 needs square images");
  {
    const int32_t* o = h_spec->color_order;
    STP_REQUIRE(o[0] >= 0 && o[0] <= 2 && o[1] >= 0 && o[1] <= 2 && o[2] >= 0 && o[2] <= 2 && o[0] != o[1] && o[0] != o[2] && o[1] != o[2],
                "augment_draw: color_order must be a permutation of {0, 1, 2}");
  }
  augment_draw_kernel<<<(n + 63) / 64, 64, 0, (cudaStream_t)stream>>>(*h_spec, seed, d_step, n, pool, h, w,
                                                                      (DevSample*)d_out);
  return check_launch("augment_draw");
}

extern "C" int stp_augment_apply(const uint8_t* img_pool, const uint8_t* mask_pool, const stp_aug_sample* d_params,
                                 uint8_t* img_out, uint8_t* mask_out, int32_t n, int32_t h, int32_t w, int32_t c_img,
                                 int32_t c_mask, int32_t mul_rint, stp_stream stream) {
  STP_REQUIRE(img_pool && d_params && img_out && n > 0, "augment_apply: bad args");
  STP_REQUIRE(c_img == 3 || c_img == 1 || c_img == 4, "augment_apply: c_img must be 1, 3 or 4");
  STP_REQUIRE(c_mask >= 0 && c_mask <= 4, "augment_apply: c_mask must be <= 4");
  STP_REQUIRE((mask_pool == nullptr) == (mask_out == nullptr), "augment_apply: mask in/out must both be given");
  const bool v4 = (w % 4 == 0) && ((reinterpret_cast<uintptr_t>(img_out) & 3) == 0) &&
                  (!mask_out || (reinterpret_cast<uintptr_t>(mask_out) & 3) == 0);
  int64_t total = (int64_t)n * h * (v4 ? w / 4 : w);
  int64_t nb = (total + 255) / 256;
  int grid = (int)(nb < (int64_t)kNumSMs * 16 ? nb : (int64_t)kNumSMs * 16);
  cudaStream_t st = (cudaStream_t)stream;
  const DevSample* P = (const DevSample*)d_params;
  const bool skipc = (mul_rint & 2) != 0;   // bit 1: leave the colour stage to stp_augment_pixel_ops
  mul_rint &= 1;
#define LAUNCH(PX, CI)                                                                                                          \
  do {                                                                                                                          \
    if (skipc)                                                                                                                  \
      augment_apply_kernel<PX, CI, true><<<grid, 256, 0, st>>>(img_pool, mask_pool, P, img_out, mask_out, n, h, w, c_mask, mul_rint); \
    else                                                                                                                        \
      augment_apply_kernel<PX, CI, false><<<grid, 256, 0, st>>>(img_pool, mask_pool, P, img_out, mask_out, n, h, w, c_mask, mul_rint); \
  } while (0)
  if (v4) {
    if (c_img == 3) LAUNCH(4, 3); else if (c_img == 1) LAUNCH(4, 1); else LAUNCH(4, 4);
  } else {
    if (c_img == 3) LAUNCH(1, 3); else if (c_img == 1) LAUNCH(1, 1); else LAUNCH(1, 4);
  }
#undef LAUNCH
  return check_launch("augment_apply");
}
