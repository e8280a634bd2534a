// Neighbourhood augmenters on the uint8 image batch (schemas/augmenters.raml:97-112, 117-119 -> imgaug 0.3.0 GaussianBlur /
// AverageBlur / MedianBlur / Sharpen / Emboss / EdgeDetect / DirectedEdgeDetect, which call cv2.GaussianBlur / cv2.blur / cv2.medianBlur /
// cv2.filter2D [DEP]).  The cv2 arithmetic is restated exactly and pinned against the real cv2 (tests/test_cpu_oracle.py):
//   GaussianBlur (uint8): 8.8 fixed-point separable kernel -- coefficients round(k_i * 256) with error diffusion from the border
//       inwards, the centre taking the remainder so that the sum is 256 -- integer products, out = (sum + 2^15) >> 16;
//       BORDER_REFLECT_101; imgaug's kernel size: 3.3 / 2.9 / 2.6 * sigma (sigma < 3 / < 5 / else), at least 5, odd.
//   AverageBlur: integer window sum s, out = s / k^2 rounded half up -- except that for k a power of two cv2.blur rounds up one
//       residue earlier, out = (s + k^2/2 + 1) >> log2(k^2) (measured on OpenCV 4.13, k = 2..17, every pixel consistent);
//       anchor k / 2; BORDER_REFLECT_101.
//   MedianBlur: median of the k x k window (k odd, <= 7), BORDER_REPLICATE.
//   Sharpen / Emboss / EdgeDetect: cv2.filter2D with the float32 3x3 matrix (1 - alpha) * identity + alpha * effect:
//       acc = fma(k, x, acc) in float32, row-major over the non-zero coefficients (OpenCV's AVX2 / FMA3 build fuses them; separate
//       multiply and add differs by 1 on ~0.1 % of the pixels), saturate(rint(.)); BORDER_REFLECT_101.
// Per-image parameters: Philox4x32-10, counter (step, sample id, 32 + k_index, step >> 32) -- first parameter from words 0-1,
// second from words 2-3 (53-bit uniforms); OneOf membership as in augment_pixel_ops_kernel.  Masks are untouched (imgaug does
// not blur segmentation maps).  Two kernels: `nb_prep` turns the draws into per-sample tap tables, `nb_apply` is one thread per
// pixel (L1/L2-resident taps).
#include "common.cuh"

namespace stp {
namespace {

constexpr int kNbMaxK = 25;   // Gaussian kernel sizes up to 25 (sigma < ~9.6)

struct NbPrep {   // per sample
  int32_t active, ksize;
  int32_t kfix[kNbMaxK];
  float mat[9];
};

__device__ __forceinline__ void philox_nb(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t* out) {
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
__device__ __forceinline__ double u53n(uint32_t hi, uint32_t lo) { return (double)((((uint64_t)hi << 32) | lo) >> 11) * 1.1102230246251565e-16; }
__device__ __forceinline__ int reflect101(int i, int n) {
  if (n == 1) return 0;
  while (i < 0 || i >= n) i = i < 0 ? -i : 2 * (n - 1) - i;
  return i;
}

__global__ void nb_prep_kernel(const stp_aug_nb_op op, const stp_aug_sample* __restrict__ params, uint64_t seed,
                               const int64_t* __restrict__ d_step, int n, NbPrep* __restrict__ out) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= n) return;
  const int64_t step = *d_step;
  const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32), s_lo = (uint32_t)step, s_hi = (uint32_t)(step >> 32);
  const uint32_t sid = (uint32_t)params[b].src_index;
  NbPrep p;
  p.active = 1;
  p.ksize = 0;
  if (op.group_size > 0) {
    uint32_t rg[4];
    philox_nb(s_lo, sid, 32u + 16u + (uint32_t)op.group_id, s_hi, k0, k1, rg);
    int pick = (int)floor(__dmul_rn(u53n(rg[0], rg[1]), (double)op.group_size));
    if (pick >= op.group_size) pick = op.group_size - 1;
    if (pick != op.group_member) p.active = 0;
  }
  uint32_t ri[4];
  philox_nb(s_lo, sid, 32u + (uint32_t)op.k_index, s_hi, k0, k1, ri);
  const double u1 = u53n(ri[0], ri[1]), u2 = u53n(ri[2], ri[3]);
  const double p1 = __dadd_rn((double)op.a, __dmul_rn(u1, __dsub_rn((double)op.b, (double)op.a)));
  const double p2 = __dadd_rn((double)op.c, __dmul_rn(u2, __dsub_rn((double)op.d, (double)op.c)));
  if (op.kind == STP_NB_GAUSSIAN_BLUR) {
    const double sigma = p1;
    if (sigma < 1e-3) {
      p.active = 0;
    } else {
      double kf = sigma < 3.0 ? 3.3 * sigma : (sigma < 5.0 ? 2.9 * sigma : 2.6 * sigma);
      int ks = (int)(kf > 5.0 ? kf : 5.0);
      if (ks % 2 == 0) ks += 1;
      if (ks > kNbMaxK) ks = kNbMaxK;   // (the host rejects sigma ranges that could exceed it)
      p.ksize = ks;
      double kd[kNbMaxK], sum = 0.0;
      const double scale2x = -0.5 / (sigma * sigma);
      for (int i = 0; i < ks; ++i) {
        const double x = (double)i - (double)(ks - 1) * 0.5;
        kd[i] = exp(scale2x * x * x);
        sum += kd[i];
      }
      const double inv = 1.0 / sum;
      double err = 0.0;
      int s2 = 0;
      for (int i = 0; i < ks / 2; ++i) {
        const double adj = __dadd_rn(__dmul_rn(__dmul_rn(kd[i], inv), 256.0), err);
        const int v = (int)floor(adj + 0.5);
        err = adj - (double)v;
        p.kfix[i] = p.kfix[ks - 1 - i] = v;
        s2 += 2 * v;
      }
      p.kfix[ks / 2] = 256 - s2;
    }
  } else if (op.kind == STP_NB_AVERAGE_BLUR || op.kind == STP_NB_MEDIAN_BLUR) {
    const int lo = (int)op.a, hi = (int)op.b;
    int k = lo + (int)floor(__dmul_rn(u1, (double)(hi - lo + 1)));
    if (k > hi) k = hi;
    if (op.kind == STP_NB_MEDIAN_BLUR && k % 2 == 0) k += 1;
    p.ksize = k;
    if (k <= 1) p.active = 0;
  } else {
    const float alpha = (float)p1, c1 = (float)(1.0 - p1);
    float eff[9];
    if (op.kind == STP_NB_SHARPEN) {
      const float mid = (float)(8.0 + p2);
      const float e[9] = {-1.f, -1.f, -1.f, -1.f, mid, -1.f, -1.f, -1.f, -1.f};
      for (int i = 0; i < 9; ++i) eff[i] = e[i];
    } else if (op.kind == STP_NB_EMBOSS) {
      const float s = (float)p2;   // effect matrix entries are float32 of (-1 - s), (0 - s), ...
      const float e[9] = {(float)(-1.0 - p2), (float)(0.0 - p2), 0.f, (float)(0.0 - p2), 1.f, (float)(0.0 + p2), 0.f, (float)(0.0 + p2), (float)(1.0 + p2)};
      (void)s;
      for (int i = 0; i < 9; ++i) eff[i] = e[i];
    } else if (op.kind == STP_NB_DIRECTED_EDGE_DETECT) {
      // imgaug computes the effect matrix from deg = int(direction * 360) % 360 with double-precision trigonometry: the 360
      // possible float32 matrices are tabulated on the host (no libm / CUDA libdevice last-bit differences)
      int deg = (int)(p2 * 360.0) % 360;
      if (deg < 0) deg += 360;
      const float* t = (const float*)op.d_table + deg * 9;
      for (int i = 0; i < 9; ++i) eff[i] = t[i];
    } else {
      const float e[9] = {0.f, 1.f, 0.f, 1.f, -4.f, 1.f, 0.f, 1.f, 0.f};
      for (int i = 0; i < 9; ++i) eff[i] = e[i];
    }
    for (int i = 0; i < 9; ++i) p.mat[i] = __fadd_rn(__fmul_rn(c1, i == 4 ? 1.f : 0.f), __fmul_rn(alpha, eff[i]));
    p.ksize = 3;
  }
  out[b] = p;
}

template <int CI>
__global__ void __launch_bounds__(256) nb_apply_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, const NbPrep* __restrict__ prep,
                                                       int kind, int n, int H, int W) {
  const int64_t total = (int64_t)n * H * W;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int b = (int)(idx / ((int64_t)H * W));
    const int rem = (int)(idx - (int64_t)b * H * W);
    const int y = rem / W, x = rem - y * W;
    const NbPrep* p = prep + b;
    const uint8_t* im = src + (int64_t)b * H * W * CI;
    if (!p->active) {
#pragma unroll
      for (int c = 0; c < CI; ++c) dst[idx * CI + c] = src[idx * CI + c];
      continue;
    }
    const int ks = p->ksize;
    if (kind == STP_NB_GAUSSIAN_BLUR) {
      const int r = ks / 2;
      int acc[CI];
#pragma unroll
      for (int c = 0; c < CI; ++c) acc[c] = 0;
      for (int a = 0; a < ks; ++a) {
        const int yy = reflect101(y + a - r, H);
        const int ka = p->kfix[a];
        int row[CI];
#pragma unroll
        for (int c = 0; c < CI; ++c) row[c] = 0;
        for (int bq = 0; bq < ks; ++bq) {
          const int xx = reflect101(x + bq - r, W);
          const int kb = p->kfix[bq];
#pragma unroll
          for (int c = 0; c < CI; ++c) row[c] += kb * (int)im[((int64_t)yy * W + xx) * CI + c];
        }
#pragma unroll
        for (int c = 0; c < CI; ++c) acc[c] += ka * row[c];
      }
#pragma unroll
      for (int c = 0; c < CI; ++c) {
        const int v = (acc[c] + 32768) >> 16;
        dst[idx * CI + c] = (uint8_t)(v > 255 ? 255 : v);
      }
    } else if (kind == STP_NB_AVERAGE_BLUR) {
      const int r = ks / 2;
      int acc[CI];
#pragma unroll
      for (int c = 0; c < CI; ++c) acc[c] = 0;
      for (int a = 0; a < ks; ++a) {
        const int yy = reflect101(y + a - r, H);
        for (int bq = 0; bq < ks; ++bq) {
          const int xx = reflect101(x + bq - r, W);
#pragma unroll
          for (int c = 0; c < CI; ++c) acc[c] += (int)im[((int64_t)yy * W + xx) * CI + c];
        }
      }
      const int kk = ks * ks;
      const int up_from = (ks & (ks - 1)) == 0 ? kk / 2 - 1 : (kk + 1) / 2;   // cv2.blur's rounding, see the header
#pragma unroll
      for (int c = 0; c < CI; ++c) {
        const int v = acc[c] / kk + (acc[c] % kk >= up_from ? 1 : 0);
        dst[idx * CI + c] = (uint8_t)(v > 255 ? 255 : v);
      }
    } else if (kind == STP_NB_MEDIAN_BLUR) {
      const int r = ks / 2, cnt = ks * ks;
#pragma unroll
      for (int c = 0; c < CI; ++c) {
        // counting selection over the 256 possible values would need a histogram; k <= 7 -> insertion sort of <= 49 bytes
        uint8_t v[49];
        int m = 0;
        for (int a = 0; a < ks; ++a) {
          int yy = y + a - r;
          yy = yy < 0 ? 0 : (yy >= H ? H - 1 : yy);
          for (int bq = 0; bq < ks; ++bq) {
            int xx = x + bq - r;
            xx = xx < 0 ? 0 : (xx >= W ? W - 1 : xx);
            const uint8_t t = im[((int64_t)yy * W + xx) * CI + c];
            int j = m++;
            while (j > 0 && v[j - 1] > t) { v[j] = v[j - 1]; --j; }
            v[j] = t;
          }
        }
        dst[idx * CI + c] = v[cnt / 2];
      }
    } else {   // filter2D 3x3
      float acc[CI];
#pragma unroll
      for (int c = 0; c < CI; ++c) acc[c] = 0.f;
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        const int yy = reflect101(y + a - 1, H);
#pragma unroll
        for (int bq = 0; bq < 3; ++bq) {
          const float kf = p->mat[a * 3 + bq];
          if (kf == 0.f) continue;
          const int xx = reflect101(x + bq - 1, W);
#pragma unroll
          for (int c = 0; c < CI; ++c) acc[c] = __fmaf_rn(kf, (float)im[((int64_t)yy * W + xx) * CI + c], acc[c]);
        }
      }
#pragma unroll
      for (int c = 0; c < CI; ++c) {
        const int v = __float2int_rn(acc[c]);
        dst[idx * CI + c] = (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
      }
    }
  }
}

}  // namespace
}  // namespace stp

using namespace stp;

extern "C" size_t stp_augment_neighbourhood_workspace(int32_t n) { return (size_t)(n > 0 ? n : 0) * sizeof(NbPrep); }

extern "C" int stp_augment_neighbourhood(const uint8_t* d_src, uint8_t* d_dst, const stp_aug_sample* d_params, const stp_aug_nb_op* h_op,
                                         uint64_t seed, const int64_t* d_step, int32_t n, int32_t h, int32_t w, int32_t c_img,
                                         void* d_work, size_t work_bytes, stp_stream stream) {
  STP_REQUIRE(d_src && d_dst && d_src != d_dst && d_params && h_op && d_step && d_work && n > 0 && h > 0 && w > 0,
              "augment_neighbourhood: bad args (src and dst must differ)");
  STP_REQUIRE(c_img == 1 || c_img == 3 || c_img == 4, "augment_neighbourhood: c_img must be 1, 3 or 4");
  STP_REQUIRE(h_op->kind >= STP_NB_GAUSSIAN_BLUR && h_op->kind <= STP_NB_DIRECTED_EDGE_DETECT, "augment_neighbourhood: unknown op kind");
  STP_REQUIRE(h_op->kind != STP_NB_DIRECTED_EDGE_DETECT || h_op->d_table, "augment_neighbourhood: DirectedEdgeDetect needs d_table (360 x 9 floats)");
  STP_REQUIRE(h_op->k_index >= 0 && h_op->k_index < 16, "augment_neighbourhood: k_index in [0, 16)");
  if (h_op->kind == STP_NB_GAUSSIAN_BLUR)
    STP_REQUIRE(h_op->a >= 0.f && h_op->b >= h_op->a && 2.6 * h_op->b < (double)kNbMaxK, "augment_neighbourhood: GaussianBlur sigma in [0, %.1f)", kNbMaxK / 2.6);
  if (h_op->kind == STP_NB_AVERAGE_BLUR) STP_REQUIRE(h_op->a >= 0.f && h_op->b >= h_op->a && h_op->b <= 31.f, "augment_neighbourhood: AverageBlur k in [0, 31]");
  if (h_op->kind == STP_NB_MEDIAN_BLUR) STP_REQUIRE(h_op->a >= 1.f && h_op->b >= h_op->a && h_op->b <= 7.f, "augment_neighbourhood: MedianBlur k in [1, 7]");
  if (work_bytes < stp_augment_neighbourhood_workspace(n)) {
    set_error("augment_neighbourhood: workspace too small");
    return STP_E_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  nb_prep_kernel<<<(n + 63) / 64, 64, 0, st>>>(*h_op, d_params, seed, d_step, n, (NbPrep*)d_work);
  int rc = check_launch("augment_neighbourhood (prep)");
  if (rc) return rc;
  const int64_t total = (int64_t)n * h * w;
  int64_t nb = (total + 255) / 256;
  const int grid = (int)(nb < (int64_t)kNumSMs * 16 ? nb : (int64_t)kNumSMs * 16);
  const NbPrep* P = (const NbPrep*)d_work;
  if (c_img == 3) nb_apply_kernel<3><<<grid, 256, 0, st>>>(d_src, d_dst, P, h_op->kind, n, h, w);
  else if (c_img == 1) nb_apply_kernel<1><<<grid, 256, 0, st>>>(d_src, d_dst, P, h_op->kind, n, h, w);
  else nb_apply_kernel<4><<<grid, 256, 0, st>>>(d_src, d_dst, P, h_op->kind, n, h, w);
  return check_launch("augment_neighbourhood");
}
