// K11b: Lovasz hinge loss (binary, per image) on LOGITS -- replaces musket_core.losses.lovasz_loss, registered at
// reference segmentation.py:15-22 and suggested by schemas/segmentation.raml:12-21 [DEP]; musket's compile strips the
// trailing Activation so the loss sees logits.  Berman's lovasz_hinge(per_image=True):
//     e = 1 - z*(2t-1);  sort e descending (stable);  g = delta Jaccard(cumsum of sorted t);  L_img = sum act(e_sorted)*g
// with act = elu(e)+1 (the Kaggle-TGS variant musket is believed to copy) or relu (Berman's original); mean over images.
//
// One radix sort for the whole batch: 64-bit keys (image index << 32 | order-inverted float bits) keep the images
// contiguous and each image's errors descending; stable, so ties keep pixel order like torch.sort(stable=True).
// The sort and the prefix sum are CUB device primitives (library code, like cuBLAS for a plain GEMM); the Jaccard
// gradient, the loss reduction (deterministic two stage) and the gradient scatter are kernels below.
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include "common.cuh"

namespace stp {

constexpr int kLovBlocks = kNumSMs * 4;

__device__ __forceinline__ uint32_t float_desc_key(float f) {
  uint32_t b = __float_as_uint(f);
  uint32_t asc = b ^ ((b >> 31) ? 0xFFFFFFFFu : 0x80000000u);  // ascending-order image of the float
  return ~asc;                                                  // ascending sort of this == descending floats
}

__global__ void __launch_bounds__(256) lovasz_keys_kernel(const float* __restrict__ logits, const uint8_t* __restrict__ mask,
                                                          int64_t total, int64_t per_img, int classes,
                                                          uint64_t* __restrict__ keys, uint32_t* __restrict__ vals) {
  // element i of the [image][pixel][class] tensor belongs to sort group image*classes + class (classes == 1: the image)
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const float sign = mask[i] ? 1.f : -1.f;
    const float e = 1.f - logits[i] * sign;
    const int64_t group = classes == 1 ? i / per_img : (i / (per_img * classes)) * classes + i % classes;
    keys[i] = ((uint64_t)group << 32) | float_desc_key(e);
    vals[i] = (uint32_t)i;
  }
}

__global__ void __launch_bounds__(256) lovasz_gt_kernel(const uint32_t* __restrict__ perm, const uint8_t* __restrict__ mask,
                                                        int64_t total, int32_t* __restrict__ gt_sorted) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x)
    gt_sorted[i] = mask[perm[i]] ? 1 : 0;
}

// csum = inclusive prefix sum of gt_sorted over the WHOLE batch; per image the local cumsum is csum[i] - base(image).
__global__ void __launch_bounds__(256) lovasz_grad_kernel(const float* __restrict__ logits, const uint8_t* __restrict__ mask,
                                                          const uint32_t* __restrict__ perm, const int32_t* __restrict__ csum,
                                                          int64_t total, int64_t per_img, int act_elu, float inv_images,
                                                          float* __restrict__ g_sorted, float* __restrict__ partial) {
  float acc = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t img = i / per_img, pos = i - img * per_img;
    const int64_t first = img * per_img;
    const int base = first > 0 ? csum[first - 1] : 0;
    // Jaccard increments in double: consecutive values differ by ~1/pixels, far below fp32 resolution of jac itself
    const double gts = (double)(csum[first + per_img - 1] - base);
    const double c = (double)(csum[i] - base);                     // positives among sorted[0..pos]
    const double jac = 1.0 - (gts - c) / (gts + ((double)(pos + 1) - c));
    double jprev = 0.0;
    if (pos > 0) {
      const double cp = (double)(csum[i - 1] - base);
      jprev = 1.0 - (gts - cp) / (gts + ((double)pos - cp));
    }
    const float grad = (float)(jac - jprev);
    const uint32_t j = perm[i];
    const float sign = mask[j] ? 1.f : -1.f;
    const float e = 1.f - logits[j] * sign;
    float a, da;
    if (act_elu) {  // elu(e) + 1
      a = e > 0.f ? e + 1.f : expf(e);
      da = e > 0.f ? 1.f : expf(e);
    } else {
      a = fmaxf(e, 0.f);
      da = e > 0.f ? 1.f : 0.f;
    }
    acc += a * grad;
    g_sorted[i] = -sign * da * grad * inv_images;  // d(mean over images)/d logit[perm[i]]
  }
  __shared__ float sm[8];
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.f;
    for (int w = 0; w < 8; ++w) a += sm[w];
    partial[blockIdx.x] = a;
  }
}

__global__ void lovasz_finalize_kernel(const float* __restrict__ partial, int nblk, float inv_images, float weight,
                                       int accumulate, float* __restrict__ result16, int slot_total, int slot_lovasz) {
  if (threadIdx.x == 0) {
    double a = 0.0;
    for (int b = 0; b < nblk; ++b) a += (double)partial[b];
    const float l = (float)(a * (double)inv_images);
    result16[slot_lovasz] = l;
    result16[slot_total] = (accumulate ? result16[slot_total] : 0.f) + weight * l;
  }
}

__global__ void __launch_bounds__(256) lovasz_scatter_kernel(const float* __restrict__ g_sorted, const uint32_t* __restrict__ perm,
                                                             int64_t total, float weight, int accumulate,
                                                             float* __restrict__ dlogits) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const uint32_t j = perm[i];
    const float g = weight * g_sorted[i];
    dlogits[j] = accumulate ? dlogits[j] + g : g;
  }
}

struct LovLayout {
  size_t keys_a, keys_b, vals_a, vals_b, gt, csum, gs, partial, temp, temp_bytes, total;
};
static LovLayout lov_layout(int64_t total) {
  auto al = [](size_t x) { return (x + 255) / 256 * 256; };
  LovLayout l;
  size_t off = 0;
  l.keys_a = off; off += al(total * 8);
  l.keys_b = off; off += al(total * 8);
  l.vals_a = off; off += al(total * 4);
  l.vals_b = off; off += al(total * 4);
  l.gt = off; off += al(total * 4);
  l.csum = off; off += al(total * 4);
  l.gs = off; off += al(total * 4);
  l.partial = off; off += al(kLovBlocks * 4);
  size_t t1 = 0, t2 = 0;
  cub::DeviceRadixSort::SortPairs((void*)nullptr, t1, (const uint64_t*)nullptr, (uint64_t*)nullptr, (const uint32_t*)nullptr,
                                  (uint32_t*)nullptr, (int)total, 0, 64);
  cub::DeviceScan::InclusiveSum((void*)nullptr, t2, (const int32_t*)nullptr, (int32_t*)nullptr, (int)total);
  l.temp_bytes = t1 > t2 ? t1 : t2;
  l.temp = off; off += al(l.temp_bytes);
  l.total = off;
  return l;
}

}  // namespace stp

using namespace stp;

extern "C" size_t stp_lovasz_workspace(int32_t images, int64_t pixels_per_image) {
  if (images <= 0 || pixels_per_image <= 0) return 0;
  return lov_layout((int64_t)images * pixels_per_image).total;
}

extern "C" int stp_lovasz_fwd(const float* logits, const uint8_t* mask, int32_t images, int64_t pixels_per_image,
                              int32_t act_elu, float weight, int32_t accumulate, void* workspace, size_t workspace_bytes,
                              float* result16, stp_stream stream) {
  return stp_lovasz_fwd_mc(logits, mask, images, pixels_per_image, 1, act_elu, weight, accumulate, workspace, workspace_bytes,
                           result16, stream);
}

extern "C" int stp_lovasz_fwd_mc(const float* logits, const uint8_t* mask, int32_t n_images, int64_t pixels_per_image,
                                 int32_t classes, int32_t act_elu, float weight, int32_t accumulate, void* workspace,
                                 size_t workspace_bytes, float* result16, stp_stream stream) {
  STP_REQUIRE(logits && mask && workspace && result16 && n_images > 0 && pixels_per_image > 0 && classes >= 1,
              "lovasz_fwd: bad args");
  const int32_t images = n_images * classes;  // sort groups: one hinge per (image, class)
  const int64_t total = (int64_t)images * pixels_per_image;
  STP_REQUIRE(total < 0x7fffffff, "lovasz_fwd: too many elements");
  LovLayout l = lov_layout(total);
  if (workspace_bytes < l.total) {
    set_error("lovasz_fwd: workspace too small (%zu < %zu)", workspace_bytes, l.total);
    return STP_E_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  char* w = (char*)workspace;
  uint64_t* keys_a = (uint64_t*)(w + l.keys_a);
  uint64_t* keys_b = (uint64_t*)(w + l.keys_b);
  uint32_t* vals_a = (uint32_t*)(w + l.vals_a);
  uint32_t* perm = (uint32_t*)(w + l.vals_b);
  int32_t* gt = (int32_t*)(w + l.gt);
  int32_t* csum = (int32_t*)(w + l.csum);
  float* gs = (float*)(w + l.gs);
  float* partial = (float*)(w + l.partial);
  const int grid = (int)((total + 255) / 256 < kLovBlocks ? (total + 255) / 256 : kLovBlocks);
  lovasz_keys_kernel<<<grid, 256, 0, st>>>(logits, mask, total, pixels_per_image, classes, keys_a, vals_a);
  int rc = check_launch("lovasz_keys");
  if (rc) return rc;
  int img_bits = 1;
  while ((1ll << img_bits) < images) ++img_bits;
  size_t tb = l.temp_bytes;
  cudaError_t e = cub::DeviceRadixSort::SortPairs((void*)(w + l.temp), tb, (const uint64_t*)keys_a, keys_b, (const uint32_t*)vals_a,
                                                  perm, (int)total, 0, 32 + img_bits, st);
  if (e != cudaSuccess) {
    set_error("lovasz_fwd: radix sort: %s", cudaGetErrorString(e));
    return STP_E_CUDA;
  }
  lovasz_gt_kernel<<<grid, 256, 0, st>>>(perm, mask, total, gt);
  rc = check_launch("lovasz_gt");
  if (rc) return rc;
  tb = l.temp_bytes;
  e = cub::DeviceScan::InclusiveSum((void*)(w + l.temp), tb, (const int32_t*)gt, csum, (int)total, st);
  if (e != cudaSuccess) {
    set_error("lovasz_fwd: scan: %s", cudaGetErrorString(e));
    return STP_E_CUDA;
  }
  const float inv_images = 1.f / (float)images;
  lovasz_grad_kernel<<<grid, 256, 0, st>>>(logits, mask, perm, csum, total, pixels_per_image, act_elu, inv_images, gs, partial);
  rc = check_launch("lovasz_grad");
  if (rc) return rc;
  lovasz_finalize_kernel<<<1, 32, 0, st>>>(partial, grid, inv_images, weight, accumulate, result16, STP_L_LOSS, STP_L_LOVASZ);
  return check_launch("lovasz_finalize");
}

extern "C" int stp_lovasz_bwd(const void* workspace, size_t workspace_bytes, int32_t images, int64_t pixels_per_image,
                              float weight, int32_t accumulate, float* dlogits, stp_stream stream) {
  STP_REQUIRE(workspace && dlogits && images > 0 && pixels_per_image > 0, "lovasz_bwd: bad args");
  const int64_t total = (int64_t)images * pixels_per_image;
  LovLayout l = lov_layout(total);
  if (workspace_bytes < l.total) {
    set_error("lovasz_bwd: workspace too small");
    return STP_E_WORKSPACE;
  }
  const char* w = (const char*)workspace;
  const int grid = (int)((total + 255) / 256 < kLovBlocks ? (total + 255) / 256 : kLovBlocks);
  lovasz_scatter_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const float*)(w + l.gs), (const uint32_t*)(w + l.vals_b), total,
                                                                weight, accumulate, dlogits);
  return check_launch("lovasz_scatter");
}
