// Probe of the tcgen05 shared-memory descriptor address generation for K-major SWIZZLE_128B operands whose start address
// is NOT aligned to the 1024-byte swizzle atom (1-pixel column shifts inside a halo box) and whose 8-row-group stride (SBO)
// is not a multiple of 1024 bytes.  Decides how conv_tc3 can take all three filter-column taps from ONE halo box.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -I segmentation_training_pipeline_b200/csrc \
//        scripts/umma_probe.cu -o scripts/bin/umma_probe && scripts/bin/umma_probe
//
// Method: A region = 512 rows x 128 B written the way TMA SWIZZLE_128B writes a dense box (16-byte chunk c of absolute row
// j lands at chunk c ^ (j & 7)); B = 64x64 identity.  D[m][n] = A_fetched[m][n], so with A holding (a) the row index and
// (b) the logical channel the host reads back WHICH shared-memory element the tensor core fetched for every (m, k).
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "tc_common.cuh"

using namespace stp::tc;

struct Cfg {
  int start_row;    // descriptor start address = A base + start_row * 128
  int base_offset;  // descriptor bits 49-51
  int sbo_bytes;    // stride between 8-row groups
};

__global__ void __launch_bounds__(128, 1) probe_kernel(Cfg cfg, int pass, float* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
  uint8_t* sA = smem;                 // 512 rows x 128 B = 64 KB
  uint8_t* sB = smem + 512 * 128;     // 64 rows x 128 B
  __shared__ uint64_t bar;
  __shared__ uint32_t tslot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int idx = tid; idx < 512 * 64; idx += 128) {
    const int j = idx >> 6, k = idx & 63, c = k >> 3, e = k & 7;
    float v = pass == 0 ? (float)(j & 255) : pass == 1 ? (float)(j >> 8) : (float)k;
    *reinterpret_cast<__nv_bfloat16*>(sA + j * 128 + ((c ^ (j & 7)) << 4) + e * 2) = __float2bfloat16(v);
  }
  for (int idx = tid; idx < 64 * 64; idx += 128) {
    const int n = idx >> 6, k = idx & 63, c = k >> 3, e = k & 7;
    *reinterpret_cast<__nv_bfloat16*>(sB + n * 128 + ((c ^ (n & 7)) << 4) + e * 2) = __float2bfloat16(n == k ? 1.f : 0.f);
  }
  if (tid == 0) {
    mbar_init(&bar, 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    tmem_alloc(&tslot, 64);
    tmem_relinquish();
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tslot;
  if (warp == 0) {
    uint64_t ad = desc_kmajor(smem_u32(sA + cfg.start_row * 128), 128);
    ad &= ~((uint64_t)0x3FFF << 32);
    ad |= (uint64_t)((cfg.sbo_bytes >> 4) & 0x3FFF) << 32;
    ad |= (uint64_t)(cfg.base_offset & 7) << 49;
    const uint64_t bd = desc_kmajor(smem_u32(sB), 128);
    constexpr uint32_t idesc = idesc_bf16(128, 64, 0, 0);
    if (elect_one()) {
#pragma unroll
      for (int k = 0; k < 4; ++k) umma_bf16(tmem, ad + (uint64_t)(k * 2), bd + (uint64_t)(k * 2), idesc, k != 0);
      umma_commit(&bar);
    }
    __syncwarp();
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  for (int c0 = 0; c0 < 64; c0 += 32) {
    uint32_t rr[32];
    tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + c0, rr);
    tmem_ld_wait();
    for (int i = 0; i < 32; ++i) out[(warp * 32 + lane) * 64 + c0 + i] = __uint_as_float(rr[i]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, 64);
  }
}

int main() {
  const Cfg cfgs[] = {
      {0, 0, 1024}, {1, 0, 1024}, {1, 1, 1024}, {3, 3, 1024}, {3, 0, 1024}, {9, 1, 1024},
      {0, 0, 1280}, {1, 1, 1280}, {1, 0, 1280}, {2, 2, 1280}, {11, 3, 1280}, {12, 4, 1280},
      {0, 0, 2304}, {1, 1, 2304}, {2, 2, 2304}, {20, 4, 2304},
      {0, 0, 1152}, {1, 1, 1152},
  };
  const int smem_bytes = 512 * 128 + 64 * 128 + 1024;
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
  float* d;
  cudaMalloc(&d, 128 * 64 * 4);
  std::vector<float> h[3];
  for (auto& v : h) v.resize(128 * 64);
  for (const Cfg& c : cfgs) {
    for (int pass = 0; pass < 3; ++pass) {
      probe_kernel<<<1, 128, smem_bytes>>>(c, pass, d);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) {
        printf("cfg start=%d bo=%d sbo=%d: CUDA error %s\n", c.start_row, c.base_offset, c.sbo_bytes, cudaGetErrorString(e));
        return 1;
      }
      cudaMemcpy(h[pass].data(), d, 128 * 64 * 4, cudaMemcpyDeviceToHost);
    }
    // expectation under "absolute row" semantics: row(m) = start_row + (m/8) * (sbo/128) + (m%8); channel(m, n) = n
    int row_ok = 0, chan_ok = 0, row_uniform = 0;
    int xors[128];
    for (int m = 0; m < 128; ++m) {
      const int want = c.start_row + (m >> 3) * (c.sbo_bytes >> 7) + (m & 7);
      bool uni = true, rok = true, cok = true;
      int x = -1;
      for (int n = 0; n < 64; ++n) {
        const int j = (int)h[0][m * 64 + n] + 256 * (int)h[1][m * 64 + n];
        const int k = (int)h[2][m * 64 + n];
        if (j != (int)h[0][m * 64] + 256 * (int)h[1][m * 64]) uni = false;
        if (j != want) rok = false;
        if (k != n) cok = false;
        const int xx = (k >> 3) ^ (n >> 3);
        if (x < 0) x = xx; else if (x != xx) x = 99;
      }
      row_ok += rok; chan_ok += cok; row_uniform += uni;
      xors[m] = x;
    }
    printf("start=%2d bo=%d sbo=%4d : rows as expected %3d/128, channels in order %3d/128, rows uniform %3d/128\n", c.start_row,
           c.base_offset, c.sbo_bytes, row_ok, chan_ok, row_uniform);
    printf("   fetched row of m=0..23 :");
    for (int m = 0; m < 24; ++m) printf(" %d", (int)h[0][m * 64] + 256 * (int)h[1][m * 64]);
    printf("\n   chunk xor error m=0..23:");
    for (int m = 0; m < 24; ++m) printf(" %d", xors[m]);
    printf("\n");
  }
  cudaFree(d);
  return 0;
}
