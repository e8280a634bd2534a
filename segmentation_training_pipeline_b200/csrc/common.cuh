// Shared helpers for libstp kernels (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>

#include "../../include/stp.h"

namespace stp {

void set_error(const char* fmt, ...);
extern std::atomic<int64_t> g_launches;
extern std::atomic<int64_t> g_tc_launches;
extern std::atomic<int64_t> g_tc3_launches;

inline int check_launch(const char* what) {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return STP_E_CUDA;
  }
  return STP_OK;
}

#define STP_REQUIRE(cond, ...)                 \
  do {                                         \
    if (!(cond)) {                             \
      ::stp::set_error(__VA_ARGS__);           \
      return STP_E_INVALID;                    \
    }                                          \
  } while (0)

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

inline int64_t pixels(const stp_tensor* t) { return (int64_t)t->n * t->h * t->w; }

// A bf16 NHWC tensor usable with 16-byte vector access.
inline bool vec_ok(const stp_tensor* t) {
  return t && t->ptr && t->dtype == STP_BF16 && t->c % 8 == 0 && t->ld % 8 == 0 && t->ld >= t->c && aligned16(t->ptr);
}

constexpr int kNumSMs = 148;

// ---- programmatic dependent launch (PDL) -----------------------------------------------------------------------
// Kernels launched through launch_pdl() may become resident while the previous kernel of the stream is still draining:
// they signal launch_dependents at once, run their prologue (barrier init, TMEM allocation, descriptor prefetch) and
// only then execute griddepcontrol.wait, which returns when every prerequisite grid has completed and flushed.  No
// global memory written by an earlier kernel is touched before pdl_wait().  Hides launch latency and prologues
// behind the previous kernel's tail inside the step graph; `stp_set_option("pdl", 0)` turns the attribute off.
extern std::atomic<int> g_pdl_enabled;
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = g_pdl_enabled.load(std::memory_order_relaxed) ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

struct alignas(16) bf16x8 {
  __nv_bfloat162 v[4];
};

__device__ __forceinline__ void unpack8(const bf16x8& a, float* f) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 t = __bfloat1622float2(a.v[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}
__device__ __forceinline__ bf16x8 pack8(const float* f) {
  bf16x8 r;
#pragma unroll
  for (int i = 0; i < 4; ++i) r.v[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
  return r;
}
__device__ __forceinline__ bf16x8 ld8(const __nv_bfloat16* p) { return *reinterpret_cast<const bf16x8*>(p); }
__device__ __forceinline__ void st8(__nv_bfloat16* p, const bf16x8& v) { *reinterpret_cast<bf16x8*>(p) = v; }

// activation codes of the BatchNorm(+activation) kernels: 0 none, 1 ReLU, 2 ReLU6 (keras relu(max_value=6), MobileNetV2 blocks of
// the reference's DeepLabV3+, impl/deeplab/model.py:38-39 / :248-262).  Gradient passes on 0 < t <= 6 (tf.clip_by_value's
// gradient includes the upper boundary; ReLU's excludes 0).
constexpr int kActRelu6 = 2;
__device__ __forceinline__ float relu_act(float t, int act) {
  if (act == 0) return t;
  if (!(t > 0.f)) return 0.f;
  return (act == kActRelu6 && t > 6.f) ? 6.f : t;
}
__device__ __forceinline__ bool relu_pass(float t, int act) { return t > 0.f && (act != kActRelu6 || t <= 6.f); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace stp
