// K12: Keras-formulation optimizers over the flat fp32 parameter / gradient buffers (multi-tensor by
// construction: ONE launch per step).  keras.optimizers.Adam 2.2.4:
//   t += 1; lr_t = lr*sqrt(1-b2^t)/(1-b1^t); m = b1 m + (1-b1) g; v = b2 v + (1-b2) g^2;
//   p -= lr_t * m / (sqrt(v) + eps)          (eps OUTSIDE the bias correction, eps = 1e-7)
// HBM bound: reads p,g,m,v, writes p,m,v = 28 B / parameter.
#include "common.cuh"

namespace stp {

struct GX {
  float scale, clipnorm, clipvalue;
  const float* sumsq;
  const float* lr_scale;  // device scalar multiplying lr (learning-rate schedules without re-capturing the step graph)
};
__device__ __forceinline__ float gx_lr(float lr, const GX& gx) { return gx.lr_scale ? lr * (*gx.lr_scale) : lr; }

__device__ __forceinline__ float gx_scale(const GX& gx) {
  float s = gx.scale;
  if (gx.clipnorm > 0.f && gx.sumsq) {
    float norm = sqrtf(*gx.sumsq) * fabsf(gx.scale);
    if (norm > gx.clipnorm) s *= gx.clipnorm / norm;
  }
  return s;
}
__device__ __forceinline__ float gx_apply(float g, float s, const GX& gx) {
  g *= s;
  if (gx.clipvalue > 0.f) g = fminf(fmaxf(g, -gx.clipvalue), gx.clipvalue);
  return g;
}

__global__ void __launch_bounds__(256) adam_kernel(float4* __restrict__ p, const float4* __restrict__ g,
                                                   float4* __restrict__ m, float4* __restrict__ v, int64_t n4,
                                                   float lr, float b1, float b2, float eps, GX gx,
                                                   const int64_t* __restrict__ d_step) {
  // d_step holds the number of COMPLETED steps; this update is step t = *d_step + 1
  const double t = (double)(*d_step + 1);
  const float lr_t = (float)((double)gx_lr(lr, gx) * sqrt(1.0 - pow((double)b2, t)) / (1.0 - pow((double)b1, t)));
  const float s = gx_scale(gx);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    float4 pp = p[i], gg = g[i], mm = m[i], vv = v[i];
    float* P = &pp.x;
    float* G = &gg.x;
    float* M = &mm.x;
    float* V = &vv.x;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float gk = gx_apply(G[k], s, gx);
      M[k] = b1 * M[k] + (1.f - b1) * gk;
      V[k] = b2 * V[k] + (1.f - b2) * gk * gk;
      P[k] -= lr_t * M[k] / (sqrtf(V[k]) + eps);
    }
    p[i] = pp;
    m[i] = mm;
    v[i] = vv;
  }
}

__global__ void __launch_bounds__(256) sgd_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                  float* __restrict__ v, int64_t n, float lr, float mu, int nesterov,
                                                  GX gx) {
  const float s = gx_scale(gx);
  lr = gx_lr(lr, gx);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float gk = gx_apply(g[i], s, gx);
    float vk = mu * v[i] - lr * gk;
    v[i] = vk;
    p[i] += nesterov ? (mu * vk - lr * gk) : vk;
  }
}

__global__ void __launch_bounds__(256) rmsprop_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                      float* __restrict__ a, int64_t n, float lr, float rho,
                                                      float eps, GX gx) {
  const float s = gx_scale(gx);
  lr = gx_lr(lr, gx);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float gk = gx_apply(g[i], s, gx);
    float ak = rho * a[i] + (1.f - rho) * gk * gk;
    a[i] = ak;
    p[i] -= lr * gk / (sqrtf(ak) + eps);
  }
}

// keras Nadam: sched[0] = running product of the momentum schedule (1.0 initially); this single-thread kernel, launched
// before nadam_kernel, publishes {m_schedule_new, m_schedule_next, mu_t, mu_t1} in sched[1..4] and advances sched[0].
__global__ void nadam_schedule_kernel(float* sched, const int64_t* d_step, float b1, float schedule_decay) {
  const double t = (double)(*d_step + 1);
  const double mu_t = (double)b1 * (1.0 - 0.5 * pow(0.96, t * (double)schedule_decay));
  const double mu_t1 = (double)b1 * (1.0 - 0.5 * pow(0.96, (t + 1.0) * (double)schedule_decay));
  const double ms_new = (double)sched[0] * mu_t;
  sched[0] = (float)ms_new;
  sched[1] = (float)ms_new;
  sched[2] = (float)(ms_new * mu_t1);
  sched[3] = (float)mu_t;
  sched[4] = (float)mu_t1;
}

__global__ void __launch_bounds__(256) nadam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                    float* __restrict__ v, int64_t n, float lr, float b1, float b2, float eps,
                                                    GX gx, const int64_t* __restrict__ d_step, const float* __restrict__ sched) {
  const double t = (double)(*d_step + 1);
  const float ms_new = sched[1], ms_next = sched[2], mu_t = sched[3], mu_t1 = sched[4];
  const float vcorr = (float)(1.0 - pow((double)b2, t));
  const float s = gx_scale(gx);
  lr = gx_lr(lr, gx);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float gk = gx_apply(g[i], s, gx);
    const float mk = b1 * m[i] + (1.f - b1) * gk;
    const float vk = b2 * v[i] + (1.f - b2) * gk * gk;
    m[i] = mk;
    v[i] = vk;
    const float m_bar = (1.f - mu_t) * (gk / (1.f - ms_new)) + mu_t1 * (mk / (1.f - ms_next));
    p[i] -= lr * m_bar / (sqrtf(vk / vcorr) + eps);
  }
}

__global__ void step_advance_kernel(int64_t* s) { *s += 1; }

constexpr int kSumsqBlocks = 1024;
__global__ void __launch_bounds__(256) sumsq_kernel(const float* __restrict__ g, int64_t n, float* __restrict__ partial) {
  float s = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    s += g[i] * g[i];
  __shared__ float sm[8];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.f;
    for (int w = 0; w < 8; ++w) a += sm[w];
    partial[blockIdx.x] = a;
  }
}
__global__ void sumsq_final_kernel(const float* partial, int nblk, float* out) {
  if (threadIdx.x == 0) {
    double a = 0.0;
    for (int b = 0; b < nblk; ++b) a += (double)partial[b];
    *out = (float)a;
  }
}

static GX make_gx(const stp_grad_xform* h) {
  GX g{1.f, 0.f, 0.f, nullptr, nullptr};
  if (h) {
    g.scale = h->scale;
    g.clipnorm = h->clipnorm;
    g.clipvalue = h->clipvalue;
    g.sumsq = h->d_sumsq;
    g.lr_scale = h->d_lr_scale;
  }
  return g;
}
static int opt_grid(int64_t n) {
  int64_t b = (n + 255) / 256;
  int64_t cap = (int64_t)kNumSMs * 8;
  return (int)(b < cap ? (b < 1 ? 1 : b) : cap);
}

}  // namespace stp

using namespace stp;

extern "C" int stp_adam(float* p, const float* g, float* m, float* v, int64_t count, float lr, float beta1,
                        float beta2, float eps, const stp_grad_xform* h_gx, const int64_t* d_step, stp_stream stream) {
  STP_REQUIRE(p && g && m && v && d_step && count > 0, "adam: bad args");
  STP_REQUIRE(count % 4 == 0 && aligned16(p) && aligned16(g) && aligned16(m) && aligned16(v),
              "adam: flat buffers must be 16B aligned and a multiple of 4 floats");
  adam_kernel<<<opt_grid(count / 4), 256, 0, (cudaStream_t)stream>>>((float4*)p, (const float4*)g, (float4*)m,
                                                                     (float4*)v, count / 4, lr, beta1, beta2, eps,
                                                                     make_gx(h_gx), d_step);
  return check_launch("adam");
}

extern "C" int stp_sgd(float* p, const float* g, float* v, int64_t count, float lr, float momentum, int32_t nesterov,
                       const stp_grad_xform* h_gx, stp_stream stream) {
  STP_REQUIRE(p && g && v && count > 0, "sgd: bad args");
  sgd_kernel<<<opt_grid(count), 256, 0, (cudaStream_t)stream>>>(p, g, v, count, lr, momentum, nesterov, make_gx(h_gx));
  return check_launch("sgd");
}

extern "C" int stp_rmsprop(float* p, const float* g, float* a, int64_t count, float lr, float rho, float eps,
                           const stp_grad_xform* h_gx, stp_stream stream) {
  STP_REQUIRE(p && g && a && count > 0, "rmsprop: bad args");
  rmsprop_kernel<<<opt_grid(count), 256, 0, (cudaStream_t)stream>>>(p, g, a, count, lr, rho, eps, make_gx(h_gx));
  return check_launch("rmsprop");
}

extern "C" int stp_nadam(float* p, const float* g, float* m, float* v, float* sched5, int64_t count, float lr, float beta1,
                         float beta2, float eps, float schedule_decay, const stp_grad_xform* h_gx, const int64_t* d_step,
                         stp_stream stream) {
  STP_REQUIRE(p && g && m && v && sched5 && d_step && count > 0, "nadam: bad args");
  nadam_schedule_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(sched5, d_step, beta1, schedule_decay);
  int rc = check_launch("nadam_schedule");
  if (rc) return rc;
  nadam_kernel<<<opt_grid(count), 256, 0, (cudaStream_t)stream>>>(p, g, m, v, count, lr, beta1, beta2, eps, make_gx(h_gx),
                                                                  d_step, sched5);
  return check_launch("nadam");
}

extern "C" int stp_step_advance(int64_t* d_step, stp_stream stream) {
  STP_REQUIRE(d_step, "step_advance: null");
  step_advance_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(d_step);
  return check_launch("step_advance");
}

extern "C" int stp_sumsq(const float* g, int64_t count, float* partial, float* out, stp_stream stream) {
  STP_REQUIRE(g && partial && out && count > 0, "sumsq: bad args");
  int64_t nb = (count + 255) / 256;
  int nblk = (int)(nb < kSumsqBlocks ? nb : kSumsqBlocks);
  sumsq_kernel<<<nblk, 256, 0, (cudaStream_t)stream>>>(g, count, partial);
  int rc = check_launch("sumsq");
  if (rc) return rc;
  sumsq_final_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(partial, nblk, out);
  return check_launch("sumsq_final");
}
