// Optimiser steps on the flat parameter / gradient buffers.
// Replaces, for the benchmarked search step, torch.nn.utils.clip_grad_norm_ + torch.optim.SGD
// (generalNAS_tools/train_and_validate.py:146-152, train_genomicDARTS.py:271-276) and the
// torch.optim.Adam over the alphas owned by Architect (darts_tools/architect.py:42,66).
#include "common.cuh"
#include "../../include/gnas.h"

namespace gnas {

static __global__ void __launch_bounds__(256) k_sumsq(const float* __restrict__ g, long long n, double* out) {
  double s = 0.0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float v = g[i];
    s += (double)v * (double)v;
  }
  s = warp_sum_d(s);
  __shared__ double red[8];
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int i = 0; i < 8; ++i) t += red[i];
    atomicAdd(out, t);
  }
}

// p <- p - lr * buf,  buf = momentum*buf + d (first step: buf = d),  d = coef*g + wd*p,
// coef = min(1, clip / (sqrt(sumsq) + 1e-6))  -- torch clip_grad_norm_ + SGD
static __global__ void __launch_bounds__(256) k_sgd(float* __restrict__ p, const float* __restrict__ g,
                                                    float* __restrict__ mom, long long n, float lr, float wd,
                                                    float momentum, int first, const double* sumsq, float clip) {
  float coef = 1.f;
  if (sumsq != nullptr) {
    const float total = (float)sqrt(*sumsq);
    const float c = clip / (total + 1e-6f);
    coef = c < 1.f ? c : 1.f;
  }
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float pv = p[i];
    float d = fmaf(wd, pv, coef * g[i]);
    if (mom != nullptr) {
      const float b = first ? d : fmaf(momentum, mom[i], d);
      mom[i] = b;
      d = b;
    }
    p[i] = pv - lr * d;
  }
}

static __global__ void k_inc_step(int* step) { if (threadIdx.x == 0 && blockIdx.x == 0) *step += 1; }

static __global__ void __launch_bounds__(256) k_adam(float* __restrict__ p, const float* __restrict__ g,
                                                     float* __restrict__ m, float* __restrict__ v, long long n,
                                                     float lr, float wd, float b1, float b2, float eps,
                                                     float bc1, float bc2_sqrt, const int* step_dev,
                                                     const double* sumsq, float clip) {
  float coef = 1.f;          // clip_grad_norm_ over the architecture parameters (train_and_validate.py:97)
  if (sumsq != nullptr) {
    const float c = clip / ((float)sqrt(*sumsq) + 1e-6f);
    coef = c < 1.f ? c : 1.f;
  }
  if (step_dev != nullptr) {   // step counter lives on the device so that a captured CUDA graph stays valid
    const float t = (float)*step_dev;
    bc1 = 1.f - powf(b1, t);
    bc2_sqrt = sqrtf(1.f - powf(b2, t));
  }
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float pv = p[i];
    const float gv = fmaf(wd, pv, coef * g[i]);
    const float mv = b1 * m[i] + (1.f - b1) * gv;
    const float vv = b2 * v[i] + (1.f - b2) * gv * gv;
    m[i] = mv; v[i] = vv;
    const float denom = sqrtf(vv) / bc2_sqrt + eps;
    p[i] = pv - (lr / bc1) * (mv / denom);
  }
}

static inline int blocks_for(long long n) {
  long long b = (n + 255) / 256;
  return (int)(b < 1 ? 1 : (b > 148 * 8 ? 148 * 8 : b));
}

}  // namespace gnas

using namespace gnas;

extern "C" int gnas_grad_sumsq(const float* g, int64_t n, double* out, void* stream) {
  if (!g || !out || n < 0) { set_error("grad_sumsq: bad argument"); return GNAS_ERR_ARG; }
  if (n == 0) return 0;
  GNAS_LAUNCH(k_sumsq, dim3(blocks_for(n)), dim3(256), 0, (cudaStream_t)stream, g, (long long)n, out);
  return check_launch("gnas_grad_sumsq");
}

extern "C" int gnas_sgd_step(float* p, const float* g, float* mom, int64_t n, float lr, float wd,
                             float momentum, int first_step, const double* sumsq, float clip, void* stream) {
  if (!p || !g || n < 0) { set_error("sgd_step: bad argument"); return GNAS_ERR_ARG; }
  if (n == 0) return 0;
  GNAS_LAUNCH(k_sgd, dim3(blocks_for(n)), dim3(256), 0, (cudaStream_t)stream, p, g, mom, (long long)n, lr, wd,
              momentum, first_step, sumsq, clip);
  return check_launch("gnas_sgd_step");
}

extern "C" int gnas_adam_step(float* p, const float* g, float* m, float* v, int64_t n, float lr, float wd,
                              float beta1, float beta2, float eps, int step, int* step_dev,
                              const double* sumsq, float clip, void* stream) {
  if (!p || !g || !m || !v || n < 0 || (step < 1 && !step_dev)) { set_error("adam_step: bad argument"); return GNAS_ERR_ARG; }
  if (n == 0) return 0;
  const float bc1 = 1.f - powf(beta1, (float)(step < 1 ? 1 : step));
  const float bc2 = sqrtf(1.f - powf(beta2, (float)(step < 1 ? 1 : step)));
  if (step_dev) GNAS_LAUNCH(k_inc_step, dim3(1), dim3(32), 0, (cudaStream_t)stream, step_dev);
  GNAS_LAUNCH(k_adam, dim3(blocks_for(n)), dim3(256), 0, (cudaStream_t)stream, p, g, m, v, (long long)n, lr, wd,
              beta1, beta2, eps, bc1, bc2, (const int*)step_dev, sumsq, clip);
  return check_launch("gnas_adam_step");
}
