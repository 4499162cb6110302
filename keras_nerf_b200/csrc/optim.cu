// a13: Keras Adam; a12: MSE reductions  (keras_nerf/model/nerf/nerf.py:163-165,455-471)
#include "common.cuh"

namespace knerf {

// theta -= lr_t * m/(sqrt(v)+eps), lr_t = lr*sqrt(1-b2^t)/(1-b1^t)  [TF-sem: Keras Adam, no amsgrad]
__global__ void __launch_bounds__(256)
adam_kernel(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
            int64_t n, float lr_t, float one_minus_b1, float one_minus_b2, float eps, int zero_grads) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float gi = g[i];
    const float mi = m[i] + (gi - m[i]) * one_minus_b1;
    const float vi = v[i] + (gi * gi - v[i]) * one_minus_b2;
    m[i] = mi;
    v[i] = vi;
    p[i] = p[i] - lr_t * mi / (sqrtf(vi) + eps);
    if (zero_grads) g[i] = 0.f;
  }
}

// out[0] += scale * sum(x[0..n))   -- single block, deterministic order
__global__ void __launch_bounds__(1024) sum_scale_kernel(const float* __restrict__ x, int64_t n, float scale,
                                                         float* __restrict__ out, int accumulate) {
  __shared__ float red[32];
  float s = 0.f;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) s += x[i];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    float t = (threadIdx.x < (blockDim.x >> 5)) ? red[threadIdx.x] : 0.f;
    t = warp_sum(t);
    if (threadIdx.x == 0) out[0] = (accumulate ? out[0] : 0.f) + scale * t;
  }
}

// out[0] = sum((a-b)^2) / n   -- single block, deterministic order (metrics path, off the hot loop)
__global__ void __launch_bounds__(1024) mse_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                   int64_t n, float* __restrict__ out) {
  __shared__ float red[32];
  float s = 0.f;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
    const float d = a[i] - b[i];
    s += d * d;
  }
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    float t = red[threadIdx.x];
    t = warp_sum(t);
    if (threadIdx.x == 0) out[0] = t / (float)n;
  }
}

int launch_sum_scale(const float* x, int64_t n, float scale, float* out, int accumulate, cudaStream_t st) {
  sum_scale_kernel<<<1, 1024, 0, st>>>(x, n, scale, out, accumulate);
  KN_LAUNCH_CHECK();
  return KNERF_OK;
}

}  // namespace knerf

using namespace knerf;

extern "C" int knerf_adam_step(float* params, float* grads, float* m, float* v, int64_t n, float lr, float beta1,
                               float beta2, float epsilon, int64_t step, int zero_grads, void* stream) {
  KN_CHECK_ARG(params && grads && m && v && n >= 0 && step >= 1, "knerf_adam_step: bad arguments");
  if (n == 0) return KNERF_OK;
  const double t = (double)step;
  const float lr_t = (float)((double)lr * sqrt(1.0 - pow((double)beta2, t)) / (1.0 - pow((double)beta1, t)));
  const int grid = (int)std::min<int64_t>(cdiv(n, 256), (int64_t)kNumSMs * 8);
  adam_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(params, grads, m, v, n, lr_t, (float)(1.0 - (double)beta1),
                                                      (float)(1.0 - (double)beta2), epsilon, zero_grads);
  KN_LAUNCH_CHECK();
  return KNERF_OK;
}

extern "C" int knerf_mse(const float* a, const float* b, int64_t n, float* out, void* stream) {
  KN_CHECK_ARG(a && b && out && n > 0, "knerf_mse: bad arguments");
  mse_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(a, b, n, out);
  KN_LAUNCH_CHECK();
  return KNERF_OK;
}
