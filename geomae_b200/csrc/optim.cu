// Fused optimiser step on the flat parameter / gradient buffers: global-norm gradient clipping
// (mmcv OptimizerHook grad_clip max_norm, torch clip_grad_norm_ semantics) + AdamW, with the
// 1/world_size gradient averaging folded in.  No host synchronisation: the clip coefficient is
// derived on the device from the partial sums.  HBM-bound elementwise work.
#include <math.h>

#include "common.cuh"

namespace {

constexpr int TPB = 256;
constexpr int MAX_PARTIALS = 1024;

__global__ void __launch_bounds__(TPB) k_sumsq(const float* __restrict__ g, int64_t n, double* partials) {
  __shared__ double sm[TPB / 32];
  double acc = 0.0;
  const int64_t n4 = n >> 2;
  const float4* g4 = reinterpret_cast<const float4*>(g);
  for (int64_t i = (int64_t)blockIdx.x * TPB + threadIdx.x; i < n4; i += (int64_t)gridDim.x * TPB) {
    const float4 v = __ldg(g4 + i);
    acc += (double)v.x * v.x + (double)v.y * v.y + (double)v.z * v.z + (double)v.w * v.w;
  }
  if (blockIdx.x == 0)
    for (int64_t i = (n4 << 2) + threadIdx.x; i < n; i += TPB) acc += (double)g[i] * g[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int i = 0; i < TPB / 32; ++i) t += sm[i];
    partials[blockIdx.x] = t;
  }
}

__global__ void __launch_bounds__(TPB) k_adamw(float* __restrict__ p, const float* __restrict__ g,
                                               float* __restrict__ m, float* __restrict__ v, int64_t n,
                                               int64_t n_decay, const double* __restrict__ partials, int n_partials,
                                               float grad_scale, float max_norm, float lr, float beta1, float beta2,
                                               float eps, float weight_decay, float bias1, float bias2_sqrt,
                                               float* stats) {
  __shared__ float s_coef;
  if (threadIdx.x < 32) {
    double t = 0.0;
    for (int i = threadIdx.x; i < n_partials; i += 32) t += partials[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    if (threadIdx.x == 0) {
      const float norm = (float)sqrt(t) * grad_scale;   // norm of the averaged gradient
      float clip = 1.0f;
      if (max_norm > 0.f) clip = fminf(1.0f, max_norm / (norm + 1e-6f));
      s_coef = clip * grad_scale;
      if (blockIdx.x == 0 && stats) { stats[0] = norm; stats[1] = clip; }
    }
  }
  __syncthreads();
  const float coef = s_coef;
  for (int64_t i = (int64_t)blockIdx.x * TPB + threadIdx.x; i < n; i += (int64_t)gridDim.x * TPB) {
    const float gi = g[i] * coef;
    float pi = p[i];
    if (i < n_decay) pi *= 1.0f - lr * weight_decay;
    const float mi = beta1 * m[i] + (1.0f - beta1) * gi;
    const float vi = beta2 * v[i] + (1.0f - beta2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    const float denom = sqrtf(vi) / bias2_sqrt + eps;
    p[i] = pi - (lr / bias1) * (mi / denom);
  }
}

}  // namespace

extern "C" int geomae_adamw_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n,
                                 int64_t n_decay, double* partials, float grad_scale, float max_norm, float lr,
                                 float beta1, float beta2, float eps, float weight_decay, int64_t step, float* stats,
                                 void* stream_) {
  GM_REQUIRE(params && grads && exp_avg && exp_avg_sq && partials, "adamw_step: null argument");
  GM_REQUIRE(n >= 0 && n_decay >= 0 && n_decay <= n && step >= 1, "adamw_step: bad sizes / step");
  if (n == 0) return GEOMAE_OK;
  cudaStream_t stream = (cudaStream_t)stream_;
  int nblk = gm_div_up(n, (int64_t)TPB * 16);
  if (nblk > MAX_PARTIALS) nblk = MAX_PARTIALS;
  k_sumsq<<<nblk, TPB, 0, stream>>>(grads, n, partials);
  // bias corrections in double like torch.optim.AdamW (Python floats): fp32 powf is off by 5e-5 relative at step 1
  const float bias1 = (float)(1.0 - pow((double)beta1, (double)step));
  const float bias2_sqrt = (float)sqrt(1.0 - pow((double)beta2, (double)step));
  int ublk = gm_div_up(n, (int64_t)TPB * 4);
  if (ublk > GM_NUM_SMS * 8) ublk = GM_NUM_SMS * 8;
  k_adamw<<<ublk, TPB, 0, stream>>>(params, grads, exp_avg, exp_avg_sq, n, n_decay, partials, nblk, grad_scale,
                                    max_norm, lr, beta1, beta2, eps, weight_decay, bias1, bias2_sqrt, stats);
  GM_LAUNCH_CHECK();
  return GEOMAE_OK;
}
