// DynamicScatterVFE of the GeoMAE configs as a handful of fused kernels:
//   decorate(11) -> Linear(11->64, no bias) -> BN -> ReLU -> scatter-max -> [point | pillar-max] ->
//   Linear(128->128) [tcgen05, sra_layer.cu] -> BN -> ReLU -> scatter-max          (voxel_encoder.py:358-419,
//   utils.py:107-144, ops/norm.py:55-86), forward and backward.
//
// What is fused away compared with the op-by-op path (library sgemm, ATen batch-norm x2 kernels, clamp, 5-launch
// scatter, index_select, cat and their autograd mirrors: ~45 launches, ~20 full passes over [P,128] rows):
//   * the 11-channel decorated features are never written: layer 0 recomputes them from the raw 20-byte records
//     in forward and in its weight-gradient pass;
//   * BatchNorm is two numbers per channel: the kernels that consume a pre-BN tensor apply scale/shift + ReLU on
//     the fly, so post-BN / post-ReLU tensors are never materialised;
//   * scatter-max is ONE pass of 64-bit atomicMax on (order-preserving value key << 32 | ~point index): the high
//     word is the max, the low word the arg-max with the smallest-index tie rule of the in-repo op
//     (scatter_points_cuda.cu:154-158);
//   * BatchNorm statistics (sum, sum of squares; and the two backward sums) are accumulated in fp64 by the kernel
//     that already streams the tensor; the cross-rank exchange of naiveSyncBN1d is a 2C-double all-reduce on the
//     host side between two launches (equal weight per rank, the reference's rule).
// HBM-bound streaming + L2 atomics: no tensor cores here by design (the 128->128 layer is the only GEMM).
#include "common.cuh"

namespace {

constexpr int TPB = 256;
constexpr int C0 = 64;       // layer-0 output channels
constexpr int F0 = 11;       // decorated input channels (5 raw + 3 cluster offset + 3 voxel-centre offset)
constexpr int C1 = 128;      // layer-1 channels

__device__ __forceinline__ uint32_t f2key_pos(float f) {   // f >= 0 (post-ReLU): order-preserving, > 0 for every input
  return __float_as_uint(f) | 0x80000000u;
}
__device__ __forceinline__ float key2f_pos(uint32_t k) { return __uint_as_float(k & 0x7fffffffu); }

struct Decor {
  const float* pts; int64_t n; int stride;
  const int32_t* point_pillar; const float* pillar_mean; const int32_t* pillar_coors;
  float vx, vy, vz, ox, oy, oz;
};

// decorated features of point p (voxel_encoder.py:371-398): [raw channels | xyz - pillar mean | xyz - pillar centre]
__device__ __forceinline__ void decorate(const Decor& d, int64_t p, float* f) {
  const float* src = d.pts + p * d.stride;
  const int pid = __ldg(d.point_pillar + p);
  const float4 mean = __ldg(reinterpret_cast<const float4*>(d.pillar_mean) + pid);
  const int4 pc = __ldg(reinterpret_cast<const int4*>(d.pillar_coors) + pid);
  const float x = __ldg(src), y = __ldg(src + 1), z = __ldg(src + 2);
  f[0] = x; f[1] = y; f[2] = z; f[3] = __ldg(src + 3); f[4] = __ldg(src + 4);
  f[5] = __fsub_rn(x, mean.x);
  f[6] = __fsub_rn(y, mean.y);
  f[7] = __fsub_rn(z, mean.z);
  f[8] = __fsub_rn(x, __fadd_rn(__fmul_rn((float)pc.w, d.vx), d.ox));
  f[9] = __fsub_rn(y, __fadd_rn(__fmul_rn((float)pc.z, d.vy), d.oy));
  f[10] = __fsub_rn(z, __fadd_rn(__fmul_rn((float)pc.y, d.vz), d.oz));
}

// BatchNorm of one layer as per-channel (scale, shift, mean) in shared memory, from the global moments
// mom = [E x (C) | E x^2 (C)] (fp64, already averaged over ranks), gamma, beta.
struct BNArgs { const double* mom; const float* gamma; const float* beta; float eps; };

template <int C>
__device__ __forceinline__ void bn_setup(const BNArgs& b, float* s_scale, float* s_shift, float* s_mean) {
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const double mu = b.mom[c], var = b.mom[C + c] - mu * mu;
    const float sc = __ldg(b.gamma + c) * (float)rsqrt(var + (double)b.eps);
    s_scale[c] = sc;
    s_shift[c] = __ldg(b.beta + c) - (float)mu * sc;
    if (s_mean) s_mean[c] = (float)mu;
  }
  __syncthreads();
}

// ------------------------------------------------------------------------------------------------ forward, layer 0
// x1[p, c] = sum_k W0[c, k] f_k(p);  stats[c] += x1, stats[C0 + c] += x1^2.
// Phase A: one thread decorates one point into shared memory.  Phase B: one warp walks 32 points, lane l owns output
// channels l and l + 32 (their W0 rows live in registers), so every store is a full 128-byte line and the
// statistics need no cross-lane reduction.
__global__ void __launch_bounds__(TPB) k_vfe0_fwd(Decor d, const float* __restrict__ W0, float* x1, double* stats) {
  __shared__ float sf[TPB][F0 + 1];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float w0[F0], w1[F0];
#pragma unroll
  for (int k = 0; k < F0; ++k) { w0[k] = __ldg(W0 + lane * F0 + k); w1[k] = __ldg(W0 + (lane + 32) * F0 + k); }
  float s0 = 0.f, q0 = 0.f, s1 = 0.f, q1 = 0.f;
  for (int64_t base = (int64_t)blockIdx.x * TPB; base < d.n; base += (int64_t)gridDim.x * TPB) {
    __syncthreads();
    const int64_t p = base + threadIdx.x;
    if (p < d.n) {
      float f[F0];
      decorate(d, p, f);
#pragma unroll
      for (int k = 0; k < F0; ++k) sf[threadIdx.x][k] = f[k];
    }
    __syncthreads();
    const int cnt = (int)min((int64_t)32, d.n - (base + warp * 32));
    for (int i = 0; i < cnt; ++i) {
      const float* f = sf[warp * 32 + i];
      float a0 = 0.f, a1 = 0.f;
#pragma unroll
      for (int k = 0; k < F0; ++k) { a0 = fmaf(w0[k], f[k], a0); a1 = fmaf(w1[k], f[k], a1); }
      float* o = x1 + (base + warp * 32 + i) * C0;
      o[lane] = a0;
      o[lane + 32] = a1;
      s0 += a0; q0 = fmaf(a0, a0, q0); s1 += a1; q1 = fmaf(a1, a1, q1);
    }
  }
  atomicAdd(stats + lane, (double)s0);
  atomicAdd(stats + lane + 32, (double)s1);
  atomicAdd(stats + C0 + lane, (double)q0);
  atomicAdd(stats + C0 + lane + 32, (double)q1);
}

// ------------------------------------------------------------------------------------------------ column statistics
// stats[c] += sum_p x[p, c], stats[C1 + c] += sum_p x^2   (x: [n, 128]); warp per row, float4 per lane.
__global__ void __launch_bounds__(TPB) k_colstats128(const float* __restrict__ x, int64_t n, double* stats) {
  __shared__ float red[TPB / 32][2][C1];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f), q = s;
  for (int64_t r = (int64_t)blockIdx.x * (TPB / 32) + warp; r < n; r += (int64_t)gridDim.x * (TPB / 32)) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(x + r * C1) + lane);
    s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    q.x = fmaf(v.x, v.x, q.x); q.y = fmaf(v.y, v.y, q.y); q.z = fmaf(v.z, v.z, q.z); q.w = fmaf(v.w, v.w, q.w);
  }
  reinterpret_cast<float4*>(red[warp][0])[lane] = s;
  reinterpret_cast<float4*>(red[warp][1])[lane] = q;
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * C1; i += TPB) {
    double t = 0.0;
#pragma unroll
    for (int w = 0; w < TPB / 32; ++w) t += (double)red[w][i / C1][i % C1];
    atomicAdd(stats + i, t);
  }
}

// ------------------------------------------------------------------------------------------------ BN + ReLU + scatter-max
// vmax[pillar, c] = max over the pillar's points of (key(relu(bn(x[p, c]))) << 32 | ~p).  Block 0 also updates the
// running statistics (momentum rule of nn.BatchNorm1d; `unbias` = N/(N-1) on one rank, 1 in the synchronised path).
template <int C>
__global__ void __launch_bounds__(TPB) k_bn_relu_max(const float* __restrict__ x, int64_t n, BNArgs bn,
                                                     const int32_t* __restrict__ point_pillar,
                                                     unsigned long long* vmax, float* running_mean, float* running_var,
                                                     float momentum, float unbias) {
  __shared__ float s_scale[C], s_shift[C];
  bn_setup<C>(bn, s_scale, s_shift, nullptr);
  if (blockIdx.x == 0 && running_mean)
    for (int c = threadIdx.x; c < C; c += TPB) {
      const double mu = bn.mom[c], var = bn.mom[C + c] - mu * mu;
      running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)mu;
      running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)var * unbias;
    }
  constexpr int Q = C / 4;                 // float4 per row
  const int64_t total = n * Q;
  for (int64_t i = (int64_t)blockIdx.x * TPB + threadIdx.x; i < total; i += (int64_t)gridDim.x * TPB) {
    const int64_t p = i / Q;
    const int c = (int)(i - p * Q) * 4;
    const int pid = __ldg(point_pillar + p);
    const float4 v = __ldg(reinterpret_cast<const float4*>(x) + i);
    const float y[4] = {fmaxf(fmaf(v.x, s_scale[c], s_shift[c]), 0.f), fmaxf(fmaf(v.y, s_scale[c + 1], s_shift[c + 1]), 0.f),
                        fmaxf(fmaf(v.z, s_scale[c + 2], s_shift[c + 2]), 0.f), fmaxf(fmaf(v.w, s_scale[c + 3], s_shift[c + 3]), 0.f)};
    unsigned long long* dst = vmax + (int64_t)pid * C + c;
    const unsigned long long tag = 0xffffffffull - (unsigned long long)p;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const unsigned long long val = ((unsigned long long)f2key_pos(y[k]) << 32) | tag;
      if (*(volatile unsigned long long*)(dst + k) < val) atomicMax(dst + k, val);
    }
  }
}

// feat1[p] = [relu(bn0(x1[p])) (64) | max of the point's pillar (64)]      (the layer-1 GEMM operand)
__global__ void __launch_bounds__(TPB) k_vfe_cat(const float* __restrict__ x1, int64_t n, BNArgs bn,
                                                 const int32_t* __restrict__ point_pillar,
                                                 const unsigned long long* __restrict__ vmax1, float* feat1) {
  __shared__ float s_scale[C0], s_shift[C0];
  bn_setup<C0>(bn, s_scale, s_shift, nullptr);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int64_t p = (int64_t)blockIdx.x * (TPB / 32) + warp; p < n; p += (int64_t)gridDim.x * (TPB / 32)) {
    float4 o;
    if (lane < 16) {
      const int c = lane * 4;
      const float4 v = __ldg(reinterpret_cast<const float4*>(x1 + p * C0) + lane);
      o = make_float4(fmaxf(fmaf(v.x, s_scale[c], s_shift[c]), 0.f), fmaxf(fmaf(v.y, s_scale[c + 1], s_shift[c + 1]), 0.f),
                      fmaxf(fmaf(v.z, s_scale[c + 2], s_shift[c + 2]), 0.f), fmaxf(fmaf(v.w, s_scale[c + 3], s_shift[c + 3]), 0.f));
    } else {
      const int pid = __ldg(point_pillar + p);
      const ulonglong2* src = reinterpret_cast<const ulonglong2*>(vmax1 + (int64_t)pid * C0 + (lane - 16) * 4);
      const ulonglong2 a = __ldg(src), b = __ldg(src + 1);
      o = make_float4(key2f_pos((uint32_t)(a.x >> 32)), key2f_pos((uint32_t)(a.y >> 32)), key2f_pos((uint32_t)(b.x >> 32)),
                      key2f_pos((uint32_t)(b.y >> 32)));
    }
    reinterpret_cast<float4*>(feat1 + p * C1)[lane] = o;
  }
}

__global__ void __launch_bounds__(TPB) k_vmax_decode(const unsigned long long* __restrict__ vmax, int64_t total, float* out) {
  for (int64_t i = (int64_t)blockIdx.x * TPB + threadIdx.x; i < total; i += (int64_t)gridDim.x * TPB)
    out[i] = key2f_pos((uint32_t)(__ldg(vmax + i) >> 32));
}

// ------------------------------------------------------------------------------------------------ backward, layer 1
// g[p, c] = (relu(bn(x)) > 0) * (arg-max of the pillar is p ? d_vox[pillar, c] : 0)
__device__ __forceinline__ float routed(const unsigned long long v, int64_t p, float dv) {
  return (uint32_t)v == (uint32_t)(0xffffffffull - (unsigned long long)p) ? dv : 0.f;
}

// MODE 0: sums[c] += g, sums[C1 + c] += g * (x - mean).   MODE 1: dx = scale * g + (A + 2 B x) * inv_wn.
template <int MODE>
__global__ void __launch_bounds__(TPB) k_vfe1_bwd(const float* __restrict__ x2, int64_t n, BNArgs bn,
                                                  const int32_t* __restrict__ point_pillar,
                                                  const unsigned long long* __restrict__ vmax2,
                                                  const float* __restrict__ d_vox, double* sums,
                                                  const double* __restrict__ ab, float inv_wn, float* dx2) {
  __shared__ float s_scale[C1], s_shift[C1], s_mean[C1];
  __shared__ float red[MODE == 0 ? TPB / 32 : 1][2][C1];
  __shared__ float s_a[MODE == 1 ? C1 : 1], s_b[MODE == 1 ? C1 : 1];
  if (MODE == 1)
    for (int c = threadIdx.x; c < C1; c += TPB) { s_a[c] = (float)ab[c] * inv_wn; s_b[c] = 2.f * (float)ab[C1 + c] * inv_wn; }
  bn_setup<C1>(bn, s_scale, s_shift, s_mean);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c = lane * 4;
  float4 sg = make_float4(0.f, 0.f, 0.f, 0.f), sgx = sg;
  for (int64_t p = (int64_t)blockIdx.x * (TPB / 32) + warp; p < n; p += (int64_t)gridDim.x * (TPB / 32)) {
    const int pid = __ldg(point_pillar + p);
    const float4 v = __ldg(reinterpret_cast<const float4*>(x2 + p * C1) + lane);
    const float4 dv = __ldg(reinterpret_cast<const float4*>(d_vox + (int64_t)pid * C1) + lane);
    const ulonglong2* am = reinterpret_cast<const ulonglong2*>(vmax2 + (int64_t)pid * C1 + c);
    const ulonglong2 a0 = __ldg(am), a1 = __ldg(am + 1);
    const float xs[4] = {v.x, v.y, v.z, v.w};
    const float up[4] = {routed(a0.x, p, dv.x), routed(a0.y, p, dv.y), routed(a1.x, p, dv.z), routed(a1.y, p, dv.w)};
    float g[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) g[k] = fmaf(xs[k], s_scale[c + k], s_shift[c + k]) > 0.f ? up[k] : 0.f;
    if (MODE == 0) {
      sg.x += g[0]; sg.y += g[1]; sg.z += g[2]; sg.w += g[3];
      sgx.x = fmaf(g[0], xs[0] - s_mean[c], sgx.x); sgx.y = fmaf(g[1], xs[1] - s_mean[c + 1], sgx.y);
      sgx.z = fmaf(g[2], xs[2] - s_mean[c + 2], sgx.z); sgx.w = fmaf(g[3], xs[3] - s_mean[c + 3], sgx.w);
    } else {
      float4 o;
      o.x = fmaf(s_scale[c], g[0], fmaf(s_b[c], xs[0], s_a[c]));
      o.y = fmaf(s_scale[c + 1], g[1], fmaf(s_b[c + 1], xs[1], s_a[c + 1]));
      o.z = fmaf(s_scale[c + 2], g[2], fmaf(s_b[c + 2], xs[2], s_a[c + 2]));
      o.w = fmaf(s_scale[c + 3], g[3], fmaf(s_b[c + 3], xs[3], s_a[c + 3]));
      reinterpret_cast<float4*>(dx2 + p * C1)[lane] = o;
    }
  }
  if (MODE == 0) {
    reinterpret_cast<float4*>(red[warp][0])[lane] = sg;
    reinterpret_cast<float4*>(red[warp][1])[lane] = sgx;
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * C1; i += TPB) {
      double t = 0.0;
#pragma unroll
      for (int w = 0; w < TPB / 32; ++w) t += (double)red[w][i / C1][i % C1];
      atomicAdd(sums + i, t);
    }
  }
}

// BatchNorm backward, the per-channel part.  From sums = [sum g | sum g (x - mean)]:
//   d_gamma += sum g xhat, d_beta += sum g,
//   ab = [dL/d(mean) | dL/d(mean of squares)] of THIS rank (all-reduced over ranks by the host before k_*_bwd<1>):
//   a = -s sum g + s mean / (var + eps) * sum g (x - mean),   b = -s / (2 (var + eps)) * sum g (x - mean).
__global__ void k_bn_ab(int C, const double* __restrict__ sums, BNArgs bn, double* ab, float* d_gamma, float* d_beta) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const double mu = bn.mom[c], var = bn.mom[C + c] - mu * mu, ve = var + (double)bn.eps;
  const double rstd = rsqrt(ve), s = (double)bn.gamma[c] * rstd;
  const double sg = sums[c], sgx = sums[C + c];
  d_gamma[c] += (float)(sgx * rstd);
  d_beta[c] += (float)sg;
  ab[c] = -s * sg + s * mu / ve * sgx;
  ab[C + c] = -0.5 * s / ve * sgx;
}

// d_vmax1[pillar, c] += dfeat1[p, 64 + c]   (backward of the pillar-max gather of layer 1's operand)
__global__ void __launch_bounds__(TPB) k_vfe_gather_bwd(const float* __restrict__ dfeat1, int64_t n,
                                                        const int32_t* __restrict__ point_pillar, float* d_vmax1) {
  const int64_t total = n * 16;
  for (int64_t i = (int64_t)blockIdx.x * TPB + threadIdx.x; i < total; i += (int64_t)gridDim.x * TPB) {
    const int64_t p = i >> 4;
    const int q = (int)(i & 15);
    const int pid = __ldg(point_pillar + p);
    const float4 v = __ldg(reinterpret_cast<const float4*>(dfeat1 + p * C1 + C0) + q);
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(d_vmax1 + (int64_t)pid * C0 + q * 4), "f"(v.x),
                 "f"(v.y), "f"(v.z), "f"(v.w)
                 : "memory");
  }
}

// ------------------------------------------------------------------------------------------------ backward, layer 0
// up[p, c] = dfeat1[p, c] + (arg-max of the pillar is p ? d_vmax1[pillar, c] : 0);  g = (relu(bn0(x1)) > 0) * up.
// MODE 0: the two BatchNorm sums.  MODE 1: dx1 = scale g + (A + 2 B x1) inv_wn, consumed on the spot by the weight
// gradient dW0[c, k] += sum_p dx1[p, c] f_k(p) with the decorated features recomputed (dx1 is never written: raw
// points carry no gradient).  Same lane = channel-pair mapping as k_vfe0_fwd.
template <int MODE>
__global__ void __launch_bounds__(TPB) k_vfe0_bwd(Decor d, const float* __restrict__ x1, BNArgs bn,
                                                  const unsigned long long* __restrict__ vmax1,
                                                  const float* __restrict__ d_vmax1, const float* __restrict__ dfeat1,
                                                  double* sums, const double* __restrict__ ab, float inv_wn, float* dW0) {
  __shared__ float sf[MODE == 1 ? TPB : 1][F0 + 1];
  __shared__ float s_scale[C0], s_shift[C0], s_mean[C0];
  __shared__ float red[MODE == 1 ? TPB / 32 : 1][C0 * F0];
  bn_setup<C0>(bn, s_scale, s_shift, s_mean);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float sc0 = s_scale[lane], sh0 = s_shift[lane], mu0 = s_mean[lane];
  const float sc1 = s_scale[lane + 32], sh1 = s_shift[lane + 32], mu1 = s_mean[lane + 32];
  float a0 = 0.f, b0 = 0.f, a1 = 0.f, b1 = 0.f;
  if (MODE == 1) {
    a0 = (float)ab[lane] * inv_wn; b0 = 2.f * (float)ab[C0 + lane] * inv_wn;
    a1 = (float)ab[lane + 32] * inv_wn; b1 = 2.f * (float)ab[C0 + lane + 32] * inv_wn;
  }
  float acc0[F0], acc1[F0];
#pragma unroll
  for (int k = 0; k < F0; ++k) { acc0[k] = 0.f; acc1[k] = 0.f; }
  float sg0 = 0.f, sgx0 = 0.f, sg1 = 0.f, sgx1 = 0.f;
  for (int64_t base = (int64_t)blockIdx.x * TPB; base < d.n; base += (int64_t)gridDim.x * TPB) {
    if (MODE == 1) {
      __syncthreads();
      const int64_t p = base + threadIdx.x;
      if (p < d.n) {
        float f[F0];
        decorate(d, p, f);
#pragma unroll
        for (int k = 0; k < F0; ++k) sf[threadIdx.x][k] = f[k];
      }
      __syncthreads();
    }
    const int cnt = (int)min((int64_t)32, d.n - (base + warp * 32));
    for (int i = 0; i < cnt; ++i) {
      const int64_t p = base + warp * 32 + i;
      const int pid = __ldg(d.point_pillar + p);
      const float v0 = __ldg(x1 + p * C0 + lane), v1 = __ldg(x1 + p * C0 + lane + 32);
      float u0 = __ldg(dfeat1 + p * C1 + lane), u1 = __ldg(dfeat1 + p * C1 + lane + 32);
      u0 += routed(__ldg(vmax1 + (int64_t)pid * C0 + lane), p, __ldg(d_vmax1 + (int64_t)pid * C0 + lane));
      u1 += routed(__ldg(vmax1 + (int64_t)pid * C0 + lane + 32), p, __ldg(d_vmax1 + (int64_t)pid * C0 + lane + 32));
      const float g0 = fmaf(v0, sc0, sh0) > 0.f ? u0 : 0.f, g1 = fmaf(v1, sc1, sh1) > 0.f ? u1 : 0.f;
      if (MODE == 0) {
        sg0 += g0; sgx0 = fmaf(g0, v0 - mu0, sgx0);
        sg1 += g1; sgx1 = fmaf(g1, v1 - mu1, sgx1);
      } else {
        const float dx0 = fmaf(sc0, g0, fmaf(b0, v0, a0)), dx1 = fmaf(sc1, g1, fmaf(b1, v1, a1));
        const float* f = sf[warp * 32 + i];
#pragma unroll
        for (int k = 0; k < F0; ++k) { acc0[k] = fmaf(dx0, f[k], acc0[k]); acc1[k] = fmaf(dx1, f[k], acc1[k]); }
      }
    }
  }
  if (MODE == 0) {
    atomicAdd(sums + lane, (double)sg0);
    atomicAdd(sums + lane + 32, (double)sg1);
    atomicAdd(sums + C0 + lane, (double)sgx0);
    atomicAdd(sums + C0 + lane + 32, (double)sgx1);
  } else {
#pragma unroll
    for (int k = 0; k < F0; ++k) { red[warp][lane * F0 + k] = acc0[k]; red[warp][(lane + 32) * F0 + k] = acc1[k]; }
    __syncthreads();
    for (int i = threadIdx.x; i < C0 * F0; i += TPB) {
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < TPB / 32; ++w) t += red[w][i];
      atomicAdd(dW0 + i, t);
    }
  }
}

inline int grid_rows(int64_t rows_per_block_units) {
  const int64_t b = rows_per_block_units;
  return (int)(b < 1 ? 1 : (b < GM_NUM_SMS * 8 ? b : GM_NUM_SMS * 8));
}

Decor make_decor(const float* points, int64_t n, int32_t channels, const int32_t* point_pillar, const float* pillar_mean,
                 const int32_t* pillar_coors, const float* v, const float* o) {
  Decor d;
  d.pts = points; d.n = n; d.stride = channels; d.point_pillar = point_pillar; d.pillar_mean = pillar_mean;
  d.pillar_coors = pillar_coors; d.vx = v[0]; d.vy = v[1]; d.vz = v[2]; d.ox = o[0]; d.oy = o[1]; d.oz = o[2];
  return d;
}

}  // namespace

extern "C" int geomae_vfe0_forward(const float* points, int64_t n, int32_t channels, const int32_t* point_pillar,
                                   const float* pillar_mean, const int32_t* pillar_coors, const float voxel_xyz[3],
                                   const float centre_offset_xyz[3], const float* W0, float* x1, double* stats,
                                   void* stream_) {
  GM_REQUIRE(channels == 5, "vfe0_forward: built for 5 raw channels (x,y,z,intensity,dt), got %d", channels);
  GM_REQUIRE(stats, "vfe0_forward: null stats");
  cudaStream_t stream = (cudaStream_t)stream_;
  GM_CUDA(cudaMemsetAsync(stats, 0, 2 * C0 * sizeof(double), stream));
  if (n == 0) return GEOMAE_OK;
  GM_REQUIRE(points && point_pillar && pillar_mean && pillar_coors && voxel_xyz && centre_offset_xyz && W0 && x1,
             "vfe0_forward: null argument");
  const Decor d = make_decor(points, n, channels, point_pillar, pillar_mean, pillar_coors, voxel_xyz, centre_offset_xyz);
  k_vfe0_fwd<<<grid_rows(gm_div_up(n, TPB)), TPB, 0, stream>>>(d, W0, x1, stats);
  GM_LAUNCH_CHECK();
  return GEOMAE_OK;
}

extern "C" int geomae_colstats(const float* x, int64_t n, int32_t channels, double* stats, void* stream_) {
  GM_REQUIRE(channels == C1, "colstats: built for 128 channels (got %d)", channels);
  GM_REQUIRE(stats, "colstats: null stats");
  cudaStream_t stream = (cudaStream_t)stream_;
  GM_CUDA(cudaMemsetAsync(stats, 0, 2 * C1 * sizeof(double), stream));
  if (n == 0) return GEOMAE_OK;
  GM_REQUIRE(x, "colstats: null input");
  k_colstats128<<<grid_rows(gm_div_up(n, 8 * 8)), TPB, 0, stream>>>(x, n, stats);
  GM_LAUNCH_CHECK();
  return GEOMAE_OK;
}

extern "C" int geomae_vfe_bn_relu_max(const float* x, int64_t n, int32_t channels, const int32_t* point_pillar,
                                      const double* mom, const float* gamma, const float* beta, float eps,
                                      float* running_mean, float* running_var, float momentum, float unbias,
                                      uint64_t* vmax, int64_t n_pillars, void* stream_) {
  GM_REQUIRE(channels == C0 || channels == C1, "vfe_bn_relu_max: 64 or 128 channels (got %d)", channels);
  GM_REQUIRE(mom && gamma && beta && (vmax || n_pillars == 0), "vfe_bn_relu_max: null argument");
  cudaStream_t stream = (cudaStream_t)stream_;
  if (n_pillars > 0) GM_CUDA(cudaMemsetAsync(vmax, 0, (size_t)n_pillars * channels * 8, stream));
  GM_REQUIRE(n == 0 || (x && point_pillar), "vfe_bn_relu_max: null input");
  const BNArgs bn{mom, gamma, beta, eps};
  const int grid = grid_rows(gm_div_up(n * (channels / 4), TPB * 2));
  unsigned long long* vm = reinterpret_cast<unsigned long long*>(vmax);
  if (channels == C0)
    k_bn_relu_max<C0><<<grid, TPB, 0, stream>>>(x, n, bn, point_pillar, vm, running_mean, running_var, momentum, unbias);
  else
    k_bn_relu_max<C1><<<grid, TPB, 0, stream>>>(x, n, bn, point_pillar, vm, running_mean, running_var, momentum, unbias);
  GM_LAUNCH_CHECK();
  return GEOMAE_OK;
}

extern "C" int geomae_vfe_cat(const float* x1, int64_t n, const int32_t* point_pillar, const double* mom,
                              const float* gamma, const float* beta, float eps, const uint64_t* vmax1, float* feat1,
                              void* stream) {
  if (n == 0) return GEOMAE_OK;
  GM_REQUIRE(x1 && point_pillar && mom && gamma && beta && vmax1 && feat1, "vfe_cat: null argument");
  const BNArgs bn{mom, gamma, beta, eps};
  k_vfe_cat<<<grid_rows(gm_div_up(n, 8 * 4)), TPB, 0, (cudaStream_t)stream>>>(
      x1, n, bn, point_pillar, reinterpret_cast<const unsigned long long*>(vmax1), feat1);
  GM_LAUNCH_CHECK();
  return GEOMAE_OK;
}

extern "C" int geomae_vmax_decode(const uint64_t* vmax, int64_t total, float* out, void* stream) {
  if (total == 0) return GEOMAE_OK;
  GM_REQUIRE(vmax && out, "vmax_decode: null argument");
  k_vmax_decode<<<grid_rows(gm_div_up(total, TPB * 4)), TPB, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const unsigned long long*>(vmax), total, out);
  GM_LAUNCH_CHECK();
  return GEOMAE_OK;
}

// mode 0: sums (zeroed here) ; mode 1: dx2
extern "C" int geomae_vfe1_backward(int32_t mode, const float* x2, int64_t n, const int32_t* point_pillar,
                                    const double* mom, const float* gamma, const float* beta, float eps,
                                    const uint64_t* vmax2, const float* d_vox, double* sums, const double* ab,
                                    float inv_wn, float* dx2, void* stream_) {
  GM_REQUIRE(mode == 0 || mode == 1, "vfe1_backward: mode %d", mode);
  cudaStream_t stream = (cudaStream_t)stream_;
  if (mode == 0) {
    GM_REQUIRE(sums, "vfe1_backward: null sums");
    GM_CUDA(cudaMemsetAsync(sums, 0, 2 * C1 * sizeof(double), stream));
  }
  if (n == 0) return GEOMAE_OK;
  GM_REQUIRE(x2 && point_pillar && mom && gamma && beta && vmax2 && d_vox && (mode == 0 || (ab && dx2)),
             "vfe1_backward: null argument");
  const BNArgs bn{mom, gamma, beta, eps};
  const int grid = grid_rows(gm_div_up(n, 8 * 4));
  const unsigned long long* vm = reinterpret_cast<const unsigned long long*>(vmax2);
  if (mode == 0)
    k_vfe1_bwd<0><<<grid, TPB, 0, stream>>>(x2, n, bn, point_pillar, vm, d_vox, sums, nullptr, 0.f, nullptr);
  else
    k_vfe1_bwd<1><<<grid, TPB, 0, stream>>>(x2, n, bn, point_pillar, vm, d_vox, nullptr, ab, inv_wn, dx2);
  GM_LAUNCH_CHECK();
  return GEOMAE_OK;
}

extern "C" int geomae_bn_backward_coeffs(int32_t channels, const double* sums, const double* mom, const float* gamma,
                                         float eps, double* ab, float* d_gamma, float* d_beta, void* stream) {
  GM_REQUIRE(channels > 0 && sums && mom && gamma && ab && d_gamma && d_beta, "bn_backward_coeffs: null argument");
  const BNArgs bn{mom, gamma, nullptr, eps};
  k_bn_ab<<<gm_div_up(channels, 128), 128, 0, (cudaStream_t)stream>>>(channels, sums, bn, ab, d_gamma, d_beta);
  GM_LAUNCH_CHECK();
  return GEOMAE_OK;
}

extern "C" int geomae_vfe_gather_backward(const float* dfeat1, int64_t n, const int32_t* point_pillar, float* d_vmax1,
                                          int64_t n_pillars, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (n_pillars > 0) {
    GM_REQUIRE(d_vmax1, "vfe_gather_backward: null output");
    GM_CUDA(cudaMemsetAsync(d_vmax1, 0, (size_t)n_pillars * C0 * 4, stream));
  }
  if (n == 0) return GEOMAE_OK;
  GM_REQUIRE(dfeat1 && point_pillar, "vfe_gather_backward: null argument");
  k_vfe_gather_bwd<<<grid_rows(gm_div_up(n * 16, TPB * 2)), TPB, 0, stream>>>(dfeat1, n, point_pillar, d_vmax1);
  GM_LAUNCH_CHECK();
  return GEOMAE_OK;
}

// mode 0: sums (zeroed here); mode 1: dW0 += dx1^T f
extern "C" int geomae_vfe0_backward(int32_t mode, const float* points, int64_t n, int32_t channels,
                                    const int32_t* point_pillar, const float* pillar_mean, const int32_t* pillar_coors,
                                    const float voxel_xyz[3], const float centre_offset_xyz[3], const float* x1,
                                    const double* mom, const float* gamma, const float* beta, float eps,
                                    const uint64_t* vmax1, const float* d_vmax1, const float* dfeat1, double* sums,
                                    const double* ab, float inv_wn, float* dW0, void* stream_) {
  GM_REQUIRE(mode == 0 || mode == 1, "vfe0_backward: mode %d", mode);
  GM_REQUIRE(channels == 5, "vfe0_backward: built for 5 raw channels, got %d", channels);
  cudaStream_t stream = (cudaStream_t)stream_;
  if (mode == 0) {
    GM_REQUIRE(sums, "vfe0_backward: null sums");
    GM_CUDA(cudaMemsetAsync(sums, 0, 2 * C0 * sizeof(double), stream));
  }
  if (n == 0) return GEOMAE_OK;
  GM_REQUIRE(points && point_pillar && pillar_mean && pillar_coors && voxel_xyz && centre_offset_xyz && x1 && mom && gamma &&
                 beta && vmax1 && d_vmax1 && dfeat1 && (mode == 0 || (ab && dW0)),
             "vfe0_backward: null argument");
  const Decor d = make_decor(points, n, channels, point_pillar, pillar_mean, pillar_coors, voxel_xyz, centre_offset_xyz);
  const BNArgs bn{mom, gamma, beta, eps};
  const int grid = grid_rows(gm_div_up(n, TPB));
  const unsigned long long* vm = reinterpret_cast<const unsigned long long*>(vmax1);
  if (mode == 0)
    k_vfe0_bwd<0><<<grid, TPB, 0, stream>>>(d, x1, bn, vm, d_vmax1, dfeat1, sums, nullptr, 0.f, nullptr);
  else
    k_vfe0_bwd<1><<<grid, TPB, 0, stream>>>(d, x1, bn, vm, d_vmax1, dfeat1, nullptr, ab, inv_wn, dW0);
  GM_LAUNCH_CHECK();
  return GEOMAE_OK;
}
