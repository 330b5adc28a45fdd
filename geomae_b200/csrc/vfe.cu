// DynamicScatterVFE building blocks: point decoration and the
// point->pillar max/mean reduction with its backward.
//
// The reference calls torch.unique(dim=0) again for every scatter (3 row sorts in the VFE alone,
// unique_once=False) and reduces with torch_scatter atomics.  Here the point->pillar row map comes
// from the bitmap-rank stage, so a scatter is a single pass of vector atomics; the arg-max needed by
// the backward is resolved deterministically (smallest point index among ties, the rule of the
// in-repo op, scatter_points_cuda.cu:154-158).
#include "common.cuh"

namespace {

constexpr int TPB = 256;

// feats[p] = [point channels (C) | xyz - pillar mean | xyz - pillar centre]   (voxel_encoder.py:371-398)
__global__ void __launch_bounds__(TPB) k_decorate(const float* __restrict__ pts, int64_t n, int C,
                                                  const int32_t* __restrict__ point_pillar,
                                                  const float* __restrict__ pillar_mean,
                                                  const int32_t* __restrict__ pillar_coors, float vx, float vy,
                                                  float vz, float ox, float oy, float oz, float* out) {
  const int64_t p = (int64_t)blockIdx.x * TPB + threadIdx.x;
  if (p >= n) return;
  const int W = C + 6;
  const float* src = pts + p * C;
  float* dst = out + p * W;
  const int pid = __ldg(point_pillar + p);
  const float4 mean = __ldg(reinterpret_cast<const float4*>(pillar_mean) + pid);
  const int4 pc = __ldg(reinterpret_cast<const int4*>(pillar_coors) + pid);
  const float x = __ldg(src), y = __ldg(src + 1), z = __ldg(src + 2);
  for (int c = 0; c < C; ++c) dst[c] = __ldg(src + c);
  dst[C + 0] = __fsub_rn(x, mean.x);
  dst[C + 1] = __fsub_rn(y, mean.y);
  dst[C + 2] = __fsub_rn(z, mean.z);
  dst[C + 3] = __fsub_rn(x, __fadd_rn(__fmul_rn((float)pc.w, vx), ox));
  dst[C + 4] = __fsub_rn(y, __fadd_rn(__fmul_rn((float)pc.z, vy), oy));
  dst[C + 5] = __fsub_rn(z, __fadd_rn(__fmul_rn((float)pc.y, vz), oz));
}

// order-preserving float <-> uint key so one atomicMax works for any sign
__device__ __forceinline__ uint32_t f2key(float f) {
  if (f == 0.f) f = 0.f;  // -0.0 and +0.0 compare equal in the reference: one key for both
  const uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float key2f(uint32_t k) {
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

__global__ void __launch_bounds__(TPB) k_fill_u32(uint32_t* p, int64_t n, uint32_t v) {
  for (int64_t i = (int64_t)blockIdx.x * TPB + threadIdx.x; i < n; i += (int64_t)gridDim.x * TPB) p[i] = v;
}

__global__ void __launch_bounds__(TPB) k_max_keys(const float* __restrict__ feat, int64_t total, int C,
                                                  const int32_t* __restrict__ point_pillar, uint32_t* keys) {
  for (int64_t i = (int64_t)blockIdx.x * TPB + threadIdx.x; i < total; i += (int64_t)gridDim.x * TPB) {
    const int64_t p = i / C;
    const int c = (int)(i - p * C);
    const int pid = __ldg(point_pillar + p);
    if (pid < 0) continue;
    const uint32_t k = f2key(feat[i]);
    uint32_t* dst = keys + (int64_t)pid * C + c;
    if (*(volatile uint32_t*)dst < k) atomicMax(dst, k);
  }
}

__global__ void __launch_bounds__(TPB) k_max_arg(const float* __restrict__ feat, int64_t total, int C,
                                                 const int32_t* __restrict__ point_pillar,
                                                 const uint32_t* __restrict__ keys, int32_t* arg) {
  for (int64_t i = (int64_t)blockIdx.x * TPB + threadIdx.x; i < total; i += (int64_t)gridDim.x * TPB) {
    const int64_t p = i / C;
    const int c = (int)(i - p * C);
    const int pid = __ldg(point_pillar + p);
    if (pid < 0) continue;
    if (f2key(feat[i]) == keys[(int64_t)pid * C + c]) atomicMin(arg + (int64_t)pid * C + c, (int32_t)p);
  }
}

__global__ void __launch_bounds__(TPB) k_keys_to_float(uint32_t* keys, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * TPB + threadIdx.x; i < n; i += (int64_t)gridDim.x * TPB)
    reinterpret_cast<float*>(keys)[i] = key2f(keys[i]);
}

__global__ void __launch_bounds__(TPB) k_max_bwd(const float* __restrict__ d_out, int64_t total, int C,
                                                 const int32_t* __restrict__ point_pillar,
                                                 const int32_t* __restrict__ arg, float* d_feat) {
  for (int64_t i = (int64_t)blockIdx.x * TPB + threadIdx.x; i < total; i += (int64_t)gridDim.x * TPB) {
    const int64_t p = i / C;
    const int c = (int)(i - p * C);
    const int pid = __ldg(point_pillar + p);
    float g = 0.f;
    if (pid >= 0 && arg[(int64_t)pid * C + c] == (int32_t)p) g = d_out[(int64_t)pid * C + c];
    d_feat[i] = g;
  }
}

__global__ void __launch_bounds__(TPB) k_sum_fwd(const float* __restrict__ feat, int64_t total, int C,
                                                 const int32_t* __restrict__ point_pillar, float* out) {
  for (int64_t i = (int64_t)blockIdx.x * TPB + threadIdx.x; i < total; i += (int64_t)gridDim.x * TPB) {
    const int64_t p = i / C;
    const int pid = __ldg(point_pillar + p);
    if (pid >= 0) atomicAdd(out + (int64_t)pid * C + (i - p * C), feat[i]);
  }
}

__global__ void __launch_bounds__(TPB) k_mean_div(float* out, int64_t total, int C,
                                                  const float* __restrict__ pillar_mean) {
  for (int64_t i = (int64_t)blockIdx.x * TPB + threadIdx.x; i < total; i += (int64_t)gridDim.x * TPB)
    out[i] = __fdiv_rn(out[i], __ldg(pillar_mean + (i / C) * 4 + 3));
}

__global__ void __launch_bounds__(TPB) k_mean_bwd(const float* __restrict__ d_out, int64_t total, int C,
                                                  const int32_t* __restrict__ point_pillar,
                                                  const float* __restrict__ pillar_mean, int mean, float* d_feat) {
  for (int64_t i = (int64_t)blockIdx.x * TPB + threadIdx.x; i < total; i += (int64_t)gridDim.x * TPB) {
    const int64_t p = i / C;
    const int pid = __ldg(point_pillar + p);
    float g = 0.f;
    if (pid >= 0) {
      g = d_out[(int64_t)pid * C + (i - p * C)];
      if (mean) g = __fdiv_rn(g, __ldg(pillar_mean + (int64_t)pid * 4 + 3));
    }
    d_feat[i] = g;
  }
}

inline int grid_for(int64_t total) {
  const int64_t b = (total + TPB - 1) / TPB;
  return (int)(b < GM_NUM_SMS * 16 ? b : GM_NUM_SMS * 16);
}

}  // namespace

extern "C" int geomae_vfe_decorate(const float* points, int64_t n, int32_t channels, const int32_t* point_pillar,
                                   const float* pillar_mean, const int32_t* pillar_coors, const float voxel_xyz[3],
                                   const float centre_offset_xyz[3], float* out, void* stream) {
  GM_REQUIRE(channels >= 3 && channels <= 8, "vfe_decorate: channels %d not in 3..8", channels);
  if (n == 0) return GEOMAE_OK;
  GM_REQUIRE(points && point_pillar && pillar_mean && pillar_coors && voxel_xyz && centre_offset_xyz && out,
             "vfe_decorate: null argument");
  k_decorate<<<gm_div_up(n, TPB), TPB, 0, (cudaStream_t)stream>>>(
      points, n, channels, point_pillar, pillar_mean, pillar_coors, voxel_xyz[0], voxel_xyz[1], voxel_xyz[2],
      centre_offset_xyz[0], centre_offset_xyz[1], centre_offset_xyz[2], out);
  GM_LAUNCH_CHECK();
  return GEOMAE_OK;
}

extern "C" int geomae_scatter_reduce_fwd(const float* feat, int64_t n_points, int32_t channels,
                                         const int32_t* point_pillar, const float* pillar_mean, int64_t n_pillars,
                                         int32_t mode, float* out, int32_t* arg, void* stream_) {
  GM_REQUIRE(mode >= 0 && mode <= 2, "scatter_reduce: mode %d (0 sum, 1 mean, 2 max)", mode);
  GM_REQUIRE(channels >= 1, "scatter_reduce: channels");
  cudaStream_t stream = (cudaStream_t)stream_;
  const int64_t total = n_points * channels, vtotal = n_pillars * channels;
  if (vtotal == 0) return GEOMAE_OK;
  GM_REQUIRE(out && point_pillar && (feat || n_points == 0), "scatter_reduce: null argument");
  if (mode == 2) {
    GM_REQUIRE(arg, "scatter_reduce(max): arg buffer required");
    k_fill_u32<<<grid_for(vtotal), TPB, 0, stream>>>(reinterpret_cast<uint32_t*>(out), vtotal, 0u);
    k_fill_u32<<<grid_for(vtotal), TPB, 0, stream>>>(reinterpret_cast<uint32_t*>(arg), vtotal, 0x7fffffffu);
    if (total > 0) {
      k_max_keys<<<grid_for(total), TPB, 0, stream>>>(feat, total, channels, point_pillar,
                                                      reinterpret_cast<uint32_t*>(out));
      k_max_arg<<<grid_for(total), TPB, 0, stream>>>(feat, total, channels, point_pillar,
                                                     reinterpret_cast<const uint32_t*>(out), arg);
    }
    k_keys_to_float<<<grid_for(vtotal), TPB, 0, stream>>>(reinterpret_cast<uint32_t*>(out), vtotal);
  } else {
    GM_REQUIRE(mode == 0 || pillar_mean, "scatter_reduce(mean): pillar_mean (counts) required");
    GM_CUDA(cudaMemsetAsync(out, 0, (size_t)vtotal * 4, stream));
    if (total > 0) k_sum_fwd<<<grid_for(total), TPB, 0, stream>>>(feat, total, channels, point_pillar, out);
    if (mode == 1) k_mean_div<<<grid_for(vtotal), TPB, 0, stream>>>(out, vtotal, channels, pillar_mean);
  }
  GM_LAUNCH_CHECK();
  return GEOMAE_OK;
}

extern "C" int geomae_scatter_reduce_bwd(const float* d_out, int64_t n_points, int32_t channels,
                                         const int32_t* point_pillar, const float* pillar_mean, const int32_t* arg,
                                         int32_t mode, float* d_feat, void* stream_) {
  GM_REQUIRE(mode >= 0 && mode <= 2, "scatter_reduce_bwd: mode %d", mode);
  cudaStream_t stream = (cudaStream_t)stream_;
  const int64_t total = n_points * channels;
  if (total == 0) return GEOMAE_OK;
  GM_REQUIRE(d_out && point_pillar && d_feat, "scatter_reduce_bwd: null argument");
  if (mode == 2) {
    GM_REQUIRE(arg, "scatter_reduce_bwd(max): arg required");
    k_max_bwd<<<grid_for(total), TPB, 0, stream>>>(d_out, total, channels, point_pillar, arg, d_feat);
  } else {
    GM_REQUIRE(mode == 0 || pillar_mean, "scatter_reduce_bwd(mean): pillar_mean required");
    k_mean_bwd<<<grid_for(total), TPB, 0, stream>>>(d_out, total, channels, point_pillar, pillar_mean, mode == 1,
                                                    d_feat);
  }
  GM_LAUNCH_CHECK();
  return GEOMAE_OK;
}
