// Fused GeoMAE pre-training losses (reference forward_loss …_ssl.py:837-902,
// mse_loss / cls_sub_voxel branch) evaluated straight from the CSR sub-voxel representation:
// no dense [M,128,3] targets, no boolean-mask gathers (each of which is a host sync in the
// reference), one warp per masked pillar.  Forward returns the six weighted losses; backward
// returns the gradients of all six prediction tensors in one pass.
#include "common.cuh"
#include "voxel_geom.cuh"

namespace {

struct LossIO {
  const int64_t* rows; int64_t m;                    // masked pillar rows
  const float* reg_low; const float* reg_med; const float* reg_top; const float* nor_top;
  const float* cls_low; const float* cls_med;        // [m,S,3] / [m,3] / [m,S,2]
  const float* normal;                               // [V,3] targets (z,y,x)
  float w_low, w_med, w_top, w_nor, w_cls_low, w_cls_med;
  int ld_low, ld_med, ld_top, ld_nor, ld_cls_low, ld_cls_med;   // row strides (floats) of the predictions / gradients
};

__device__ __forceinline__ float norm_coord(float c, int coor, float vs, float lo) {
  return __fdiv_rn(__fsub_rn(c, __fadd_rn(__fmul_rn((float)coor, vs), lo)), vs);
}
__device__ __forceinline__ float bce_logit(float x, float t) { return fmaxf(x, 0.f) - x * t + log1pf(expf(-fabsf(x))); }
__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

__global__ void k_count_present(const int64_t* __restrict__ rows, int64_t m, const uint32_t* __restrict__ med_mask,
                                const uint32_t* __restrict__ low_mask, int32_t* counts /*[2] low, med*/) {
  int cl = 0, cm = 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t v = rows[i];
    const uint4 w = __ldg(reinterpret_cast<const uint4*>(low_mask) + v);
    cl += __popc(w.x) + __popc(w.y) + __popc(w.z) + __popc(w.w);
    cm += __popc(__ldg(med_mask + v));
  }
  cl = gm_warp_sum_i(cl);
  cm = gm_warp_sum_i(cm);
  if ((threadIdx.x & 31) == 0) {
    if (cl) atomicAdd(counts + 0, cl);
    if (cm) atomicAdd(counts + 1, cm);
  }
}

// BWD = false: accumulate the six loss sums into acc[6] (double).  BWD = true: write the gradients.
template <bool BWD>
__global__ void __launch_bounds__(256) k_loss(VoxGeom g, LossIO io, const int32_t* __restrict__ pillar_coors,
                                              const float* __restrict__ pillar_mean,
                                              const uint32_t* __restrict__ med_mask,
                                              const uint32_t* __restrict__ low_mask,
                                              const int32_t* __restrict__ med_ptr, const int32_t* __restrict__ low_ptr,
                                              const float* __restrict__ med_mean, const float* __restrict__ low_mean,
                                              const int32_t* __restrict__ counts, double* acc,
                                              const float* __restrict__ gscale /*[6] upstream grads*/, float* d_low,
                                              float* d_med, float* d_top, float* d_nor, float* d_cls_low,
                                              float* d_cls_med) {
  __shared__ double sacc[8][6];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float n_low = (float)counts[0], n_med = (float)counts[1], mf = (float)io.m;
  float part[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (int64_t i = (int64_t)blockIdx.x * 8 + warp; i < io.m; i += (int64_t)gridDim.x * 8) {
    const int64_t v = io.rows[i];
    const int4 pc = __ldg(reinterpret_cast<const int4*>(pillar_coors) + v);
#pragma unroll
    for (int s = 1; s <= 2; ++s) {
      const int rz = g.ratio[s][0], ry = g.ratio[s][1], rx = g.ratio[s][2];
      const int slots = rz * ry * rx;
      uint4 mask;
      int base;
      const float* mean;
      const float *reg, *cls;
      float *dreg, *dcls;
      float wreg, wcls, cnt;
      int ldr, ldc;
      if (s == 1) {
        mask = make_uint4(__ldg(med_mask + v), 0u, 0u, 0u); base = __ldg(med_ptr + v); mean = med_mean;
        reg = io.reg_med; cls = io.cls_med; dreg = d_med; dcls = d_cls_med; wreg = io.w_med; wcls = io.w_cls_med; cnt = n_med;
        ldr = io.ld_med; ldc = io.ld_cls_med;
      } else {
        mask = __ldg(reinterpret_cast<const uint4*>(low_mask) + v); base = __ldg(low_ptr + v); mean = low_mean;
        reg = io.reg_low; cls = io.cls_low; dreg = d_low; dcls = d_cls_low; wreg = io.w_low; wcls = io.w_cls_low; cnt = n_low;
        ldr = io.ld_low; ldc = io.ld_cls_low;
      }
      const uint32_t w[4] = {mask.x, mask.y, mask.z, mask.w};
      const float greg = BWD ? gscale[s == 1 ? 2 : 1] * wreg * 2.0f / (3.0f * cnt) : 0.f;
      const float gcls = BWD ? gscale[s == 1 ? 5 : 4] * wcls / (mf * slots * 2.0f) : 0.f;
      for (int slot = lane; slot < slots; slot += 32) {
        const bool present = (w[slot >> 5] >> (slot & 31)) & 1u;
        const int64_t er = i * ldr + slot * 3, ec = i * ldc + slot * 2;
        const float2 lg = __ldg(reinterpret_cast<const float2*>(cls + ec));
        const float t0 = present ? 0.f : 1.f, t1 = present ? 1.f : 0.f;     // one-hot of the occupancy label
        if (BWD) {
          *reinterpret_cast<float2*>(dcls + ec) = make_float2((sigmoidf_(lg.x) - t0) * gcls, (sigmoidf_(lg.y) - t1) * gcls);
        } else {
          part[s == 1 ? 5 : 4] += bce_logit(lg.x, t0) + bce_logit(lg.y, t1);
        }
        float dz = 0.f, dy = 0.f, dx = 0.f;
        if (present) {
          const float4 c = __ldg(reinterpret_cast<const float4*>(mean) + base + rank128(mask, slot));
          const int cz = slot / (ry * rx), cy = pc.z * ry + (slot / rx) % ry, cx = pc.w * rx + slot % rx;
          dz = reg[er + 0] - norm_coord(c.z, cz, g.vs[s][2], g.lo[2]);
          dy = reg[er + 1] - norm_coord(c.y, cy, g.vs[s][1], g.lo[1]);
          dx = reg[er + 2] - norm_coord(c.x, cx, g.vs[s][0], g.lo[0]);
          if (!BWD) part[s == 1 ? 2 : 1] += (dz * dz + dy * dy + dx * dx) * (1.0f / 3.0f);
        }
        if (BWD) { dreg[er + 0] = dz * greg; dreg[er + 1] = dy * greg; dreg[er + 2] = dx * greg; }
      }
    }
    if (lane < 3) {
      const float4 c = __ldg(reinterpret_cast<const float4*>(pillar_mean) + v);
      const float cen = lane == 0 ? norm_coord(c.z, pc.y, g.vs[0][2], g.lo[2])
                                  : (lane == 1 ? norm_coord(c.y, pc.z, g.vs[0][1], g.lo[1])
                                               : norm_coord(c.x, pc.w, g.vs[0][0], g.lo[0]));
      const float dt = io.reg_top[i * io.ld_top + lane] - cen;
      const float dn = io.nor_top[i * io.ld_nor + lane] - __ldg(io.normal + v * 3 + lane);
      if (BWD) {
        d_top[i * io.ld_top + lane] = dt * gscale[3] * io.w_top * 2.0f / (3.0f * mf);
        d_nor[i * io.ld_nor + lane] = dn * gscale[0] * io.w_nor * 2.0f / (3.0f * mf);
      } else {
        part[3] += dt * dt * (1.0f / 3.0f);
        part[0] += dn * dn * (1.0f / 3.0f);
      }
    }
  }
  if (BWD) return;
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    const float t = gm_warp_sum(part[k]);
    if (lane == 0) sacc[warp][k] = (double)t;
  }
  __syncthreads();
  if (threadIdx.x < 6) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += sacc[w][threadIdx.x];
    atomicAdd(acc + threadIdx.x, t);
  }
}

__global__ void k_loss_finish(const double* __restrict__ acc, const int32_t* __restrict__ counts, LossIO io,
                              int slots_low, int slots_med, float* out) {
  const int k = threadIdx.x;
  if (k >= 6) return;
  const double mf = (double)io.m;
  const double denom[6] = {mf, (double)counts[0], (double)counts[1], mf, mf * slots_low * 2.0, mf * slots_med * 2.0};
  const double w[6] = {io.w_nor, io.w_low, io.w_med, io.w_top, io.w_cls_low, io.w_cls_med};
  out[k] = (float)(acc[k] / denom[k] * w[k]);
}

int fill(LossIO* io, const geomae_loss_args* a, const geomae_voxel_cfg* cfg) {
  const int sl = cfg->ratio_low[0] * cfg->ratio_low[1] * cfg->ratio_low[2];
  const int sm = cfg->ratio_med[0] * cfg->ratio_med[1] * cfg->ratio_med[2];
  const int dflt[6] = {sl * 3, sm * 3, 3, 3, sl * 2, sm * 2};
  int ld[6];
  for (int k = 0; k < 6; ++k) {
    ld[k] = a->ld[k] ? a->ld[k] : dflt[k];
    GM_REQUIRE(ld[k] >= dflt[k], "geom_loss: row stride %d of prediction %d is smaller than its row (%d)", ld[k], k, dflt[k]);
  }
  GM_REQUIRE(ld[4] % 2 == 0 && ld[5] % 2 == 0, "geom_loss: occupancy-logit row strides must be even");
  io->ld_low = ld[0]; io->ld_med = ld[1]; io->ld_top = ld[2]; io->ld_nor = ld[3]; io->ld_cls_low = ld[4]; io->ld_cls_med = ld[5];
  GM_REQUIRE(a->rows && a->reg_low && a->reg_med && a->reg_top && a->nor_top && a->cls_low && a->cls_med && a->normal,
             "geom_loss: null prediction / target pointer");
  io->rows = a->rows; io->m = a->m;
  io->reg_low = a->reg_low; io->reg_med = a->reg_med; io->reg_top = a->reg_top; io->nor_top = a->nor_top;
  io->cls_low = a->cls_low; io->cls_med = a->cls_med; io->normal = a->normal;
  io->w_low = a->w_low; io->w_med = a->w_med; io->w_top = a->w_top; io->w_nor = a->w_nor;
  io->w_cls_low = a->w_cls_low; io->w_cls_med = a->w_cls_med;
  return GEOMAE_OK;
}

}  // namespace

// losses out[6]: curv_around (normal), centroid_low, centroid_med, centroid_top, cls_low, cls_med
extern "C" int geomae_geom_loss_fwd(const geomae_voxel_cfg* cfg, const geomae_scatter_io* sc, const geomae_loss_args* a,
                                    int32_t* counts /*[2] scratch+out*/, double* acc /*[6] scratch*/, float* out,
                                    void* stream_) {
  GM_REQUIRE(cfg && sc && a && counts && acc && out, "geom_loss_fwd: null argument");
  GM_REQUIRE(a->m > 0, "geom_loss_fwd: no masked pillars");
  cudaStream_t stream = (cudaStream_t)stream_;
  VoxGeom g;
  int rc = gm_make_geom(cfg, sc->n_frames, &g);
  if (rc) return rc;
  // k_loss rebuilds every sub-voxel coordinate from its parent pillar (cz = slot / (ry rx): one z cell per pillar) and
  // relies on every sub-voxel's parent cell being the point's own pillar (nested power-of-two scales); geometries
  // outside that (several pillar cells in z, non-nested ratios) would silently normalise against the wrong origin.
  GM_REQUIRE(g.parent_is_top && g.grid[0][2] == 1,
             "geom_loss: needs nested sub-voxel scales and a single pillar cell in z (got grid z = %d)", g.grid[0][2]);
  LossIO io;
  rc = fill(&io, a, cfg);
  if (rc) return rc;
  GM_CUDA(cudaMemsetAsync(counts, 0, 8, stream));
  GM_CUDA(cudaMemsetAsync(acc, 0, 48, stream));
  int blocks = gm_div_up(a->m, 8);
  if (blocks > GM_NUM_SMS * 8) blocks = GM_NUM_SMS * 8;
  k_count_present<<<gm_div_up(a->m, 256) > 592 ? 592 : gm_div_up(a->m, 256), 256, 0, stream>>>(a->rows, a->m, sc->med_mask,
                                                                                              sc->low_mask, counts);
  k_loss<false><<<blocks, 256, 0, stream>>>(g, io, sc->pillar_coors, sc->pillar_mean, sc->med_mask, sc->low_mask,
                                            sc->med_ptr, sc->low_ptr, sc->med_mean, sc->low_mean, counts, acc, nullptr,
                                            nullptr, nullptr, nullptr, nullptr, nullptr, nullptr);
  const int sl = cfg->ratio_low[0] * cfg->ratio_low[1] * cfg->ratio_low[2];
  const int sm = cfg->ratio_med[0] * cfg->ratio_med[1] * cfg->ratio_med[2];
  k_loss_finish<<<1, 32, 0, stream>>>(acc, counts, io, sl, sm, out);
  GM_LAUNCH_CHECK();
  return GEOMAE_OK;
}

extern "C" int geomae_geom_loss_bwd(const geomae_voxel_cfg* cfg, const geomae_scatter_io* sc, const geomae_loss_args* a,
                                    const int32_t* counts, const float* d_losses /*[6]*/, float* d_reg_low,
                                    float* d_reg_med, float* d_reg_top, float* d_nor_top, float* d_cls_low,
                                    float* d_cls_med, void* stream) {
  GM_REQUIRE(cfg && sc && a && counts && d_losses && d_reg_low && d_reg_med && d_reg_top && d_nor_top && d_cls_low &&
                 d_cls_med, "geom_loss_bwd: null argument");
  VoxGeom g;
  int rc = gm_make_geom(cfg, sc->n_frames, &g);
  if (rc) return rc;
  // k_loss rebuilds every sub-voxel coordinate from its parent pillar (cz = slot / (ry rx): one z cell per pillar) and
  // relies on every sub-voxel's parent cell being the point's own pillar (nested power-of-two scales); geometries
  // outside that (several pillar cells in z, non-nested ratios) would silently normalise against the wrong origin.
  GM_REQUIRE(g.parent_is_top && g.grid[0][2] == 1,
             "geom_loss: needs nested sub-voxel scales and a single pillar cell in z (got grid z = %d)", g.grid[0][2]);
  LossIO io;
  rc = fill(&io, a, cfg);
  if (rc) return rc;
  int blocks = gm_div_up(a->m, 8);
  if (blocks > GM_NUM_SMS * 8) blocks = GM_NUM_SMS * 8;
  k_loss<true><<<blocks, 256, 0, (cudaStream_t)stream>>>(g, io, sc->pillar_coors, sc->pillar_mean, sc->med_mask,
                                                         sc->low_mask, sc->med_ptr, sc->low_ptr, sc->med_mean,
                                                         sc->low_mean, counts, nullptr, d_losses, d_reg_low, d_reg_med,
                                                         d_reg_top, d_nor_top, d_cls_low, d_cls_med);
  GM_LAUNCH_CHECK();
  return GEOMAE_OK;
}
