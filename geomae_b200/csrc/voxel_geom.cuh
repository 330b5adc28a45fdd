// Voxel geometry shared by the scatter and geometric-target kernels.
#pragma once
#include "common.cuh"

namespace {

struct VoxGeom {
  float lo[3];        // range min x,y,z
  float vs[3][3];     // [scale: 0 top, 1 med, 2 low][x,y,z]
  int grid[3][3];     // [scale][x,y,z]
  int ratio[3][3];    // [scale][z,y,x]
  int n_frames;
  // shift[s][a] >= 0: on axis a the voxel of scale s is exactly 2^shift low-scale voxels (sizes and grids), so its
  // coordinate is the low-scale coordinate >> shift — bit-exact with the independent IEEE divide, because dividing by
  // v * 2^k only changes the exponent of the quotient (every GeoMAE config qualifies).  -1: divide.
  int shift[3][3];
  int parent_is_top;  // the parent BEV cell of every sub-voxel is the point's own pillar cell (x, y shifts consistent)
  // fast path (every GeoMAE config): all scales are power-of-two multiples of the low scale and all sub-voxel ratios
  // are powers of two.  Then one low-scale coordinate per axis yields everything by shifts and masks, and that
  // coordinate comes from a reciprocal multiply that is PROVEN equal to floor of the IEEE quotient whenever the
  // product's fraction is further than qeps from an integer (the product is within |q| * 2^-22 < qeps / 2 of the
  // real quotient, and rounding the quotient to fp32 cannot cross an integer it is that far from); otherwise the
  // point takes the IEEE divide.
  int fast;
  float rvs[3];       // 1 / low voxel size, x,y,z
  float qeps[3];
  int smask[3][3];    // [scale 1,2][x,y,z] ratio - 1
  int sshift[3][3];   // [scale 1,2][x,y,z] left shift of the axis inside the slot id
};

__device__ __forceinline__ int vox_coord(float p, float lo, float vs, int g) {
  // fp32 subtract, IEEE divide (no fast-math, no reciprocal), floor, clamp — bit-exact with
  // voxelization_cpu.cpp:22-31 / voxelization_cuda.cu:35-57.
  int c = (int)floorf(__fdiv_rn(__fsub_rn(p, lo), vs));
  return min(max(c, 0), g - 1);
}

struct PointKeys {
  int b;
  int c[3][3];  // [scale][x,y,z]
};

__device__ __forceinline__ void point_keys(const VoxGeom& g, const float* p, PointKeys& k) {
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const int lowc = vox_coord(p[a], g.lo[a], g.vs[2][a], g.grid[2][a]);
    k.c[2][a] = lowc;
#pragma unroll
    for (int s = 0; s < 2; ++s)
      k.c[s][a] = g.shift[s][a] >= 0 ? (lowc >> g.shift[s][a]) : vox_coord(p[a], g.lo[a], g.vs[s][a], g.grid[s][a]);
  }
}

// low-scale coordinate on the fast path (see VoxGeom::fast): branch-free candidate + "needs the IEEE divide" flag
__device__ __forceinline__ int vox_coord_try(float p, float lo, float rvs, float qeps, int g, bool& redo) {
  const float q = __fmul_rn(__fsub_rn(p, lo), rvs);
  const float f = floorf(q);
  const float fr = q - f;
  redo = !(fr > qeps && fr < 1.0f - qeps);
  return min(max((int)f, 0), g - 1);
}

__device__ __forceinline__ int frame_of(const int32_t* __restrict__ off, int n_frames, int64_t idx) {
  int lo = 0, hi = n_frames;  // find largest b with off[b] <= idx
  while (hi - lo > 1) {
    int mid = (lo + hi) >> 1;
    if ((int64_t)__ldg(off + mid) <= idx) lo = mid; else hi = mid;
  }
  return lo;
}

__device__ __forceinline__ int64_t top_cell(const VoxGeom& g, int b, int y, int x) {
  return ((int64_t)b * g.grid[0][1] + y) * g.grid[0][0] + x;
}

// parent BEV cell + slot of a sub-voxel (…_ssl.py:659-665): parent = (y//ry, x//rx), pillar z ignored.
__device__ __forceinline__ void sub_parent(const VoxGeom& g, int s, const PointKeys& k, int64_t& cell, int& slot) {
  const int rz = g.ratio[s][0], ry = g.ratio[s][1], rx = g.ratio[s][2];
  const int cx = k.c[s][0], cy = k.c[s][1], cz = k.c[s][2];
  cell = top_cell(g, k.b, min(cy / ry, g.grid[0][1] - 1), min(cx / rx, g.grid[0][0] - 1));
  slot = (cz % rz) * (ry * rx) + (cy % ry) * rx + (cx % rx);
}

__device__ __forceinline__ int cell_rank(const uint32_t* __restrict__ bitmap, const int32_t* __restrict__ word_rank,
                                         int64_t cell) {
  const uint32_t w = __ldg(bitmap + (cell >> 5));
  const uint32_t bit = 1u << (cell & 31);
  if (!(w & bit)) return -1;
  return __ldg(word_rank + (cell >> 5)) + __popc(w & (bit - 1));
}

__device__ __forceinline__ int rank128(const uint4 m, int slot) {
  const uint32_t w[4] = {m.x, m.y, m.z, m.w};
  int r = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    if (i < (slot >> 5)) r += __popc(w[i]);
    else if (i == (slot >> 5)) r += __popc(w[i] & ((1u << (slot & 31)) - 1u));
  }
  return r;
}

// fast path: the middle-scale slot that contains low-scale slot s (bit fields, see gm_make_geom)
__device__ __forceinline__ int med_slot_of_low(const VoxGeom& g, int s) {
  const int lx = s & g.smask[2][0], ly = (s >> g.sshift[2][1]) & g.smask[2][1], lz = (s >> g.sshift[2][2]) & g.smask[2][2];
  return ((lz >> g.shift[1][2]) << g.sshift[1][2]) | ((ly >> g.shift[1][1]) << g.sshift[1][1]) | (lx >> g.shift[1][0]);
}

__device__ __forceinline__ uint32_t med_mask_of_low(const VoxGeom& g, const uint4 m) {
  const uint32_t w[4] = {m.x, m.y, m.z, m.w};
  uint32_t out = 0u;
#pragma unroll
  for (int i = 0; i < 4; ++i)
    for (uint32_t b = w[i]; b; b &= b - 1) out |= 1u << med_slot_of_low(g, i * 32 + __ffs(b) - 1);
  return out;
}

inline int gm_make_geom(const geomae_voxel_cfg* cfg, int n_frames, VoxGeom* g) {
  const float* sizes[3] = {cfg->voxel_top, cfg->voxel_med, cfg->voxel_low};
  for (int a = 0; a < 3; ++a) g->lo[a] = cfg->range_min[a];
  for (int s = 0; s < 3; ++s) {
    int32_t grid[3];
    int rc = geomae_grid_size(cfg->range_min, cfg->range_max, sizes[s], grid);
    if (rc) return rc;
    for (int a = 0; a < 3; ++a) {
      g->vs[s][a] = sizes[s][a];
      g->grid[s][a] = grid[a];
    }
  }
  for (int a = 0; a < 3; ++a) {
    g->ratio[0][a] = 1;
    g->ratio[1][a] = cfg->ratio_med[a];
    g->ratio[2][a] = cfg->ratio_low[a];
  }
  g->n_frames = n_frames;
  for (int s = 0; s < 3; ++s)
    for (int a = 0; a < 3; ++a) {
      g->shift[s][a] = -1;
      for (int k = 0; k <= 8; ++k)      // exact fp32 comparisons: v_s == v_low * 2^k and grid_low == grid_s * 2^k
        if (g->vs[s][a] == g->vs[2][a] * (float)(1 << k) && g->grid[2][a] == g->grid[s][a] * (1 << k)) {
          g->shift[s][a] = k;
          break;
        }
    }
  // parent cell (cy / ry, cx / rx) of a sub-voxel == the pillar cell when both are shifts of the same low coordinate
  g->parent_is_top = 1;
  for (int s = 1; s < 3; ++s)
    for (int a = 0; a < 2; ++a) {       // x, y
      const int r = g->ratio[s][2 - a]; // ratio is stored (z, y, x)
      if (g->shift[0][a] < 0 || g->shift[s][a] < 0 || (1 << (g->shift[0][a] - g->shift[s][a])) != r) g->parent_is_top = 0;
    }
  g->fast = g->parent_is_top;
  for (int s = 0; s < 2; ++s)
    for (int a = 0; a < 3; ++a)
      if (g->shift[s][a] < 0) g->fast = 0;
  for (int s = 1; s < 3; ++s) {
    int sh = 0;
    for (int a = 0; a < 3; ++a) {         // x, y, z; ratio is stored (z, y, x)
      const int r = g->ratio[s][2 - a];
      if (r < 1 || (r & (r - 1))) g->fast = 0;
      // nested scales on every axis: the pillar is exactly ratio sub-voxels wide (so a sub-voxel's in-pillar
      // coordinate is a bit field of the low-scale coordinate, and a low sub-voxel lies inside one middle sub-voxel)
      if (g->fast && (1 << (g->shift[0][a] - g->shift[s][a])) != r) g->fast = 0;
      g->smask[s][a] = r - 1;
      g->sshift[s][a] = sh;
      while ((1 << sh) < r * (1 << g->sshift[s][a])) ++sh;
    }
  }
  if (g->fast && g->shift[1][0] + g->shift[1][1] + g->shift[1][2] > 5) g->fast = 0;  // <= 32 low per middle sub-voxel
  for (int a = 0; a < 3; ++a) {
    g->smask[0][a] = g->sshift[0][a] = 0;
    g->rvs[a] = 1.0f / g->vs[2][a];
    // |fl(d * rvs) - d / vs| <= |q| * (2^-24 + 2^-24 + 2^-48) < |q| * 2^-22; in-range |q| <= grid
    g->qeps[a] = (float)(g->grid[2][a] + 2) * 4.8e-7f + 1e-6f;
    if (!(g->qeps[a] < 0.25f)) g->fast = 0;
  }
  const int slots_med = cfg->ratio_med[0] * cfg->ratio_med[1] * cfg->ratio_med[2];
  const int slots_low = cfg->ratio_low[0] * cfg->ratio_low[1] * cfg->ratio_low[2];
  GM_REQUIRE(slots_med >= 1 && slots_med <= 32, "sub_voxel_ratio_med has %d slots, supported 1..32", slots_med);
  GM_REQUIRE(slots_low >= 1 && slots_low <= 128, "sub_voxel_ratio_low has %d slots, supported 1..128", slots_low);
  GM_REQUIRE(g->grid[0][2] == 1, "pillar grid must be one voxel deep (z grid = %d)", g->grid[0][2]);
  return GEOMAE_OK;
}


}  // namespace
