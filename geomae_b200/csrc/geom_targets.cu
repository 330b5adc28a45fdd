// Per-pillar geometric targets: neighbour lookup, 3x3 scatter matrix,
// symmetric eigen-solve -> unit normal + curvature, and the dense slot targets the loss consumes.
//
// The reference materialises a [9,V,16,3] gather twice, runs a batched bmm and a cuSOLVER SVD.
// Here one thread walks its pillar's 3x3 neighbourhood through the occupancy bitmap (the same
// structure that ranks the pillars — it doubles as spconv's hash table), reads only the occupied
// middle-scale centroids from their CSR rows, keeps the 6 unique moments in registers and solves
// the 3x3 eigenproblem with cyclic Jacobi in fp64.  HBM-bound; no tensor cores by design.
#include "common.cuh"
#include "voxel_geom.cuh"

namespace {

constexpr int TPB = 128;

__device__ __forceinline__ void jacobi3(double a[3][3], double v[3][3]) {
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) v[i][j] = (i == j) ? 1.0 : 0.0;
  for (int sweep = 0; sweep < 16; ++sweep) {
    const double off = a[0][1] * a[0][1] + a[0][2] * a[0][2] + a[1][2] * a[1][2];
    const double diag = a[0][0] * a[0][0] + a[1][1] * a[1][1] + a[2][2] * a[2][2];
    if (off <= 1e-30 * diag || off == 0.0) break;
#pragma unroll
    for (int pq = 0; pq < 3; ++pq) {
      const int p = pq == 2 ? 1 : 0, q = pq == 0 ? 1 : 2;
      const double apq = a[p][q];
      if (apq == 0.0) continue;
      const double theta = (a[q][q] - a[p][p]) / (2.0 * apq);
      const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
      const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
      const int r = 3 - p - q;
      const double app = a[p][p], aqq = a[q][q], arp = a[r][p], arq = a[r][q];
      a[p][p] = app - t * apq;
      a[q][q] = aqq + t * apq;
      a[p][q] = a[q][p] = 0.0;
      a[r][p] = a[p][r] = c * arp - s * arq;
      a[r][q] = a[q][r] = s * arp + c * arq;
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        const double vip = v[i][p], viq = v[i][q];
        v[i][p] = c * vip - s * viq;
        v[i][q] = s * vip + c * viq;
      }
    }
  }
}

__global__ void __launch_bounds__(TPB) k_geom(VoxGeom g, const uint32_t* __restrict__ bitmap,
                                              const int32_t* __restrict__ word_rank,
                                              const int32_t* __restrict__ pillar_coors,
                                              const float* __restrict__ pillar_mean,
                                              const uint32_t* __restrict__ med_mask,
                                              const int32_t* __restrict__ med_ptr, const float* __restrict__ med_mean,
                                              int64_t n, float* normal, double* curvature, float* cov6,
                                              float* singular, int32_t* pair) {
  const int64_t v = (int64_t)blockIdx.x * TPB + threadIdx.x;
  if (v >= n) return;
  const int4 pc = __ldg(reinterpret_cast<const int4*>(pillar_coors) + v);
  const float4 ctr = __ldg(reinterpret_cast<const float4*>(pillar_mean) + v);
  float zz = 0.f, zy = 0.f, zx = 0.f, yy = 0.f, yx = 0.f, xx = 0.f;
  // Three phases so that the loads of all nine neighbours are in flight together (the walk used to be nine serial
  // chains of four dependent loads): (1) occupancy words + ranks, (2) slot masks + CSR offsets, (3) centroid rows.
  // Accumulation order (neighbour k, then slot) is unchanged, so the moments are bit-identical.
  int nid[9];
  uint32_t bw[9];
  int wr[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) {
    const int ny = pc.z + k / 3 - 1, nx = pc.w + k % 3 - 1;
    const bool in = ny >= 0 && ny < g.grid[0][1] && nx >= 0 && nx < g.grid[0][0];
    const int64_t cell = in ? top_cell(g, pc.x, ny, nx) : 0;
    bw[k] = in ? __ldg(bitmap + (cell >> 5)) : 0u;
    wr[k] = in ? __ldg(word_rank + (cell >> 5)) : 0;
    const uint32_t bit = 1u << (cell & 31);
    nid[k] = (in && (bw[k] & bit)) ? wr[k] + __popc(bw[k] & (bit - 1)) : -1;
  }
  uint32_t msk[9];
  int ptr[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) {
    if (pair) pair[(int64_t)k * n + v] = nid[k];
    msk[k] = nid[k] >= 0 ? __ldg(med_mask + nid[k]) : 0u;
    ptr[k] = nid[k] >= 0 ? __ldg(med_ptr + nid[k]) : 0;
  }
#pragma unroll
  for (int k = 0; k < 9; ++k) {
    const float4* row = reinterpret_cast<const float4*>(med_mean) + ptr[k];
    for (uint32_t m = msk[k]; m; m &= m - 1, ++row) {
      const float4 c = __ldg(row);
      const float dz = __fsub_rn(c.z, ctr.z), dyy = __fsub_rn(c.y, ctr.y), dxx = __fsub_rn(c.x, ctr.x);
      zz = fmaf(dz, dz, zz); zy = fmaf(dz, dyy, zy); zx = fmaf(dz, dxx, zx);
      yy = fmaf(dyy, dyy, yy); yx = fmaf(dyy, dxx, yx); xx = fmaf(dxx, dxx, xx);
    }
  }
  if (cov6) {
    float* o = cov6 + v * 6;
    o[0] = zz; o[1] = zy; o[2] = zx; o[3] = yy; o[4] = yx; o[5] = xx;
  }
  double a[3][3] = {{zz, zy, zx}, {zy, yy, yx}, {zx, yx, xx}}, ev[3][3];
  jacobi3(a, ev);
  double lam[3] = {fabs(a[0][0]), fabs(a[1][1]), fabs(a[2][2])};
  int idx[3] = {0, 1, 2};
  // stable descending sort of three values (ties keep index order -> zero matrix gives (0,0,1))
  if (lam[idx[1]] > lam[idx[0]]) { int t = idx[0]; idx[0] = idx[1]; idx[1] = t; }
  if (lam[idx[2]] > lam[idx[1]]) { int t = idx[1]; idx[1] = idx[2]; idx[2] = t; }
  if (lam[idx[1]] > lam[idx[0]]) { int t = idx[0]; idx[0] = idx[1]; idx[1] = t; }
  double nz = ev[0][idx[2]], ny_ = ev[1][idx[2]], nx_ = ev[2][idx[2]];
  const double inv = 1.0 / sqrt(nz * nz + ny_ * ny_ + nx_ * nx_);
  nz *= inv; ny_ *= inv; nx_ *= inv;
  // documented sign convention: first non-zero component of (z,y,x) is positive
  const double lead = (fabs(nz) > 1e-12) ? nz : ((fabs(ny_) > 1e-12) ? ny_ : nx_);
  if (lead < 0) { nz = -nz; ny_ = -ny_; nx_ = -nx_; }
  normal[v * 3 + 0] = (float)nz;
  normal[v * 3 + 1] = (float)ny_;
  normal[v * 3 + 2] = (float)nx_;
  const float s0 = (float)lam[idx[0]], s1 = (float)lam[idx[1]], s2 = (float)lam[idx[2]];
  if (singular) { singular[v * 3 + 0] = s0; singular[v * 3 + 1] = s1; singular[v * 3 + 2] = s2; }
  const double e0 = (double)s0 + 1e-9, e1 = (double)s1 + 1e-9, e2 = (double)s2 + 1e-9;
  const double sum = e0 + e1 + e2;
  curvature[v * 3 + 0] = e0 / sum;
  curvature[v * 3 + 1] = e1 / sum;
  curvature[v * 3 + 2] = e2 / sum;
}

// (c - (coor*size + min)) / size, each step rounded like the reference's separate torch ops
__device__ __forceinline__ float norm_coord(float c, int coor, float vs, float lo) {
  return __fdiv_rn(__fsub_rn(c, __fadd_rn(__fmul_rn((float)coor, vs), lo)), vs);
}

// one warp per selected pillar
__global__ void __launch_bounds__(256) k_dense_targets(VoxGeom g, const int32_t* __restrict__ pillar_coors,
                                                       const float* __restrict__ pillar_mean,
                                                       const uint32_t* __restrict__ med_mask,
                                                       const uint32_t* __restrict__ low_mask,
                                                       const int32_t* __restrict__ med_ptr,
                                                       const int32_t* __restrict__ low_ptr,
                                                       const float* __restrict__ med_mean,
                                                       const float* __restrict__ low_mean,
                                                       const int64_t* __restrict__ rows, int64_t m, int raw, float* low,
                                                       uint8_t* low_m, float* med, uint8_t* med_m, float* top) {
  const int lane = threadIdx.x & 31;
  const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (i >= m) return;
  const int64_t v = rows[i];
  const int4 pc = __ldg(reinterpret_cast<const int4*>(pillar_coors) + v);
  for (int s = 1; s <= 2; ++s) {
    float* out = s == 1 ? med : low;
    uint8_t* outm = s == 1 ? med_m : low_m;
    if (!out && !outm) continue;
    const int rz = g.ratio[s][0], ry = g.ratio[s][1], rx = g.ratio[s][2];
    const int slots = rz * ry * rx;
    uint4 mask;
    int base;
    const float* mean;
    if (s == 1) {
      mask = make_uint4(__ldg(med_mask + v), 0u, 0u, 0u);
      base = __ldg(med_ptr + v);
      mean = med_mean;
    } else {
      mask = __ldg(reinterpret_cast<const uint4*>(low_mask) + v);
      base = __ldg(low_ptr + v);
      mean = low_mean;
    }
    const uint32_t w[4] = {mask.x, mask.y, mask.z, mask.w};
    for (int slot = lane; slot < slots; slot += 32) {
      const bool present = (w[slot >> 5] >> (slot & 31)) & 1u;
      float oz = 0.f, oy = 0.f, ox = 0.f;
      if (present) {
        const float4 c = __ldg(reinterpret_cast<const float4*>(mean) + base + rank128(mask, slot));
        if (raw) {
          oz = c.z; oy = c.y; ox = c.x;
        } else {
          const int cz = slot / (ry * rx), cy = pc.z * ry + (slot / rx) % ry, cx = pc.w * rx + slot % rx;
          oz = norm_coord(c.z, cz, g.vs[s][2], g.lo[2]);
          oy = norm_coord(c.y, cy, g.vs[s][1], g.lo[1]);
          ox = norm_coord(c.x, cx, g.vs[s][0], g.lo[0]);
        }
      }
      if (out) {
        float* o = out + (i * slots + slot) * 3;
        o[0] = oz; o[1] = oy; o[2] = ox;
      }
      if (outm) outm[i * slots + slot] = present ? 1 : 0;
    }
  }
  if (top && lane == 0) {
    const float4 c = __ldg(reinterpret_cast<const float4*>(pillar_mean) + v);
    if (raw) {
      top[i * 3 + 0] = c.z; top[i * 3 + 1] = c.y; top[i * 3 + 2] = c.x;
    } else {
      top[i * 3 + 0] = norm_coord(c.z, pc.y, g.vs[0][2], g.lo[2]);
      top[i * 3 + 1] = norm_coord(c.y, pc.z, g.vs[0][1], g.lo[1]);
      top[i * 3 + 2] = norm_coord(c.x, pc.w, g.vs[0][0], g.lo[0]);
    }
  }
}

}  // namespace

extern "C" int geomae_geom_targets(const geomae_voxel_cfg* cfg, const geomae_scatter_io* io, int64_t n_pillars,
                                   float* normal, double* curvature, float* cov6, float* singular, int32_t* pair,
                                   void* stream) {
  GM_REQUIRE(cfg && io && normal && curvature, "geom_targets: null argument");
  GM_REQUIRE(n_pillars >= 0 && n_pillars <= io->cap, "geom_targets: n_pillars %lld out of range", (long long)n_pillars);
  if (n_pillars == 0) return GEOMAE_OK;
  VoxGeom g;
  int rc = gm_make_geom(cfg, io->n_frames, &g);
  if (rc) return rc;
  k_geom<<<gm_div_up(n_pillars, TPB), TPB, 0, (cudaStream_t)stream>>>(
      g, io->bitmap, io->word_rank, io->pillar_coors, io->pillar_mean, io->med_mask, io->med_ptr, io->med_mean,
      n_pillars, normal, curvature, cov6, singular, pair);
  GM_LAUNCH_CHECK();
  return GEOMAE_OK;
}

extern "C" int geomae_dense_targets(const geomae_voxel_cfg* cfg, const geomae_scatter_io* io, const int64_t* rows,
                                    int64_t m, int32_t raw, float* low, uint8_t* low_mask, float* med,
                                    uint8_t* med_mask, float* top, void* stream) {
  GM_REQUIRE(cfg && io && (rows || m == 0), "dense_targets: null argument");
  if (m == 0) return GEOMAE_OK;
  VoxGeom g;
  int rc = gm_make_geom(cfg, io->n_frames, &g);
  if (rc) return rc;
  k_dense_targets<<<gm_div_up(m * 32, 256), 256, 0, (cudaStream_t)stream>>>(
      g, io->pillar_coors, io->pillar_mean, io->med_mask, io->low_mask, io->med_ptr, io->low_ptr, io->med_mean,
      io->low_mean, rows, m, raw, low, low_mask, med, med_mask, top);
  GM_LAUNCH_CHECK();
  return GEOMAE_OK;
}
