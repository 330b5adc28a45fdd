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

// Cyclic Jacobi on a symmetric 3x3 matrix in fp32 — the arithmetic the reference's own solver uses (torch.svd of an
// fp32 batch runs cuSOLVER's batched Jacobi).  a = {a00, a01, a02, a11, a12, a22}; v's COLUMNS converge to the
// eigenvectors.  Rotations use the stable small-angle root t = sgn(theta) / (|theta| + sqrt(theta^2 + 1)).
__device__ __forceinline__ void jacobi3f(float a[3][3], float v[3][3]) {
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) v[i][j] = (i == j) ? 1.f : 0.f;
  for (int sweep = 0; sweep < 8; ++sweep) {
    const float off = a[0][1] * a[0][1] + a[0][2] * a[0][2] + a[1][2] * a[1][2];
    const float diag = a[0][0] * a[0][0] + a[1][1] * a[1][1] + a[2][2] * a[2][2];
    if (off <= 1e-15f * diag || off == 0.f) break;  // off-diagonal norm below 3e-8 of the diagonal's: fp32 converged
#pragma unroll
    for (int pq = 0; pq < 3; ++pq) {
      const int p = pq == 2 ? 1 : 0, q = pq == 0 ? 1 : 2;
      const float apq = a[p][q];
      if (fabsf(apq) < 1e-30f) {  // also keeps the reciprocal below out of the flushed-denormal range
        a[p][q] = a[q][p] = 0.f;
        continue;
      }
      // the rotation only has to be a rotation (c, s below) by roughly the annihilating angle: reciprocal-unit
      // divides instead of IEEE ones; theta^2 overflowing to inf gives t = 0, the correct limit
      const float theta = __fdividef(a[q][q] - a[p][p], 2.f * apq);
      const float t = copysignf(__fdividef(1.f, fabsf(theta) + sqrtf(fmaf(theta, theta, 1.f))), theta);
      const float c = rsqrtf(fmaf(t, t, 1.f)), s = t * c;
      const int r = 3 - p - q;
      const float arp = a[r][p], arq = a[r][q];
      a[p][p] = fmaf(-t, apq, a[p][p]);
      a[q][q] = fmaf(t, apq, a[q][q]);
      a[p][q] = a[q][p] = 0.f;
      a[r][p] = a[p][r] = c * arp - s * arq;
      a[r][q] = a[q][r] = s * arp + c * arq;
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        const float vip = v[i][p], viq = v[i][q];
        v[i][p] = c * vip - s * viq;
        v[i][q] = s * vip + c * viq;
      }
    }
  }
}

// One thread per pillar.  (1) the 3x3 BEV neighbourhood through the occupancy bitmap: per neighbour row ONE 64-bit
// window of the bitmap + one word rank cover all three cells; (2) slot masks + CSR offsets of the occupied
// neighbours; (3) their middle-scale centroid rows -> six moments in registers, accumulated in (neighbour, slot)
// order; (4) fp32 Jacobi eigenvectors, eigenvalues re-evaluated as fp64 Rayleigh quotients of the fp32 matrix (second
// order in the eigenvector error), curvature in fp64 like the reference's .double() tail (…_ssl.py:604-607).
__global__ void __launch_bounds__(TPB) k_geom(VoxGeom g, const uint32_t* __restrict__ bitmap, int n_words,
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
  const int gx = g.grid[0][0], gy = g.grid[0][1];
  int nid[9];
  {
    uint64_t win[3];
    int wr[3], o0[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const int ny = pc.z + r - 1;
      const bool in = ny >= 0 && ny < gy;
      const int64_t cm = in ? top_cell(g, pc.x, ny, pc.w) : 0;  // centre cell of the row
      const int64_t c0 = cm > 0 ? cm - 1 : 0;
      const int wi = (int)(c0 >> 5);
      o0[r] = (int)(cm - 1 - ((int64_t)wi << 5));               // bit offset of cell (ny, x-1) in the window (may be -1)
      const uint32_t lo = in ? __ldg(bitmap + wi) : 0u;
      const uint32_t hi = (in && wi + 1 < n_words) ? __ldg(bitmap + wi + 1) : 0u;
      wr[r] = in ? __ldg(word_rank + wi) : 0;
      win[r] = ((uint64_t)hi << 32) | lo;
    }
#pragma unroll
    for (int k = 0; k < 9; ++k) {
      const int r = k / 3, nx = pc.w + k % 3 - 1;
      const int o = o0[r] + k % 3;
      const bool in = nx >= 0 && nx < gx && o >= 0;               // row validity is folded into win == 0
      const uint64_t bit = 1ull << (o & 63);
      nid[k] = (in && (win[r] & bit)) ? wr[r] + __popcll(win[r] & (bit - 1)) : -1;
    }
  }
  uint32_t msk[9];
  int ptr[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) {
    if (pair) pair[(int64_t)k * n + v] = nid[k];
    msk[k] = nid[k] >= 0 ? __ldg(med_mask + nid[k]) : 0u;
    ptr[k] = nid[k] >= 0 ? __ldg(med_ptr + nid[k]) : 0;
  }
  float zz = 0.f, zy = 0.f, zx = 0.f, yy = 0.f, yx = 0.f, xx = 0.f;
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
  float a[3][3] = {{zz, zy, zx}, {zy, yy, yx}, {zx, yx, xx}}, ev[3][3];
  jacobi3f(a, ev);
  // eigenvalues: Rayleigh quotients of the ORIGINAL matrix in fp64 (removes the rounding the fp32 rotations
  // accumulate on the diagonal); the matrix is a Gram matrix, so they are the singular values
  double lam[3];
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    const double e0 = ev[0][j], e1 = ev[1][j], e2 = ev[2][j];
    const double q = e0 * ((double)zz * e0 + (double)zy * e1 + (double)zx * e2) +
                     e1 * ((double)zy * e0 + (double)yy * e1 + (double)yx * e2) +
                     e2 * ((double)zx * e0 + (double)yx * e1 + (double)xx * e2);
    const double nn = e0 * e0 + e1 * e1 + e2 * e2;
    lam[j] = fabs(q / nn);
  }
  // stable descending sort of the three (value, vector) pairs (ties keep index order -> zero matrix gives (0,0,1))
  float e[3][3];  // e[j] = eigenvector j as (z, y, x)
#pragma unroll
  for (int j = 0; j < 3; ++j) { e[j][0] = ev[0][j]; e[j][1] = ev[1][j]; e[j][2] = ev[2][j]; }
  auto cswap = [&](int i, int j) {
    if (lam[j] > lam[i]) {
      const double tl = lam[i]; lam[i] = lam[j]; lam[j] = tl;
#pragma unroll
      for (int c = 0; c < 3; ++c) { const float tv = e[i][c]; e[i][c] = e[j][c]; e[j][c] = tv; }
    }
  };
  cswap(0, 1);
  cswap(1, 2);
  cswap(0, 1);
  float nz = e[2][0], ny_ = e[2][1], nx_ = e[2][2];
  const float inv = rsqrtf(nz * nz + ny_ * ny_ + nx_ * nx_);
  nz *= inv; ny_ *= inv; nx_ *= inv;
  {  // one Newton step on the normalisation: rsqrtf is a 2-ulp approximation
    const float n2 = nz * nz + ny_ * ny_ + nx_ * nx_;
    const float fix = fmaf(-0.5f, n2 - 1.f, 1.f);
    nz *= fix; ny_ *= fix; nx_ *= fix;
  }
  // documented sign convention: first non-zero component of (z,y,x) is positive
  const float lead = (fabsf(nz) > 1e-12f) ? nz : ((fabsf(ny_) > 1e-12f) ? ny_ : nx_);
  if (lead < 0.f) { nz = -nz; ny_ = -ny_; nx_ = -nx_; }
  normal[v * 3 + 0] = nz;
  normal[v * 3 + 1] = ny_;
  normal[v * 3 + 2] = nx_;
  const float s0 = (float)lam[0], s1 = (float)lam[1], s2 = (float)lam[2];
  if (singular) { singular[v * 3 + 0] = s0; singular[v * 3 + 1] = s1; singular[v * 3 + 2] = s2; }
  const double e0 = (double)s0 + 1e-9, e1 = (double)s1 + 1e-9, e2 = (double)s2 + 1e-9;
  const double rsum = 1.0 / (e0 + e1 + e2);
  curvature[v * 3 + 0] = e0 * rsum;
  curvature[v * 3 + 1] = e1 * rsum;
  curvature[v * 3 + 2] = e2 * rsum;
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
      g, io->bitmap, gm_div_up((int64_t)io->n_frames * g.grid[0][0] * g.grid[0][1], 32), io->word_rank, io->pillar_coors, io->pillar_mean, io->med_mask, io->med_ptr, io->med_mean,
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
