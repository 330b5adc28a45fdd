// Fused three-scale dynamic voxelisation + point->voxel scatter (SURVEY.md §8 rows a1-a3,a6,a7,a11).
//
// Design (B200-first, not a port of voxelization_cuda.cu / torch.unique / torch_scatter):
//   * the raw [P,5] fp32 records are streamed with 128-bit loads into shared memory and
//     de-interleaved there (a 20-byte record cannot be loaded coalesced per thread);
//   * the lexicographically sorted pillar list the reference gets from torch.unique(dim=0)
//     (6 row sorts per step) is produced WITHOUT a sort: one occupancy bit per BEV cell,
//     popcount prefix over the bitmap words, pillar row = rank of the cell's bit;
//   * sub-voxels never get a global id: each pillar owns a slot bit-mask (16 / 128 bits) and a
//     CSR range; a sub-voxel's row is ptr[pillar] + popcount(mask below its slot);
//   * sums use vector float4 reductions in L2 (red.global.add.v4.f32), one per point and scale.
// HBM-bound integer/byte work: no tensor cores here by design.
#include "common.cuh"
#include "voxel_geom.cuh"

namespace {

constexpr int TPB = 256;
constexpr int MAX_STRIDE = 8;

// Stage TPB consecutive point records through shared memory with 128-bit loads.
__device__ __forceinline__ bool load_tile(const float* __restrict__ pts, int64_t n, int stride, float* tile,
                                          int64_t& idx, float* p) {
  const int64_t p0 = (int64_t)blockIdx.x * TPB;
  const int nvalid = (int)min((int64_t)TPB, n - p0);
  const int nfloat = nvalid * stride;
  const float* src = pts + p0 * stride;
  if ((reinterpret_cast<uintptr_t>(src) & 15u) == 0) {
    const int nvec = nfloat >> 2;
    const float4* src4 = reinterpret_cast<const float4*>(src);
    float4* dst4 = reinterpret_cast<float4*>(tile);
    for (int i = threadIdx.x; i < nvec; i += TPB) dst4[i] = __ldg(src4 + i);
    for (int i = (nvec << 2) + threadIdx.x; i < nfloat; i += TPB) tile[i] = __ldg(src + i);
  } else {
    for (int i = threadIdx.x; i < nfloat; i += TPB) tile[i] = __ldg(src + i);
  }
  __syncthreads();
  idx = p0 + threadIdx.x;
  if ((int)threadIdx.x >= nvalid) return false;
  p[0] = tile[threadIdx.x * stride + 0];
  p[1] = tile[threadIdx.x * stride + 1];
  p[2] = tile[threadIdx.x * stride + 2];
  return true;
}

__device__ __forceinline__ void red_add4(float* addr, float x, float y, float z, float w) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(x), "f"(y), "f"(z), "f"(w)
               : "memory");
}

// ---------------------------------------------------------------- pass 1: mark occupancy
__global__ void __launch_bounds__(TPB) k_mark(VoxGeom g, const float* __restrict__ pts, int64_t n, int stride,
                                              const int32_t* __restrict__ frame_off, uint32_t* bitmap,
                                              int32_t* coors_top, int32_t* coors_med, int32_t* coors_low) {
  __shared__ __align__(16) float tile[TPB * MAX_STRIDE];
  int64_t idx;
  float p[3];
  if (!load_tile(pts, n, stride, tile, idx, p)) return;
  PointKeys k;
  k.b = frame_of(frame_off, g.n_frames, idx);
  point_keys(g, p, k);
  const int64_t cell = top_cell(g, k.b, k.c[0][1], k.c[0][0]);
  const uint32_t bit = 1u << (cell & 31);
  uint32_t* w = bitmap + (cell >> 5);
  if (!(*(volatile uint32_t*)w & bit)) atomicOr(w, bit);
  int32_t* outs[3] = {coors_top, coors_med, coors_low};
#pragma unroll
  for (int s = 0; s < 3; ++s)
    if (outs[s]) reinterpret_cast<int4*>(outs[s])[idx] = make_int4(k.b, k.c[s][2], k.c[s][1], k.c[s][0]);
}

// ---------------------------------------------------------------- pass 2: rank the bitmap
constexpr int SCAN_ITEMS = 4;
constexpr int SCAN_CHUNK = TPB * SCAN_ITEMS;
constexpr int SCAN_MAX_BLOCKS = 16384;

__device__ __forceinline__ int block_base(const int32_t* __restrict__ sums, int* smem) {
  int v = 0;
  for (int i = threadIdx.x; i < (int)blockIdx.x; i += TPB) v += sums[i];
  v = gm_warp_sum_i(v);
  if ((threadIdx.x & 31) == 0) smem[threadIdx.x >> 5] = v;
  __syncthreads();
  int t = 0;
  for (int i = 0; i < TPB / 32; ++i) t += smem[i];
  __syncthreads();
  return t;
}

__global__ void __launch_bounds__(TPB) k_bitmap_sums(const uint32_t* __restrict__ bitmap, int n_words, int32_t* sums) {
  __shared__ int smem[TPB / 32];
  int v = 0;
  const int base = blockIdx.x * SCAN_CHUNK + threadIdx.x * SCAN_ITEMS;
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; ++i)
    if (base + i < n_words) v += __popc(bitmap[base + i]);
  v = gm_warp_sum_i(v);
  if ((threadIdx.x & 31) == 0) smem[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    for (int i = 0; i < TPB / 32; ++i) t += smem[i];
    sums[blockIdx.x] = t;
  }
}

__global__ void __launch_bounds__(TPB) k_bitmap_rank(VoxGeom g, const uint32_t* __restrict__ bitmap, int n_words,
                                                     const int32_t* __restrict__ sums, int32_t* word_rank,
                                                     int32_t* counts, int32_t* pillar_coors, float* pillar_mean,
                                                     uint32_t* med_mask, uint32_t* low_mask, int64_t cap) {
  __shared__ int smem[40];
  const int blk_base = block_base(sums, smem);
  const int w0 = blockIdx.x * SCAN_CHUNK + threadIdx.x * SCAN_ITEMS;
  uint32_t words[SCAN_ITEMS];
  int mine = 0;
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; ++i) {
    words[i] = (w0 + i < n_words) ? bitmap[w0 + i] : 0u;
    mine += __popc(words[i]);
  }
  int total;
  int rank = blk_base + gm_block_excl_scan(mine, &total, smem);
  if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) {
    const int v = blk_base + total;
    counts[0] = (int)min((int64_t)v, cap);
    counts[3] = v > cap ? 1 : 0;
  }
  const int cells_per_frame = g.grid[0][0] * g.grid[0][1];
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; ++i) {
    if (w0 + i >= n_words) break;
    word_rank[w0 + i] = rank;
    uint32_t w = words[i];
    while (w) {
      const int bit = __ffs(w) - 1;
      w &= w - 1;
      if (rank < cap) {
        const int64_t cell = ((int64_t)(w0 + i) << 5) + bit;
        const int b = (int)(cell / cells_per_frame);
        const int r = (int)(cell - (int64_t)b * cells_per_frame);
        if (pillar_coors)
          reinterpret_cast<int4*>(pillar_coors)[rank] = make_int4(b, 0, r / g.grid[0][0], r % g.grid[0][0]);
        if (pillar_mean) reinterpret_cast<float4*>(pillar_mean)[rank] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (med_mask) med_mask[rank] = 0u;
        if (low_mask) reinterpret_cast<uint4*>(low_mask)[rank] = make_uint4(0u, 0u, 0u, 0u);
      }
      ++rank;
    }
  }
}

// first pillar row of every frame: counts[4 + b] (b = 0..n_frames), so the host learns the per-sample pillar
// counts (needed by the random mask split) from the same single read that returns the totals
__global__ void k_frame_starts(VoxGeom g, const uint32_t* __restrict__ bitmap, const int32_t* __restrict__ word_rank,
                               int32_t* counts) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b > g.n_frames) return;
  if (b == g.n_frames) { counts[4 + b] = counts[0]; return; }
  const int64_t cell = (int64_t)b * g.grid[0][0] * g.grid[0][1];
  const uint32_t w = bitmap[cell >> 5];
  counts[4 + b] = word_rank[cell >> 5] + __popc(w & ((1u << (cell & 31)) - 1u));
}

// ---------------------------------------------------------------- pass 3: point -> pillar, pillar sums, slot masks
__global__ void __launch_bounds__(TPB) k_assign(VoxGeom g, const float* __restrict__ pts, int64_t n, int stride,
                                                const int32_t* __restrict__ frame_off,
                                                const uint32_t* __restrict__ bitmap,
                                                const int32_t* __restrict__ word_rank, int64_t cap,
                                                int32_t* point_pillar, float* pillar_mean, uint32_t* med_mask,
                                                uint32_t* low_mask) {
  __shared__ __align__(16) float tile[TPB * MAX_STRIDE];
  int64_t idx;
  float p[3];
  if (!load_tile(pts, n, stride, tile, idx, p)) return;
  PointKeys k;
  k.b = frame_of(frame_off, g.n_frames, idx);
  point_keys(g, p, k);
  const int pid = cell_rank(bitmap, word_rank, top_cell(g, k.b, k.c[0][1], k.c[0][0]));
  point_pillar[idx] = pid;
  if (pid < 0 || pid >= cap) return;
  red_add4(pillar_mean + 4 * (int64_t)pid, p[0], p[1], p[2], 1.0f);
  int64_t cell;
  int slot;
  sub_parent(g, 1, k, cell, slot);
  // missing parent aliases row 0 like the reference's zero table; with consistent power-of-two scales the parent is the
  // point's own pillar and the two extra bitmap look-ups disappear
  int par = g.parent_is_top ? pid : max(cell_rank(bitmap, word_rank, cell), 0);
  if (par < cap) {
    const uint32_t bit = 1u << slot;
    if (!(*(volatile uint32_t*)(med_mask + par) & bit)) atomicOr(med_mask + par, bit);
  }
  sub_parent(g, 2, k, cell, slot);
  par = g.parent_is_top ? pid : max(cell_rank(bitmap, word_rank, cell), 0);
  if (par < cap) {
    uint32_t* w = low_mask + 4 * (int64_t)par + (slot >> 5);
    const uint32_t bit = 1u << (slot & 31);
    if (!(*(volatile uint32_t*)w & bit)) atomicOr(w, bit);
  }
}

// ---------------------------------------------------------------- pass 4: CSR offsets of the sub-voxels
__global__ void __launch_bounds__(TPB) k_sub_sums(const int32_t* __restrict__ counts, const uint32_t* __restrict__ med_mask,
                                                  const uint32_t* __restrict__ low_mask, int32_t* sums) {
  __shared__ int smem[2][TPB / 32];
  const int n = counts[0];
  int vm = 0, vl = 0;
  const int base = blockIdx.x * SCAN_CHUNK + threadIdx.x * SCAN_ITEMS;
  if (blockIdx.x * SCAN_CHUNK < n) {
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i)
      if (base + i < n) {
        vm += __popc(med_mask[base + i]);
        const uint4 m = reinterpret_cast<const uint4*>(low_mask)[base + i];
        vl += __popc(m.x) + __popc(m.y) + __popc(m.z) + __popc(m.w);
      }
  }
  vm = gm_warp_sum_i(vm);
  vl = gm_warp_sum_i(vl);
  if ((threadIdx.x & 31) == 0) { smem[0][threadIdx.x >> 5] = vm; smem[1][threadIdx.x >> 5] = vl; }
  __syncthreads();
  if (threadIdx.x < 2) {
    int t = 0;
    for (int i = 0; i < TPB / 32; ++i) t += smem[threadIdx.x][i];
    sums[threadIdx.x * SCAN_MAX_BLOCKS + blockIdx.x] = t;
  }
}

__global__ void __launch_bounds__(TPB) k_sub_ptr(int32_t* counts, const uint32_t* __restrict__ med_mask,
                                                 const uint32_t* __restrict__ low_mask, const int32_t* __restrict__ sums,
                                                 float* pillar_mean, int32_t* med_ptr, int32_t* low_ptr, float* med_mean,
                                                 float* low_mean, int64_t sub_cap) {
  __shared__ int smem[40];
  const int n = counts[0];
  if (blockIdx.x * SCAN_CHUNK >= n && blockIdx.x != 0) return;
  const int base_m = block_base(sums, smem);
  const int base_l = block_base(sums + SCAN_MAX_BLOCKS, smem);
  const int v0 = blockIdx.x * SCAN_CHUNK + threadIdx.x * SCAN_ITEMS;
  int cm[SCAN_ITEMS], cl[SCAN_ITEMS], sm_ = 0, sl_ = 0;
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; ++i) {
    cm[i] = cl[i] = 0;
    if (v0 + i < n) {
      cm[i] = __popc(med_mask[v0 + i]);
      const uint4 m = reinterpret_cast<const uint4*>(low_mask)[v0 + i];
      cl[i] = __popc(m.x) + __popc(m.y) + __popc(m.z) + __popc(m.w);
    }
    sm_ += cm[i];
    sl_ += cl[i];
  }
  int tot_m, tot_l;
  int pm = base_m + gm_block_excl_scan(sm_, &tot_m, smem);
  int pl = base_l + gm_block_excl_scan(sl_, &tot_l, smem);
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; ++i) {
    const int v = v0 + i;
    if (v < n) {
      med_ptr[v] = pm;
      low_ptr[v] = pl;
      // finish the pillar mean (sum -> mean), zero this pillar's sub-voxel accumulators
      float4* pmn = reinterpret_cast<float4*>(pillar_mean) + v;
      float4 a = *pmn;
      a.x = __fdiv_rn(a.x, a.w); a.y = __fdiv_rn(a.y, a.w); a.z = __fdiv_rn(a.z, a.w);
      *pmn = a;
      for (int j = 0; j < cm[i]; ++j)
        if (pm + j < sub_cap) reinterpret_cast<float4*>(med_mean)[pm + j] = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int j = 0; j < cl[i]; ++j)
        if (pl + j < sub_cap) reinterpret_cast<float4*>(low_mean)[pl + j] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    pm += cm[i];
    pl += cl[i];
    if (v == n - 1) {
      med_ptr[n] = pm;
      low_ptr[n] = pl;
      counts[1] = pm;
      counts[2] = pl;
    }
  }
  if (n == 0 && blockIdx.x == 0 && threadIdx.x == 0) {
    med_ptr[0] = low_ptr[0] = 0;
    counts[1] = counts[2] = 0;
  }
}

// ---------------------------------------------------------------- pass 5: sub-voxel sums
__global__ void __launch_bounds__(TPB) k_sub_accum(VoxGeom g, const float* __restrict__ pts, int64_t n, int stride,
                                                   const int32_t* __restrict__ frame_off,
                                                   const uint32_t* __restrict__ bitmap,
                                                   const int32_t* __restrict__ word_rank, int64_t cap, int64_t sub_cap,
                                                   const uint32_t* __restrict__ med_mask,
                                                   const uint32_t* __restrict__ low_mask,
                                                   const int32_t* __restrict__ med_ptr,
                                                   const int32_t* __restrict__ low_ptr, float* med_mean,
                                                   float* low_mean) {
  __shared__ __align__(16) float tile[TPB * MAX_STRIDE];
  int64_t idx;
  float p[3];
  if (!load_tile(pts, n, stride, tile, idx, p)) return;
  PointKeys k;
  k.b = frame_of(frame_off, g.n_frames, idx);
  point_keys(g, p, k);
  int64_t cell;
  int slot;
  sub_parent(g, 1, k, cell, slot);
  int par = max(cell_rank(bitmap, word_rank, cell), 0);
  const int par_med = par;
  if (par < cap) {
    const int row = __ldg(med_ptr + par) + __popc(__ldg(med_mask + par) & ((1u << slot) - 1u));
    if (row < sub_cap) red_add4(med_mean + 4 * (int64_t)row, p[0], p[1], p[2], 1.0f);
  }
  sub_parent(g, 2, k, cell, slot);
  par = g.parent_is_top ? par_med : max(cell_rank(bitmap, word_rank, cell), 0);
  if (par < cap) {
    const uint4 m = __ldg(reinterpret_cast<const uint4*>(low_mask) + par);
    const int row = __ldg(low_ptr + par) + rank128(m, slot);
    if (row < sub_cap) red_add4(low_mean + 4 * (int64_t)row, p[0], p[1], p[2], 1.0f);
  }
}

__global__ void __launch_bounds__(TPB) k_sub_finalize(const int32_t* __restrict__ counts, float* med_mean,
                                                      float* low_mean, int64_t sub_cap) {
  const int nm = (int)min((int64_t)counts[1], sub_cap), nl = (int)min((int64_t)counts[2], sub_cap);
  for (int i = blockIdx.x * TPB + threadIdx.x; i < nm + nl; i += gridDim.x * TPB) {
    float4* q = reinterpret_cast<float4*>(i < nm ? med_mean : low_mean) + (i < nm ? i : i - nm);
    float4 a = *q;
    a.x = __fdiv_rn(a.x, a.w); a.y = __fdiv_rn(a.y, a.w); a.z = __fdiv_rn(a.z, a.w);
    *q = a;
  }
}

// ---------------------------------------------------------------- occupancy from explicit (b,z,y,x) rows
__global__ void __launch_bounds__(TPB) k_mark_coors(VoxGeom g, const int32_t* __restrict__ coors, int64_t n,
                                                    uint32_t* bitmap) {
  const int64_t i = (int64_t)blockIdx.x * TPB + threadIdx.x;
  if (i >= n) return;
  const int4 c = __ldg(reinterpret_cast<const int4*>(coors) + i);
  const int64_t cell = top_cell(g, c.x, c.z, c.w);
  atomicOr(bitmap + (cell >> 5), 1u << (cell & 31));
}

__global__ void __launch_bounds__(TPB) k_token_of_cell(VoxGeom g, const int32_t* __restrict__ coors, int64_t n,
                                                       const uint32_t* __restrict__ bitmap,
                                                       const int32_t* __restrict__ word_rank, int32_t* tok_of_pillar) {
  const int64_t i = (int64_t)blockIdx.x * TPB + threadIdx.x;
  if (i >= n) return;
  const int4 c = __ldg(reinterpret_cast<const int4*>(coors) + i);
  tok_of_pillar[cell_rank(bitmap, word_rank, top_cell(g, c.x, c.z, c.w))] = (int32_t)i;
}

// ---------------------------------------------------------------- standalone dynamic_voxelize (a1)
__global__ void __launch_bounds__(TPB) k_dynamic_voxelize(const float* __restrict__ pts, int64_t n, int stride, float lx,
                                                          float ly, float lz, float vx, float vy, float vz, int gx,
                                                          int gy, int gz, int32_t* coors) {
  __shared__ __align__(16) float tile[TPB * MAX_STRIDE];
  int64_t idx;
  float p[3];
  if (!load_tile(pts, n, stride, tile, idx, p)) return;
  coors[idx * 3 + 0] = vox_coord(p[2], lz, vz, gz);
  coors[idx * 3 + 1] = vox_coord(p[1], ly, vy, gy);
  coors[idx * 3 + 2] = vox_coord(p[0], lx, vx, gx);
}

}  // namespace

extern "C" int geomae_grid_size(const float range_min[3], const float range_max[3], const float voxel[3],
                                int32_t grid_xyz[3]) {
  GM_REQUIRE(range_min && range_max && voxel && grid_xyz, "geomae_grid_size: null argument");
  for (int a = 0; a < 3; ++a) {
    GM_REQUIRE(voxel[a] > 0.f && range_max[a] > range_min[a], "geomae_grid_size: empty range/voxel on axis %d", a);
    volatile float q = (range_max[a] - range_min[a]) / voxel[a];
    grid_xyz[a] = (int32_t)ceilf(q);
  }
  return GEOMAE_OK;
}

extern "C" int geomae_dynamic_voxelize(const float* points, int64_t n, int32_t stride, const float voxel_xyz[3],
                                       const float range_min[3], const float range_max[3], int32_t* coors,
                                       void* stream) {
  GM_REQUIRE(n >= 0 && stride >= 3 && stride <= MAX_STRIDE, "dynamic_voxelize: stride %d not in 3..%d", stride,
             MAX_STRIDE);
  if (n == 0) return GEOMAE_OK;
  GM_REQUIRE(points && coors, "dynamic_voxelize: null buffer");
  int32_t grid[3];
  int rc = geomae_grid_size(range_min, range_max, voxel_xyz, grid);
  if (rc) return rc;
  k_dynamic_voxelize<<<gm_div_up(n, TPB), TPB, 0, (cudaStream_t)stream>>>(
      points, n, stride, range_min[0], range_min[1], range_min[2], voxel_xyz[0], voxel_xyz[1], voxel_xyz[2], grid[0],
      grid[1], grid[2], coors);
  GM_LAUNCH_CHECK();
  return GEOMAE_OK;
}

extern "C" int geomae_voxel_scatter(const geomae_voxel_cfg* cfg, const geomae_scatter_io* io, void* stream_) {
  GM_REQUIRE(cfg && io, "voxel_scatter: null cfg/io");
  cudaStream_t stream = (cudaStream_t)stream_;
  GM_REQUIRE(io->stride >= 3 && io->stride <= MAX_STRIDE, "voxel_scatter: stride %d not in 3..%d", io->stride,
             MAX_STRIDE);
  GM_REQUIRE(io->n_frames >= 1 && io->n_points >= 0 && io->cap >= 1, "voxel_scatter: bad sizes");
  GM_REQUIRE(io->n_points < (int64_t)1 << 31, "voxel_scatter: more than 2^31 points");
  GM_REQUIRE(io->points && io->frame_offsets && io->bitmap && io->word_rank && io->scan_tmp && io->counts &&
                 io->pillar_coors && io->pillar_mean && io->point_pillar && io->med_mask && io->low_mask &&
                 io->med_ptr && io->low_ptr && io->med_mean && io->low_mean,
             "voxel_scatter: null buffer");
  VoxGeom g;
  int rc = gm_make_geom(cfg, io->n_frames, &g);
  if (rc) return rc;
  const int64_t n_cells = (int64_t)io->n_frames * g.grid[0][0] * g.grid[0][1];
  const int n_words = gm_div_up(n_cells, 32);
  const int scan_blocks = gm_div_up(n_words, SCAN_CHUNK);
  const int sub_blocks = gm_div_up(io->cap, SCAN_CHUNK);
  GM_REQUIRE(scan_blocks <= SCAN_MAX_BLOCKS && sub_blocks <= SCAN_MAX_BLOCKS,
             "voxel_scatter: grid too large for one launch (%d / %d scan blocks)", scan_blocks, sub_blocks);
  const int64_t n = io->n_points;
  const int pblocks = gm_div_up(n > 0 ? n : 1, TPB);
  GM_CUDA(cudaMemsetAsync(io->bitmap, 0, (size_t)n_words * 4, stream));
  if (n > 0)
    k_mark<<<pblocks, TPB, 0, stream>>>(g, io->points, n, io->stride, io->frame_offsets, io->bitmap, io->coors_top,
                                        io->coors_med, io->coors_low);
  k_bitmap_sums<<<scan_blocks, TPB, 0, stream>>>(io->bitmap, n_words, io->scan_tmp);
  k_bitmap_rank<<<scan_blocks, TPB, 0, stream>>>(g, io->bitmap, n_words, io->scan_tmp, io->word_rank, io->counts,
                                                 io->pillar_coors, io->pillar_mean, io->med_mask, io->low_mask,
                                                 io->cap);
  k_frame_starts<<<gm_div_up(io->n_frames + 1, 64), 64, 0, stream>>>(g, io->bitmap, io->word_rank, io->counts);
  if (n > 0)
    k_assign<<<pblocks, TPB, 0, stream>>>(g, io->points, n, io->stride, io->frame_offsets, io->bitmap, io->word_rank,
                                          io->cap, io->point_pillar, io->pillar_mean, io->med_mask, io->low_mask);
  int32_t* sub_sums = io->scan_tmp + SCAN_MAX_BLOCKS;  // scan_tmp holds 3*SCAN_MAX_BLOCKS ints
  k_sub_sums<<<sub_blocks, TPB, 0, stream>>>(io->counts, io->med_mask, io->low_mask, sub_sums);
  k_sub_ptr<<<sub_blocks, TPB, 0, stream>>>(io->counts, io->med_mask, io->low_mask, sub_sums, io->pillar_mean,
                                            io->med_ptr, io->low_ptr, io->med_mean, io->low_mean, io->n_points);
  if (n > 0) {
    k_sub_accum<<<pblocks, TPB, 0, stream>>>(g, io->points, n, io->stride, io->frame_offsets, io->bitmap,
                                             io->word_rank, io->cap, io->n_points, io->med_mask, io->low_mask,
                                             io->med_ptr, io->low_ptr, io->med_mean, io->low_mean);
    k_sub_finalize<<<GM_NUM_SMS * 4, TPB, 0, stream>>>(io->counts, io->med_mean, io->low_mean, io->n_points);
  }
  GM_LAUNCH_CHECK();
  return GEOMAE_OK;
}

extern "C" int geomae_coors_bitmap(const geomae_voxel_cfg* cfg, const int32_t* coors, int64_t n, int32_t n_frames,
                                   uint32_t* bitmap, int32_t* word_rank, int32_t* scan_tmp, int32_t* counts,
                                   int32_t* tok_of_pillar, void* stream_) {
  GM_REQUIRE(cfg && bitmap && word_rank && scan_tmp && counts && tok_of_pillar && (coors || n == 0),
             "coors_bitmap: null argument");
  GM_REQUIRE(n_frames >= 1 && n >= 0, "coors_bitmap: bad sizes");
  cudaStream_t stream = (cudaStream_t)stream_;
  VoxGeom g;
  int rc = gm_make_geom(cfg, n_frames, &g);
  if (rc) return rc;
  const int64_t n_cells = (int64_t)n_frames * g.grid[0][0] * g.grid[0][1];
  const int n_words = gm_div_up(n_cells, 32);
  const int scan_blocks = gm_div_up(n_words, SCAN_CHUNK);
  GM_REQUIRE(scan_blocks <= SCAN_MAX_BLOCKS, "coors_bitmap: grid too large (%d scan blocks)", scan_blocks);
  GM_CUDA(cudaMemsetAsync(bitmap, 0, (size_t)n_words * 4, stream));
  if (n > 0) k_mark_coors<<<gm_div_up(n, TPB), TPB, 0, stream>>>(g, coors, n, bitmap);
  k_bitmap_sums<<<scan_blocks, TPB, 0, stream>>>(bitmap, n_words, scan_tmp);
  k_bitmap_rank<<<scan_blocks, TPB, 0, stream>>>(g, bitmap, n_words, scan_tmp, word_rank, counts, nullptr, nullptr,
                                                 nullptr, nullptr, n > 0 ? n : 1);
  if (n > 0) k_token_of_cell<<<gm_div_up(n, TPB), TPB, 0, stream>>>(g, coors, n, bitmap, word_rank, tok_of_pillar);
  GM_LAUNCH_CHECK();
  return GEOMAE_OK;
}
