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
// Point passes: one CTA stages a tile of TPB*PPT consecutive records (20 KB at the 5-float nuScenes record) with all
// of a thread's 128-bit loads in flight together, then every thread owns PPT points (tile-strided, so the shared
// reads of a 5-word record are bank-conflict free and the per-point global writes are coalesced).
constexpr int PPT = 4;
constexpr int TILE = TPB * PPT;
constexpr int TILE_VEC = TILE * MAX_STRIDE / 4 / TPB;  // 128-bit loads per thread at the widest record

struct TileInfo {
  int64_t p0;   // first point of the tile
  int nvalid;   // points in the tile
  int b0;       // frame of the first point
  bool one_frame;  // every point of the tile belongs to frame b0
};

__device__ __forceinline__ void tile_frame(TileInfo& t, const int32_t* __restrict__ frame_off, int n_frames) {
  t.b0 = 0;
  t.one_frame = true;
  if (frame_off) {
    t.b0 = frame_of(frame_off, n_frames, t.p0);
    t.one_frame = t.p0 + t.nvalid <= (int64_t)__ldg(frame_off + t.b0 + 1);
  }
}

__device__ __forceinline__ TileInfo load_tile(const float* __restrict__ pts, int64_t n, int stride,
                                              const int32_t* __restrict__ frame_off, int n_frames, float* tile) {
  TileInfo t;
  t.p0 = (int64_t)blockIdx.x * TILE;
  t.nvalid = (int)min((int64_t)TILE, n - t.p0);
  const int nfloat = t.nvalid * stride;
  const float* src = pts + t.p0 * stride;
  if ((reinterpret_cast<uintptr_t>(src) & 15u) == 0) {
    const int nvec = nfloat >> 2;
    const float4* src4 = reinterpret_cast<const float4*>(src);
    float4 v[TILE_VEC];
#pragma unroll
    for (int k = 0; k < TILE_VEC; ++k) {
      const int i = threadIdx.x + k * TPB;
      if (i < nvec) v[k] = __ldg(src4 + i);
    }
    // the frame of the tile's first point (one binary search per thread on broadcast addresses) resolves while the
    // tile loads are in flight; points then advance linearly from it
    tile_frame(t, frame_off, n_frames);
#pragma unroll
    for (int k = 0; k < TILE_VEC; ++k) {
      const int i = threadIdx.x + k * TPB;
      if (i < nvec) reinterpret_cast<float4*>(tile)[i] = v[k];
    }
    for (int i = (nvec << 2) + threadIdx.x; i < nfloat; i += TPB) tile[i] = __ldg(src + i);
  } else {
    for (int i = threadIdx.x; i < nfloat; i += TPB) tile[i] = __ldg(src + i);
    tile_frame(t, frame_off, n_frames);
  }
  __syncthreads();
  return t;
}

// frame of point idx, walking forward from the tile's first frame (empty frames are skipped like frame_of does)
__device__ __forceinline__ int frame_from(const int32_t* __restrict__ off, int n_frames, const TileInfo& t, int l) {
  int b = t.b0;
  if (!t.one_frame)
    while (b + 1 < n_frames && (int64_t)__ldg(off + b + 1) <= t.p0 + l) ++b;
  return b;
}

// Everything the passes need to know about one point.
struct PointInfo {
  int cell;            // BEV cell of the pillar
  int slot_m, slot_l;  // slot ids inside the parent pillar
  int cell_m, cell_l;  // parent BEV cells (== cell on the fast path)
};

// General path: independent IEEE divides per scale, parents by integer division of the sub-voxel coordinates.
__device__ __forceinline__ PointInfo point_info(const VoxGeom& g, const float* rec, int b, int4* c_top, int4* c_med,
                                                int4* c_low) {
  PointInfo r;
  PointKeys k;
  k.b = b;
  point_keys(g, rec, k);
  r.cell = (int)top_cell(g, b, k.c[0][1], k.c[0][0]);
  int64_t cell;
  sub_parent(g, 1, k, cell, r.slot_m);
  r.cell_m = (int)cell;
  sub_parent(g, 2, k, cell, r.slot_l);
  r.cell_l = (int)cell;
  if (c_top) *c_top = make_int4(b, k.c[0][2], k.c[0][1], k.c[0][0]);
  if (c_med) *c_med = make_int4(b, k.c[1][2], k.c[1][1], k.c[1][0]);
  if (c_low) *c_low = make_int4(b, k.c[2][2], k.c[2][1], k.c[2][0]);
  return r;
}

// Fast path (VoxGeom::fast): shifts and masks of one low-scale coordinate per axis; the sub-voxel parents are the
// point's own pillar.  Straight-line integer code.
__device__ __forceinline__ PointInfo point_info_fast(const VoxGeom& g, int cx, int cy, int cz, int b, int4* c_top,
                                                     int4* c_med, int4* c_low) {
  PointInfo r;
  const int tx = cx >> g.shift[0][0], ty = cy >> g.shift[0][1];
  const int mx = cx >> g.shift[1][0], my = cy >> g.shift[1][1], mz = cz >> g.shift[1][2];
  r.cell = r.cell_m = r.cell_l = (b * g.grid[0][1] + ty) * g.grid[0][0] + tx;
  r.slot_m = ((mz & g.smask[1][2]) << g.sshift[1][2]) | ((my & g.smask[1][1]) << g.sshift[1][1]) | (mx & g.smask[1][0]);
  r.slot_l = ((cz & g.smask[2][2]) << g.sshift[2][2]) | ((cy & g.smask[2][1]) << g.sshift[2][1]) | (cx & g.smask[2][0]);
  if (c_top) *c_top = make_int4(b, cz >> g.shift[0][2], ty, tx);
  if (c_med) *c_med = make_int4(b, mz, my, mx);
  if (c_low) *c_low = make_int4(b, cz, cy, cx);
  return r;
}

// PointInfo of the thread's PPT points.  The fast path first runs branch-free over all of them (reciprocal multiply,
// twelve independent chains the scheduler can interleave), then repairs the rare coordinates that sit within qeps of
// a voxel boundary with the IEEE divide.
template <bool FAST>
__device__ __forceinline__ void tile_point_info(const VoxGeom& g, const float* tile, int stride, const TileInfo& t,
                                                const int32_t* __restrict__ frame_off, bool need_frame,
                                                PointInfo* info, int32_t* coors_top, int32_t* coors_med,
                                                int32_t* coors_low) {
  int b[PPT];
#pragma unroll
  for (int j = 0; j < PPT; ++j) b[j] = t.b0;
  if (need_frame && !t.one_frame) {
#pragma unroll
    for (int j = 0; j < PPT; ++j) {
      const int l = threadIdx.x + j * TPB;
      if (l < t.nvalid) b[j] = frame_from(frame_off, g.n_frames, t, l);
    }
  }
  int cx[PPT], cy[PPT], cz[PPT];
  if (FAST) {
    unsigned redo = 0;
#pragma unroll
    for (int j = 0; j < PPT; ++j) {
      const int l = min(threadIdx.x + j * TPB, t.nvalid - 1);  // clamped: out-of-tile lanes compute a dummy
      bool rx, ry, rz;
      cx[j] = vox_coord_try(tile[l * stride], g.lo[0], g.rvs[0], g.qeps[0], g.grid[2][0], rx);
      cy[j] = vox_coord_try(tile[l * stride + 1], g.lo[1], g.rvs[1], g.qeps[1], g.grid[2][1], ry);
      cz[j] = vox_coord_try(tile[l * stride + 2], g.lo[2], g.rvs[2], g.qeps[2], g.grid[2][2], rz);
      redo |= (rx ? 1u : 0u) << (3 * j) | (ry ? 2u : 0u) << (3 * j) | (rz ? 4u : 0u) << (3 * j);
    }
    if (redo) {
#pragma unroll
      for (int j = 0; j < PPT; ++j) {
        const int l = min(threadIdx.x + j * TPB, t.nvalid - 1);
        if (redo >> (3 * j) & 1u) cx[j] = vox_coord(tile[l * stride], g.lo[0], g.vs[2][0], g.grid[2][0]);
        if (redo >> (3 * j) & 2u) cy[j] = vox_coord(tile[l * stride + 1], g.lo[1], g.vs[2][1], g.grid[2][1]);
        if (redo >> (3 * j) & 4u) cz[j] = vox_coord(tile[l * stride + 2], g.lo[2], g.vs[2][2], g.grid[2][2]);
      }
    }
  }
#pragma unroll
  for (int j = 0; j < PPT; ++j) {
    const int l = threadIdx.x + j * TPB;
    const bool ok = l < t.nvalid;
    const int64_t i = t.p0 + l;
    int4* ct = (ok && coors_top) ? reinterpret_cast<int4*>(coors_top) + i : nullptr;
    int4* cm = (ok && coors_med) ? reinterpret_cast<int4*>(coors_med) + i : nullptr;
    int4* cl = (ok && coors_low) ? reinterpret_cast<int4*>(coors_low) + i : nullptr;
    if (FAST) {
      info[j] = point_info_fast(g, cx[j], cy[j], cz[j], b[j], ct, cm, cl);
    } else if (ok) {
      const float rec[3] = {tile[l * stride], tile[l * stride + 1], tile[l * stride + 2]};
      info[j] = point_info(g, rec, b[j], ct, cm, cl);
    }
    if (!ok) info[j].cell = -1;
  }
}

__device__ __forceinline__ void red_add4(float* addr, float x, float y, float z, float w) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(x), "f"(y), "f"(z), "f"(w)
               : "memory");
}

// ---------------------------------------------------------------- pass 1: mark occupancy
template <bool FAST>
__global__ void __launch_bounds__(TPB, 6) k_mark(VoxGeom g, const float* __restrict__ pts, int64_t n, int stride,
                                              const int32_t* __restrict__ frame_off, uint32_t* bitmap,
                                              int32_t* coors_top, int32_t* coors_med, int32_t* coors_low) {
  extern __shared__ __align__(16) float tile[];
  const TileInfo t = load_tile(pts, n, stride, frame_off, g.n_frames, tile);
  PointInfo info[PPT];
  tile_point_info<FAST>(g, tile, stride, t, frame_off, true, info, coors_top, coors_med, coors_low);
  int cell[PPT];
#pragma unroll
  for (int j = 0; j < PPT; ++j) cell[j] = info[j].cell;
  uint32_t seen[PPT];
#pragma unroll
  for (int j = 0; j < PPT; ++j) seen[j] = cell[j] >= 0 ? *(volatile uint32_t*)(bitmap + (cell[j] >> 5)) : ~0u;
#pragma unroll
  for (int j = 0; j < PPT; ++j) {
    const uint32_t bit = 1u << (cell[j] & 31);
    if (!(seen[j] & bit)) atomicOr(bitmap + (cell[j] >> 5), bit);
  }
}

// ---------------------------------------------------------------- pass 2: rank the bitmap
constexpr int SCAN_ITEMS = 4;
constexpr int SCAN_CHUNK = TPB * SCAN_ITEMS;
constexpr int SCAN_MAX_BLOCKS = 16384;
constexpr int RANK_LIST = 2048;  // set cells expanded per round of k_bitmap_rank

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

// word ranks, the pillar rows (coordinates + zeroed accumulators, written row-coalesced from a shared list of the
// CTA's set cells) and counts[4 + b] = first pillar row of frame b (b = 0..n_frames) — the host learns the per-sample
// pillar counts (needed by the random mask split) from the same single read that returns the totals
__global__ void __launch_bounds__(TPB) k_bitmap_rank(VoxGeom g, const uint32_t* __restrict__ bitmap, int n_words,
                                                     const int32_t* __restrict__ sums, int32_t* word_rank,
                                                     int32_t* counts, int32_t* pillar_coors, float* pillar_mean,
                                                     uint32_t* med_mask, uint32_t* low_mask, int64_t cap,
                                                     int write_frame_starts, int zero_acc) {
  __shared__ int smem[40];
  __shared__ uint32_t list[RANK_LIST];
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
  const int local = gm_block_excl_scan(mine, &total, smem);  // rank inside the CTA of this thread's first set cell
  const uint32_t cells_per_frame = (uint32_t)(g.grid[0][0] * g.grid[0][1]);
  if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) {
    const int v = blk_base + total;
    counts[0] = (int)min((int64_t)v, cap);
    counts[3] = v > cap ? 1 : 0;
    if (write_frame_starts) counts[4 + g.n_frames] = counts[0];
  }
  {
    int r = blk_base + local;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
      if (w0 + i < n_words) word_rank[w0 + i] = r;
      r += __popc(words[i]);
    }
  }
  if (write_frame_starts && w0 < n_words) {
    // frames whose first cell lies in this thread's SCAN_ITEMS*32 cells
    const uint32_t c0 = (uint32_t)w0 << 5, c1 = c0 + SCAN_ITEMS * 32;
    for (uint32_t b = (c0 + cells_per_frame - 1) / cells_per_frame; b < (uint32_t)g.n_frames; ++b) {
      const uint32_t cell = b * cells_per_frame;
      if (cell >= c1) break;
      int r = blk_base + local;
      const int wi = (int)((cell - c0) >> 5);
#pragma unroll
      for (int i = 0; i < SCAN_ITEMS; ++i)
        if (i < wi) r += __popc(words[i]);
        else if (i == wi) r += __popc(words[i] & ((1u << (cell & 31)) - 1u));
      counts[4 + b] = (int)min((int64_t)r, cap);
    }
  }
  if (!pillar_coors) return;
  for (int r0 = 0; r0 < total; r0 += RANK_LIST) {
    int r = local;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
      uint32_t w = words[i];
      while (w) {
        const int bit = __ffs(w) - 1;
        w &= w - 1;
        if (r >= r0 && r < r0 + RANK_LIST) list[r - r0] = ((uint32_t)(w0 + i) << 5) + bit;
        ++r;
      }
    }
    __syncthreads();
    const int cnt = min(RANK_LIST, total - r0);
    for (int i = threadIdx.x; i < cnt; i += TPB) {
      const int64_t row = (int64_t)blk_base + r0 + i;
      if (row >= cap) break;
      const uint32_t cell = list[i];
      const uint32_t b = cell / cells_per_frame, rem = cell - b * cells_per_frame;
      const uint32_t y = rem / (uint32_t)g.grid[0][0], x = rem - y * (uint32_t)g.grid[0][0];
      reinterpret_cast<int4*>(pillar_coors)[row] = make_int4((int)b, 0, (int)y, (int)x);
      if (zero_acc) {  // the hierarchical (fast) path never accumulates into these
        reinterpret_cast<float4*>(pillar_mean)[row] = make_float4(0.f, 0.f, 0.f, 0.f);
        med_mask[row] = 0u;
      }
      reinterpret_cast<uint4*>(low_mask)[row] = make_uint4(0u, 0u, 0u, 0u);
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------- pass 3: point -> pillar, pillar sums, slot masks
__device__ __forceinline__ int cell_rank32(const uint32_t* __restrict__ bitmap, const int32_t* __restrict__ word_rank,
                                           int cell) {
  const uint32_t w = __ldg(bitmap + (cell >> 5));
  const uint32_t bit = 1u << (cell & 31);
  return (w & bit) ? __ldg(word_rank + (cell >> 5)) + __popc(w & (bit - 1)) : -1;
}

template <bool FAST>
__global__ void __launch_bounds__(TPB, 6) k_assign(VoxGeom g, const float* __restrict__ pts, int64_t n, int stride,
                                                const int32_t* __restrict__ frame_off,
                                                const uint32_t* __restrict__ bitmap,
                                                const int32_t* __restrict__ word_rank, int64_t cap,
                                                int32_t* point_pillar, float* pillar_mean, uint32_t* med_mask,
                                                uint32_t* low_mask) {
  extern __shared__ __align__(16) float tile[];
  const TileInfo t = load_tile(pts, n, stride, frame_off, g.n_frames, tile);
  // phase 1: all look-ups of the thread's points in flight together (the reductions below are ordering barriers)
  PointInfo info[PPT];
  tile_point_info<FAST>(g, tile, stride, t, frame_off, true, info, nullptr, nullptr, nullptr);
  int pid[PPT], par_m[PPT], par_l[PPT], slot_m[PPT], slot_l[PPT];
#pragma unroll
  for (int j = 0; j < PPT; ++j) {
    pid[j] = par_m[j] = par_l[j] = -1;
    slot_m[j] = info[j].slot_m;
    slot_l[j] = info[j].slot_l;
    if (info[j].cell >= 0) {
      pid[j] = cell_rank32(bitmap, word_rank, info[j].cell);
      // missing parent aliases row 0 like the reference's zero table; on the fast path the parent is the point's own
      // pillar and the two extra bitmap look-ups disappear
      par_m[j] = FAST ? pid[j] : max(cell_rank32(bitmap, word_rank, info[j].cell_m), 0);
      par_l[j] = FAST ? pid[j] : max(cell_rank32(bitmap, word_rank, info[j].cell_l), 0);
    }
  }
  uint32_t seen_m[PPT], seen_l[PPT];
#pragma unroll
  for (int j = 0; j < PPT; ++j) {
    const bool ok = pid[j] >= 0 && pid[j] < cap;
    seen_m[j] = (!FAST && ok && par_m[j] < cap) ? *(volatile uint32_t*)(med_mask + par_m[j]) : ~0u;
    seen_l[j] = (ok && par_l[j] < cap) ? *(volatile uint32_t*)(low_mask + 4 * (int64_t)par_l[j] + (slot_l[j] >> 5)) : ~0u;
  }
  // phase 2: writes.  Fast path: only the low-scale slot bit — the middle-scale masks and all three levels of sums
  // follow hierarchically from the low-scale sub-voxels (k_sub_sums / k_sub_reduce), which halves the L2 reductions
  // these passes are bound by.
#pragma unroll
  for (int j = 0; j < PPT; ++j) {
    const int l = threadIdx.x + j * TPB;
    if (l >= t.nvalid) continue;
    point_pillar[t.p0 + l] = pid[j];
    if (pid[j] < 0 || pid[j] >= cap) continue;
    if (!FAST) {
      red_add4(pillar_mean + 4 * (int64_t)pid[j], tile[l * stride], tile[l * stride + 1], tile[l * stride + 2], 1.0f);
      const uint32_t bm = 1u << slot_m[j];
      if (!(seen_m[j] & bm)) atomicOr(med_mask + par_m[j], bm);
    }
    const uint32_t bl = 1u << (slot_l[j] & 31);
    if (!(seen_l[j] & bl)) atomicOr(low_mask + 4 * (int64_t)par_l[j] + (slot_l[j] >> 5), bl);
  }
}

// ---------------------------------------------------------------- pass 4: CSR offsets of the sub-voxels
__global__ void __launch_bounds__(TPB) k_sub_sums(VoxGeom g, const int32_t* __restrict__ counts, uint32_t* med_mask,
                                                  const uint32_t* __restrict__ low_mask, int32_t* sums) {
  __shared__ int smem[2][TPB / 32];
  const int n = counts[0];
  int vm = 0, vl = 0;
  const int base = blockIdx.x * SCAN_CHUNK + threadIdx.x * SCAN_ITEMS;
  if (blockIdx.x * SCAN_CHUNK < n) {
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i)
      if (base + i < n) {
        const uint4 m = reinterpret_cast<const uint4*>(low_mask)[base + i];
        vl += __popc(m.x) + __popc(m.y) + __popc(m.z) + __popc(m.w);
        uint32_t mm;
        if (g.fast) {  // the middle-scale occupancy follows from the low-scale one
          mm = med_mask_of_low(g, m);
          med_mask[base + i] = mm;
        } else {
          mm = med_mask[base + i];
        }
        vm += __popc(mm);
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
                                                 float* low_mean, int64_t sub_cap, int hierarchical) {
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
  const int pm0 = pm;
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; ++i) {
    const int v = v0 + i;
    if (v < n) {
      med_ptr[v] = pm;
      low_ptr[v] = pl;
      if (!hierarchical) {  // finish the pillar mean (sum -> mean); the hierarchical path does it in k_sub_reduce
        float4* pmn = reinterpret_cast<float4*>(pillar_mean) + v;
        float4 a = *pmn;
        a.x = __fdiv_rn(a.x, a.w); a.y = __fdiv_rn(a.y, a.w); a.z = __fdiv_rn(a.z, a.w);
        *pmn = a;
      }
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
  if (hierarchical) {
    // the hierarchical path only accumulates into the low-scale rows; each middle-scale row is tagged with its pillar
    // (in the count lane, overwritten by k_med_reduce) so that pass can run one thread per row
    int row = pm0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i)
      for (int j = 0; j < cm[i]; ++j, ++row)
        if (row < sub_cap) med_mean[4 * (int64_t)row + 3] = __int_as_float(v0 + i);
  }
  // zero the CTA's contiguous range of sub-voxel accumulators, row-coalesced
  for (int i = threadIdx.x; i < (hierarchical ? 0 : tot_m); i += TPB)
    if (base_m + i < sub_cap) reinterpret_cast<float4*>(med_mean)[base_m + i] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int i = threadIdx.x; i < tot_l; i += TPB)
    if (base_l + i < sub_cap) reinterpret_cast<float4*>(low_mean)[base_l + i] = make_float4(0.f, 0.f, 0.f, 0.f);
  if (n == 0 && blockIdx.x == 0 && threadIdx.x == 0) {
    med_ptr[0] = low_ptr[0] = 0;
    counts[1] = counts[2] = 0;
  }
}

// ---------------------------------------------------------------- pass 5: sub-voxel sums
template <bool FAST>
__global__ void __launch_bounds__(TPB, 6) k_sub_accum(VoxGeom g, const float* __restrict__ pts, int64_t n, int stride,
                                                   const int32_t* __restrict__ frame_off,
                                                   const uint32_t* __restrict__ bitmap,
                                                   const int32_t* __restrict__ word_rank,
                                                   const int32_t* __restrict__ point_pillar, int64_t cap,
                                                   int64_t sub_cap, const uint32_t* __restrict__ med_mask,
                                                   const uint32_t* __restrict__ low_mask,
                                                   const int32_t* __restrict__ med_ptr,
                                                   const int32_t* __restrict__ low_ptr, float* med_mean,
                                                   float* low_mean) {
  extern __shared__ __align__(16) float tile[];
  // on the fast path the parent is the stored point -> pillar row: no frame index, no bitmap look-up
  const TileInfo t = load_tile(pts, n, stride, FAST ? nullptr : frame_off, g.n_frames, tile);
  int row_m[PPT], row_l[PPT];
  int par_m[PPT], par_l[PPT], slot_m[PPT], slot_l[PPT];
  // phase 1a: parents
  PointInfo info[PPT];
  tile_point_info<FAST>(g, tile, stride, t, frame_off, !FAST, info, nullptr, nullptr, nullptr);
#pragma unroll
  for (int j = 0; j < PPT; ++j) {
    const int l = threadIdx.x + j * TPB;
    par_m[j] = par_l[j] = -1;
    slot_m[j] = info[j].slot_m;
    slot_l[j] = info[j].slot_l;
    if (info[j].cell >= 0) {
      if (FAST) {
        par_m[j] = par_l[j] = __ldg(point_pillar + t.p0 + l);
      } else {
        par_m[j] = max(cell_rank32(bitmap, word_rank, info[j].cell_m), 0);
        par_l[j] = max(cell_rank32(bitmap, word_rank, info[j].cell_l), 0);
      }
      if (par_m[j] >= cap) par_m[j] = -1;
      if (par_l[j] >= cap) par_l[j] = -1;
    }
  }
  // phase 1b: CSR rows
#pragma unroll
  for (int j = 0; j < PPT; ++j) {
    row_m[j] = row_l[j] = -1;
    if (!FAST && par_m[j] >= 0)
      row_m[j] = __ldg(med_ptr + par_m[j]) + __popc(__ldg(med_mask + par_m[j]) & ((1u << slot_m[j]) - 1u));
    if (par_l[j] >= 0) {
      const uint4 m = __ldg(reinterpret_cast<const uint4*>(low_mask) + par_l[j]);
      row_l[j] = __ldg(low_ptr + par_l[j]) + rank128(m, slot_l[j]);
    }
  }
  // phase 2: reductions
#pragma unroll
  for (int j = 0; j < PPT; ++j) {
    const int l = threadIdx.x + j * TPB;
    if (l >= t.nvalid) continue;
    const float x = tile[l * stride], y = tile[l * stride + 1], z = tile[l * stride + 2];
    if (!FAST && row_m[j] >= 0 && row_m[j] < sub_cap) red_add4(med_mean + 4 * (int64_t)row_m[j], x, y, z, 1.0f);
    if (row_l[j] >= 0 && row_l[j] < sub_cap) red_add4(low_mean + 4 * (int64_t)row_l[j], x, y, z, 1.0f);
  }
}

// Fast path, after the low-scale sums are complete.  k_med_reduce, one thread per middle-scale row: its sum = the sum
// of its (nested) low-scale children, visited in slot order; the children turn from (sum, count) into (mean, count)
// on the way.  k_top_reduce, one thread per pillar: pillar sum = sum of its middle-scale rows in slot order; both
// levels become means.  The result depends only on the low-scale sums.
// (sum, count) -> (mean, count).  One correctly rounded reciprocal of the (small integer) count serves the three
// quotients: q = x * r, then q + fma(-w, q, x) * r — Markstein's correction, which rounds like the IEEE divide.
__device__ __forceinline__ float4 mean_of(const float4 a) {
  const float r = __frcp_rn(a.w);
  const float qx = a.x * r, qy = a.y * r, qz = a.z * r;
  return make_float4(fmaf(fmaf(-a.w, qx, a.x), r, qx), fmaf(fmaf(-a.w, qy, a.y), r, qy),
                     fmaf(fmaf(-a.w, qz, a.z), r, qz), a.w);
}

__global__ void __launch_bounds__(TPB) k_med_reduce(VoxGeom g, const int32_t* __restrict__ counts,
                                                    const uint32_t* __restrict__ med_mask,
                                                    const uint32_t* __restrict__ low_mask,
                                                    const int32_t* __restrict__ med_ptr,
                                                    const int32_t* __restrict__ low_ptr, float* med_mean,
                                                    float* low_mean, int64_t sub_cap) {
  const int64_t row = (int64_t)blockIdx.x * TPB + threadIdx.x;
  if (row >= min((int64_t)counts[1], sub_cap)) return;
  const int v = __float_as_int(med_mean[4 * row + 3]);  // pillar tag written by k_sub_ptr
  const uint4 ml = __ldg(reinterpret_cast<const uint4*>(low_mask) + v);
  const int lp = __ldg(low_ptr + v);
  const int ms = (int)__fns(__ldg(med_mask + v), 0, (int)(row - __ldg(med_ptr + v)) + 1);  // slot of this row
  const int mx = ms & g.smask[1][0], my = (ms >> g.sshift[1][1]) & g.smask[1][1], mz = (ms >> g.sshift[1][2]) & g.smask[1][2];
  const int sx = g.shift[1][0], sy = g.shift[1][1], sz = g.shift[1][2];
  const int nchild = 1 << (sx + sy + sz);               // low sub-voxels per middle one (<= 32 on the fast path)
  const int p1 = __popc(ml.x), p2 = p1 + __popc(ml.y), p3 = p2 + __popc(ml.z);
  // slot of child c = s0 + (dz, dy, dx) placed in the low-scale bit fields; the fields of s0 have their low bits clear
  const int s0 = ((mz << sz) << g.sshift[2][2]) | ((my << sy) << g.sshift[2][1]) | (mx << sx);
  auto child_slot = [&](int c) {
    return s0 + ((c >> (sx + sy)) << g.sshift[2][2]) + (((c >> sx) & ((1 << sy) - 1)) << g.sshift[2][1]) + (c & ((1 << sx) - 1));
  };
  auto word_of = [&](int wi) { return wi == 0 ? ml.x : wi == 1 ? ml.y : wi == 2 ? ml.z : ml.w; };
  // present children first (cheap bit tests), then a loop that only visits those: a warp iterates max-over-lanes of
  // the PRESENT children (typically 2-4), not over all candidates
  uint32_t present = 0u;
  for (int c = 0; c < nchild; ++c) {
    const int s = child_slot(c);
    present |= ((word_of(s >> 5) >> (s & 31)) & 1u) << c;
  }
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (; present; present &= present - 1) {               // ascending child index = slot order
    const int s = child_slot(__ffs(present) - 1);
    const int wi = s >> 5;
    const int64_t lrow = (int64_t)lp + (wi == 0 ? 0 : wi == 1 ? p1 : wi == 2 ? p2 : p3) +
                         __popc(word_of(wi) & ((1u << (s & 31)) - 1u));
    if (lrow >= sub_cap) continue;
    float4* q = reinterpret_cast<float4*>(low_mean) + lrow;
    const float4 a = *q;
    acc.x += a.x; acc.y += a.y; acc.z += a.z; acc.w += a.w;
    *q = mean_of(a);
  }
  reinterpret_cast<float4*>(med_mean)[row] = acc;
}

__global__ void __launch_bounds__(TPB) k_top_reduce(const int32_t* __restrict__ counts,
                                                    const int32_t* __restrict__ med_ptr, float* pillar_mean,
                                                    float* med_mean, int64_t sub_cap) {
  const int v = blockIdx.x * TPB + threadIdx.x;
  if (v >= counts[0]) return;
  const int r0 = __ldg(med_ptr + v), r1 = (int)min((int64_t)__ldg(med_ptr + v + 1), sub_cap);
  float4 top = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int r = r0; r < r1; r += 8) {
    float4 a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
      a[i] = r + i < r1 ? reinterpret_cast<const float4*>(med_mean)[r + i] : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int i = 0; i < 8; ++i)
      if (r + i < r1) {
        top.x += a[i].x; top.y += a[i].y; top.z += a[i].z; top.w += a[i].w;
        reinterpret_cast<float4*>(med_mean)[r + i] = mean_of(a[i]);
      }
  }
  reinterpret_cast<float4*>(pillar_mean)[v] = mean_of(top);
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

__global__ void __launch_bounds__(TPB) k_rank_of_row(VoxGeom g, const int32_t* __restrict__ coors, int64_t n,
                                                     const uint32_t* __restrict__ bitmap,
                                                     const int32_t* __restrict__ word_rank, int32_t* rank_of_row,
                                                     int32_t* first_row) {
  const int64_t i = (int64_t)blockIdx.x * TPB + threadIdx.x;
  if (i >= n) return;
  const int4 c = __ldg(reinterpret_cast<const int4*>(coors) + i);
  const int r = cell_rank(bitmap, word_rank, top_cell(g, c.x, c.z, c.w));
  rank_of_row[i] = r;
  if (first_row) atomicMin(first_row + r, (int32_t)i);
}

// ---------------------------------------------------------------- standalone dynamic_voxelize (a1)
__global__ void __launch_bounds__(TPB) k_dynamic_voxelize(const float* __restrict__ pts, int64_t n, int stride, float lx,
                                                          float ly, float lz, float vx, float vy, float vz, int gx,
                                                          int gy, int gz, int32_t* coors) {
  extern __shared__ __align__(16) float tile[];
  const TileInfo t = load_tile(pts, n, stride, nullptr, 0, tile);
  int32_t* out = coors + t.p0 * 3;
#pragma unroll
  for (int j = 0; j < PPT; ++j) {
    const int l = threadIdx.x + j * TPB;
    if (l < t.nvalid) {
      out[l * 3 + 0] = vox_coord(tile[l * stride + 2], lz, vz, gz);
      out[l * 3 + 1] = vox_coord(tile[l * stride + 1], ly, vy, gy);
      out[l * 3 + 2] = vox_coord(tile[l * stride + 0], lx, vx, gx);
    }
  }
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
  k_dynamic_voxelize<<<gm_div_up(n, TILE), TPB, (size_t)TILE * stride * sizeof(float), (cudaStream_t)stream>>>(
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
  GM_REQUIRE(n_cells < ((int64_t)1 << 31), "voxel_scatter: %lld BEV cells in the batch, supported < 2^31",
             (long long)n_cells);
  const int64_t n = io->n_points;
  const int pblocks = gm_div_up(n > 0 ? n : 1, TILE);
  const size_t tile_bytes = (size_t)TILE * io->stride * sizeof(float);
  GM_CUDA(cudaMemsetAsync(io->bitmap, 0, (size_t)n_words * 4, stream));
  GM_CUDA(cudaMemsetAsync(io->counts, 0, sizeof(int32_t) * (size_t)(4 + io->n_frames + 1), stream));
  if (n > 0)
    (g.fast ? k_mark<true> : k_mark<false>)<<<pblocks, TPB, tile_bytes, stream>>>(
        g, io->points, n, io->stride, io->frame_offsets, io->bitmap, io->coors_top, io->coors_med, io->coors_low);
  k_bitmap_sums<<<scan_blocks, TPB, 0, stream>>>(io->bitmap, n_words, io->scan_tmp);
  k_bitmap_rank<<<scan_blocks, TPB, 0, stream>>>(g, io->bitmap, n_words, io->scan_tmp, io->word_rank, io->counts,
                                                 io->pillar_coors, io->pillar_mean, io->med_mask, io->low_mask,
                                                 io->cap, 1, g.fast ? 0 : 1);
  if (n > 0)
    (g.fast ? k_assign<true> : k_assign<false>)<<<pblocks, TPB, tile_bytes, stream>>>(
        g, io->points, n, io->stride, io->frame_offsets, io->bitmap, io->word_rank, io->cap, io->point_pillar,
        io->pillar_mean, io->med_mask, io->low_mask);
  int32_t* sub_sums = io->scan_tmp + SCAN_MAX_BLOCKS;  // scan_tmp holds 3*SCAN_MAX_BLOCKS ints
  k_sub_sums<<<sub_blocks, TPB, 0, stream>>>(g, io->counts, io->med_mask, io->low_mask, sub_sums);
  k_sub_ptr<<<sub_blocks, TPB, 0, stream>>>(io->counts, io->med_mask, io->low_mask, sub_sums, io->pillar_mean,
                                            io->med_ptr, io->low_ptr, io->med_mean, io->low_mean, io->n_points,
                                            g.fast);
  if (n > 0) {
    (g.fast ? k_sub_accum<true> : k_sub_accum<false>)<<<pblocks, TPB, tile_bytes, stream>>>(
        g, io->points, n, io->stride, io->frame_offsets, io->bitmap, io->word_rank, io->point_pillar, io->cap,
        io->n_points, io->med_mask, io->low_mask, io->med_ptr, io->low_ptr, io->med_mean, io->low_mean);
    if (g.fast) {
      // sub-voxel and pillar counts are bounded by the point count; surplus CTAs exit on the device-side counts
      k_med_reduce<<<gm_div_up(io->n_points, TPB), TPB, 0, stream>>>(g, io->counts, io->med_mask, io->low_mask,
                                                                     io->med_ptr, io->low_ptr, io->med_mean,
                                                                     io->low_mean, io->n_points);
      k_top_reduce<<<gm_div_up(io->cap, TPB), TPB, 0, stream>>>(io->counts, io->med_ptr, io->pillar_mean,
                                                                io->med_mean, io->n_points);
    }
    else
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
  GM_CUDA(cudaMemsetAsync(counts, 0, sizeof(int32_t) * 4, stream));
  if (n > 0) k_mark_coors<<<gm_div_up(n, TPB), TPB, 0, stream>>>(g, coors, n, bitmap);
  k_bitmap_sums<<<scan_blocks, TPB, 0, stream>>>(bitmap, n_words, scan_tmp);
  k_bitmap_rank<<<scan_blocks, TPB, 0, stream>>>(g, bitmap, n_words, scan_tmp, word_rank, counts, nullptr, nullptr,
                                                 nullptr, nullptr, n > 0 ? n : 1, 0, 0);
  if (n > 0) k_token_of_cell<<<gm_div_up(n, TPB), TPB, 0, stream>>>(g, coors, n, bitmap, word_rank, tok_of_pillar);
  GM_LAUNCH_CHECK();
  return GEOMAE_OK;
}

extern "C" int geomae_coors_rank(const geomae_voxel_cfg* cfg, const int32_t* coors, int64_t n, int32_t n_frames,
                                 uint32_t* bitmap, int32_t* word_rank, int32_t* scan_tmp, int32_t* counts,
                                 int32_t* rank_of_row, int32_t* first_row, void* stream_) {
  GM_REQUIRE(cfg && bitmap && word_rank && scan_tmp && counts && rank_of_row && (coors || n == 0),
             "coors_rank: null argument");
  GM_REQUIRE(n_frames >= 1 && n >= 0 && n < ((int64_t)1 << 31), "coors_rank: bad sizes");
  cudaStream_t stream = (cudaStream_t)stream_;
  VoxGeom g;
  int rc = gm_make_geom(cfg, n_frames, &g);
  if (rc) return rc;
  const int64_t n_cells = (int64_t)n_frames * g.grid[0][0] * g.grid[0][1];
  GM_REQUIRE(n_cells < ((int64_t)1 << 31), "coors_rank: %lld cells, supported < 2^31", (long long)n_cells);
  const int n_words = gm_div_up(n_cells, 32);
  const int scan_blocks = gm_div_up(n_words, SCAN_CHUNK);
  GM_REQUIRE(scan_blocks <= SCAN_MAX_BLOCKS, "coors_rank: grid too large (%d scan blocks)", scan_blocks);
  GM_CUDA(cudaMemsetAsync(bitmap, 0, (size_t)n_words * 4, stream));
  GM_CUDA(cudaMemsetAsync(counts, 0, sizeof(int32_t) * 4, stream));
  if (first_row && n > 0) GM_CUDA(cudaMemsetAsync(first_row, 0x7f, (size_t)n * 4, stream));
  if (n > 0) k_mark_coors<<<gm_div_up(n, TPB), TPB, 0, stream>>>(g, coors, n, bitmap);
  k_bitmap_sums<<<scan_blocks, TPB, 0, stream>>>(bitmap, n_words, scan_tmp);
  k_bitmap_rank<<<scan_blocks, TPB, 0, stream>>>(g, bitmap, n_words, scan_tmp, word_rank, counts, nullptr, nullptr,
                                                 nullptr, nullptr, n > 0 ? n : 1, 0, 0);
  if (n > 0)
    k_rank_of_row<<<gm_div_up(n, TPB), TPB, 0, stream>>>(g, coors, n, bitmap, word_rank, rank_of_row, first_row);
  GM_LAUNCH_CHECK();
  return GEOMAE_OK;
}
