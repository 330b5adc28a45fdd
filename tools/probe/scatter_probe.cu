// Microbenchmark variants of the point passes of voxel_scatter.cu (NOT part of the product library): isolates where a
// pass's time goes — streaming structure, frame search, bitmap look-up, atomics.  Built by tools/probe/build.sh into
// tools/probe/scatter_probe.so and driven by tools/probe_scatter.py.
#include "../../geomae_b200/csrc/api.cu"
#include "../../geomae_b200/csrc/voxel_scatter.cu"

namespace {

// V0: plain streaming read, 4 x 128-bit loads per thread in flight, nothing else
__global__ void __launch_bounds__(256) p_stream(const float4* __restrict__ src, int64_t nvec, float* sink) {
  const int64_t base = (int64_t)blockIdx.x * 1024 + threadIdx.x;
  float acc = 0.f;
  float4 v[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) v[k] = base + k * 256 < nvec ? __ldg(src + base + k * 256) : make_float4(0, 0, 0, 0);
#pragma unroll
  for (int k = 0; k < 4; ++k) acc += v[k].x + v[k].y + v[k].z + v[k].w;
  if (acc == 1234.5f) sink[0] = acc;
}

// V1..V5: the tile structure of the product with pieces switched on
//   bit 0: frame look-up   bit 1: coordinates + cell   bit 2: bitmap look-up (volatile)   bit 3: atomicOr
//   bit 4: bitmap look-up with ld.cg instead of volatile   bit 5: unconditional atomicOr (no look-up needed)
template <int MODE>
__global__ void __launch_bounds__(TPB, 6) p_tile(VoxGeom g, const float* __restrict__ pts, int64_t n, int stride,
                                                 const int32_t* __restrict__ frame_off, uint32_t* bitmap, float* sink) {
  extern __shared__ __align__(16) float tile[];
  const TileInfo t = load_tile(pts, n, stride, (MODE & 1) ? frame_off : nullptr, g.n_frames, tile);
  float acc = 0.f;
  if (!(MODE & 2)) {
#pragma unroll
    for (int j = 0; j < PPT; ++j) {
      const int l = threadIdx.x + j * TPB;
      if (l < t.nvalid) acc += tile[l * stride] + tile[l * stride + 1] + tile[l * stride + 2];
    }
    acc += (float)t.b0;
    if (acc == 1234.5f) sink[0] = acc;
    return;
  }
  PointInfo info[PPT];
  tile_point_info<true>(g, tile, stride, t, frame_off, (MODE & 1) != 0, info, nullptr, nullptr, nullptr);
  int cell[PPT];
#pragma unroll
  for (int j = 0; j < PPT; ++j) cell[j] = info[j].cell;
  if (!(MODE & (4 | 16 | 32))) {
#pragma unroll
    for (int j = 0; j < PPT; ++j) acc += (float)cell[j];
    if (acc == 1234.5f) sink[0] = acc;
    return;
  }
  uint32_t seen[PPT];
#pragma unroll
  for (int j = 0; j < PPT; ++j) {
    seen[j] = ~0u;
    if (cell[j] >= 0) {
      if (MODE & 4) seen[j] = *(volatile uint32_t*)(bitmap + (cell[j] >> 5));
      if (MODE & 16) seen[j] = __ldcg(bitmap + (cell[j] >> 5));
      if (MODE & 32) seen[j] = 0u;
    }
  }
  if (MODE & (8 | 32)) {
#pragma unroll
    for (int j = 0; j < PPT; ++j) {
      const uint32_t bit = 1u << (cell[j] & 31);
      if (!(seen[j] & bit)) atomicOr(bitmap + (cell[j] >> 5), bit);
    }
  } else {
#pragma unroll
    for (int j = 0; j < PPT; ++j) acc += (float)seen[j];
    if (acc == 1234.5f) sink[0] = acc;
  }
}

// V6: no shared memory — a thread owns 4 consecutive 5-float records = 5 x 128-bit loads straight into registers
template <int MODE>
__global__ void __launch_bounds__(256) p_reg5(VoxGeom g, const float* __restrict__ pts, int64_t n,
                                              const int32_t* __restrict__ frame_off, uint32_t* bitmap, float* sink) {
  const int64_t q = (int64_t)blockIdx.x * 256 + threadIdx.x;  // quad of points
  if (q * 4 + 3 >= n) return;                                // probe only: tail ignored
  const float4* src = reinterpret_cast<const float4*>(pts) + q * 5;
  const float4 a = __ldg(src), b = __ldg(src + 1), c = __ldg(src + 2), d = __ldg(src + 3), e = __ldg(src + 4);
  const float px[4] = {a.x, b.y, c.z, d.w}, py[4] = {a.y, b.z, c.w, e.x}, pz[4] = {a.z, b.w, d.x, e.y};
  int fb = 0;
  if (MODE & 1) fb = frame_of(frame_off, g.n_frames, q * 4);
  int cell[4];
  unsigned redo = 0;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    bool rx, ry, rz;
    const int cx = vox_coord_try(px[j], g.lo[0], g.rvs[0], g.qeps[0], g.grid[2][0], rx);
    const int cy = vox_coord_try(py[j], g.lo[1], g.rvs[1], g.qeps[1], g.grid[2][1], ry);
    const int cz = vox_coord_try(pz[j], g.lo[2], g.rvs[2], g.qeps[2], g.grid[2][2], rz);
    redo |= (rx || ry || rz) ? 1u << j : 0u;
    cell[j] = (fb * g.grid[0][1] + (cy >> g.shift[0][1])) * g.grid[0][0] + (cx >> g.shift[0][0]) + (cz >> 8);
  }
  if (redo) {
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (redo >> j & 1u) {
        const int cx = vox_coord(px[j], g.lo[0], g.vs[2][0], g.grid[2][0]);
        const int cy = vox_coord(py[j], g.lo[1], g.vs[2][1], g.grid[2][1]);
        cell[j] = (fb * g.grid[0][1] + (cy >> g.shift[0][1])) * g.grid[0][0] + (cx >> g.shift[0][0]);
      }
  }
  uint32_t seen[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) seen[j] = (MODE & 4) ? *(volatile uint32_t*)(bitmap + (cell[j] >> 5)) : __ldcg(bitmap + (cell[j] >> 5));
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const uint32_t bit = 1u << (cell[j] & 31);
    if (!(seen[j] & bit)) atomicOr(bitmap + (cell[j] >> 5), bit);
  }
}

}  // namespace

extern "C" int probe_run(int variant, const geomae_voxel_cfg* cfg, const float* pts, int64_t n, int stride,
                         const int32_t* frame_off, int n_frames, uint32_t* bitmap, float* sink, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  VoxGeom g;
  int rc = gm_make_geom(cfg, n_frames, &g);
  if (rc) return rc;
  if (!g.fast) return -100;
  const int blocks = gm_div_up(n, TILE);
  const size_t smem = (size_t)TILE * stride * sizeof(float);
  switch (variant) {
    case 0: p_stream<<<gm_div_up(n * stride / 4, 1024), 256, 0, stream>>>((const float4*)pts, n * stride / 4, sink); break;
    case 1: p_tile<0><<<blocks, TPB, smem, stream>>>(g, pts, n, stride, frame_off, bitmap, sink); break;
    case 2: p_tile<1><<<blocks, TPB, smem, stream>>>(g, pts, n, stride, frame_off, bitmap, sink); break;
    case 3: p_tile<1 | 2><<<blocks, TPB, smem, stream>>>(g, pts, n, stride, frame_off, bitmap, sink); break;
    case 4: p_tile<1 | 2 | 4><<<blocks, TPB, smem, stream>>>(g, pts, n, stride, frame_off, bitmap, sink); break;
    case 5: p_tile<1 | 2 | 4 | 8><<<blocks, TPB, smem, stream>>>(g, pts, n, stride, frame_off, bitmap, sink); break;
    case 6: p_tile<1 | 2 | 16 | 8><<<blocks, TPB, smem, stream>>>(g, pts, n, stride, frame_off, bitmap, sink); break;
    case 7: p_tile<1 | 2 | 32><<<blocks, TPB, smem, stream>>>(g, pts, n, stride, frame_off, bitmap, sink); break;
    case 8: p_reg5<1 | 4><<<gm_div_up(n / 4, 256), 256, 0, stream>>>(g, pts, n, frame_off, bitmap, sink); break;
    case 9: p_reg5<1><<<gm_div_up(n / 4, 256), 256, 0, stream>>>(g, pts, n, frame_off, bitmap, sink); break;
    case 10: p_reg5<0><<<gm_div_up(n / 4, 256), 256, 0, stream>>>(g, pts, n, frame_off, bitmap, sink); break;
    default: return -101;
  }
  return cudaGetLastError() == cudaSuccess ? 0 : -102;
}
