// GPU-side data step in front of the voxel scatter: the per-sample augmentations of the
// pretraining pipeline and the range filter, fused with a stable compaction over the whole batch.
//
// The reference runs these per sample on the data-loader workers' CPU cores (configs/mae_sst/…6x_1e-5.py:181-193):
//   GlobalRotScaleTrans  points[:, :3] @ [[c, s, 0], [-s, c, 0], [0, 0, 1]], then *= scale, translation std 0
//                        (datasets/pipelines/transforms_3d.py:670-718, core/points/base_points.py:139-179,263-269)
//   RandomFlip3D         horizontal: y = -y, vertical: x = -x          (transforms_3d.py:95-123, lidar_points.py:28-33)
//   PointsRangeFilter    strict  min < p < max  on x, y, z             (transforms_3d.py:849-883, base_points.py:207-229)
//   PointShuffle         a permutation — everything downstream is order-free, not reproduced.
// Here: one batch-wide pass counts the survivors per 1024-point tile, a single-CTA scan turns the counts into output
// offsets, a second pass writes the transformed survivors in input order (what boolean-mask indexing gives) and the
// new per-frame offsets.  Each product is rounded separately (no contraction), like a non-FMA fp32 matmul.
#include "common.cuh"
#include "voxel_geom.cuh"

namespace {

constexpr int ATPB = 256;
constexpr int APPT = 4;
constexpr int ATILE = ATPB * APPT;

struct AugArgs {
  const float* pts;
  int64_t n;
  int stride;
  const int32_t* off;
  int n_frames;
  const float* params;   // [n_frames, 4]: cos, sin, scale, flip bits (1 horizontal, 2 vertical) as a float
  float lo[3], hi[3];
  const double* seg;     // sweep-merge mode (non-null): [n_frames, 16] doubles per SEGMENT (= one sweep file):
                         // R row-major (9), t (3), time lag, close radius (< 0: keep everything), 2 unused
};

// Sweep merge (LoadPointsFromMultiSweeps, datasets/pipelines/loading.py:160-181,218-226): a raw sweep point inside
// the |x| < r, |y| < r box is dropped; the others move into the key frame: the float32 row times the float64
// sensor->lidar rotation (numpy promotes to float64), rounded to float32 on assignment, then the float64 translation
// added in place (computed in float64, rounded to float32 again); channel 4 becomes the sweep's time lag.
__device__ __forceinline__ bool merge_point(const AugArgs& a, int64_t idx, int s, float& x, float& y, float& z) {
  const float* rec = a.pts + idx * a.stride;
  const float px = __ldg(rec), py = __ldg(rec + 1), pz = __ldg(rec + 2);
  const double* q = a.seg + (int64_t)s * 16;
  const double r = __ldg(q + 13);
  if (r >= 0.0 && fabsf(px) < (float)r && fabsf(py) < (float)r) return false;
  const double dx = px, dy = py, dz = pz;
  // row . R^T: x' = p . R[0], in the operand order of a row-major dot product
  const float rx = (float)__dadd_rn(__dadd_rn(__dmul_rn(dx, __ldg(q + 0)), __dmul_rn(dy, __ldg(q + 1))), __dmul_rn(dz, __ldg(q + 2)));
  const float ry = (float)__dadd_rn(__dadd_rn(__dmul_rn(dx, __ldg(q + 3)), __dmul_rn(dy, __ldg(q + 4))), __dmul_rn(dz, __ldg(q + 5)));
  const float rz = (float)__dadd_rn(__dadd_rn(__dmul_rn(dx, __ldg(q + 6)), __dmul_rn(dy, __ldg(q + 7))), __dmul_rn(dz, __ldg(q + 8)));
  x = (float)__dadd_rn((double)rx, __ldg(q + 9));
  y = (float)__dadd_rn((double)ry, __ldg(q + 10));
  z = (float)__dadd_rn((double)rz, __ldg(q + 11));
  return true;
}

// transformed xyz of point idx and whether it survives the range filter
__device__ __forceinline__ bool aug_point(const AugArgs& a, int64_t idx, int b, float& x, float& y, float& z) {
  const float* rec = a.pts + idx * a.stride;
  const float px = __ldg(rec), py = __ldg(rec + 1), pz = __ldg(rec + 2);
  const float4 q = __ldg(reinterpret_cast<const float4*>(a.params) + b);
  x = __fmul_rn(__fsub_rn(__fmul_rn(px, q.x), __fmul_rn(py, q.y)), q.z);   // (x c - y s) * scale
  y = __fmul_rn(__fadd_rn(__fmul_rn(px, q.y), __fmul_rn(py, q.x)), q.z);   // (x s + y c) * scale
  z = __fmul_rn(pz, q.z);
  const int flags = (int)q.w;
  if (flags & 1) y = -y;
  if (flags & 2) x = -x;
  return x > a.lo[0] && y > a.lo[1] && z > a.lo[2] && x < a.hi[0] && y < a.hi[1] && z < a.hi[2];
}

// survivors of the thread's APPT points (tile-strided: chunk j of the tile is 256 consecutive points) as a bit mask
__device__ __forceinline__ unsigned aug_flags(const AugArgs& a, int64_t p0, int nvalid, int b0, float (*xyz)[3]) {
  unsigned keep = 0;
#pragma unroll
  for (int j = 0; j < APPT; ++j) {
    const int l = threadIdx.x + j * ATPB;
    if (l < nvalid && p0 + l < (int64_t)__ldg(a.off + a.n_frames)) {   // rows past the last frame's end are not points
      int b = b0;
      while (b + 1 < a.n_frames && (int64_t)__ldg(a.off + b + 1) <= p0 + l) ++b;
      const bool ok = a.seg ? merge_point(a, p0 + l, b, xyz[j][0], xyz[j][1], xyz[j][2])
                            : aug_point(a, p0 + l, b, xyz[j][0], xyz[j][1], xyz[j][2]);
      if (ok) keep |= 1u << j;
    }
  }
  return keep;
}

__global__ void __launch_bounds__(ATPB) k_aug_count(AugArgs a, int32_t* tile_sums) {
  __shared__ int smem[ATPB / 32];
  const int64_t p0 = (int64_t)blockIdx.x * ATILE;
  const int nvalid = (int)min((int64_t)ATILE, a.n - p0);
  float xyz[APPT][3];
  const unsigned keep = aug_flags(a, p0, nvalid, frame_of(a.off, a.n_frames, p0), xyz);
  int c = gm_warp_sum_i(__popc(keep));
  if ((threadIdx.x & 31) == 0) smem[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    for (int i = 0; i < ATPB / 32; ++i) t += smem[i];
    tile_sums[blockIdx.x] = t;
  }
}

// in-place exclusive scan of tile_sums[0..n_tiles), total appended at [n_tiles]; one CTA, chunks of 1024 with a carry
__global__ void __launch_bounds__(1024) k_aug_scan(int32_t* tile_sums, int n_tiles) {
  __shared__ int smem[40];
  __shared__ int carry_s;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  for (int base = 0; base < n_tiles; base += 1024) {
    const int i = base + threadIdx.x;
    const int v = i < n_tiles ? tile_sums[i] : 0;
    int total;
    const int ex = gm_block_excl_scan(v, &total, smem);
    const int carry = carry_s;
    if (i < n_tiles) tile_sums[i] = carry + ex;
    __syncthreads();
    if (threadIdx.x == 0) carry_s = carry + total;
    __syncthreads();
  }
  if (threadIdx.x == 0) tile_sums[n_tiles] = carry_s;
}

__global__ void __launch_bounds__(ATPB) k_aug_write(AugArgs a, const int32_t* __restrict__ tile_base, float* out,
                                                    int32_t* out_off) {
  __shared__ int warp_cnt[APPT][ATPB / 32];
  __shared__ int rank_s[ATILE];   // exclusive rank of every point of the tile among the tile's survivors
  const int64_t p0 = (int64_t)blockIdx.x * ATILE;
  const int nvalid = (int)min((int64_t)ATILE, a.n - p0);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float xyz[APPT][3];
  const unsigned keep = aug_flags(a, p0, nvalid, frame_of(a.off, a.n_frames, p0), xyz);
  unsigned ballot[APPT];
#pragma unroll
  for (int j = 0; j < APPT; ++j) {
    ballot[j] = __ballot_sync(0xffffffffu, (keep >> j) & 1u);
    if (lane == 0) warp_cnt[j][warp] = __popc(ballot[j]);
  }
  __syncthreads();
  // point order inside the tile: chunk j (256 consecutive points), then warp, then lane
  int before[APPT];
  {
    int run = 0;
#pragma unroll
    for (int j = 0; j < APPT; ++j) {
      int mine = 0;
      for (int w = 0; w < ATPB / 32; ++w) {
        if (w == warp) mine = run;
        run += warp_cnt[j][w];
      }
      before[j] = mine;
    }
  }
  const int base = tile_base[blockIdx.x];
#pragma unroll
  for (int j = 0; j < APPT; ++j) {
    const int l = threadIdx.x + j * ATPB;
    const int r = before[j] + __popc(ballot[j] & ((1u << lane) - 1u));
    rank_s[l] = r;
    if ((keep >> j) & 1u) {
      float* o = out + ((int64_t)base + r) * a.stride;
      const float* rec = a.pts + (p0 + l) * a.stride;
      o[0] = xyz[j][0];
      o[1] = xyz[j][1];
      o[2] = xyz[j][2];
      for (int c = 3; c < a.stride; ++c) o[c] = __ldg(rec + c);
      if (a.seg && a.stride > 4) {       // the time channel of the sweep this point came from
        int sg = frame_of(a.off, a.n_frames, p0 + l);
        o[4] = (float)__ldg(a.seg + (int64_t)sg * 16 + 12);
      }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    // frames that start inside this tile (several when frames are empty): first output row = rank of their first point
    int lo = 0, hi = a.n_frames + 1;   // first b in [0, n_frames] with off[b] >= p0
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if ((int64_t)__ldg(a.off + mid) < p0) lo = mid + 1; else hi = mid;
    }
    int b = lo;
    for (; b <= a.n_frames && (int64_t)__ldg(a.off + b) < p0 + nvalid; ++b)
      out_off[b] = base + rank_s[(int)((int64_t)__ldg(a.off + b) - p0)];
    if (p0 + nvalid == a.n) {           // the last tile also closes the batch (and any trailing empty frames)
      const int total = tile_base[gridDim.x];
      for (; b <= a.n_frames; ++b) out_off[b] = total;
    }
  }
}

}  // namespace

extern "C" int geomae_augment_filter(const float* points, int64_t n_points, int32_t stride,
                                     const int32_t* frame_offsets, int32_t n_frames, const float* frame_params,
                                     const float range_min[3], const float range_max[3], float* out_points,
                                     int32_t* out_frame_offsets, int32_t* scan_tmp, int64_t scan_tmp_len, void* stream_) {
  GM_REQUIRE(n_points >= 0 && n_frames >= 1 && stride >= 3 && stride <= 16, "augment_filter: bad sizes (stride %d)", stride);
  GM_REQUIRE(frame_offsets && frame_params && range_min && range_max && out_frame_offsets && scan_tmp,
             "augment_filter: null argument");
  GM_REQUIRE(n_points < ((int64_t)1 << 31), "augment_filter: more than 2^31 points");
  cudaStream_t stream = (cudaStream_t)stream_;
  if (n_points == 0) {
    GM_CUDA(cudaMemsetAsync(out_frame_offsets, 0, sizeof(int32_t) * (size_t)(n_frames + 1), stream));
    return GEOMAE_OK;
  }
  GM_REQUIRE(points && out_points, "augment_filter: null point buffer");
  const int n_tiles = gm_div_up(n_points, ATILE);
  GM_REQUIRE(scan_tmp_len >= (int64_t)n_tiles + 1, "augment_filter: scan_tmp holds %lld ints, %d needed",
             (long long)scan_tmp_len, n_tiles + 1);
  AugArgs a;
  a.pts = points; a.n = n_points; a.stride = stride; a.off = frame_offsets; a.n_frames = n_frames; a.params = frame_params;
  a.seg = nullptr;
  for (int i = 0; i < 3; ++i) { a.lo[i] = range_min[i]; a.hi[i] = range_max[i]; }
  k_aug_count<<<n_tiles, ATPB, 0, stream>>>(a, scan_tmp);
  k_aug_scan<<<1, 1024, 0, stream>>>(scan_tmp, n_tiles);
  k_aug_write<<<n_tiles, ATPB, 0, stream>>>(a, scan_tmp, out_points, out_frame_offsets);
  GM_LAUNCH_CHECK();
  return GEOMAE_OK;
}

extern "C" int geomae_sweep_merge(const float* points, int64_t n_points, int32_t stride, const int32_t* seg_offsets,
                                  int32_t n_segments, const double* seg_params, float* out_points,
                                  int32_t* out_seg_offsets, int32_t* scan_tmp, int64_t scan_tmp_len, void* stream_) {
  GM_REQUIRE(n_points >= 0 && n_segments >= 1 && stride >= 3 && stride <= 16, "sweep_merge: bad sizes (stride %d)", stride);
  GM_REQUIRE(seg_offsets && seg_params && out_seg_offsets && scan_tmp, "sweep_merge: null argument");
  GM_REQUIRE(n_points < ((int64_t)1 << 31), "sweep_merge: more than 2^31 points");
  cudaStream_t stream = (cudaStream_t)stream_;
  if (n_points == 0) {
    GM_CUDA(cudaMemsetAsync(out_seg_offsets, 0, sizeof(int32_t) * (size_t)(n_segments + 1), stream));
    return GEOMAE_OK;
  }
  GM_REQUIRE(points && out_points, "sweep_merge: null point buffer");
  const int n_tiles = gm_div_up(n_points, ATILE);
  GM_REQUIRE(scan_tmp_len >= (int64_t)n_tiles + 1, "sweep_merge: scan_tmp holds %lld ints, %d needed",
             (long long)scan_tmp_len, n_tiles + 1);
  AugArgs a{};
  a.pts = points; a.n = n_points; a.stride = stride; a.off = seg_offsets; a.n_frames = n_segments; a.seg = seg_params;
  k_aug_count<<<n_tiles, ATPB, 0, stream>>>(a, scan_tmp);
  k_aug_scan<<<1, 1024, 0, stream>>>(scan_tmp, n_tiles);
  k_aug_write<<<n_tiles, ATPB, 0, stream>>>(a, scan_tmp, out_points, out_seg_offsets);
  GM_LAUNCH_CHECK();
  return GEOMAE_OK;
}
