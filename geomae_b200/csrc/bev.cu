// Pillar rows -> dense BEV canvas and back (the step between the SRA encoder and the SECOND
// convolutions of the fine-tune consumer).
//
// The reference fills one [C, ny*nx] canvas per sample with an indexed assignment of the transposed rows
// (backbones/sst_second_pretrained_v1.py:246-276).  Here one launch covers the batch: a block stages 32 rows x C
// channels through shared memory so that the row reads are coalesced along C and the canvas writes run along x
// (pillar rows arrive in cell order, so neighbouring rows are mostly neighbouring x of one canvas line).
#include "common.cuh"

namespace {

constexpr int BEV_ROWS = 32;

template <bool FWD>
__global__ void __launch_bounds__(256) k_bev(const float* __restrict__ src, const int32_t* __restrict__ coors, int64_t n,
                                             int channels, int ny, int nx, float* __restrict__ dst) {
  extern __shared__ float tile[];  // [BEV_ROWS][channels + 1]
  __shared__ int64_t s_base[BEV_ROWS];
  const int64_t row0 = (int64_t)blockIdx.x * BEV_ROWS;
  const int rows = (int)min((int64_t)BEV_ROWS, n - row0);
  const int ld = channels + 1;
  const int64_t plane = (int64_t)ny * nx;
  if (threadIdx.x < rows) {
    const int32_t* c = coors + (row0 + threadIdx.x) * 4;
    s_base[threadIdx.x] = (int64_t)c[0] * channels * plane + (int64_t)c[2] * nx + c[3];
  }
  if (FWD)
    for (int i = threadIdx.x; i < rows * channels; i += blockDim.x)
      tile[(i / channels) * ld + i % channels] = src[row0 * channels + i];
  __syncthreads();
  // thread -> (row = i % 32, channel = i / 32): a warp walks 32 neighbouring rows of one channel plane
  for (int i = threadIdx.x; i < BEV_ROWS * channels; i += blockDim.x) {
    const int r = i % BEV_ROWS, ch = i / BEV_ROWS;
    if (r >= rows) continue;
    const int64_t at = s_base[r] + (int64_t)ch * plane;
    if (FWD) dst[at] = tile[r * ld + ch];
    else tile[r * ld + ch] = src[at];
  }
  if (!FWD) {
    __syncthreads();
    for (int i = threadIdx.x; i < rows * channels; i += blockDim.x)
      dst[row0 * channels + i] = tile[(i / channels) * ld + i % channels];
  }
}

int check(const void* a, const int32_t* coors, const void* b, int64_t n, int channels, int ny, int nx, const char* who) {
  GM_REQUIRE(b && (n == 0 || (a && coors)), "%s: null argument", who);
  GM_REQUIRE(n >= 0 && channels > 0 && channels <= 1024 && ny > 0 && nx > 0, "%s: bad shape", who);
  return GEOMAE_OK;
}

}  // namespace

extern "C" int geomae_recover_bev(const float* feat, const int32_t* coors, int64_t n, int32_t channels,
                                  int32_t n_frames, int32_t ny, int32_t nx, float* canvas, void* stream_) {
  int rc = check(feat, coors, canvas, n, channels, ny, nx, "recover_bev");
  if (rc) return rc;
  GM_REQUIRE(n_frames > 0, "recover_bev: n_frames must be positive");
  cudaStream_t stream = (cudaStream_t)stream_;
  GM_CUDA(cudaMemsetAsync(canvas, 0, (size_t)n_frames * channels * ny * nx * sizeof(float), stream));
  if (n > 0)
    k_bev<true><<<gm_div_up(n, BEV_ROWS), 256, BEV_ROWS * (channels + 1) * sizeof(float), stream>>>(
        feat, coors, n, channels, ny, nx, canvas);
  GM_LAUNCH_CHECK();
  return GEOMAE_OK;
}

extern "C" int geomae_recover_bev_bwd(const float* d_canvas, const int32_t* coors, int64_t n, int32_t channels,
                                      int32_t ny, int32_t nx, float* d_feat, void* stream_) {
  int rc = check(d_canvas, coors, d_feat, n, channels, ny, nx, "recover_bev_bwd");
  if (rc) return rc;
  if (n > 0)
    k_bev<false><<<gm_div_up(n, BEV_ROWS), 256, BEV_ROWS * (channels + 1) * sizeof(float), (cudaStream_t)stream_>>>(
        d_canvas, coors, n, channels, ny, nx, d_feat);
  GM_LAUNCH_CHECK();
  return GEOMAE_OK;
}
