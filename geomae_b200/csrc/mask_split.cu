// Random visible / masked split of the pillars of each frame, entirely on the device
// (detectors/multi_sub_voxel_dynamic_voxelnet_ssl.py:287-304).
//
// The reference draws torch.randperm(L) per sample and keeps the first int(L * (1 - ratio)) entries.
// What the model consumes is the random SUBSET (attention inside a window does not depend on token
// order), so this kernel selects exactly k = int(L * keep_frac) pillars uniformly at random without a
// sort: every pillar gets a 32-bit hash of (seed, frame, index); a CTA per frame finds the k-th smallest
// hash with a 4-pass radix select (8 bits per pass, histogram in shared memory), and a final pass
// compacts the pillars below / above the threshold into ids_keep / ids_mask in ascending pillar order
// (ties on the threshold value are resolved by index, so the count is always exact).
// One launch replaces 4 x (arange, random keys, 4 radix-sort passes, fix-up) of torch.randperm.
#include "common.cuh"

namespace {

__device__ __forceinline__ uint32_t hash_u32(uint64_t seed, uint32_t frame, uint32_t i) {
  uint64_t z = seed + 0x9E3779B97F4A7C15ull * ((uint64_t)frame << 32 | (uint64_t)i + 1ull);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;      // splitmix64 finaliser
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z ^= z >> 31;
  return (uint32_t)(z >> 32);
}

__global__ void __launch_bounds__(1024) k_mask_split(const int32_t* __restrict__ frame_starts, int n_frames,
                                                      double keep_frac, uint64_t seed, int64_t* ids_keep,
                                                      int64_t* ids_mask) {
  __shared__ int hist[256];
  __shared__ int scan_buf[3][33];
  __shared__ uint32_t s_prefix;
  __shared__ int s_need;
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int base0 = frame_starts[0];
  const int s = frame_starts[b], L = frame_starts[b + 1] - s;
  int keep_off = 0;
  for (int f = 0; f < b; ++f) keep_off += (int)((double)(frame_starts[f + 1] - frame_starts[f]) * keep_frac);
  const int mask_off = (s - base0) - keep_off;
  const int k = (int)((double)L * keep_frac);        // int(L * (1 - ratio)) in float64, as the reference

  // ---- radix select: threshold T = k-th smallest hash (1-based), need = how many elements == T are kept
  uint32_t prefix = 0, pmask = 0;
  int need = k;
  if (k > 0) {
    for (int pass = 0; pass < 4; ++pass) {
      const int shift = 24 - 8 * pass;
      if (tid < 256) hist[tid] = 0;
      __syncthreads();
      for (int i = tid; i < L; i += 1024) {
        const uint32_t hsh = hash_u32(seed, b, i);
        if ((hsh & pmask) == prefix) atomicAdd(&hist[(hsh >> shift) & 255], 1);
      }
      __syncthreads();
      if (tid == 0) {
        int cum = 0, d = 0;
        for (; d < 256; ++d) {
          if (cum + hist[d] >= need) break;
          cum += hist[d];
        }
        s_prefix = prefix | ((uint32_t)d << shift);
        s_need = need - cum;
      }
      __syncthreads();
      prefix = s_prefix;
      need = s_need;
      pmask |= 255u << shift;
    }
  }
  const uint32_t T = prefix;

  // ---- compaction in pillar order
  int run_keep = 0, run_mask = 0, run_tie = 0;
  for (int base = 0; base < L; base += 1024) {
    const int i = base + tid;
    const bool valid = i < L;
    const uint32_t hsh = valid ? hash_u32(seed, b, i) : 0u;
    const bool lt = valid && k > 0 && hsh < T;
    const bool eq = valid && k > 0 && hsh == T;
    // exclusive ranks of three predicates via ballots + per-warp totals
    const uint32_t below = (1u << lane) - 1u;
    const uint32_t m_eq = __ballot_sync(0xffffffffu, eq);
    if (lane == 0) scan_buf[0][warp] = __popc(m_eq);
    __syncthreads();
    int tie_before = run_tie;
    for (int w = 0; w < warp; ++w) tie_before += scan_buf[0][w];
    int tie_total = 0;
    for (int w = 0; w < 32; ++w) tie_total += scan_buf[0][w];
    const int tie_rank = tie_before + __popc(m_eq & below);
    const bool keep = lt || (eq && tie_rank < need);
    const bool msk = valid && !keep;
    const uint32_t m_k = __ballot_sync(0xffffffffu, keep), m_m = __ballot_sync(0xffffffffu, msk);
    if (lane == 0) { scan_buf[1][warp] = __popc(m_k); scan_buf[2][warp] = __popc(m_m); }
    __syncthreads();
    int kb = run_keep, mb = run_mask, kt = 0, mt = 0;
    for (int w = 0; w < 32; ++w) {
      if (w < warp) { kb += scan_buf[1][w]; mb += scan_buf[2][w]; }
      kt += scan_buf[1][w]; mt += scan_buf[2][w];
    }
    if (keep) ids_keep[keep_off + kb + __popc(m_k & below)] = (int64_t)(s - base0) + i;
    if (msk) ids_mask[mask_off + mb + __popc(m_m & below)] = (int64_t)(s - base0) + i;
    run_keep += kt; run_mask += mt; run_tie += tie_total;
    __syncthreads();
  }
}

}  // namespace

extern "C" int geomae_mask_split(const int32_t* frame_starts, int32_t n_frames, double keep_frac, uint64_t seed,
                                 int64_t* ids_keep, int64_t* ids_mask, void* stream) {
  GM_REQUIRE(n_frames >= 0, "mask_split: negative frame count");
  if (n_frames == 0) return GEOMAE_OK;
  GM_REQUIRE(frame_starts, "mask_split: null frame_starts");   // ids_keep / ids_mask may be empty (null) lists
  GM_REQUIRE(keep_frac >= 0.0 && keep_frac <= 1.0, "mask_split: keep fraction %f not in [0,1]", keep_frac);
  k_mask_split<<<n_frames, 1024, 0, (cudaStream_t)stream>>>(frame_starts, n_frames, keep_frac, seed, ids_keep, ids_mask);
  GM_LAUNCH_CHECK();
  return GEOMAE_OK;
}
