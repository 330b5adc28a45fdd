// SST window partition + bucketing in CSR form.
//
// The reference pads every window to a bucket size (56/144), building the padded index with
// sorts, bincounts, uniques and ~6 host syncs per bucket per shift.  Here a window is just a CSR
// row: one warp owns one candidate window (b, wx, wy), probes its win_x*win_y BEV cells in the
// pillar-occupancy bitmap, and compacts the occupied ones with ballot/popcount prefix sums.
// Tokens inside a window come out in cell order, windows in ascending batch_win_inds order (the
// order the reference's sorted-unique "continuous" window index has).  No sort, no atomics, no
// host sync.  Both shifts are processed by the same launches (blockIdx.y = shift).
#include "common.cuh"
#include "voxel_geom.cuh"

namespace {

struct WinGeom {
  int win_x, win_y, n_shifts;
  int off_x[2], off_y[2];  // what is ADDED to a pillar coordinate before the division
  int nwx, nwy;            // candidate windows per frame along x / y
  int n_cand;              // n_frames * nwx * nwy
};

__device__ __forceinline__ int probe(const VoxGeom& g, const WinGeom& w, const uint32_t* __restrict__ bitmap,
                                     const int32_t* __restrict__ word_rank, const int32_t* __restrict__ tok_of_pillar,
                                     int shift, int cand, int c) {
  const int per_frame = w.nwx * w.nwy;
  const int b = cand / per_frame, r = cand % per_frame;
  const int wx = r / w.nwy, wy = r % w.nwy;
  const int cx = c / w.win_y, cy = c % w.win_y;
  const int x = wx * w.win_x + cx - w.off_x[shift], y = wy * w.win_y + cy - w.off_y[shift];
  if (x < 0 || y < 0 || x >= g.grid[0][0] || y >= g.grid[0][1]) return -1;
  const int pid = cell_rank(bitmap, word_rank, top_cell(g, b, y, x));
  return pid < 0 ? -1 : __ldg(tok_of_pillar + pid);
}

// one warp per candidate window
__global__ void __launch_bounds__(256) k_win_count(VoxGeom g, WinGeom w, const uint32_t* __restrict__ bitmap,
                                                   const int32_t* __restrict__ word_rank,
                                                   const int32_t* __restrict__ tok_of_pillar, int32_t* cand_count) {
  const int lane = threadIdx.x & 31;
  const int cand = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int shift = blockIdx.y;
  if (cand >= w.n_cand) return;
  const int cells = w.win_x * w.win_y;
  int cnt = 0;
  for (int c0 = 0; c0 < cells; c0 += 32) {
    const int c = c0 + lane;
    const bool ok = c < cells && probe(g, w, bitmap, word_rank, tok_of_pillar, shift, cand, c) >= 0;
    cnt += __popc(__ballot_sync(0xffffffffu, ok));
  }
  if (lane == 0) cand_count[shift * w.n_cand + cand] = cnt;
}

// one block per shift: exclusive scan of token counts and of the non-empty flags
__global__ void __launch_bounds__(1024) k_win_scan(WinGeom w, const int32_t* __restrict__ cand_count,
                                                   int32_t* cand_tok_off, int32_t* cand_win_idx, int32_t* win_ptr,
                                                   int32_t* win_id, int32_t* n_windows, int64_t ptr_stride) {
  __shared__ int smem[40];
  const int shift = blockIdx.x;
  const int32_t* cnt = cand_count + shift * w.n_cand;
  int32_t* toff = cand_tok_off + shift * w.n_cand;
  int32_t* widx = cand_win_idx + shift * w.n_cand;
  int32_t* wp = win_ptr + shift * ptr_stride;
  int32_t* wi = win_id + shift * ptr_stride;
  int run_tok = 0, run_win = 0;
  for (int base = 0; base < w.n_cand; base += blockDim.x) {
    const int i = base + threadIdx.x;
    const int c = i < w.n_cand ? cnt[i] : 0;
    int tot_t, tot_w;
    const int et = gm_block_excl_scan(c, &tot_t, smem);
    const int ew = gm_block_excl_scan(c > 0 ? 1 : 0, &tot_w, smem);
    if (i < w.n_cand) {
      toff[i] = run_tok + et;
      widx[i] = run_win + ew;
      if (c > 0) {
        wp[run_win + ew] = run_tok + et;
        wi[run_win + ew] = i;
      }
    }
    run_tok += tot_t;
    run_win += tot_w;
  }
  if (threadIdx.x == 0) {
    wp[run_win] = run_tok;
    n_windows[shift] = run_win;
  }
}

__global__ void __launch_bounds__(256) k_win_fill(VoxGeom g, WinGeom w, const uint32_t* __restrict__ bitmap,
                                                  const int32_t* __restrict__ word_rank,
                                                  const int32_t* __restrict__ tok_of_pillar,
                                                  const int32_t* __restrict__ cand_count,
                                                  const int32_t* __restrict__ cand_tok_off,
                                                  const int32_t* __restrict__ cand_win_idx, int64_t n_tokens,
                                                  int32_t* win_tok, int32_t* tok_cell, int32_t* tok_win,
                                                  int32_t* tok_pos) {
  const int lane = threadIdx.x & 31;
  const int cand = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int shift = blockIdx.y;
  if (cand >= w.n_cand) return;
  if (cand_count[shift * w.n_cand + cand] == 0) return;
  const int cells = w.win_x * w.win_y;
  int pos = cand_tok_off[shift * w.n_cand + cand];
  const int widx = cand_win_idx[shift * w.n_cand + cand];
  int32_t* o_tok = win_tok + shift * n_tokens;
  int32_t* o_cell = tok_cell + shift * n_tokens;
  int32_t* o_win = tok_win + shift * n_tokens;
  int32_t* o_pos = tok_pos + shift * n_tokens;
  for (int c0 = 0; c0 < cells; c0 += 32) {
    const int c = c0 + lane;
    const int tok = c < cells ? probe(g, w, bitmap, word_rank, tok_of_pillar, shift, cand, c) : -1;
    const uint32_t m = __ballot_sync(0xffffffffu, tok >= 0);
    if (tok >= 0) {
      const int p = pos + __popc(m & ((1u << lane) - 1u));
      o_tok[p] = tok;
      o_cell[tok] = c;
      o_win[tok] = widx;
      o_pos[tok] = p;
    }
    pos += __popc(m);
  }
}

// Region batching with voxel drop (middle_encoders/sst_input_layer.py:211-275).  One warp per candidate window of
// ONE shift: the window's live tokens are counted (n), the bucket with lower < n <= upper gives the level and the
// token budget, and when n exceeds the budget the `budget` tokens with the smallest key survive.  key = token index
// (seed 0: what the reference keeps with shuffle_voxels=False and a stable sort) or a hash of (seed, token)
// (shuffle_voxels=True: a uniformly random subset, which is all the reference's randperm + sort provides).
// A window that matches no bucket keeps nothing and reports level -1 (:219-228 leave target 0 / level -1).
struct DropLevels {
  int n;
  int max_tokens[8], lower[8], upper[8];
};
constexpr int DROP_MAX_CELLS = 1024;

__device__ __forceinline__ uint32_t drop_key(uint64_t seed, int tok) {
  if (seed == 0) return (uint32_t)tok;
  uint64_t z = seed + 0x9E3779B97F4A7C15ull * (uint64_t)(tok + 1);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return (uint32_t)((z ^ (z >> 31)) >> 32);
}

__global__ void __launch_bounds__(128) k_win_drop(VoxGeom g, WinGeom w, DropLevels lv, int shift, uint64_t seed,
                                                  const uint32_t* __restrict__ bitmap,
                                                  const int32_t* __restrict__ word_rank,
                                                  const int32_t* __restrict__ tok_of_pillar, uint8_t* alive,
                                                  int32_t* level /* this shift */) {
  __shared__ int s_tok[4][DROP_MAX_CELLS];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int cand = blockIdx.x * 4 + wib;
  if (cand >= w.n_cand) return;
  const int cells = w.win_x * w.win_y;
  int n = 0;
  for (int c0 = 0; c0 < cells; c0 += 32) {
    const int c = c0 + lane;
    int tok = c < cells ? probe(g, w, bitmap, word_rank, tok_of_pillar, shift, cand, c) : -1;
    if (tok >= 0 && !alive[tok]) tok = -1;
    const uint32_t m = __ballot_sync(0xffffffffu, tok >= 0);
    if (tok >= 0) s_tok[wib][n + __popc(m & ((1u << lane) - 1u))] = tok;
    n += __popc(m);
  }
  if (n == 0) return;
  __syncwarp();
  int lvl = -1, budget = 0;
  for (int l = 0; l < lv.n; ++l)   // later buckets override earlier ones, as the reference's loop of masked writes does
    if (n > lv.lower[l] && n <= lv.upper[l]) { lvl = l; budget = lv.max_tokens[l]; }
  for (int i = lane; i < n; i += 32) {
    const int tok = s_tok[wib][i];
    bool keep = budget >= n;
    if (!keep && budget > 0) {
      const uint32_t key = drop_key(seed, tok);
      int rank = 0;
      for (int j = 0; j < n; ++j) {
        const int tj = s_tok[wib][j];
        const uint32_t kj = drop_key(seed, tj);
        rank += (kj < key || (kj == key && tj < tok)) ? 1 : 0;
      }
      keep = rank < budget;
    }
    level[tok] = lvl;
    if (!keep) alive[tok] = 0;
  }
}

__global__ void k_token_map(const int64_t* __restrict__ rows, int64_t n, int32_t* tok_of_pillar) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) tok_of_pillar[rows[i]] = (int32_t)i;
}

// Sinusoidal in-window position table, row = cx*win_y + cy (backbones/…top_only.py:361-394):
// inv_freq_i = T^(2*(i//2)/half); e = (c - win/2)/inv_freq; even i -> sin, odd i -> cos; x block then y block.
__global__ void k_pos_table(int win_x, int win_y, int d_model, float temperature, float* table) {
  const int half = d_model / 2;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= win_x * win_y * d_model) return;
  const int row = i / d_model, col = i % d_model;
  const int cx = row / win_y, cy = row % win_y;
  const int j = col < half ? col : col - half;
  const float coord = col < half ? (float)cx - win_x * 0.5f : (float)cy - win_y * 0.5f;
  const float expo = __fdiv_rn((float)(2 * (j / 2)), (float)half);
  const float inv_freq = powf(temperature, expo);
  const float e = __fdiv_rn(coord, inv_freq);
  table[i] = (j & 1) ? cosf(e) : sinf(e);
}

int make_win_geom(const VoxGeom& g, const geomae_window_cfg* wc, int n_frames, WinGeom* w) {
  GM_REQUIRE(wc->win_x > 0 && wc->win_y > 0 && wc->n_shifts >= 1 && wc->n_shifts <= 2,
             "window cfg: need 1..2 shifts and a positive window shape");
  w->win_x = wc->win_x;
  w->win_y = wc->win_y;
  w->n_shifts = wc->n_shifts;
  for (int s = 0; s < wc->n_shifts; ++s) {
    // backbones/…top_only.py:646-648: shifted = coord + (win - shift if shift > 0 else 0)
    w->off_x[s] = wc->shift_x[s] > 0 ? wc->win_x - wc->shift_x[s] : 0;
    w->off_y[s] = wc->shift_y[s] > 0 ? wc->win_y - wc->shift_y[s] : 0;
  }
  w->nwx = (g.grid[0][0] + wc->win_x - 1) / wc->win_x + 1;  // :640-641, "+1 to meet the needs of shift"
  w->nwy = (g.grid[0][1] + wc->win_y - 1) / wc->win_y + 1;
  w->n_cand = n_frames * w->nwx * w->nwy;
  return GEOMAE_OK;
}

}  // namespace

extern "C" int geomae_window_candidates(const geomae_voxel_cfg* cfg, const geomae_window_cfg* wcfg, int32_t n_frames,
                                        int32_t* n_cand, int32_t* nwx, int32_t* nwy) {
  GM_REQUIRE(cfg && wcfg && n_cand, "window_candidates: null argument");
  VoxGeom g;
  int rc = gm_make_geom(cfg, n_frames, &g);
  if (rc) return rc;
  WinGeom w;
  rc = make_win_geom(g, wcfg, n_frames, &w);
  if (rc) return rc;
  *n_cand = w.n_cand;
  if (nwx) *nwx = w.nwx;
  if (nwy) *nwy = w.nwy;
  return GEOMAE_OK;
}

extern "C" int geomae_token_map(const int64_t* rows, int64_t n_tokens, int32_t* tok_of_pillar, int64_t n_pillars,
                                void* stream) {
  GM_REQUIRE(tok_of_pillar && (rows || n_tokens == 0), "token_map: null argument");
  GM_CUDA(cudaMemsetAsync(tok_of_pillar, 0xff, (size_t)n_pillars * 4, (cudaStream_t)stream));
  if (n_tokens > 0) k_token_map<<<gm_div_up(n_tokens, 256), 256, 0, (cudaStream_t)stream>>>(rows, n_tokens, tok_of_pillar);
  GM_LAUNCH_CHECK();
  return GEOMAE_OK;
}

extern "C" int geomae_window_csr(const geomae_voxel_cfg* cfg, const geomae_window_cfg* wcfg,
                                 const geomae_scatter_io* io, const int32_t* tok_of_pillar, int64_t n_tokens,
                                 const geomae_window_io* out, void* stream_) {
  GM_REQUIRE(cfg && wcfg && io && tok_of_pillar && out, "window_csr: null argument");
  GM_REQUIRE(out->cand_count && out->cand_tok_off && out->cand_win_idx && out->win_ptr && out->win_id &&
                 out->n_windows && out->win_tok && out->tok_cell && out->tok_win && out->tok_pos,
             "window_csr: null output buffer");
  cudaStream_t stream = (cudaStream_t)stream_;
  VoxGeom g;
  int rc = gm_make_geom(cfg, io->n_frames, &g);
  if (rc) return rc;
  WinGeom w;
  rc = make_win_geom(g, wcfg, io->n_frames, &w);
  if (rc) return rc;
  GM_REQUIRE(out->ptr_stride >= w.n_cand + 1, "window_csr: ptr_stride %lld < n_cand+1 = %d",
             (long long)out->ptr_stride, w.n_cand + 1);
  const dim3 grid(gm_div_up((int64_t)w.n_cand * 32, 256), w.n_shifts);
  k_win_count<<<grid, 256, 0, stream>>>(g, w, io->bitmap, io->word_rank, tok_of_pillar, out->cand_count);
  k_win_scan<<<w.n_shifts, 1024, 0, stream>>>(w, out->cand_count, out->cand_tok_off, out->cand_win_idx, out->win_ptr,
                                              out->win_id, out->n_windows, out->ptr_stride);
  if (n_tokens > 0)
    k_win_fill<<<grid, 256, 0, stream>>>(g, w, io->bitmap, io->word_rank, tok_of_pillar, out->cand_count,
                                         out->cand_tok_off, out->cand_win_idx, n_tokens, out->win_tok, out->tok_cell,
                                         out->tok_win, out->tok_pos);
  GM_LAUNCH_CHECK();
  return GEOMAE_OK;
}

extern "C" int geomae_window_drop(const geomae_voxel_cfg* cfg, const geomae_window_cfg* wcfg,
                                  const geomae_scatter_io* io, const int32_t* tok_of_pillar, int64_t n_tokens,
                                  int32_t n_levels, const int32_t* max_tokens, const int32_t* lower,
                                  const int32_t* upper, uint64_t seed, uint8_t* keep, int32_t* level, void* stream_) {
  GM_REQUIRE(cfg && wcfg && io && tok_of_pillar && max_tokens && lower && upper && keep && level,
             "window_drop: null argument");
  GM_REQUIRE(n_levels >= 1 && n_levels <= 8, "window_drop: 1..8 drop levels, got %d", n_levels);
  cudaStream_t stream = (cudaStream_t)stream_;
  VoxGeom g;
  int rc = gm_make_geom(cfg, io->n_frames, &g);
  if (rc) return rc;
  WinGeom w;
  rc = make_win_geom(g, wcfg, io->n_frames, &w);
  if (rc) return rc;
  GM_REQUIRE(w.win_x * w.win_y <= DROP_MAX_CELLS, "window_drop: window of %d cells > %d", w.win_x * w.win_y,
             DROP_MAX_CELLS);
  DropLevels lv;
  lv.n = n_levels;
  for (int l = 0; l < n_levels; ++l) {
    lv.max_tokens[l] = max_tokens[l];
    lv.lower[l] = lower[l];
    lv.upper[l] = upper[l];
  }
  if (n_tokens == 0) return GEOMAE_OK;
  GM_CUDA(cudaMemsetAsync(keep, 1, (size_t)n_tokens, stream));
  // shift 1 buckets the survivors of shift 0 (sst_input_layer.py:252-262); a token leaves if either shift drops it
  for (int s = 0; s < w.n_shifts; ++s)
    k_win_drop<<<gm_div_up(w.n_cand, 4), 128, 0, stream>>>(g, w, lv, s, seed, io->bitmap, io->word_rank,
                                                           tok_of_pillar, keep, level + (int64_t)s * n_tokens);
  GM_LAUNCH_CHECK();
  return GEOMAE_OK;
}

extern "C" int geomae_pos_table(int32_t win_x, int32_t win_y, int32_t d_model, float temperature, float* table,
                                void* stream) {
  GM_REQUIRE(table && win_x > 0 && win_y > 0 && d_model > 0 && d_model % 4 == 0, "pos_table: bad argument");
  const int n = win_x * win_y * d_model;
  k_pos_table<<<gm_div_up(n, 256), 256, 0, (cudaStream_t)stream>>>(win_x, win_y, d_model, temperature, table);
  GM_LAUNCH_CHECK();
  return GEOMAE_OK;
}
