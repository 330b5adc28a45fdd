/* geomae_b200 — C ABI of the B200-native GeoMAE pretraining hot path.
 *
 * Every entry point takes plain device pointers + sizes + a CUDA stream (passed
 * as void* so this header needs no CUDA include), never allocates, never
 * synchronises the device, and returns 0 on success or a negative GEOMAE_ERR_*
 * code; geomae_last_error() returns the message for the calling thread.  The
 * caller owns every buffer including workspaces.  Nothing throws across the ABI.
 *
 * Paths in "replaces:" comments are relative to the reference tree
 * (Tsinghua-MARS-Lab/GeoMAE); they name the interface each symbol stands in for.
 */
#ifndef GEOMAE_B200_H
#define GEOMAE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GEOMAE_OK 0
#define GEOMAE_ERR_INVALID (-1) /* bad argument (null pointer, size, alignment, config) */
#define GEOMAE_ERR_CUDA (-2)    /* a CUDA runtime call / kernel launch failed           */
#define GEOMAE_ERR_CAPACITY (-3)

const char* geomae_last_error(void);
int geomae_abi_version(void);

/* ---------------------------------------------------------------- voxelise */

/* Three-scale voxel geometry of one config (configs/mae_sst/…6x_1e-5.py:14-24). */
typedef struct geomae_voxel_cfg {
  float range_min[3];   /* x, y, z */
  float range_max[3];   /* x, y, z */
  float voxel_top[3];   /* pillar size x, y, z                    (voxel_size)          */
  float voxel_med[3];   /* sub-voxel size, middle scale           (sub_voxel_size_med)  */
  float voxel_low[3];   /* sub-voxel size, finest scale           (sub_voxel_size_low)  */
  int32_t ratio_med[3]; /* z, y, x sub-voxels per pillar, middle  (sub_voxel_ratio_med) */
  int32_t ratio_low[3]; /* z, y, x sub-voxels per pillar, finest  (sub_voxel_ratio_low) */
} geomae_voxel_cfg;

/* grid[x,y,z] = ceil((max-min)/size) in fp32.
 * replaces: mmdet3d/ops/voxel/src/voxelization_cuda.cu:375-377 */
int geomae_grid_size(const float range_min[3], const float range_max[3], const float voxel[3],
                     int32_t grid_xyz[3]);

/* coors[i] = (z,y,x) = clamp(floor((p - min)/size), 0, grid-1), fp32 IEEE divide.
 * points: [n, stride] fp32 row-major (stride >= 3); coors: [n,3] int32, caller-allocated.
 * replaces: voxel_layer.dynamic_voxelize  (mmdet3d/ops/voxel/src/voxelization.h:97-109,
 *           kernel voxelization_cuda.cu:22-63) */
int geomae_dynamic_voxelize(const float* points, int64_t n, int32_t stride,
                            const float voxel_xyz[3], const float range_min[3],
                            const float range_max[3], int32_t* coors, void* stream);

/* Workspace/outputs of the fused three-scale voxelise + scatter stage.  All device
 * pointers; "cap" = capacity in pillars (>= number of non-empty pillars; n_points is
 * always enough).  Arrays marked [opt] may be NULL. */
typedef struct geomae_scatter_io {
  /* inputs */
  const float* points;          /* [n_points, stride] fp32, frames concatenated            */
  const int32_t* frame_offsets; /* [n_frames+1] first point of each frame                  */
  int64_t n_points;
  int32_t stride;
  int32_t n_frames;
  int64_t cap;
  /* scratch (caller zeroing not required) */
  uint32_t* bitmap;             /* [ceil(n_frames*gy*gx/32)] pillar occupancy              */
  int32_t* word_rank;           /* [same] exclusive popcount prefix                        */
  int32_t* scan_tmp;            /* [3*16384] block sums                                     */
  /* outputs */
  int32_t* counts;              /* [4+n_frames+1] n_pillars, n_med, n_low, overflow flag,  */
                                /*   then the first pillar row of each frame (and V last)  */
  int32_t* pillar_coors;        /* [cap,4] (b,z,y,x), lexicographically sorted             */
  float* pillar_mean;           /* [cap,4] mean x,y,z and point count (as float)           */
  int32_t* point_pillar;        /* [n_points] pillar row of each point (unq_inv)           */
  uint32_t* med_mask;           /* [cap]   bit s set = middle sub-voxel slot s occupied     */
  uint32_t* low_mask;           /* [cap,4] 128 slot bits                                   */
  int32_t* med_ptr;             /* [cap+1] CSR offsets into med_mean (slot order)          */
  int32_t* low_ptr;             /* [cap+1]                                                 */
  float* med_mean;              /* [n_points,4] mean x,y,z,count per middle sub-voxel      */
  float* low_mean;              /* [n_points,4]                                            */
  int32_t* coors_top;           /* [opt] [n_points,4] (b,z,y,x) per point, pillar scale    */
  int32_t* coors_med;           /* [opt]                                                   */
  int32_t* coors_low;           /* [opt]                                                   */
} geomae_scatter_io;

/* One pass structure over the raw points: voxelise at three scales, build the sorted
 * pillar list without a sort (occupancy bitmap + popcount ranks), point->pillar map,
 * and per-pillar / per-sub-voxel centroids.  n_frames * grid_y * grid_x must stay below 2^31.
 * When the three scales are nested power-of-two multiples (every GeoMAE config) the voxel
 * coordinates of all scales derive from one low-scale coordinate per axis (bit-exact with the
 * reference's independent IEEE divides) and the middle-scale / pillar sums are formed
 * hierarchically from the low-scale sums; other geometries take the general path (independent
 * divides, one reduction per point and scale).
 * replaces: MultiSubVoxelDynamicVoxelNetSSL.voxelize / sub_voxelize_low / sub_voxelize_med
 *           (detectors/multi_sub_voxel_dynamic_voxelnet_ssl.py:307-377), scatter_v2(mode='avg')
 *           (ops/sst/sst_ops.py:8-39), get_centroid_per_voxel x3 (…_ssl.py:726-768) and the
 *           slot bookkeeping of get_multi_voxel_id_to_tensor_id_* (…_ssl.py:643-722). */
int geomae_voxel_scatter(const geomae_voxel_cfg* cfg, const geomae_scatter_io* io, void* stream);

/* ------------------------------------------------------- geometric targets */

/* Per pillar: 3x3 scatter matrix of the neighbourhood's middle-scale centroids about the
 * pillar centroid, symmetric eigen-solve, unit normal (z,y,x; sign: first non-zero
 * component positive), singular values (descending) and curvature (S+1e-9)/sum in f64.
 * replaces: spconv get_indice_pairs_implicit_gemm (call …_ssl.py:192-207) +
 *           cal_regular_voxel_nor_and_curv (…_ssl.py:575-610). */
int geomae_geom_targets(const geomae_voxel_cfg* cfg, const geomae_scatter_io* io, int64_t n_pillars,
                        float* normal /*[n,3]*/, double* curvature /*[n,3]*/,
                        float* cov6 /*[opt][n,6] zz,zy,zx,yy,yx,xx*/,
                        float* singular /*[opt][n,3]*/, int32_t* pair /*[opt][9,n]*/, void* stream);

/* Dense targets of the selected (masked) pillars, the layout the reference's loss consumes.
 * rows: [m] pillar rows.  Any output may be NULL.
 * replaces: normalize_centroid_sub_voxel (…_ssl.py:626-641) +
 *           get_multi_voxel_id_to_tensor_id_ori (…_ssl.py:673-722) +
 *           get_multi_voxel_id_to_tensor_id_for_curv (…_ssl.py:643-671, raw=1). */
int geomae_dense_targets(const geomae_voxel_cfg* cfg, const geomae_scatter_io* io, const int64_t* rows,
                         int64_t m, int32_t raw, float* low /*[m,slots_low,3]*/,
                         uint8_t* low_mask /*[m,slots_low]*/, float* med /*[m,slots_med,3]*/,
                         uint8_t* med_mask /*[m,slots_med]*/, float* top /*[m,3]*/, void* stream);

/* ------------------------------------------------------ window partition */

/* window_shape / shifts_list of the backbone (configs/mae_sst/…6x_1e-5.py:15,57). */
typedef struct geomae_window_cfg {
  int32_t win_x, win_y;
  int32_t n_shifts;       /* 1 or 2 */
  int32_t shift_x[2], shift_y[2];
} geomae_window_cfg;

/* CSR window layout of one token set, both shifts.  Arrays are [n_shifts, stride] with the
 * stride given next to each; all device pointers, caller-allocated.  n_cand comes from
 * geomae_window_candidates(). */
typedef struct geomae_window_io {
  int64_t ptr_stride;    /* >= n_cand + 1                                                      */
  int32_t* cand_count;   /* scratch [n_shifts, n_cand]                                          */
  int32_t* cand_tok_off; /* scratch [n_shifts, n_cand]                                          */
  int32_t* cand_win_idx; /* scratch [n_shifts, n_cand]                                          */
  int32_t* n_windows;    /* [n_shifts] number of non-empty windows                              */
  int32_t* win_ptr;      /* [n_shifts, ptr_stride] CSR offsets, first n_windows+1 valid         */
  int32_t* win_id;       /* [n_shifts, ptr_stride] batch_win_inds value of each window (sorted) */
  int32_t* win_tok;      /* [n_shifts, n_tokens] token index, grouped by window, cell order     */
  int32_t* tok_cell;     /* [n_shifts, n_tokens] in-window cell cx*win_y+cy of each token (position-table row) */
  int32_t* tok_win;      /* [n_shifts, n_tokens] window row of each token                       */
  int32_t* tok_pos;      /* [n_shifts, n_tokens] CSR position of each token                     */
} geomae_window_io;

/* Number of candidate windows n_frames*nwx*nwy, nw = ceil(grid/win)+1.
 * replaces: MultiMAESSTSPChoose.window_partition bookkeeping
 *           (backbones/multi_mae_sst_spearate_top_only.py:637-642). */
int geomae_window_candidates(const geomae_voxel_cfg* cfg, const geomae_window_cfg* wcfg, int32_t n_frames,
                             int32_t* n_cand, int32_t* nwx, int32_t* nwy);

/* Occupancy bitmap + ranks from explicit token rows coors [n,4] (b,z,y,x) (unique cells), and
 * tok_of_pillar[rank(cell_i)] = i.  Lets geomae_window_csr serve callers that only hold coordinates,
 * i.e. the reference signature backbone.forward(voxel_feat, coors, coors_mask, batch_size)
 * (…top_only.py:136-141) and SSTInputLayer.forward (middle_encoders/sst_input_layer.py:51-103).
 * bitmap/word_rank: [ceil(n_frames*gy*gx/32)], scan_tmp [3*16384], counts [4] (cleared here), tok_of_pillar [n]. */
int geomae_coors_bitmap(const geomae_voxel_cfg* cfg, const int32_t* coors, int64_t n, int32_t n_frames,
                        uint32_t* bitmap, int32_t* word_rank, int32_t* scan_tmp, int32_t* counts,
                        int32_t* tok_of_pillar, void* stream);

/* Duplicate-tolerant variant for point-level rows: rank_of_row[i] = rank of coors[i]'s cell among the distinct cells
 * (cells in (b, y, x) order, i.e. the row order of torch.unique(dim=0) for a grid with one z level), counts[0] =
 * number of distinct cells, first_row[r] = smallest i with rank r (optional; [n] ints).  A (b, z, y, x) grid with Z
 * levels is served by folding z into y (y' = z*Y + y, grid Z*Y rows).  Scratch buffers as for geomae_coors_bitmap.
 * replaces: the torch.unique(coors, return_inverse, dim=0) of scatter_v2 (ops/sst/sst_ops.py:15-17) — a library
 *           lexicographic sort — by one bitmap pass + a prefix sum. */
int geomae_coors_rank(const geomae_voxel_cfg* cfg, const int32_t* coors, int64_t n, int32_t n_frames, uint32_t* bitmap,
                      int32_t* word_rank, int32_t* scan_tmp, int32_t* counts, int32_t* rank_of_row,
                      int32_t* first_row, void* stream);

/* Random visible / masked split of each frame's pillars: frame f (pillars frame_starts[f] .. frame_starts[f+1])
 * keeps exactly k_f = (int)(L_f * keep_frac) pillars chosen uniformly at random (hash of (seed, f, index), k-th
 * smallest found by radix select — no sort); ids_keep [sum k_f] and ids_mask [sum (L_f - k_f)] receive the pillar
 * rows frame by frame, ascending inside a frame.  frame_starts: device int32 [n_frames + 1].
 * replaces: MultiSubVoxelDynamicVoxelNetSSL.get_vanilla_mask_index
 *           (detectors/multi_sub_voxel_dynamic_voxelnet_ssl.py:287-304; per-sample torch.randperm). */
int geomae_mask_split(const int32_t* frame_starts, int32_t n_frames, double keep_frac, uint64_t seed,
                      int64_t* ids_keep, int64_t* ids_mask, void* stream);

/* tok_of_pillar[rows[i]] = i, every other pillar -1 (rows = ids_keep, or [ids_keep; ids_mask]). */
int geomae_token_map(const int64_t* rows, int64_t n_tokens, int32_t* tok_of_pillar, int64_t n_pillars,
                     void* stream);

/* Window CSR for the token set described by tok_of_pillar, using the scatter stage's bitmap.
 * replaces: window_partition (…top_only.py:628-659), drop_single_shift / get_voxel_keep_inds
 *           (:519-541,562-626; no token is ever dropped when win_x*win_y <= the largest bucket),
 *           get_flat2win_inds / make_continuous_inds / get_inner_win_inds (:413-507,661-681),
 *           and the flat2window / window2flat padding round trips (ops/sst/sst_ops.py:98-135,225-251). */
int geomae_window_csr(const geomae_voxel_cfg* cfg, const geomae_window_cfg* wcfg, const geomae_scatter_io* io,
                      const int32_t* tok_of_pillar, int64_t n_tokens, const geomae_window_io* out, void* stream);

/* Region batching with voxel drop for one token set, both shifts: keep[t] = 1 when token t survives, level[s][t] =
 * drop level of t's shift-s window (-1: no bucket matched, the token is dropped).  A window with n live tokens takes
 * the level l with lower[l] < n <= upper[l] and keeps max_tokens[l] of them; shift 1 buckets the survivors of
 * shift 0.  seed 0 keeps the lowest token indices (the reference with shuffle_voxels=False), any other seed a
 * uniformly random subset (shuffle_voxels=True).  max_tokens / lower / upper are HOST arrays [n_levels <= 8].
 * replaces: SSTInputLayer.drop_single_shift / get_voxel_keep_inds (middle_encoders/sst_input_layer.py:211-275)
 *           and the shuffle at :66-74. */
int geomae_window_drop(const geomae_voxel_cfg* cfg, const geomae_window_cfg* wcfg, const geomae_scatter_io* io,
                       const int32_t* tok_of_pillar, int64_t n_tokens, int32_t n_levels, const int32_t* max_tokens,
                       const int32_t* lower, const int32_t* upper, uint64_t seed, uint8_t* keep /*[n_tokens]*/,
                       int32_t* level /*[n_shifts, n_tokens]*/, void* stream);

/* canvas[b, c, y, x] = feat[i, c] for coors[i] = (b, z, y, x), zero elsewhere; canvas [n_frames, channels, ny, nx]
 * is cleared here.  _bwd gathers d_feat[i, c] = d_canvas[b, c, y, x].
 * replaces: SSTSecondPretrainedv1.recover_bev (backbones/sst_second_pretrained_v1.py:246-276). */
int geomae_recover_bev(const float* feat, const int32_t* coors, int64_t n, int32_t channels, int32_t n_frames,
                       int32_t ny, int32_t nx, float* canvas, void* stream);
int geomae_recover_bev_bwd(const float* d_canvas, const int32_t* coors, int64_t n, int32_t channels, int32_t ny,
                           int32_t nx, float* d_feat, void* stream);

/* [win_x*win_y, d_model] sinusoidal position table, row = cx*win_y + cy.
 * replaces: MultiMAESSTSPChoose.get_pos_embed (…top_only.py:361-399). */
int geomae_pos_table(int32_t win_x, int32_t win_y, int32_t d_model, float temperature, float* table, void* stream);

/* ------------------------------------------------------ voxel feature encoder */

/* out[p] = [point channels | xyz - pillar mean | xyz - pillar centre], width channels+6.
 * centre_offset = voxel/2 + range_min evaluated in double then rounded (voxel_encoder.py:155-157).
 * replaces: the decoration block of DynamicScatterVFE.forward (voxel_encoders/voxel_encoder.py:371-398). */
int geomae_vfe_decorate(const float* points, int64_t n, int32_t channels, const int32_t* point_pillar,
                        const float* pillar_mean, const int32_t* pillar_coors, const float voxel_xyz[3],
                        const float centre_offset_xyz[3], float* out, void* stream);

/* Point -> pillar reduction over the row map produced by geomae_voxel_scatter.
 * mode: 0 sum, 1 mean, 2 max.  out: [n_pillars, channels]; arg (max only): [n_pillars, channels]
 * int32 arg-max point (smallest index among ties).  pillar_mean supplies the counts for mode 1.
 * replaces: voxel_layer.dynamic_point_to_voxel_forward (mmdet3d/ops/voxel/src/voxelization.h:122-134,
 *           scatter_points_cuda.cu:80-103) and torch_scatter.scatter/scatter_max as called by
 *           scatter_v2 (ops/sst/sst_ops.py:29-32). */
int geomae_scatter_reduce_fwd(const float* feat, int64_t n_points, int32_t channels, const int32_t* point_pillar,
                              const float* pillar_mean, int64_t n_pillars, int32_t mode, float* out, int32_t* arg,
                              void* stream);

/* d_feat[p] = d_out[pillar(p)] routed to the arg-max point (max), or spread (sum / mean).
 * replaces: voxel_layer.dynamic_point_to_voxel_backward (voxelization.h:136-154). */
int geomae_scatter_reduce_bwd(const float* d_out, int64_t n_points, int32_t channels, const int32_t* point_pillar,
                              const float* pillar_mean, const int32_t* arg, int32_t mode, float* d_feat,
                              void* stream);

/* ---- DynamicScatterVFE of the GeoMAE configs as fused kernels (csrc/vfe_fused.cu): 5 raw channels -> 11 decorated
 * -> Linear(11->64, no bias) -> BN -> ReLU -> scatter-max -> [point | pillar max] -> Linear(128->128) -> BN -> ReLU ->
 * scatter-max.  BatchNorm moments `mom` = [E x (C) | E x^2 (C)] are fp64 and already averaged over ranks (equal
 * weight per rank, mmdet3d/ops/norm.py:66-73); `vmax` holds per (pillar, channel) one uint64 =
 * (order-preserving key of the post-ReLU max << 32 | ~index of the arg-max point, smallest index on ties).
 * replaces: DynamicScatterVFE.forward (voxel_encoders/voxel_encoder.py:358-419), DynamicVFELayer (utils.py:129-144),
 *           scatter_v2(mode='max') (ops/sst/sst_ops.py:8-39), naiveSyncBN1d (ops/norm.py:55-86) and their autograd. */

/* x1[n,64] = decorate(points) W0^T; stats[128] (fp64, zeroed here) = column sums and sums of squares of x1. */
int geomae_vfe0_forward(const float* points, int64_t n, int32_t channels, const int32_t* point_pillar,
                        const float* pillar_mean, const int32_t* pillar_coors, const float voxel_xyz[3],
                        const float centre_offset_xyz[3], const float* W0, float* x1, double* stats, void* stream);
/* stats[2*128] (fp64, zeroed here) of x [n,128]. */
int geomae_colstats(const float* x, int64_t n, int32_t channels, double* stats, void* stream);
/* vmax[n_pillars, C] (zeroed here) = scatter-max of relu(bn(x)); also the running-statistics update of the layer
 * (running_* may be NULL; unbias = N/(N-1) for the single-rank nn.BatchNorm1d rule, 1 for the synchronised one). */
int geomae_vfe_bn_relu_max(const float* x, int64_t n, int32_t channels, const int32_t* point_pillar, const double* mom,
                           const float* gamma, const float* beta, float eps, float* running_mean, float* running_var,
                           float momentum, float unbias, uint64_t* vmax, int64_t n_pillars, void* stream);
/* feat1[n,128] = [relu(bn0(x1)) | max of the point's pillar]. */
int geomae_vfe_cat(const float* x1, int64_t n, const int32_t* point_pillar, const double* mom, const float* gamma,
                   const float* beta, float eps, const uint64_t* vmax1, float* feat1, void* stream);
/* out[i] = value part of vmax[i]. */
int geomae_vmax_decode(const uint64_t* vmax, int64_t total, float* out, void* stream);
/* Layer-1 backward.  mode 0: sums[256] (zeroed here) = [sum g | sum g (x - mean)] with g the gradient reaching the
 * BN output (arg-max routing of d_vox, ReLU mask).  mode 1: dx2 = scale g + (ab[c] + 2 ab[128+c] x) * inv_wn, where
 * ab = geomae_bn_backward_coeffs summed over ranks and inv_wn = 1 / (world_size * n). */
int geomae_vfe1_backward(int32_t mode, const float* x2, int64_t n, const int32_t* point_pillar, const double* mom,
                         const float* gamma, const float* beta, float eps, const uint64_t* vmax2, const float* d_vox,
                         double* sums, const double* ab, float inv_wn, float* dx2, void* stream);
/* Per-channel BatchNorm backward terms from `sums`: d_gamma += sum g xhat, d_beta += sum g,
 * ab = [dL/d(mean) | dL/d(mean of squares)] of this rank. */
int geomae_bn_backward_coeffs(int32_t channels, const double* sums, const double* mom, const float* gamma, float eps,
                              double* ab, float* d_gamma, float* d_beta, void* stream);
/* d_vmax1[n_pillars,64] (zeroed here) += dfeat1[p, 64:128] of the pillar's points. */
int geomae_vfe_gather_backward(const float* dfeat1, int64_t n, const int32_t* point_pillar, float* d_vmax1,
                               int64_t n_pillars, void* stream);
/* Layer-0 backward.  mode 0: the two BatchNorm sums[128] (zeroed here).  mode 1: dW0[64,11] += dx1^T decorate(points)
 * (dx1 is formed on the fly and never stored: raw points carry no gradient). */
int geomae_vfe0_backward(int32_t mode, const float* points, int64_t n, int32_t channels, const int32_t* point_pillar,
                         const float* pillar_mean, const int32_t* pillar_coors, const float voxel_xyz[3],
                         const float centre_offset_xyz[3], const float* x1, const double* mom, const float* gamma,
                         const float* beta, float eps, const uint64_t* vmax1, const float* d_vmax1,
                         const float* dfeat1, double* sums, const double* ab, float inv_wn, float* dW0, void* stream);

/* ------------------------------------------------- sparse regional attention */

/* Multi-head attention inside CSR windows (no padding, no key_padding_mask needed).
 * qkv: [n_tokens, 3*d_model] fp32, rows in flat token order: q | k | v, q NOT yet scaled.
 * win_ptr / win_tok / tok_win: one shift of a geomae_window_io.
 * out: [n_tokens, d_model]; lse: [n_tokens, n_heads] log-sum-exp saved for backward.
 * head_dim must be 16 (d_model = 16*n_heads), n_heads <= 8.
 * replaces: nn.MultiheadAttention core inside WindowAttention.forward
 *           (models/sst/sst_basic_block.py:26-61) on padded [T,W,C] buckets. */
int geomae_sra_attention_fwd(const float* qkv, int64_t n_tokens, int32_t n_heads, const int32_t* win_ptr,
                             const int32_t* win_tok, const int32_t* tok_win, float* out, float* lse, void* stream);

/* Backward of the above: d_qkv [n_tokens, 3*d_model] from d_out, recomputing the probabilities.
 * scratch: [n_tokens * n_heads] floats. */
int geomae_sra_attention_bwd(const float* qkv, const float* out, const float* lse, const float* d_out,
                             int64_t n_tokens, int32_t n_heads, const int32_t* win_ptr, const int32_t* win_tok,
                             const int32_t* tok_win, float* d_qkv, float* scratch, void* stream);

/* The same attention with bf16 operands on the tensor cores (mma.sync m16n8k16, fp32 accumulate and softmax):
 * windows are packed block-diagonally into 16-query x 16-key tiles; K|V of the contiguous CSR range a CTA can
 * see are converted to bf16 while being staged into shared memory.  Same arguments and outputs as the fp32
 * entry points above (no scratch: D = dO.O is read from `dd` when given, else recomputed from the staged rows).  Used by the SRA stack executor
 * when precision == 1 (the bf16 benchmark mode); precision == 3 keeps the fp32 kernels.
 * replaces: nn.MultiheadAttention core inside WindowAttention.forward
 *           (models/sst/sst_basic_block.py:26-61) and its autograd backward. */
/* io_flags: bit 0: qkv rows are bf16 [n, 3*d_model]; bit 1: d_out rows are bf16 (needs dd); bit 2: d_qkv is written
 * as bf16; bit 3 (forward): out is written as bf16 [n, d_model].  bf16 rows are moved with 16-byte cp.async copies
 * (no registers, no conversion). */
int geomae_sra_attention_tc_fwd(const float* qkv, int64_t n_tokens, int32_t n_heads, const int32_t* win_ptr,
                                const int32_t* win_tok, const int32_t* tok_win, float* out, float* lse, int32_t io_flags,
                                void* stream);
int geomae_sra_attention_tc_bwd(const float* qkv, const float* out, const float* lse, const float* d_out,
                                int64_t n_tokens, int32_t n_heads, const int32_t* win_ptr, const int32_t* win_tok,
                                const int32_t* tok_win, float* d_qkv, const float* dd /* [n_tokens, n_heads] D = dO.O, or NULL */,
                                int32_t io_flags, void* stream);

/* ------------------------------------------------ tensor-core dense layers */

/* out[128-token tile, N] = prologue(A)[., K] x W + epilogue, on tcgen05 tensor cores (bf16 operands,
 * fp32 accumulation in TMEM).  precision 1 = bf16, 3 = bf16x3 split (fp32-equivalent parity mode).
 *   prologue : A += pos_table[tok_cell[row]] for the first pos_slabs 128-column slabs of the output
 *              (q and k of the in-projection use x+pos, v uses x); a_gelu: A = gelu(A)
 *   weights  : w_mn_major 0: W[n][k] row-major (y = x W^T, nn.Linear forward)
 *              w_mn_major 1: W[k][n] row-major (dX = dY W, nn.Linear input gradient) — no transposed copy
 *   epilogue : 0  out = acc (+ bias) (+ add_src)
 *              1  out = LayerNorm(acc + bias + add_src) * gamma + beta; ln_in / ln_stats (mean, rstd) saved
 *              2  out = acc * gelu'(gelu_u)
 * replaces: the nn.MultiheadAttention in/out projections, EncoderLayer.linear1/linear2, norm1/norm2 and the
 *           activation of models/sst/sst_basic_block.py:55,94-100 (library GEMM + 5-6 elementwise kernels). */
typedef struct geomae_linear_args {
  const float* A; int32_t lda; int32_t n_rows; int32_t K;
  const float* pos_table; const int32_t* tok_cell; int32_t pos_slabs; int32_t a_gelu;
  const float* W; int32_t ldw; int32_t w_rows; int32_t w_mn_major;
  const float* bias; int32_t N_total;
  float* out; int32_t ldo;
  const float* add_src; int32_t ld_add;
  const float* ln_gamma; const float* ln_beta; float ln_eps; float* ln_in; float* ln_stats;
  const float* gelu_u; int32_t ldu;
  int32_t epilogue; int32_t precision;
  const void* Wp_hi; const void* Wp_lo;   /* optional: images of W from geomae_pack_weights (ldw = cols of W) */
  /* optional side output of epilogue 0 with N_total == 128: dot_out[row, h] = sum over the 16 columns of head h of
   * out[row, .] * dot_src[row, .]  (h = 0..7).  The attention backward needs D = dO . O per (token, head); the
   * GEMM that produces dO computes it on the way out instead of both attention passes re-reading dO and O. */
  const float* dot_src; int32_t ld_dot; float* dot_out;
  /* epilogue 3 (N_total == 128, no bias): LayerNorm BACKWARD.  acc (+ add_src) is the gradient dz reaching a LayerNorm
   * output; with ln_in (saved pre-LN rows, INPUT here), ln_stats (mean, rstd) and ln_gamma the kernel writes
   * out = d(pre-LN rows) and accumulates (+=) ln_dgamma, ln_dbeta and (optional) ln_dcolsum = column sums of out.
   * replaces: ATen layer_norm backward after the matching dX GEMM — dz is never stored. */
  float* ln_dgamma; float* ln_dbeta; float* ln_dcolsum;
  /* bf16 storage of tensors that only ever serve as tensor-core operands (precision 1): a_bf16: A rows are bf16
   * (lda in elements; no position / GELU prologue); out_bf16: epilogue 0 / 2 writes bf16 rows (ldo in elements).
   * Numerically identical to the fp32-storage path, which rounds the same values to bf16 when staging them. */
  int32_t a_bf16; int32_t out_bf16;
} geomae_linear_args;

int geomae_tc_linear(const geomae_linear_args* args, void* stream);

/* Pre-pack weights for the tensor-core kernels: fp32 W[k] [rows, cols] (row-major, cols % 64 == 0) ->
 * bf16 "hi" image and bf16 residual "lo" image (lo may be NULL), each ceil(rows/128)*128*cols*2 bytes, laid
 * out as [128 x 64] 128B-swizzled blocks so a CTA loads its weight tile with bulk copies.  One launch per
 * 64 items. */
int geomae_pack_weights(int32_t n_items, const float* const* W, const int32_t* rows, const int32_t* cols,
                        void* const* hi, void* const* lo, void* stream);

/* dW[M_total, N_total] += dY^T X over the token rows, db[M_total] += column sums of dY (db may be NULL).
 * X prologue: + pos_table[tok_cell] for the first pos_slabs 128-row slabs of dW (q,k rows of in_proj), gelu.
 * Accumulates with fp32 vector reductions: the destination must hold the running gradient (or zeros).
 * replaces: the weight/bias gradient GEMMs autograd runs for nn.Linear / MultiheadAttention in
 *           models/sst/sst_basic_block.py (library sgemm + column-sum reductions). */
typedef struct geomae_wgrad_args {
  const float* dY; int32_t ldy; const float* X; int32_t ldx; int32_t n_rows;
  const float* pos_table; const int32_t* tok_cell; int32_t pos_slabs; int32_t x_gelu;
  float* dW; int32_t ldw; float* db; int32_t M_total; int32_t N_total;
  int32_t precision;
  int32_t dy_bf16;     /* dY rows are bf16 (ldy in elements), precision 1 */
} geomae_wgrad_args;

int geomae_tc_wgrad(const geomae_wgrad_args* args, void* stream);

/* LayerNorm backward from the saved pre-LN rows and (mean, rstd): d_in, and d_gamma / d_beta accumulated
 * (+=) into their buffers.  channels must be 128.  d_in_colsum (optional, [channels], +=) receives the column
 * sums of d_in, which are the bias gradient of the linear layer whose output (+ residual) was normalised
 * (out_proj for norm1, linear2 for norm2) — that bias reduction then costs nothing extra.
 * replaces: ATen layer_norm backward (GammaBetaBackward + grad_input kernels) under nn.LayerNorm
 *           (sst_basic_block.py:76-77,96,100) and the bias-gradient column reductions of linear2 / out_proj. */
int geomae_layernorm_bwd(const float* d_out, const float* ln_in, const float* ln_stats, const float* gamma,
                         int64_t n_rows, int32_t channels, float* d_in, float* d_gamma, float* d_beta,
                         float* d_in_colsum, void* stream);

/* ------------------------------------------------------- SRA layer stacks */

typedef struct geomae_sra_windows {   /* one shift of a geomae_window_io */
  const int32_t* win_ptr; const int32_t* win_tok; const int32_t* tok_win; const int32_t* tok_cell;
} geomae_sra_windows;

typedef struct geomae_sra_ctx {
  int64_t n_tokens; int32_t d_model; int32_t n_heads; int32_t ffn; int32_t precision;   /* 1 bf16, 3 bf16x3 */
  const float* pos_table;             /* [win_x*win_y, d_model] */
  geomae_sra_windows shift[2];
  void* pos16[2];                     /* precision 1: scratch [n_tokens, d_model] bf16 per shift (gathered position rows) */
} geomae_sra_ctx;

/* Parameters (and their fp32 gradient accumulators) of one EncoderLayer, reference names in comments. */
typedef struct geomae_sra_layer {
  int32_t shift; float ln_eps;
  const float *in_proj_w, *in_proj_b;     /* win_attn.self_attn.in_proj_{weight,bias}  [3d,d],[3d] */
  const float *out_proj_w, *out_proj_b;   /* win_attn.self_attn.out_proj.{weight,bias} [d,d],[d]   */
  const float *lin1_w, *lin1_b;           /* linear1.{weight,bias} [f,d],[f] */
  const float *lin2_w, *lin2_b;           /* linear2.{weight,bias} [d,f],[d] */
  const float *norm1_w, *norm1_b, *norm2_w, *norm2_b;
  float *g_in_proj_w, *g_in_proj_b, *g_out_proj_w, *g_out_proj_b, *g_lin1_w, *g_lin1_b, *g_lin2_w, *g_lin2_b;
  float *g_norm1_w, *g_norm1_b, *g_norm2_w, *g_norm2_b;
  /* scratch for the packed bf16 images of the four weight matrices (hi, lo): in_proj 2*3d*d bytes each,
   * out_proj 2*d*d, lin1 2*f*d, lin2 2*d*f.  (Re)filled by geomae_sra_stack_forward. */
  void *p_in_proj[2], *p_out_proj[2], *p_lin1[2], *p_lin2[2];
} geomae_sra_layer;

/* Activations one layer keeps for its backward (caller-allocated, n = n_tokens). */
typedef struct geomae_sra_saved {
  float* qkv;  /* [n,3d] */ float* attn; /* [n,d] */ float* lse; /* [n,heads] */
  float* s1;   /* [n,d] pre-LN1 */ float* st1; /* [n,2] */ float* y; /* [n,d] */
  float* u;    /* [n,f] pre-GELU */ float* s2; /* [n,d] pre-LN2 */ float* st2; /* [n,2] */ float* z; /* [n,d] output */
  /* precision 1 (bf16 mode): qkv, attn, u hold bf16 rows; s1 / s2 hold the bf16 NORMALISED rows xhat1 / xhat2 (y is
   * not stored); g = gelu(u) [n,f] bf16; xb (layer 0 only) = bf16 copy of the stack input; xp is unused. */
  void* g; void* xp; void* xb;
} geomae_sra_saved;

/* Forward of n_layers EncoderLayers: layer l reads layer l-1's z (layer 0 reads x_in); 5 kernels per layer.
 * replaces: BasicShiftBlock / EncoderLayer / WindowAttention forward (models/sst/sst_basic_block.py:26-147)
 *           as driven by forward_encoder / forward_decoder (backbones/…top_only.py:230-232,271-277). */
int geomae_sra_stack_forward(const geomae_sra_ctx* ctx, int32_t n_layers, const geomae_sra_layer* layers,
                             const geomae_sra_saved* saved, const float* x_in, void* stream);

/* Backward: d_out = gradient w.r.t. the last layer's z, d_in receives the gradient w.r.t. x_in, parameter
 * gradients are ACCUMULATED into the g_* buffers.  scratch: geomae_sra_scratch_floats(ctx, 1) floats.
 * The weight-gradient GEMMs run on an internal side stream overlapping the dX chain; the call returns with
 * all work ordered after `stream`. */
int geomae_sra_stack_backward(const geomae_sra_ctx* ctx, int32_t n_layers, const geomae_sra_layer* layers,
                              const geomae_sra_saved* saved, const float* x_in, const float* d_out, float* d_in,
                              float* scratch, void* stream);
int64_t geomae_sra_scratch_floats(const geomae_sra_ctx* ctx, int32_t n_stacks);

/* Two stacks reading the same input (decoder_centroid_blocks / decoder_density_blocks,
 * backbones/…top_only.py:269-277), run concurrently on two streams.  scratch: geomae_sra_scratch_floats(ctx, 2). */
int geomae_sra_stack2_forward(const geomae_sra_ctx* ctx, int32_t n_layers, const geomae_sra_layer* layers_a,
                              const geomae_sra_saved* saved_a, const geomae_sra_layer* layers_b,
                              const geomae_sra_saved* saved_b, const float* x_in, void* stream);
int geomae_sra_stack2_backward(const geomae_sra_ctx* ctx, int32_t n_layers, const geomae_sra_layer* layers_a,
                               const geomae_sra_saved* saved_a, const geomae_sra_layer* layers_b,
                               const geomae_sra_saved* saved_b, const float* x_in, const float* d_out_a,
                               const float* d_out_b, float* d_in_a, float* d_in_b, float* scratch, void* stream);

/* ---------------------------------------- fused token-local layer chains (bf16 mode) */

/* Forward chain of one EncoderLayer for every 128-token tile in ONE persistent, warp-specialised tcgen05 kernel
 * (csrc/sra_chain.cu):  s1 = x + attn Wo^T + bo ; y = LN1(s1) ; u = y W1^T + b1 ; g = gelu(u) ; s2 = y + g W2^T + b2 ;
 * z = LN2(s2) ; and (mode bit 1) the in-projection of the NEXT layer, q|k = (z + pos) Wqk^T + b, v = z Wv^T + b.
 * mode: bit 0 = this layer's chain, bit 1 = next in-projection (mode 2 alone = in-projection of x: the stack prologue).
 * p_*: packed bf16 "hi" images from geomae_pack_weights.  Saved for the backward, bf16 row-major [n, cols]:
 *   xh1_16, xh2_16 [n,128] = the NORMALISED rows (s - mean) * rstd of the two LayerNorms (their outputs y, z are
 *   affine in them per column, so neither y nor the pre-LN rows are stored), u16 [n,256] (pre-GELU), g16 [n,256]
 *   (gelu(u)), qkv16_next [n,384]; fp32: st1, st2 [n,2] (mean, rstd), z [n,128] (the residual stream stays fp32).
 *   xb16 [n,128] (mode 2 only): bf16 copy of x, the operand of layer 0's in-projection weight gradient.
 * attn [n,128] bf16 = the attention output.  Every tile tensor moves by TMA box loads / stores.
 * replaces: out_proj of nn.MultiheadAttention + EncoderLayer.forward (models/sst/sst_basic_block.py:55,85-102) and
 *           the in_proj of the next layer's WindowAttention (:41-55) — library GEMMs + ~8 elementwise kernels. */
typedef struct geomae_chain_fwd_args {
  int64_t n_tokens; int32_t mode;
  const float* x; const void* attn;
  const void *p_out_proj, *p_lin1, *p_lin2, *p_in_proj_next;
  const float *out_proj_b, *lin1_b, *lin2_b, *in_proj_b_next, *norm1_w, *norm1_b, *norm2_w, *norm2_b; float ln_eps;
  const float* pos_table; const int32_t* tok_cell_next;
  float *st1, *st2, *z;
  void *xh1_16, *xh2_16, *u16, *g16, *xb16, *qkv16_next;
} geomae_chain_fwd_args;

int geomae_sra_chain_fwd(const geomae_chain_fwd_args* args, void* stream);

/* Backward of the same chain (k_sra_chain_bwd), gradients flowing DOWN through one EncoderLayer per 128-token tile:
 *   mode bit 0: dz = dqkv16_up Win_up + ds1_up  (in-projection backward of the layer ABOVE; else dz = dz_in, fp32)
 *   mode bit 1: ds2 = LN2-bwd(dz; xh2_16, st2) ; du = (ds2 W2) gelu'(u) ; dy = du W1 + ds2 ; ds1 = LN1-bwd(dy; xh1_16, st1) ;
 *               dattn = ds1 Wo ;
 *               dd[n,8] = per-head dot(dattn, attn)  (the attention backward's row term D)
 *   mode 1 alone writes dx = dz (the input gradient of the stack's first layer).
 * Outputs: ds2_16, ds1_16, dattn16 [n,128], du16 [n,256] bf16 (operands of geomae_sra_wgrad_layer and of the attention
 * backward), ds1 [n,128] fp32 (residual-gradient term of the layer below); g_norm* accumulate (+=).  The bias
 * gradients of linear2 / out_proj are column sums of ds2 / ds1 and are produced by geomae_sra_wgrad_layer.
 * replaces: autograd of EncoderLayer.forward + the MultiheadAttention projections
 *           (models/sst/sst_basic_block.py:55,85-102): 5 dX GEMMs, 2 LayerNorm backwards, GELU backward. */
typedef struct geomae_chain_bwd_args {
  int64_t n_tokens; int32_t mode;
  const void* dqkv16_up; const float* ds1_up; const void* p_in_proj_up; const float* dz_in;
  const void* xh2_16; const float* st2; const void* xh1_16; const float* st1; const void *u16, *attn16;
  const void *p_lin2, *p_lin1, *p_out_proj; const float *norm2_w, *norm1_w;
  void *ds2_16, *du16, *ds1_16, *dattn16; float *ds1, *dd, *dx;
  float *g_norm2_w, *g_norm2_b, *g_norm1_w, *g_norm1_b;
} geomae_chain_bwd_args;

int geomae_sra_chain_bwd(const geomae_chain_bwd_args* args, void* stream);

/* Weight / bias gradients of one EncoderLayer in ONE TMA-fed tcgen05 launch (csrc/sra_wgrad.cu), accumulated (+=)
 * with vector reductions:  g_lin2_w [128,256] += ds2^T g ; g_lin1_w [256,128] += du^T y, g_lin1_b += colsum du ;
 * g_out_proj_w [128,128] += ds1^T attn ; g_in_proj_w [384,128] += dq|dk ^T (x + pos), dv^T x ; g_in_proj_b += colsum dqkv ;
 * g_lin2_b += colsum ds2 ; g_out_proj_b += colsum ds1.
 * Operands are bf16 row-major (16-byte aligned): ds2_16, ds1_16, attn16 [n,128]; g16, du16 [n,256]; dqkv16 [n,384];
 *   xh1_16 [n,128]: y = xh1 * norm1_w + norm1_b is applied in the flush (per-column scale + rank-1 term);
 *   xin16 [n,128] with in_scale / in_shift: the layer input x = xin * in_scale + in_shift (the xhat2 tile and norm2
 *   parameters of the layer below), or x itself with in_scale = in_shift = NULL (first layer of a stack);
 *   pos16 [n,128]: the gathered position rows of this layer's shift (geomae_pos_rows_bf16).
 * Bias buffers may be NULL.
 * replaces: the weight-gradient GEMMs + bias column sums autograd runs for linear1 / linear2 / out_proj / in_proj
 *           of models/sst/sst_basic_block.py:55,94-100. */
typedef struct geomae_wgrad_layer_args {
  int64_t n_tokens;
  const void *ds2_16, *g16, *du16, *xh1_16, *ds1_16, *attn16, *dqkv16, *xin16, *pos16;
  const float *norm1_w, *norm1_b, *in_scale, *in_shift;
  float *g_lin2_w, *g_lin1_w, *g_lin1_b, *g_out_proj_w, *g_in_proj_w, *g_in_proj_b, *g_lin2_b, *g_out_proj_b;
} geomae_wgrad_layer_args;

/* out16[i, :] = bf16(pos_table[tok_cell[i], :]), [n,128]: once per token set and shift. */
int geomae_pos_rows_bf16(const float* pos_table, const int32_t* tok_cell, int64_t n_tokens, void* out16, void* stream);

int geomae_sra_wgrad_layer(const geomae_wgrad_layer_args* args, void* stream);

/* -------------------------------------------------------------------- losses */

typedef struct geomae_loss_args {
  const int64_t* rows; int64_t m;                 /* masked pillar rows (ids_mask) */
  const float* reg_low;  /* [m,slots_low,3] */ const float* reg_med; /* [m,slots_med,3] */
  const float* reg_top;  /* [m,3] */           const float* nor_top; /* [m,3] */
  const float* cls_low;  /* [m,slots_low,2] */ const float* cls_med; /* [m,slots_med,2] */
  const float* normal;   /* [V,3] per-pillar normal targets (geomae_geom_targets) */
  float w_low, w_med, w_top, w_nor, w_cls_low, w_cls_med;   /* loss_ratio_* of the config */
  /* row strides (floats) of reg_low, reg_med, reg_top, nor_top, cls_low, cls_med and of their gradients; 0 = the
   * dense default (slots*3, slots*3, 3, 3, slots*2, slots*2).  Non-default strides let the six predictions be
   * column slices of ONE fused head GEMM output [m, 768] (backbone heads, …top_only.py:279-300). */
  int32_t ld[6];
} geomae_loss_args;

/* out[6] = loss_curv_around, loss_centroid_low, loss_centroid_med, loss_centroid_top, loss_cls_low, loss_cls_med.
 * Masked MSE (mean over xyz, mean over occupied slots) and occupancy BCE-with-logits against the one-hot
 * occupancy label, evaluated from the CSR sub-voxel lists.  counts [2] i32 and acc [6] f64 are scratch; counts is
 * re-used by the backward.
 * replaces: forward_loss (detectors/…_ssl.py:837-902) incl. mmdet CrossEntropyLoss(use_sigmoid=True), and the
 *           dense-target construction it consumes. */
int geomae_geom_loss_fwd(const geomae_voxel_cfg* cfg, const geomae_scatter_io* io, const geomae_loss_args* args,
                         int32_t* counts, double* acc, float* out, void* stream);

/* Gradients of the six prediction tensors given d_losses[6] (upstream gradient of each loss). */
int geomae_geom_loss_bwd(const geomae_voxel_cfg* cfg, const geomae_scatter_io* io, const geomae_loss_args* args,
                         const int32_t* counts, const float* d_losses, float* d_reg_low, float* d_reg_med,
                         float* d_reg_top, float* d_nor_top, float* d_cls_low, float* d_cls_med, void* stream);

/* Per-kernel-family device timing of the stack executors (CUDA events on the launching streams), used by
 * bench.py for the roofline entry.  Families: 0 tc_linear, 1 tc_wgrad, 2 attention fwd, 3 attention bwd
 * (2 kernels per span), 4 layernorm bwd.  geomae_profile_read synchronises the device and resets the log. */
int geomae_profile_enable(int32_t on);
int geomae_profile_read(double* ms /*[5]*/, int64_t* launches /*[5]*/, double* flops /*[5]*/,
                        double* bytes /*[5] algorithmic bytes of the launches, may be NULL*/);

/* ---------------------------------------------------------------- optimiser */

/* One fused step over flat fp32 buffers: g' = g*grad_scale (1/world_size), clip by global L2 norm
 * to max_norm (<=0 disables), AdamW update (decoupled weight decay on the first n_decay elements
 * only).  partials: [1024] f64 scratch.  stats [opt]: [2] = grad norm, clip coefficient.
 * replaces: mmcv OptimizerHook(grad_clip) + torch.optim.AdamW as configured by
 *           configs/_base_/schedules/cosine_2x.py:1-9 (paramwise 'norm' decay_mult=0). */
int geomae_adamw_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n,
                      int64_t n_decay, double* partials, float grad_scale, float max_norm, float lr, float beta1,
                      float beta2, float eps, float weight_decay, int64_t step, float* stats, void* stream);

/* ------------------------------------------------- data step in front of the path */

/* Per-sample augmentation + range filter of a concatenated batch, fused with a stable compaction:
 *   (x, y, z) -> ((x c - y s) * scale, (x s + y c) * scale, z * scale), y = -y if flip bit 0, x = -x if flip bit 1,
 *   keep iff range_min < (x, y, z) < range_max (strict); further channels are copied unchanged.
 * frame_params: [n_frames, 4] = cos, sin, scale, flip bits (as a float).  Survivors keep their input order;
 * out_frame_offsets [n_frames+1] are the new frame starts (last = number of survivors).  out_points must hold
 * n_points rows, scan_tmp ceil(n_points/1024)+1 ints.  Caller-allocated, no device sync.
 * replaces: GlobalRotScaleTrans, RandomFlip3D and PointsRangeFilter of the pretraining pipeline
 *           (configs/mae_sst/…6x_1e-5.py:181-193; datasets/pipelines/transforms_3d.py:95-123,670-718,849-883;
 *           core/points/base_points.py:139-179,207-229,263-269; core/points/lidar_points.py:28-33), which the reference
 *           applies per sample on the data-loader workers.  PointShuffle (transforms_3d.py:771-790) is not reproduced:
 *           everything downstream is order-free. */
/* (rows of `points` at or beyond frame_offsets[n_frames] are ignored: the buffer may be the over-sized output of an
 * earlier compaction whose length only the device knows.) */
int geomae_augment_filter(const float* points, int64_t n_points, int32_t stride, const int32_t* frame_offsets,
                          int32_t n_frames, const float* frame_params, const float range_min[3],
                          const float range_max[3], float* out_points, int32_t* out_frame_offsets,
                          int32_t* scan_tmp, int64_t scan_tmp_len, void* stream);

/* ------------------------------------------------------ peer-memory exchange (multi-GPU, one node) */

/* mailbox[r] = rank r's mailbox as mapped into THIS process (own allocation for r == rank, CUDA-IPC mapping of the
 * peer's allocation otherwise), geomae_peer_mailbox_doubles(world) doubles each, zero-initialised once.
 * timeout_flag: optional device int32 set to 1 when a peer did not arrive within ~10 s. */
typedef struct geomae_peer_ctx {
  int32_t rank, world;
  void* mailbox[8];
  void* timeout_flag;
} geomae_peer_ctx;

int64_t geomae_peer_mailbox_doubles(int32_t world);
/* Own mailbox: device memory of the current device, zeroed, with its 64-byte CUDA IPC handle (send it to the other
 * ranks of the node by any means).  _open maps a peer's mailbox from the CURRENT device (peer access is enabled
 * lazily by the driver); _close unmaps (owned = 0) or frees (owned = 1). */
int geomae_peer_mailbox_create(int32_t world, void** mailbox, void* ipc_handle_64);
/* Any buffer meant to be mapped by the other ranks (the flat gradient buffer): zeroed device memory + IPC handle;
 * map with geomae_peer_mailbox_open, release with geomae_peer_mailbox_close. */
int geomae_peer_buffer_create(int64_t bytes, void** buffer, void* ipc_handle_64);
int geomae_peer_mailbox_open(const void* ipc_handle_64, void** mapped);
int geomae_peer_mailbox_close(void* ptr, int32_t owned);
/* Kernels of the CURRENT device may dereference memory of peer_device from now on (cudaDeviceEnablePeerAccess). */
int geomae_peer_enable_access(int32_t peer_device);

/* In place: buf[i] = post_scale * sum over ranks of (pre_scale_r * buf_r[i]), count <= 512 doubles, ONE single-CTA
 * kernel per rank: peer stores into every rank's mailbox over NVLink, a system-scope release flag, acquire-spin on the
 * own mailbox, sum in rank order.  Every rank must issue the same sequence of calls with epoch = 1, 2, 3, ...
 * world == 1: just the scaling (no mailbox needed).
 * replaces: the four dist.all_reduce / all_gather calls per step of naiveSyncBN1d (mmdet3d/ops/norm.py:28-86) and the
 *           stats / n normalisation in front of them. */
int geomae_peer_allreduce_f64(const geomae_peer_ctx* ctx, double* buf, int32_t count, double pre_scale,
                              double post_scale, uint64_t epoch, void* stream);

/* Gradient exchange over peer memory: grads[r] = rank r's gradient buffer as mapped here (float32).  This rank sums
 * its 1/world slice of the element range [lo, hi) (multiples of 4) over all ranks, in rank order, and stores the sum
 * into every rank's buffer — reduce-scatter + all-gather in one launch, bit-identical results on all ranks.  The
 * caller orders it between two geomae_peer_allreduce_f64 calls used as barriers (all gradients complete / all slices
 * written).
 * replaces: the DDP gradient all-reduce (MMDistributedDataParallel; reference apis/train.py:117-127). */
int geomae_peer_reduce_shard(const geomae_peer_ctx* ctx, void* const* grads, int64_t lo, int64_t hi, void* stream);

/* tokens [n_vis + n_mask, d_model] = [visible rows ; mask_token repeated n_mask times].
 * replaces: torch.cat([visible_voxel_feat, self.mask_token.repeat(n_mask, 1)]) (…top_only.py:204-210). */
int geomae_decoder_tokens(const float* visible, int64_t n_vis, const float* mask_token, int64_t n_mask, int32_t d_model,
                          float* tokens, void* stream);
/* g_mask_token[c] += sum over the n_mask trailing rows of d_tokens[:, c] (the gradient of the repeat); the visible
 * rows' gradient is the leading slice of d_tokens itself.  d_model must be 128. */
int geomae_mask_token_grad(const float* d_tokens, int64_t n_vis, int64_t n_mask, int32_t d_model, float* g_mask_token,
                           void* stream);

/* Multi-sweep merge on the device: `points` holds the raw records of S segments back to back (segment = one
 * `.pcd.bin` file: the key frame first, then its earlier sweeps; several samples may follow each other),
 * seg_offsets [S+1].  seg_params [S,16] doubles per segment: sensor->key-frame rotation R row-major (9), translation
 * (3), time lag in seconds, close radius (< 0: keep all points), 2 unused.  Per point: dropped when |x| < r and
 * |y| < r; else xyz = f32(f32(row . R^T in float64) + t in float64), channel 4 = (float)time lag; survivors are
 * compacted in input order, out_seg_offsets [S+1] are the segments' new starts (a sample's frame offsets are the
 * entries of its first segment).  scan_tmp: int32 [ceil(n/1024)+1].
 * replaces: LoadPointsFromMultiSweeps.__call__ / _remove_close (datasets/pipelines/loading.py:160-235), numpy on the
 *           data-loader workers in the reference. */
int geomae_sweep_merge(const float* points, int64_t n_points, int32_t stride, const int32_t* seg_offsets,
                       int32_t n_segments, const double* seg_params, float* out_points, int32_t* out_seg_offsets,
                       int32_t* scan_tmp, int64_t scan_tmp_len, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GEOMAE_B200_H */
