// Weight gradients of one SRA EncoderLayer in ONE launch, fed by TMA (bf16 mode):
//   dW2  += ds2^T gelu(u)      dW1 += du^T y  (+ db1)      dWo += ds1^T O      dWin += dqkv^T (x+pos | x)  (+ dbin)
// The LayerNorm outputs y = xhat1*g1 + b1 and x = xhat2'*g2' + b2' (' = the layer below) are never stored: the chain
// kernels save only the bf16 xhat tiles the LayerNorm backward needs anyway, and because the affine map is per COLUMN of
// the operand it moves into the flush:  dY^T (xhat*g + b) = g[n] * (dY^T xhat)[m][n] + (sum_tok dY)[m] * b[n],  the second
// factor being the bias-gradient column this kernel accumulates anyway.  The position term of the q|k rows,
// dqk^T pos[cell], is two more slabs against the gathered position rows (one bf16 tensor per token set and shift).
// i.e. the weight / bias gradient GEMMs autograd runs for nn.Linear / nn.MultiheadAttention in
// models/sst/sst_basic_block.py:55,94-100.  Every operand is a token-major bf16 tensor [n, C] that the forward /
// backward chain kernels saved; dW = dY^T X contracts over TOKENS, so both MMA operands are MN-major views of
// [128 tokens x 64 channels] boxes, which is exactly what a 2-D tiled TMA load with the 128-byte swizzle writes into
// shared memory — no thread ever touches an operand.  A CTA owns one [128 x 128] slab of one dW and a range of
// 128-token tiles: warp 0 issues cp.async.bulk.tensor loads (UTMALDG) into a 3-stage ring, one thread of warp 1
// issues tcgen05.mma (+ an N = 16 MMA against a constant "ones" block whose result column is the bias gradient),
// the accumulator stays in TMEM over the whole token range and four warps flush it once with vector reductions.
// The round-1 kernel re-staged fp32 rows through registers per slab (2-6 % tensor-pipe activity, 25 % of the step).
#include <cuda_bf16.h>

#include "common.cuh"
#include "tc_common.cuh"
#include "tma_host.cuh"

namespace {

constexpr int WT = 128;                       // tokens per stage = K of one stage
constexpr int STAGES = 3;
constexpr int BLK = 16384;                    // one [128 x 64] bf16 box
constexpr int STAGE_BYTES = 4 * BLK;          // dY box pair + X box pair
constexpr int OFF_ONES = STAGES * STAGE_BYTES;
constexpr int OFF_BAR = OFF_ONES + BLK;
constexpr int SMEM_BYTES = OFF_BAR + 128 + 1024;
constexpr int MAX_SLABS = 10;
constexpr int NTHR = 192;

struct Slab { int a_map, a_col, b_map, b_col; float* dW; int ldw; float* db; const float* scale; const float* shift; };
struct WgArgs { int n_tiles; int tiles_per_cta; Slab slab[MAX_SLABS]; };
struct WgMaps { CUtensorMap m[9]; };

__global__ void __launch_bounds__(NTHR, 1) k_wgrad_tma(const __grid_constant__ WgMaps maps, const WgArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* sm = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sOnes = sm + OFF_ONES;
  uint64_t* full = reinterpret_cast<uint64_t*>(sm + OFF_BAR);
  uint64_t* empty = full + STAGES;
  uint64_t* acc_bar = empty + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_bar + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const Slab s = a.slab[blockIdx.y];
  const int tile_begin = blockIdx.x * a.tiles_per_cta;
  const int tile_end = min(tile_begin + a.tiles_per_cta, a.n_tiles);
  if (tile_begin >= tile_end) return;
  if (threadIdx.x == 0) {
    for (int i = 0; i < STAGES; ++i) { tc::mbar_init(&full[i], 1); tc::mbar_init(&empty[i], 1); }
    tc::mbar_init(acc_bar, 1);
  }
  if (warp == 1) tc::tmem_alloc(tmem_slot, 256);
  // constant "ones" block: element (token row, column 0) = 1.0 -> D[m][0] = sum over tokens of dY[tok][m]
  for (int i = threadIdx.x; i < WT * 8; i += NTHR) {
    const int r = i >> 3, c = i & 7;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (c == 0) v.x = 0x00003F80u;
    *reinterpret_cast<uint4*>(sOnes + tc::swz(r, c)) = v;
  }
  tc::fence_async_smem();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = *tmem_slot;
  gm_pdl_wait();
  gm_pdl_trigger();
  if (warp == 0) {
    if (lane == 0) {
      const CUtensorMap* ma = &maps.m[s.a_map];
      const CUtensorMap* mb = &maps.m[s.b_map];
      uint32_t cnt = 0;
      for (int t = tile_begin; t < tile_end; ++t, ++cnt) {
        const uint32_t slot = cnt % STAGES, ph = (cnt / STAGES) & 1;
        uint8_t* st = sm + slot * STAGE_BYTES;
        tc::mbar_wait(&empty[slot], ph ^ 1);
        tc::mbar_expect_tx(&full[slot], STAGE_BYTES);
        tma::load_2d(st, ma, s.a_col, t * WT, &full[slot]);
        tma::load_2d(st + BLK, ma, s.a_col + 64, t * WT, &full[slot]);
        tma::load_2d(st + 2 * BLK, mb, s.b_col, t * WT, &full[slot]);
        tma::load_2d(st + 3 * BLK, mb, s.b_col + 64, t * WT, &full[slot]);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = tc::make_idesc_bf16(128, 128, 1, 1);
      const uint32_t idesc_b = tc::make_idesc_bf16(128, 16, 1, 1);
      const uint32_t ones = tc::smem_u32(sOnes);
      const bool bias = s.db != nullptr || s.shift != nullptr;
      uint32_t cnt = 0;
      for (int t = tile_begin; t < tile_end; ++t, ++cnt) {
        const uint32_t slot = cnt % STAGES, ph = (cnt / STAGES) & 1;
        tc::mbar_wait(&full[slot], ph);
        tc::fence_after_sync();
        const uint32_t A = tc::smem_u32(sm + slot * STAGE_BYTES), B = A + 2 * BLK;
#pragma unroll
        for (int j = 0; j < WT / 16; ++j) {            // 16 tokens per MMA: two 8-row swizzle atoms
          const uint32_t off = (uint32_t)j * 2 * tc::ATOM_BYTES;
          const uint64_t da = tc::make_desc(A + off, BLK, tc::ATOM_BYTES);
          const bool acc = cnt > 0 || j > 0;
          tc::mma_bf16(tmem, da, tc::make_desc(B + off, BLK, tc::ATOM_BYTES), idesc, acc);
          if (bias) tc::mma_bf16(tmem + 128, da, tc::make_desc(ones + off, BLK, tc::ATOM_BYTES), idesc_b, acc);
        }
        tc::mma_commit(&empty[slot]);
      }
      tc::mma_commit(acc_bar);
    }
  } else {
    // ---- epilogue warps 2..5: TMEM lane quarter warp % 4 -> dW rows, one vector reduction per 4 columns;
    // out = scale[n] * acc[m][n] + colsum[m] * shift[n] when the operand was a LayerNorm's xhat (see the file header)
    tc::mbar_wait(acc_bar, 0);
    tc::fence_after_sync();
    const int q = warp & 3;
    const int m = q * 32 + lane;
    const uint32_t t_lane = tmem + ((uint32_t)(q * 32) << 16);
    float* o = s.dW + (int64_t)m * s.ldw;
    float colsum = 0.f;
    if (s.db || s.shift) {
      float v[32];
      tc::tmem_ld32(t_lane + 128, v);
      tc::tmem_ld_wait();
      colsum = v[0];
      if (s.db) atomicAdd(s.db + m, colsum);
    }
#pragma unroll 1
    for (int c0 = 0; c0 < 128; c0 += 32) {
      float v[32];
      tc::tmem_ld32(t_lane + c0, v);
      tc::tmem_ld_wait();
      if (s.scale) {
#pragma unroll
        for (int c = 0; c < 32; ++c) v[c] = fmaf(v[c], __ldg(s.scale + c0 + c), colsum * __ldg(s.shift + c0 + c));
      }
#pragma unroll
      for (int c = 0; c < 32; c += 4)
        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(o + c0 + c), "f"(v[c]), "f"(v[c + 1]),
                     "f"(v[c + 2]), "f"(v[c + 3])
                     : "memory");
    }
    tc::fence_before_sync();
  }
  __syncthreads();
  if (warp == 1) {
    tc::fence_after_sync();
    tc::tmem_free(tmem, 256);
  }
}

// out[i, :] = bf16(pos_table[tok_cell[i], :]) — the position rows of a token set and shift, gathered once per step
__global__ void __launch_bounds__(256) k_pos_rows(const float* __restrict__ table, const int32_t* __restrict__ cell, int n,
                                                  __nv_bfloat16* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  for (int i = blockIdx.x * 8 + (threadIdx.x >> 5); i < n; i += gridDim.x * 8) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(table + (int64_t)__ldg(cell + i) * 128) + lane);
    const __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
    reinterpret_cast<uint2*>(out + (int64_t)i * 128)[lane] =
        make_uint2(*reinterpret_cast<const uint32_t*>(&a), *reinterpret_cast<const uint32_t*>(&b));
  }
}

}  // namespace

extern "C" int geomae_pos_rows_bf16(const float* pos_table, const int32_t* tok_cell, int64_t n_tokens, void* out16, void* stream) {
  GM_REQUIRE(n_tokens >= 0 && n_tokens < ((int64_t)1 << 31), "pos_rows_bf16: bad token count");
  if (n_tokens == 0) return GEOMAE_OK;
  GM_REQUIRE(pos_table && tok_cell && out16, "pos_rows_bf16: null argument");
  int blocks = gm_div_up(n_tokens, 8);
  if (blocks > GM_NUM_SMS * 8) blocks = GM_NUM_SMS * 8;
  k_pos_rows<<<blocks, 256, 0, (cudaStream_t)stream>>>(pos_table, tok_cell, (int)n_tokens, (__nv_bfloat16*)out16);
  GM_LAUNCH_CHECK();
  return GEOMAE_OK;
}

extern "C" int geomae_sra_wgrad_layer(const geomae_wgrad_layer_args* p, void* stream) {
  GM_REQUIRE(p, "sra_wgrad_layer: null argument");
  GM_REQUIRE(p->n_tokens >= 0 && p->n_tokens < ((int64_t)1 << 31) - 256, "sra_wgrad_layer: bad token count");
  if (p->n_tokens == 0) return GEOMAE_OK;
  GM_REQUIRE(p->ds2_16 && p->g16 && p->du16 && p->xh1_16 && p->ds1_16 && p->attn16 && p->dqkv16 && p->xin16 && p->pos16,
             "sra_wgrad_layer: null operand tensor");
  GM_REQUIRE(p->g_lin2_w && p->g_lin1_w && p->g_out_proj_w && p->g_in_proj_w && p->norm1_w && p->norm1_b,
             "sra_wgrad_layer: null gradient buffer / LayerNorm-1 parameters");
  GM_REQUIRE((p->in_scale == nullptr) == (p->in_shift == nullptr), "sra_wgrad_layer: in_scale and in_shift go together");
  const int64_t n = p->n_tokens;
  WgMaps maps;
  const void* base[9] = {p->ds2_16, p->g16, p->du16, p->xh1_16, p->ds1_16, p->attn16, p->dqkv16, p->xin16, p->pos16};
  const int cols[9] = {128, 256, 256, 128, 128, 128, 384, 128, 128};
  for (int i = 0; i < 9; ++i) {
    const int rc = tma::make_map(&maps.m[i], base[i], n, cols[i], 2);
    if (rc) return rc;
  }
  WgArgs a;
  a.n_tiles = gm_div_up(n, WT);
  int splits = GM_NUM_SMS / MAX_SLABS;                  // 14 token ranges x 10 slabs = 140 CTAs
  if (splits > a.n_tiles) splits = a.n_tiles;
  a.tiles_per_cta = gm_div_up(a.n_tiles, splits);
  splits = gm_div_up(a.n_tiles, a.tiles_per_cta);
  float* bi = p->g_in_proj_b;
  // slab = [128 dY columns] x [128 X columns]:  dY map, col, X map, col, dW, ld, db, scale, shift
  a.slab[0] = Slab{0, 0, 1, 0, p->g_lin2_w, 256, p->g_lin2_b, nullptr, nullptr};              // dW2[:, 0:128] = ds2^T g[:, 0:128]
  a.slab[1] = Slab{0, 0, 1, 128, p->g_lin2_w + 128, 256, nullptr, nullptr, nullptr};          // dW2[:, 128:256]
  a.slab[2] = Slab{2, 0, 3, 0, p->g_lin1_w, 128, p->g_lin1_b, p->norm1_w, p->norm1_b};        // dW1[0:128] = du[:, 0:128]^T y
  a.slab[3] = Slab{2, 128, 3, 0, p->g_lin1_w + 128 * 128, 128, p->g_lin1_b ? p->g_lin1_b + 128 : nullptr, p->norm1_w, p->norm1_b};
  a.slab[4] = Slab{4, 0, 5, 0, p->g_out_proj_w, 128, p->g_out_proj_b, nullptr, nullptr};      // dWo = ds1^T O
  a.slab[5] = Slab{6, 0, 7, 0, p->g_in_proj_w, 128, bi, p->in_scale, p->in_shift};            // dWin[q rows] = dq^T x
  a.slab[6] = Slab{6, 128, 7, 0, p->g_in_proj_w + 128 * 128, 128, bi ? bi + 128 : nullptr, p->in_scale, p->in_shift};
  a.slab[7] = Slab{6, 256, 7, 0, p->g_in_proj_w + 256 * 128, 128, bi ? bi + 256 : nullptr, p->in_scale, p->in_shift};
  a.slab[8] = Slab{6, 0, 8, 0, p->g_in_proj_w, 128, nullptr, nullptr, nullptr};               // += dq^T pos[cell]
  a.slab[9] = Slab{6, 128, 8, 0, p->g_in_proj_w + 128 * 128, 128, nullptr, nullptr, nullptr}; // += dk^T pos[cell]
  static bool configured = false;
  if (!configured) {
    GM_CUDA(cudaFuncSetAttribute(k_wgrad_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    configured = true;
  }
  GM_CUDA(gm_launch_pdl(k_wgrad_tma, dim3(splits, MAX_SLABS), dim3(NTHR), (size_t)SMEM_BYTES, (cudaStream_t)stream, maps, a));
  return GEOMAE_OK;
}
