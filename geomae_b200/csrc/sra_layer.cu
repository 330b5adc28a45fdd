// Dense layers of the SRA block on the 5th-gen tensor cores:
// token-tile GEMMs with fused prologues (position add, GELU) and epilogues (bias, residual +
// LayerNorm, GELU-gradient), and the weight-gradient GEMM, all with tcgen05.mma accumulating in TMEM.
//
// One CTA = 128 tokens x NT outputs.  Operands are converted fp32 -> bf16 while being staged into the
// swizzled shared-memory layout of tc_common.cuh; `precision 3` additionally stages the bf16 residuals
// and issues hi*hi + hi*lo + lo*hi (fp32-equivalent to ~2^-16, the parity mode), `precision 1` is plain
// bf16.  The accumulator row of a token lives in one TMEM lane = one thread of the epilogue, so the
// LayerNorm statistics need no cross-thread reduction at all.
#include <stdlib.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace {

constexpr int TM = 128;       // tokens per CTA
constexpr int KC = 128;       // K elements staged per chunk
constexpr int NTHREADS = 256;    // k_tc_wgrad: 8 warps stage; warps w and w+4 share TMEM lane quarter w%4 and split the columns

struct LinArgs {
  const float* A; int lda; int n_rows; int K;
  const float* pos_table; const int32_t* tok_cell; int pos_slabs;   // A += pos_table[tok_cell[row]] for slabs < pos_slabs
  int a_gelu;                                                       // A = gelu(A)
  const float* W; int ldw; int w_rows; int w_mn_major;              // 0: W[n][k] (y = x W^T), 1: W[k][n] (dX = dY W)
  const uint8_t* Wp_hi; const uint8_t* Wp_lo; int wp_cols;          // optional pre-packed bf16 images of W (geomae_pack_weights)
  const float* bias; int N_total;
  float* out; int ldo;
  const float* add_src; int ld_add;                                 // out += add_src (residual / gradient sum)
  const float* ln_gamma; const float* ln_beta; float ln_eps;        // EPI 1
  float* ln_in; float* ln_stats;                                    // EPI 1: saved pre-LN rows [n,N] and (mean, rstd) [n,2]
  const float* gelu_u; int ldu;                                     // EPI 2: out = acc * gelu'(u)
  int precision;
  int w_early;      // packed weights may be fetched before gm_pdl_wait() (see common.cuh)
  const float* dot_src; int ld_dot; float* dot_out;   // EPI 0, N = 128: dot_out[row, h] = sum_head out * dot_src
  float* ln_dgamma; float* ln_dbeta; float* ln_dcolsum;   // EPI 3 accumulators (+=)
  int a_bf16;       // A rows are bf16 (lda in elements): staged with plain 16-byte copies, precision 1 only
  int out_bf16;     // EPI 0 / 2: `out` rows are written as bf16 (ldo in elements)
};

// Stage a [ROWS x COLS] fp32 tile (global rows row0.., columns col0..) as bf16 into swizzled 64-column blocks.
// Loads are issued UNROLL items ahead of the conversions so one thread keeps 2*UNROLL 128-bit loads in
// flight (the un-pipelined version spent ~20 us per tile waiting on one load at a time).
template <int ROWS, int COLS, int NTHR = NTHREADS, bool POS = true>
__device__ __forceinline__ void stage_tile(uint8_t* dst_hi, uint8_t* dst_lo, const float* __restrict__ src, int ld,
                                           int row0, int row_end, int col0, const float* __restrict__ pos_table_,
                                           const int32_t* __restrict__ tok_cell, int pos_ld, bool do_gelu) {
  constexpr int CPR = COLS / 8;                 // 16-byte chunks per row
  constexpr int BLOCK_BYTES = ROWS * tc::LINE_BYTES;
  constexpr int ITEMS = ROWS * CPR / NTHR;      // per thread
  constexpr int UNROLL = ITEMS < (POS ? 4 : 8) ? ITEMS : (POS ? 4 : 8);   // the position prologue doubles the loads per item
  const float* pos_table = POS ? pos_table_ : nullptr;
  static_assert(ITEMS % UNROLL == 0, "tile size must be a multiple of the staging unroll");
#pragma unroll 1
  for (int it = 0; it < ITEMS; it += UNROLL) {
    float4 a[UNROLL], b[UNROLL], c[UNROLL], d[UNROLL];
#pragma unroll
    for (int k = 0; k < UNROLL; ++k) {
      const int i = (it + k) * NTHR + threadIdx.x;
      const int r = i / CPR, c8 = i % CPR;
      const int grow = row0 + r;
      if (grow < row_end) {
        const float4* p = reinterpret_cast<const float4*>(src + (int64_t)grow * ld + col0 + c8 * 8);
        a[k] = __ldg(p);
        b[k] = __ldg(p + 1);
        if (pos_table) {
          const float4* q = reinterpret_cast<const float4*>(pos_table + (int64_t)__ldg(tok_cell + grow) * pos_ld + col0 + c8 * 8);
          c[k] = __ldg(q);
          d[k] = __ldg(q + 1);
        }
      } else {
        a[k] = b[k] = make_float4(0.f, 0.f, 0.f, 0.f);
        c[k] = d[k] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
#pragma unroll
    for (int k = 0; k < UNROLL; ++k) {
      const int i = (it + k) * NTHR + threadIdx.x;
      const int r = i / CPR, c8 = i % CPR;
      float f[8] = {a[k].x, a[k].y, a[k].z, a[k].w, b[k].x, b[k].y, b[k].z, b[k].w};
      if (pos_table && row0 + r < row_end) {
        f[0] += c[k].x; f[1] += c[k].y; f[2] += c[k].z; f[3] += c[k].w;
        f[4] += d[k].x; f[5] += d[k].y; f[6] += d[k].z; f[7] += d[k].w;
      }
      if (do_gelu) {
#pragma unroll
        for (int e = 0; e < 8; ++e) f[e] = gelu_f(f[e]);
      }
      uint4 lo;
      const uint4 hi = tc::pack8(f, dst_lo ? &lo : nullptr);
      const uint32_t off = (uint32_t)(c8 >> 3) * BLOCK_BYTES + tc::swz(r, c8 & 7);
      *reinterpret_cast<uint4*>(dst_hi + off) = hi;
      if (dst_lo) *reinterpret_cast<uint4*>(dst_lo + off) = lo;
    }
  }
}

// The same tile from bf16 rows (a tensor that only ever serves as an MMA operand is stored in bf16 by its producer:
// numerically identical to rounding at staging time, half the bytes, no conversion): 16-byte chunks straight into
// the swizzled layout.
template <int ROWS, int COLS, int NTHR>
__device__ __forceinline__ void stage_tile_bf16(uint8_t* dst, const __nv_bfloat16* __restrict__ src, int ld, int row0,
                                                int row_end, int col0) {
  constexpr int CPR = COLS / 8;
  constexpr int BLOCK_BYTES = ROWS * tc::LINE_BYTES;
  constexpr int ITEMS = ROWS * CPR / NTHR;
  uint4 v[ITEMS];
#pragma unroll
  for (int k = 0; k < ITEMS; ++k) {
    const int i = k * NTHR + threadIdx.x;
    const int r = i / CPR, c8 = i % CPR;
    const int grow = row0 + r;
    v[k] = grow < row_end ? __ldg(reinterpret_cast<const uint4*>(src + (int64_t)grow * ld + col0 + c8 * 8))
                          : make_uint4(0u, 0u, 0u, 0u);
  }
#pragma unroll
  for (int k = 0; k < ITEMS; ++k) {
    const int i = k * NTHR + threadIdx.x;
    const int r = i / CPR, c8 = i % CPR;
    *reinterpret_cast<uint4*>(dst + (uint32_t)(c8 >> 3) * BLOCK_BYTES + tc::swz(r, c8 & 7)) = v[k];
  }
}

constexpr int LTHREADS = 256;  // k_tc_linear: 8 warps stage; warps w and w+4 share TMEM lane quarter w%4 and split the columns

// fp32 accumulator tile in shared memory, [128 rows][C4 float4], XOR-swizzled so that both the row-per-thread
// TMEM drain and the row-contiguous (coalesced) read-back are bank-conflict free
template <int C4>
__device__ __forceinline__ float4* ctile(float* base, int row, int c4) {
  return reinterpret_cast<float4*>(base) + row * C4 + (c4 ^ (row & (C4 - 1)));
}

// EPI: 0 = acc (+bias) (+add_src); 1 = LayerNorm(acc + bias + add_src); 2 = acc * gelu'(u)
// TMR = token rows per CTA = the M of the MMA.  128: accumulator row r lives in TMEM lane r.  64: the accumulator uses
// 16 lanes of each 32-lane quarter (row r -> lane 32*(r/16) + r%16), so twice as many, half as long CTAs cover the
// token set — these kernels are latency-bound per CTA, and the encoder has only 57 tiles of 128 tokens for 148 SMs.
template <int NT, int EPI, bool POS, int TMR>
__global__ void __launch_bounds__(LTHREADS, 2) k_tc_linear(const LinArgs a) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t mbar;
  __shared__ __align__(8) uint64_t mbar_w;      // weight-image bulk copies
  __shared__ uint32_t tmem_slot;
  uint8_t* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
  constexpr int TM = TMR;                       // shadows the file-level 128 inside this kernel
  constexpr int A_BYTES = TM * KC * 2;          // 32 KB (16 KB for 64-row tiles)
  constexpr int B_BYTES = NT * KC * 2;
  constexpr int BLOCK16K = 128 * tc::LINE_BYTES;   // one packed [128 x 64] bf16 block
  const bool x3 = a.precision == 3;
  const bool packed = a.Wp_hi != nullptr;
  uint8_t* sA = smem;
  uint8_t* sAlo = sA + A_BYTES;
  uint8_t* sB = sA + (x3 ? 2 : 1) * A_BYTES;
  uint8_t* sBlo = sB + B_BYTES;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row0 = blockIdx.x * TM;
  const int n0 = blockIdx.y * NT;
  if (warp == 0) tc::tmem_alloc(&tmem_slot, NT);
  if (threadIdx.x == 0) { tc::mbar_init(&mbar, 1); tc::mbar_init(&mbar_w, 1); }
  const bool use_pos = a.pos_table && (int)blockIdx.y < a.pos_slabs;
  const int n_chunks = a.K / KC;
  const int ncb = a.wp_cols / 64;       // 64-column blocks per 128-row band of the packed image
  // weights: bulk copies of the pre-swizzled bf16 image of K-chunk kc, no thread work
  auto fetch_weights = [&](int kc) {
    const int k0 = kc * KC;
    tc::mbar_expect_tx(&mbar_w, (uint32_t)((x3 ? 2 : 1) * B_BYTES));
    for (int part = 0; part < (x3 ? 2 : 1); ++part) {
      const uint8_t* img = part ? a.Wp_lo : a.Wp_hi;
      uint8_t* dst = part ? sBlo : sB;
      if (a.w_mn_major) {             // band k0/128, blocks n0/64 .. : contiguous
        tc::bulk_g2s(dst, img + ((size_t)(k0 / 128) * ncb + n0 / 64) * BLOCK16K, B_BYTES, &mbar_w);
      } else {                        // bands n0/128 + r, blocks k0/64, k0/64+1 : 32 KB each
        for (int r = 0; r < NT / 128; ++r)
          tc::bulk_g2s(dst + r * 2 * BLOCK16K, img + ((size_t)(n0 / 128 + r) * ncb + k0 / 64) * BLOCK16K,
                       2 * BLOCK16K, &mbar_w);
      }
    }
  };
  const bool w_early = packed && a.w_early;
  if (w_early && threadIdx.x == 0) fetch_weights(0);   // overlaps the predecessor kernel's tail (PDL)
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = tmem_slot;
  gm_pdl_wait();                        // everything below reads what earlier kernels wrote
  gm_pdl_trigger();
  for (int kc = 0; kc < n_chunks; ++kc) {
    const int k0 = kc * KC;
    if (kc > 0) {                       // operands of the previous chunk must be consumed before overwriting
      tc::mbar_wait(&mbar, (kc - 1) & 1);
      tc::fence_after_sync();
    }
    if (packed && threadIdx.x == 0 && !(w_early && kc == 0)) fetch_weights(kc);
    if (!POS && a.a_bf16)
      stage_tile_bf16<TM, KC, LTHREADS>(sA, reinterpret_cast<const __nv_bfloat16*>(a.A), a.lda, row0, a.n_rows, k0);
    else
      stage_tile<TM, KC, LTHREADS, POS>(sA, x3 ? sAlo : nullptr, a.A, a.lda, row0, a.n_rows, k0,
                                        use_pos ? a.pos_table : nullptr, a.tok_cell, a.K, a.a_gelu != 0);
    if (!packed) {
      if (a.w_mn_major)   // rows = k, columns = n
        stage_tile<KC, NT, LTHREADS, false>(sB, x3 ? sBlo : nullptr, a.W, a.ldw, k0, a.w_rows, n0, nullptr, nullptr, 0, false);
      else                // rows = n, columns = k
        stage_tile<NT, KC, LTHREADS, false>(sB, x3 ? sBlo : nullptr, a.W, a.ldw, n0, a.w_rows, k0, nullptr, nullptr, 0, false);
    }
    tc::fence_async_smem();
    tc::fence_before_sync();
    __syncthreads();
    if (threadIdx.x == 0) {
      if (packed) tc::mbar_wait(&mbar_w, kc & 1);
      tc::fence_after_sync();
      // packed K-major images are organised in 128-row bands: one N=128 MMA per band
      const int n_sub = (packed && !a.w_mn_major) ? NT / 128 : 1;
      const int n_mma = NT / n_sub;
      const uint32_t idesc = tc::make_idesc_bf16(TM, n_mma, 0, a.w_mn_major);
      const uint32_t a_hi = tc::smem_u32(sA), a_lo = tc::smem_u32(sAlo), b_hi = tc::smem_u32(sB), b_lo = tc::smem_u32(sBlo);
      const bool acc0 = kc > 0;
#pragma unroll
      for (int j = 0; j < KC / 16; ++j) {
        const uint32_t a_off = (uint32_t)(j >> 2) * (TM * tc::LINE_BYTES) + (uint32_t)(j & 3) * 32;
        const uint64_t da_hi = tc::make_desc(a_hi + a_off, 16, tc::ATOM_BYTES);
        const uint64_t da_lo = tc::make_desc(a_lo + a_off, 16, tc::ATOM_BYTES);
        for (int sub = 0; sub < n_sub; ++sub) {
          uint32_t b_off, b_lbo;
          if (a.w_mn_major) { b_off = (uint32_t)j * 2 * tc::ATOM_BYTES; b_lbo = KC * tc::LINE_BYTES; }
          else if (packed) { b_off = (uint32_t)sub * 2 * BLOCK16K + (uint32_t)(j >> 2) * BLOCK16K + (uint32_t)(j & 3) * 32; b_lbo = 16; }
          else { b_off = (uint32_t)(j >> 2) * (NT * tc::LINE_BYTES) + (uint32_t)(j & 3) * 32; b_lbo = 16; }
          const uint64_t db_hi = tc::make_desc(b_hi + b_off, b_lbo, tc::ATOM_BYTES);
          const uint32_t td = tmem + sub * 128;
          const bool acc = acc0 || j > 0;
          tc::mma_bf16(td, da_hi, db_hi, idesc, acc);
          if (x3) {
            const uint64_t db_lo = tc::make_desc(b_lo + b_off, b_lbo, tc::ATOM_BYTES);
            tc::mma_bf16(td, da_hi, db_lo, idesc, true);
            tc::mma_bf16(td, da_lo, db_hi, idesc, true);
          }
        }
      }
      tc::mma_commit(&mbar);
    }
  }
  // ---- epilogue.  The accumulator row of a token is one TMEM lane; it is drained row-per-thread into a
  // swizzled fp32 tile in (now free) operand shared memory and read back row-contiguously, so every global
  // access of the epilogue (bias, residual, GELU input, outputs) is a coalesced 128-bit access.  The epilogue's
  // global operands are PREFETCHED into registers before waiting on the MMA barrier (and panel p+1's while panel
  // p is written), so their L2/DRAM latency overlaps the tensor-core work instead of serialising behind it.
  float* sC = reinterpret_cast<float*>(smem);
  const int q = warp & 3, hsel = warp >> 2;                    // TMEM lane quarter, column half
  constexpr int LPQ = TM / 4;                                  // accumulator lanes in use per quarter (32 or 16)
  const int trow = q * LPQ + lane;                             // tile row drained by this thread (if lane < LPQ)
  const bool drains = lane < LPQ;
  const uint32_t t_lane = tmem + ((uint32_t)(q * 32) << 16);
  if constexpr (EPI == 1) {
    static_assert(EPI != 1 || NT == 128, "LayerNorm epilogue needs the whole row in one CTA");
    constexpr int C4 = 32;                                     // 128 columns
    constexpr int RPW = TM / (LTHREADS / 32);                  // rows per warp (16 or 8), handled in batches of 8
    constexpr int NB = RPW / 8;
    const float4 b4 = __ldg(reinterpret_cast<const float4*>(a.bias) + lane);
    const float4 g4 = __ldg(reinterpret_cast<const float4*>(a.ln_gamma) + lane);
    const float4 be4 = __ldg(reinterpret_cast<const float4*>(a.ln_beta) + lane);
    float4 res[8];
    auto fetch = [&](int batch) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int row = row0 + warp + (LTHREADS / 32) * (batch * 8 + j);
        res[j] = row < a.n_rows ? __ldg(reinterpret_cast<const float4*>(a.add_src + (int64_t)row * a.ld_add) + lane)
                                : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    };
    fetch(0);
    tc::mbar_wait(&mbar, (n_chunks - 1) & 1);
    tc::fence_after_sync();
#pragma unroll
    for (int half = 0; half < 2; ++half) {                     // this warp: columns hsel*64 + half*32 ..
      float v[32];
      const int c0 = hsel * 64 + half * 32;
      tc::tmem_ld32(t_lane + c0, v);
      tc::tmem_ld_wait();
      if (drains) {
#pragma unroll
        for (int c = 0; c < 32; c += 4) *ctile<C4>(sC, trow, (c0 + c) >> 2) = make_float4(v[c], v[c + 1], v[c + 2], v[c + 3]);
      }
    }
    __syncthreads();
    // one warp per row, one float4 per lane: s = acc + bias + residual; exact two-pass statistics; write s, LN(s)
#pragma unroll
    for (int batch = 0; batch < NB; ++batch) {
      // eight rows per batch, every step written for all eight at once so their shuffle reductions interleave
      constexpr int RB = 8;
      float4 v[RB];
      float mean[RB], rstd[RB];
#pragma unroll
      for (int j = 0; j < RB; ++j) {
        const int r = warp + (LTHREADS / 32) * (batch * RB + j);
        v[j] = *ctile<C4>(sC, r, lane);
        v[j].x += b4.x + res[j].x; v[j].y += b4.y + res[j].y; v[j].z += b4.z + res[j].z; v[j].w += b4.w + res[j].w;
      }
      if (batch + 1 < NB) fetch(batch + 1);
#pragma unroll
      for (int j = 0; j < RB; ++j) mean[j] = (v[j].x + v[j].y) + (v[j].z + v[j].w);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1)
#pragma unroll
        for (int j = 0; j < RB; ++j) mean[j] += __shfl_xor_sync(0xffffffffu, mean[j], o);
#pragma unroll
      for (int j = 0; j < RB; ++j) {
        mean[j] *= (1.0f / NT);
        const float dx = v[j].x - mean[j], dy = v[j].y - mean[j], dz = v[j].z - mean[j], dw = v[j].w - mean[j];
        rstd[j] = (dx * dx + dy * dy) + (dz * dz + dw * dw);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1)
#pragma unroll
        for (int j = 0; j < RB; ++j) rstd[j] += __shfl_xor_sync(0xffffffffu, rstd[j], o);
#pragma unroll
      for (int j = 0; j < RB; ++j) {
        const int row = row0 + warp + (LTHREADS / 32) * (batch * RB + j);
        if (row < a.n_rows) {
          const float rs = rsqrtf(rstd[j] * (1.0f / NT) + a.ln_eps);
          const float dx = v[j].x - mean[j], dy = v[j].y - mean[j], dz = v[j].z - mean[j], dw = v[j].w - mean[j];
          if (a.ln_in) reinterpret_cast<float4*>(a.ln_in + (int64_t)row * NT)[lane] = v[j];
          if (a.ln_stats && lane == 0) *reinterpret_cast<float2*>(a.ln_stats + 2 * (int64_t)row) = make_float2(mean[j], rs);
          reinterpret_cast<float4*>(a.out + (int64_t)row * a.ldo)[lane] =
              make_float4(dx * rs * g4.x + be4.x, dy * rs * g4.y + be4.y, dz * rs * g4.z + be4.z, dw * rs * g4.w + be4.w);
        }
      }
    }
  } else if constexpr (EPI == 3) {
    // LayerNorm BACKWARD as the epilogue: the GEMM result (+ add_src) is dz, the gradient reaching a LayerNorm output;
    // with the saved pre-LN rows s = ln_in, (mean, rstd) = ln_stats and gamma this writes
    //   ds = rstd (g - mean_c(g) - xhat mean_c(g xhat)),  g = dz gamma,  xhat = (s - mean) rstd
    // to `out` and accumulates d_gamma += sum dz xhat, d_beta += sum dz, d_colsum += sum ds (the bias gradient of the
    // linear layer in front of the LayerNorm).  dz itself never reaches memory.
    static_assert(EPI != 3 || NT == 128, "LayerNorm-backward epilogue needs the whole row in one CTA");
    constexpr int C4 = 32;
    constexpr int RPW = TM / (LTHREADS / 32);                  // rows per warp (16 or 8)
    constexpr int RB = 4, NB = RPW / RB;                       // rows per batch
    const float4 g4 = __ldg(reinterpret_cast<const float4*>(a.ln_gamma) + lane);
    float4 res[RB], sx[RB];
    float2 st[RB];
    auto fetch = [&](int batch) {
#pragma unroll
      for (int j = 0; j < RB; ++j) {
        const int row = row0 + warp + (LTHREADS / 32) * (batch * RB + j);
        const bool ok = row < a.n_rows;
        res[j] = (ok && a.add_src) ? __ldg(reinterpret_cast<const float4*>(a.add_src + (int64_t)row * a.ld_add) + lane)
                                   : make_float4(0.f, 0.f, 0.f, 0.f);
        sx[j] = ok ? __ldg(reinterpret_cast<const float4*>(a.ln_in + (int64_t)row * NT) + lane) : make_float4(0.f, 0.f, 0.f, 0.f);
        st[j] = ok ? __ldg(reinterpret_cast<const float2*>(a.ln_stats + 2 * (int64_t)row)) : make_float2(0.f, 0.f);
      }
    };
    fetch(0);
    tc::mbar_wait(&mbar, (n_chunks - 1) & 1);
    tc::fence_after_sync();
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      float v[32];
      const int c0 = hsel * 64 + half * 32;
      tc::tmem_ld32(t_lane + c0, v);
      tc::tmem_ld_wait();
      if (drains) {
#pragma unroll
        for (int c = 0; c < 32; c += 4) *ctile<C4>(sC, trow, (c0 + c) >> 2) = make_float4(v[c], v[c + 1], v[c + 2], v[c + 3]);
      }
    }
    __syncthreads();
    float4 accg = make_float4(0.f, 0.f, 0.f, 0.f), accb = accg, accs = accg;
#pragma unroll 1
    for (int batch = 0; batch < NB; ++batch) {
      float4 dz[RB], xh[RB];
      float p1[RB], p2[RB], rs[RB];
#pragma unroll
      for (int j = 0; j < RB; ++j) {
        const int r = warp + (LTHREADS / 32) * (batch * RB + j);
        dz[j] = *ctile<C4>(sC, r, lane);
        dz[j].x += res[j].x; dz[j].y += res[j].y; dz[j].z += res[j].z; dz[j].w += res[j].w;
        rs[j] = st[j].y;
        xh[j] = make_float4((sx[j].x - st[j].x) * rs[j], (sx[j].y - st[j].x) * rs[j], (sx[j].z - st[j].x) * rs[j],
                            (sx[j].w - st[j].x) * rs[j]);
      }
      if (batch + 1 < NB) fetch(batch + 1);
#pragma unroll
      for (int j = 0; j < RB; ++j) {
        const float gx = dz[j].x * g4.x, gy = dz[j].y * g4.y, gz = dz[j].z * g4.z, gw = dz[j].w * g4.w;
        p1[j] = (gx + gy) + (gz + gw);
        p2[j] = (gx * xh[j].x + gy * xh[j].y) + (gz * xh[j].z + gw * xh[j].w);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1)
#pragma unroll
        for (int j = 0; j < RB; ++j) {
          p1[j] += __shfl_xor_sync(0xffffffffu, p1[j], o);
          p2[j] += __shfl_xor_sync(0xffffffffu, p2[j], o);
        }
#pragma unroll
      for (int j = 0; j < RB; ++j) {
        const int row = row0 + warp + (LTHREADS / 32) * (batch * RB + j);
        if (row < a.n_rows) {
          const float m1 = p1[j] * (1.0f / NT), m2 = p2[j] * (1.0f / NT);
          const float4 o = make_float4(rs[j] * (dz[j].x * g4.x - m1 - xh[j].x * m2), rs[j] * (dz[j].y * g4.y - m1 - xh[j].y * m2),
                                       rs[j] * (dz[j].z * g4.z - m1 - xh[j].z * m2), rs[j] * (dz[j].w * g4.w - m1 - xh[j].w * m2));
          reinterpret_cast<float4*>(a.out + (int64_t)row * a.ldo)[lane] = o;
          accg.x += dz[j].x * xh[j].x; accg.y += dz[j].y * xh[j].y; accg.z += dz[j].z * xh[j].z; accg.w += dz[j].w * xh[j].w;
          accb.x += dz[j].x; accb.y += dz[j].y; accb.z += dz[j].z; accb.w += dz[j].w;
          accs.x += o.x; accs.y += o.y; accs.z += o.z; accs.w += o.w;
        }
      }
    }
    __syncthreads();                                           // the accumulator tile is dead: reuse it for the column sums
    float* part = sC;                                          // [8 warps][3][128]
    reinterpret_cast<float4*>(part + (warp * 3 + 0) * NT)[lane] = accg;
    reinterpret_cast<float4*>(part + (warp * 3 + 1) * NT)[lane] = accb;
    reinterpret_cast<float4*>(part + (warp * 3 + 2) * NT)[lane] = accs;
    __syncthreads();
    for (int i = threadIdx.x; i < 3 * NT; i += LTHREADS) {
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < LTHREADS / 32; ++w) t += part[w * 3 * NT + i];
      float* dst = i < NT ? a.ln_dgamma : (i < 2 * NT ? a.ln_dbeta : a.ln_dcolsum);
      if (dst) atomicAdd(dst + (i % NT), t);
    }
  } else {
    constexpr int C4 = 16;                                     // 64-column panels
    constexpr int NP = NT / 64;                                // panels
    constexpr int RI = TM * C4 / LTHREADS;                     // rows per thread and panel (8)
    const int c4 = threadIdx.x % C4, rbase = threadIdx.x / C4; // this thread: column quad c4 of rows rbase + 16 k
    const bool dot_mode = EPI == 0 && NT == 128 && !POS && a.dot_src != nullptr;   // `pre` then carries dot_src, not add_src
    const float* esrc = EPI == 2 ? a.gelu_u : (dot_mode ? a.dot_src : a.add_src);  // per-element global operand (may be null)
    const int eld = EPI == 2 ? a.ldu : (dot_mode ? a.ld_dot : a.ld_add);
    float4 pre[RI], bias4;
    auto fetch = [&](int pc) {
      bias4 = (EPI == 0 && a.bias) ? __ldg(reinterpret_cast<const float4*>(a.bias + n0 + pc * 64 + c4 * 4))
                                   : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int k = 0; k < RI; ++k) {
        const int row = row0 + rbase + (LTHREADS / C4) * k;
        pre[k] = (esrc && row < a.n_rows)
                     ? __ldg(reinterpret_cast<const float4*>(esrc + (int64_t)row * eld + n0 + pc * 64 + c4 * 4))
                     : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    };
    fetch(0);
    tc::mbar_wait(&mbar, (n_chunks - 1) & 1);
    tc::fence_after_sync();
    float4 dotv[(EPI == 0 && NT == 128 && !POS) ? RI : 1];
#pragma unroll 1
    for (int pc = 0; pc < NP; ++pc) {
      float v[32];
      tc::tmem_ld32(t_lane + pc * 64 + hsel * 32, v);
      tc::tmem_ld_wait();
      if (pc > 0) __syncthreads();                             // previous panel fully read back
      if (drains) {
#pragma unroll
        for (int c = 0; c < 32; c += 4) *ctile<C4>(sC, trow, (hsel * 32 + c) >> 2) = make_float4(v[c], v[c + 1], v[c + 2], v[c + 3]);
      }
      __syncthreads();
#pragma unroll
      for (int k = 0; k < RI; ++k) {
        const int r = rbase + (LTHREADS / C4) * k;
        const int row = row0 + r;
        if (row < a.n_rows) {
          float4 o = *ctile<C4>(sC, r, c4);
          if constexpr (EPI == 2) {
            const float4 u = pre[k];
            o.x *= gelu_grad_f(u.x); o.y *= gelu_grad_f(u.y); o.z *= gelu_grad_f(u.z); o.w *= gelu_grad_f(u.w);
          } else {
            if (dot_mode) { o.x += bias4.x; o.y += bias4.y; o.z += bias4.z; o.w += bias4.w; }
            else {
              o.x += bias4.x + pre[k].x; o.y += bias4.y + pre[k].y;
              o.z += bias4.z + pre[k].z; o.w += bias4.w + pre[k].w;
            }
          }
          if (a.out_bf16) {
            __nv_bfloat16* o16 = reinterpret_cast<__nv_bfloat16*>(a.out) + (int64_t)row * a.ldo + n0 + pc * 64 + c4 * 4;
            const __nv_bfloat162 lo2 = __floats2bfloat162_rn(o.x, o.y), hi2 = __floats2bfloat162_rn(o.z, o.w);
            *reinterpret_cast<uint2*>(o16) = make_uint2(*reinterpret_cast<const uint32_t*>(&lo2), *reinterpret_cast<const uint32_t*>(&hi2));
          } else {
            *reinterpret_cast<float4*>(a.out + (int64_t)row * a.ldo + n0 + pc * 64 + c4 * 4) = o;
          }
          if constexpr (EPI == 0 && NT == 128 && !POS) dotv[k] = o;
        }
      }
      if constexpr (EPI == 0 && NT == 128 && !POS) {
        if (a.dot_src) {                 // per-(row, head) dot with dot_src: 4 adjacent lanes hold one head's 16 columns
#pragma unroll
          for (int k = 0; k < RI; ++k) {
            const int row = row0 + rbase + (LTHREADS / C4) * k;
            float part = 0.f;
            if (row < a.n_rows) {
              const float4 s4 = pre[k];       // dot_src, prefetched with the panel
              part = (dotv[k].x * s4.x + dotv[k].y * s4.y) + (dotv[k].z * s4.z + dotv[k].w * s4.w);
            }
            part += __shfl_xor_sync(0xffffffffu, part, 1);
            part += __shfl_xor_sync(0xffffffffu, part, 2);
            if (row < a.n_rows && (c4 & 3) == 0) a.dot_out[(int64_t)row * 8 + ((n0 + pc * 64) >> 4) + (c4 >> 2)] = part;
          }
        }
      }
      if (pc + 1 < NP) fetch(pc + 1);          // in flight while the next panel is drained from TMEM
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_free(tmem, NT);
}

template <int NT, int EPI, bool POS, int TMR>
int launch_linear_impl(const LinArgs& a, cudaStream_t stream) {
  constexpr int per_prec = TMR * KC * 2 + NT * KC * 2;
  constexpr int epi_bytes = TMR * ((EPI == 1 || EPI == 3) ? 128 : 64) * 4;          // fp32 staging tile of the epilogue
  const int operands = (a.precision == 3 ? 2 : 1) * per_prec;
  const int smem = (operands > epi_bytes ? operands : epi_bytes) + 1024;
  static bool configured = false;
  if (!configured) {
    GM_CUDA(cudaFuncSetAttribute(k_tc_linear<NT, EPI, POS, TMR>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (2 * per_prec > epi_bytes ? 2 * per_prec : epi_bytes) + 1024));
    configured = true;
  }
  const dim3 grid(gm_div_up(a.n_rows, TMR), a.N_total / NT);
  GM_CUDA(gm_launch_pdl(k_tc_linear<NT, EPI, POS, TMR>, grid, dim3(LTHREADS), (size_t)smem, stream, a));
  return GEOMAE_OK;
}

// Rows per CTA.  Token sets that give fewer 128-row tiles than SMs (the encoder: 57 tiles) run 64-row tiles — twice
// the CTAs, each half as long (measured: 25-35 % faster per kernel there); large sets keep 128 rows (fewer weight
// fetches; measured 10-30 % faster at 190+ tiles).  GEOMAE_TC_TILE_M=64|128 forces one (read once).
inline int tile_rows(int n_rows) {
  static int forced = -1;
  if (forced < 0) {
    const char* e = getenv("GEOMAE_TC_TILE_M");
    forced = e ? atoi(e) : 0;
  }
  if (forced == 64 || forced == 128) return forced;
  return gm_div_up(n_rows, 128) < GM_NUM_SMS ? 64 : 128;
}

template <int NT, int EPI>
int launch_linear(const LinArgs& a, cudaStream_t stream) {
  const bool pos = NT == 128 && EPI == 0 && a.pos_table && a.pos_slabs > 0;
  if (tile_rows(a.n_rows) == 64) {
    if (pos) return launch_linear_impl<128, 0, true, 64>(a, stream);
    return launch_linear_impl<NT, EPI, false, 64>(a, stream);
  }
  if (pos) return launch_linear_impl<128, 0, true, 128>(a, stream);
  return launch_linear_impl<NT, EPI, false, 128>(a, stream);
}


// ------------------------------------------------------------------------------------------------
// Weight gradient: dW[m, n] += sum_tok dY[tok, m] * X[tok, n]  (+ db[m] += sum_tok dY[tok, m]).
// Both operands are MN-major views of token-major activations (K = tokens), so they are staged
// exactly like a forward operand.  The bias gradient rides along as one extra N=16 MMA against a
// constant block whose first column is 1.  A CTA accumulates its token range in TMEM and finishes
// with vector reductions into the fp32 gradient buffer.
struct WgradArgs {
  const float* dY; int ldy; const float* X; int ldx; int n_rows;
  const float* pos_table; const int32_t* tok_cell; int pos_slabs; int x_gelu;
  float* dW; int ldw; float* db; int M_total; int N_total;
  int tiles_per_cta; int precision;
  int dy_bf16;      // dY rows are bf16 (ldy in elements), precision 1 only
};

template <int NT>
__global__ void __launch_bounds__(NTHREADS, 2) k_tc_wgrad(const WgradArgs a) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t mbar;
  __shared__ uint32_t tmem_slot;
  uint8_t* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
  constexpr int A_BYTES = TM * 128 * 2;             // dY tile: 128 tokens x 128 m-columns
  constexpr int B_BYTES = TM * NT * 2;              // X tile : 128 tokens x NT n-columns
  constexpr int ONES_BYTES = TM * tc::LINE_BYTES;   // one 64-column block
  constexpr int TCOLS = 256;                        // NT = 128: accumulator + 16 bias columns; NT = 256: accumulator only
  const bool x3 = a.precision == 3;
  uint8_t* sA = smem;
  uint8_t* sAlo = sA + A_BYTES;
  uint8_t* sB = sA + (x3 ? 2 : 1) * A_BYTES;
  uint8_t* sBlo = sB + B_BYTES;
  uint8_t* sOnes = sB + (x3 ? 2 : 1) * B_BYTES;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.y * 128, n0 = blockIdx.z * NT;
  const int tile_begin = blockIdx.x * a.tiles_per_cta;
  const int n_tiles_total = (a.n_rows + TM - 1) / TM;
  const int tile_end = min(tile_begin + a.tiles_per_cta, n_tiles_total);
  if (tile_begin >= tile_end) return;
  gm_pdl_wait();
  gm_pdl_trigger();
  if (warp == 0) tc::tmem_alloc(&tmem_slot, TCOLS);
  if (threadIdx.x == 0) tc::mbar_init(&mbar, 1);
  const bool want_bias = NT == 128 && a.db != nullptr && blockIdx.z == 0;
  // constant "ones" block: element (row, col 0) = 1.0
  for (int i = threadIdx.x; i < TM * 8; i += NTHREADS) {
    const int r = i >> 3, c = i & 7;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (c == 0) v.x = 0x00003F80u;    // bf16(1.0) in the low half
    *reinterpret_cast<uint4*>(sOnes + tc::swz(r, c)) = v;
  }
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = tmem_slot;
  const bool use_pos = a.pos_table && (int)blockIdx.y < a.pos_slabs;
  int it = 0;
  for (int tile = tile_begin; tile < tile_end; ++tile, ++it) {
    const int row0 = tile * TM;
    if (it > 0) {
      tc::mbar_wait(&mbar, (it - 1) & 1);
      tc::fence_after_sync();
    }
    if (a.dy_bf16)
      stage_tile_bf16<TM, 128, NTHREADS>(sA, reinterpret_cast<const __nv_bfloat16*>(a.dY), a.ldy, row0, a.n_rows, m0);
    else
      stage_tile<TM, 128, NTHREADS, false>(sA, x3 ? sAlo : nullptr, a.dY, a.ldy, row0, a.n_rows, m0, nullptr, nullptr, 0, false);
    if (use_pos)
      stage_tile<TM, NT, NTHREADS, true>(sB, x3 ? sBlo : nullptr, a.X, a.ldx, row0, a.n_rows, n0, a.pos_table, a.tok_cell,
                                         a.N_total, false);
    else
      stage_tile<TM, NT, NTHREADS, false>(sB, x3 ? sBlo : nullptr, a.X, a.ldx, row0, a.n_rows, n0, nullptr, nullptr, 0,
                                          a.x_gelu != 0);
    tc::fence_async_smem();
    tc::fence_before_sync();
    __syncthreads();
    if (threadIdx.x == 0) {
      tc::fence_after_sync();
      const uint32_t idesc = tc::make_idesc_bf16(128, NT, 1, 1);
      const uint32_t idesc_b = tc::make_idesc_bf16(128, 16, 1, 1);
      const uint32_t lbo = TM * tc::LINE_BYTES;     // stride between 64-column blocks
      bool acc = it > 0;
#pragma unroll
      for (int j = 0; j < TM / 16; ++j) {
        const uint32_t off = (uint32_t)j * 2 * tc::ATOM_BYTES;
        const uint64_t da_hi = tc::make_desc(tc::smem_u32(sA) + off, lbo, tc::ATOM_BYTES);
        const uint64_t db_hi = tc::make_desc(tc::smem_u32(sB) + off, lbo, tc::ATOM_BYTES);
        tc::mma_bf16(tmem, da_hi, db_hi, idesc, acc);
        uint64_t da_lo = 0;
        if (x3) {
          da_lo = tc::make_desc(tc::smem_u32(sAlo) + off, lbo, tc::ATOM_BYTES);
          const uint64_t db_lo = tc::make_desc(tc::smem_u32(sBlo) + off, lbo, tc::ATOM_BYTES);
          tc::mma_bf16(tmem, da_hi, db_lo, idesc, true);
          tc::mma_bf16(tmem, da_lo, db_hi, idesc, true);
        }
        if (want_bias) {
          const uint64_t d1 = tc::make_desc(tc::smem_u32(sOnes) + off, lbo, tc::ATOM_BYTES);
          tc::mma_bf16(tmem + NT, da_hi, d1, idesc_b, acc);
          if (x3) tc::mma_bf16(tmem + NT, da_lo, d1, idesc_b, true);
        }
        acc = true;
      }
      tc::mma_commit(&mbar);
    }
  }
  tc::mbar_wait(&mbar, (it - 1) & 1);
  tc::fence_after_sync();
  const int qw = warp & 3, chalf = warp >> 2;          // TMEM lane quarter, column half
  const int m = m0 + qw * 32 + lane;                   // this thread's dW row
  const uint32_t t_lane = tmem + ((uint32_t)(qw * 32) << 16);
#pragma unroll 1
  for (int c0 = chalf * (NT / 2); c0 < (chalf + 1) * (NT / 2); c0 += 32) {
    float v[32];
    tc::tmem_ld32(t_lane + c0, v);
    tc::tmem_ld_wait();
    if (m < a.M_total) {
      float* o = a.dW + (int64_t)m * a.ldw + n0 + c0;
#pragma unroll
      for (int c = 0; c < 32; c += 4)
        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(o + c), "f"(v[c]), "f"(v[c + 1]),
                     "f"(v[c + 2]), "f"(v[c + 3])
                     : "memory");
    }
  }
  if (want_bias && chalf == 0) {
    float v[32];
    tc::tmem_ld32(t_lane + NT, v);
    tc::tmem_ld_wait();
    if (m < a.M_total) atomicAdd(a.db + m, v[0]);
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_free(tmem, TCOLS);
}

template <int NT>
int launch_wgrad(const WgradArgs& a, cudaStream_t stream) {
  const int max_smem = 2 * (TM * 128 * 2 + TM * NT * 2) + TM * tc::LINE_BYTES + 1024;
  const int smem = (a.precision == 3 ? 2 : 1) * (TM * 128 * 2 + TM * NT * 2) + TM * tc::LINE_BYTES + 1024;
  static bool configured = false;
  if (!configured) {
    GM_CUDA(cudaFuncSetAttribute(k_tc_wgrad<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
    configured = true;
  }
  const int n_tiles = gm_div_up(a.n_rows, TM);
  const dim3 grid(gm_div_up(n_tiles, a.tiles_per_cta), a.M_total / 128, a.N_total / NT);
  GM_CUDA(gm_launch_pdl(k_tc_wgrad<NT>, grid, dim3(NTHREADS), (size_t)smem, stream, a));
  return GEOMAE_OK;
}

// ------------------------------------------------------------------------------------------------
// LayerNorm backward over saved (pre-LN row, mean, rstd): one warp per row, float4 per lane (C = 128).
__global__ void __launch_bounds__(256) k_ln_bwd(const float* __restrict__ dz, const float* __restrict__ s,
                                                const float* __restrict__ stats, const float* __restrict__ gamma,
                                                int n_rows, float* ds, float* dgamma, float* dbeta, float* dsum) {
  __shared__ float sg[8][128], sb[8][128];
  __shared__ float ss[8][128];                  // column sums of ds (= bias gradient of the linear that produced the LN input)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  gm_pdl_wait();
  gm_pdl_trigger();
  const float4 g4 = __ldg(reinterpret_cast<const float4*>(gamma) + lane);
  float4 accg = make_float4(0.f, 0.f, 0.f, 0.f), accb = accg, accs = accg;
  for (int row = blockIdx.x * 8 + warp; row < n_rows; row += gridDim.x * 8) {
    const float4 d = __ldg(reinterpret_cast<const float4*>(dz + (int64_t)row * 128) + lane);
    const float4 x = __ldg(reinterpret_cast<const float4*>(s + (int64_t)row * 128) + lane);
    const float mean = __ldg(stats + 2 * (int64_t)row), rstd = __ldg(stats + 2 * (int64_t)row + 1);
    const float4 xh = make_float4((x.x - mean) * rstd, (x.y - mean) * rstd, (x.z - mean) * rstd, (x.w - mean) * rstd);
    const float4 g = make_float4(d.x * g4.x, d.y * g4.y, d.z * g4.z, d.w * g4.w);
    float s1 = (g.x + g.y) + (g.z + g.w);
    float s2 = (g.x * xh.x + g.y * xh.y) + (g.z * xh.z + g.w * xh.w);
    s1 = gm_warp_sum(s1) * (1.0f / 128.f);
    s2 = gm_warp_sum(s2) * (1.0f / 128.f);
    const float4 o = make_float4(rstd * (g.x - s1 - xh.x * s2), rstd * (g.y - s1 - xh.y * s2), rstd * (g.z - s1 - xh.z * s2),
                                 rstd * (g.w - s1 - xh.w * s2));
    reinterpret_cast<float4*>(ds + (int64_t)row * 128)[lane] = o;
    accs.x += o.x; accs.y += o.y; accs.z += o.z; accs.w += o.w;
    accg.x += d.x * xh.x; accg.y += d.y * xh.y; accg.z += d.z * xh.z; accg.w += d.w * xh.w;
    accb.x += d.x; accb.y += d.y; accb.z += d.z; accb.w += d.w;
  }
  reinterpret_cast<float4*>(sg[warp])[lane] = accg;
  reinterpret_cast<float4*>(sb[warp])[lane] = accb;
  reinterpret_cast<float4*>(ss[warp])[lane] = accs;
  __syncthreads();
  if (threadIdx.x < 128) {
    float tg = 0.f, tb = 0.f, ts = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) { tg += sg[w][threadIdx.x]; tb += sb[w][threadIdx.x]; ts += ss[w][threadIdx.x]; }
    atomicAdd(dgamma + threadIdx.x, tg);
    atomicAdd(dbeta + threadIdx.x, tb);
    if (dsum) atomicAdd(dsum + threadIdx.x, ts);
  }
}


// ------------------------------------------------------------------------------------------------
// Weight pre-packing: fp32 [rows, cols] -> bf16 hi (+ lo residual) images made of [128 x 64] swizzled blocks
// (16 KB each, block index = band * (cols/64) + column block), i.e. exactly the bytes the tensor-core kernels
// want in shared memory, so a CTA fetches its weight tile with one or two bulk copies and no thread work.
struct PackItem { const float* W; int rows; int cols; uint8_t* hi; uint8_t* lo; };
struct PackBatch { PackItem item[64]; int n; };

__global__ void __launch_bounds__(256) k_pack_weights(const PackBatch b) {
  const PackItem it = b.item[blockIdx.y];
  const int bands = (it.rows + 127) / 128, cpr = it.cols / 8;
  const int total = bands * 128 * cpr;
  for (int i = blockIdx.x * 256 + threadIdx.x; i < total; i += gridDim.x * 256) {
    const int r = i / cpr, c8 = i % cpr;
    float f[8];
    if (r < it.rows) {
      const float4* p = reinterpret_cast<const float4*>(it.W + (int64_t)r * it.cols + c8 * 8);
      const float4 x = __ldg(p), y = __ldg(p + 1);
      f[0] = x.x; f[1] = x.y; f[2] = x.z; f[3] = x.w; f[4] = y.x; f[5] = y.y; f[6] = y.z; f[7] = y.w;
    } else {
#pragma unroll
      for (int k = 0; k < 8; ++k) f[k] = 0.f;
    }
    uint4 lo;
    const uint4 hi = tc::pack8(f, &lo);
    const size_t off = ((size_t)(r / 128) * (it.cols / 64) + (c8 >> 3)) * (128 * tc::LINE_BYTES) + tc::swz(r & 127, c8 & 7);
    *reinterpret_cast<uint4*>(it.hi + off) = hi;
    if (it.lo) *reinterpret_cast<uint4*>(it.lo + off) = lo;
  }
}

}  // namespace

extern "C" int geomae_tc_linear(const geomae_linear_args* p, void* stream_) {
  GM_REQUIRE(p && p->A && p->W && p->out, "tc_linear: null argument");
  GM_REQUIRE(p->K > 0 && p->K % KC == 0, "tc_linear: K=%d must be a positive multiple of %d", p->K, KC);
  GM_REQUIRE(p->N_total > 0 && p->N_total % 128 == 0, "tc_linear: N=%d must be a multiple of 128", p->N_total);
  GM_REQUIRE(p->precision == 1 || p->precision == 3, "tc_linear: precision must be 1 (bf16) or 3 (bf16x3)");
  GM_REQUIRE(p->epilogue >= 0 && p->epilogue <= 3, "tc_linear: unknown epilogue %d", p->epilogue);
  GM_REQUIRE(p->lda % 4 == 0 && p->ldw % 4 == 0 && p->ldo % 4 == 0, "tc_linear: leading dimensions must be multiples of 4");
  if (p->n_rows == 0) return GEOMAE_OK;
  LinArgs a;
  a.A = p->A; a.lda = p->lda; a.n_rows = p->n_rows; a.K = p->K;
  a.pos_table = p->pos_table; a.tok_cell = p->tok_cell; a.pos_slabs = p->pos_slabs; a.a_gelu = p->a_gelu;
  a.W = p->W; a.ldw = p->ldw; a.w_rows = p->w_rows; a.w_mn_major = p->w_mn_major;
  a.Wp_hi = (const uint8_t*)p->Wp_hi; a.Wp_lo = (const uint8_t*)p->Wp_lo; a.wp_cols = p->ldw;
  GM_REQUIRE(!p->Wp_hi || (p->precision == 1 || p->Wp_lo), "tc_linear: bf16x3 with packed weights needs the lo image");
  GM_REQUIRE(!p->Wp_hi || p->ldw % 64 == 0, "tc_linear: packed weights need cols %% 64 == 0");
  a.bias = p->bias; a.N_total = p->N_total; a.out = p->out; a.ldo = p->ldo;
  a.add_src = p->add_src; a.ld_add = p->ld_add;
  a.ln_gamma = p->ln_gamma; a.ln_beta = p->ln_beta; a.ln_eps = p->ln_eps; a.ln_in = p->ln_in; a.ln_stats = p->ln_stats;
  a.gelu_u = p->gelu_u; a.ldu = p->ldu; a.precision = p->precision;
  a.w_early = gm_weights_stable() ? 1 : 0;
  a.dot_src = p->dot_src; a.ld_dot = p->ld_dot; a.dot_out = p->dot_out;
  a.ln_dgamma = a.ln_dbeta = a.ln_dcolsum = nullptr;
  a.a_bf16 = p->a_bf16; a.out_bf16 = p->out_bf16;
  GM_REQUIRE(!p->a_bf16 || (p->precision == 1 && !p->pos_table && !p->a_gelu && p->lda % 8 == 0),
             "tc_linear: bf16 A rows need precision 1, no position / GELU prologue, lda %% 8 == 0");
  GM_REQUIRE(!p->out_bf16 || (p->epilogue == 0 || p->epilogue == 2), "tc_linear: bf16 output only for epilogues 0 and 2");
  GM_REQUIRE(!p->dot_src || (p->dot_out && p->epilogue == 0 && p->N_total == 128 && p->pos_slabs == 0 && p->ld_dot % 4 == 0),
             "tc_linear: the per-head dot side output needs epilogue 0, N = 128, no position prologue");
  GM_REQUIRE(!p->dot_src || !p->add_src, "tc_linear: the per-head dot side output cannot be combined with add_src");
  cudaStream_t stream = (cudaStream_t)stream_;
  if (p->epilogue == 1) {
    GM_REQUIRE(p->N_total == 128 && p->bias && p->add_src && p->ln_gamma && p->ln_beta,
               "tc_linear: LayerNorm epilogue needs N=128, bias, residual, gamma, beta");
    return launch_linear<128, 1>(a, stream);
  }
  if (p->epilogue == 3) {
    GM_REQUIRE(p->N_total == 128 && p->ln_gamma && p->ln_in && p->ln_stats && p->ln_dgamma && p->ln_dbeta && !p->bias,
               "tc_linear: LayerNorm-backward epilogue needs N=128, gamma, saved rows + stats, gradient accumulators, no bias");
    a.ln_dgamma = p->ln_dgamma; a.ln_dbeta = p->ln_dbeta; a.ln_dcolsum = p->ln_dcolsum;
    return launch_linear<128, 3>(a, stream);
  }
  if (p->epilogue == 2) {
    GM_REQUIRE(p->gelu_u, "tc_linear: gelu-grad epilogue needs u");
    return (p->N_total % 256 == 0) ? launch_linear<256, 2>(a, stream) : launch_linear<128, 2>(a, stream);
  }
  return (p->N_total % 256 == 0 && p->pos_slabs == 0) ? launch_linear<256, 0>(a, stream) : launch_linear<128, 0>(a, stream);
}

extern "C" int geomae_tc_wgrad(const geomae_wgrad_args* p, void* stream_) {
  GM_REQUIRE(p && p->dY && p->X && p->dW, "tc_wgrad: null argument");
  GM_REQUIRE(p->M_total > 0 && p->M_total % 128 == 0 && p->N_total > 0 && p->N_total % 128 == 0,
             "tc_wgrad: M=%d and N=%d must be multiples of 128", p->M_total, p->N_total);
  GM_REQUIRE(p->precision == 1 || p->precision == 3, "tc_wgrad: precision must be 1 or 3");
  GM_REQUIRE(p->ldy % 4 == 0 && p->ldx % 4 == 0 && p->ldw % 4 == 0, "tc_wgrad: leading dimensions must be multiples of 4");
  if (p->n_rows == 0) return GEOMAE_OK;
  WgradArgs a;
  a.dY = p->dY; a.ldy = p->ldy; a.X = p->X; a.ldx = p->ldx; a.n_rows = p->n_rows;
  a.pos_table = p->pos_table; a.tok_cell = p->tok_cell; a.pos_slabs = p->pos_slabs; a.x_gelu = p->x_gelu;
  a.dW = p->dW; a.ldw = p->ldw; a.db = p->db; a.M_total = p->M_total; a.N_total = p->N_total;
  a.precision = p->precision;
  a.dy_bf16 = p->dy_bf16;
  GM_REQUIRE(!p->dy_bf16 || (p->precision == 1 && p->ldy % 8 == 0), "tc_wgrad: bf16 dY rows need precision 1, ldy %% 8 == 0");
  const int n_tiles = gm_div_up(p->n_rows, TM);
  const int slabs = (p->M_total / 128) * ((p->N_total % 256 == 0 && !p->db) ? p->N_total / 256 : p->N_total / 128);
  // target CTAs per SM x 2 (GEOMAE_WGRAD_CTAS_X2 overrides).  Measured on the 4-frame step: 1 per SM 5.09 ms, 2 per SM
  // 5.14 ms, 1 per 2 SMs 5.23 ms — more CTAs mean more red.add flushes of the same dW addresses, fewer mean longer chains.
  static int ctas_x2 = -1;
  if (ctas_x2 < 0) { const char* e = getenv("GEOMAE_WGRAD_CTAS_X2"); ctas_x2 = e ? atoi(e) : 2; if (ctas_x2 < 1) ctas_x2 = 2; }
  int splits = (ctas_x2 * GM_NUM_SMS / 2 + slabs - 1) / slabs;
  if (splits > n_tiles) splits = n_tiles;
  a.tiles_per_cta = gm_div_up(n_tiles, splits);
  // the 256-wide variant fills its TMEM allocation with the accumulator: no bias column there
  return (p->N_total % 256 == 0 && !p->db) ? launch_wgrad<256>(a, (cudaStream_t)stream_)
                                           : launch_wgrad<128>(a, (cudaStream_t)stream_);
}

extern "C" int geomae_layernorm_bwd(const float* d_out, const float* ln_in, const float* ln_stats, const float* gamma,
                                    int64_t n_rows, int32_t channels, float* d_in, float* d_gamma, float* d_beta,
                                    float* d_in_colsum, void* stream) {
  GM_REQUIRE(channels == 128, "layernorm_bwd: specialised for 128 channels (got %d)", channels);
  if (n_rows == 0) return GEOMAE_OK;
  GM_REQUIRE(d_out && ln_in && ln_stats && gamma && d_in && d_gamma && d_beta, "layernorm_bwd: null argument");
  int blocks = gm_div_up(n_rows, 8);
  if (blocks > GM_NUM_SMS * 4) blocks = GM_NUM_SMS * 4;
  GM_CUDA(gm_launch_pdl(k_ln_bwd, dim3(blocks), dim3(256), (size_t)0, (cudaStream_t)stream, d_out, ln_in, ln_stats, gamma,
                        (int)n_rows, d_in, d_gamma, d_beta, d_in_colsum));
  return GEOMAE_OK;
}

extern "C" int geomae_pack_weights(int32_t n_items, const float* const* W, const int32_t* rows, const int32_t* cols,
                                   void* const* hi, void* const* lo, void* stream) {
  GM_REQUIRE(n_items >= 0 && (n_items == 0 || (W && rows && cols && hi)), "pack_weights: null argument");
  int done = 0;
  while (done < n_items) {
    PackBatch b;
    b.n = n_items - done < 64 ? n_items - done : 64;
    int max_total = 0;
    for (int i = 0; i < b.n; ++i) {
      const int k = done + i;
      GM_REQUIRE(W[k] && hi[k] && rows[k] > 0 && cols[k] > 0 && cols[k] % 64 == 0,
                 "pack_weights: item %d needs cols %% 64 == 0 and non-null buffers", k);
      b.item[i] = PackItem{W[k], rows[k], cols[k], (uint8_t*)hi[k], lo ? (uint8_t*)lo[k] : nullptr};
      const int total = (rows[k] + 127) / 128 * 128 * (cols[k] / 8);
      if (total > max_total) max_total = total;
    }
    int gx = gm_div_up(max_total, 256 * 4);
    if (gx < 1) gx = 1;
    k_pack_weights<<<dim3(gx, b.n), 256, 0, (cudaStream_t)stream>>>(b);
    GM_LAUNCH_CHECK();
    done += b.n;
  }
  return GEOMAE_OK;
}
