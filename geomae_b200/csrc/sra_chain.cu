// Token-local chain of one SRA EncoderLayer as ONE persistent, warp-specialised tcgen05 kernel (bf16 mode):
//
//   forward  (k_sra_chain_fwd):  s1 = x + O Wo^T + bo ; y = LN1(s1) ; u = y W1^T + b1 ; g = gelu(u) ;
//                                s2 = y + g W2^T + b2 ; z = LN2(s2) ; q|k|v(next layer) = (z + pos | z) Win'^T + bin'
//
// models/sst/sst_basic_block.py:26-61 (out_proj of nn.MultiheadAttention, the in-projection of the NEXT layer's
// attention), :85-102 (post-norm residual blocks, exact-erf GELU feed-forward).  Everything here is token-local: the
// only window-local step of a layer, the attention core, stays in sra_attention_tc.cu and hands over O in bf16.
//
// One CTA walks 128-token tiles (flat token order, so every global row block is contiguous).  Warp roles:
//   warps 0-7  "compute"  thread (quarter q = warp%4, lane) owns accumulator row q*32+lane (one TMEM lane) and the
//                         column half warp/4: epilogues (bias, residual, LayerNorm, GELU), operand tiles for the next
//                         GEMM written straight into the 128B-swizzled layout, coalesced copy-out of everything the
//                         backward needs
//   warp 8     "producer" streams the layer's pre-packed bf16 weight blocks (16 KB each, 16 per tile) from L2 through
//                         a 3-slot shared-memory ring with bulk async copies (UBLKCP) and full/empty mbarriers
//   warp 9     "mma"      one thread issues tcgen05.mma (M=128, N=128, K=16) from the operand tiles and the ring
//                         into TMEM and commits to the ring's empty barriers / the accumulator-full barriers
// so weight traffic, tensor-core work and the SIMT epilogues of a tile overlap, and nothing but the saved tensors
// ever goes to HBM (the five-kernel version round-tripped every intermediate in fp32).
#include <cuda_bf16.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace {

constexpr int CT = 128;            // token rows per tile = M of every MMA
constexpr int NCOMP = 256;         // compute threads
constexpr int NTHR = 320;          // + producer warp + mma warp
constexpr int R_LD = 132;          // floats per row of the fp32 staging tile (528 B: conflict-free both ways)
constexpr int ST_LD = 272;         // bytes per row of the bf16 staging tile (256 B + 16 B pad)
constexpr int RING = 3;
constexpr int BLK = 16384;         // one packed [128 x 64] bf16 block

// shared-memory carve-up (from a 1024-byte aligned base)
constexpr int OFF_OP = 0;                          // operand tiles: A [0,32K), A2 [32K,64K); G = all 64 KB
constexpr int OFF_RING = OFF_OP + 4 * BLK;
constexpr int OFF_R = OFF_RING + RING * BLK;
constexpr int OFF_ST = OFF_R + CT * R_LD * 4;
constexpr int OFF_PAR = OFF_ST + CT * ST_LD;       // biases and LayerNorm parameters (1408 floats)
constexpr int PAR_FLOATS = 1408;
constexpr int OFF_RED = OFF_PAR + PAR_FLOATS * 4;  // [2][128][2] partial row statistics
constexpr int OFF_BAR = OFF_RED + 2 * CT * 2 * 4;
constexpr int SMEM_BYTES = OFF_BAR + 256 + 1024;   // + alignment slack

struct FwdArgs {
  int n; int mode;
  const float* x; const __nv_bfloat16* attn;
  const uint8_t *Wo, *W1, *W2, *Win;
  const float *bo, *b1, *b2, *bin, *g1, *be1, *g2, *be2; float eps;
  const float* pos; const int32_t* cell_next;
  float *s1, *st1, *s2, *st2, *z;
  __nv_bfloat16 *y16, *u16, *g16, *xp16, *xb16, *qkv16;
};

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tc::smem_u32(bar)) : "memory");
}
// operand tile / TMEM pre-load of this warp complete: one arrival per warp (the barriers expect NCOMP / 32 arrivals)
__device__ __forceinline__ void warp_arrive(uint64_t* bar) {
  tc::fence_async_smem();
  tc::fence_before_sync();
  __syncwarp();
  if ((threadIdx.x & 31) == 0) mbar_arrive(bar);
}
__device__ __forceinline__ void compute_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }
__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(tc::smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;\n" ::: "memory");
}
__device__ __forceinline__ uint32_t pack2(float a, float b) {
  const __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&v);
}
__device__ __forceinline__ uint4 pack8f(const float* f) {
  return make_uint4(pack2(f[0], f[1]), pack2(f[2], f[3]), pack2(f[4], f[5]), pack2(f[6], f[7]));
}
// byte offset of (row, 16-byte chunk c of a 128-column bf16 row) in a two-block swizzled operand tile
__device__ __forceinline__ uint32_t op_off(int row, int c) { return (uint32_t)(c >> 3) * BLK + tc::swz(row, c & 7); }

// ---- cooperative (all 256 compute threads) tile <-> global row copies; rows >= m are skipped
__device__ __forceinline__ void store_rows_f32(const float* sR, float* dst, int row0, int m) {
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll 4
  for (int r = w; r < m; r += 8)
    reinterpret_cast<float4*>(dst + (int64_t)(row0 + r) * 128)[lane] = *reinterpret_cast<const float4*>(sR + r * R_LD + lane * 4);
}
// bf16 rows of 128 columns from the padded staging tile to dst[(row0+r)*ld + col0 ..]
__device__ __forceinline__ void store_rows_st(const uint8_t* sT, __nv_bfloat16* dst, int ld, int col0, int row0, int m) {
  const int c = threadIdx.x & 15;
#pragma unroll 4
  for (int r = threadIdx.x >> 4; r < m; r += 16)
    *reinterpret_cast<uint4*>(dst + (int64_t)(row0 + r) * ld + col0 + c * 8) = *reinterpret_cast<const uint4*>(sT + r * ST_LD + c * 16);
}
// bf16 rows of a swizzled operand tile (NB 64-column blocks) to dst[(row0+r)*ld ..]
template <int NB>
__device__ __forceinline__ void store_rows_op(const uint8_t* sOp, __nv_bfloat16* dst, int ld, int row0, int m) {
  constexpr int CPR = NB * 8;                       // 16-byte chunks per row
  for (int i = threadIdx.x; i < m * CPR; i += NCOMP) {
    const int r = i / CPR, c = i % CPR;
    *reinterpret_cast<uint4*>(dst + (int64_t)(row0 + r) * ld + c * 8) = *reinterpret_cast<const uint4*>(sOp + op_off(r, c));
  }
}

// exclusive per-row LayerNorm statistics from the two column halves of a row (threads of warps w and w+4)
__device__ __forceinline__ void row_stats(const float* v, float* red, int r, int hsel, float eps, float& mean, float& rstd) {
  float s = 0.f;
#pragma unroll
  for (int c = 0; c < 64; ++c) s += v[c];
  red[r * 2 + hsel] = s;
  compute_sync();
  mean = (red[r * 2] + red[r * 2 + 1]) * (1.0f / 128.f);
  float q = 0.f;
#pragma unroll
  for (int c = 0; c < 64; ++c) { const float d = v[c] - mean; q = fmaf(d, d, q); }
  red[2 * CT + r * 2 + hsel] = q;
  compute_sync();
  rstd = rsqrtf((red[2 * CT + r * 2] + red[2 * CT + r * 2 + 1]) * (1.0f / 128.f) + eps);
}

__global__ void __launch_bounds__(NTHR, 1) k_sra_chain_fwd(const FwdArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* sm = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sA = sm + OFF_OP;
  uint8_t* sA2 = sA + 2 * BLK;
  uint8_t* sG = sA;
  uint8_t* sRing = sm + OFF_RING;
  float* sR = reinterpret_cast<float*>(sm + OFF_R);
  uint8_t* sT = sm + OFF_ST;
  float* sPar = reinterpret_cast<float*>(sm + OFF_PAR);
  float* sRed = reinterpret_cast<float*>(sm + OFF_RED);
  uint64_t* w_full = reinterpret_cast<uint64_t*>(sm + OFF_BAR);
  uint64_t* w_empty = w_full + RING;
  uint64_t* a_ready = w_empty + RING;      // [4] compute -> mma: operand tile of GEMM g is complete
  uint64_t* acc_full = a_ready + 4;        // [4] mma -> compute: accumulators of GEMM g are complete
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 4);
  // parameter cache: bo | b1 | b2 | bin | g1 | be1 | g2 | be2
  float* p_bo = sPar; float* p_b1 = sPar + 128; float* p_b2 = sPar + 384; float* p_bin = sPar + 512;
  float* p_g1 = sPar + 896; float* p_be1 = sPar + 1024; float* p_g2 = sPar + 1152; float* p_be2 = sPar + 1280;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_tiles = (a.n + CT - 1) / CT;
  const bool chain = a.mode & 1, next = (a.mode & 2) != 0;
  if (threadIdx.x == 0) {
    for (int i = 0; i < RING; ++i) { tc::mbar_init(&w_full[i], 1); tc::mbar_init(&w_empty[i], 1); }
    for (int i = 0; i < 4; ++i) { tc::mbar_init(&a_ready[i], NCOMP / 32); tc::mbar_init(&acc_full[i], 1); }
  }
  if (warp == 9) tc::tmem_alloc(tmem_slot, 512);
  if (threadIdx.x < NCOMP) {                       // parameters are never written by a kernel of this stream's chain
    for (int i = threadIdx.x; i < PAR_FLOATS; i += NCOMP) {
      float v = 0.f;
      if (i < 128) { if (chain) v = a.bo[i]; }
      else if (i < 384) { if (chain) v = a.b1[i - 128]; }
      else if (i < 512) { if (chain) v = a.b2[i - 384]; }
      else if (i < 896) { if (next) v = a.bin[i - 512]; }
      else if (i < 1024) { if (chain) v = a.g1[i - 896]; }
      else if (i < 1152) { if (chain) v = a.be1[i - 1024]; }
      else if (i < 1280) { if (chain) v = a.g2[i - 1152]; }
      else { if (chain) v = a.be2[i - 1280]; }
      sPar[i] = v;
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = *tmem_slot;
  gm_pdl_wait();                                   // everything below reads what earlier kernels wrote
  gm_pdl_trigger();

  if (warp == 8) {
    // ------------------------------------------------------------------ producer: weight blocks through the ring
    if (lane == 0) {
      uint32_t cnt = 0;
      auto push = [&](const uint8_t* img, int nblk) {
        for (int b = 0; b < nblk; ++b, ++cnt) {
          const uint32_t slot = cnt % RING, ph = (cnt / RING) & 1;
          tc::mbar_wait(&w_empty[slot], ph ^ 1);
          tc::mbar_expect_tx(&w_full[slot], BLK);
          tc::bulk_g2s(sRing + slot * BLK, img + (size_t)b * BLK, BLK, &w_full[slot]);
        }
      };
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        if (chain) { push(a.Wo, 2); push(a.W1, 4); push(a.W2, 4); }
        if (next) push(a.Win, 6);
      }
    }
  } else if (warp == 9) {
    // ------------------------------------------------------------------ mma issuer
    if (lane == 0) {
      const uint32_t idesc = tc::make_idesc_bf16(128, 128, 0, 0);
      uint32_t cnt = 0;
      // one ring block = 64 K-columns = 4 MMAs of K = 16 against the operand block at `a_addr`
      auto mma_block = [&](uint32_t a_addr, uint32_t tcol, bool acc) {
        const uint32_t slot = cnt % RING, ph = (cnt / RING) & 1;
        tc::mbar_wait(&w_full[slot], ph);
        tc::fence_after_sync();
        const uint32_t b_addr = tc::smem_u32(sRing + slot * BLK);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)
          tc::mma_bf16(tmem + tcol, tc::make_desc(a_addr + kk * 32, 16, tc::ATOM_BYTES),
                       tc::make_desc(b_addr + kk * 32, 16, tc::ATOM_BYTES), idesc, acc || kk > 0);
        tc::mma_commit(&w_empty[slot]);
        ++cnt;
      };
      const uint32_t A = tc::smem_u32(sA), A2 = tc::smem_u32(sA2);
      int it = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
        const uint32_t par = it & 1;
        if (chain) {
          tc::mbar_wait(&a_ready[0], par);            // O in A: s1 acc = O Wo^T -> columns [0,128)
          tc::fence_after_sync();
          for (int j = 0; j < 2; ++j) mma_block(A + j * BLK, 0, j > 0);
          tc::mma_commit(&acc_full[0]);
          tc::mbar_wait(&a_ready[1], par);            // y in A: u acc = y W1^T -> columns [128,384)
          tc::fence_after_sync();
          for (int b = 0; b < 2; ++b)
            for (int j = 0; j < 2; ++j) mma_block(A + j * BLK, 128 + b * 128, j > 0);
          tc::mma_commit(&acc_full[1]);
          tc::mbar_wait(&a_ready[2], par);            // gelu(u) in G: s2 acc = g W2^T -> columns [0,128)
          tc::fence_after_sync();
          for (int j = 0; j < 4; ++j) mma_block(A + j * BLK, 0, j > 0);
          tc::mma_commit(&acc_full[2]);
        }
        if (next) {
          tc::mbar_wait(&a_ready[3], par);            // z+pos in A, z in A2: q|k|v acc -> columns [128,512)
          tc::fence_after_sync();
          for (int b = 0; b < 3; ++b)
            for (int j = 0; j < 2; ++j) mma_block((b < 2 ? A : A2) + j * BLK, 128 + b * 128, j > 0);
          tc::mma_commit(&acc_full[3]);
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ compute warps
    const int q = warp & 3, hsel = warp >> 2;
    const int r = q * 32 + lane;                     // tile row = TMEM lane of this thread
    const int c0 = hsel * 64;                        // its column half
    const uint32_t t_lane = tmem + ((uint32_t)(q * 32) << 16);
    float* myR = sR + r * R_LD + c0;
    // x rows -> R, O rows -> A (swizzled); rows past the end are zero-filled
    auto prefetch = [&](int tile) {
      const int row0 = tile * CT, m = min(CT, a.n - row0);
      for (int i = threadIdx.x; i < CT * 32; i += NCOMP) {
        const int rr = i >> 5, c4 = i & 31;
        float* dst = sR + rr * R_LD + c4 * 4;
        if (rr < m) cp_async16(dst, a.x + (int64_t)(row0 + rr) * 128 + c4 * 4);
        else *reinterpret_cast<float4*>(dst) = make_float4(0.f, 0.f, 0.f, 0.f);
      }
      if (chain) {
        for (int i = threadIdx.x; i < CT * 16; i += NCOMP) {
          const int rr = i >> 4, c = i & 15;
          uint8_t* dst = sA + op_off(rr, c);
          if (rr < m) cp_async16(dst, a.attn + (int64_t)(row0 + rr) * 128 + c * 8);
          else *reinterpret_cast<uint4*>(dst) = make_uint4(0u, 0u, 0u, 0u);
        }
      }
    };
    int it = 0;
    if ((int)blockIdx.x < n_tiles) prefetch(blockIdx.x);
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
      const uint32_t par = it & 1;
      const int row0 = tile * CT, m = min(CT, a.n - row0);
      const int grow = row0 + r;
      float v[64];
      cp_async_wait_all();
      if (chain) {
        warp_arrive(&a_ready[0]);
      }
      compute_sync();                                // R (and A) complete for every thread
      if (chain) {
        // ---------------- E1: s1 = acc + bo + x ; y = LN1(s1)
        tc::mbar_wait(&acc_full[0], par);
        tc::fence_after_sync();
        tc::tmem_ld32(t_lane + c0, v);
        tc::tmem_ld32(t_lane + c0 + 32, v + 32);
        tc::tmem_ld_wait();
#pragma unroll
        for (int c = 0; c < 64; c += 4) {
          const float4 xr = *reinterpret_cast<const float4*>(myR + c);
          v[c] += p_bo[c0 + c] + xr.x; v[c + 1] += p_bo[c0 + c + 1] + xr.y;
          v[c + 2] += p_bo[c0 + c + 2] + xr.z; v[c + 3] += p_bo[c0 + c + 3] + xr.w;
        }
        float mean, rstd;
        row_stats(v, sRed, r, hsel, a.eps, mean, rstd);
#pragma unroll
        for (int c = 0; c < 64; c += 4) *reinterpret_cast<float4*>(myR + c) = make_float4(v[c], v[c + 1], v[c + 2], v[c + 3]);
        if (hsel == 0 && r < m) *reinterpret_cast<float2*>(a.st1 + 2 * (int64_t)grow) = make_float2(mean, rstd);
#pragma unroll
        for (int c = 0; c < 64; ++c) v[c] = fmaf((v[c] - mean) * rstd, p_g1[c0 + c], p_be1[c0 + c]);     // y
#pragma unroll
        for (int c = 0; c < 64; c += 8) *reinterpret_cast<uint4*>(sA + op_off(r, (c0 + c) >> 3)) = pack8f(v + c);
        warp_arrive(&a_ready[1]);
        compute_sync();
        store_rows_f32(sR, a.s1, row0, m);           // saved pre-LN1 rows
        store_rows_op<2>(sA, a.y16, 128, row0, m);   // bf16 y: operand of the lin1 weight gradient
        compute_sync();
#pragma unroll
        for (int c = 0; c < 64; c += 4) *reinterpret_cast<float4*>(myR + c) = make_float4(v[c], v[c + 1], v[c + 2], v[c + 3]);
        // ---------------- E2: u = acc + b1 ; g = gelu(u)   (two 128-column bands)
        tc::mbar_wait(&acc_full[1], par);
        tc::fence_after_sync();
#pragma unroll 1
        for (int band = 0; band < 2; ++band) {
          tc::tmem_ld32(t_lane + 128 + band * 128 + c0, v);
          tc::tmem_ld32(t_lane + 128 + band * 128 + c0 + 32, v + 32);
          tc::tmem_ld_wait();
#pragma unroll
          for (int c = 0; c < 64; c += 8) {
            float u8[8], g8[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              u8[e] = v[c + e] + p_b1[band * 128 + c0 + c + e];
              g8[e] = u8[e] * gelu_parts(u8[e]).cdf;
            }
            *reinterpret_cast<uint4*>(sT + r * ST_LD + (c0 + c) * 2) = pack8f(u8);
            const int col = band * 128 + c0 + c;     // column of g in [0,256): block col/64, chunk (col%64)/8
            *reinterpret_cast<uint4*>(sG + (uint32_t)(col >> 6) * BLK + tc::swz(r, (col & 63) >> 3)) = pack8f(g8);
          }
          if (band == 1) {
            warp_arrive(&a_ready[2]);
          }
          compute_sync();
          store_rows_st(sT, a.u16, 256, band * 128, row0, m);    // saved pre-GELU rows (bf16)
          if (band == 1) store_rows_op<4>(sG, a.g16, 256, row0, m);   // bf16 gelu(u): operand of the lin2 weight gradient
          compute_sync();
        }
        // ---------------- E3: s2 = acc + b2 + y ; z = LN2(s2)
        tc::mbar_wait(&acc_full[2], par);
        tc::fence_after_sync();
        tc::tmem_ld32(t_lane + c0, v);
        tc::tmem_ld32(t_lane + c0 + 32, v + 32);
        tc::tmem_ld_wait();
#pragma unroll
        for (int c = 0; c < 64; c += 4) {
          const float4 yr = *reinterpret_cast<const float4*>(myR + c);
          v[c] += p_b2[c0 + c] + yr.x; v[c + 1] += p_b2[c0 + c + 1] + yr.y;
          v[c + 2] += p_b2[c0 + c + 2] + yr.z; v[c + 3] += p_b2[c0 + c + 3] + yr.w;
        }
        row_stats(v, sRed, r, hsel, a.eps, mean, rstd);
#pragma unroll
        for (int c = 0; c < 64; c += 4) *reinterpret_cast<float4*>(myR + c) = make_float4(v[c], v[c + 1], v[c + 2], v[c + 3]);
        if (hsel == 0 && r < m) *reinterpret_cast<float2*>(a.st2 + 2 * (int64_t)grow) = make_float2(mean, rstd);
        compute_sync();
        store_rows_f32(sR, a.s2, row0, m);           // saved pre-LN2 rows
        compute_sync();
#pragma unroll
        for (int c = 0; c < 64; c += 4)
          *reinterpret_cast<float4*>(myR + c) =
              make_float4(fmaf((v[c] - mean) * rstd, p_g2[c0 + c], p_be2[c0 + c]),
                          fmaf((v[c + 1] - mean) * rstd, p_g2[c0 + c + 1], p_be2[c0 + c + 1]),
                          fmaf((v[c + 2] - mean) * rstd, p_g2[c0 + c + 2], p_be2[c0 + c + 2]),
                          fmaf((v[c + 3] - mean) * rstd, p_g2[c0 + c + 3], p_be2[c0 + c + 3]));
        compute_sync();
      }
      // ---------------- row pass over R (= z, or the stack input x in prologue mode): z out, operands of the next in-proj
      {
        const int w = threadIdx.x >> 5;
#pragma unroll 2
        for (int rr = w; rr < CT; rr += 8) {
          const float4 z4 = *reinterpret_cast<const float4*>(sR + rr * R_LD + lane * 4);
          const bool ok = rr < m;
          if (chain && ok) reinterpret_cast<float4*>(a.z + (int64_t)(row0 + rr) * 128)[lane] = z4;
          if (next) {
            const int cell = ok ? __ldg(a.cell_next + row0 + rr) : 0;
            const float4 p4 = __ldg(reinterpret_cast<const float4*>(a.pos + (int64_t)cell * 128) + lane);
            const uint2 xp = make_uint2(pack2(z4.x + p4.x, z4.y + p4.y), pack2(z4.z + p4.z, z4.w + p4.w));
            const uint2 xb = make_uint2(pack2(z4.x, z4.y), pack2(z4.z, z4.w));
            const uint32_t off = op_off(rr, lane >> 1) + (lane & 1) * 8;
            *reinterpret_cast<uint2*>(sA + off) = xp;
            *reinterpret_cast<uint2*>(sA2 + off) = xb;
            if (ok) {
              reinterpret_cast<uint2*>(a.xp16 + (int64_t)(row0 + rr) * 128)[lane] = xp;
              reinterpret_cast<uint2*>(a.xb16 + (int64_t)(row0 + rr) * 128)[lane] = xb;
            }
          }
        }
      }
      if (next) {
        warp_arrive(&a_ready[3]);
      }
      compute_sync();                                // R is free
      const int ntile = tile + gridDim.x;
      if (next) {
        // ---------------- E4: q|k|v of the next layer = acc + bin  (three 128-column bands, bf16)
        tc::mbar_wait(&acc_full[3], par);
        tc::fence_after_sync();
        if (ntile < n_tiles) prefetch(ntile);        // A is free too: overlaps this epilogue
#pragma unroll 1
        for (int band = 0; band < 3; ++band) {
          tc::tmem_ld32(t_lane + 128 + band * 128 + c0, v);
          tc::tmem_ld32(t_lane + 128 + band * 128 + c0 + 32, v + 32);
          tc::tmem_ld_wait();
#pragma unroll
          for (int c = 0; c < 64; c += 8) {
            float o8[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) o8[e] = v[c + e] + p_bin[band * 128 + c0 + c + e];
            *reinterpret_cast<uint4*>(sT + r * ST_LD + (c0 + c) * 2) = pack8f(o8);
          }
          compute_sync();
          store_rows_st(sT, a.qkv16, 384, band * 128, row0, m);
          compute_sync();
        }
      } else if (ntile < n_tiles) {
        prefetch(ntile);
      }
    }
    tc::fence_before_sync();
  }
  __syncthreads();
  if (warp == 9) {
    tc::fence_after_sync();
    tc::tmem_free(tmem, 512);
  }
}


// ================================================================================================================
//   backward (k_sra_chain_bwd), per 128-token tile, gradients flowing down through the same token-local chain:
//     dz  = dqkv' Win' + ds1'            in-projection backward of the layer ABOVE (' = that layer), or dz given
//     ds2 = LN2-backward(dz; s2)         d_gamma2 += sum dz xhat2, d_beta2 += sum dz
//     du  = (ds2 W2) * gelu'(u)
//     dy  = du W1 + ds2                  ds2 is PRE-LOADED into the TMEM accumulator (tcgen05.st), the MMA adds onto it
//     ds1 = LN1-backward(dy; s1)         d_gamma1, d_beta1
//     dO  = ds1 Wo ; D = dO . O per head (the attention backward's row term)
//   dX = dY W uses the weights as MN-major B operands straight from the same packed images ([128 out-rows x 64 in-cols]
//   blocks, N = 64 per MMA).  bf16 copies of ds2, du, ds1 feed the TMA weight-gradient kernel (sra_wgrad.cu); the bias
//   gradients of linear2 / out_proj are column sums of ds2 / ds1 and ride along there.
constexpr int B_OFF_OP = 0;
constexpr int B_ST_BYTES = 36864;                       // band-2 operand tile | 2 x [128][144 B] u sub-bands | [128][272 B] O rows
constexpr int B_OFF_ST = B_OFF_OP + 4 * BLK;
constexpr int B_OFF_RING = B_OFF_ST + B_ST_BYTES;
constexpr int B_OFF_R = B_OFF_RING + RING * BLK;
constexpr int B_OFF_PAR = B_OFF_R + CT * R_LD * 4;      // gamma2 | gamma1
constexpr int B_OFF_RED = B_OFF_PAR + 256 * 4;          // [128][2] float2
constexpr int B_OFF_BAR = B_OFF_RED + CT * 2 * 8;
constexpr int B_SMEM_BYTES = B_OFF_BAR + 256 + 1024;
constexpr int UH_LD = 144;                              // bytes per row of a u sub-band tile (128 B + 16 B pad)
constexpr int UH_BYTES = CT * UH_LD;
static_assert(B_OFF_ST % 1024 == 0 && B_OFF_RING % 1024 == 0, "swizzled tiles need 1024-byte alignment");
static_assert(2 * UH_BYTES <= B_ST_BYTES && CT * ST_LD <= B_ST_BYTES && 2 * BLK <= B_ST_BYTES, "staging tile too small");

struct BwdArgs {
  int n; int mode;
  const __nv_bfloat16* dqkv_up; const float* ds1_up; const uint8_t* Win_up; const float* dz_in;
  const float *s2, *st2, *s1, *st1; const __nv_bfloat16 *u16, *attn16;
  const uint8_t *W2, *W1, *Wo; const float *g2, *g1;
  __nv_bfloat16 *ds2_16, *du16, *ds1_16, *dO16; float *ds1, *dd, *dx;
  float *d_g2, *d_be2, *d_g1, *d_be1;
};

__device__ __forceinline__ void tmem_st32(uint32_t taddr, const float* v) {
  const uint32_t* r = reinterpret_cast<const uint32_t*>(v);
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};\n" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]),
      "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),
      "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

// LayerNorm backward of one row half held in v (= dz), xhat recomputed from the saved pre-LN row in R (overwritten in
// place with dz * xhat for the d_gamma column sums).  On return v holds d(pre-LN row), xh the untouched dz.
__device__ __forceinline__ void ln_backward_row(float* v, float* xh, float* myR, const float* gamma, float mean, float rstd,
                                                float2* red, int r, int hsel) {
  float p1 = 0.f, p2 = 0.f;
#pragma unroll
  for (int c = 0; c < 64; c += 4) {
    const float4 s4 = *reinterpret_cast<const float4*>(myR + c);
    const float s[4] = {s4.x, s4.y, s4.z, s4.w};
    float t[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      xh[c + e] = (s[e] - mean) * rstd;
      const float g = v[c + e] * gamma[c + e];
      p1 += g;
      p2 = fmaf(g, xh[c + e], p2);
      t[e] = v[c + e] * xh[c + e];
    }
    *reinterpret_cast<float4*>(myR + c) = make_float4(t[0], t[1], t[2], t[3]);
  }
  red[r * 2 + hsel] = make_float2(p1, p2);
  compute_sync();
  const float2 ra = red[r * 2], rb = red[r * 2 + 1];
  const float m1 = (ra.x + rb.x) * (1.0f / 128.f), m2 = (ra.y + rb.y) * (1.0f / 128.f);
#pragma unroll
  for (int c = 0; c < 64; ++c) {
    const float dzv = v[c];
    v[c] = rstd * (dzv * gamma[c] - m1 - xh[c] * m2);
    xh[c] = dzv;                                      // keep dz for the d_beta column sums
  }
}

__global__ void __launch_bounds__(NTHR, 1) k_sra_chain_bwd(const BwdArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* sm = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sA = sm + B_OFF_OP;
  uint8_t* sA2 = sA + 2 * BLK;
  uint8_t* sG = sA;
  uint8_t* sT = sm + B_OFF_ST;
  uint8_t* sRing = sm + B_OFF_RING;
  float* sR = reinterpret_cast<float*>(sm + B_OFF_R);
  float* sPar = reinterpret_cast<float*>(sm + B_OFF_PAR);
  float2* sRed = reinterpret_cast<float2*>(sm + B_OFF_RED);
  uint64_t* w_full = reinterpret_cast<uint64_t*>(sm + B_OFF_BAR);
  uint64_t* w_empty = w_full + RING;
  uint64_t* a_ready = w_empty + RING;      // [4]: 0 dqkv' bands (+ ds1' pre-load), 1 ds2, 2 du (+ ds2 pre-load), 3 ds1
  uint64_t* acc_full = a_ready + 4;        // [4]: 0 dz, 1 du_pre, 2 dy, 3 dO
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 4);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_tiles = (a.n + CT - 1) / CT;
  const bool up = a.mode & 1, chain = (a.mode & 2) != 0;
  if (threadIdx.x == 0) {
    for (int i = 0; i < RING; ++i) { tc::mbar_init(&w_full[i], 1); tc::mbar_init(&w_empty[i], 1); }
    for (int i = 0; i < 4; ++i) { tc::mbar_init(&a_ready[i], NCOMP / 32); tc::mbar_init(&acc_full[i], 1); }
  }
  if (warp == 9) tc::tmem_alloc(tmem_slot, 512);
  if (threadIdx.x < NCOMP && chain) sPar[threadIdx.x] = threadIdx.x < 128 ? a.g2[threadIdx.x] : a.g1[threadIdx.x - 128];
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = *tmem_slot;
  gm_pdl_wait();
  gm_pdl_trigger();

  if (warp == 8) {
    if (lane == 0) {
      uint32_t cnt = 0;
      auto push = [&](const uint8_t* img, int nblk) {
        for (int b = 0; b < nblk; ++b, ++cnt) {
          const uint32_t slot = cnt % RING, ph = (cnt / RING) & 1;
          tc::mbar_wait(&w_empty[slot], ph ^ 1);
          tc::mbar_expect_tx(&w_full[slot], BLK);
          tc::bulk_g2s(sRing + slot * BLK, img + (size_t)b * BLK, BLK, &w_full[slot]);
        }
      };
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        if (up) push(a.Win_up, 6);
        if (chain) { push(a.W2, 4); push(a.W1, 4); push(a.Wo, 2); }
      }
    }
  } else if (warp == 9) {
    if (lane == 0) {
      const uint32_t idesc = tc::make_idesc_bf16(128, 64, 0, 1);     // A K-major, B MN-major, N = 64
      uint32_t cnt = 0;
      // one ring block = W[128 K-rows (out features) x 64 N-cols (in features)]: 8 MMAs of K = 16 against the
      // K = 128 operand band at `a_addr` (two 64-column blocks)
      auto mma_block = [&](uint32_t a_addr, uint32_t tcol, bool acc) {
        const uint32_t slot = cnt % RING, ph = (cnt / RING) & 1;
        tc::mbar_wait(&w_full[slot], ph);
        tc::fence_after_sync();
        const uint32_t b_addr = tc::smem_u32(sRing + slot * BLK);
#pragma unroll
        for (int j = 0; j < 8; ++j)
          tc::mma_bf16(tmem + tcol, tc::make_desc(a_addr + (uint32_t)(j >> 2) * BLK + (j & 3) * 32, 16, tc::ATOM_BYTES),
                       tc::make_desc(b_addr + (uint32_t)j * 2 * tc::ATOM_BYTES, BLK, tc::ATOM_BYTES), idesc, acc || j > 0);
        tc::mma_commit(&w_empty[slot]);
        ++cnt;
      };
      const uint32_t A = tc::smem_u32(sA), A2 = tc::smem_u32(sA2), A3 = tc::smem_u32(sT);
      int it = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
        const uint32_t par = it & 1;
        if (up) {
          tc::mbar_wait(&a_ready[0], par);            // dz acc [0,128) (pre-loaded with ds1') += dqkv' Win'
          tc::fence_after_sync();
          for (int b = 0; b < 3; ++b)
            for (int c = 0; c < 2; ++c) mma_block(b == 0 ? A : (b == 1 ? A2 : A3), c * 64, true);
          tc::mma_commit(&acc_full[0]);
        }
        if (chain) {
          tc::mbar_wait(&a_ready[1], par);            // du_pre acc [128,384) = ds2 W2
          tc::fence_after_sync();
          for (int c = 0; c < 4; ++c) mma_block(A, 128 + c * 64, false);
          tc::mma_commit(&acc_full[1]);
          tc::mbar_wait(&a_ready[2], par);            // dy acc [0,128) (pre-loaded with ds2) += du W1
          tc::fence_after_sync();
          for (int b = 0; b < 2; ++b)
            for (int c = 0; c < 2; ++c) mma_block(A + b * 2 * BLK, c * 64, true);
          tc::mma_commit(&acc_full[2]);
          tc::mbar_wait(&a_ready[3], par);            // dO acc [384,512) = ds1 Wo
          tc::fence_after_sync();
          for (int c = 0; c < 2; ++c) mma_block(A, 384 + c * 64, false);
          tc::mma_commit(&acc_full[3]);
        }
      }
    }
  } else {
    const int q = warp & 3, hsel = warp >> 2;
    const int r = q * 32 + lane;
    const int c0 = hsel * 64;
    const uint32_t t_lane = tmem + ((uint32_t)(q * 32) << 16);
    float* myR = sR + r * R_LD + c0;
    const int cs_col = threadIdx.x & 127, cs_row0 = (threadIdx.x >> 7) * 64;     // column-sum ownership
    float acc_dg2 = 0.f, acc_db2 = 0.f, acc_dg1 = 0.f, acc_db1 = 0.f;
    auto load_rows_f32 = [&](const float* src, int row0, int m) {
      for (int i = threadIdx.x; i < CT * 32; i += NCOMP) {
        const int rr = i >> 5, c4 = i & 31;
        float* dst = sR + rr * R_LD + c4 * 4;
        if (rr < m) cp_async16(dst, src + (int64_t)(row0 + rr) * 128 + c4 * 4);
        else *reinterpret_cast<float4*>(dst) = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    };
    // 128 columns [col0, col0+128) of a bf16 row-major [n, ld] tensor -> swizzled two-block operand tile
    auto load_band = [&](uint8_t* tile, const __nv_bfloat16* src, int ld, int col0, int row0, int m) {
      for (int i = threadIdx.x; i < CT * 16; i += NCOMP) {
        const int rr = i >> 4, c = i & 15;
        uint8_t* dst = tile + op_off(rr, c);
        if (rr < m) cp_async16(dst, src + (int64_t)(row0 + rr) * ld + col0 + c * 8);
        else *reinterpret_cast<uint4*>(dst) = make_uint4(0u, 0u, 0u, 0u);
      }
    };
    auto load_u = [&](int sb, int row0, int m) {       // u columns [64 sb, 64 sb + 64) -> half (sb & 1) of the staging tile
      uint8_t* half = sT + (sb & 1) * UH_BYTES;
      for (int i = threadIdx.x; i < CT * 8; i += NCOMP) {
        const int rr = i >> 3, c = i & 7;
        uint8_t* dst = half + rr * UH_LD + c * 16;
        if (rr < m) cp_async16(dst, a.u16 + (int64_t)(row0 + rr) * 256 + sb * 64 + c * 8);
        else *reinterpret_cast<uint4*>(dst) = make_uint4(0u, 0u, 0u, 0u);
      }
    };
    auto load_o = [&](int row0, int m) {
      for (int i = threadIdx.x; i < CT * 16; i += NCOMP) {
        const int rr = i >> 4, c = i & 15;
        uint8_t* dst = sT + rr * ST_LD + c * 16;
        if (rr < m) cp_async16(dst, a.attn16 + (int64_t)(row0 + rr) * 128 + c * 8);
        else *reinterpret_cast<uint4*>(dst) = make_uint4(0u, 0u, 0u, 0u);
      }
    };
    auto prefetch_early = [&](int tile) {              // R and the first two dqkv' bands
      const int row0 = tile * CT, m = min(CT, a.n - row0);
      load_rows_f32(up ? a.ds1_up : a.dz_in, row0, m);
      if (up) {
        load_band(sA, a.dqkv_up, 384, 0, row0, m);
        load_band(sA2, a.dqkv_up, 384, 128, row0, m);
      }
      cp_commit();
    };
    auto prefetch_late = [&](int tile) {               // third band into the staging tile
      const int row0 = tile * CT, m = min(CT, a.n - row0);
      if (up) load_band(sT, a.dqkv_up, 384, 256, row0, m);
      cp_commit();
    };
    // column sums of R over this thread's 64 tile rows (two threads per column)
    auto column_sum = [&](float& acc) {
      float s0 = 0.f, s1 = 0.f;
#pragma unroll 8
      for (int rr = cs_row0; rr < cs_row0 + 64; rr += 2) {
        s0 += sR[rr * R_LD + cs_col];
        s1 += sR[(rr + 1) * R_LD + cs_col];
      }
      acc += s0 + s1;
    };
    // R <- v (this thread's half row)
    auto put_row = [&](const float* v) {
#pragma unroll
      for (int c = 0; c < 64; c += 4) *reinterpret_cast<float4*>(myR + c) = make_float4(v[c], v[c + 1], v[c + 2], v[c + 3]);
    };
    int it = 0;
    if ((int)blockIdx.x < n_tiles) { prefetch_early(blockIdx.x); prefetch_late(blockIdx.x); }
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
      const uint32_t par = it & 1;
      const int row0 = tile * CT, m = min(CT, a.n - row0);
      const int grow = row0 + r;
      const int ntile = tile + gridDim.x;
      float v[64], xh[64];
      cp_wait<0>();
      compute_sync();                                  // R = ds1' (or dz), dqkv' bands complete
#pragma unroll
      for (int c = 0; c < 64; c += 4) {
        const float4 t4 = *reinterpret_cast<const float4*>(myR + c);
        v[c] = t4.x; v[c + 1] = t4.y; v[c + 2] = t4.z; v[c + 3] = t4.w;
      }
      if (up) {
        tmem_st32(t_lane + c0, v);                     // accumulator starts at ds1' (fp32 residual-gradient path)
        tmem_st32(t_lane + c0 + 32, v + 32);
        tmem_st_wait();
        warp_arrive(&a_ready[0]);
      }
      compute_sync();                                  // every thread has read its R row
      if (chain) { load_rows_f32(a.s2, row0, m); cp_commit(); }
      if (up) {
        tc::mbar_wait(&acc_full[0], par);
        tc::fence_after_sync();
        tc::tmem_ld32(t_lane + c0, v);
        tc::tmem_ld32(t_lane + c0 + 32, v + 32);
        tc::tmem_ld_wait();
      }
      if (!chain) {
        // ---------------- input gradient of the stack: dx = dz
#pragma unroll
        for (int c = 0; c < 64; c += 4) *reinterpret_cast<float4*>(myR + c) = make_float4(v[c], v[c + 1], v[c + 2], v[c + 3]);
        compute_sync();
        store_rows_f32(sR, a.dx, row0, m);
        compute_sync();
        if (ntile < n_tiles) { prefetch_early(ntile); prefetch_late(ntile); }
        continue;
      }
      // ---------------- LayerNorm-2 backward
      load_u(0, row0, m);
      cp_commit();
      float2 st = r < m ? __ldg(reinterpret_cast<const float2*>(a.st2 + 2 * (int64_t)grow)) : make_float2(0.f, 0.f);
      cp_wait<0>();
      compute_sync();                                  // s2 in R, u sub-band 0 in the staging tile
      ln_backward_row(v, xh, myR, sPar + c0, st.x, st.y, sRed, r, hsel);       // v = ds2, xh = dz, R = dz * xhat2
#pragma unroll
      for (int c = 0; c < 64; c += 8) *reinterpret_cast<uint4*>(sA + op_off(r, (c0 + c) >> 3)) = pack8f(v + c);
      tmem_st32(t_lane + c0, v);                       // dy accumulator starts at ds2
      tmem_st32(t_lane + c0 + 32, v + 32);
      tmem_st_wait();
      warp_arrive(&a_ready[1]);
      compute_sync();
      column_sum(acc_dg2);                             // d_gamma2 += sum dz * xhat2
      store_rows_op<2>(sA, a.ds2_16, 128, row0, m);
      compute_sync();
      put_row(xh);
      compute_sync();
      column_sum(acc_db2);                             // d_beta2 += sum dz
      compute_sync();
      load_rows_f32(a.s1, row0, m);
      cp_commit();
      // ---------------- du = (ds2 W2) * gelu'(u): four 64-column sub-bands, 32 columns per thread
      tc::mbar_wait(&acc_full[1], par);
      tc::fence_after_sync();
      // (the dy accumulator pre-load above is ordered before a_ready[2] below)
#pragma unroll 1
      for (int sb = 0; sb < 4; ++sb) {
        if (sb + 1 < 4) { load_u(sb + 1, row0, m); cp_commit(); cp_wait<1>(); }
        else cp_wait<0>();
        compute_sync();                                // u sub-band sb visible
        float w[32];
        tc::tmem_ld32(t_lane + 128 + sb * 64 + hsel * 32, w);
        tc::tmem_ld_wait();
        const uint8_t* urow = sT + (sb & 1) * UH_BYTES + r * UH_LD + hsel * 64;
#pragma unroll
        for (int c = 0; c < 32; c += 8) {
          const uint4 u4 = *reinterpret_cast<const uint4*>(urow + c * 2);
          const __nv_bfloat162* u2 = reinterpret_cast<const __nv_bfloat162*>(&u4);
          float d8[8];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float2 uf = __bfloat1622float2(u2[e]);
            d8[2 * e] = w[c + 2 * e] * gelu_grad_f(uf.x);
            d8[2 * e + 1] = w[c + 2 * e + 1] * gelu_grad_f(uf.y);
          }
          *reinterpret_cast<uint4*>(sG + (uint32_t)sb * BLK + tc::swz(r, hsel * 4 + (c >> 3))) = pack8f(d8);
        }
        if (sb == 3) {
          warp_arrive(&a_ready[2]);
        }
        compute_sync();                                // sub-band tile free again, du columns complete
      }
      store_rows_op<4>(sG, a.du16, 256, row0, m);
      load_o(row0, m);
      cp_commit();
      // ---------------- LayerNorm-1 backward on dy
      tc::mbar_wait(&acc_full[2], par);
      tc::fence_after_sync();
      tc::tmem_ld32(t_lane + c0, v);
      tc::tmem_ld32(t_lane + c0 + 32, v + 32);
      tc::tmem_ld_wait();
      st = r < m ? __ldg(reinterpret_cast<const float2*>(a.st1 + 2 * (int64_t)grow)) : make_float2(0.f, 0.f);
      cp_wait<0>();
      compute_sync();                                  // s1 in R, O rows in the staging tile; du16 copy-out finished
      ln_backward_row(v, xh, myR, sPar + 128 + c0, st.x, st.y, sRed, r, hsel);  // v = ds1, xh = dy, R = dy * xhat1
#pragma unroll
      for (int c = 0; c < 64; c += 8) *reinterpret_cast<uint4*>(sA + op_off(r, (c0 + c) >> 3)) = pack8f(v + c);
      warp_arrive(&a_ready[3]);
      compute_sync();
      column_sum(acc_dg1);                             // d_gamma1 += sum dy * xhat1
      store_rows_op<2>(sA, a.ds1_16, 128, row0, m);
      compute_sync();
      put_row(xh);
      compute_sync();
      column_sum(acc_db1);                             // d_beta1 += sum dy
      compute_sync();
      put_row(v);
      compute_sync();
      store_rows_f32(sR, a.ds1, row0, m);              // fp32 ds1: the residual-gradient term of the layer below
      compute_sync();
      // ---------------- dO = ds1 Wo ; D = dO . O per head
      tc::mbar_wait(&acc_full[3], par);
      tc::fence_after_sync();
      if (ntile < n_tiles) prefetch_early(ntile);      // R, A, A2 are free
      tc::tmem_ld32(t_lane + 384 + c0, v);
      tc::tmem_ld32(t_lane + 384 + c0 + 32, v + 32);
      tc::tmem_ld_wait();
      {
        uint8_t* orow = sT + r * ST_LD + c0 * 2;
        float dh[4];
#pragma unroll
        for (int h = 0; h < 4; ++h) {
          float d = 0.f;
#pragma unroll
          for (int c = 0; c < 16; c += 8) {
            const uint4 o4 = *reinterpret_cast<const uint4*>(orow + (h * 16 + c) * 2);
            const __nv_bfloat162* o2 = reinterpret_cast<const __nv_bfloat162*>(&o4);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float2 of = __bfloat1622float2(o2[e]);
              d = fmaf(v[h * 16 + c + 2 * e], of.x, d);
              d = fmaf(v[h * 16 + c + 2 * e + 1], of.y, d);
            }
            *reinterpret_cast<uint4*>(orow + (h * 16 + c) * 2) = pack8f(v + h * 16 + c);      // dO over O, in place
          }
          dh[h] = d;
        }
        if (r < m) *reinterpret_cast<float4*>(a.dd + (int64_t)grow * 8 + hsel * 4) = make_float4(dh[0], dh[1], dh[2], dh[3]);
      }
      compute_sync();
      store_rows_st(sT, a.dO16, 128, 0, row0, m);
      compute_sync();
      if (ntile < n_tiles) prefetch_late(ntile);
    }
    if (chain) {
      atomicAdd(a.d_g2 + cs_col, acc_dg2);
      atomicAdd(a.d_be2 + cs_col, acc_db2);
      atomicAdd(a.d_g1 + cs_col, acc_dg1);
      atomicAdd(a.d_be1 + cs_col, acc_db1);
    }
    tc::fence_before_sync();
  }
  __syncthreads();
  if (warp == 9) {
    tc::fence_after_sync();
    tc::tmem_free(tmem, 512);
  }
}

}  // namespace

extern "C" int geomae_sra_chain_fwd(const geomae_chain_fwd_args* p, void* stream) {
  GM_REQUIRE(p, "sra_chain_fwd: null argument");
  GM_REQUIRE(p->n_tokens >= 0 && p->n_tokens < ((int64_t)1 << 31) - 256, "sra_chain_fwd: bad token count");
  GM_REQUIRE((p->mode & 3) != 0 && (p->mode & ~3) == 0, "sra_chain_fwd: mode must be 1 (chain), 2 (next in-proj) or 3");
  if (p->n_tokens == 0) return GEOMAE_OK;
  const bool chain = p->mode & 1, next = (p->mode & 2) != 0;
  GM_REQUIRE(p->x, "sra_chain_fwd: x is null");
  if (chain)
    GM_REQUIRE(p->attn && p->p_out_proj && p->p_lin1 && p->p_lin2 && p->out_proj_b && p->lin1_b && p->lin2_b &&
                   p->norm1_w && p->norm1_b && p->norm2_w && p->norm2_b && p->s1 && p->st1 && p->s2 && p->st2 && p->z &&
                   p->y16 && p->u16 && p->g16,
               "sra_chain_fwd: the layer chain needs attn, packed weights, biases, norms and every saved-tensor buffer");
  if (next)
    GM_REQUIRE(p->p_in_proj_next && p->in_proj_b_next && p->pos_table && p->tok_cell_next && p->xp16_next &&
                   p->xb16_next && p->qkv16_next,
               "sra_chain_fwd: the next in-projection needs packed weights, bias, position table, cells and outputs");
  FwdArgs a;
  a.n = (int)p->n_tokens; a.mode = p->mode; a.x = p->x; a.attn = (const __nv_bfloat16*)p->attn;
  a.Wo = (const uint8_t*)p->p_out_proj; a.W1 = (const uint8_t*)p->p_lin1; a.W2 = (const uint8_t*)p->p_lin2;
  a.Win = (const uint8_t*)p->p_in_proj_next;
  a.bo = p->out_proj_b; a.b1 = p->lin1_b; a.b2 = p->lin2_b; a.bin = p->in_proj_b_next;
  a.g1 = p->norm1_w; a.be1 = p->norm1_b; a.g2 = p->norm2_w; a.be2 = p->norm2_b; a.eps = p->ln_eps;
  a.pos = p->pos_table; a.cell_next = p->tok_cell_next;
  a.s1 = p->s1; a.st1 = p->st1; a.s2 = p->s2; a.st2 = p->st2; a.z = p->z;
  a.y16 = (__nv_bfloat16*)p->y16; a.u16 = (__nv_bfloat16*)p->u16; a.g16 = (__nv_bfloat16*)p->g16;
  a.xp16 = (__nv_bfloat16*)p->xp16_next; a.xb16 = (__nv_bfloat16*)p->xb16_next; a.qkv16 = (__nv_bfloat16*)p->qkv16_next;
  static bool configured = false;
  if (!configured) {
    GM_CUDA(cudaFuncSetAttribute(k_sra_chain_fwd, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    configured = true;
  }
  const int n_tiles = gm_div_up(a.n, CT);
  const int grid = n_tiles < GM_NUM_SMS ? n_tiles : GM_NUM_SMS;
  GM_CUDA(gm_launch_pdl(k_sra_chain_fwd, dim3(grid), dim3(NTHR), (size_t)SMEM_BYTES, (cudaStream_t)stream, a));
  return GEOMAE_OK;
}

extern "C" int geomae_sra_chain_bwd(const geomae_chain_bwd_args* p, void* stream) {
  GM_REQUIRE(p, "sra_chain_bwd: null argument");
  GM_REQUIRE(p->n_tokens >= 0 && p->n_tokens < ((int64_t)1 << 31) - 256, "sra_chain_bwd: bad token count");
  GM_REQUIRE((p->mode & 3) != 0 && (p->mode & ~3) == 0, "sra_chain_bwd: mode must be 1, 2 or 3");
  if (p->n_tokens == 0) return GEOMAE_OK;
  const bool up = p->mode & 1, chain = (p->mode & 2) != 0;
  if (up) GM_REQUIRE(p->dqkv16_up && p->ds1_up && p->p_in_proj_up, "sra_chain_bwd: the in-projection backward needs dqkv, ds1 and packed weights of the layer above");
  else GM_REQUIRE(p->dz_in, "sra_chain_bwd: dz_in is null");
  if (chain)
    GM_REQUIRE(p->s2 && p->st2 && p->s1 && p->st1 && p->u16 && p->attn16 && p->p_lin2 && p->p_lin1 && p->p_out_proj &&
                   p->norm2_w && p->norm1_w && p->ds2_16 && p->du16 && p->ds1_16 && p->dattn16 && p->ds1 && p->dd &&
                   p->g_norm2_w && p->g_norm2_b && p->g_norm1_w && p->g_norm1_b,
               "sra_chain_bwd: the layer chain needs every saved tensor, packed weight, output and gradient buffer");
  else GM_REQUIRE(p->dx, "sra_chain_bwd: dx is null");
  BwdArgs a;
  a.n = (int)p->n_tokens; a.mode = p->mode;
  a.dqkv_up = (const __nv_bfloat16*)p->dqkv16_up; a.ds1_up = p->ds1_up; a.Win_up = (const uint8_t*)p->p_in_proj_up; a.dz_in = p->dz_in;
  a.s2 = p->s2; a.st2 = p->st2; a.s1 = p->s1; a.st1 = p->st1;
  a.u16 = (const __nv_bfloat16*)p->u16; a.attn16 = (const __nv_bfloat16*)p->attn16;
  a.W2 = (const uint8_t*)p->p_lin2; a.W1 = (const uint8_t*)p->p_lin1; a.Wo = (const uint8_t*)p->p_out_proj;
  a.g2 = p->norm2_w; a.g1 = p->norm1_w;
  a.ds2_16 = (__nv_bfloat16*)p->ds2_16; a.du16 = (__nv_bfloat16*)p->du16; a.ds1_16 = (__nv_bfloat16*)p->ds1_16;
  a.dO16 = (__nv_bfloat16*)p->dattn16; a.ds1 = p->ds1; a.dd = p->dd; a.dx = p->dx;
  a.d_g2 = p->g_norm2_w; a.d_be2 = p->g_norm2_b; a.d_g1 = p->g_norm1_w; a.d_be1 = p->g_norm1_b;
  static bool configured = false;
  if (!configured) {
    GM_CUDA(cudaFuncSetAttribute(k_sra_chain_bwd, cudaFuncAttributeMaxDynamicSharedMemorySize, B_SMEM_BYTES));
    configured = true;
  }
  const int n_tiles = gm_div_up(a.n, CT);
  const int grid = n_tiles < GM_NUM_SMS ? n_tiles : GM_NUM_SMS;
  GM_CUDA(gm_launch_pdl(k_sra_chain_bwd, dim3(grid), dim3(NTHR), (size_t)B_SMEM_BYTES, (cudaStream_t)stream, a));
  return GEOMAE_OK;
}
