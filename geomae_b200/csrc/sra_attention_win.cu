// Sparse Regional Attention core, window-resident variant for the all-bf16 path (SURVEY.md §8 row a18,
// models/sst/sst_basic_block.py:26-61): one WARP owns one (window, head) pair end to end.
//
// A CTA is 8 warps = the 8 heads of one CSR window; CTAs walk the window list.  The warp stages the window's K and V
// head slices (L x 16 bf16 each, L <= 144) into its private shared-memory slab, then for every 16-query tile computes
//     S = Q K^T (mma.sync m16n8k16, k = head_dim = 16)  ->  online softmax in the accumulator fragments  ->  O += P V
// with Q fragments loaded straight from global (prefetched one tile ahead) and O / LSE stored straight from the
// fragments — no CTA-wide barrier, no block-diagonal packing of foreign windows, and the only masked columns are the
// padding of the window's last 16-key block.  The round-1 kernel (sra_attention_tc.cu) gave a CTA 64 consecutive CSR
// positions and walked every key block any of their windows touched: ~30 % of the score entries it computed were valid
// and its CTA-wide staging barriers left it latency-bound (35 us per 24.5 k-token decoder layer, 12 % issue utilisation).
// The backward runs the same structure twice per warp: queries as rows (dQ), then keys as rows (dK, dV), recomputing
// S from the staged operands; D = dO . O per (token, head) comes from the chain kernel (sra_chain.cu).
#include <cuda_bf16.h>

#include "common.cuh"

namespace {

constexpr int NH = 8;               // heads = warps per CTA
constexpr int LMAX = 144;           // tokens per window (12 x 12 cells)
constexpr int ROW = 32;             // bytes of one head slice (16 bf16)
constexpr float LOG2E = 1.4426950408889634f;
constexpr float QSCALE = 0.25f * LOG2E;     // 1/sqrt(head_dim), exp2 domain

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  const __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<const uint32_t*>(&v);
}
__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// D[16x8] += A[16x16] * B[16x8], bf16 operands, fp32 accumulate
__device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// A staged slab holds one 32-byte row per window position; the two 16-byte halves of rows 4..7 (mod 8) are swapped
// so that both the 4-byte fragment loads and ldmatrix see eight different bank groups.
__device__ __forceinline__ uint32_t slab_off(int row, int half) { return (uint32_t)row * ROW + (uint32_t)((half ^ (row >> 2)) & 1) * 16; }

// B fragment (k = the 16 dims, n = rows r0 + g): B[k][n] = slab[r0 + n][k]
__device__ __forceinline__ void ldb(uint32_t& b0, uint32_t& b1, const uint8_t* slab, int r0) {
  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  b0 = *reinterpret_cast<const uint32_t*>(slab + slab_off(r0 + g, 0) + 4 * t);
  b1 = *reinterpret_cast<const uint32_t*>(slab + slab_off(r0 + g, 1) + 4 * t);
}
// A fragment of rows r0..r0+15 from a slab
__device__ __forceinline__ void lda_slab(uint32_t (&a)[4], const uint8_t* slab, int r0) {
  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  a[0] = *reinterpret_cast<const uint32_t*>(slab + slab_off(r0 + g, 0) + 4 * t);
  a[1] = *reinterpret_cast<const uint32_t*>(slab + slab_off(r0 + g + 8, 0) + 4 * t);
  a[2] = *reinterpret_cast<const uint32_t*>(slab + slab_off(r0 + g, 1) + 4 * t);
  a[3] = *reinterpret_cast<const uint32_t*>(slab + slab_off(r0 + g + 8, 1) + 4 * t);
}
// transposed B fragments (k = rows r0..r0+15 of the slab, n = dims 0..7 | 8..15): r[0], r[1] for dims 0..7, r[2], r[3] for 8..15
__device__ __forceinline__ void ldsm_bt(uint32_t (&r)[4], const uint8_t* slab, int r0) {
  const int lane = threadIdx.x & 31;
  const int mi = lane >> 3, rr = lane & 7;
  const uint32_t addr = smem_addr(slab + slab_off(r0 + (mi & 1) * 8 + rr, mi >> 1));
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
// stage the head slices (16 bf16 at column `col`) of the window's rows into a slab; rows [L, Lpad) are zero
__device__ __forceinline__ void stage_slab(uint8_t* slab, const __nv_bfloat16* __restrict__ src, int ld, int col,
                                           const int32_t* sTok, int L, int Lpad) {
  const int lane = threadIdx.x & 31;
  for (int i = lane; i < Lpad; i += 32) {
    uint4 lo = make_uint4(0u, 0u, 0u, 0u), hi = lo;
    if (i < L) {
      const uint4* p = reinterpret_cast<const uint4*>(src + (int64_t)sTok[i] * ld + col);
      lo = __ldg(p);
      hi = __ldg(p + 1);
    }
    *reinterpret_cast<uint4*>(slab + slab_off(i, 0)) = lo;
    *reinterpret_cast<uint4*>(slab + slab_off(i, 1)) = hi;
  }
}
// A fragment of window rows r0+g, r0+g+8 straight from global rows (zeros past the window)
__device__ __forceinline__ void lda_global(uint32_t (&a)[4], const __nv_bfloat16* __restrict__ src, int ld, int col,
                                           const int32_t* sTok, int r0, int L) {
  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  a[0] = a[1] = a[2] = a[3] = 0u;
  if (r0 + g < L) {
    const uint32_t* p = reinterpret_cast<const uint32_t*>(src + (int64_t)sTok[r0 + g] * ld + col) + t;
    a[0] = __ldg(p);
    a[2] = __ldg(p + 4);
  }
  if (r0 + g + 8 < L) {
    const uint32_t* p = reinterpret_cast<const uint32_t*>(src + (int64_t)sTok[r0 + g + 8] * ld + col) + t;
    a[1] = __ldg(p);
    a[3] = __ldg(p + 4);
  }
}

struct WarpSmem {
  uint8_t x[LMAX * ROW];            // K (forward, dQ pass) | Q (dK/dV pass)
  uint8_t y[LMAX * ROW];            // V (forward, dQ pass) | dO (dK/dV pass)
  int32_t tok[LMAX];
  float lse2[LMAX];                 // backward: log2-domain LSE and D of the window's queries
  float dd[LMAX];
};

__device__ __forceinline__ int window_count(const int32_t* __restrict__ win_tok, const int32_t* __restrict__ tok_win, int n) {
  return __ldg(tok_win + __ldg(win_tok + n - 1)) + 1;     // the last CSR position belongs to the last window
}

constexpr int CHUNKS = 5;           // a window has at most 9 16-row tiles: 5 chunks of 2

// one 16-query tile against one 16-key block: scores, online softmax update, O += P V
struct FwdTile {
  float mx0 = -INFINITY, mx1 = -INFINITY, ls0 = 0.f, ls1 = 0.f;
  float o0[4] = {0.f, 0.f, 0.f, 0.f}, o1[4] = {0.f, 0.f, 0.f, 0.f};
};
__device__ __forceinline__ void fwd_block(FwdTile& f, const uint32_t (&qa)[4], const uint8_t* sK, const uint8_t* sV, int kb,
                                          int L, bool last_padded) {
  const int t = threadIdx.x & 3;
  float s0[4] = {0.f, 0.f, 0.f, 0.f}, s1[4] = {0.f, 0.f, 0.f, 0.f};
  uint32_t b0, b1;
  ldb(b0, b1, sK, kb * 16);
  mma16816(s0, qa, b0, b1);
  ldb(b0, b1, sK, kb * 16 + 8);
  mma16816(s1, qa, b0, b1);
#pragma unroll
  for (int i = 0; i < 4; ++i) { s0[i] *= QSCALE; s1[i] *= QSCALE; }
  if (last_padded) {                                     // padding keys of the window's last block
    const int k0 = kb * 16 + 2 * t;
    if (k0 >= L) { s0[0] = -INFINITY; s0[2] = -INFINITY; }
    if (k0 + 1 >= L) { s0[1] = -INFINITY; s0[3] = -INFINITY; }
    if (k0 + 8 >= L) { s1[0] = -INFINITY; s1[2] = -INFINITY; }
    if (k0 + 9 >= L) { s1[1] = -INFINITY; s1[3] = -INFINITY; }
  }
  float m0 = fmaxf(fmaxf(s0[0], s0[1]), fmaxf(s1[0], s1[1]));
  float m1 = fmaxf(fmaxf(s0[2], s0[3]), fmaxf(s1[2], s1[3]));
  m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 1));
  m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 2));
  m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 1));
  m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 2));
  const float n0 = fmaxf(f.mx0, m0), n1 = fmaxf(f.mx1, m1);       // finite: key 0 of block 0 is always a real key
  const float c0 = ex2(f.mx0 - n0), c1 = ex2(f.mx1 - n1);
  f.mx0 = n0; f.mx1 = n1;
  const float p00 = ex2(s0[0] - n0), p01 = ex2(s0[1] - n0), p02 = ex2(s1[0] - n0), p03 = ex2(s1[1] - n0);
  const float p10 = ex2(s0[2] - n1), p11 = ex2(s0[3] - n1), p12 = ex2(s1[2] - n1), p13 = ex2(s1[3] - n1);
  f.ls0 = f.ls0 * c0 + ((p00 + p01) + (p02 + p03));
  f.ls1 = f.ls1 * c1 + ((p10 + p11) + (p12 + p13));
  f.o0[0] *= c0; f.o0[1] *= c0; f.o1[0] *= c0; f.o1[1] *= c0;
  f.o0[2] *= c1; f.o0[3] *= c1; f.o1[2] *= c1; f.o1[3] *= c1;
  const uint32_t pa[4] = {pack_bf16(p00, p01), pack_bf16(p10, p11), pack_bf16(p02, p03), pack_bf16(p12, p13)};
  uint32_t vb[4];
  ldsm_bt(vb, sV, kb * 16);
  mma16816(f.o0, pa, vb[0], vb[1]);
  mma16816(f.o1, pa, vb[2], vb[3]);
}
__device__ __forceinline__ void fwd_store(FwdTile& f, int mt, int L, const int32_t* sTok, int h, __nv_bfloat16* out, float* lse) {
  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  float ls0 = f.ls0, ls1 = f.ls1;
  ls0 += __shfl_xor_sync(0xffffffffu, ls0, 1); ls0 += __shfl_xor_sync(0xffffffffu, ls0, 2);
  ls1 += __shfl_xor_sync(0xffffffffu, ls1, 1); ls1 += __shfl_xor_sync(0xffffffffu, ls1, 2);
  const float i0 = 1.0f / ls0, i1 = 1.0f / ls1;
  const int r0 = mt * 16 + g, r1 = r0 + 8;
  if (r0 < L) {
    const int64_t tok = sTok[r0];
    uint32_t* o = reinterpret_cast<uint32_t*>(out + tok * 128 + h * 16) + t;
    o[0] = pack_bf16(f.o0[0] * i0, f.o0[1] * i0);
    o[4] = pack_bf16(f.o1[0] * i0, f.o1[1] * i0);
    if (t == 0) lse[tok * NH + h] = (f.mx0 + log2f(ls0)) * 0.6931471805599453f;
  }
  if (r1 < L) {
    const int64_t tok = sTok[r1];
    uint32_t* o = reinterpret_cast<uint32_t*>(out + tok * 128 + h * 16) + t;
    o[0] = pack_bf16(f.o0[2] * i1, f.o0[3] * i1);
    o[4] = pack_bf16(f.o1[2] * i1, f.o1[3] * i1);
    if (t == 0) lse[tok * NH + h] = (f.mx1 + log2f(ls1)) * 0.6931471805599453f;
  }
}

// Work item = (window, chunk of two 16-query tiles): the largest window (9 tiles) is five items, so no CTA is stuck
// behind an 81-block chain; the two tiles of a chunk run interleaved (two independent MMA -> softmax -> MMA chains).
__global__ void __launch_bounds__(256) k_sra_win_fwd(const __nv_bfloat16* __restrict__ qkv, int n,
                                                     const int32_t* __restrict__ win_ptr,
                                                     const int32_t* __restrict__ win_tok,
                                                     const int32_t* __restrict__ tok_win, __nv_bfloat16* __restrict__ out,
                                                     float* __restrict__ lse) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  const int lane = threadIdx.x & 31, h = threadIdx.x >> 5;
  WarpSmem& ws = reinterpret_cast<WarpSmem*>(smem_raw)[h];
  gm_pdl_wait();
  gm_pdl_trigger();
  const int W = window_count(win_tok, tok_win, n);
  const int mt0 = blockIdx.y * 2;                        // this CTA's chunk of two query tiles, for every window it visits
  // 32 windows per round: one batched read of their extents, then only the windows that HAVE this chunk are visited
  for (int base = blockIdx.x; base < W; base += gridDim.x * 32) {
    const int wl = base + lane * gridDim.x;
    int beg_l = 0, len_l = 0;
    if (wl < W) { beg_l = __ldg(win_ptr + wl); len_l = __ldg(win_ptr + wl + 1) - beg_l; }
    unsigned todo = __ballot_sync(0xffffffffu, ((len_l + 15) >> 4) > mt0);
  while (todo) {
    const int src = __ffs(todo) - 1;
    todo &= todo - 1;
    const int beg = __shfl_sync(0xffffffffu, beg_l, src), L = __shfl_sync(0xffffffffu, len_l, src);
    const int T = (L + 15) >> 4, Lpad = T * 16;
    const bool two = mt0 + 1 < T;
    __syncwarp();                                        // previous item's slabs fully consumed
    for (int i = lane; i < Lpad; i += 32) ws.tok[i] = i < L ? __ldg(win_tok + beg + i) : 0;
    __syncwarp();
    uint32_t qa[4], qb[4];
    lda_global(qa, qkv, 384, h * 16, ws.tok, mt0 * 16, L);
    lda_global(qb, qkv, 384, h * 16, ws.tok, (mt0 + 1) * 16, two ? L : 0);
    stage_slab(ws.x, qkv, 384, 128 + h * 16, ws.tok, L, Lpad);
    stage_slab(ws.y, qkv, 384, 256 + h * 16, ws.tok, L, Lpad);
    __syncwarp();
    FwdTile fa, fb;
    const bool padded = Lpad != L;
    if (two) {
      for (int kb = 0; kb < T; ++kb) {
        const bool lp = padded && kb == T - 1;
        fwd_block(fa, qa, ws.x, ws.y, kb, L, lp);
        fwd_block(fb, qb, ws.x, ws.y, kb, L, lp);
      }
      fwd_store(fa, mt0, L, ws.tok, h, out, lse);
      fwd_store(fb, mt0 + 1, L, ws.tok, h, out, lse);
    } else {
      for (int kb = 0; kb < T; ++kb) fwd_block(fa, qa, ws.x, ws.y, kb, L, padded && kb == T - 1);
      fwd_store(fa, mt0, L, ws.tok, h, out, lse);
    }
  }
  }
}

// Backward.  Pass A: queries as rows -> dQ (K, V staged).  Pass B: keys as rows -> dK, dV (Q, dO staged).  The two passes
// of a window are independent work items.
__global__ void __launch_bounds__(256) k_sra_win_bwd(const __nv_bfloat16* __restrict__ qkv,
                                                     const float* __restrict__ lse, const __nv_bfloat16* __restrict__ d_out,
                                                     const float* __restrict__ dd, int n,
                                                     const int32_t* __restrict__ win_ptr,
                                                     const int32_t* __restrict__ win_tok,
                                                     const int32_t* __restrict__ tok_win, __nv_bfloat16* __restrict__ d_qkv) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  const int lane = threadIdx.x & 31, h = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
  WarpSmem& ws = reinterpret_cast<WarpSmem*>(smem_raw)[h];
  gm_pdl_wait();
  gm_pdl_trigger();
  const int W = window_count(win_tok, tok_win, n);
  // work item = (window, chunk of two 16-row tiles, pass): blockIdx.y = pass * CHUNKS + chunk, windows as in k_sra_win_fwd
  const bool pass_b = (int)blockIdx.y >= CHUNKS;
  const int t_lo = ((int)blockIdx.y - (pass_b ? CHUNKS : 0)) * 2;
  for (int base = blockIdx.x; base < W; base += gridDim.x * 32) {
    const int wl = base + lane * gridDim.x;
    int beg_l = 0, len_l = 0;
    if (wl < W) { beg_l = __ldg(win_ptr + wl); len_l = __ldg(win_ptr + wl + 1) - beg_l; }
    unsigned todo = __ballot_sync(0xffffffffu, ((len_l + 15) >> 4) > t_lo);
  while (todo) {
    const int src = __ffs(todo) - 1;
    todo &= todo - 1;
    const int beg = __shfl_sync(0xffffffffu, beg_l, src), L = __shfl_sync(0xffffffffu, len_l, src);
    const int T = (L + 15) >> 4, Lpad = T * 16;
    const int t_hi = min(t_lo + 2, T);
    __syncwarp();
    for (int i = lane; i < Lpad; i += 32) {
      const int tok = i < L ? __ldg(win_tok + beg + i) : 0;
      ws.tok[i] = tok;
      ws.lse2[i] = i < L ? __ldg(lse + (int64_t)tok * NH + h) * LOG2E : 0.f;
      ws.dd[i] = i < L ? __ldg(dd + (int64_t)tok * NH + h) : 0.f;
    }
    __syncwarp();
    if (!pass_b) {
    // ------------------------------------------------------------ pass A: dQ
    stage_slab(ws.x, qkv, 384, 128 + h * 16, ws.tok, L, Lpad);      // K
    stage_slab(ws.y, qkv, 384, 256 + h * 16, ws.tok, L, Lpad);      // V
    __syncwarp();
    for (int mt = t_lo; mt < t_hi; ++mt) {
      uint32_t qa[4], ga[4];
      lda_global(qa, qkv, 384, h * 16, ws.tok, mt * 16, L);
      lda_global(ga, d_out, 128, h * 16, ws.tok, mt * 16, L);
      const int r0 = mt * 16 + g, r1 = r0 + 8;
      const float l0 = ws.lse2[r0], l1 = ws.lse2[r1], d0 = ws.dd[r0], d1 = ws.dd[r1];
      float a0[4] = {0.f, 0.f, 0.f, 0.f}, a1[4] = {0.f, 0.f, 0.f, 0.f};
      for (int kb = 0; kb < T; ++kb) {
        float s0[4] = {0.f, 0.f, 0.f, 0.f}, s1[4] = {0.f, 0.f, 0.f, 0.f};
        float q0[4] = {0.f, 0.f, 0.f, 0.f}, q1[4] = {0.f, 0.f, 0.f, 0.f};
        uint32_t b0, b1;
        ldb(b0, b1, ws.x, kb * 16);
        mma16816(s0, qa, b0, b1);
        ldb(b0, b1, ws.x, kb * 16 + 8);
        mma16816(s1, qa, b0, b1);
        ldb(b0, b1, ws.y, kb * 16);
        mma16816(q0, ga, b0, b1);                        // dP = dO V^T
        ldb(b0, b1, ws.y, kb * 16 + 8);
        mma16816(q1, ga, b0, b1);
        const int k0 = kb * 16 + 2 * t;
        const bool v0 = k0 < L, v1 = k0 + 1 < L, v8 = k0 + 8 < L, v9 = k0 + 9 < L;
        // dS = P (dP - D), P = exp2(S * scale - lse2); padding keys contribute nothing
        const float e00 = v0 ? ex2(fmaf(s0[0], QSCALE, -l0)) * (q0[0] - d0) : 0.f;
        const float e01 = v1 ? ex2(fmaf(s0[1], QSCALE, -l0)) * (q0[1] - d0) : 0.f;
        const float e02 = v8 ? ex2(fmaf(s1[0], QSCALE, -l0)) * (q1[0] - d0) : 0.f;
        const float e03 = v9 ? ex2(fmaf(s1[1], QSCALE, -l0)) * (q1[1] - d0) : 0.f;
        const float e10 = v0 ? ex2(fmaf(s0[2], QSCALE, -l1)) * (q0[2] - d1) : 0.f;
        const float e11 = v1 ? ex2(fmaf(s0[3], QSCALE, -l1)) * (q0[3] - d1) : 0.f;
        const float e12 = v8 ? ex2(fmaf(s1[2], QSCALE, -l1)) * (q1[2] - d1) : 0.f;
        const float e13 = v9 ? ex2(fmaf(s1[3], QSCALE, -l1)) * (q1[3] - d1) : 0.f;
        const uint32_t ds[4] = {pack_bf16(e00, e01), pack_bf16(e10, e11), pack_bf16(e02, e03), pack_bf16(e12, e13)};
        uint32_t bt[4];
        ldsm_bt(bt, ws.x, kb * 16);                      // dQ += dS K
        mma16816(a0, ds, bt[0], bt[1]);
        mma16816(a1, ds, bt[2], bt[3]);
      }
      if (r0 < L) {
        uint32_t* o = reinterpret_cast<uint32_t*>(d_qkv + (int64_t)ws.tok[r0] * 384 + h * 16) + t;
        o[0] = pack_bf16(a0[0] * 0.25f, a0[1] * 0.25f);
        o[4] = pack_bf16(a1[0] * 0.25f, a1[1] * 0.25f);
      }
      if (r1 < L) {
        uint32_t* o = reinterpret_cast<uint32_t*>(d_qkv + (int64_t)ws.tok[r1] * 384 + h * 16) + t;
        o[0] = pack_bf16(a0[2] * 0.25f, a0[3] * 0.25f);
        o[4] = pack_bf16(a1[2] * 0.25f, a1[3] * 0.25f);
      }
    }
    } else {
    // ------------------------------------------------------------ pass B: dK, dV (rows = keys, columns = queries)
    stage_slab(ws.x, qkv, 384, h * 16, ws.tok, L, Lpad);            // Q
    stage_slab(ws.y, d_out, 128, h * 16, ws.tok, L, Lpad);          // dO
    __syncwarp();
    for (int kt = t_lo; kt < t_hi; ++kt) {
      uint32_t ka[4], va[4];
      lda_global(ka, qkv, 384, 128 + h * 16, ws.tok, kt * 16, L);
      lda_global(va, qkv, 384, 256 + h * 16, ws.tok, kt * 16, L);
      float dk0[4] = {0.f, 0.f, 0.f, 0.f}, dk1[4] = {0.f, 0.f, 0.f, 0.f};
      float dv0[4] = {0.f, 0.f, 0.f, 0.f}, dv1[4] = {0.f, 0.f, 0.f, 0.f};
      for (int qb = 0; qb < T; ++qb) {
        float s0[4] = {0.f, 0.f, 0.f, 0.f}, s1[4] = {0.f, 0.f, 0.f, 0.f};
        float q0[4] = {0.f, 0.f, 0.f, 0.f}, q1[4] = {0.f, 0.f, 0.f, 0.f};
        uint32_t b0, b1;
        ldb(b0, b1, ws.x, qb * 16);
        mma16816(s0, ka, b0, b1);                        // S^T = K Q^T
        ldb(b0, b1, ws.x, qb * 16 + 8);
        mma16816(s1, ka, b0, b1);
        ldb(b0, b1, ws.y, qb * 16);
        mma16816(q0, va, b0, b1);                        // dP^T = V dO^T
        ldb(b0, b1, ws.y, qb * 16 + 8);
        mma16816(q1, va, b0, b1);
        const int c0 = qb * 16 + 2 * t;                  // this thread's query columns: c0, c0+1 | c0+8, c0+9
        const bool v0 = c0 < L, v1 = c0 + 1 < L, v8 = c0 + 8 < L, v9 = c0 + 9 < L;
        const float l0 = ws.lse2[c0], l1 = ws.lse2[c0 + 1], l8 = ws.lse2[c0 + 8], l9 = ws.lse2[c0 + 9];
        const float d0 = ws.dd[c0], d1 = ws.dd[c0 + 1], d8 = ws.dd[c0 + 8], d9 = ws.dd[c0 + 9];
        const float p00 = v0 ? ex2(fmaf(s0[0], QSCALE, -l0)) : 0.f, p01 = v1 ? ex2(fmaf(s0[1], QSCALE, -l1)) : 0.f;
        const float p02 = v8 ? ex2(fmaf(s1[0], QSCALE, -l8)) : 0.f, p03 = v9 ? ex2(fmaf(s1[1], QSCALE, -l9)) : 0.f;
        const float p10 = v0 ? ex2(fmaf(s0[2], QSCALE, -l0)) : 0.f, p11 = v1 ? ex2(fmaf(s0[3], QSCALE, -l1)) : 0.f;
        const float p12 = v8 ? ex2(fmaf(s1[2], QSCALE, -l8)) : 0.f, p13 = v9 ? ex2(fmaf(s1[3], QSCALE, -l9)) : 0.f;
        const uint32_t pp[4] = {pack_bf16(p00, p01), pack_bf16(p10, p11), pack_bf16(p02, p03), pack_bf16(p12, p13)};
        const uint32_t ds[4] = {pack_bf16(p00 * (q0[0] - d0), p01 * (q0[1] - d1)), pack_bf16(p10 * (q0[2] - d0), p11 * (q0[3] - d1)),
                                pack_bf16(p02 * (q1[0] - d8), p03 * (q1[1] - d9)), pack_bf16(p12 * (q1[2] - d8), p13 * (q1[3] - d9))};
        uint32_t bt[4];
        ldsm_bt(bt, ws.x, qb * 16);                      // dK += dS^T Q
        mma16816(dk0, ds, bt[0], bt[1]);
        mma16816(dk1, ds, bt[2], bt[3]);
        ldsm_bt(bt, ws.y, qb * 16);                      // dV += P^T dO
        mma16816(dv0, pp, bt[0], bt[1]);
        mma16816(dv1, pp, bt[2], bt[3]);
      }
      const int r0 = kt * 16 + g, r1 = r0 + 8;
      if (r0 < L) {
        uint32_t* o = reinterpret_cast<uint32_t*>(d_qkv + (int64_t)ws.tok[r0] * 384 + 128 + h * 16) + t;
        o[0] = pack_bf16(dk0[0] * 0.25f, dk0[1] * 0.25f);
        o[4] = pack_bf16(dk1[0] * 0.25f, dk1[1] * 0.25f);
        o[64] = pack_bf16(dv0[0], dv0[1]);               // + 128 columns = 64 words: the v block
        o[68] = pack_bf16(dv1[0], dv1[1]);
      }
      if (r1 < L) {
        uint32_t* o = reinterpret_cast<uint32_t*>(d_qkv + (int64_t)ws.tok[r1] * 384 + 128 + h * 16) + t;
        o[0] = pack_bf16(dk0[2] * 0.25f, dk0[3] * 0.25f);
        o[4] = pack_bf16(dk1[2] * 0.25f, dk1[3] * 0.25f);
        o[64] = pack_bf16(dv0[2], dv0[3]);
        o[68] = pack_bf16(dv1[2], dv1[3]);
      }
    }
    }
  }
  }
}

constexpr int WIN_SMEM = NH * (int)sizeof(WarpSmem);

}  // namespace

// All-bf16 window-resident attention (used by geomae_sra_attention_tc_fwd / _bwd when every operand is bf16).
int gm_sra_win_fwd(const void* qkv, int64_t n, const int32_t* win_ptr, const int32_t* win_tok, const int32_t* tok_win,
                   void* out, float* lse, cudaStream_t st) {
  static bool configured = false;
  if (!configured) {
    GM_CUDA(cudaFuncSetAttribute(k_sra_win_fwd, cudaFuncAttributeMaxDynamicSharedMemorySize, WIN_SMEM));
    configured = true;
  }
  GM_CUDA(gm_launch_pdl(k_sra_win_fwd, dim3(GM_NUM_SMS * 3, CHUNKS), dim3(256), (size_t)WIN_SMEM, st, (const __nv_bfloat16*)qkv, (int)n,
                        win_ptr, win_tok, tok_win, (__nv_bfloat16*)out, lse));
  return GEOMAE_OK;
}

int gm_sra_win_bwd(const void* qkv, const float* lse, const void* d_out, const float* dd, int64_t n, const int32_t* win_ptr,
                   const int32_t* win_tok, const int32_t* tok_win, void* d_qkv, cudaStream_t st) {
  static bool configured = false;
  if (!configured) {
    GM_CUDA(cudaFuncSetAttribute(k_sra_win_bwd, cudaFuncAttributeMaxDynamicSharedMemorySize, WIN_SMEM));
    configured = true;
  }
  GM_CUDA(gm_launch_pdl(k_sra_win_bwd, dim3(GM_NUM_SMS * 3, 2 * CHUNKS), dim3(256), (size_t)WIN_SMEM, st, (const __nv_bfloat16*)qkv, lse,
                        (const __nv_bfloat16*)d_out, dd, (int)n, win_ptr, win_tok, tok_win, (__nv_bfloat16*)d_qkv));
  return GEOMAE_OK;
}
