// Sparse Regional Attention core on tensor cores (bf16 operands, fp32 accumulate) — the `precision 1`
// path of SURVEY.md §8 row a18 (models/sst/sst_basic_block.py:26-61).
//
// Windows are short (mean 5-14 tokens, max 144) and head_dim is 16, so one window is far too small for
// an MMA tile.  Several windows are therefore packed into one tile with block-diagonal masking: a CTA
// owns TQ consecutive CSR positions (64 or 32 queries, 4 or 2 m16 tiles) for all 8 heads, one warp per
// head.  The keys those positions can see form ONE contiguous CSR range [lo, hi); its K|V rows are
// converted to bf16 while being staged into shared memory (chunks of KC rows), and every warp walks only
// the 16-key blocks that intersect the windows of the m-tile at hand:
//     S  = Q K^T      mma.sync.m16n8k16 (k = head_dim = 16: exactly one k-step)
//     P  = online softmax on the accumulator fragments (exp2 domain, window mask per element)
//     O += P V        the S accumulators re-packed as the A fragment, V through ldmatrix.trans
// No padding to 56/144 buckets, no key_padding_mask, no [W,T,T] attention map in memory.
// The backward is the same structure twice in one launch (blockIdx.y): as queries (dQ) and as keys
// (dK, dV), each output element written exactly once — no atomics.  D_i = dO_i.O_i is recomputed from
// the staged rows in both passes so the two passes have no ordering between them.
//
// tcgen05 is not used here on purpose: its minimum tile is M = 64/128 rows x one TMEM accumulator per
// head, and with K = 16 a single k-step per tile cannot amortise the descriptor / TMEM round trip; the
// warp-level MMA keeps S and P in registers.  The K = 128/256 projections around it are on tcgen05
// (sra_layer.cu).  fp32 parity mode (precision 3) keeps the SIMT kernel of sra_attention.cu.
#include <cuda_bf16.h>
#include <stdlib.h>

#include "common.cuh"

namespace {

constexpr int NH = 8;             // heads (one warp each)
constexpr int DM = 128;           // d_model = NH * 16
constexpr int RSB = 272;          // bytes per staged bf16 row: 128 bf16 + 16 B pad (conflict-free fragment loads)
constexpr int OS = 132;           // floats per row of the fp32 output staging tile
constexpr float LOG2E = 1.4426950408889634f;
constexpr float QSCALE = 0.25f * LOG2E;     // 1/sqrt(head_dim), exp2 domain

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// exp2 on the special-function unit: one MUFU.EX2 (exp2f without fast-math adds a denormal-range rescale per call)
__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  const __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<const uint32_t*>(&v);
}

// D[16x8] += A[16x16] * B[16x8], bf16 operands, fp32 accumulate
__device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// Four transposed 8x8 bf16 matrices: B fragments (k = row index of the staged tile, n = column) for two n-tiles.
// tile rows r0..r0+15, columns c0..c0+15 (bf16 elements) of a staged [rows][RSB] tile.
__device__ __forceinline__ void ldsm_bt(uint32_t (&r)[4], const uint8_t* tile, int r0, int c0) {
  const int lane = threadIdx.x & 31;
  const int mi = lane >> 3, rr = lane & 7;
  const uint32_t addr = smem_addr(tile + (r0 + (mi & 1) * 8 + rr) * RSB + (c0 + (mi >> 1) * 8) * 2);
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}

// A fragment of rows r0..r0+15, columns c0..c0+15 of a staged tile (row-major, K-contiguous)
__device__ __forceinline__ void lda(uint32_t (&a)[4], const uint8_t* tile, int r0, int c0) {
  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const uint8_t* p = tile + (r0 + g) * RSB + (c0 + 2 * t) * 2;
  a[0] = *reinterpret_cast<const uint32_t*>(p);
  a[1] = *reinterpret_cast<const uint32_t*>(p + 8 * RSB);
  a[2] = *reinterpret_cast<const uint32_t*>(p + 16);
  a[3] = *reinterpret_cast<const uint32_t*>(p + 8 * RSB + 16);
}

// B fragment (k = the 16 columns c0.., n = rows r0..r0+7) of a staged row-major tile: B[k][n] = tile[r0+n][c0+k]
__device__ __forceinline__ void ldb(uint32_t& b0, uint32_t& b1, const uint8_t* tile, int r0, int c0) {
  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const uint8_t* p = tile + (r0 + g) * RSB + (c0 + 2 * t) * 2;
  b0 = *reinterpret_cast<const uint32_t*>(p);
  b1 = *reinterpret_cast<const uint32_t*>(p + 16);
}

// Stage `rows` rows (tokens win_tok[p_start + r]) of 128 floats src[tok*ld + col ..] as bf16 (x scale) into a
// [rows_pad][RSB] tile; rows in [rows, rows_pad) are zero-filled.  One warp moves a whole 512-byte row per load
// instruction, four rows in flight per thread.
template <int NSRC>
__device__ __forceinline__ void stage_rows_bf16(uint8_t* dst0, const float* __restrict__ src0, int ld0, int col0,
                                                float scale0, uint8_t* dst1, const float* __restrict__ src1, int ld1,
                                                int col1, const int32_t* __restrict__ win_tok, int p_start, int rows,
                                                int rows_pad) {
  const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
#pragma unroll 1
  for (int r0 = wrp * 4; r0 < rows_pad; r0 += 32) {
    float4 v[4], w[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int r = r0 + k;
      v[k] = make_float4(0.f, 0.f, 0.f, 0.f);
      w[k] = v[k];
      if (r < rows) {
        const int64_t tok = __ldg(win_tok + p_start + r);
        v[k] = __ldg(reinterpret_cast<const float4*>(src0 + tok * ld0 + col0) + lane);
        if (NSRC == 2) w[k] = __ldg(reinterpret_cast<const float4*>(src1 + tok * ld1 + col1) + lane);
      }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int r = r0 + k;
      if (r < rows_pad) {
        uint2 o;
        o.x = pack_bf16(v[k].x * scale0, v[k].y * scale0);
        o.y = pack_bf16(v[k].z * scale0, v[k].w * scale0);
        *reinterpret_cast<uint2*>(dst0 + r * RSB + lane * 8) = o;
        if (NSRC == 2) {
          o.x = pack_bf16(w[k].x, w[k].y);
          o.y = pack_bf16(w[k].z, w[k].w);
          *reinterpret_cast<uint2*>(dst1 + r * RSB + lane * 8) = o;
        }
      }
    }
  }
}

// The same staging from bf16 rows in global memory (rows of 128 bf16 = 256 B at src[tok*ld + col ..]): pure 16-byte
// cp.async copies, no registers, no conversion; rows in [rows, rows_pad) are zero-filled.  16 consecutive threads move
// one 256-byte row.  Completion: cp_async_wait_all() + __syncthreads() by the caller.
__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(smem_addr(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;\n" ::: "memory");
}
template <int NSRC>
__device__ __forceinline__ void stage_rows_cp(uint8_t* dst0, const __nv_bfloat16* __restrict__ src0, int ld0, int col0,
                                              uint8_t* dst1, const __nv_bfloat16* __restrict__ src1, int ld1, int col1,
                                              const int32_t* __restrict__ win_tok, int p_start, int rows, int rows_pad) {
  constexpr int PER_ROW = 16 * NSRC;
  for (int idx = threadIdx.x; idx < rows_pad * PER_ROW; idx += 256) {
    const int r = idx / PER_ROW, c = idx % PER_ROW;
    const int which = c >> 4, ch = c & 15;
    uint8_t* dst = (which ? dst1 : dst0) + r * RSB + ch * 16;
    if (r < rows) {
      const int64_t tok = __ldg(win_tok + p_start + r);
      const __nv_bfloat16* src = which ? src1 + tok * ld1 + col1 : src0 + tok * ld0 + col0;
      cp_async16(dst, src + ch * 8);
    } else {
      *reinterpret_cast<uint4*>(dst) = make_uint4(0u, 0u, 0u, 0u);
    }
  }
}

// per-(row, head) D = dO . O (fp32, from global) and log2-domain LSE for `rows` CSR positions starting at p_start
__device__ __forceinline__ void stage_lse_d(float* sL, float* sD, const float* __restrict__ lse,
                                            const float* __restrict__ d_out, const float* __restrict__ out,
                                            const float* __restrict__ dd, const int32_t* __restrict__ win_tok,
                                            int p_start, int rows) {
  for (int idx = threadIdx.x; idx < rows * NH; idx += 256) {
    const int r = idx >> 3, hh = idx & 7;
    const int64_t tok = __ldg(win_tok + p_start + r);
    if (dd) {                             // D was produced by the GEMM that wrote d_out (geomae_linear_args.dot_out)
      sD[idx] = __ldg(dd + tok * NH + hh);
      sL[idx] = __ldg(lse + tok * NH + hh) * LOG2E;
      continue;
    }
    const float4* g4 = reinterpret_cast<const float4*>(d_out + tok * DM + hh * 16);
    const float4* o4 = reinterpret_cast<const float4*>(out + tok * DM + hh * 16);
    float d = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float4 a = __ldg(g4 + i), b = __ldg(o4 + i);
      d = fmaf(a.x, b.x, d); d = fmaf(a.y, b.y, d); d = fmaf(a.z, b.z, d); d = fmaf(a.w, b.w, d);
    }
    sD[idx] = d;
    sL[idx] = __ldg(lse + tok * NH + hh) * LOG2E;
  }
}

struct RowMeta {
  int tok[64], beg[64], end[64];
};

// window [beg, end) (CSR positions) and token id of the CTA's TQ positions; rows past n get an empty window
template <int TQ>
__device__ __forceinline__ void locate_rows(RowMeta& m, int p0, int n, const int32_t* __restrict__ win_ptr,
                                            const int32_t* __restrict__ win_tok, const int32_t* __restrict__ tok_win) {
  if (threadIdx.x < TQ) {
    const int p = p0 + threadIdx.x;
    int tok = 0, b = 0, e = 0;
    if (p < n) {
      tok = __ldg(win_tok + p);
      const int w = __ldg(tok_win + tok);
      b = __ldg(win_ptr + w);
      e = __ldg(win_ptr + w + 1);
    }
    m.tok[threadIdx.x] = tok; m.beg[threadIdx.x] = b; m.end[threadIdx.x] = e;
  }
}

// write a [rows][128] fp32 tile staged in shared memory (row stride OS) to dst[tok * ld + col ..], 512 B per row
__device__ __forceinline__ void flush_tile(const float* sO, const RowMeta& m, int rows, float* dst, int ld, int col,
                                           bool as_bf16 = false) {
  const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
  for (int r = wrp; r < rows; r += 8) {
    const float4 v = *reinterpret_cast<const float4*>(sO + r * OS + lane * 4);
    if (as_bf16) {
      __nv_bfloat16* d16 = reinterpret_cast<__nv_bfloat16*>(dst) + (int64_t)m.tok[r] * ld + col;
      reinterpret_cast<uint2*>(d16)[lane] = make_uint2(pack_bf16(v.x, v.y), pack_bf16(v.z, v.w));
    } else {
      reinterpret_cast<float4*>(dst + (int64_t)m.tok[r] * ld + col)[lane] = v;
    }
  }
}

// accumulator fragments of one head (two n-tiles of 8 dims) -> staging tile rows r0+g, r0+g+8
__device__ __forceinline__ void put_frag(float* sO, int r0, int h, const float (&o0)[4], const float (&o1)[4], float s0,
                                         float s1) {
  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  float* a = sO + (r0 + g) * OS + h * 16 + 2 * t;
  float* b = a + 8 * OS;
  *reinterpret_cast<float2*>(a) = make_float2(o0[0] * s0, o0[1] * s0);
  *reinterpret_cast<float2*>(a + 8) = make_float2(o1[0] * s0, o1[1] * s0);
  *reinterpret_cast<float2*>(b) = make_float2(o0[2] * s1, o0[3] * s1);
  *reinterpret_cast<float2*>(b + 8) = make_float2(o1[2] * s1, o1[3] * s1);
}

template <int TQ, int KC, int MINB = 2>
__global__ void __launch_bounds__(256, MINB) k_sra_tc_fwd(const float* __restrict__ qkv, int n,
                                                       const int32_t* __restrict__ win_ptr,
                                                       const int32_t* __restrict__ win_tok,
                                                       const int32_t* __restrict__ tok_win, float* out, float* lse,
                                                       int io_flags) {
  extern __shared__ __align__(16) uint8_t smem[];
  __shared__ RowMeta meta;
  __shared__ float s_lse[TQ * NH];
  const bool qkv16 = io_flags & 1;
  const __nv_bfloat16* qkv_h = reinterpret_cast<const __nv_bfloat16*>(qkv);
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + TQ * RSB;
  uint8_t* sV = sK + KC * RSB;
  constexpr int MT = TQ / 16;
  const int lane = threadIdx.x & 31, h = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
  const int p0 = blockIdx.x * TQ;
  const int nq = min(TQ, n - p0);
  gm_pdl_wait();
  gm_pdl_trigger();
  locate_rows<TQ>(meta, p0, n, win_ptr, win_tok, tok_win);
  if (qkv16) stage_rows_cp<1>(sQ, qkv_h, 3 * DM, 0, nullptr, nullptr, 0, 0, win_tok, p0, nq, TQ);
  else stage_rows_bf16<1>(sQ, qkv, 3 * DM, 0, 1.0f, nullptr, nullptr, 0, 0, win_tok, p0, nq, TQ);
  cp_async_wait_all();
  __syncthreads();
  const int lo = meta.beg[0], hi = meta.end[nq - 1];
  float mx[MT][2], ls[MT][2], o0[MT][4], o1[MT][4];
#pragma unroll
  for (int mt = 0; mt < MT; ++mt) {
    mx[mt][0] = mx[mt][1] = -INFINITY;
    ls[mt][0] = ls[mt][1] = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) { o0[mt][i] = 0.f; o1[mt][i] = 0.f; }
  }
  for (int c0 = lo; c0 < hi; c0 += KC) {
    const int rows = min(KC, hi - c0);
    const int rows_pad = (rows + 15) & ~15;
    if (c0 > lo) __syncthreads();          // previous chunk fully consumed
    if (qkv16) stage_rows_cp<2>(sK, qkv_h, 3 * DM, DM, sV, qkv_h, 3 * DM, 2 * DM, win_tok, c0, rows, rows_pad);
    else stage_rows_bf16<2>(sK, qkv, 3 * DM, DM, 1.0f, sV, qkv, 3 * DM, 2 * DM, win_tok, c0, rows, rows_pad);
    cp_async_wait_all();
    __syncthreads();
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
      const int r0 = mt * 16;
      if (r0 < nq) {
        const int rl = min(r0 + 15, nq - 1);
        const int j_lo = max(meta.beg[r0], c0) - c0, j_hi = min(meta.end[rl], c0 + rows) - c0;
        if (j_lo < j_hi) {
          uint32_t qa[4];
          lda(qa, sQ, r0, h * 16);
          const int b0 = meta.beg[r0 + g] - c0, e0 = meta.end[r0 + g] - c0;
          const int b1 = meta.beg[r0 + g + 8] - c0, e1 = meta.end[r0 + g + 8] - c0;
          // all 16 rows of the m-tile valid and in ONE window [wb, we) (chunk-relative): interior key blocks need no mask
          const bool one_window = r0 + 15 < nq && meta.beg[r0] == meta.beg[r0 + 15];
          const int wb = meta.beg[r0] - c0, we = meta.end[r0] - c0;
          for (int kb = j_lo >> 4; kb * 16 < j_hi; ++kb) {
            float s0[4] = {0.f, 0.f, 0.f, 0.f}, s1[4] = {0.f, 0.f, 0.f, 0.f};
            uint32_t kb0, kb1;
            ldb(kb0, kb1, sK, kb * 16, h * 16);
            mma16816(s0, qa, kb0, kb1);
            ldb(kb0, kb1, sK, kb * 16 + 8, h * 16);
            mma16816(s1, qa, kb0, kb1);
            const int k0 = kb * 16 + 2 * t, k1 = k0 + 8;       // scores -> exp2 domain (scale 1/sqrt(hd) * log2 e), window mask
            if (one_window && kb * 16 >= wb && kb * 16 + 16 <= we) {   // whole block inside the tile's single window: no mask
#pragma unroll
              for (int i = 0; i < 4; ++i) { s0[i] *= QSCALE; s1[i] *= QSCALE; }
            } else {
              s0[0] = (k0 >= b0 && k0 < e0) ? s0[0] * QSCALE : -INFINITY;
              s0[1] = (k0 + 1 >= b0 && k0 + 1 < e0) ? s0[1] * QSCALE : -INFINITY;
              s1[0] = (k1 >= b0 && k1 < e0) ? s1[0] * QSCALE : -INFINITY;
              s1[1] = (k1 + 1 >= b0 && k1 + 1 < e0) ? s1[1] * QSCALE : -INFINITY;
              s0[2] = (k0 >= b1 && k0 < e1) ? s0[2] * QSCALE : -INFINITY;
              s0[3] = (k0 + 1 >= b1 && k0 + 1 < e1) ? s0[3] * QSCALE : -INFINITY;
              s1[2] = (k1 >= b1 && k1 < e1) ? s1[2] * QSCALE : -INFINITY;
              s1[3] = (k1 + 1 >= b1 && k1 + 1 < e1) ? s1[3] * QSCALE : -INFINITY;
            }
            float m0 = fmaxf(fmaxf(s0[0], s0[1]), fmaxf(s1[0], s1[1]));
            float m1 = fmaxf(fmaxf(s0[2], s0[3]), fmaxf(s1[2], s1[3]));
            m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 1));
            m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 2));
            m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 1));
            m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 2));
            const float n0 = fmaxf(mx[mt][0], m0), n1 = fmaxf(mx[mt][1], m1);
            const float u0 = n0 == -INFINITY ? 0.f : n0, u1 = n1 == -INFINITY ? 0.f : n1;
            const float c0f = ex2(mx[mt][0] - u0), c1f = ex2(mx[mt][1] - u1);
            mx[mt][0] = n0; mx[mt][1] = n1;
            const float p00 = ex2(s0[0] - u0), p01 = ex2(s0[1] - u0), p02 = ex2(s1[0] - u0), p03 = ex2(s1[1] - u0);
            const float p10 = ex2(s0[2] - u1), p11 = ex2(s0[3] - u1), p12 = ex2(s1[2] - u1), p13 = ex2(s1[3] - u1);
            ls[mt][0] = ls[mt][0] * c0f + ((p00 + p01) + (p02 + p03));
            ls[mt][1] = ls[mt][1] * c1f + ((p10 + p11) + (p12 + p13));
            o0[mt][0] *= c0f; o0[mt][1] *= c0f; o1[mt][0] *= c0f; o1[mt][1] *= c0f;
            o0[mt][2] *= c1f; o0[mt][3] *= c1f; o1[mt][2] *= c1f; o1[mt][3] *= c1f;
            const uint32_t pa[4] = {pack_bf16(p00, p01), pack_bf16(p10, p11), pack_bf16(p02, p03), pack_bf16(p12, p13)};
            uint32_t vb[4];
            ldsm_bt(vb, sV, kb * 16, h * 16);
            mma16816(o0[mt], pa, vb[0], vb[1]);
            mma16816(o1[mt], pa, vb[2], vb[3]);
          }
        }
      }
    }
  }
  __syncthreads();                         // K/V tiles free: reuse them as the fp32 output staging tile
  float* sO = reinterpret_cast<float*>(sK);
#pragma unroll
  for (int mt = 0; mt < MT; ++mt) {
    float l0 = ls[mt][0], l1 = ls[mt][1];
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    const float i0 = l0 > 0.f ? 1.0f / l0 : 0.f, i1 = l1 > 0.f ? 1.0f / l1 : 0.f;
    put_frag(sO, mt * 16, h, o0[mt], o1[mt], i0, i1);
    if (t == 0) {
      s_lse[(mt * 16 + g) * NH + h] = (mx[mt][0] + log2f(l0)) * 0.6931471805599453f;
      s_lse[(mt * 16 + g + 8) * NH + h] = (mx[mt][1] + log2f(l1)) * 0.6931471805599453f;
    }
  }
  __syncthreads();
  flush_tile(sO, meta, nq, out, DM, 0, (io_flags & 8) != 0);
  for (int idx = threadIdx.x; idx < nq * NH; idx += 256) lse[(int64_t)meta.tok[idx >> 3] * NH + (idx & 7)] = s_lse[idx];
}

// Backward.  AS_KEYS == false: the CTA's TQ positions are QUERIES -> dQ.   AS_KEYS == true: they are KEYS -> dK, dV.
template <int TQ, int KC, bool AS_KEYS>
__device__ __forceinline__ void sra_bwd_body(uint8_t* smem, RowMeta& meta, const float* __restrict__ qkv,
                                             const float* __restrict__ out, const float* __restrict__ lse,
                                             const float* __restrict__ d_out, int n, const int32_t* __restrict__ win_ptr,
                                             const int32_t* __restrict__ win_tok, const int32_t* __restrict__ tok_win,
                                             float* d_qkv, const float* __restrict__ dd, int io_flags) {
  constexpr int MT = TQ / 16;
  const bool qkv16 = io_flags & 1, dout16 = io_flags & 2, dqkv16 = io_flags & 4;
  const __nv_bfloat16* qkv_h = reinterpret_cast<const __nv_bfloat16*>(qkv);
  const __nv_bfloat16* dout_h = reinterpret_cast<const __nv_bfloat16*>(d_out);
  uint8_t* sA = smem;                       // own rows, operand 1: Q (queries) | K (keys)
  uint8_t* sB = sA + TQ * RSB;              // own rows, operand 2: dO       (queries) | V       (keys)
  uint8_t* sC = sB + TQ * RSB;              // chunk rows, operand 1: K      (queries) | Q       (keys)
  uint8_t* sD = sC + KC * RSB;              // chunk rows, operand 2: V      (queries) | dO      (keys)
  float* sL = reinterpret_cast<float*>(sD + KC * RSB);   // [KC][8] log2-domain LSE of the query rows
  float* sDd = sL + KC * NH;                             // [KC][8] D = dO.O of the query rows
  static_assert(KC >= TQ, "chunk must be at least one row tile");
  const int lane = threadIdx.x & 31, h = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
  const int p0 = blockIdx.x * TQ;
  const int nq = min(TQ, n - p0);
  gm_pdl_wait();
  gm_pdl_trigger();
  locate_rows<TQ>(meta, p0, n, win_ptr, win_tok, tok_win);
  // rows of 128 channels from qkv (column block cb) or d_out, fp32 (converted while staged) or bf16 (cp.async)
  auto stage_q = [&](uint8_t* dst, int cb, int p_start, int rows, int rows_pad) {
    if (qkv16) stage_rows_cp<1>(dst, qkv_h, 3 * DM, cb * DM, nullptr, nullptr, 0, 0, win_tok, p_start, rows, rows_pad);
    else stage_rows_bf16<1>(dst, qkv, 3 * DM, cb * DM, 1.0f, nullptr, nullptr, 0, 0, win_tok, p_start, rows, rows_pad);
  };
  auto stage_g = [&](uint8_t* dst, int p_start, int rows, int rows_pad) {
    if (dout16) stage_rows_cp<1>(dst, dout_h, DM, 0, nullptr, nullptr, 0, 0, win_tok, p_start, rows, rows_pad);
    else stage_rows_bf16<1>(dst, d_out, DM, 0, 1.0f, nullptr, nullptr, 0, 0, win_tok, p_start, rows, rows_pad);
  };
  if (!AS_KEYS) {
    stage_q(sA, 0, p0, nq, TQ);
    stage_g(sB, p0, nq, TQ);
    stage_lse_d(sL, sDd, lse, d_out, out, dd, win_tok, p0, nq);
  } else {
    stage_q(sA, 1, p0, nq, TQ);
    stage_q(sB, 2, p0, nq, TQ);
  }
  cp_async_wait_all();
  __syncthreads();
  const int lo = meta.beg[0], hi = meta.end[nq - 1];
  float a0[MT][4], a1[MT][4];                              // dQ | dK
  float v0[AS_KEYS ? MT : 1][4], v1[AS_KEYS ? MT : 1][4];  // dV (keys pass only)
#pragma unroll
  for (int mt = 0; mt < MT; ++mt)
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      a0[mt][i] = 0.f; a1[mt][i] = 0.f;
      if (AS_KEYS) { v0[mt][i] = 0.f; v1[mt][i] = 0.f; }
    }
  for (int c0 = lo; c0 < hi; c0 += KC) {
    const int rows = min(KC, hi - c0);
    const int rows_pad = (rows + 15) & ~15;
    if (c0 > lo || !AS_KEYS) __syncthreads();     // previous chunk consumed (queries pass: own-row sL/sDd published)
    if (!AS_KEYS) {
      stage_q(sC, 1, c0, rows, rows_pad);
      stage_q(sD, 2, c0, rows, rows_pad);
    } else {
      stage_q(sC, 0, c0, rows, rows_pad);
      stage_g(sD, c0, rows, rows_pad);
      stage_lse_d(sL, sDd, lse, d_out, out, dd, win_tok, c0, rows);
    }
    cp_async_wait_all();
    __syncthreads();
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
      const int r0 = mt * 16;
      if (r0 < nq) {
        const int rl = min(r0 + 15, nq - 1);
        const int j_lo = max(meta.beg[r0], c0) - c0, j_hi = min(meta.end[rl], c0 + rows) - c0;
        if (j_lo < j_hi) {
          uint32_t fa[4], fb[4];
          lda(fa, sA, r0, h * 16);
          lda(fb, sB, r0, h * 16);
          const int b0 = meta.beg[r0 + g] - c0, e0 = meta.end[r0 + g] - c0;
          const int b1 = meta.beg[r0 + g + 8] - c0, e1 = meta.end[r0 + g + 8] - c0;
          const bool one_window = r0 + 15 < nq && meta.beg[r0] == meta.beg[r0 + 15];
          const int wb = meta.beg[r0] - c0, we = meta.end[r0] - c0;
          float lr0 = 0.f, lr1 = 0.f, dr0 = 0.f, dr1 = 0.f;
          if (!AS_KEYS) {                  // LSE and D belong to the ROWS (queries) of this m-tile
            lr0 = sL[(r0 + g) * NH + h]; lr1 = sL[(r0 + g + 8) * NH + h];
            dr0 = sDd[(r0 + g) * NH + h]; dr1 = sDd[(r0 + g + 8) * NH + h];
          }
          for (int kb = j_lo >> 4; kb * 16 < j_hi; ++kb) {
            float s0[4] = {0.f, 0.f, 0.f, 0.f}, s1[4] = {0.f, 0.f, 0.f, 0.f};
            float q0[4] = {0.f, 0.f, 0.f, 0.f}, q1[4] = {0.f, 0.f, 0.f, 0.f};
            uint32_t x0, x1;
            ldb(x0, x1, sC, kb * 16, h * 16);
            mma16816(s0, fa, x0, x1);                     // scores (log2 domain)
            ldb(x0, x1, sC, kb * 16 + 8, h * 16);
            mma16816(s1, fa, x0, x1);
            ldb(x0, x1, sD, kb * 16, h * 16);
            mma16816(q0, fb, x0, x1);                     // dP
            ldb(x0, x1, sD, kb * 16 + 8, h * 16);
            mma16816(q1, fb, x0, x1);
            const int k0 = kb * 16 + 2 * t, k1 = k0 + 8;  // this thread's columns: k0, k0+1 | k1, k1+1
            float lc[4], dc[4];                           // LSE / D per column (keys pass) or per row (queries pass)
            if (AS_KEYS) {
              lc[0] = sL[k0 * NH + h]; lc[1] = sL[(k0 + 1) * NH + h]; lc[2] = sL[k1 * NH + h]; lc[3] = sL[(k1 + 1) * NH + h];
              dc[0] = sDd[k0 * NH + h]; dc[1] = sDd[(k0 + 1) * NH + h]; dc[2] = sDd[k1 * NH + h]; dc[3] = sDd[(k1 + 1) * NH + h];
            }
            const int kk[4] = {k0, k0 + 1, k1, k1 + 1};
            const bool inside = one_window && kb * 16 >= wb && kb * 16 + 16 <= we;    // warp-uniform: no per-element mask
            const float sc[2][4] = {{s0[0], s0[1], s1[0], s1[1]}, {s0[2], s0[3], s1[2], s1[3]}};
            const float dp[2][4] = {{q0[0], q0[1], q1[0], q1[1]}, {q0[2], q0[3], q1[2], q1[3]}};
            float pr[2][4], dsv[2][4];
#pragma unroll
            for (int r = 0; r < 2; ++r) {
              const int bb = r ? b1 : b0, ee = r ? e1 : e0;
#pragma unroll
              for (int c = 0; c < 4; ++c) {
                const bool valid = inside || (kk[c] >= bb && kk[c] < ee);
                const float lz = AS_KEYS ? lc[c] : (r ? lr1 : lr0);
                const float dz = AS_KEYS ? dc[c] : (r ? dr1 : dr0);
                const float p = valid ? ex2(fmaf(sc[r][c], QSCALE, -lz)) : 0.f;
                pr[r][c] = p;
                dsv[r][c] = valid ? p * (dp[r][c] - dz) : 0.f;   // dS = P * (dP - D)
              }
            }
            const uint32_t ds[4] = {pack_bf16(dsv[0][0], dsv[0][1]), pack_bf16(dsv[1][0], dsv[1][1]),
                                    pack_bf16(dsv[0][2], dsv[0][3]), pack_bf16(dsv[1][2], dsv[1][3])};
            uint32_t bt[4];
            ldsm_bt(bt, sC, kb * 16, h * 16);             // queries pass: K (dQ += dS K);  keys pass: Q (dK += dS^T Q)
            mma16816(a0[mt], ds, bt[0], bt[1]);
            mma16816(a1[mt], ds, bt[2], bt[3]);
            if (AS_KEYS) {
              const uint32_t pp[4] = {pack_bf16(pr[0][0], pr[0][1]), pack_bf16(pr[1][0], pr[1][1]),
                                      pack_bf16(pr[0][2], pr[0][3]), pack_bf16(pr[1][2], pr[1][3])};
              ldsm_bt(bt, sD, kb * 16, h * 16);           // dO (dV += P^T dO)
              mma16816(v0[mt], pp, bt[0], bt[1]);
              mma16816(v1[mt], pp, bt[2], bt[3]);
            }
          }
        }
      }
    }
  }
  __syncthreads();
  float* sO = reinterpret_cast<float*>(sC);     // chunk tiles are free: fp32 staging for coalesced row stores
#pragma unroll
  for (int mt = 0; mt < MT; ++mt) put_frag(sO, mt * 16, h, a0[mt], a1[mt], 0.25f, 0.25f);
  __syncthreads();
  flush_tile(sO, meta, nq, d_qkv, 3 * DM, AS_KEYS ? DM : 0, dqkv16);
  if (AS_KEYS) {
    __syncthreads();
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) put_frag(sO, mt * 16, h, v0[mt], v1[mt], 1.0f, 1.0f);
    __syncthreads();
    flush_tile(sO, meta, nq, d_qkv, 3 * DM, 2 * DM, dqkv16);
  }
}

template <int TQ, int KC>
__global__ void __launch_bounds__(256, 2) k_sra_tc_bwd(const float* __restrict__ qkv, const float* __restrict__ out,
                                                       const float* __restrict__ lse, const float* __restrict__ d_out,
                                                       int n, const int32_t* __restrict__ win_ptr,
                                                       const int32_t* __restrict__ win_tok,
                                                       const int32_t* __restrict__ tok_win, float* d_qkv,
                                                       const float* __restrict__ dd, int io_flags) {
  extern __shared__ __align__(16) uint8_t smem[];
  __shared__ RowMeta meta;
  if (blockIdx.y == 0)
    sra_bwd_body<TQ, KC, false>(smem, meta, qkv, out, lse, d_out, n, win_ptr, win_tok, tok_win, d_qkv, dd, io_flags);
  else
    sra_bwd_body<TQ, KC, true>(smem, meta, qkv, out, lse, d_out, n, win_ptr, win_tok, tok_win, d_qkv, dd, io_flags);
}

constexpr int KC_FWD = 128, KC_BWD = 112;
template <int TQ> constexpr int smem_fwd() { return (TQ + 2 * KC_FWD) * RSB; }
template <int TQ> constexpr int smem_bwd() { return (2 * TQ + 2 * KC_BWD) * RSB + 2 * KC_BWD * NH * 4; }
static_assert(64 * OS * 4 <= 2 * KC_FWD * RSB && 64 * OS * 4 <= 2 * KC_BWD * RSB, "output staging tile must fit");

// 32-query tiles, 112-key chunks, three CTAs per SM (<= 85 registers, 70 KB of shared memory)
int launch_fwd_occ3(const float* qkv, int n, const int32_t* win_ptr, const int32_t* win_tok, const int32_t* tok_win,
                    float* out, float* lse, int io_flags, cudaStream_t st) {
  constexpr int TQ = 32, KC = 112, SM = (TQ + 2 * KC) * RSB;
  static bool configured = false;
  if (!configured) {
    GM_CUDA(cudaFuncSetAttribute(k_sra_tc_fwd<TQ, KC, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM));
    configured = true;
  }
  GM_CUDA(gm_launch_pdl(k_sra_tc_fwd<TQ, KC, 3>, dim3(gm_div_up(n, TQ)), dim3(256), (size_t)SM, st, qkv, n, win_ptr,
                        win_tok, tok_win, out, lse, io_flags));
  return GEOMAE_OK;
}

template <int TQ>
int launch_fwd(const float* qkv, int n, const int32_t* win_ptr, const int32_t* win_tok, const int32_t* tok_win, float* out,
               float* lse, int io_flags, cudaStream_t st) {
  static bool configured = false;
  if (!configured) {
    GM_CUDA(cudaFuncSetAttribute(k_sra_tc_fwd<TQ, KC_FWD>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_fwd<TQ>()));
    configured = true;
  }
  GM_CUDA(gm_launch_pdl(k_sra_tc_fwd<TQ, KC_FWD>, dim3(gm_div_up(n, TQ)), dim3(256), (size_t)smem_fwd<TQ>(), st, qkv, n,
                        win_ptr, win_tok, tok_win, out, lse, io_flags));
  return GEOMAE_OK;
}

template <int TQ>
int launch_bwd(const float* qkv, const float* out, const float* lse, const float* d_out, int n, const int32_t* win_ptr,
               const int32_t* win_tok, const int32_t* tok_win, float* d_qkv, const float* dd, int io_flags,
               cudaStream_t st) {
  static bool configured = false;
  if (!configured) {
    GM_CUDA(cudaFuncSetAttribute(k_sra_tc_bwd<TQ, KC_BWD>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bwd<TQ>()));
    configured = true;
  }
  GM_CUDA(gm_launch_pdl(k_sra_tc_bwd<TQ, KC_BWD>, dim3(gm_div_up(n, TQ), 2), dim3(256), (size_t)smem_bwd<TQ>(), st, qkv, out,
                        lse, d_out, n, win_ptr, win_tok, tok_win, d_qkv, dd, io_flags));
  return GEOMAE_OK;
}

}  // namespace

extern "C" int geomae_sra_attention_tc_fwd(const float* qkv, int64_t n_tokens, int32_t n_heads, const int32_t* win_ptr,
                                           const int32_t* win_tok, const int32_t* tok_win, float* out, float* lse,
                                           int32_t io_flags, void* stream) {
  GM_REQUIRE(n_heads == NH, "sra_attention_tc: built for %d heads of 16 channels (got %d)", NH, n_heads);
  if (n_tokens == 0) return GEOMAE_OK;
  GM_REQUIRE(qkv && win_ptr && win_tok && tok_win && out && lse, "sra_attention_tc_fwd: null argument");
  GM_REQUIRE(n_tokens < (int64_t)1 << 31, "sra_attention_tc: too many tokens");
  const int n = (int)n_tokens;
  static const int force_tq = getenv("GEOMAE_ATTN_TQ") ? atoi(getenv("GEOMAE_ATTN_TQ")) : 0;   // tuning switch
  // small token sets (the encoder sees 30 % of the pillars): 32-query CTAs so the grid still covers the SMs (one
  // wave: bound by the per-CTA latency chain).  Large sets are throughput-bound and gain from a third resident CTA
  // per SM (32 queries, 112-key chunks, <= 85 registers): 38.9 -> 34 us on 24.7 k decoder tokens
  // (tools/bench_attention.py); GEOMAE_ATTN_TQ = 32 | 64 forces the older tilings.
  if (force_tq == 32 || (force_tq != 64 && gm_div_up(n, 64) < 2 * GM_NUM_SMS))
    return launch_fwd<32>(qkv, n, win_ptr, win_tok, tok_win, out, lse, io_flags, (cudaStream_t)stream);
  if (force_tq == 64)
    return launch_fwd<64>(qkv, n, win_ptr, win_tok, tok_win, out, lse, io_flags, (cudaStream_t)stream);
  return launch_fwd_occ3(qkv, n, win_ptr, win_tok, tok_win, out, lse, io_flags, (cudaStream_t)stream);
}

extern "C" int geomae_sra_attention_tc_bwd(const float* qkv, const float* out, const float* lse, const float* d_out,
                                           int64_t n_tokens, int32_t n_heads, const int32_t* win_ptr,
                                           const int32_t* win_tok, const int32_t* tok_win, float* d_qkv, const float* dd,
                                           int32_t io_flags, void* stream) {
  GM_REQUIRE(n_heads == NH, "sra_attention_tc: built for %d heads of 16 channels (got %d)", NH, n_heads);
  if (n_tokens == 0) return GEOMAE_OK;
  GM_REQUIRE(qkv && out && lse && d_out && win_ptr && win_tok && tok_win && d_qkv, "sra_attention_tc_bwd: null argument");
  GM_REQUIRE(!(io_flags & 2) || dd, "sra_attention_tc_bwd: a bf16 d_out needs the precomputed D = dO.O (dd)");
  GM_REQUIRE(n_tokens < (int64_t)1 << 31, "sra_attention_tc: too many tokens");
  const int n = (int)n_tokens;
  static const int force_tq = getenv("GEOMAE_ATTN_TQ") ? atoi(getenv("GEOMAE_ATTN_TQ")) : 0;   // tuning switch
  if (force_tq == 32 || (force_tq != 64 && gm_div_up(n, 64) < GM_NUM_SMS))
    return launch_bwd<32>(qkv, out, lse, d_out, n, win_ptr, win_tok, tok_win, d_qkv, dd, io_flags, (cudaStream_t)stream);
  return launch_bwd<64>(qkv, out, lse, d_out, n, win_ptr, win_tok, tok_win, d_qkv, dd, io_flags, (cudaStream_t)stream);
}
