// Sparse Regional Attention core over CSR windows, fp32.
//
// The reference scatters tokens into zero-padded [W, 56|144, C] buckets, runs nn.MultiheadAttention
// with a key_padding_mask (11.5x padding in the encoder) and gathers the result back.  Windows here
// are short (mean 5 tokens in the encoder, 14 in the decoders, max 144) and head_dim is 16, so the
// work per window is far too small for a CTA: the kernel is latency-bound, not FLOP-bound (a whole
// decoder launch is ~0.3 GFLOP).  Mapping: ONE THREAD = one (query token, head); a CTA = 32
// consecutive CSR positions x 8 heads (warp = head).  The windows those positions belong to form one
// contiguous CSR range; its K|V rows (1 KB contiguous per token) are staged chunk-wise into shared
// memory by all 256 threads with coalesced 128-bit loads, and every thread then walks only the keys
// of its own window with broadcast shared-memory reads.  No padding, no mask, no [W,T,T] attention
// map, online softmax in the exp2 domain.  The backward is two such passes (as query: dQ; as key:
// dK, dV), each output element written exactly once — no atomics.
// head_dim = 16: one score costs 16 FMAs against one exp — SFU/issue bound, not a tensor-core shape;
// the K=128 projections around it are where tcgen05 is used (sra_layer.cu).
#include "common.cuh"

namespace {

constexpr int HD = 16;      // head_dim
constexpr int MAX_H = 8;    // heads per CTA (one warp each)

__device__ __forceinline__ void load_row16(const float* __restrict__ p, float* r) {
  const float4* p4 = reinterpret_cast<const float4*>(p);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float4 t = __ldg(p4 + i);
    r[4 * i + 0] = t.x; r[4 * i + 1] = t.y; r[4 * i + 2] = t.z; r[4 * i + 3] = t.w;
  }
}
__device__ __forceinline__ void store_row16(float* p, const float* r) {
  float4* p4 = reinterpret_cast<float4*>(p);
#pragma unroll
  for (int i = 0; i < 4; ++i) p4[i] = make_float4(r[4 * i], r[4 * i + 1], r[4 * i + 2], r[4 * i + 3]);
}
__device__ __forceinline__ float dot16(const float* a, const float* b) {
  float s = 0.f;
#pragma unroll
  for (int d = 0; d < HD; ++d) s = fmaf(a[d], b[d], s);
  return s;
}

struct Pos {
  bool valid;
  int tok, beg, len;
};

__device__ __forceinline__ Pos locate(int64_t n, const int32_t* __restrict__ win_ptr, const int32_t* __restrict__ win_tok,
                                      const int32_t* __restrict__ tok_win) {
  Pos p;
  const int64_t pos = (int64_t)blockIdx.x * 32 + (threadIdx.x & 31);
  p.valid = pos < n;
  p.tok = p.valid ? __ldg(win_tok + pos) : 0;
  const int w = p.valid ? __ldg(tok_win + p.tok) : 0;
  p.beg = p.valid ? __ldg(win_ptr + w) : 0;
  p.len = p.valid ? __ldg(win_ptr + w + 1) - p.beg : 0;
  return p;
}

constexpr int CH = 64;          // CSR positions staged per chunk
constexpr int RS = 132;         // shared-memory row stride in floats (128 + 4: neighbouring rows land in different banks)

__device__ __forceinline__ float dot16s(const float* a, const float* __restrict__ s) {
  const float4* s4 = reinterpret_cast<const float4*>(s);
  float r = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float4 t = s4[i];
    r = fmaf(a[4 * i], t.x, r); r = fmaf(a[4 * i + 1], t.y, r); r = fmaf(a[4 * i + 2], t.z, r); r = fmaf(a[4 * i + 3], t.w, r);
  }
  return r;
}

// CSR range [lo, hi) spanned by the windows of this CTA's 32 positions (identical in every warp)
__device__ __forceinline__ void cta_range(const Pos& p, int& lo, int& hi) {
  const uint32_t valid = __ballot_sync(0xffffffffu, p.valid);
  const int last = 31 - __clz(valid);
  lo = __shfl_sync(0xffffffffu, p.beg, 0);
  hi = __shfl_sync(0xffffffffu, p.beg + p.len, last);
}

// cooperative, coalesced staging of `cols4` float4 per row from two row-major sources into shared memory
template <int NSRC>
__device__ __forceinline__ void stage_rows(float* dst0, float* dst1, const float* __restrict__ src0, int ld0,
                                           const float* __restrict__ src1, int ld1, const int32_t* __restrict__ win_tok,
                                           int c0, int rows) {
  // each row: 32 float4 from src0 (+ 32 float4 from src1 when NSRC == 2)
  const int per_row = 32 * NSRC;
#pragma unroll 4
  for (int idx = threadIdx.x; idx < rows * per_row; idx += MAX_H * 32) {
    const int r = idx / per_row, c = idx % per_row;
    const int tok = __ldg(win_tok + c0 + r);
    if (NSRC == 1 || c < 32) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(src0 + (int64_t)tok * ld0) + c);
      *reinterpret_cast<float4*>(dst0 + r * RS + c * 4) = v;
    } else {
      const float4 v = __ldg(reinterpret_cast<const float4*>(src1 + (int64_t)tok * ld1) + (c - 32));
      *reinterpret_cast<float4*>(dst1 + r * RS + (c - 32) * 4) = v;
    }
  }
}

__global__ void __launch_bounds__(MAX_H * 32) k_sra_fwd(const float* __restrict__ qkv, int64_t n, int n_heads,
                                                        const int32_t* __restrict__ win_ptr,
                                                        const int32_t* __restrict__ win_tok,
                                                        const int32_t* __restrict__ tok_win, float* out, float* lse) {
  extern __shared__ __align__(16) float smem[];
  float* sK = smem;
  float* sV = smem + CH * RS;
  const int h = threadIdx.x >> 5;
  const int D = n_heads * HD, ld = 3 * D;
  const Pos p = locate(n, win_ptr, win_tok, tok_win);
  int lo, hi;
  cta_range(p, lo, hi);
  float q[HD], o[HD];
  if (p.valid) load_row16(qkv + (int64_t)p.tok * ld + h * HD, q);
#pragma unroll
  for (int d = 0; d < HD; ++d) { q[d] = p.valid ? q[d] * (0.25f * 1.4426950408889634f) : 0.f; o[d] = 0.f; }   // log2(e) folded in
  float m = -INFINITY, l = 0.f;
  for (int c0 = lo; c0 < hi; c0 += CH) {
    const int rows = min(CH, hi - c0);
    __syncthreads();
    stage_rows<2>(sK, sV, qkv + D, ld, qkv + 2 * D, ld, win_tok, c0, rows);
    __syncthreads();
    const int j0 = max(p.beg, c0) - c0, j1 = min(p.beg + p.len, c0 + rows) - c0;
    for (int j = j0; j < j1; ++j) {
      const float s = dot16s(q, sK + j * RS + h * HD);
      if (s > m) {                        // rescale only when the running max moves
        const float corr = exp2f(m - s);  // m = -inf on the first key -> 0
        l *= corr;
#pragma unroll
        for (int d = 0; d < HD; ++d) o[d] *= corr;
        m = s;
      }
      const float pr = exp2f(s - m);
      l += pr;
      const float4* v4 = reinterpret_cast<const float4*>(sV + j * RS + h * HD);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float4 t = v4[i];
        o[4 * i] = fmaf(pr, t.x, o[4 * i]); o[4 * i + 1] = fmaf(pr, t.y, o[4 * i + 1]);
        o[4 * i + 2] = fmaf(pr, t.z, o[4 * i + 2]); o[4 * i + 3] = fmaf(pr, t.w, o[4 * i + 3]);
      }
    }
  }
  if (p.valid) {
    const float inv = 1.0f / l;
#pragma unroll
    for (int d = 0; d < HD; ++d) o[d] *= inv;
    store_row16(out + (int64_t)p.tok * D + h * HD, o);
    lse[(int64_t)p.tok * n_heads + h] = (m + log2f(l)) * 0.6931471805599453f;     // natural-log LSE
  }
}

// pass A, thread = (query, head): dQ_i = 0.25 * sum_j P_ij (dO_i.v_j - D_i) k_j ; also publishes D_i = dO_i.O_i
__global__ void __launch_bounds__(MAX_H * 32) k_sra_bwd_q(const float* __restrict__ qkv, const float* __restrict__ out,
                                                          const float* __restrict__ lse,
                                                          const float* __restrict__ d_out, int64_t n, int n_heads,
                                                          const int32_t* __restrict__ win_ptr,
                                                          const int32_t* __restrict__ win_tok,
                                                          const int32_t* __restrict__ tok_win, float* d_qkv,
                                                          float* dd_out) {
  extern __shared__ __align__(16) float smem[];
  float* sK = smem;
  float* sV = smem + CH * RS;
  const int h = threadIdx.x >> 5;
  const int D = n_heads * HD, ld = 3 * D;
  const Pos p = locate(n, win_ptr, win_tok, tok_win);
  int lo, hi;
  cta_range(p, lo, hi);
  float q[HD], go[HD], dq[HD];
  float lse_i = 0.f, dd = 0.f;
  if (p.valid) {
    load_row16(qkv + (int64_t)p.tok * ld + h * HD, q);
    load_row16(d_out + (int64_t)p.tok * D + h * HD, go);
    float o[HD];
    load_row16(out + (int64_t)p.tok * D + h * HD, o);
    dd = dot16(go, o);
    lse_i = __ldg(lse + (int64_t)p.tok * n_heads + h) * 1.4426950408889634f;
    dd_out[(int64_t)p.tok * n_heads + h] = dd;
  }
#pragma unroll
  for (int d = 0; d < HD; ++d) {
    if (!p.valid) { q[d] = 0.f; go[d] = 0.f; }
    q[d] *= 0.25f * 1.4426950408889634f;
    dq[d] = 0.f;
  }
  for (int c0 = lo; c0 < hi; c0 += CH) {
    const int rows = min(CH, hi - c0);
    __syncthreads();
    stage_rows<2>(sK, sV, qkv + D, ld, qkv + 2 * D, ld, win_tok, c0, rows);
    __syncthreads();
    const int j0 = max(p.beg, c0) - c0, j1 = min(p.beg + p.len, c0 + rows) - c0;
    for (int j = j0; j < j1; ++j) {
      const float* kr = sK + j * RS + h * HD;
      const float pr = exp2f(dot16s(q, kr) - lse_i);
      const float ds = pr * (dot16s(go, sV + j * RS + h * HD) - dd);
      const float4* k4 = reinterpret_cast<const float4*>(kr);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float4 t = k4[i];
        dq[4 * i] = fmaf(ds, t.x, dq[4 * i]); dq[4 * i + 1] = fmaf(ds, t.y, dq[4 * i + 1]);
        dq[4 * i + 2] = fmaf(ds, t.z, dq[4 * i + 2]); dq[4 * i + 3] = fmaf(ds, t.w, dq[4 * i + 3]);
      }
    }
  }
  if (p.valid) {
#pragma unroll
    for (int d = 0; d < HD; ++d) dq[d] *= 0.25f;
    store_row16(d_qkv + (int64_t)p.tok * ld + h * HD, dq);
  }
}

// pass B, thread = (key, head): dV_j = sum_i P_ij dO_i ; dK_j = 0.25 * sum_i P_ij (dO_i.v_j - D_i) q_i
__global__ void __launch_bounds__(MAX_H * 32) k_sra_bwd_kv(const float* __restrict__ qkv, const float* __restrict__ lse,
                                                           const float* __restrict__ d_out,
                                                           const float* __restrict__ dd_in, int64_t n, int n_heads,
                                                           const int32_t* __restrict__ win_ptr,
                                                           const int32_t* __restrict__ win_tok,
                                                           const int32_t* __restrict__ tok_win, float* d_qkv) {
  extern __shared__ __align__(16) float smem[];
  float* sQ = smem;
  float* sG = smem + CH * RS;
  float* sL = sG + CH * RS;          // [CH][MAX_H] log2-domain LSE
  float* sD = sL + CH * MAX_H;       // [CH][MAX_H]
  const int h = threadIdx.x >> 5;
  const int D = n_heads * HD, ld = 3 * D;
  const Pos p = locate(n, win_ptr, win_tok, tok_win);
  int lo, hi;
  cta_range(p, lo, hi);
  float k[HD], v[HD], dk[HD], dv[HD];
  if (p.valid) {
    load_row16(qkv + (int64_t)p.tok * ld + D + h * HD, k);
    load_row16(qkv + (int64_t)p.tok * ld + 2 * D + h * HD, v);
  }
#pragma unroll
  for (int d = 0; d < HD; ++d) {
    if (!p.valid) { k[d] = 0.f; v[d] = 0.f; }
    k[d] *= 0.25f * 1.4426950408889634f;
    dk[d] = 0.f; dv[d] = 0.f;
  }
  for (int c0 = lo; c0 < hi; c0 += CH) {
    const int rows = min(CH, hi - c0);
    __syncthreads();
    stage_rows<2>(sQ, sG, qkv, ld, d_out, D, win_tok, c0, rows);
    for (int idx = threadIdx.x; idx < rows * n_heads; idx += MAX_H * 32) {
      const int r = idx / n_heads, hh = idx % n_heads;
      const int tok = __ldg(win_tok + c0 + r);
      sL[r * MAX_H + hh] = __ldg(lse + (int64_t)tok * n_heads + hh) * 1.4426950408889634f;
      sD[r * MAX_H + hh] = __ldg(dd_in + (int64_t)tok * n_heads + hh);
    }
    __syncthreads();
    const int i0 = max(p.beg, c0) - c0, i1 = min(p.beg + p.len, c0 + rows) - c0;
    for (int i = i0; i < i1; ++i) {
      const float* qr = sQ + i * RS + h * HD;
      const float* gr = sG + i * RS + h * HD;
      const float pr = exp2f(dot16s(k, qr) - sL[i * MAX_H + h]);
      const float ds = pr * (dot16s(v, gr) - sD[i * MAX_H + h]) * 0.25f;
      const float4* q4 = reinterpret_cast<const float4*>(qr);
      const float4* g4 = reinterpret_cast<const float4*>(gr);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float4 tq = q4[e], tg = g4[e];
        dv[4 * e] = fmaf(pr, tg.x, dv[4 * e]); dv[4 * e + 1] = fmaf(pr, tg.y, dv[4 * e + 1]);
        dv[4 * e + 2] = fmaf(pr, tg.z, dv[4 * e + 2]); dv[4 * e + 3] = fmaf(pr, tg.w, dv[4 * e + 3]);
        dk[4 * e] = fmaf(ds, tq.x, dk[4 * e]); dk[4 * e + 1] = fmaf(ds, tq.y, dk[4 * e + 1]);
        dk[4 * e + 2] = fmaf(ds, tq.z, dk[4 * e + 2]); dk[4 * e + 3] = fmaf(ds, tq.w, dk[4 * e + 3]);
      }
    }
  }
  if (p.valid) {
    store_row16(d_qkv + (int64_t)p.tok * ld + D + h * HD, dk);
    store_row16(d_qkv + (int64_t)p.tok * ld + 2 * D + h * HD, dv);
  }
}

constexpr int SMEM_FWD = 2 * CH * RS * 4;
constexpr int SMEM_KV = (2 * CH * RS + 2 * CH * MAX_H) * 4;

int configure_attention() {
  static bool done = false;
  if (!done) {
    GM_CUDA(cudaFuncSetAttribute(k_sra_fwd, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_FWD));
    GM_CUDA(cudaFuncSetAttribute(k_sra_bwd_q, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_FWD));
    GM_CUDA(cudaFuncSetAttribute(k_sra_bwd_kv, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_KV));
    done = true;
  }
  return GEOMAE_OK;
}

}  // namespace

extern "C" int geomae_sra_attention_fwd(const float* qkv, int64_t n_tokens, int32_t n_heads, const int32_t* win_ptr,
                                        const int32_t* win_tok, const int32_t* tok_win, float* out, float* lse,
                                        void* stream) {
  GM_REQUIRE(n_heads >= 1 && n_heads <= MAX_H, "sra_attention: n_heads %d not in 1..%d", n_heads, MAX_H);
  if (n_tokens == 0) return GEOMAE_OK;
  GM_REQUIRE(qkv && win_ptr && win_tok && tok_win && out && lse, "sra_attention_fwd: null argument");
  GM_REQUIRE(n_heads == MAX_H, "sra_attention: the staged kernels are built for %d heads (got %d)", MAX_H, n_heads);
  int rc = configure_attention();
  if (rc) return rc;
  k_sra_fwd<<<gm_div_up(n_tokens, 32), n_heads * 32, SMEM_FWD, (cudaStream_t)stream>>>(qkv, n_tokens, n_heads, win_ptr,
                                                                                win_tok, tok_win, out, lse);
  GM_LAUNCH_CHECK();
  return GEOMAE_OK;
}

extern "C" int geomae_sra_attention_bwd(const float* qkv, const float* out, const float* lse, const float* d_out,
                                        int64_t n_tokens, int32_t n_heads, const int32_t* win_ptr,
                                        const int32_t* win_tok, const int32_t* tok_win, float* d_qkv, float* scratch,
                                        void* stream) {
  GM_REQUIRE(n_heads >= 1 && n_heads <= MAX_H, "sra_attention: n_heads %d not in 1..%d", n_heads, MAX_H);
  if (n_tokens == 0) return GEOMAE_OK;
  GM_REQUIRE(qkv && out && lse && d_out && win_ptr && win_tok && tok_win && d_qkv && scratch,
             "sra_attention_bwd: null argument");
  GM_REQUIRE(n_heads == MAX_H, "sra_attention: the staged kernels are built for %d heads (got %d)", MAX_H, n_heads);
  int rc = configure_attention();
  if (rc) return rc;
  const int grid = gm_div_up(n_tokens, 32);
  k_sra_bwd_q<<<grid, n_heads * 32, SMEM_FWD, (cudaStream_t)stream>>>(qkv, out, lse, d_out, n_tokens, n_heads, win_ptr,
                                                               win_tok, tok_win, d_qkv, scratch);
  k_sra_bwd_kv<<<grid, n_heads * 32, SMEM_KV, (cudaStream_t)stream>>>(qkv, lse, d_out, scratch, n_tokens, n_heads, win_ptr,
                                                                win_tok, tok_win, d_qkv);
  GM_LAUNCH_CHECK();
  return GEOMAE_OK;
}
