// Sparse Regional Attention core over CSR windows, fp32.
//
// The reference scatters tokens into zero-padded [W, 56|144, C] buckets, runs nn.MultiheadAttention
// with a key_padding_mask (11.5x padding in the encoder) and gathers the result back.  Windows here
// are short (mean 5 tokens in the encoder, 14 in the decoders, max 144) and head_dim is 16, so the
// work per window is far too small for a CTA: the kernel is latency-bound, not FLOP-bound (a whole
// decoder launch is ~0.3 GFLOP).  Mapping: ONE THREAD = one (query token, head); a warp = 32
// consecutive CSR positions of one head, so its lanes sit in 2-3 neighbouring windows and read the
// same K/V rows (hardware broadcast, L1-resident); no shared memory, no barriers, no padding, no
// mask, no [W,T,T] attention map.  The backward is two such passes (as query: dQ; as key: dK, dV),
// each output element written exactly once — no atomics.
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

__global__ void __launch_bounds__(MAX_H * 32) k_sra_fwd(const float* __restrict__ qkv, int64_t n, int n_heads,
                                                        const int32_t* __restrict__ win_ptr,
                                                        const int32_t* __restrict__ win_tok,
                                                        const int32_t* __restrict__ tok_win, float* out, float* lse) {
  const int h = threadIdx.x >> 5;
  const int D = n_heads * HD, ld = 3 * D;
  const Pos p = locate(n, win_ptr, win_tok, tok_win);
  float q[HD], o[HD];
  if (p.valid) load_row16(qkv + (int64_t)p.tok * ld + h * HD, q);
#pragma unroll
  for (int d = 0; d < HD; ++d) { q[d] = p.valid ? q[d] * 0.25f : 0.f; o[d] = 0.f; }
  float m = -INFINITY, l = 0.f;
  int maxlen = p.len;
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) maxlen = max(maxlen, __shfl_xor_sync(0xffffffffu, maxlen, s));
  int tk = p.len > 0 ? __ldg(win_tok + p.beg) : 0;
  for (int j = 0; j < maxlen; ++j) {
    const bool on = j < p.len;
    const int tk_next = (j + 1 < p.len) ? __ldg(win_tok + p.beg + j + 1) : 0;   // prefetch the next key's row id
    if (on) {
      float k[HD], v[HD];
      load_row16(qkv + (int64_t)tk * ld + D + h * HD, k);
      load_row16(qkv + (int64_t)tk * ld + 2 * D + h * HD, v);
      const float s = dot16(q, k);
      if (s > m) {                       // rescale only when the running max moves
        const float corr = expf(m - s);  // m = -inf on the first key -> 0
        l *= corr;
#pragma unroll
        for (int d = 0; d < HD; ++d) o[d] *= corr;
        m = s;
      }
      const float pr = expf(s - m);
      l += pr;
#pragma unroll
      for (int d = 0; d < HD; ++d) o[d] = fmaf(pr, v[d], o[d]);
    }
    tk = tk_next;
  }
  if (p.valid) {
    const float inv = 1.0f / l;
#pragma unroll
    for (int d = 0; d < HD; ++d) o[d] *= inv;
    store_row16(out + (int64_t)p.tok * D + h * HD, o);
    lse[(int64_t)p.tok * n_heads + h] = m + logf(l);
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
  const int h = threadIdx.x >> 5;
  const int D = n_heads * HD, ld = 3 * D;
  const Pos p = locate(n, win_ptr, win_tok, tok_win);
  float q[HD], go[HD], dq[HD];
  float lse_i = 0.f, dd = 0.f;
  if (p.valid) {
    load_row16(qkv + (int64_t)p.tok * ld + h * HD, q);
    load_row16(d_out + (int64_t)p.tok * D + h * HD, go);
    float o[HD];
    load_row16(out + (int64_t)p.tok * D + h * HD, o);
    dd = dot16(go, o);
    lse_i = __ldg(lse + (int64_t)p.tok * n_heads + h);
    dd_out[(int64_t)p.tok * n_heads + h] = dd;
  }
#pragma unroll
  for (int d = 0; d < HD; ++d) { if (!p.valid) { q[d] = 0.f; go[d] = 0.f; } q[d] *= 0.25f; dq[d] = 0.f; }
  int maxlen = p.len;
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) maxlen = max(maxlen, __shfl_xor_sync(0xffffffffu, maxlen, s));
  int tk = p.len > 0 ? __ldg(win_tok + p.beg) : 0;
  for (int j = 0; j < maxlen; ++j) {
    const int tk_next = (j + 1 < p.len) ? __ldg(win_tok + p.beg + j + 1) : 0;
    if (j < p.len) {
      float k[HD], v[HD];
      load_row16(qkv + (int64_t)tk * ld + D + h * HD, k);
      load_row16(qkv + (int64_t)tk * ld + 2 * D + h * HD, v);
      const float pr = expf(dot16(q, k) - lse_i);
      const float ds = pr * (dot16(go, v) - dd);
#pragma unroll
      for (int d = 0; d < HD; ++d) dq[d] = fmaf(ds, k[d], dq[d]);
    }
    tk = tk_next;
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
  const int h = threadIdx.x >> 5;
  const int D = n_heads * HD, ld = 3 * D;
  const Pos p = locate(n, win_ptr, win_tok, tok_win);
  float k[HD], v[HD], dk[HD], dv[HD];
  if (p.valid) {
    load_row16(qkv + (int64_t)p.tok * ld + D + h * HD, k);
    load_row16(qkv + (int64_t)p.tok * ld + 2 * D + h * HD, v);
  }
#pragma unroll
  for (int d = 0; d < HD; ++d) { if (!p.valid) { k[d] = 0.f; v[d] = 0.f; } k[d] *= 0.25f; dk[d] = 0.f; dv[d] = 0.f; }
  int maxlen = p.len;
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) maxlen = max(maxlen, __shfl_xor_sync(0xffffffffu, maxlen, s));
  int tq = p.len > 0 ? __ldg(win_tok + p.beg) : 0;
  for (int i = 0; i < maxlen; ++i) {
    const int tq_next = (i + 1 < p.len) ? __ldg(win_tok + p.beg + i + 1) : 0;
    if (i < p.len) {
      float q[HD], go[HD];
      load_row16(qkv + (int64_t)tq * ld + h * HD, q);
      load_row16(d_out + (int64_t)tq * D + h * HD, go);
      const float lse_i = __ldg(lse + (int64_t)tq * n_heads + h);
      const float dd = __ldg(dd_in + (int64_t)tq * n_heads + h);
      const float pr = expf(dot16(q, k) - lse_i);          // k already carries the 1/sqrt(hd) scale
      const float ds = pr * (dot16(go, v) - dd) * 0.25f;
#pragma unroll
      for (int d = 0; d < HD; ++d) {
        dv[d] = fmaf(pr, go[d], dv[d]);
        dk[d] = fmaf(ds, q[d], dk[d]);
      }
    }
    tq = tq_next;
  }
  if (p.valid) {
    store_row16(d_qkv + (int64_t)p.tok * ld + D + h * HD, dk);
    store_row16(d_qkv + (int64_t)p.tok * ld + 2 * D + h * HD, dv);
  }
}

}  // namespace

extern "C" int geomae_sra_attention_fwd(const float* qkv, int64_t n_tokens, int32_t n_heads, const int32_t* win_ptr,
                                        const int32_t* win_tok, const int32_t* tok_win, float* out, float* lse,
                                        void* stream) {
  GM_REQUIRE(n_heads >= 1 && n_heads <= MAX_H, "sra_attention: n_heads %d not in 1..%d", n_heads, MAX_H);
  if (n_tokens == 0) return GEOMAE_OK;
  GM_REQUIRE(qkv && win_ptr && win_tok && tok_win && out && lse, "sra_attention_fwd: null argument");
  k_sra_fwd<<<gm_div_up(n_tokens, 32), n_heads * 32, 0, (cudaStream_t)stream>>>(qkv, n_tokens, n_heads, win_ptr,
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
  const int grid = gm_div_up(n_tokens, 32);
  k_sra_bwd_q<<<grid, n_heads * 32, 0, (cudaStream_t)stream>>>(qkv, out, lse, d_out, n_tokens, n_heads, win_ptr,
                                                               win_tok, tok_win, d_qkv, scratch);
  k_sra_bwd_kv<<<grid, n_heads * 32, 0, (cudaStream_t)stream>>>(qkv, lse, d_out, scratch, n_tokens, n_heads, win_ptr,
                                                                win_tok, tok_win, d_qkv);
  GM_LAUNCH_CHECK();
  return GEOMAE_OK;
}
