// Sparse Regional Attention core over CSR windows, fp32.
//
// The reference scatters tokens into zero-padded [W, 56|144, C] buckets, runs nn.MultiheadAttention
// with a key_padding_mask (11.5x padding in the encoder) and gathers the result back.  Here a CTA
// owns one variable-length window straight from the CSR list; warp h owns head h; lanes are query
// rows; K/V rows are staged in 32-key chunks through shared memory (conflict-free stride 20) and the
// softmax runs online over the chunks, so there is no padding, no mask and no [W,T,T] attention map.
// head_dim = 16: a score costs 16 FMAs against one exp, so this kernel is SFU/FMA bound, not
// tensor-pipe bound; the K=128 projections around it are where the tensor cores go.
#include "common.cuh"

namespace {

constexpr int HD = 16;      // head_dim
constexpr int RS = 20;      // shared-memory row stride in floats (16 + 4 pad: conflict-free float4 row stores)
constexpr int MAX_H = 8;    // heads per CTA (one warp each)
constexpr int CH = 32;      // keys (or queries) staged per chunk

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

__global__ void __launch_bounds__(MAX_H * 32) k_sra_fwd(const float* __restrict__ qkv, int n_heads,
                                                        const int32_t* __restrict__ win_ptr,
                                                        const int32_t* __restrict__ win_tok,
                                                        const int32_t* __restrict__ n_windows, float* out,
                                                        float* lse) {
  __shared__ __align__(16) float sk[MAX_H][CH * RS];
  __shared__ __align__(16) float sv[MAX_H][CH * RS];
  const int w = blockIdx.x;
  if (w >= *n_windows) return;
  const int h = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int D = n_heads * HD, ld = 3 * D;
  const int beg = win_ptr[w], L = win_ptr[w + 1] - beg;
  float* myk = sk[h];
  float* myv = sv[h];
  for (int q0 = 0; q0 < L; q0 += CH) {
    const bool vq = q0 + lane < L;
    const int tq = vq ? win_tok[beg + q0 + lane] : 0;
    float q[HD], o[HD];
    if (vq) load_row16(qkv + (int64_t)tq * ld + h * HD, q);
#pragma unroll
    for (int d = 0; d < HD; ++d) { q[d] = vq ? q[d] * 0.25f : 0.f; o[d] = 0.f; }
    float m = -INFINITY, l = 0.f;
    for (int k0 = 0; k0 < L; k0 += CH) {
      const int nk = min(CH, L - k0);
      __syncwarp();
      if (lane < nk) {
        const int tk = win_tok[beg + k0 + lane];
        float r[HD];
        load_row16(qkv + (int64_t)tk * ld + D + h * HD, r);
        store_row16(myk + lane * RS, r);
        load_row16(qkv + (int64_t)tk * ld + 2 * D + h * HD, r);
        store_row16(myv + lane * RS, r);
      }
      __syncwarp();
      float s[CH];
      float cmax = -INFINITY;
#pragma unroll
      for (int j = 0; j < CH; ++j) {
        s[j] = (j < nk) ? dot16(q, myk + j * RS) : -INFINITY;
        cmax = fmaxf(cmax, s[j]);
      }
      const float m_new = fmaxf(m, cmax);
      const float corr = expf(m - m_new);  // m = -inf on the first chunk -> 0
      l *= corr;
#pragma unroll
      for (int d = 0; d < HD; ++d) o[d] *= corr;
#pragma unroll
      for (int j = 0; j < CH; ++j) {
        if (j < nk) {
          const float p = expf(s[j] - m_new);
          l += p;
#pragma unroll
          for (int d = 0; d < HD; ++d) o[d] = fmaf(p, myv[j * RS + d], o[d]);
        }
      }
      m = m_new;
    }
    if (vq) {
      const float inv = 1.0f / l;
#pragma unroll
      for (int d = 0; d < HD; ++d) o[d] *= inv;
      store_row16(out + (int64_t)tq * D + h * HD, o);
      lse[(int64_t)tq * n_heads + h] = m + logf(l);
    }
  }
}

// Backward: pass A (lanes = queries) produces dQ, pass B (lanes = keys) produces dK and dV; both
// recompute P from the saved log-sum-exp, so every output element is written exactly once.
__global__ void __launch_bounds__(MAX_H * 32) k_sra_bwd(const float* __restrict__ qkv, const float* __restrict__ out,
                                                        const float* __restrict__ lse,
                                                        const float* __restrict__ d_out, int n_heads,
                                                        const int32_t* __restrict__ win_ptr,
                                                        const int32_t* __restrict__ win_tok,
                                                        const int32_t* __restrict__ n_windows, float* d_qkv) {
  __shared__ __align__(16) float sa[MAX_H][CH * RS];
  __shared__ __align__(16) float sb[MAX_H][CH * RS];
  __shared__ float s_lse[MAX_H][CH], s_dd[MAX_H][CH];
  const int w = blockIdx.x;
  if (w >= *n_windows) return;
  const int h = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int D = n_heads * HD, ld = 3 * D;
  const int beg = win_ptr[w], L = win_ptr[w + 1] - beg;
  float* a_ = sa[h];
  float* b_ = sb[h];
  // ---- pass A: dQ_i = 0.25 * sum_j P_ij (dO_i.v_j - D_i) k_j
  for (int q0 = 0; q0 < L; q0 += CH) {
    const bool vq = q0 + lane < L;
    const int tq = vq ? win_tok[beg + q0 + lane] : 0;
    float q[HD], go[HD], dq[HD];
    float lse_i = 0.f, dd = 0.f;
    if (vq) {
      load_row16(qkv + (int64_t)tq * ld + h * HD, q);
      load_row16(d_out + (int64_t)tq * D + h * HD, go);
      float o[HD];
      load_row16(out + (int64_t)tq * D + h * HD, o);
      dd = dot16(go, o);
      lse_i = lse[(int64_t)tq * n_heads + h];
    }
#pragma unroll
    for (int d = 0; d < HD; ++d) { if (!vq) { q[d] = 0.f; go[d] = 0.f; } q[d] *= 0.25f; dq[d] = 0.f; }
    for (int k0 = 0; k0 < L; k0 += CH) {
      const int nk = min(CH, L - k0);
      __syncwarp();
      if (lane < nk) {
        const int tk = win_tok[beg + k0 + lane];
        float r[HD];
        load_row16(qkv + (int64_t)tk * ld + D + h * HD, r);
        store_row16(a_ + lane * RS, r);
        load_row16(qkv + (int64_t)tk * ld + 2 * D + h * HD, r);
        store_row16(b_ + lane * RS, r);
      }
      __syncwarp();
      if (vq) {
        for (int j = 0; j < nk; ++j) {
          const float p = expf(dot16(q, a_ + j * RS) - lse_i);
          const float ds = p * (dot16(go, b_ + j * RS) - dd);
#pragma unroll
          for (int d = 0; d < HD; ++d) dq[d] = fmaf(ds, a_[j * RS + d], dq[d]);
        }
      }
    }
    if (vq) {
#pragma unroll
      for (int d = 0; d < HD; ++d) dq[d] *= 0.25f;
      store_row16(d_qkv + (int64_t)tq * ld + h * HD, dq);
    }
  }
  // ---- pass B: dV_j = sum_i P_ij dO_i ; dK_j = 0.25 * sum_i P_ij (dO_i.v_j - D_i) q_i
  for (int k0 = 0; k0 < L; k0 += CH) {
    const bool vk = k0 + lane < L;
    const int tk = vk ? win_tok[beg + k0 + lane] : 0;
    float k[HD], v[HD], dk[HD], dv[HD];
    if (vk) {
      load_row16(qkv + (int64_t)tk * ld + D + h * HD, k);
      load_row16(qkv + (int64_t)tk * ld + 2 * D + h * HD, v);
    }
#pragma unroll
    for (int d = 0; d < HD; ++d) { if (!vk) { k[d] = 0.f; v[d] = 0.f; } dk[d] = 0.f; dv[d] = 0.f; }
    for (int q0 = 0; q0 < L; q0 += CH) {
      const int nq = min(CH, L - q0);
      __syncwarp();
      if (lane < nq) {
        const int tq = win_tok[beg + q0 + lane];
        float r[HD], g[HD], o[HD];
        load_row16(qkv + (int64_t)tq * ld + h * HD, r);
        load_row16(d_out + (int64_t)tq * D + h * HD, g);
        load_row16(out + (int64_t)tq * D + h * HD, o);
        s_dd[h][lane] = dot16(g, o);
        s_lse[h][lane] = lse[(int64_t)tq * n_heads + h];
#pragma unroll
        for (int d = 0; d < HD; ++d) r[d] *= 0.25f;
        store_row16(a_ + lane * RS, r);
        store_row16(b_ + lane * RS, g);
      }
      __syncwarp();
      if (vk) {
        for (int i = 0; i < nq; ++i) {
          const float p = expf(dot16(a_ + i * RS, k) - s_lse[h][i]);
          const float ds = p * (dot16(b_ + i * RS, v) - s_dd[h][i]);
#pragma unroll
          for (int d = 0; d < HD; ++d) {
            dv[d] = fmaf(p, b_[i * RS + d], dv[d]);
            dk[d] = fmaf(ds, a_[i * RS + d], dk[d]);  // a_ holds 0.25*q
          }
        }
      }
    }
    if (vk) {
      store_row16(d_qkv + (int64_t)tk * ld + D + h * HD, dk);
      store_row16(d_qkv + (int64_t)tk * ld + 2 * D + h * HD, dv);
    }
  }
}

}  // namespace

extern "C" int geomae_sra_attention_fwd(const float* qkv, int64_t n_tokens, int32_t n_heads, const int32_t* win_ptr,
                                        const int32_t* win_tok, const int32_t* n_windows, int32_t max_windows,
                                        float* out, float* lse, void* stream) {
  GM_REQUIRE(n_heads >= 1 && n_heads <= MAX_H, "sra_attention: n_heads %d not in 1..%d", n_heads, MAX_H);
  if (n_tokens == 0 || max_windows == 0) return GEOMAE_OK;
  GM_REQUIRE(qkv && win_ptr && win_tok && n_windows && out && lse, "sra_attention_fwd: null argument");
  k_sra_fwd<<<max_windows, n_heads * 32, 0, (cudaStream_t)stream>>>(qkv, n_heads, win_ptr, win_tok, n_windows, out,
                                                                    lse);
  GM_LAUNCH_CHECK();
  return GEOMAE_OK;
}

extern "C" int geomae_sra_attention_bwd(const float* qkv, const float* out, const float* lse, const float* d_out,
                                        int64_t n_tokens, int32_t n_heads, const int32_t* win_ptr,
                                        const int32_t* win_tok, const int32_t* n_windows, int32_t max_windows,
                                        float* d_qkv, void* stream) {
  GM_REQUIRE(n_heads >= 1 && n_heads <= MAX_H, "sra_attention: n_heads %d not in 1..%d", n_heads, MAX_H);
  if (n_tokens == 0 || max_windows == 0) return GEOMAE_OK;
  GM_REQUIRE(qkv && out && lse && d_out && win_ptr && win_tok && n_windows && d_qkv,
             "sra_attention_bwd: null argument");
  k_sra_bwd<<<max_windows, n_heads * 32, 0, (cudaStream_t)stream>>>(qkv, out, lse, d_out, n_heads, win_ptr, win_tok,
                                                                    n_windows, d_qkv);
  GM_LAUNCH_CHECK();
  return GEOMAE_OK;
}
