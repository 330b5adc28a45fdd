// Decoder token assembly (backbones/multi_mae_sst_spearate_top_only.py:204-210): tokens = [visible features ;
// mask_token repeated for every masked pillar], and its backward: the gradient of the visible rows is the leading
// slice of the token gradient (no copy), the mask token's gradient is the column sum of the trailing rows.
// One launch each instead of repeat + cat / reduce + accumulate through the tensor library.
#include "common.cuh"

namespace {

__global__ void __launch_bounds__(256) k_tokens_fwd(const float4* __restrict__ vis, int64_t n_vis4,
                                                    const float4* __restrict__ mask_token, int d4, int64_t total4,
                                                    float4* out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = i < n_vis4 ? __ldg(vis + i) : __ldg(mask_token + (int)((i - n_vis4) % d4));
}

// grad[c] += sum over rows of x[r][c]; 256 threads = 8 row lanes x 32 column groups of 4 floats (d = 128)
__global__ void __launch_bounds__(256) k_colsum_add128(const float4* __restrict__ x, int64_t n_rows, float* grad) {
  __shared__ float4 part[8][32];
  const int cg = threadIdx.x & 31, rl = threadIdx.x >> 5;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int64_t r = (int64_t)blockIdx.x * 8 + rl; r < n_rows; r += (int64_t)gridDim.x * 8) {
    const float4 v = __ldg(x + r * 32 + cg);
    acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
  }
  part[rl][cg] = acc;
  __syncthreads();
  if (rl == 0) {
#pragma unroll
    for (int k = 1; k < 8; ++k) {
      const float4 v = part[k][cg];
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    atomicAdd(grad + cg * 4 + 0, acc.x);
    atomicAdd(grad + cg * 4 + 1, acc.y);
    atomicAdd(grad + cg * 4 + 2, acc.z);
    atomicAdd(grad + cg * 4 + 3, acc.w);
  }
}

}  // namespace

extern "C" int geomae_decoder_tokens(const float* visible, int64_t n_vis, const float* mask_token, int64_t n_mask,
                                     int32_t d_model, float* tokens, void* stream) {
  GM_REQUIRE(tokens && mask_token && (visible || n_vis == 0), "decoder_tokens: null argument");
  GM_REQUIRE(d_model > 0 && d_model % 4 == 0 && n_vis >= 0 && n_mask >= 0, "decoder_tokens: bad sizes");
  const int64_t total4 = (n_vis + n_mask) * (d_model / 4);
  if (total4 == 0) return GEOMAE_OK;
  const int blocks = (int)min((int64_t)GM_NUM_SMS * 8, (total4 + 255) / 256);
  k_tokens_fwd<<<blocks, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float4*>(visible), n_vis * (d_model / 4),
                                                         reinterpret_cast<const float4*>(mask_token), d_model / 4, total4,
                                                         reinterpret_cast<float4*>(tokens));
  GM_LAUNCH_CHECK();
  return GEOMAE_OK;
}

extern "C" int geomae_mask_token_grad(const float* d_tokens, int64_t n_vis, int64_t n_mask, int32_t d_model,
                                      float* g_mask_token, void* stream) {
  GM_REQUIRE(g_mask_token && (d_tokens || n_mask == 0), "mask_token_grad: null argument");
  GM_REQUIRE(d_model == 128, "mask_token_grad: d_model %d, this kernel is built for 128", d_model);
  if (n_mask == 0) return GEOMAE_OK;
  const int blocks = (int)min((int64_t)GM_NUM_SMS * 2, (n_mask + 63) / 64);
  k_colsum_add128<<<blocks, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float4*>(d_tokens + n_vis * d_model),
                                                            n_mask, g_mask_token);
  GM_LAUNCH_CHECK();
  return GEOMAE_OK;
}
