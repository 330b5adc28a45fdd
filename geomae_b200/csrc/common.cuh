// Shared helpers for the geomae_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/geomae_b200.h"

void gm_set_error(const char* fmt, ...);

#define GM_REQUIRE(cond, ...)          \
  do {                                 \
    if (!(cond)) {                     \
      gm_set_error(__VA_ARGS__);       \
      return GEOMAE_ERR_INVALID;       \
    }                                  \
  } while (0)

#define GM_CUDA(call)                                                          \
  do {                                                                         \
    cudaError_t e__ = (call);                                                  \
    if (e__ != cudaSuccess) {                                                  \
      gm_set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call,                \
                   cudaGetErrorString(e__));                                   \
      return GEOMAE_ERR_CUDA;                                                  \
    }                                                                          \
  } while (0)

#define GM_LAUNCH_CHECK() GM_CUDA(cudaGetLastError())

static inline int gm_div_up(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// ---- programmatic dependent launch (PDL).  Kernels of the SRA chain are launched with
// cudaLaunchAttributeProgrammaticStreamSerialization: the next kernel's CTAs may become resident and run their
// prologue (barrier init, TMEM allocation, weight-image bulk copies) while the previous kernel drains.  Every such
// kernel executes gm_pdl_wait() BEFORE it touches memory written by an earlier kernel and only then
// gm_pdl_trigger() — so at most one successor is ever in flight and a successor's pre-wait work can only race with
// its direct predecessor.
__device__ __forceinline__ void gm_pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void gm_pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// true while the caller guarantees that the packed weight images a dense kernel reads were written before its
// PREDECESSOR kernel started (the stack executor sets it for every kernel but the first after the packing launch);
// only then may a kernel fetch weights ahead of gm_pdl_wait().
void gm_set_weights_stable(bool stable);
bool gm_weights_stable();

template <typename... KArgs, typename... Args>
static inline cudaError_t gm_launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                                        Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

// B200: 148 SMs.  Grid-stride kernels are sized in multiples of this.
constexpr int GM_NUM_SMS = 148;

__device__ __forceinline__ float gm_warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float gm_warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ int gm_warp_sum_i(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Block-wide exclusive scan of one int per thread (blockDim.x <= 1024, multiple of 32).
// Returns the exclusive prefix; *total receives the block sum (valid in every thread).
__device__ __forceinline__ int gm_block_excl_scan(int v, int* total, int* smem /*>=33 ints*/) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) smem[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    int w = lane < nwarp ? smem[lane] : 0;
    int winc = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int t = __shfl_up_sync(0xffffffffu, winc, o);
      if (lane >= o) winc += t;
    }
    smem[lane] = winc - w;  // exclusive warp offsets
    if (lane == 31) smem[32] = winc;
  }
  __syncthreads();
  int res = smem[warp] + inc - v;
  *total = smem[32];
  __syncthreads();
  return res;
}
