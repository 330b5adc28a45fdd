// Host-side construction of 2-D TMA tensor maps (cuTensorMapEncodeTiled through the runtime's driver entry point, so
// the library has no link dependency on libcuda) and the device-side bulk-tensor copy wrappers that use them.
#pragma once
#include <cuda.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace tma {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

// Row-major [n_rows, cols] tensor of 2-byte (bf16) or 4-byte (fp32) elements, boxes of [128 rows x 128 bytes] with the
// 128-byte swizzle (the layout of every tensor-core operand tile in this library); rows past n_rows read as zeros and
// are dropped on stores.
inline int make_map(CUtensorMap* map, const void* base, int64_t n_rows, int cols, int elem_bytes) {
  EncodeTiledFn fn = encode_fn();
  GM_REQUIRE(fn, "tma: cuTensorMapEncodeTiled is not available from this driver");
  GM_REQUIRE(((uintptr_t)base & 15) == 0, "tma: tensors must be 16-byte aligned");
  GM_REQUIRE(elem_bytes == 2 || elem_bytes == 4, "tma: 2- or 4-byte elements");
  const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)n_rows};
  const cuuint64_t strides[1] = {(cuuint64_t)cols * elem_bytes};
  const cuuint32_t box[2] = {(cuuint32_t)(128 / elem_bytes), 128};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult rc = fn(map, elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
                         const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  GM_REQUIRE(rc == CUDA_SUCCESS, "tma: cuTensorMapEncodeTiled failed (%d)", (int)rc);
  return GEOMAE_OK;
}

// global -> shared box load (UTMALDG), completion counted in bytes on an mbarrier
__device__ __forceinline__ void load_2d(void* smem_dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
          tc::smem_u32(smem_dst)),
      "l"(map), "r"(c0), "r"(c1), "r"(tc::smem_u32(bar))
      : "memory");
}
// shared -> global box store (UTMASTG), tracked by the issuing thread's bulk async-groups
__device__ __forceinline__ void store_2d(const CUtensorMap* map, int c0, int c1, const void* smem_src) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%1, %2}], [%3];" ::"l"(map), "r"(c0),
               "r"(c1), "r"(tc::smem_u32(smem_src))
               : "memory");
}
__device__ __forceinline__ void store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// all of this thread's committed stores have finished READING shared memory (the source may be overwritten)
__device__ __forceinline__ void store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
// ... have completed entirely
__device__ __forceinline__ void store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

}  // namespace tma
