// tcgen05 / TMEM / mbarrier building blocks for the SRA dense layers (sm_100a inline PTX).
//
// Operand staging convention used by every tensor-core kernel in this library: a [rows x 64] bf16
// block is stored as `rows` consecutive 128-byte lines with the hardware 128B swizzle (16-byte chunk
// index XOR (row & 7)); blocks of further 64 columns follow at a fixed block stride.  The SAME bytes
// serve two descriptor views:
//   * K-major  (rows = M or N index, the 64 columns = a K slice)           -> y = x W^T style products
//   * MN-major (rows = K index,      the 64 columns = an M or N slice)     -> dX = dY W and dW = dY^T X
// so activations [tokens, channels] and weights [out, in] are staged exactly as they lie in memory and
// no transposed copy ever exists.
#pragma once
#include <cuda_bf16.h>

#include "common.cuh"

namespace tc {

constexpr int BLK_COLS = 64;          // bf16 columns per 128-byte line
constexpr int LINE_BYTES = 128;
constexpr int ATOM_BYTES = 1024;      // 8 lines

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout)
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);          // [0,14)  start address >> 4
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;     // [16,30) leading byte offset >> 4
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;     // [32,46) stride byte offset >> 4
  d |= (uint64_t)1 << 46;                                // [46,48) version = 1 (Blackwell)
  d |= (uint64_t)2 << 61;                                // [61,64) layout = SWIZZLE_128B
  return d;
}

// ---- instruction descriptor for kind::f16, bf16 x bf16 -> f32 (cute::UMMA::InstrDescriptor bit layout)
__device__ __forceinline__ uint32_t make_idesc_bf16(int M, int N, int a_mn_major, int b_mn_major) {
  return (1u << 4)                      // c_format = F32
         | (1u << 7)                    // a_format = BF16
         | (1u << 10)                   // b_format = BF16
         | ((uint32_t)a_mn_major << 15) // 0 = K-major
         | ((uint32_t)b_mn_major << 16)
         | ((uint32_t)(N >> 3) << 17)
         | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ void mma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                         bool accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"((uint32_t)accumulate)
      : "memory");
}

// arrive on an mbarrier once every MMA issued so far by this thread has completed
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// make generic-proxy smem writes visible to the async proxy (tcgen05.mma operand reads)
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t a = smem_u32(bar);
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}\n" ::"r"(a),
      "r"(parity)
      : "memory");
}

// ---- bulk asynchronous copy global -> shared (UBLKCP), completion counted on an mbarrier
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// ---- TMEM allocation (one full warp), power-of-two columns >= 32
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_slot, int cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_slot)),
               "r"(cols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_free(uint32_t taddr, int cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}

// 32 lanes x 32 consecutive fp32 columns: thread t of the warp receives lane (warp%4)*32+t
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---- staging: byte offset of (row, 16-byte chunk) inside one [rows x 64] swizzled block
__device__ __forceinline__ uint32_t swz(int row, int chunk) {
  return (uint32_t)row * LINE_BYTES + (uint32_t)((chunk ^ (row & 7)) << 4);
}

// pack 8 floats into 8 bf16 (16 bytes); optionally also the residuals a - bf16(a) ("lo" operand)
__device__ __forceinline__ uint4 pack8(const float* f, uint4* lo) {
  __nv_bfloat162 h[4], l[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    h[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
    if (lo) {
      const float2 back = __bfloat1622float2(h[i]);
      l[i] = __floats2bfloat162_rn(f[2 * i] - back.x, f[2 * i + 1] - back.y);
    }
  }
  if (lo) *lo = *reinterpret_cast<uint4*>(l);
  return *reinterpret_cast<uint4*>(h);
}

}  // namespace tc

// Exact-erf GELU (F.gelu default, sst_basic_block.py:153-154) and its derivative, evaluated with the
// Abramowitz-Stegun 7.1.26 rational form erf(z) = 1 - (a1 t + .. + a5 t^5) exp(-z^2), t = 1/(1 + p z)
// (|error| <= 1.5e-7, i.e. fp32 rounding level): one MUFU.RCP + one MUFU.EX2 + 8 FMAs.  libdevice erff + expf
// cost ~70 dependent instructions per element, which made the GELU prologue / GELU-gradient epilogue the
// longest phase of three kernels (128 evaluations per thread at 8 warps per SM).  exp(-z^2) = exp(-x^2/2) is
// shared between erf and the Gaussian density of the derivative.
struct GeluParts { float cdf; float pdf_x; };    // Phi(x), x * phi(x)
__device__ __forceinline__ GeluParts gelu_parts(float x) {
  const float z = fabsf(x) * 0.70710678118654752f;
  float t, e;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, z, 1.0f)));   // argument in [1, inf): 1 ulp
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-z * z * 1.4426950408889634f));  // 2 ulp; flushes to 0 below 2^-126
  float p = fmaf(t, 1.061405429f, -1.453152027f);
  p = fmaf(t, p, 1.421413741f);
  p = fmaf(t, p, -0.284496736f);
  p = fmaf(t, p, 0.254829592f);
  const float erf_abs = fmaf(-p * t, e, 1.0f);
  GeluParts r;
  r.cdf = 0.5f * (1.0f + copysignf(erf_abs, x));
  r.pdf_x = x * e * 0.3989422804014327f;
  return r;
}
__device__ __forceinline__ float gelu_f(float x) { return x * gelu_parts(x).cdf; }
__device__ __forceinline__ float gelu_grad_f(float x) {
  const GeluParts g = gelu_parts(x);
  return g.cdf + g.pdf_x;
}
