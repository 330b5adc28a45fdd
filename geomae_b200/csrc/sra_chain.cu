// Token-local chain of one SRA EncoderLayer as ONE persistent, warp-specialised tcgen05 kernel (bf16 mode):
//
//   forward  (k_sra_chain_fwd):  s1 = x + O Wo^T + bo ; y = LN1(s1) ; u = y W1^T + b1 ; g = gelu(u) ;
//                                s2 = y + g W2^T + b2 ; z = LN2(s2) ; q|k|v(next layer) = (z + pos | z) Win'^T + bin'
//
// models/sst/sst_basic_block.py:26-61 (out_proj of nn.MultiheadAttention, the in-projection of the NEXT layer's
// attention), :85-102 (post-norm residual blocks, exact-erf GELU feed-forward).  Everything here is token-local: the
// only window-local step of a layer, the attention core, stays in sra_attention_tc.cu and hands over O in bf16.
//
// One CTA walks 128-token tiles (flat token order, so every global row block is a dense 2-D box).  Warp roles:
//   warps 0-15 "compute"  thread (quarter q = warp%4, lane) owns accumulator row q*32+lane (one TMEM lane) and the
//                         32-column group warp/4: epilogues (bias, residual, LayerNorm, GELU) and the operand tiles of
//                         the next GEMM, written straight into the 128B-swizzled layout.  Thread 0 also drives the
//                         tile's TMA traffic: box loads of the inputs (UTMALDG) and box stores (UTMASTG) of every
//                         saved tensor straight out of the staging / operand tiles — no thread copies a row, and the
//                         stores drain in the background while the next phase computes.  (Measured: the per-thread
//                         LDS+STG copy-out of the first version sustained 27 B/clk per SM, 19 k of a tile's 45 k cycles.)
//   warp 16    "producer" streams the layer's pre-packed bf16 weight blocks (16 KB each, 16 per tile) from L2 through
//                         a 3-slot shared-memory ring with bulk async copies (UBLKCP) and full/empty mbarriers
//   warp 17    "mma"      one thread issues tcgen05.mma (M=128, K=16) from the operand tiles and the ring into TMEM
//                         and commits to the ring's empty barriers / the accumulator-full barriers
// Residuals never sit in shared memory: y is pre-loaded into the FFN2 accumulator with tcgen05.st and the MMA adds
// onto it.  Nothing but the saved tensors goes to HBM (the five-kernel version round-tripped every intermediate).
#include <cuda_bf16.h>

#include "common.cuh"
#include "tc_common.cuh"
#include "tma_host.cuh"

namespace {

constexpr int CT = 128;            // token rows per tile = M of every MMA
constexpr int NCW = 16;            // compute warps
constexpr int NCOMP = NCW * 32;    // compute threads
constexpr int NTHR = NCOMP + 64;   // + producer warp + mma warp
constexpr int CW = 32;             // accumulator columns per compute thread (of a 128-column band)
constexpr int RING = 3;
constexpr int BLK = 16384;         // one [128 rows x 128 bytes] swizzled box: 64 bf16 or 32 fp32 columns

// shared-memory carve-up (from a 1024-byte aligned base); every tile is made of 16 KB swizzled boxes
constexpr int OFF_OP = 0;                          // operand tiles: A [0,32K), A2 [32K,64K); G = all 64 KB
constexpr int OFF_ST = OFF_OP + 4 * BLK;           // bf16 staging tile (2 boxes)
constexpr int OFF_RING = OFF_ST + 2 * BLK;
constexpr int OFF_R = OFF_RING + RING * BLK;       // fp32 tile (4 boxes of 32 columns)
constexpr int OFF_PAR = OFF_R + 4 * BLK;           // biases and LayerNorm parameters (1408 floats)
constexpr int PAR_FLOATS = 1408;
constexpr int OFF_CELL = OFF_PAR + PAR_FLOATS * 4; // position-table row of each tile row
constexpr int OFF_RED = OFF_CELL + CT * 4;         // [2][128][4] partial row statistics
constexpr int OFF_BAR = OFF_RED + 2 * CT * 4 * 4;
constexpr int SMEM_BYTES = OFF_BAR + 256 + 1024;   // + alignment slack
static_assert(SMEM_BYTES <= 232448, "forward chain: shared memory over the 227 KB limit");

enum { FX, FATTN, FXH1, FXH2, FZ, FU16, FG16, FXB, FQKV, F_MAPS };
struct FwdMaps { CUtensorMap m[F_MAPS]; };
struct FwdArgs {
  int n; int mode;
  const uint8_t *Wo, *W1, *W2, *Win;
  const float *bo, *b1, *b2, *bin, *g1, *be1, *g2, *be2; float eps;
  const float* pos; const int32_t* cell_next;
  float *st1, *st2;
  long long* dbg;       // optional [32] clock64 stamps of CTA 0's first tile (tools/chain_phase_timing.py)
};

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tc::smem_u32(bar)) : "memory");
}
// operand tile / TMEM pre-load of this warp complete: one arrival per warp (the barriers expect NCW arrivals)
__device__ __forceinline__ void warp_arrive(uint64_t* bar) {
  tc::fence_async_smem();
  tc::fence_before_sync();
  __syncwarp();
  if ((threadIdx.x & 31) == 0) mbar_arrive(bar);
}
// generic-proxy writes of every compute thread -> visible to the TMA store engine, then one barrier
__device__ __forceinline__ void publish_sync() {
  tc::fence_async_smem();
  asm volatile("bar.sync 1, %0;" ::"n"(NCOMP) : "memory");
}
__device__ __forceinline__ void compute_sync() { asm volatile("bar.sync 1, %0;" ::"n"(NCOMP) : "memory"); }
__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(tc::smem_u32(dst)), "l"(src) : "memory");
}
// completion of all prior cp.async of this thread counted as ONE pending arrival on the mbarrier (count pre-armed at init)
__device__ __forceinline__ void cp_async_arrive(uint64_t* bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(tc::smem_u32(bar)) : "memory");
}
// one 16 KB weight block global -> ring slot with 16-byte cp.async by the 32 lanes of the producer warp.  The weights
// deliberately do NOT go through the bulk-copy / TMA engine: it serves requests in order, and a block queued behind a
// tile's 64-192 KB of output stores stalled the next GEMM for 4-6 k cycles (measured with the phase stamps).
__device__ __forceinline__ void fetch_block(uint8_t* slot, const uint8_t* src, uint64_t* full_bar, int lane) {
#pragma unroll 8
  for (int i = lane; i < BLK / 16; i += 32) cp_async16(slot + i * 16, src + i * 16);
  cp_async_arrive(full_bar);
}
__device__ __forceinline__ void cp_wait_all() { asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;\n" ::: "memory"); }
__device__ __forceinline__ uint32_t pack2(float a, float b) {
  const __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&v);
}
__device__ __forceinline__ uint4 pack8f(const float* f) {
  return make_uint4(pack2(f[0], f[1]), pack2(f[2], f[3]), pack2(f[4], f[5]), pack2(f[6], f[7]));
}
// byte offset of (row, 16-byte chunk c) in a tile of swizzled boxes: bf16 -> chunk = 8 columns, fp32 -> 4 columns
__device__ __forceinline__ uint32_t op_off(int row, int c) { return (uint32_t)(c >> 3) * BLK + tc::swz(row, c & 7); }
// float4 k (0..7) of this thread's 32-column group in the fp32 tile
__device__ __forceinline__ float4* rq(float* sR, int cq, int r, int k) {
  return reinterpret_cast<float4*>(reinterpret_cast<uint8_t*>(sR) + cq * BLK + tc::swz(r, k));
}
__device__ __forceinline__ void get_row(float* sR, int cq, int r, float* v) {
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const float4 t = *rq(sR, cq, r, k);
    v[4 * k] = t.x; v[4 * k + 1] = t.y; v[4 * k + 2] = t.z; v[4 * k + 3] = t.w;
  }
}
__device__ __forceinline__ void put_row(float* sR, int cq, int r, const float* v) {
#pragma unroll
  for (int k = 0; k < 8; ++k) *rq(sR, cq, r, k) = make_float4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
}
// ---- TMA traffic of a tile (issued by one thread)
__device__ __forceinline__ void load_f32(float* sR, const CUtensorMap* map, int row0, uint64_t* bar) {
#pragma unroll
  for (int b = 0; b < 4; ++b) tma::load_2d(reinterpret_cast<uint8_t*>(sR) + b * BLK, map, 32 * b, row0, bar);
}
__device__ __forceinline__ void store_f32(const float* sR, const CUtensorMap* map, int row0) {
#pragma unroll
  for (int b = 0; b < 4; ++b) tma::store_2d(map, 32 * b, row0, reinterpret_cast<const uint8_t*>(sR) + b * BLK);
}
template <int NB>
__device__ __forceinline__ void load_b16(uint8_t* tile, const CUtensorMap* map, int col0, int row0, uint64_t* bar) {
#pragma unroll
  for (int b = 0; b < NB; ++b) tma::load_2d(tile + b * BLK, map, col0 + 64 * b, row0, bar);
}
template <int NB>
__device__ __forceinline__ void store_b16(const uint8_t* tile, const CUtensorMap* map, int col0, int row0) {
#pragma unroll
  for (int b = 0; b < NB; ++b) tma::store_2d(map, col0 + 64 * b, row0, tile + b * BLK);
}
__device__ __forceinline__ float sum32(const float* v) {          // four independent chains
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
  for (int c = 0; c < CW; c += 4) { s0 += v[c]; s1 += v[c + 1]; s2 += v[c + 2]; s3 += v[c + 3]; }
  return (s0 + s1) + (s2 + s3);
}
__device__ __forceinline__ float quad(const float* red, int r) {  // the four column groups' partials of row r
  const float4 p = *reinterpret_cast<const float4*>(red + r * 4);
  return (p.x + p.y) + (p.z + p.w);
}
// exact two-pass LayerNorm statistics of a row from its four 32-column groups (threads of warps w, w+4, w+8, w+12)
__device__ __forceinline__ void row_stats(const float* v, float* red, int r, int cq, float eps, float& mean, float& rstd) {
  red[r * 4 + cq] = sum32(v);
  compute_sync();
  mean = quad(red, r) * (1.0f / 128.f);
  float q0 = 0.f, q1 = 0.f, q2 = 0.f, q3 = 0.f;
#pragma unroll
  for (int c = 0; c < CW; c += 4) {
    const float d0 = v[c] - mean, d1 = v[c + 1] - mean, d2 = v[c + 2] - mean, d3 = v[c + 3] - mean;
    q0 = fmaf(d0, d0, q0); q1 = fmaf(d1, d1, q1); q2 = fmaf(d2, d2, q2); q3 = fmaf(d3, d3, q3);
  }
  red[CT * 4 + r * 4 + cq] = (q0 + q1) + (q2 + q3);
  compute_sync();
  rstd = rsqrtf(quad(red + CT * 4, r) * (1.0f / 128.f) + eps);
}
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

__global__ void __launch_bounds__(NTHR, 1) k_sra_chain_fwd(const __grid_constant__ FwdMaps maps, const FwdArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* sm = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sA = sm + OFF_OP;
  uint8_t* sA2 = sA + 2 * BLK;
  uint8_t* sG = sA;
  uint8_t* sT = sm + OFF_ST;
  uint8_t* sRing = sm + OFF_RING;
  float* sR = reinterpret_cast<float*>(sm + OFF_R);
  float* sPar = reinterpret_cast<float*>(sm + OFF_PAR);
  int32_t* sCell = reinterpret_cast<int32_t*>(sm + OFF_CELL);
  float* sRed = reinterpret_cast<float*>(sm + OFF_RED);
  uint64_t* w_full = reinterpret_cast<uint64_t*>(sm + OFF_BAR);
  uint64_t* w_empty = w_full + RING;
  uint64_t* a_ready = w_empty + RING;      // [4] compute -> mma: operand tile of GEMM g is complete
  uint64_t* acc_full = a_ready + 4;        // [4] mma -> compute: accumulators of GEMM g are complete
  uint64_t* in_full = acc_full + 4;        // TMA -> compute: x (and O) of the tile have landed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(in_full + 1);
  // parameter cache: bo | b1 | b2 | bin | g1 | be1 | g2 | be2
  float* p_bo = sPar; float* p_b1 = sPar + 128; float* p_b2 = sPar + 384; float* p_bin = sPar + 512;
  float* p_g1 = sPar + 896; float* p_be1 = sPar + 1024; float* p_g2 = sPar + 1152; float* p_be2 = sPar + 1280;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_tiles = (a.n + CT - 1) / CT;
  const bool chain = a.mode & 1, next = (a.mode & 2) != 0;
  if (threadIdx.x == 0) {
    for (int i = 0; i < RING; ++i) { tc::mbar_init(&w_full[i], 32); tc::mbar_init(&w_empty[i], 1); }
    for (int i = 0; i < 4; ++i) { tc::mbar_init(&a_ready[i], NCW); tc::mbar_init(&acc_full[i], 1); }
    tc::mbar_init(in_full, 1);
  }
  if (warp == NCW + 1) tc::tmem_alloc(tmem_slot, 512);
  if (threadIdx.x < PAR_FLOATS / 4) {              // parameters (never written inside a step's kernel chain): 16-byte
    const int i = threadIdx.x * 4;                 // cp.async pieces, waited for before the first tile
    const float* src = nullptr;
    if (i < 128) { if (chain) src = a.bo + i; }
    else if (i < 384) { if (chain) src = a.b1 + (i - 128); }
    else if (i < 512) { if (chain) src = a.b2 + (i - 384); }
    else if (i < 896) { if (next) src = a.bin + (i - 512); }
    else if (i < 1024) { if (chain) src = a.g1 + (i - 896); }
    else if (i < 1152) { if (chain) src = a.be1 + (i - 1024); }
    else if (i < 1280) { if (chain) src = a.g2 + (i - 1152); }
    else { if (chain) src = a.be2 + (i - 1280); }
    if (src) cp_async16(sPar + i, src);
  }
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = *tmem_slot;
  gm_pdl_wait();                                   // everything below reads what earlier kernels wrote
  gm_pdl_trigger();

  if (warp == NCW) {
    // ------------------------------------------------------------------ producer: weight blocks through the ring
    {
      uint32_t cnt = 0;
      auto push = [&](const uint8_t* img, int nblk) {
        for (int b = 0; b < nblk; ++b, ++cnt) {
          const uint32_t slot = cnt % RING, ph = (cnt / RING) & 1;
          tc::mbar_wait(&w_empty[slot], ph ^ 1);
          fetch_block(sRing + slot * BLK, img + (size_t)b * BLK, &w_full[slot], lane);
        }
      };
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        if (chain) { push(a.Wo, 2); push(a.W1, 4); push(a.W2, 4); }
        if (next) push(a.Win, 6);
      }
    }
  } else if (warp == NCW + 1) {
    // ------------------------------------------------------------------ mma issuer
    if (lane == 0) {
      const uint32_t idesc = tc::make_idesc_bf16(128, 128, 0, 0);
      uint32_t cnt = 0;
      // one ring block = 64 K-columns = 4 MMAs of K = 16 against the operand block at `a_addr`
      auto mma_block = [&](uint32_t a_addr, uint32_t tcol, bool acc) {
        const uint32_t slot = cnt % RING, ph = (cnt / RING) & 1;
        tc::mbar_wait(&w_full[slot], ph);
        tc::fence_async_smem();                       // the block was written by cp.async (generic proxy)
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
          tc::mbar_wait(&a_ready[2], par);            // gelu(u) in G: s2 acc [0,128) (pre-loaded with y + b2) += g W2^T
          tc::fence_after_sync();
          for (int j = 0; j < 4; ++j) mma_block(A + j * BLK, 0, true);
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
    const int q = warp & 3, cq = warp >> 2;
    const int r = q * 32 + lane;                     // tile row = TMEM lane of this thread
    const int c0 = cq * CW;                          // its 32-column group
    const bool t0 = threadIdx.x == 0;                // drives the tile's TMA loads and stores
    const uint32_t t_lane = tmem + ((uint32_t)(q * 32) << 16);
    // x -> R and O -> A by TMA (rows past the end read as zeros); position rows of the tile for the next in-projection
    auto issue_inputs = [&](int tile) {
      const int row0 = tile * CT;
      if (t0) {
        tc::mbar_expect_tx(in_full, (chain ? 6 : 4) * BLK);
        load_f32(sR, &maps.m[FX], row0, in_full);
        if (chain) load_b16<2>(sA, &maps.m[FATTN], 0, row0, in_full);
      }
      if (next && threadIdx.x < CT) sCell[threadIdx.x] = row0 + (int)threadIdx.x < a.n ? __ldg(a.cell_next + row0 + threadIdx.x) : 0;
    };
    int it = 0;
    int stamp_i = 0;
    auto stamp = [&]() {
      if (a.dbg && blockIdx.x == 0 && threadIdx.x == 0 && stamp_i < 32) a.dbg[stamp_i++] = clock64();
    };
    stamp();                                         // 0: kernel body reached (after griddepcontrol.wait)
    issue_inputs(blockIdx.x);
    cp_wait_all();                                   // parameters
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
      const uint32_t par = it & 1;
      const int row0 = tile * CT;
      const int grow = row0 + r;
      const int ntile = tile + gridDim.x;
      float v[CW];
      tc::mbar_wait(in_full, par);                   // x in R, O in A (written by the async proxy)
      if (chain) warp_arrive(&a_ready[0]);
      compute_sync();                                // parameters / cells of every thread visible
      stamp();                                       // 1: tile inputs on chip
      if (chain) {
        // ---------------- E1: s1 = acc + bo + x ; y = LN1(s1)
        tc::mbar_wait(&acc_full[0], par);
        tc::fence_after_sync();
        stamp();                                     // 2: out-proj accumulators complete
        tc::tmem_ld32(t_lane + c0, v);
        tc::tmem_ld_wait();
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const float4 xr = *rq(sR, cq, r, k);
          const float4 b4 = *reinterpret_cast<const float4*>(p_bo + c0 + 4 * k);
          v[4 * k] += b4.x + xr.x; v[4 * k + 1] += b4.y + xr.y; v[4 * k + 2] += b4.z + xr.z; v[4 * k + 3] += b4.w + xr.w;
        }
        float mean, rstd;
        row_stats(v, sRed, r, cq, a.eps, mean, rstd);          // (its barriers: every thread has read its x row)
        if (cq == 0 && grow < a.n) *reinterpret_cast<float2*>(a.st1 + 2 * (int64_t)grow) = make_float2(mean, rstd);
#pragma unroll
        for (int c = 0; c < CW; ++c) v[c] = (v[c] - mean) * rstd;                                       // xhat1
#pragma unroll
        for (int c = 0; c < CW; c += 8)               // saved for the backward (bf16): staged in the lower half of R
          *reinterpret_cast<uint4*>(reinterpret_cast<uint8_t*>(sR) + op_off(r, (c0 + c) >> 3)) = pack8f(v + c);
#pragma unroll
        for (int c = 0; c < CW; ++c) v[c] = fmaf(v[c], p_g1[c0 + c], p_be1[c0 + c]);                    // y
#pragma unroll
        for (int c = 0; c < CW; c += 8) *reinterpret_cast<uint4*>(sA + op_off(r, (c0 + c) >> 3)) = pack8f(v + c);
#pragma unroll
        for (int c = 0; c < CW; ++c) v[c] += p_b2[c0 + c];
        tmem_st32(t_lane + c0, v);                   // FFN2 accumulator starts at y + b2 (fp32 residual path)
        tmem_st_wait();
        warp_arrive(&a_ready[1]);
        compute_sync();
        if (t0) {
          store_b16<2>(reinterpret_cast<uint8_t*>(sR), &maps.m[FXH1], 0, row0);       // xhat1: LN1 backward + lin1 weight gradient
          tma::store_commit();
        }
        // ---------------- E2: u = acc + b1 ; g = gelu(u)   (two 128-column bands; u staged in T, then in the upper half of R)
        stamp();                                     // 3: E1 done
        tc::mbar_wait(&acc_full[1], par);
        tc::fence_after_sync();
        stamp();                                     // 4: FFN1 accumulators complete
#pragma unroll 1
        for (int band = 0; band < 2; ++band) {
          uint8_t* ust = band == 0 ? sT : reinterpret_cast<uint8_t*>(sR) + 2 * BLK;
          tc::tmem_ld32(t_lane + 128 + band * 128 + c0, v);
          tc::tmem_ld_wait();
#pragma unroll
          for (int c = 0; c < CW; c += 8) {
            float u8[8], g8[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              u8[e] = v[c + e] + p_b1[band * 128 + c0 + c + e];
              g8[e] = u8[e] * gelu_parts(u8[e]).cdf;
            }
            *reinterpret_cast<uint4*>(ust + op_off(r, (c0 + c) >> 3)) = pack8f(u8);
            const int col = band * 128 + c0 + c;     // column of g in [0,256): block col/64, chunk (col%64)/8
            *reinterpret_cast<uint4*>(sG + (uint32_t)(col >> 6) * BLK + tc::swz(r, (col & 63) >> 3)) = pack8f(g8);
          }
        }
        warp_arrive(&a_ready[2]);
        compute_sync();
        if (t0) {
          store_b16<2>(sT, &maps.m[FU16], 0, row0);                                   // saved pre-GELU rows (bf16)
          store_b16<2>(reinterpret_cast<uint8_t*>(sR) + 2 * BLK, &maps.m[FU16], 128, row0);
          store_b16<4>(sG, &maps.m[FG16], 0, row0);                                   // bf16 gelu(u): lin2 weight-gradient operand
          tma::store_commit();
        }
        // ---------------- E3: s2 = acc (= y + b2 + g W2^T) ; z = LN2(s2)
        stamp();                                     // 5: E2 done
        tc::mbar_wait(&acc_full[2], par);
        tc::fence_after_sync();
        stamp();                                     // 6: FFN2 accumulators complete
        tc::tmem_ld32(t_lane + c0, v);
        tc::tmem_ld_wait();
        if (t0) tma::store_wait_read();              // R, T, G have been read by their stores (the barrier is inside row_stats)
        row_stats(v, sRed, r, cq, a.eps, mean, rstd);
        if (cq == 0 && grow < a.n) *reinterpret_cast<float2*>(a.st2 + 2 * (int64_t)grow) = make_float2(mean, rstd);
#pragma unroll
        for (int c = 0; c < CW; ++c) v[c] = (v[c] - mean) * rstd;                                       // xhat2
#pragma unroll
        for (int c = 0; c < CW; c += 8) *reinterpret_cast<uint4*>(sT + op_off(r, (c0 + c) >> 3)) = pack8f(v + c);
#pragma unroll
        for (int c = 0; c < CW; ++c) v[c] = fmaf(v[c], p_g2[c0 + c], p_be2[c0 + c]);                    // z
        put_row(sR, cq, r, v);
        publish_sync();
        if (t0) {
          store_b16<2>(sT, &maps.m[FXH2], 0, row0);  // xhat2: LN2 backward + the next layer's in-projection weight gradient
          store_f32(sR, &maps.m[FZ], row0);          // the layer output (next layer's residual input)
          tma::store_commit();
        }
      }
      stamp();                                       // 7: E3 done
      if (next) {
        // ---------------- operands of the next in-projection: A = bf16(z + pos[cell]), A2 = bf16(z); one warp per row,
        // the eight position rows of a warp fetched before any is used
        const int w = threadIdx.x >> 5;
        float4 p4[CT / NCW];
#pragma unroll
        for (int i = 0; i < CT / NCW; ++i)
          p4[i] = __ldg(reinterpret_cast<const float4*>(a.pos + (int64_t)sCell[w + NCW * i] * 128) + lane);
#pragma unroll
        for (int i = 0; i < CT / NCW; ++i) {
          const int rr = w + NCW * i;
          const float4 z4 = *rq(sR, lane >> 3, rr, lane & 7);
          const uint2 xp = make_uint2(pack2(z4.x + p4[i].x, z4.y + p4[i].y), pack2(z4.z + p4[i].z, z4.w + p4[i].w));
          const uint2 xb = make_uint2(pack2(z4.x, z4.y), pack2(z4.z, z4.w));
          const uint32_t off = op_off(rr, lane >> 1) + (lane & 1) * 8;
          *reinterpret_cast<uint2*>(sA + off) = xp;
          *reinterpret_cast<uint2*>(sA2 + off) = xb;
        }
        warp_arrive(&a_ready[3]);
        compute_sync();
        if (t0 && !chain) {
          store_b16<2>(sA2, &maps.m[FXB], 0, row0);  // bf16 copy of the stack input: operand of layer 0's in-projection weight gradient
          tma::store_commit();
        }
        stamp();                                     // 8: next-layer operands done
        // ---------------- E4: q|k|v of the next layer = acc + bin  (three 128-column bands, bf16, staged in T)
        tc::mbar_wait(&acc_full[3], par);
        tc::fence_after_sync();
        stamp();                                     // 9: in-proj accumulators complete
#pragma unroll 1
        for (int band = 0; band < 3; ++band) {
          tc::tmem_ld32(t_lane + 128 + band * 128 + c0, v);
          tc::tmem_ld_wait();
          if (t0) tma::store_wait_read();            // xhat2 / the previous band has left T
          compute_sync();
#pragma unroll
          for (int c = 0; c < CW; c += 8) {
            float o8[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) o8[e] = v[c + e] + p_bin[band * 128 + c0 + c + e];
            *reinterpret_cast<uint4*>(sT + op_off(r, (c0 + c) >> 3)) = pack8f(o8);
          }
          publish_sync();
          if (t0) {
            store_b16<2>(sT, &maps.m[FQKV], band * 128, row0);
            tma::store_commit();
          }
        }
      }
      if (t0) tma::store_wait_read();                // R, A, A2, T free again (G4 is complete: acc_full[3] was waited)
      if (ntile < n_tiles) issue_inputs(ntile);
      stamp();                                       // 10: tile done
    }
    if (t0) tma::store_wait_all();
    tc::fence_before_sync();
  }
  __syncthreads();
  if (warp == NCW + 1) {
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
//   gradients of linear2 / out_proj are column sums of ds2 / ds1 and ride along there.  Same warp roles and TMA tile
//   traffic as the forward kernel.
constexpr int B_OFF_OP = 0;                             // A | A2 (dqkv' bands 0, 1; ds2 / ds1; u band 1 -> du in place); G = both
constexpr int B_OFF_ST = B_OFF_OP + 4 * BLK;            // dqkv' band 2 | u band 0 | O rows -> dO
constexpr int B_OFF_RING = B_OFF_ST + 2 * BLK;
constexpr int B_OFF_R = B_OFF_RING + RING * BLK;
constexpr int B_OFF_PAR = B_OFF_R + 4 * BLK;            // gamma2 | gamma1
constexpr int B_OFF_RED = B_OFF_PAR + 256 * 4;          // [128][4] float2
constexpr int B_OFF_BAR = B_OFF_RED + CT * 4 * 8;
constexpr int B_SMEM_BYTES = B_OFF_BAR + 256 + 1024;
static_assert(B_SMEM_BYTES <= 232448, "backward chain: shared memory over the 227 KB limit");

enum { BDIN, BDQKV, BXH2, BXH1, BU16, BATTN, BDS2, BDU, BDS1H, BDS1, BDO, BDX, B_MAPS };
struct BwdMaps { CUtensorMap m[B_MAPS]; };
struct BwdArgs {
  int n; int mode;
  const uint8_t *Win_up, *W2, *W1, *Wo;
  const float *st2, *st1, *g2, *g1;
  float* dd;
  float *d_g2, *d_be2, *d_g1, *d_be1;
};

// Column sums over the 32 rows of a warp of 32 per-lane values: after five exchange-and-halve stages lane l holds the sum
// over the warp's rows of column l (62 selects + 31 shuffles + 31 adds, no shared memory, no barrier).
__device__ __forceinline__ float warp_column_sums(float* a) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int k = 16; k > 0; k >>= 1) {
    const bool up = (lane & k) != 0;
#pragma unroll
    for (int i = 0; i < k; ++i) {
      const float send = up ? a[i] : a[i + k];
      const float keep = up ? a[i + k] : a[i];
      a[i] = keep + __shfl_xor_sync(0xffffffffu, send, k);
    }
  }
  return a[0];
}

// LayerNorm backward of one 32-column group of a row: v = dz in, d(pre-LN row) out; xhat comes from the saved bf16 tile.
// acc_dg / acc_db (lane l = column c0 + l) accumulate this warp's share of d_gamma = sum dz xhat and d_beta = sum dz.
__device__ __forceinline__ void ln_backward_row(float* v, const uint8_t* xh_tile, const float* gamma, float rstd, float2* red,
                                                int r, int cq, float& acc_dg, float& acc_db) {
  const int c0 = cq * CW;
  float xh[CW], t[CW];
  float p1a = 0.f, p1b = 0.f, p2a = 0.f, p2b = 0.f;
#pragma unroll
  for (int c = 0; c < CW; c += 8) {
    const uint4 x4 = *reinterpret_cast<const uint4*>(xh_tile + op_off(r, (c0 + c) >> 3));
    const __nv_bfloat162* x2 = reinterpret_cast<const __nv_bfloat162*>(&x4);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float2 xf = __bfloat1622float2(x2[e]);
      xh[c + 2 * e] = xf.x; xh[c + 2 * e + 1] = xf.y;
      const float g0 = v[c + 2 * e] * gamma[c + 2 * e], g1 = v[c + 2 * e + 1] * gamma[c + 2 * e + 1];
      p1a += g0; p1b += g1;
      p2a = fmaf(g0, xf.x, p2a); p2b = fmaf(g1, xf.y, p2b);
    }
  }
  red[r * 4 + cq] = make_float2(p1a + p1b, p2a + p2b);
#pragma unroll
  for (int c = 0; c < CW; ++c) t[c] = v[c] * xh[c];
  acc_dg += warp_column_sums(t);
#pragma unroll
  for (int c = 0; c < CW; ++c) t[c] = v[c];
  acc_db += warp_column_sums(t);
  compute_sync();
  const float4 ra = *reinterpret_cast<const float4*>(red + r * 4), rb = *reinterpret_cast<const float4*>(red + r * 4 + 2);
  const float m1 = ((ra.x + ra.z) + (rb.x + rb.z)) * (1.0f / 128.f), m2 = ((ra.y + ra.w) + (rb.y + rb.w)) * (1.0f / 128.f);
#pragma unroll
  for (int c = 0; c < CW; ++c) v[c] = rstd * (v[c] * gamma[c] - m1 - xh[c] * m2);
}

__global__ void __launch_bounds__(NTHR, 1) k_sra_chain_bwd(const __grid_constant__ BwdMaps maps, const BwdArgs a) {
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
  uint64_t* in_full = acc_full + 4;        // ds1' / dz rows and the dqkv' bands
  uint64_t* r_full = in_full + 1;          // the bf16 xhat2 | xhat1 tiles in R (once per tile)
  uint64_t* u_full = r_full + 1;           // u bands, then O rows (two uses per tile)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(u_full + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_tiles = (a.n + CT - 1) / CT;
  const bool up = a.mode & 1, chain = (a.mode & 2) != 0;
  if (threadIdx.x == 0) {
    for (int i = 0; i < RING; ++i) { tc::mbar_init(&w_full[i], 32); tc::mbar_init(&w_empty[i], 1); }
    for (int i = 0; i < 4; ++i) { tc::mbar_init(&a_ready[i], NCW); tc::mbar_init(&acc_full[i], 1); }
    tc::mbar_init(in_full, 1); tc::mbar_init(r_full, 1); tc::mbar_init(u_full, 1);
  }
  if (warp == NCW + 1) tc::tmem_alloc(tmem_slot, 512);
  if (threadIdx.x < 64 && chain) cp_async16(sPar + threadIdx.x * 4, (threadIdx.x < 32 ? a.g2 : a.g1 - 128) + threadIdx.x * 4);
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = *tmem_slot;
  gm_pdl_wait();
  gm_pdl_trigger();

  if (warp == NCW) {
    {
      uint32_t cnt = 0;
      auto push = [&](const uint8_t* img, int nblk) {
        for (int b = 0; b < nblk; ++b, ++cnt) {
          const uint32_t slot = cnt % RING, ph = (cnt / RING) & 1;
          tc::mbar_wait(&w_empty[slot], ph ^ 1);
          fetch_block(sRing + slot * BLK, img + (size_t)b * BLK, &w_full[slot], lane);
        }
      };
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        if (up) push(a.Win_up, 6);
        if (chain) { push(a.W2, 4); push(a.W1, 4); push(a.Wo, 2); }
      }
    }
  } else if (warp == NCW + 1) {
    if (lane == 0) {
      const uint32_t idesc = tc::make_idesc_bf16(128, 64, 0, 1);     // A K-major, B MN-major, N = 64
      uint32_t cnt = 0;
      // one ring block = W[128 K-rows (out features) x 64 N-cols (in features)]: 8 MMAs of K = 16 against the
      // K = 128 operand band at `a_addr` (two 64-column blocks)
      auto mma_block = [&](uint32_t a_addr, uint32_t tcol, bool acc) {
        const uint32_t slot = cnt % RING, ph = (cnt / RING) & 1;
        tc::mbar_wait(&w_full[slot], ph);
        tc::fence_async_smem();                       // the block was written by cp.async (generic proxy)
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
    const int q = warp & 3, cq = warp >> 2;
    const int r = q * 32 + lane;
    const int c0 = cq * CW;
    const bool t0 = threadIdx.x == 0;
    const uint32_t t_lane = tmem + ((uint32_t)(q * 32) << 16);
    float acc_dg2 = 0.f, acc_db2 = 0.f, acc_dg1 = 0.f, acc_db1 = 0.f;     // lane l: column c0 + l, this warp's 32 rows of every tile
    auto issue_inputs = [&](int tile) {                // ds1' (or dz) -> R, the three dqkv' bands -> A, A2, T
      const int row0 = tile * CT;
      tc::mbar_expect_tx(in_full, (up ? 10 : 4) * BLK);
      load_f32(sR, &maps.m[BDIN], row0, in_full);
      if (up) {
        load_b16<2>(sA, &maps.m[BDQKV], 0, row0, in_full);
        load_b16<2>(sA2, &maps.m[BDQKV], 128, row0, in_full);
        load_b16<2>(sT, &maps.m[BDQKV], 256, row0, in_full);
      }
    };
    int it = 0;
    if (t0) issue_inputs(blockIdx.x);
    cp_wait_all();                                     // LayerNorm weights
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
      const uint32_t par = it & 1;
      const int row0 = tile * CT;
      const int grow = row0 + r;
      const int ntile = tile + gridDim.x;
      float v[CW];
      tc::mbar_wait(in_full, par);                     // R = ds1' (or dz), dqkv' bands on chip
      get_row(sR, cq, r, v);
      if (up) {
        tmem_st32(t_lane + c0, v);                     // accumulator starts at ds1' (fp32 residual-gradient path)
        tmem_st_wait();
        warp_arrive(&a_ready[0]);
      }
      compute_sync();                                  // every thread has read its R row (and sees the parameters)
      if (chain && t0) {                               // bf16 xhat2 -> lower half of R, xhat1 -> upper half
        tc::mbar_expect_tx(r_full, 4 * BLK);
        load_b16<2>(reinterpret_cast<uint8_t*>(sR), &maps.m[BXH2], 0, row0, r_full);
        load_b16<2>(reinterpret_cast<uint8_t*>(sR) + 2 * BLK, &maps.m[BXH1], 0, row0, r_full);
      }
      if (up) {
        tc::mbar_wait(&acc_full[0], par);
        tc::fence_after_sync();
        tc::tmem_ld32(t_lane + c0, v);
        tc::tmem_ld_wait();
      }
      if (!chain) {
        // ---------------- input gradient of the stack: dx = dz
        put_row(sR, cq, r, v);
        publish_sync();
        if (t0) {
          store_f32(sR, &maps.m[BDX], row0);
          tma::store_commit();
          tma::store_wait_read();
          if (ntile < n_tiles) issue_inputs(ntile);
        }
        continue;
      }
      // ---------------- LayerNorm-2 backward.  u band 0 -> T, u band 1 -> A2: du of band 1 is later written over it IN
      // PLACE (same thread, same chunk), so both bands are on chip long before du is needed
      if (t0) {
        tc::mbar_expect_tx(u_full, 4 * BLK);
        load_b16<2>(sT, &maps.m[BU16], 0, row0, u_full);
        load_b16<2>(sA2, &maps.m[BU16], 128, row0, u_full);
      }
      float rstd = grow < a.n ? __ldg(a.st2 + 2 * (int64_t)grow + 1) : 0.f;
      tc::mbar_wait(r_full, par);                      // xhat tiles in R
      ln_backward_row(v, reinterpret_cast<const uint8_t*>(sR), sPar + c0, rstd, sRed, r, cq, acc_dg2, acc_db2);       // v = ds2
#pragma unroll
      for (int c = 0; c < CW; c += 8) *reinterpret_cast<uint4*>(sA + op_off(r, (c0 + c) >> 3)) = pack8f(v + c);
      tmem_st32(t_lane + c0, v);                       // dy accumulator starts at ds2
      tmem_st_wait();
      warp_arrive(&a_ready[1]);
      compute_sync();
      if (t0) {
        store_b16<2>(sA, &maps.m[BDS2], 0, row0);      // bf16 ds2: operand of the lin2 weight gradient
        tma::store_commit();
      }
      // ---------------- du = (ds2 W2) * gelu'(u): two 128-column bands, 32 columns per thread
      tc::mbar_wait(&acc_full[1], par);
      tc::fence_after_sync();
      tc::mbar_wait(u_full, 0);
      if (t0) tma::store_wait_read();                  // A (ds2) has been read: du may be written over it
      compute_sync();
#pragma unroll 1
      for (int band = 0; band < 2; ++band) {
        float w[CW];
        tc::tmem_ld32(t_lane + 128 + band * 128 + c0, w);
        tc::tmem_ld_wait();
#pragma unroll
        for (int c = 0; c < CW; c += 8) {
          const int col = band * 128 + c0 + c;         // column of du in [0,256): block col/64, chunk (col%64)/8
          uint8_t* gdst = sG + (uint32_t)(col >> 6) * BLK + tc::swz(r, (col & 63) >> 3);
          const uint4 u4 = *reinterpret_cast<const uint4*>(band == 0 ? sT + op_off(r, (c0 + c) >> 3) : gdst);
          const __nv_bfloat162* u2 = reinterpret_cast<const __nv_bfloat162*>(&u4);
          float d8[8];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float2 uf = __bfloat1622float2(u2[e]);
            d8[2 * e] = w[c + 2 * e] * gelu_grad_f(uf.x);
            d8[2 * e + 1] = w[c + 2 * e + 1] * gelu_grad_f(uf.y);
          }
          *reinterpret_cast<uint4*>(gdst) = pack8f(d8);
        }
      }
      warp_arrive(&a_ready[2]);
      compute_sync();                                  // du complete, T free
      if (t0) {
        store_b16<4>(sG, &maps.m[BDU], 0, row0);       // bf16 du: operand of the lin1 weight gradient
        tma::store_commit();
        tc::mbar_expect_tx(u_full, 2 * BLK);
        load_b16<2>(sT, &maps.m[BATTN], 0, row0, u_full);
      }
      // ---------------- LayerNorm-1 backward on dy
      tc::mbar_wait(&acc_full[2], par);
      tc::fence_after_sync();
      tc::tmem_ld32(t_lane + c0, v);
      tc::tmem_ld_wait();
      rstd = grow < a.n ? __ldg(a.st1 + 2 * (int64_t)grow + 1) : 0.f;
      if (t0) tma::store_wait_read();                  // G (du) has been read (the barrier is inside ln_backward_row)
      ln_backward_row(v, reinterpret_cast<const uint8_t*>(sR) + 2 * BLK, sPar + 128 + c0, rstd, sRed, r, cq, acc_dg1, acc_db1);   // v = ds1
#pragma unroll
      for (int c = 0; c < CW; c += 8) *reinterpret_cast<uint4*>(sA + op_off(r, (c0 + c) >> 3)) = pack8f(v + c);
      warp_arrive(&a_ready[3]);
      put_row(sR, cq, r, v);                           // (every thread read its xhat values before the barrier above)
      publish_sync();
      if (t0) {
        store_b16<2>(sA, &maps.m[BDS1H], 0, row0);     // bf16 ds1: operand of the out_proj weight gradient
        store_f32(sR, &maps.m[BDS1], row0);            // fp32 ds1: the residual-gradient term of the layer below
        tma::store_commit();
      }
      // ---------------- dO = ds1 Wo ; D = dO . O per head
      tc::mbar_wait(&acc_full[3], par);
      tc::fence_after_sync();
      tc::tmem_ld32(t_lane + 384 + c0, v);
      tc::tmem_ld_wait();
      tc::mbar_wait(u_full, 1);                        // O rows in T
      {
        float dh[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          float d0 = 0.f, d1 = 0.f;
#pragma unroll
          for (int c = 0; c < 16; c += 8) {
            uint8_t* ochunk = sT + op_off(r, (c0 + h * 16 + c) >> 3);
            const uint4 o4 = *reinterpret_cast<const uint4*>(ochunk);
            const __nv_bfloat162* o2 = reinterpret_cast<const __nv_bfloat162*>(&o4);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float2 of = __bfloat1622float2(o2[e]);
              d0 = fmaf(v[h * 16 + c + 2 * e], of.x, d0);
              d1 = fmaf(v[h * 16 + c + 2 * e + 1], of.y, d1);
            }
            *reinterpret_cast<uint4*>(ochunk) = pack8f(v + h * 16 + c);      // dO over O, in place
          }
          dh[h] = d0 + d1;
        }
        if (grow < a.n) *reinterpret_cast<float2*>(a.dd + (int64_t)grow * 8 + cq * 2) = make_float2(dh[0], dh[1]);
      }
      publish_sync();
      if (t0) {
        store_b16<2>(sT, &maps.m[BDO], 0, row0);
        tma::store_commit();
        tma::store_wait_read();                        // R, A, T free again (G3 is complete: acc_full[3] was waited)
        if (ntile < n_tiles) issue_inputs(ntile);
      }
    }
    if (t0) tma::store_wait_all();
    if (chain) {
      atomicAdd(a.d_g2 + c0 + lane, acc_dg2);
      atomicAdd(a.d_be2 + c0 + lane, acc_db2);
      atomicAdd(a.d_g1 + c0 + lane, acc_dg1);
      atomicAdd(a.d_be1 + c0 + lane, acc_db1);
    }
    tc::fence_before_sync();
  }
  __syncthreads();
  if (warp == NCW + 1) {
    tc::fence_after_sync();
    tc::tmem_free(tmem, 512);
  }
}

}  // namespace

static long long* g_chain_dbg = nullptr;
// development aid (not in the public header): the next geomae_sra_chain_fwd launches record phase stamps of CTA 0 there
extern "C" int geomae_debug_chain_stamps(long long* dev_buf) { g_chain_dbg = dev_buf; return GEOMAE_OK; }

#define GM_MAP(dst, ptr, cols, bytes)                                          \
  do {                                                                         \
    if (ptr) {                                                                 \
      const int rc__ = tma::make_map(&(dst), (ptr), n, (cols), (bytes));       \
      if (rc__) return rc__;                                                   \
    }                                                                          \
  } while (0)

extern "C" int geomae_sra_chain_fwd(const geomae_chain_fwd_args* p, void* stream) {
  GM_REQUIRE(p, "sra_chain_fwd: null argument");
  GM_REQUIRE(p->n_tokens >= 0 && p->n_tokens < ((int64_t)1 << 31) - 256, "sra_chain_fwd: bad token count");
  GM_REQUIRE((p->mode & 3) != 0 && (p->mode & ~3) == 0, "sra_chain_fwd: mode must be 1 (chain), 2 (next in-proj) or 3");
  if (p->n_tokens == 0) return GEOMAE_OK;
  const bool chain = p->mode & 1, next = (p->mode & 2) != 0;
  GM_REQUIRE(p->x, "sra_chain_fwd: x is null");
  if (chain)
    GM_REQUIRE(p->attn && p->p_out_proj && p->p_lin1 && p->p_lin2 && p->out_proj_b && p->lin1_b && p->lin2_b &&
                   p->norm1_w && p->norm1_b && p->norm2_w && p->norm2_b && p->xh1_16 && p->st1 && p->xh2_16 && p->st2 && p->z &&
                   p->u16 && p->g16,
               "sra_chain_fwd: the layer chain needs attn, packed weights, biases, norms and every saved-tensor buffer");
  if (next)
    GM_REQUIRE(p->p_in_proj_next && p->in_proj_b_next && p->pos_table && p->tok_cell_next && p->qkv16_next && (chain || p->xb16),
               "sra_chain_fwd: the next in-projection needs packed weights, bias, position table, cells and outputs");
  const int64_t n = p->n_tokens;
  FwdMaps maps;
  GM_MAP(maps.m[FX], p->x, 128, 4);
  if (chain) {
    GM_MAP(maps.m[FATTN], p->attn, 128, 2);
    GM_MAP(maps.m[FXH1], p->xh1_16, 128, 2);
    GM_MAP(maps.m[FXH2], p->xh2_16, 128, 2);
    GM_MAP(maps.m[FZ], p->z, 128, 4);
    GM_MAP(maps.m[FU16], p->u16, 256, 2);
    GM_MAP(maps.m[FG16], p->g16, 256, 2);
  } else {
    GM_MAP(maps.m[FXB], p->xb16, 128, 2);
  }
  if (next) GM_MAP(maps.m[FQKV], p->qkv16_next, 384, 2);
  FwdArgs a;
  a.n = (int)n; a.mode = p->mode;
  a.Wo = (const uint8_t*)p->p_out_proj; a.W1 = (const uint8_t*)p->p_lin1; a.W2 = (const uint8_t*)p->p_lin2;
  a.Win = (const uint8_t*)p->p_in_proj_next;
  a.bo = p->out_proj_b; a.b1 = p->lin1_b; a.b2 = p->lin2_b; a.bin = p->in_proj_b_next;
  a.g1 = p->norm1_w; a.be1 = p->norm1_b; a.g2 = p->norm2_w; a.be2 = p->norm2_b; a.eps = p->ln_eps;
  a.pos = p->pos_table; a.cell_next = p->tok_cell_next;
  a.st1 = p->st1; a.st2 = p->st2;
  a.dbg = g_chain_dbg;
  static bool configured = false;
  if (!configured) {
    GM_CUDA(cudaFuncSetAttribute(k_sra_chain_fwd, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    configured = true;
  }
  const int n_tiles = gm_div_up(a.n, CT);
  const int grid = n_tiles < GM_NUM_SMS ? n_tiles : GM_NUM_SMS;
  GM_CUDA(gm_launch_pdl(k_sra_chain_fwd, dim3(grid), dim3(NTHR), (size_t)SMEM_BYTES, (cudaStream_t)stream, maps, a));
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
    GM_REQUIRE(p->xh2_16 && p->st2 && p->xh1_16 && p->st1 && p->u16 && p->attn16 && p->p_lin2 && p->p_lin1 && p->p_out_proj &&
                   p->norm2_w && p->norm1_w && p->ds2_16 && p->du16 && p->ds1_16 && p->dattn16 && p->ds1 && p->dd &&
                   p->g_norm2_w && p->g_norm2_b && p->g_norm1_w && p->g_norm1_b,
               "sra_chain_bwd: the layer chain needs every saved tensor, packed weight, output and gradient buffer");
  else GM_REQUIRE(p->dx, "sra_chain_bwd: dx is null");
  const int64_t n = p->n_tokens;
  BwdMaps maps;
  GM_MAP(maps.m[BDIN], up ? p->ds1_up : p->dz_in, 128, 4);
  if (up) GM_MAP(maps.m[BDQKV], p->dqkv16_up, 384, 2);
  if (chain) {
    GM_MAP(maps.m[BXH2], p->xh2_16, 128, 2);
    GM_MAP(maps.m[BXH1], p->xh1_16, 128, 2);
    GM_MAP(maps.m[BU16], p->u16, 256, 2);
    GM_MAP(maps.m[BATTN], p->attn16, 128, 2);
    GM_MAP(maps.m[BDS2], p->ds2_16, 128, 2);
    GM_MAP(maps.m[BDU], p->du16, 256, 2);
    GM_MAP(maps.m[BDS1H], p->ds1_16, 128, 2);
    GM_MAP(maps.m[BDS1], p->ds1, 128, 4);
    GM_MAP(maps.m[BDO], p->dattn16, 128, 2);
  } else {
    GM_MAP(maps.m[BDX], p->dx, 128, 4);
  }
  BwdArgs a;
  a.n = (int)n; a.mode = p->mode;
  a.Win_up = (const uint8_t*)p->p_in_proj_up;
  a.W2 = (const uint8_t*)p->p_lin2; a.W1 = (const uint8_t*)p->p_lin1; a.Wo = (const uint8_t*)p->p_out_proj;
  a.st2 = p->st2; a.st1 = p->st1; a.g2 = p->norm2_w; a.g1 = p->norm1_w;
  a.dd = p->dd;
  a.d_g2 = p->g_norm2_w; a.d_be2 = p->g_norm2_b; a.d_g1 = p->g_norm1_w; a.d_be1 = p->g_norm1_b;
  static bool configured = false;
  if (!configured) {
    GM_CUDA(cudaFuncSetAttribute(k_sra_chain_bwd, cudaFuncAttributeMaxDynamicSharedMemorySize, B_SMEM_BYTES));
    configured = true;
  }
  const int n_tiles = gm_div_up(a.n, CT);
  const int grid = n_tiles < GM_NUM_SMS ? n_tiles : GM_NUM_SMS;
  GM_CUDA(gm_launch_pdl(k_sra_chain_bwd, dim3(grid), dim3(NTHR), (size_t)B_SMEM_BYTES, (cudaStream_t)stream, maps, a));
  return GEOMAE_OK;
}
