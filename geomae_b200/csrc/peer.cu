// Small latency-bound all-reduce over NVLink / NVSwitch peer memory: the per-channel BatchNorm
// statistics of the VFE (2C doubles forward, 2C doubles backward per layer) that `naiveSyncBN1d` exchanges between
// ranks (mmdet3d/ops/norm.py:66-73) four times per training step.  Through a collective library each of them costs a
// host-side launch of ~50 us and a multi-kernel protocol for 1-2 KB; here ONE single-CTA kernel scales the local
// vector, stores it into every rank's mailbox (peer stores), publishes a release flag, waits for the other ranks'
// flags in its own mailbox and sums — the exchange is part of the same kernel that normalises the statistics.
//
// Mailbox of a rank (device memory of that rank, IPC-mapped into every process): [2 parities][world ranks] slots of
// (1 flag + PEER_MAX doubles).  Epochs increase by one per call on every rank; parity = epoch & 1.  A slot written for
// epoch e is next written for e + 2, which its writer can only reach after the reader has published e + 1, i.e. after
// the reader's epoch-e kernel (which read the slot) has finished: two parities are enough.
#include <string.h>

#include "common.cuh"

namespace {

constexpr int PEER_MAX = 512;                 // doubles per message
constexpr int PEER_SLOT = PEER_MAX + 1;       // flag + payload, in 8-byte words

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

struct PeerArgs {
  double* box[8];
  int rank, world, count;
  double pre, post;
  unsigned long long epoch;
  int32_t* timeout_flag;
};

__global__ void __launch_bounds__(256) k_peer_allreduce(PeerArgs a, double* buf) {
  const int par = (int)(a.epoch & 1ull);
  if (a.world > 1) {
    for (int dst = 0; dst < a.world; ++dst) {
      double* slot = a.box[dst] + (int64_t)(par * a.world + a.rank) * PEER_SLOT;
      for (int i = threadIdx.x; i < a.count; i += blockDim.x) slot[1 + i] = buf[i] * a.pre;
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x < a.world) {
      double* slot = a.box[threadIdx.x] + (int64_t)(par * a.world + a.rank) * PEER_SLOT;
      st_release_sys(reinterpret_cast<unsigned long long*>(slot), a.epoch);
      const unsigned long long* mine = reinterpret_cast<const unsigned long long*>(
          a.box[a.rank] + (int64_t)(par * a.world + threadIdx.x) * PEER_SLOT);
      const long long t0 = clock64();
      while (ld_acquire_sys(mine) != a.epoch) {
        if (clock64() - t0 > 20000000000ll) {       // ~10 s: a peer died; do not hang the device
          if (a.timeout_flag) atomicExch(a.timeout_flag, 1);
          break;
        }
        __nanosleep(100);
      }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < a.count; i += blockDim.x) {
      double s = 0.0;
      for (int r = 0; r < a.world; ++r)
        s += *reinterpret_cast<volatile double*>(a.box[a.rank] + (int64_t)(par * a.world + r) * PEER_SLOT + 1 + i);
      buf[i] = s * a.post;
    }
  } else {
    for (int i = threadIdx.x; i < a.count; i += blockDim.x) buf[i] = buf[i] * a.pre * a.post;
  }
}

// Gradient exchange as reduce-scatter + all-gather in ONE kernel over peer memory: this rank owns the slice
// [lo + (hi-lo)*rank/world, lo + (hi-lo)*(rank+1)/world) of every rank's gradient buffer, reads the slice from all
// ranks (peer loads), sums in rank order and stores the sum back into all ranks' buffers (peer stores).  Every element
// is summed by exactly one rank, so all ranks end up with bit-identical gradients.  The caller brackets it with two
// geomae_peer_allreduce_f64 calls acting as barriers ("every rank's gradients are complete" before, "every slice has
// been written everywhere" after).
struct ShardArgs {
  float* g[8];
  int world;
  int64_t begin4, end4;     // my slice in float4 units
};

__global__ void __launch_bounds__(256) k_peer_reduce_shard(ShardArgs a) {
  for (int64_t i = a.begin4 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.end4;
       i += (int64_t)gridDim.x * blockDim.x) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int r = 0; r < a.world; ++r) {
      const float4 v = __ldcg(reinterpret_cast<const float4*>(a.g[r]) + i);      // L2 / peer, never a stale L1 line
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    for (int r = 0; r < a.world; ++r) reinterpret_cast<float4*>(a.g[r])[i] = acc;
  }
}

}  // namespace

extern "C" int64_t geomae_peer_mailbox_doubles(int32_t world) { return (int64_t)2 * world * PEER_SLOT; }

extern "C" int geomae_peer_buffer_create(int64_t bytes_, void** buffer, void* ipc_handle_64) {
  GM_REQUIRE(bytes_ > 0 && buffer && ipc_handle_64, "peer_buffer_create: bad argument");
  const size_t bytes = (size_t)bytes_;
  void* p = nullptr;
  GM_CUDA(cudaMalloc(&p, bytes));
  GM_CUDA(cudaMemset(p, 0, bytes));
  GM_CUDA(cudaDeviceSynchronize());
  cudaIpcMemHandle_t h;
  GM_CUDA(cudaIpcGetMemHandle(&h, p));
  memcpy(ipc_handle_64, &h, 64);
  *buffer = p;
  return GEOMAE_OK;
}

extern "C" int geomae_peer_mailbox_create(int32_t world, void** mailbox, void* ipc_handle_64) {
  GM_REQUIRE(world >= 1 && world <= 8 && mailbox && ipc_handle_64, "peer_mailbox_create: bad argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  const size_t bytes = (size_t)geomae_peer_mailbox_doubles(world) * sizeof(double);
  void* p = nullptr;
  GM_CUDA(cudaMalloc(&p, bytes));
  GM_CUDA(cudaMemset(p, 0, bytes));
  GM_CUDA(cudaDeviceSynchronize());
  cudaIpcMemHandle_t h;
  GM_CUDA(cudaIpcGetMemHandle(&h, p));
  memcpy(ipc_handle_64, &h, 64);
  *mailbox = p;
  return GEOMAE_OK;
}

// Map a peer's mailbox into this process FROM THE CURRENT DEVICE: cudaIpcMemLazyEnablePeerAccess sets up peer access
// between the current device and the owner, which is what lets this device's kernels dereference the mapping.
extern "C" int geomae_peer_mailbox_open(const void* ipc_handle_64, void** mapped) {
  GM_REQUIRE(ipc_handle_64 && mapped, "peer_mailbox_open: null argument");
  cudaIpcMemHandle_t h;
  memcpy(&h, ipc_handle_64, 64);
  GM_CUDA(cudaIpcOpenMemHandle(mapped, h, cudaIpcMemLazyEnablePeerAccess));
  return GEOMAE_OK;
}

extern "C" int geomae_peer_mailbox_close(void* ptr, int32_t owned) {
  if (!ptr) return GEOMAE_OK;
  if (owned) GM_CUDA(cudaFree(ptr));
  else GM_CUDA(cudaIpcCloseMemHandle(ptr));
  return GEOMAE_OK;
}

extern "C" int geomae_peer_enable_access(int32_t peer_device) {
  int cur = -1, can = 0;
  GM_CUDA(cudaGetDevice(&cur));
  if (cur == peer_device) return GEOMAE_OK;
  GM_CUDA(cudaDeviceCanAccessPeer(&can, cur, peer_device));
  GM_REQUIRE(can, "peer_enable_access: device %d cannot access device %d", cur, peer_device);
  const cudaError_t e = cudaDeviceEnablePeerAccess(peer_device, 0);
  if (e == cudaErrorPeerAccessAlreadyEnabled) {
    (void)cudaGetLastError();
    return GEOMAE_OK;
  }
  GM_CUDA(e);
  return GEOMAE_OK;
}

extern "C" int geomae_peer_allreduce_f64(const geomae_peer_ctx* ctx, double* buf, int32_t count, double pre_scale,
                                         double post_scale, uint64_t epoch, void* stream) {
  GM_REQUIRE(ctx && buf, "peer_allreduce: null argument");
  GM_REQUIRE(count >= 0 && count <= PEER_MAX, "peer_allreduce: %d doubles, at most %d", count, PEER_MAX);
  GM_REQUIRE(ctx->world >= 1 && ctx->world <= 8 && ctx->rank >= 0 && ctx->rank < ctx->world,
             "peer_allreduce: rank %d of %d (1..8 ranks of one node)", ctx->rank, ctx->world);
  GM_REQUIRE(ctx->world == 1 || epoch > 0, "peer_allreduce: epochs start at 1");
  if (count == 0) return GEOMAE_OK;
  PeerArgs a;
  for (int r = 0; r < 8; ++r) a.box[r] = r < ctx->world ? (double*)ctx->mailbox[r] : nullptr;
  for (int r = 0; r < ctx->world && ctx->world > 1; ++r) GM_REQUIRE(a.box[r], "peer_allreduce: mailbox of rank %d missing", r);
  a.rank = ctx->rank; a.world = ctx->world; a.count = count; a.pre = pre_scale; a.post = post_scale; a.epoch = epoch;
  a.timeout_flag = (int32_t*)ctx->timeout_flag;
  k_peer_allreduce<<<1, 256, 0, (cudaStream_t)stream>>>(a, buf);
  GM_LAUNCH_CHECK();
  return GEOMAE_OK;
}

extern "C" int geomae_peer_reduce_shard(const geomae_peer_ctx* ctx, void* const* grads, int64_t lo, int64_t hi,
                                        void* stream) {
  GM_REQUIRE(ctx && grads, "peer_reduce_shard: null argument");
  GM_REQUIRE(ctx->world >= 2 && ctx->world <= 8 && ctx->rank >= 0 && ctx->rank < ctx->world, "peer_reduce_shard: bad ranks");
  GM_REQUIRE(lo >= 0 && hi >= lo && lo % 4 == 0 && hi % 4 == 0, "peer_reduce_shard: range [%lld, %lld) must be float4-aligned",
             (long long)lo, (long long)hi);
  ShardArgs a;
  for (int r = 0; r < 8; ++r) a.g[r] = r < ctx->world ? (float*)grads[r] : nullptr;
  for (int r = 0; r < ctx->world; ++r) GM_REQUIRE(a.g[r], "peer_reduce_shard: buffer of rank %d missing", r);
  a.world = ctx->world;
  const int64_t n4 = (hi - lo) / 4;
  a.begin4 = lo / 4 + n4 * ctx->rank / ctx->world;
  a.end4 = lo / 4 + n4 * (ctx->rank + 1) / ctx->world;
  if (a.end4 <= a.begin4) return GEOMAE_OK;
  const int blocks = (int)min((int64_t)GM_NUM_SMS * 4, (a.end4 - a.begin4 + 255) / 256);
  k_peer_reduce_shard<<<blocks, 256, 0, (cudaStream_t)stream>>>(a);
  GM_LAUNCH_CHECK();
  return GEOMAE_OK;
}
