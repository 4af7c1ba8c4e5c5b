// Data-parallel gradient exchange over NVLink / NVSwitch peer memory (declared in include/fpg.h, "Peer exchange").
//
// The reference trains in one process; data-parallel training shards the batch (independent tiles, per-sample
// InstanceNorm) and sums the parameter gradients of the ranks before torch.optim.Adam.step (model.py:633,646). An NCCL
// all-reduce does that sum with kernels: launched beside the backward pass they take SMs from the persistent
// one-CTA-per-SM convolution grids (the displaced CTAs run after the others), launched after it they sit on the
// critical path. Here the exchange is an ALL-GATHER DONE BY THE COPY ENGINES: as soon as a bucket of gradients is
// complete, a rank pushes it into its slot of every peer's staging buffer (cudaMemcpyAsync on IPC-mapped peer memory:
// no SM is involved, NVSwitch gives every pair full bandwidth), and the Adam kernel sums the W gradient sources in rank
// order while it reads them -- every rank forms the same sum in the same order, so the replicas stay bit-identical and
// equal to ONE process accumulating the shards in shard order. Ordering between ranks is carried by two arrays of flag
// words per rank (written by the peers with system-scope release stores, polled locally): "data of step s has landed"
// and "I have consumed step s" (the staging slot may be overwritten). Flag values are step counters read from device
// memory, so every launch has constant arguments and the whole step replays as a CUDA graph.
#include "common.cuh"
#include "host_util.h"

using namespace fpg;

namespace {

constexpr int kMaxPeers = 16;

struct FlagPtrs {
  uint32_t* p[kMaxPeers];
};

struct GradSrcs {
  const float* g[kMaxPeers];
};

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ uint64_t global_timer_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// thread i < n: *flags.p[i] = *ctr + add (release, system scope); afterwards thread 0 advances *ctr by `bump`
__global__ void peer_signal_kernel(FlagPtrs flags, int n, uint32_t* ctr, uint32_t add, uint32_t bump) {
  const uint32_t value = *ctr + add;
  __threadfence_system();
  if (threadIdx.x < n) st_release_sys(flags.p[threadIdx.x], value);
  __syncthreads();
  if (threadIdx.x == 0 && bump) *ctr = *ctr + bump;
}

// lane i < n polls local[i] until it reaches *ctr + add (peers write these words). A peer that never arrives would
// hang the GPU: after timeout_ns the kernel records the lane in status[0] and traps (a loud launch failure, not a hang).
// status[1] accumulates the nanoseconds spent polling (the part of the exchange the backward pass did not hide, plus
// the skew between the ranks): bench.py reports it per step.
__global__ void peer_wait_kernel(const uint32_t* local, int n, int skip, const uint32_t* ctr, uint32_t add,
                                 int64_t* status, uint64_t timeout_ns) {
  const uint32_t target = *ctr + add;
  const int i = threadIdx.x;
  const uint64_t t0 = global_timer_ns();
  if (i < n && i != skip) {
    // counters wrap after 2^32 steps: compare as a signed distance
    while (static_cast<int32_t>(ld_acquire_sys(local + i) - target) < 0) {
      __nanosleep(100);
      if (global_timer_ns() - t0 > timeout_ns) {
        status[0] = i + 1;
        __threadfence_system();
        __trap();
      }
    }
  }
  __syncwarp();
  if (i == 0) status[1] += static_cast<int64_t>(global_timer_ns() - t0);
  __threadfence_system();
}

struct PushDst {
  void* p[kMaxPeers];
};

// Tail of the exchange: what completes when the backward pass is over (nothing left to overlap with) is written to
// the peers by the SMs in one launch -- blockIdx.y = peer, 16-byte loads of the local gradients, 16-byte stores into
// the peer's staging slot over NVLink (posted writes) -- instead of W-1 copy-engine copies of ~14 us fixed cost each
// that the engines run one after the other.
__global__ void __launch_bounds__(256) peer_push_kernel(const uint4* __restrict__ src, PushDst dst, int64_t n16) {
  uint4* out = static_cast<uint4*>(dst.p[blockIdx.y]);
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n16; i += stride) out[i] = src[i];
  __threadfence_system();  // the remote writes are performed before the kernel (and the flag write after it) completes
}

__device__ __forceinline__ void adam_update(float& p, float g, float& m, float& v, float w1, float beta2, float eps,
                                            float step_size, float bc2_sqrt, float grad_scale) {
  adam_update_rn(p, g, m, v, w1, beta2, eps, step_size, bc2_sqrt, grad_scale);  // common.cuh
}

// streaming (evict-first) coherent load: gsum_out may alias the local source
__device__ __forceinline__ float4 ld_stream(const float4* p) { return __ldcs(p); }

// Adam over the gradient sum of n_src sources taken in source order (source r = rank r's gradient: the local buffer
// for this rank, the staged copies for the peers), 16-byte accesses; otherwise adam_dev_kernel of elementwise.cu.
__global__ void adam_multi_kernel(float* __restrict__ p, GradSrcs srcs, int n_src, float* __restrict__ m,
                                  float* __restrict__ v, int64_t count, float beta1, float beta2, float eps,
                                  const int32_t* __restrict__ state, float grad_scale, float* gsum_out) {
  const float step_size = reinterpret_cast<const float*>(state)[2];
  const float bc2_sqrt = reinterpret_cast<const float*>(state)[3];
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  const float w1 = 1.f - beta1;
  const int64_t n4 = count >> 2;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 p4 = reinterpret_cast<float4*>(p)[i], m4 = reinterpret_cast<float4*>(m)[i];
    float4 v4 = reinterpret_cast<float4*>(v)[i];
    float4 g4 = ld_stream(reinterpret_cast<const float4*>(srcs.g[0]) + i);
    for (int s = 1; s < n_src; ++s) {
      const float4 a = ld_stream(reinterpret_cast<const float4*>(srcs.g[s]) + i);
      g4.x += a.x;
      g4.y += a.y;
      g4.z += a.z;
      g4.w += a.w;
    }
    if (gsum_out != nullptr) reinterpret_cast<float4*>(gsum_out)[i] = g4;
    adam_update(p4.x, g4.x, m4.x, v4.x, w1, beta2, eps, step_size, bc2_sqrt, grad_scale);
    adam_update(p4.y, g4.y, m4.y, v4.y, w1, beta2, eps, step_size, bc2_sqrt, grad_scale);
    adam_update(p4.z, g4.z, m4.z, v4.z, w1, beta2, eps, step_size, bc2_sqrt, grad_scale);
    adam_update(p4.w, g4.w, m4.w, v4.w, w1, beta2, eps, step_size, bc2_sqrt, grad_scale);
    reinterpret_cast<float4*>(p)[i] = p4;
    reinterpret_cast<float4*>(m)[i] = m4;
    reinterpret_cast<float4*>(v)[i] = v4;
  }
  for (int64_t i = (n4 << 2) + static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < count; i += stride) {
    float g = srcs.g[0][i];
    for (int s = 1; s < n_src; ++s) g += srcs.g[s][i];
    if (gsum_out != nullptr) gsum_out[i] = g;
    adam_update(p[i], g, m[i], v[i], w1, beta2, eps, step_size, bc2_sqrt, grad_scale);
  }
}

}  // namespace

extern "C" {

int fpg_peer_alloc(void** ptr, int64_t bytes) {
  FPG_REQUIRE(ptr != nullptr && bytes > 0, "bad argument");
  // a dedicated cudaMalloc block: IPC handles address whole allocations, and nothing else may live in an exported one
  FPG_CUDA_CHECK(cudaMalloc(ptr, static_cast<size_t>(bytes)));
  FPG_CUDA_CHECK(cudaMemset(*ptr, 0, static_cast<size_t>(bytes)));
  FPG_CUDA_CHECK(cudaDeviceSynchronize());
  return 0;
}

int fpg_peer_free(void* ptr) {
  if (ptr != nullptr) FPG_CUDA_CHECK(cudaFree(ptr));
  return 0;
}

int fpg_peer_export(void* ptr, void* handle_host) {
  FPG_REQUIRE(ptr != nullptr && handle_host != nullptr, "bad argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == FPG_PEER_HANDLE_BYTES, "handle size");
  FPG_CUDA_CHECK(cudaIpcGetMemHandle(static_cast<cudaIpcMemHandle_t*>(handle_host), ptr));
  return 0;
}

int fpg_peer_open(const void* handle_host, void** ptr) {
  FPG_REQUIRE(ptr != nullptr && handle_host != nullptr, "bad argument");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle_host, sizeof(h));
  FPG_CUDA_CHECK(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return 0;
}

int fpg_peer_close(void* ptr) {
  if (ptr != nullptr) FPG_CUDA_CHECK(cudaIpcCloseMemHandle(ptr));
  return 0;
}

int fpg_peer_copy(void* dst, const void* src, int64_t bytes, void* stream) {
  FPG_REQUIRE(dst != nullptr && src != nullptr && bytes > 0, "bad argument");
  FPG_CUDA_CHECK(cudaMemcpyAsync(dst, src, static_cast<size_t>(bytes), cudaMemcpyDeviceToDevice,
                                 static_cast<cudaStream_t>(stream)));
  return 0;
}

int fpg_peer_signal(void* const* flags_host, int32_t n, uint32_t* ctr, uint32_t add, uint32_t bump, void* stream) {
  FPG_REQUIRE(n >= 0 && n <= kMaxPeers && ctr != nullptr && (n == 0 || flags_host != nullptr), "bad argument");
  FlagPtrs f;
  memset(&f, 0, sizeof(f));
  for (int i = 0; i < n; ++i) {
    FPG_REQUIRE(flags_host[i] != nullptr, "null flag pointer");
    f.p[i] = static_cast<uint32_t*>(flags_host[i]);
  }
  peer_signal_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(f, n, ctr, add, bump);
  FPG_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int fpg_peer_wait(const uint32_t* local_flags, int32_t n, int32_t skip, const uint32_t* ctr, uint32_t add,
                  int64_t* status, float timeout_s, void* stream) {
  FPG_REQUIRE(local_flags != nullptr && n >= 1 && n <= 32 && ctr != nullptr && status != nullptr && timeout_s > 0.f,
              "bad argument");
  peer_wait_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(
      local_flags, n, skip, ctr, add, status, static_cast<uint64_t>(static_cast<double>(timeout_s) * 1e9));
  FPG_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int fpg_peer_push(const void* src, void* const* dst_host, int32_t n_dst, int64_t bytes, void* stream) {
  FPG_REQUIRE(src != nullptr && dst_host != nullptr && n_dst >= 1 && n_dst <= kMaxPeers && bytes > 0 && bytes % 16 == 0 &&
                  (reinterpret_cast<uintptr_t>(src) & 15) == 0,
              "bad argument");
  PushDst d;
  memset(&d, 0, sizeof(d));
  for (int i = 0; i < n_dst; ++i) {
    FPG_REQUIRE(dst_host[i] != nullptr && (reinterpret_cast<uintptr_t>(dst_host[i]) & 15) == 0, "bad destination");
    d.p[i] = dst_host[i];
  }
  const int64_t n16 = bytes / 16;
  int64_t per_peer = (n16 + 256 * 8 - 1) / (256 * 8);  // ~8 stores in flight per thread
  const int sms = sm_count_cached();
  const int64_t cap = sms > 0 ? (2 * sms + n_dst - 1) / n_dst : 32;
  if (per_peer > cap) per_peer = cap;
  if (per_peer < 1) per_peer = 1;
  peer_push_kernel<<<dim3(static_cast<unsigned>(per_peer), static_cast<unsigned>(n_dst)), 256, 0,
                     static_cast<cudaStream_t>(stream)>>>(static_cast<const uint4*>(src), d, n16);
  FPG_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int fpg_adam_step_dev_multi(float* p, const void* const* grads_host, int32_t n_src, float* m, float* v, int64_t count,
                            float beta1, float beta2, float eps, int32_t* state, float grad_scale, float* gsum_out,
                            void* stream) {
  FPG_REQUIRE(p && grads_host && m && v && state && count > 0 && n_src >= 1 && n_src <= kMaxPeers, "bad argument");
  GradSrcs s;
  memset(&s, 0, sizeof(s));
  uintptr_t align = reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(m) | reinterpret_cast<uintptr_t>(v) |
                    reinterpret_cast<uintptr_t>(gsum_out);
  for (int i = 0; i < n_src; ++i) {
    FPG_REQUIRE(grads_host[i] != nullptr, "null gradient source");
    s.g[i] = static_cast<const float*>(grads_host[i]);
    align |= reinterpret_cast<uintptr_t>(grads_host[i]);
  }
  FPG_REQUIRE((align & 15) == 0, "Adam buffers must be 16-byte aligned");
  int rc = fpg_adam_prepare_dev(state, beta1, beta2, stream);
  if (rc) return rc;
  int64_t blocks = (count / 4 + 255) / 256;
  if (blocks > 2368) blocks = 2368;
  if (blocks < 1) blocks = 1;
  adam_multi_kernel<<<static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      p, s, n_src, m, v, count, beta1, beta2, eps, state, grad_scale, gsum_out);
  FPG_CUDA_CHECK(cudaGetLastError());
  return 0;
}

}  // extern "C"
