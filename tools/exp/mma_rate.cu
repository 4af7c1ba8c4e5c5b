// Hardware experiment: cycles per tcgen05.mma (cta_group::1, kind::f16, K = 16, operands in shared memory) as a function
// of M, N and the operand majorness -- the small-N cost decides the tile shapes of the 7x7 head kernels.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -I../../flood-prediction-gan_b200/csrc mma_rate.cu -o mma_rate.bin
#include <stdlib.h>
#include <vector>

#include "common.cuh"

using namespace fpg;

__global__ void __launch_bounds__(128, 1)
rate_kernel(int M, int N, int mn_major, int iters, int a_step, int n_acc, int n_warps, int swz, int b_shift, int b_lbo, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* done = reinterpret_cast<uint64_t*>(smem + 160 * 1024);
  uint32_t* slot = reinterpret_cast<uint32_t*>(done + 4);
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 0; i < 4; ++i) mbar_init(done + i, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if ((threadIdx.x & 31) == 0 && warp < n_warps) {
    const uint32_t idesc = make_idesc_bf16(M, N, mn_major, mn_major);
    const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem + 96 * 1024);
    const uint32_t lt = swizzle_layout_type(swz);
    const uint64_t ad0 = mn_major ? make_smem_desc(a0, 64 * swz, 8 * swz, lt) : make_smem_desc(a0, 0, 8 * swz, lt);
    const uint64_t bd0 = mn_major ? make_smem_desc(b0 + b_shift * swz, b_lbo, 8 * swz, lt) : make_smem_desc(b0 + b_shift * swz, 0, 8 * swz, lt);
    const uint32_t astep16 = a_step >> 4;
    const long long t0 = clock64();
    for (int i = 0; i < iters / n_warps; i += 16) {
#pragma unroll
      for (int u = 0; u < 16; ++u) {
        // n_acc accumulators round-robin (a_step doubles as the A advance in bytes; B advances by k-slices)
        umma_bf16(tmem + warp * 256 + (u % n_acc) * (256 / n_acc >= N ? N : 0), ad0 + u * astep16, bd0 + (swz >= 64 ? 2 * (u & (swz / 32 - 1)) : 0), idesc, 1u);
      }
    }
    umma_commit(done + warp);
    mbar_wait(done + warp, 0);
    const long long t1 = clock64();
    if (blockIdx.x == 0 && warp == 0) out[0] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

int main() {
  long long* dout;
  cudaMalloc(&dout, 8);
  cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 162 * 1024 + 1024);
  const int iters = 8192;
  for (int mn : {1, 0})
    for (int b_lbo : {8192, 9216, 128})
      for (int b_shift : {0, 1, 2, 8}) {
        const int a_step = 0, M = 128, n_acc = 1, swz = 128, n_warps = 1;
        printf("%s M=128 swizzle=128 B start +%d rows, B atom stride %5d :", mn ? "MN-major" : "K-major ", b_shift, b_lbo);
        for (int N : {64, 128, 256}) {
          rate_kernel<<<148, 128, 162 * 1024 + 1024>>>(M, N, mn, iters, a_step, n_acc, n_warps, swz, b_shift, b_lbo, dout);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf(" [%s]\n", cudaGetErrorString(e)); return 1; }
          long long cyc;
          cudaMemcpy(&cyc, dout, 8, cudaMemcpyDeviceToHost);
          printf("  N=%3d %6.1f", N, static_cast<double>(cyc) / iters);
        }
        printf("  clk/MMA\n");
      }
  return 0;
}
