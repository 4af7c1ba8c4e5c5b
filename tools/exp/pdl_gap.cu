// Hardware experiment: kernel-to-kernel gap inside a CUDA graph with and without programmatic dependent launch (PDL).
// A chain of N dependent kernels (each streams `mb` MB with 148 CTAs, or does nothing) is captured into a graph; the
// per-kernel time of the chain minus the time of one kernel alone is the boundary cost.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 pdl_gap.cu -o pdl_gap.bin
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(256, 1) stream_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst, long long n16, int pdl) {
  extern __shared__ unsigned char smem[];  // forces one CTA per SM like the product kernels
  if (pdl) {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
  }
  for (long long i = blockIdx.x * 256ll + threadIdx.x; i < n16; i += gridDim.x * 256ll) {
    uint4 v = src[i];
    v.x += 1;
    dst[i] = v;
  }
  if (n16 < 0) smem[threadIdx.x] = 1;
}

static float run_chain(int n_kernels, long long n16, bool pdl, uint4* a, uint4* b, int reps) {
  cudaStream_t st;
  cudaStreamCreate(&st);
  cudaGraph_t graph;
  cudaGraphExec_t exec;
  cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal);
  for (int k = 0; k < n_kernels; ++k) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(148);
    cfg.blockDim = dim3(256);
    cfg.dynamicSmemBytes = 160 * 1024;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = pdl ? 1 : 0;
    const uint4* s = (k & 1) ? b : a;
    uint4* d = (k & 1) ? a : b;
    int p = pdl ? 1 : 0;
    cudaError_t e = cudaLaunchKernelEx(&cfg, stream_kernel, s, d, n16, p);
    if (e != cudaSuccess) { printf("launch: %s\n", cudaGetErrorString(e)); exit(1); }
  }
  cudaStreamEndCapture(st, &graph);
  cudaError_t e = cudaGraphInstantiate(&exec, graph, 0);
  if (e != cudaSuccess) { printf("instantiate: %s\n", cudaGetErrorString(e)); exit(1); }
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  cudaGraphLaunch(exec, st);
  cudaStreamSynchronize(st);
  float best = 1e9f;
  for (int r = 0; r < reps; ++r) {
    cudaEventRecord(e0, st);
    cudaGraphLaunch(exec, st);
    cudaEventRecord(e1, st);
    cudaStreamSynchronize(st);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  e = cudaGetLastError();
  if (e != cudaSuccess) { printf("run: %s\n", cudaGetErrorString(e)); exit(1); }
  cudaGraphExecDestroy(exec);
  cudaGraphDestroy(graph);
  cudaStreamDestroy(st);
  return best;
}

int main() {
  cudaFuncSetAttribute(stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
  uint4 *a, *b;
  const long long max_bytes = 256ll << 20;
  cudaMalloc(&a, max_bytes);
  cudaMalloc(&b, max_bytes);
  cudaMemset(a, 0, max_bytes);
  cudaMemset(b, 0, max_bytes);
  for (int mb : {0, 8, 32, 128}) {
    const long long n16 = (static_cast<long long>(mb) << 20) / 16;
    for (int pdl = 0; pdl < 2; ++pdl) {
      const float t1 = run_chain(1, n16, pdl, a, b, 20);
      const float t100 = run_chain(100, n16, pdl, a, b, 10);
      printf("%3d MB per kernel, %s: one kernel %7.2f us, chain of 100: %7.2f us per kernel\n", mb,
             pdl ? "PDL  " : "plain", t1 * 1000.f, t100 * 10.f);
    }
  }
  return 0;
}
