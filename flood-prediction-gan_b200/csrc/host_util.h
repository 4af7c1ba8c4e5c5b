// Host-side helpers shared by the launchers: error reporting and TMA tensor-map encoding.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>

#include "../../include/fpg.h"

namespace fpg {

void set_error(const char* fmt, ...);
int fail(int code, const char* fmt, ...);

#define FPG_CUDA_CHECK(expr)                                                                      \
  do {                                                                                            \
    cudaError_t _e = (expr);                                                                      \
    if (_e != cudaSuccess)                                                                        \
      return fpg::fail(static_cast<int>(_e), "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                       __FILE__, __LINE__);                                                       \
  } while (0)

#define FPG_REQUIRE(cond, ...)                              \
  do {                                                      \
    if (!(cond)) return fpg::fail(FPG_EINVAL, __VA_ARGS__); \
  } while (0)

// Encode a bf16 tiled tensor map from the ABI description. Returns 0 or an error code.
int encode_tmap(const fpg_tmap* t, CUtensorMap* out);

int sm_count_cached();

inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// Launch of a persistent (one CTA per SM) kernel. FPG_CLUSTER_ALL=1 (experiment, tools/exp_sm_split.py): every such
// kernel is launched as clusters of 2 CTAs so that the block scheduler hands out SMs TPC by TPC -- two SM-capped grids
// on two streams then partition the chip into whole TPCs and the CTA-pair kernels of one stream still find free pairs.
template <typename... KArgs, typename... Args>
inline cudaError_t launch_persistent(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                     Args&&... args) {
  static const bool pair_all = getenv("FPG_CLUSTER_ALL") != nullptr;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  if (pair_all && (grid.x % 2 == 0 || grid.y % 2 == 0)) {
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = grid.x % 2 == 0 ? 2 : 1;
    attr[0].val.clusterDim.y = grid.x % 2 == 0 ? 1 : 2;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
  }
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
// floor division / positive modulo for possibly negative numerators
inline int floor_div(int a, int b) { return (a >= 0) ? a / b : -((-a + b - 1) / b); }
inline int pos_mod(int a, int b) { return a - floor_div(a, b) * b; }

}  // namespace fpg
