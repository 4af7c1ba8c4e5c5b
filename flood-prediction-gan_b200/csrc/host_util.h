// Host-side helpers shared by the launchers: error reporting and TMA tensor-map encoding.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>

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
// floor division / positive modulo for possibly negative numerators
inline int floor_div(int a, int b) { return (a >= 0) ? a / b : -((-a + b - 1) / b); }
inline int pos_mod(int a, int b) { return a - floor_div(a, b) * b; }

}  // namespace fpg
