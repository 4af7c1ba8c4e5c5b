#include <stdlib.h>

#include "host_util.h"

#include <string.h>

namespace fpg {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code == 0 ? FPG_EINVAL : code;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
  if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || p == nullptr) {
    set_error("cuTensorMapEncodeTiled not available: %s", cudaGetErrorString(e));
    (void)cudaGetLastError();
    return nullptr;
  }
  fn = reinterpret_cast<EncodeTiledFn>(p);
  return fn;
}

int encode_tmap(const fpg_tmap* t, CUtensorMap* out) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return FPG_ENOTSUP;
  if (t->rank < 1 || t->rank > 5) return fail(FPG_EINVAL, "tensor map rank %d", t->rank);
  CUtensorMapSwizzle sw;
  switch (t->swizzle_bytes) {
    case 128: sw = CU_TENSOR_MAP_SWIZZLE_128B; break;
    case 64: sw = CU_TENSOR_MAP_SWIZZLE_64B; break;
    case 32: sw = CU_TENSOR_MAP_SWIZZLE_32B; break;
    default: return fail(FPG_EINVAL, "unsupported swizzle %d", t->swizzle_bytes);
  }
  cuuint64_t dims[5];
  cuuint64_t strides[4];
  cuuint32_t box[5];
  cuuint32_t estr[5];
  for (int i = 0; i < t->rank; ++i) {
    dims[i] = t->dims[i];
    box[i] = t->box[i];
    estr[i] = 1;
    if (i > 0) strides[i - 1] = t->strides[i - 1];
  }
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, static_cast<cuuint32_t>(t->rank), t->base, dims, strides, box,
                  estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    return fail(static_cast<int>(r),
                "cuTensorMapEncodeTiled failed (%d): rank %d base %p dims [%llu %llu %llu %llu %llu] strides [%llu "
                "%llu %llu %llu] box [%u %u %u %u %u] swizzle %d",
                static_cast<int>(r), t->rank, t->base, (unsigned long long)t->dims[0], (unsigned long long)t->dims[1],
                (unsigned long long)t->dims[2], (unsigned long long)t->dims[3], (unsigned long long)t->dims[4],
                (unsigned long long)t->strides[0], (unsigned long long)t->strides[1],
                (unsigned long long)t->strides[2], (unsigned long long)t->strides[3], t->box[0], t->box[1], t->box[2],
                t->box[3], t->box[4], t->swizzle_bytes);
  }
  return 0;
}

int sm_count_cached() {
  static int cached = 0;
  if (cached > 0) return cached;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) {
    (void)cudaGetLastError();
    return -1;
  }
  int n = 0;
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) {
    (void)cudaGetLastError();
    return -1;
  }
  // FPG_SM_BUDGET=<n>: size every grid for n SMs (experiments with several concurrent streams, tools/exp_sm_split.py)
  if (const char* budget = getenv("FPG_SM_BUDGET")) {
    const int b = atoi(budget);
    if (b > 0 && b < n) n = b;
  }
  cached = n;
  return n;
}

}  // namespace fpg

extern "C" {
int fpg_abi_version(void) { return FPG_ABI_VERSION; }
const char* fpg_last_error(void) { return fpg::g_err; }
int fpg_sm_count(void) { return fpg::sm_count_cached(); }
}
