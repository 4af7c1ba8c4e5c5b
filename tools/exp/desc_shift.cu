// Hardware experiment: can a tcgen05.mma shared-memory descriptor start at an arbitrary 128-byte row of a
// TMA-written, 128B-swizzled region (start address not 1024-byte aligned)? This is what a sliding-window implicit GEMM
// needs: one halo patch in shared memory, one descriptor per filter tap pointing at a shifted sub-window.
//   mode 0: A K-major   (rows = pixels = M, 64 channels = K per row)      D[m][n] = sum_k A[m + shift][k] * B[n][k]
//   mode 1: A MN-major  (rows = pixels = K, 64 channels = M per row)      D[m][n] = sum_p X[p + shift][m] * Y[p][n]
// For every shift 0..15 and both settings of the descriptor "base offset" field the result is compared with the host.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -I../../flood-prediction-gan_b200/csrc \
//        desc_shift.cu ../../flood-prediction-gan_b200/csrc/host_util.cu -o desc_shift
#include <math.h>
#include <stdlib.h>
#include <vector>

#include "common.cuh"
#include "host_util.h"

using namespace fpg;

constexpr int ROWS = 160;  // rows of the A / X region in shared memory

__global__ void __launch_bounds__(128, 1)
shift_kernel(const __grid_constant__ CUtensorMap amap, const __grid_constant__ CUtensorMap bmap, int mode, int shift,
             int use_base_offset, int swz_bytes, float* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int row_bytes = swz_bytes;           // channels per row * 2
  uint8_t* sa = smem;                        // ROWS x row_bytes
  uint8_t* sb = smem + 32768;                // 64 x row_bytes (mode 0: N=64 rows K-major; mode 1: 64 pixel rows)
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 49152);
  uint64_t* done = bar + 1;
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 2);
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    mbar_init(done, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(slot, 64);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot;
  const uint32_t layout = swizzle_layout_type(swz_bytes);
  const int cpr = row_bytes / 2;  // channels per row
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(bar, ROWS * row_bytes + 64 * row_bytes);
    tma_load_2d(&amap, bar, sa, 0, 0);
    tma_load_2d(&bmap, bar, sb, 0, 0);
    mbar_wait(bar, 0);
    tc_fence_after();
    const uint32_t a0 = smem_u32(sa) + shift * row_bytes;
    const uint32_t b0 = smem_u32(sb);
    auto desc = [&](uint32_t addr, uint32_t lbo, uint32_t sbo) {
      uint64_t d = make_smem_desc(addr, lbo, sbo, layout);
      if (use_base_offset) d |= static_cast<uint64_t>((addr >> 7) & 7) << 49;
      return d;
    };
    if (mode == 0) {
      // M = 128 pixel rows, N = 64, K = cpr channels in steps of 16 (32 bytes)
      const uint32_t idesc = make_idesc_bf16(128, 64, 0, 0);
      for (int k = 0; k < cpr / 16; ++k)
        umma_bf16(tmem, desc(a0 + k * 32, 0, 8 * row_bytes), desc(b0 + k * 32, 0, 8 * row_bytes), idesc, k != 0);
    } else {
      // M = cpr channels of X (MN-major), N = cpr channels of Y, K = 64 pixel rows in steps of 16 rows
      const uint32_t idesc = make_idesc_bf16(cpr == 64 ? 64 : 64, cpr, 1, 1);
      for (int k = 0; k < 4; ++k)
        umma_bf16(tmem, desc(a0 + k * 16 * row_bytes, 64 * row_bytes, 8 * row_bytes),
                  desc(b0 + k * 16 * row_bytes, 64 * row_bytes, 8 * row_bytes), idesc, k != 0);
    }
    umma_commit(done);
  }
  __syncthreads();
  mbar_wait(done, 0);
  tc_fence_after();
  // D: mode 0 -> 128 lanes x 64 cols; mode 1 (M=64) -> lanes 0..15 of each quarter hold rows 16q..16q+15
  uint32_t v[16];
  const int ncol = mode == 0 ? 64 : cpr;
  for (int c = 0; c < ncol; c += 16) {
    tmem_ld16(tmem + (static_cast<uint32_t>(warp * 32) << 16) + c, v);
    tmem_ld_wait();
    for (int i = 0; i < 16; ++i) out[(warp * 32 + (threadIdx.x & 31)) * 64 + c + i] = __uint_as_float(v[i]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 64);
}

static float bf(float x) { return __bfloat162float(__float2bfloat16(x)); }

int main() {
  for (int swz : {128, 64, 32}) {
    const int cpr = swz / 2;
    std::vector<__nv_bfloat16> ha(ROWS * cpr), hb(64 * cpr);
    std::vector<float> fa(ROWS * cpr), fb(64 * cpr);
    srand(1);
    for (size_t i = 0; i < ha.size(); ++i) { fa[i] = bf((rand() % 17 - 8) / 8.f); ha[i] = __float2bfloat16(fa[i]); }
    for (size_t i = 0; i < hb.size(); ++i) { fb[i] = bf((rand() % 13 - 6) / 4.f); hb[i] = __float2bfloat16(fb[i]); }
    __nv_bfloat16 *da, *db;
    float* dout;
    cudaMalloc(&da, ha.size() * 2);
    cudaMalloc(&db, hb.size() * 2);
    cudaMalloc(&dout, 128 * 64 * 4);
    cudaMemcpy(da, ha.data(), ha.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(db, hb.data(), hb.size() * 2, cudaMemcpyHostToDevice);
    fpg_tmap ta = {}, tb = {};
    ta.base = da; ta.rank = 2; ta.swizzle_bytes = swz; ta.dims[0] = cpr; ta.dims[1] = ROWS; ta.strides[0] = cpr * 2;
    ta.box[0] = cpr; ta.box[1] = ROWS;
    tb = ta; tb.base = db; tb.dims[1] = 64; tb.box[1] = 64;
    CUtensorMap ma, mb;
    if (encode_tmap(&ta, &ma) || encode_tmap(&tb, &mb)) { printf("encode failed: %s\n", fpg_last_error()); return 1; }
    cudaFuncSetAttribute(shift_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 52 * 1024);
    for (int mode = 0; mode < 2; ++mode) {
      if (mode == 1 && swz != 128) continue;  // MN-major M = 64 needs a 64-channel atom here
      for (int ubo = 0; ubo < 2; ++ubo) {
        printf("swizzle %3d mode %d base_offset %d :", swz, mode, ubo);
        for (int shift = 0; shift < 16; ++shift) {
          cudaMemset(dout, 0, 128 * 64 * 4);
          shift_kernel<<<1, 128, 52 * 1024>>>(ma, mb, mode, shift, ubo, swz, dout);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf(" [%s]", cudaGetErrorString(e)); return 2; }
          std::vector<float> ho(128 * 64);
          cudaMemcpy(ho.data(), dout, ho.size() * 4, cudaMemcpyDeviceToHost);
          double maxerr = 0;
          if (mode == 0) {
            for (int m = 0; m < 128; ++m)
              for (int n = 0; n < 64; ++n) {
                double r = 0;
                for (int k = 0; k < cpr; ++k) r += fa[(m + shift) * cpr + k] * fb[n * cpr + k];
                maxerr = fmax(maxerr, fabs(r - ho[m * 64 + n]));
              }
          } else {
            for (int m = 0; m < 64; ++m)
              for (int n = 0; n < 64; ++n) {
                double r = 0;
                for (int p = 0; p < 64; ++p) r += fa[(p + shift) * cpr + m] * fb[p * cpr + n];
                const int lane_row = (m / 16) * 32 + (m % 16);
                maxerr = fmax(maxerr, fabs(r - ho[lane_row * 64 + n]));
              }
          }
          printf(" %s", maxerr < 1e-3 ? "ok" : "XX");
        }
        printf("\n");
      }
    }
    cudaFree(da); cudaFree(db); cudaFree(dout);
  }
  return 0;
}
