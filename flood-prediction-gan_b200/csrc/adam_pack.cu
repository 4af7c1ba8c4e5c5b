// Adam and the bf16 operand repack in ONE pass over the parameters (declared in include/fpg.h, "Optimiser").
//
// torch.optim.Adam.step (model.py:633,646) is followed, in this implementation, by the rebuild of every packed bf16 GEMM
// operand from the updated fp32 master (fprop layout [K][taps][C], data-gradient layout [C][class taps][K] per output-
// parity class). As two launches (adam_dev_kernel + pack_batched_kernel) the master was written once and read back
// twice (once per layout): 28 + 12 bytes per parameter and a transposing gather from global memory. Here a block owns a
// tile of TK x TC (output channel, input channel) pairs with all their taps: it walks the tile as contiguous runs of
// the parameter, applies the update (p, m, v read and written once, 16 independent loads in flight per thread), keeps
// the new values in shared memory, and writes every operand row of every layout as contiguous runs from there.
// Parameters that are not convolution weights (biases, normalisation affine terms) are updated by "chunk" blocks of the
// same launch, which also refresh the zero-padded fp32 bias vectors the epilogues read. The gradient may be a sum of
// several sources taken in a fixed order (data-parallel peer exchange, csrc/peer.cu).
#include "common.cuh"
#include "host_util.h"

using namespace fpg;

namespace {

constexpr int kMaxSrc = 16;
constexpr int kTileFloats = 9472;   // 37 KB of static shared memory for the tile
constexpr int kUnroll = 4;

struct GradSrcs {
  const float* g[kMaxSrc];
};

struct Hyper {
  float w1, beta2, eps, grad_scale;
};

__device__ __forceinline__ void adam_update(float& p, float g, float& m, float& v, const Hyper& h, float step_size,
                                            float bc2_sqrt) {
  adam_update_rn(p, g, m, v, h.w1, h.beta2, h.eps, step_size, bc2_sqrt, h.grad_scale);  // common.cuh
}

__global__ void __launch_bounds__(256)
adam_pack_kernel(float* __restrict__ p, GradSrcs srcs, int n_src, float* __restrict__ m, float* __restrict__ v,
                 const int32_t* __restrict__ state, Hyper hy, const fpg_adam_pack_layer* __restrict__ layers,
                 const fpg_pack_job* __restrict__ jobs, const fpg_adam_pack_chunk* __restrict__ chunks,
                 const int32_t* __restrict__ block_item, const int32_t* __restrict__ block_first, float* gsum_out) {
  __shared__ float tile[kTileFloats];
  __shared__ fpg_adam_pack_layer L;
  __shared__ fpg_pack_job job;
  const float step_size = reinterpret_cast<const float*>(state)[2];
  const float bc2_sqrt = reinterpret_cast<const float*>(state)[3];
  const int item = block_item[blockIdx.x];
  if (item < 0) {
    // ---- chunk: plain parameters [off, off + count), optionally mirrored into a padded fp32 vector
    const fpg_adam_pack_chunk ch = chunks[-1 - item];
    for (int i = threadIdx.x; i < ch.count; i += blockDim.x) {
      const int64_t idx = ch.off + i;
      float g = srcs.g[0][idx];
      for (int s = 1; s < n_src; ++s) g += srcs.g[s][idx];
      if (gsum_out != nullptr) gsum_out[idx] = g;
      float pi = p[idx], mi = m[idx], vi = v[idx];
      adam_update(pi, g, mi, vi, hy, step_size, bc2_sqrt);
      p[idx] = pi;
      m[idx] = mi;
      v[idx] = vi;
      if (ch.copy_dst != nullptr) ch.copy_dst[i] = pi;
    }
    return;
  }
  {
    const int32_t* src = reinterpret_cast<const int32_t*>(layers + item);
    int32_t* dst = reinterpret_cast<int32_t*>(&L);
    for (int i = threadIdx.x; i < static_cast<int>(sizeof(L) / 4); i += blockDim.x) dst[i] = src[i];
  }
  __syncthreads();
  const int rs = L.rs, rsp = rs | 1;
  const int tiles_c = (L.c + L.tc - 1) / L.tc;
  const int tile_idx = block_first[blockIdx.x];
  const int k0 = (tile_idx / tiles_c) * L.tk, c0 = (tile_idx % tiles_c) * L.tc;
  const int nk = min(L.tk, L.k - k0), nc = min(L.tc, L.c - c0);
  const int pitch_k = (L.tc * rsp) | 1;  // odd: lanes that walk k (data-gradient layouts) hit distinct banks
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // ---- phase 1: the update over contiguous runs of the parameter (one run per output channel of the tile)
  const int run = nc * rs;
  const float inv_rs = 1.f / static_cast<float>(rs);
  for (int kk = warp; kk < nk; kk += 8) {
    const int64_t base = L.p_off + (static_cast<int64_t>(k0 + kk) * L.c + c0) * rs;
    float* trow = tile + kk * pitch_k;
    for (int e0 = lane; e0 < run; e0 += 32 * kUnroll) {
      float pv[kUnroll], gv[kUnroll], mv[kUnroll], vv[kUnroll];
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) {
        const int e = e0 + 32 * u;
        if (e < run) {
          pv[u] = p[base + e];
          mv[u] = m[base + e];
          vv[u] = v[base + e];
          gv[u] = srcs.g[0][base + e];
        }
      }
      for (int s = 1; s < n_src; ++s) {
#pragma unroll
        for (int u = 0; u < kUnroll; ++u)
          if (e0 + 32 * u < run) gv[u] += srcs.g[s][base + e0 + 32 * u];
      }
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) {
        const int e = e0 + 32 * u;
        if (e < run) {
          if (gsum_out != nullptr) gsum_out[base + e] = gv[u];
          adam_update(pv[u], gv[u], mv[u], vv[u], hy, step_size, bc2_sqrt);
          p[base + e] = pv[u];
          m[base + e] = mv[u];
          v[base + e] = vv[u];
          const int cc = __float2int_rz((static_cast<float>(e) + 0.5f) * inv_rs), st = e - cc * rs;
          trow[cc * rsp + st] = pv[u];
        }
      }
    }
  }
  // ---- phase 2: every operand layout of the layer from the tile
  for (int j = 0; j < L.n_jobs; ++j) {
    __syncthreads();  // the tile is complete (first pass) / the previous job descriptor is no longer read
    {
      const int32_t* src = reinterpret_cast<const int32_t*>(jobs + L.job[j]);
      int32_t* dst = reinterpret_cast<int32_t*>(&job);
      for (int i = threadIdx.x; i < static_cast<int>(sizeof(job) / 4); i += blockDim.x) dst[i] = src[i];
    }
    __syncthreads();
    __nv_bfloat16* dst = static_cast<__nv_bfloat16*>(job.dst);
    const bool k_rows = job.src_stride_col < job.src_stride_row;  // fprop layout: rows = k, contiguous runs along c
    const int n_rows = k_rows ? nk : nc, n_cols = k_rows ? nc : nk;
    const int row0 = k_rows ? k0 : c0, col0 = k_rows ? c0 : k0;
    // an item = one (row, tap) run of the operand; a lane writes 4 consecutive columns (8 bytes), so a run of the
    // tile's 32 / 16 / 8 / 4 columns takes 8 / 4 / 2 / 1 lanes and a warp handles 4 / 8 / 16 / 32 items per iteration
    // (every lane busy: the first version gave a warp one item at a time and was instruction bound)
    const int tile_cols = k_rows ? L.tc : L.tk;
    const int lanes_per_item = tile_cols >= 32 ? 8 : (tile_cols >= 16 ? 4 : (tile_cols >= 8 ? 2 : 1));
    const int items_per_iter = 32 / lanes_per_item;
    const int sub = lane / lanes_per_item, c4 = (lane - sub * lanes_per_item) * 4;
    const int n_items = n_rows * job.taps;
    const float inv_taps = 1.f / static_cast<float>(job.taps);
    const int col_step = lanes_per_item * 4;  // tiles wider than 32 columns (1x1 filters) loop over the columns
    const int s_r = k_rows ? pitch_k : rsp, s_c = k_rows ? rsp : pitch_k;
    for (int it = warp * items_per_iter + sub; it < n_items; it += 8 * items_per_iter) {
      const int r = __float2int_rz((static_cast<float>(it) + 0.5f) * inv_taps), t = it - r * job.taps;
      const int st = job.src_tap[t];
      if (st < 0) continue;  // padding taps were zeroed when the table was built and never change
      __nv_bfloat16* drow = dst + (static_cast<int64_t>(row0 + r) * job.taps + t) * job.cols + col0;
      const float* trow = tile + r * s_r + st;
      for (int c = c4; c < n_cols; c += col_step) {
        float f[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) f[q] = c + q < n_cols ? trow[(c + q) * s_c] : 0.f;
        if (c + 2 < n_cols) {
          *reinterpret_cast<uint2*>(drow + c) = make_uint2(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]));
        } else {
          *reinterpret_cast<uint32_t*>(drow + c) = pack_bf16x2(f[0], f[1]);
        }
      }
    }
  }
}

}  // namespace

extern "C" {

int fpg_adam_pack_tile(int32_t rs, int32_t* tk, int32_t* tc) {
  FPG_REQUIRE(rs >= 1 && tk && tc, "bad argument");
  const int rsp = rs | 1;
  int k = 32, c = 32;
  // shrink the input-channel side first (its runs are the parameter's own contiguous direction), keep both even
  while (k * ((c * rsp) | 1) > kTileFloats && c > 4) c >>= 1;  // >= 4: a lane writes 4 columns with one 8-byte store
  while (k * ((c * rsp) | 1) > kTileFloats && k > 4) k >>= 1;
  FPG_REQUIRE(k * ((c * rsp) | 1) <= kTileFloats, "filter of %d taps does not fit a tile", rs);
  *tk = k;
  *tc = c;
  return 0;
}

int fpg_adam_pack_step(float* p, const void* const* grads_host, int32_t n_src, float* m, float* v, float beta1,
                       float beta2, float eps, int32_t* state, float grad_scale, float* gsum_out,
                       const fpg_adam_pack_layer* layers_dev, const fpg_pack_job* jobs_dev,
                       const fpg_adam_pack_chunk* chunks_dev, const int32_t* block_item_dev,
                       const int32_t* block_first_dev, int32_t n_blocks, void* stream) {
  FPG_REQUIRE(p && grads_host && m && v && state && n_src >= 1 && n_src <= kMaxSrc && block_item_dev && block_first_dev &&
                  n_blocks > 0,
              "bad argument");
  GradSrcs s;
  memset(&s, 0, sizeof(s));
  for (int i = 0; i < n_src; ++i) {
    FPG_REQUIRE(grads_host[i] != nullptr, "null gradient source");
    s.g[i] = static_cast<const float*>(grads_host[i]);
  }
  int rc = fpg_adam_prepare_dev(state, beta1, beta2, stream);
  if (rc) return rc;
  Hyper hy = {1.f - beta1, beta2, eps, grad_scale};
  adam_pack_kernel<<<static_cast<unsigned>(n_blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      p, s, n_src, m, v, state, hy, layers_dev, jobs_dev, chunks_dev, block_item_dev, block_first_dev, gsum_out);
  FPG_CUDA_CHECK(cudaGetLastError());
  return 0;
}

}  // extern "C"
