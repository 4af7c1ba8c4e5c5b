// Image-quality metrics of calculate_metrics (reference models/model.py:367-371, 404-406: torchmetrics
// PeakSignalNoiseRatio, StructuralSimilarityIndexMeasure, MultiScaleStructuralSimilarityIndexMeasure, all with
// data_range=(0, 1)) as device kernels on fp32 NCHW images.
//   SSIM (gaussian 11x11, sigma 1.5, k1 0.01, k2 0.03): torchmetrics reflect-pads by 5, filters, then CROPS 5 pixels
//   from every border -- the kept (H-10) x (W-10) values never see the padding, so the kernel evaluates the windows of
//   the interior only. Per image: mean SSIM and mean contrast sensitivity (the MS-SSIM ingredient) over C*(H-10)*(W-10).
//   MS-SSIM: five scales linked by 2x2 average pooling; the per-scale values are combined on the host (5 numbers/image).
//   PSNR: 10 log10(1 / mean squared error) over the whole batch.
// HBM-bound: one read of both images per scale (8 B/pixel/channel).
#include <math.h>

#include "common.cuh"
#include "host_util.h"

namespace fpg {

constexpr int kSsimK = 11, kSsimR = 5;
constexpr int kSsimTx = 32, kSsimTy = 8;

struct SsimArgs {
  int32_t n, c, h, w, blocks_per_image;
  float c1, c2;
  float g[kSsimK];
};

// block = 32 x 8 outputs of one (image, channel) plane of the cropped map. Separable filter: horizontal sums of
// {p, t, p^2, t^2, p t} into shared memory, vertical sums from there. partial[(image, block)][2] = {sum ssim, sum cs}.
__global__ void __launch_bounds__(kSsimTx * kSsimTy)
ssim_map_kernel(const float* __restrict__ pred, const float* __restrict__ target, float* __restrict__ partial,
                const SsimArgs a) {
  constexpr int PW = kSsimTx + kSsimK - 1, PH = kSsimTy + kSsimK - 1;  // 42 x 18 input patch
  __shared__ float sp[PH][PW], st[PH][PW];
  __shared__ float rows[PH][kSsimTx][5];
  __shared__ float red[kSsimTx * kSsimTy / 32][2];
  const int oh = a.h - 2 * kSsimR, ow = a.w - 2 * kSsimR;  // cropped map
  const int tiles_x = (ow + kSsimTx - 1) / kSsimTx, tiles_y = (oh + kSsimTy - 1) / kSsimTy;
  int blk = blockIdx.x;
  const int tx = blk % tiles_x;
  blk /= tiles_x;
  const int ty = blk % tiles_y;
  blk /= tiles_y;
  const int ch = blk % a.c, img = blk / a.c;
  const int x0 = tx * kSsimTx, y0 = ty * kSsimTy;  // cropped coordinates = image coordinates of the window origin
  const int64_t plane = (static_cast<int64_t>(img) * a.c + ch) * a.h * a.w;
  const int tid = threadIdx.y * kSsimTx + threadIdx.x;
  for (int i = tid; i < PH * PW; i += kSsimTx * kSsimTy) {
    const int py = i / PW, px = i - py * PW;
    const int y = min(y0 + py, a.h - 1), x = min(x0 + px, a.w - 1);  // clamped reads feed masked outputs only
    sp[py][px] = __ldg(pred + plane + static_cast<int64_t>(y) * a.w + x);
    st[py][px] = __ldg(target + plane + static_cast<int64_t>(y) * a.w + x);
  }
  __syncthreads();
  for (int i = tid; i < PH * kSsimTx; i += kSsimTx * kSsimTy) {
    const int py = i / kSsimTx, px = i - py * kSsimTx;
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f, s4 = 0.f;
#pragma unroll
    for (int k = 0; k < kSsimK; ++k) {
      const float p = sp[py][px + k], t = st[py][px + k], g = a.g[k];
      s0 += g * p;
      s1 += g * t;
      s2 += g * (p * p);
      s3 += g * (t * t);
      s4 += g * (p * t);
    }
    rows[py][px][0] = s0; rows[py][px][1] = s1; rows[py][px][2] = s2; rows[py][px][3] = s3; rows[py][px][4] = s4;
  }
  __syncthreads();
  float ssim = 0.f, cs = 0.f;
  {
    float m[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int k = 0; k < kSsimK; ++k) {
      const float g = a.g[k];
#pragma unroll
      for (int q = 0; q < 5; ++q) m[q] += g * rows[threadIdx.y + k][threadIdx.x][q];
    }
    const float mu_pp = m[0] * m[0], mu_tt = m[1] * m[1], mu_pt = m[0] * m[1];
    const float sig_p = m[2] - mu_pp, sig_t = m[3] - mu_tt, sig_pt = m[4] - mu_pt;
    const float upper = 2.f * sig_pt + a.c2, lower = sig_p + sig_t + a.c2;
    if (y0 + threadIdx.y < oh && x0 + threadIdx.x < ow) {
      ssim = ((2.f * mu_pt + a.c1) * upper) / ((mu_pp + mu_tt + a.c1) * lower);
      cs = upper / lower;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    ssim += __shfl_xor_sync(0xffffffffu, ssim, o);
    cs += __shfl_xor_sync(0xffffffffu, cs, o);
  }
  if ((tid & 31) == 0) {
    red[tid >> 5][0] = ssim;
    red[tid >> 5][1] = cs;
  }
  __syncthreads();
  if (tid == 0) {
    float s = 0.f, c = 0.f;
    for (int i = 0; i < kSsimTx * kSsimTy / 32; ++i) {
      s += red[i][0];
      c += red[i][1];
    }
    const int local = blockIdx.x - img * a.blocks_per_image;
    partial[(static_cast<int64_t>(img) * a.blocks_per_image + local) * 2] = s;
    partial[(static_cast<int64_t>(img) * a.blocks_per_image + local) * 2 + 1] = c;
  }
}

// out[img] = {mean ssim, mean cs}: fixed-order sum of the image's block partials (double accumulation)
__global__ void ssim_finalize_kernel(const float* __restrict__ partial, int blocks_per_image, double inv_count,
                                     float* __restrict__ out) {
  const int img = blockIdx.x, q = threadIdx.x;  // 2 threads
  double acc = 0.0;
  for (int i = 0; i < blocks_per_image; ++i) acc += partial[(static_cast<int64_t>(img) * blocks_per_image + i) * 2 + q];
  out[img * 2 + q] = static_cast<float>(acc * inv_count);
}

// y[n][c][h/2][w/2] = mean of 2x2 windows (F.avg_pool2d(x, (2, 2)))
__global__ void avgpool2_f32_kernel(const float* __restrict__ x, float* __restrict__ y, int planes, int h, int w) {
  const int oh = h / 2, ow = w / 2;
  const int64_t total = static_cast<int64_t>(planes) * oh * ow;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int ox = static_cast<int>(i % ow);
    const int oy = static_cast<int>((i / ow) % oh);
    const int64_t p = i / (static_cast<int64_t>(ow) * oh);
    const float* s = x + (p * h + 2 * oy) * w + 2 * ox;
    y[i] = (s[0] + s[1] + s[w] + s[w + 1]) * 0.25f;
  }
}

// partial[block] = sum over the block's elements of (clamp(a) - clamp(b))^2, then one thread adds the partials in order
__global__ void __launch_bounds__(256)
sq_err_partial_kernel(const float* __restrict__ a, const float* __restrict__ b, int64_t n, float lo, float hi,
                      double* __restrict__ partial) {
  __shared__ double red[8];
  double acc = 0.0;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const float d = fminf(fmaxf(__ldg(a + i), lo), hi) - fminf(fmaxf(__ldg(b + i), lo), hi);
    acc += static_cast<double>(d * d);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int i = 0; i < 8; ++i) s += red[i];
    partial[blockIdx.x] = s;
  }
}

__global__ void sq_err_finalize_kernel(const double* __restrict__ partial, int blocks, double* __restrict__ out) {
  double s = 0.0;
  for (int i = 0; i < blocks; ++i) s += partial[i];
  out[0] = s;
}

constexpr int kSqErrBlocks = 592;

}  // namespace fpg

using namespace fpg;

extern "C" int64_t fpg_ssim_scratch_bytes(int32_t n, int32_t c, int32_t h, int32_t w) {
  if (n <= 0 || c <= 0 || h <= 2 * kSsimR || w <= 2 * kSsimR) return -1;
  const int64_t tiles = static_cast<int64_t>(ceil_div(w - 2 * kSsimR, kSsimTx)) * ceil_div(h - 2 * kSsimR, kSsimTy);
  return static_cast<int64_t>(n) * c * tiles * 2 * 4;
}

extern "C" int fpg_ssim_stats(const float* pred, const float* target, int32_t n, int32_t c, int32_t h, int32_t w,
                              float data_range, float k1, float k2, float sigma, float* out, void* scratch,
                              void* stream) {
  FPG_REQUIRE(pred != nullptr && target != nullptr && out != nullptr && scratch != nullptr, "ssim: null pointer");
  FPG_REQUIRE(n > 0 && c > 0 && h > 2 * kSsimR && w > 2 * kSsimR, "ssim: images must exceed the 11 x 11 window");
  SsimArgs a;
  a.n = n;
  a.c = c;
  a.h = h;
  a.w = w;
  a.c1 = (k1 * data_range) * (k1 * data_range);
  a.c2 = (k2 * data_range) * (k2 * data_range);
  float sum = 0.f;
  for (int i = 0; i < kSsimK; ++i) {  // torchmetrics _gaussian: exp(-(d / sigma)^2 / 2), d = i - 5, normalised
    const float d = static_cast<float>(i - kSsimR) / sigma;
    a.g[i] = expf(-(d * d) / 2.f);
    sum += a.g[i];
  }
  for (int i = 0; i < kSsimK; ++i) a.g[i] /= sum;
  const int tiles = ceil_div(w - 2 * kSsimR, kSsimTx) * ceil_div(h - 2 * kSsimR, kSsimTy);
  a.blocks_per_image = c * tiles;
  const int64_t blocks = static_cast<int64_t>(n) * a.blocks_per_image;
  FPG_REQUIRE(blocks < (1ll << 31), "ssim: too many tiles");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  float* partial = static_cast<float*>(scratch);
  ssim_map_kernel<<<static_cast<unsigned>(blocks), dim3(kSsimTx, kSsimTy), 0, st>>>(pred, target, partial, a);
  const double inv = 1.0 / (static_cast<double>(c) * (h - 2 * kSsimR) * (w - 2 * kSsimR));
  ssim_finalize_kernel<<<n, 2, 0, st>>>(partial, a.blocks_per_image, inv, out);
  FPG_CUDA_CHECK(cudaGetLastError());
  return 0;
}

extern "C" int fpg_avgpool2_f32(const float* x, float* y, int32_t planes, int32_t h, int32_t w, void* stream) {
  FPG_REQUIRE(x != nullptr && y != nullptr && planes > 0 && h >= 2 && w >= 2, "avgpool2: bad arguments");
  const int64_t total = static_cast<int64_t>(planes) * (h / 2) * (w / 2);
  int blocks = static_cast<int>((total + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  avgpool2_f32_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(x, y, planes, h, w);
  FPG_CUDA_CHECK(cudaGetLastError());
  return 0;
}

extern "C" int64_t fpg_sq_err_scratch_bytes(void) { return kSqErrBlocks * 8; }

extern "C" int fpg_sq_err_sum(const float* a, const float* b, int64_t count, float clamp_lo, float clamp_hi,
                              double* out, void* scratch, void* stream) {
  FPG_REQUIRE(a != nullptr && b != nullptr && out != nullptr && scratch != nullptr && count > 0, "sq_err: bad arguments");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  double* partial = static_cast<double*>(scratch);
  sq_err_partial_kernel<<<kSqErrBlocks, 256, 0, st>>>(a, b, count, clamp_lo, clamp_hi, partial);
  sq_err_finalize_kernel<<<1, 1, 0, st>>>(partial, kSqErrBlocks, out);
  FPG_CUDA_CHECK(cudaGetLastError());
  return 0;
}
