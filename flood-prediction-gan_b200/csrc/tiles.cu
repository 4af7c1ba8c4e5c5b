// Input pipeline on the device (reference: models/utils.py:19-67 apply_transformations, models/data.py:57-78):
//   fpg_resize_bicubic_aa  decoded HWC fp32 stack -> channel selection (topography) -> optional horizontal flip ->
//                          anti-aliased bicubic resize (torchvision Resize(BICUBIC, antialias=True)) -> CHW fp32
//   fpg_tile_gather        batch assembly from the HBM-resident resized images: crop window of each sample + Normalize
// The reference repeats decode + resize on one CPU thread for every crop of every epoch; here an image is resized once
// (HBM-bound: one read of the 37.7 MB stack) and stays resident -- the whole resized dataset (2336 pairs x 12.6 MB fp32)
// is 29 GB of the 180 GB -- so a training step only gathers its crops.
#include <math.h>

#include "common.cuh"
#include "host_util.h"

namespace fpg {

// Anti-aliased bicubic resampling weights, the PIL / ATen algorithm (Keys kernel a = -0.5): output i takes the input
// samples [lo, lo + n), centre = scale * (i + 0.5), support = 2 * max(scale, 1), weights normalised to sum 1.
// Mixed float/double evaluation order as in ATen's _compute_indices_weights_aa (float scale and centre, the tap offsets
// through double) so that the tables agree with the CPU reference to the last bits.
__device__ __forceinline__ float cubic_aa(float x) {
  const float a = -0.5f;
  x = fabsf(x);
  if (x < 1.f) return __fadd_rn(__fmul_rn(__fmul_rn(__fadd_rn(__fmul_rn(a + 2.f, x), -(a + 3.f)), x), x), 1.f);
  if (x < 2.f)
    return __fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(a, x), -5.f * a), x), 8.f * a), x), -4.f * a);
  return 0.f;
}

struct AaAxis {
  int32_t in_size, out_size, max_taps;
  int32_t* lo;
  int32_t* n;
  float* w;
};

__global__ void aa_weights_kernel(const AaAxis ax0, const AaAxis ax1) {  // blockIdx.y = axis (0: width, 1: height)
  const AaAxis& ax = blockIdx.y == 0 ? ax0 : ax1;
  const int in_size = ax.in_size, out_size = ax.out_size, max_taps = ax.max_taps;
  int32_t* lo_out = ax.lo;
  int32_t* n_out = ax.n;
  float* w_out = ax.w;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= out_size) return;
  float* w = w_out + static_cast<int64_t>(i) * max_taps;
  if (in_size == out_size) {  // this axis is not resampled
    lo_out[i] = i;
    n_out[i] = 1;
    w[0] = 1.f;
    return;
  }
  const float scale = __fdiv_rn(static_cast<float>(in_size), static_cast<float>(out_size));
  const float support = scale >= 1.f ? __fmul_rn(2.f, scale) : 2.f;
  const float inv = scale >= 1.f ? __fdiv_rn(1.f, scale) : 1.f;
  const float center = static_cast<float>(static_cast<double>(scale) * (static_cast<double>(i) + 0.5));
  int lo = static_cast<int>(static_cast<long long>(static_cast<double>(__fadd_rn(center, -support)) + 0.5));
  if (lo < 0) lo = 0;
  int hi = static_cast<int>(static_cast<long long>(static_cast<double>(__fadd_rn(center, support)) + 0.5));
  if (hi > in_size) hi = in_size;
  int n = hi - lo;
  if (n > max_taps) n = max_taps;
  float total = 0.f;
  for (int j = 0; j < n; ++j) {
    const double off = (static_cast<double>(static_cast<float>(j + lo) - center) + 0.5) * static_cast<double>(inv);
    const float v = cubic_aa(static_cast<float>(off));
    w[j] = v;
    total = __fadd_rn(total, v);
  }
  if (total != 0.f)
    for (int j = 0; j < n; ++j) w[j] = __fdiv_rn(w[j], total);
  lo_out[i] = lo;
  n_out[i] = n;
}

struct ResizeArgs {
  int32_t in_h, in_w, in_c, channels, flip_w, out_h, out_w, taps_w, taps_h;
  int32_t chmap[16];
};

// pass 1 (horizontal): tmp[y][x'][c'] = sum_j w[x'][j] * src[y][lo + j][chmap[c']]  (source column mirrored if flip_w).
// One block = one source row x 128 output columns (or fewer for large scales): the source span those columns read
// (contiguous in the HWC row, all in_c channels) is staged in shared memory with coalesced loads -- reading it tap by
// tap from global memory touched 18 cache lines per warp load. A thread owns one output column and ALL its channels, so
// a weight is fetched once per tap and the tap's channels are consecutive shared-memory words (the one-output-per-
// thread form was instruction bound: 12 instructions per multiply-add). Sequential accumulation in tap order without
// FMA contraction (the reference's summation order); results go back through shared memory for contiguous stores.
template <int NC>
__global__ void __launch_bounds__(128)
resize_rows_kernel(const float* __restrict__ src, float* __restrict__ tmp, const int32_t* __restrict__ lo_t,
                   const int32_t* __restrict__ n_t, const float* __restrict__ w_t, const ResizeArgs a, int tx) {
  extern __shared__ float span[];
  const int y = blockIdx.y, x_first = blockIdx.x * tx;
  const int x_last = min(x_first + tx, a.out_w) - 1;
  const int p_lo = lo_t[x_first], p_hi = lo_t[x_last] + n_t[x_last];  // source pixels [p_lo, p_hi) (unflipped index)
  const int n_px = p_hi - p_lo;
  // flipped: logical pixel p is source pixel in_w - 1 - p, so the span is the mirrored contiguous range, read backwards
  const int s_lo = a.flip_w ? a.in_w - p_hi : p_lo;
  const float* row = src + (static_cast<int64_t>(y) * a.in_w + s_lo) * a.in_c;
  // 4-byte asynchronous copies (the span starts at a multiple of 36 bytes, not of 16): every thread's ~19 words are in
  // flight together instead of one global-load latency per word
  for (int i = threadIdx.x; i < n_px * a.in_c; i += blockDim.x)
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(span + i)), "l"(row + i) : "memory");
  asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
  __syncthreads();
  float acc[NC];
  const int x = x_first + threadIdx.x;
  const bool active = threadIdx.x < tx && x <= x_last;
  if (active) {
    const int lo = lo_t[x], n = n_t[x];
    const float* w = w_t + static_cast<int64_t>(x) * a.taps_w;
    for (int j = 0; j < n; ++j) {
      const int p = lo + j - p_lo;  // position inside the logical span
      const float* px = span + (a.flip_w ? n_px - 1 - p : p) * a.in_c;
      const float wj = __ldg(w + j);
#pragma unroll
      for (int c = 0; c < NC; ++c) {
        const float v = __fmul_rn(px[a.chmap[c]], wj);
        acc[c] = j == 0 ? v : __fadd_rn(acc[c], v);
      }
    }
  }
  __syncthreads();  // the span is dead: reuse it as the [tx][NC] output staging buffer
  if (active) {
#pragma unroll
    for (int c = 0; c < NC; ++c) span[threadIdx.x * NC + c] = acc[c];
  }
  __syncthreads();
  const int n_out = (x_last - x_first + 1) * NC;
  float* out = tmp + (static_cast<int64_t>(y) * a.out_w + x_first) * NC;
  for (int o = threadIdx.x; o < n_out; o += blockDim.x) out[o] = span[o];
}

// pass 2 (vertical): dst[c'][y'][x'] = sum_j w[y'][j] * tmp[lo + j][x'][c']. block = 32 columns x channels.
__global__ void resize_cols_kernel(const float* __restrict__ tmp, float* __restrict__ dst,
                                   const int32_t* __restrict__ lo_t, const int32_t* __restrict__ n_t,
                                   const float* __restrict__ w_t, const ResizeArgs a) {
  const int x = blockIdx.x * 32 + threadIdx.x, c = threadIdx.y, y = blockIdx.y;
  if (x >= a.out_w) return;
  const int lo = lo_t[y], n = n_t[y];
  const float* w = w_t + static_cast<int64_t>(y) * a.taps_h;
  const float* col = tmp + static_cast<int64_t>(x) * a.channels + c;
  const int64_t pitch = static_cast<int64_t>(a.out_w) * a.channels;
  float acc = 0.f;
  int j = 0;
  for (; j + 8 <= n; j += 8) {  // loads batched eight at a time (one latency per batch), summation order unchanged
    float v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) v[u] = __ldg(col + (lo + j + u) * pitch);
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const float t = __fmul_rn(v[u], w[j + u]);
      acc = (j + u) == 0 ? t : __fadd_rn(acc, t);
    }
  }
  for (; j < n; ++j) {
    const float t = __fmul_rn(__ldg(col + (lo + j) * pitch), w[j]);
    acc = j == 0 ? t : __fadd_rn(acc, t);
  }
  dst[(static_cast<int64_t>(c) * a.out_h + y) * a.out_w + x] = acc;
}

// dst[b][c][y][x] = (images[b][c][r0 + y][c0 + x] - mean) / std, (r0, c0) = window crop_index[b] of a
// divisions x divisions grid (models/utils.py:45-61). VEC = 4: 16-byte loads / stores (window width and offsets
// multiples of 4 floats), one thread per 4 pixels of a row.
template <int VEC>
__global__ void __launch_bounds__(256)
tile_gather_kernel(const float* const* __restrict__ images, int channels, int height, int width,
                   const int32_t* __restrict__ crops, int divisions, float mean, float stdv, float* __restrict__ dst) {
  const int th = height / divisions, tw = width / divisions;
  const int b = blockIdx.z, c = blockIdx.y;
  const int crop = crops[b];
  const int r0 = (crop / divisions) * th, c0 = (crop % divisions) * tw;
  const float* src = images[b] + static_cast<int64_t>(c) * height * width + static_cast<int64_t>(r0) * width + c0;
  float* out = dst + (static_cast<int64_t>(b) * channels + c) * th * tw;
  const int row_vecs = tw / VEC;
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < th * row_vecs; p += gridDim.x * blockDim.x) {
    const int y = p / row_vecs, xv = p - y * row_vecs;
    if (VEC == 4) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(src + static_cast<int64_t>(y) * width) + xv);
      float4 o;
      o.x = __fdiv_rn(__fadd_rn(v.x, -mean), stdv);
      o.y = __fdiv_rn(__fadd_rn(v.y, -mean), stdv);
      o.z = __fdiv_rn(__fadd_rn(v.z, -mean), stdv);
      o.w = __fdiv_rn(__fadd_rn(v.w, -mean), stdv);
      reinterpret_cast<float4*>(out + static_cast<int64_t>(y) * tw)[xv] = o;
    } else {
      out[static_cast<int64_t>(y) * tw + xv] =
          __fdiv_rn(__fadd_rn(__ldg(src + static_cast<int64_t>(y) * width + xv), -mean), stdv);
    }
  }
}

static int aa_max_taps(int in_size, int out_size) {
  if (in_size == out_size) return 1;
  const float scale = static_cast<float>(in_size) / static_cast<float>(out_size);
  const float support = scale >= 1.f ? 2.f * scale : 2.f;
  return static_cast<int>(ceilf(support)) * 2 + 1;
}

struct ResizeScratch {
  int64_t tmp, lo_w, n_w, w_w, lo_h, n_h, w_h, total;  // byte offsets
};

static ResizeScratch resize_scratch(int in_h, int in_w, int out_h, int out_w, int channels) {
  auto align = [](int64_t v) { return (v + 255) & ~static_cast<int64_t>(255); };
  ResizeScratch s;
  int64_t off = 0;
  s.tmp = off;
  off = align(off + static_cast<int64_t>(in_h) * out_w * channels * 4);
  s.lo_w = off;
  off = align(off + static_cast<int64_t>(out_w) * 4);
  s.n_w = off;
  off = align(off + static_cast<int64_t>(out_w) * 4);
  s.w_w = off;
  off = align(off + static_cast<int64_t>(out_w) * aa_max_taps(in_w, out_w) * 4);
  s.lo_h = off;
  off = align(off + static_cast<int64_t>(out_h) * 4);
  s.n_h = off;
  off = align(off + static_cast<int64_t>(out_h) * 4);
  s.w_h = off;
  off = align(off + static_cast<int64_t>(out_h) * aa_max_taps(in_h, out_h) * 4);
  s.total = off;
  return s;
}

}  // namespace fpg

using namespace fpg;

extern "C" int64_t fpg_resize_aa_scratch_bytes(int32_t in_h, int32_t in_w, int32_t out_h, int32_t out_w,
                                               int32_t channels) {
  if (in_h <= 0 || in_w <= 0 || out_h <= 0 || out_w <= 0 || channels <= 0 || channels > 16) return -1;
  return resize_scratch(in_h, in_w, out_h, out_w, channels).total;
}

extern "C" int fpg_resize_bicubic_aa(const float* src_hwc, int32_t in_h, int32_t in_w, int32_t in_c,
                                     const int32_t* channel_map, int32_t channels, int32_t flip_w, int32_t out_h,
                                     int32_t out_w, float* dst_chw, void* scratch, void* stream) {
  FPG_REQUIRE(src_hwc != nullptr && dst_chw != nullptr && scratch != nullptr && channel_map != nullptr,
              "resize: null pointer");
  FPG_REQUIRE(in_h > 0 && in_w > 0 && out_h > 0 && out_w > 0 && in_c > 0, "resize: empty image");
  FPG_REQUIRE(channels > 0 && channels <= 16, "resize: 1..16 selected channels");
  ResizeArgs a;
  a.in_h = in_h;
  a.in_w = in_w;
  a.in_c = in_c;
  a.channels = channels;
  a.flip_w = flip_w ? 1 : 0;
  a.out_h = out_h;
  a.out_w = out_w;
  a.taps_w = aa_max_taps(in_w, out_w);
  a.taps_h = aa_max_taps(in_h, out_h);
  for (int i = 0; i < 16; ++i) a.chmap[i] = 0;
  for (int i = 0; i < channels; ++i) {
    FPG_REQUIRE(channel_map[i] >= 0 && channel_map[i] < in_c, "resize: channel_map[%d] = %d outside [0, %d)", i,
                channel_map[i], in_c);
    a.chmap[i] = channel_map[i];
  }
  const ResizeScratch s = resize_scratch(in_h, in_w, out_h, out_w, channels);
  uint8_t* base = static_cast<uint8_t*>(scratch);
  float* tmp = reinterpret_cast<float*>(base + s.tmp);
  int32_t* lo_w = reinterpret_cast<int32_t*>(base + s.lo_w);
  int32_t* n_w = reinterpret_cast<int32_t*>(base + s.n_w);
  float* w_w = reinterpret_cast<float*>(base + s.w_w);
  int32_t* lo_h = reinterpret_cast<int32_t*>(base + s.lo_h);
  int32_t* n_h = reinterpret_cast<int32_t*>(base + s.n_h);
  float* w_h = reinterpret_cast<float*>(base + s.w_h);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const AaAxis ax_w = {in_w, out_w, a.taps_w, lo_w, n_w, w_w}, ax_h = {in_h, out_h, a.taps_h, lo_h, n_h, w_h};
  aa_weights_kernel<<<dim3(ceil_div(out_w > out_h ? out_w : out_h, 128), 2), 128, 0, st>>>(ax_w, ax_h);
  // output columns per block of the horizontal pass: the source span of the block must fit 48 KB of shared memory
  const float scale_w = in_w == out_w ? 1.f : static_cast<float>(in_w) / static_cast<float>(out_w);
  int tx = 128;
  auto span_px = [&](int t) {
    const int px = static_cast<int>(ceilf(t * scale_w)) + a.taps_w + 2;
    return px < in_w ? px : in_w;
  };
  while (tx > 1 && static_cast<int64_t>(span_px(tx)) * in_c * 4 > 48 * 1024) tx /= 2;
  size_t span_bytes = static_cast<size_t>(span_px(tx)) * in_c * 4;
  FPG_REQUIRE(span_bytes <= 48 * 1024, "resize: a source row span of %zu bytes does not fit shared memory", span_bytes);
  if (span_bytes < static_cast<size_t>(tx) * channels * 4) span_bytes = static_cast<size_t>(tx) * channels * 4;
  const dim3 rows_grid(ceil_div(out_w, tx), in_h);
#define FPG_ROWS_CASE(NC)                                                                                    \
  case NC:                                                                                                   \
    resize_rows_kernel<NC><<<rows_grid, 128, span_bytes, st>>>(src_hwc, tmp, lo_w, n_w, w_w, a, tx);        \
    break;
  switch (channels) {
    FPG_ROWS_CASE(1) FPG_ROWS_CASE(2) FPG_ROWS_CASE(3) FPG_ROWS_CASE(4) FPG_ROWS_CASE(5) FPG_ROWS_CASE(6)
    FPG_ROWS_CASE(7) FPG_ROWS_CASE(8) FPG_ROWS_CASE(9) FPG_ROWS_CASE(10) FPG_ROWS_CASE(11) FPG_ROWS_CASE(12)
    FPG_ROWS_CASE(13) FPG_ROWS_CASE(14) FPG_ROWS_CASE(15) FPG_ROWS_CASE(16)
    default: return fail(FPG_EINVAL, "resize: 1..16 selected channels");
  }
#undef FPG_ROWS_CASE
  resize_cols_kernel<<<dim3(ceil_div(out_w, 32), out_h), dim3(32, channels), 0, st>>>(tmp, dst_chw, lo_h, n_h, w_h, a);
  FPG_CUDA_CHECK(cudaGetLastError());
  return 0;
}

extern "C" int fpg_tile_gather(const float* const* images, int32_t channels, int32_t height, int32_t width,
                               const int32_t* crop_index, int32_t batch, int32_t divisions, float mean, float stdv,
                               float* dst, void* stream) {
  FPG_REQUIRE(images != nullptr && crop_index != nullptr && dst != nullptr, "gather: null pointer");
  FPG_REQUIRE(channels > 0 && height > 0 && width > 0 && divisions > 0 && stdv != 0.f, "gather: bad geometry");
  if (batch <= 0) return 0;
  FPG_REQUIRE(batch <= 65535 && channels <= 65535, "gather: batch / channels exceed the grid limits");
  const int th = height / divisions, tw = width / divisions;
  const bool vec = tw % 4 == 0 && width % 4 == 0;  // window offsets are multiples of tw; images are 256-byte aligned
  const int work = th * (vec ? tw / 4 : tw);
  int bx = ceil_div(work, 256 * 4);
  if (bx < 1) bx = 1;
  const dim3 grid(bx, channels, batch);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (vec)
    tile_gather_kernel<4><<<grid, 256, 0, st>>>(images, channels, height, width, crop_index, divisions, mean, stdv, dst);
  else
    tile_gather_kernel<1><<<grid, 256, 0, st>>>(images, channels, height, width, crop_index, divisions, mean, stdv, dst);
  FPG_CUDA_CHECK(cudaGetLastError());
  return 0;
}
