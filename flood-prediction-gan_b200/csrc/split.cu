// fp32 parity mode ("3 x bf16 split"): elementwise kernels around the tensor-core convolutions.
//
// north_star asks for a mode in which generator outputs and losses match the fp32 reference to rtol 1e-4. tcgen05 has
// no fp32 MMA, so every operand is carried as a PAIR of bf16 tensors, v = hi + lo with hi = bf16(v), lo = bf16(v - hi)
// (16 mantissa bits), and a convolution y = x * w is evaluated as  hi_x*hi_w + lo_x*hi_w + hi_x*lo_w  with fp32
// accumulation in TMEM (the dropped lo*lo term and the rounding of lo are both 2^-18 relative). The three products are
// ONE launch of the ordinary implicit-GEMM kernels: the activation tensor holds the channel blocks [hi | lo | hi]
// (3 * C channels) and the packed weight the blocks [hi_w | hi_w | lo_w] along its contraction dimension. Convolution
// outputs stay fp32; the kernels below do the normalisation / activation / residual add in fp32 and emit the next
// layer's [hi | lo | hi] operand. Simple one-CTA-per-(image, 8 channels) kernels: this mode exists to demonstrate
// parity, not speed.
#include "common.cuh"
#include "host_util.h"

namespace fpg {

__device__ __forceinline__ float split_act(float v, int act) {
  return act == FPG_ACT_RELU ? fmaxf(v, 0.f) : (act == FPG_ACT_LEAKY ? (v > 0.f ? v : 0.2f * v) : v);
}
__device__ __forceinline__ int split_reflect(int q, int n) { return q < 0 ? -q : (q >= n ? 2 * (n - 1) - q : q); }

// hi / lo halves of 8 fp32 values as two 16-byte bf16 vectors
__device__ __forceinline__ void split8(const float (&v)[8], uint4* hi, uint4* lo) {
  uint32_t h[4], l[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const __nv_bfloat16 h0 = __float2bfloat16(v[2 * k]), h1 = __float2bfloat16(v[2 * k + 1]);
    const float r0 = v[2 * k] - __bfloat162float(h0), r1 = v[2 * k + 1] - __bfloat162float(h1);
    h[k] = pack_bf16x2(__bfloat162float(h0), __bfloat162float(h1));
    l[k] = pack_bf16x2(r0, r1);
  }
  *hi = make_uint4(h[0], h[1], h[2], h[3]);
  *lo = make_uint4(l[0], l[1], l[2], l[3]);
}

// y: fp32 NHWC [n][h][w][c] (halo-free, channel stride ycs). grid (c / 8, n), 256 threads.
//   v = act(norm ? (y - mean) * rstd : y) + residual        (mean / biased variance over the h*w plane, in double)
//   skip_out[n][h][w][c] = v (fp32, optional);  out (bf16, 3*c channels, halo `halo` mirrored) = [hi(v) | lo(v) | hi(v)]
__global__ void __launch_bounds__(256)
norm_split_kernel(const float* __restrict__ y, int ycs, int h, int w, int c, int norm, float eps, int act,
                  const float* __restrict__ residual, float* __restrict__ skip_out, __nv_bfloat16* __restrict__ out,
                  int halo) {
  const int g = blockIdx.x, i = blockIdx.y;
  const int hw = h * w;
  const float* yb = y + static_cast<int64_t>(i) * hw * ycs + g * 8;
  float mean[8], rstd[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    mean[k] = 0.f;
    rstd[k] = 1.f;
  }
  if (norm) {
    double s[8], ss[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) s[k] = ss[k] = 0.0;
    for (int p = threadIdx.x; p < hw; p += blockDim.x) {
      const float4 a = *reinterpret_cast<const float4*>(yb + static_cast<int64_t>(p) * ycs);
      const float4 b = *reinterpret_cast<const float4*>(yb + static_cast<int64_t>(p) * ycs + 4);
      const float f[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        s[k] += f[k];
        ss[k] += static_cast<double>(f[k]) * f[k];
      }
    }
    __shared__ double red[256][17];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      red[threadIdx.x][k] = s[k];
      red[threadIdx.x][8 + k] = ss[k];
    }
    __syncthreads();
    for (int off = 128; off > 0; off >>= 1) {
      if (static_cast<int>(threadIdx.x) < off)
        for (int k = 0; k < 16; ++k) red[threadIdx.x][k] += red[threadIdx.x + off][k];
      __syncthreads();
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const double m = red[0][k] / hw;
      double var = red[0][8 + k] / hw - m * m;
      if (var < 0.0) var = 0.0;
      mean[k] = static_cast<float>(m);
      rstd[k] = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
    }
  }
  const int hp = h + 2 * halo, wp = w + 2 * halo;
  const int64_t ocs = 3 * static_cast<int64_t>(c);
  __nv_bfloat16* ob = out + static_cast<int64_t>(i) * hp * wp * ocs + g * 8;
  for (int q = threadIdx.x; q < hp * wp; q += blockDim.x) {
    const int py = q / wp, px = q - py * wp;
    const int sy = split_reflect(py - halo, h), sx = split_reflect(px - halo, w);
    const int64_t p = static_cast<int64_t>(sy) * w + sx;
    const float4 a = *reinterpret_cast<const float4*>(yb + p * ycs);
    const float4 b = *reinterpret_cast<const float4*>(yb + p * ycs + 4);
    float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = split_act(norm ? (v[k] - mean[k]) * rstd[k] : v[k], act);
    if (residual != nullptr) {
      const float* rp = residual + (static_cast<int64_t>(i) * hw + p) * c + g * 8;
      const float4 ra = *reinterpret_cast<const float4*>(rp), rb = *reinterpret_cast<const float4*>(rp + 4);
      v[0] += ra.x; v[1] += ra.y; v[2] += ra.z; v[3] += ra.w; v[4] += rb.x; v[5] += rb.y; v[6] += rb.z; v[7] += rb.w;
    }
    const bool interior = py - halo == sy && px - halo == sx;
    if (skip_out != nullptr && interior) {
      float* sp = skip_out + (static_cast<int64_t>(i) * hw + p) * c + g * 8;
      *reinterpret_cast<float4*>(sp) = make_float4(v[0], v[1], v[2], v[3]);
      *reinterpret_cast<float4*>(sp + 4) = make_float4(v[4], v[5], v[6], v[7]);
    }
    uint4 hi, lo;
    split8(v, &hi, &lo);
    __nv_bfloat16* op = ob + static_cast<int64_t>(q) * ocs;
    *reinterpret_cast<uint4*>(op) = hi;
    *reinterpret_cast<uint4*>(op + c) = lo;
    *reinterpret_cast<uint4*>(op + 2 * c) = hi;
  }
}

// src fp32 NCHW [n][c_src][h][w] -> dst bf16 [n][h + 2 halo][w + 2 halo][3 * cpad] = [hi | lo | hi], channels >= c_src 0
__global__ void pack_nchw_split_kernel(const float* __restrict__ src, int c_src, int h, int w, int cpad, int halo,
                                       __nv_bfloat16* __restrict__ dst, int n) {
  const int hp = h + 2 * halo, wp = w + 2 * halo;
  const int64_t total = static_cast<int64_t>(n) * hp * wp;
  const int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int px = static_cast<int>(idx % wp);
  const int64_t r = idx / wp;
  const int py = static_cast<int>(r % hp);
  const int i = static_cast<int>(r / hp);
  const int sy = split_reflect(py - halo, h), sx = split_reflect(px - halo, w);
  const int64_t hw = static_cast<int64_t>(h) * w;
  const float* sp = src + static_cast<int64_t>(i) * c_src * hw + static_cast<int64_t>(sy) * w + sx;
  __nv_bfloat16* dp = dst + idx * 3 * cpad;
  for (int c8 = 0; c8 < cpad; c8 += 8) {
    float v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = (c8 + k < c_src) ? sp[(c8 + k) * hw] : 0.f;
    uint4 hi, lo;
    split8(v, &hi, &lo);
    *reinterpret_cast<uint4*>(dp + c8) = hi;
    *reinterpret_cast<uint4*>(dp + cpad + c8) = lo;
    *reinterpret_cast<uint4*>(dp + 2 * cpad + c8) = hi;
  }
}

}  // namespace fpg

using namespace fpg;

extern "C" {

int fpg_norm_split_f32(const fpg_act* y, int norm, float eps, int act, const float* residual, float* skip_out,
                       const fpg_act* out, void* stream) {
  FPG_REQUIRE(y && out, "null argument");
  FPG_REQUIRE(y->fp32 == FPG_DT_FP32 && y->halo == 0 && y->c % 8 == 0 && y->c_stride % 4 == 0,
              "y must be a halo-free fp32 tensor with a multiple of 8 channels");
  FPG_REQUIRE(out->fp32 == FPG_DT_BF16 && out->n == y->n && out->h == y->h && out->w == y->w && out->c == 3 * y->c &&
                  out->c_stride == out->c && out->halo < y->h && out->halo < y->w,
              "out must be a bf16 tensor of 3 * c channels ([hi | lo | hi]) with y's extent");
  norm_split_kernel<<<dim3(y->c / 8, y->n), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const float*>(y->data), y->c_stride, y->h, y->w, y->c, norm, eps, act, residual, skip_out,
      static_cast<__nv_bfloat16*>(out->data), out->halo);
  FPG_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int fpg_pack_nchw_split(const float* src, int32_t c_src, const fpg_act* dst, void* stream) {
  FPG_REQUIRE(src && dst && dst->fp32 == FPG_DT_BF16 && dst->c % 24 == 0 && dst->c == dst->c_stride &&
                  c_src <= dst->c / 3,
              "dst must be a bf16 tensor of 3 * cpad channels");
  const int64_t total = static_cast<int64_t>(dst->n) * (dst->h + 2 * dst->halo) * (dst->w + 2 * dst->halo);
  pack_nchw_split_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      src, c_src, dst->h, dst->w, dst->c / 3, dst->halo, static_cast<__nv_bfloat16*>(dst->data), dst->n);
  FPG_CUDA_CHECK(cudaGetLastError());
  return 0;
}

}  // extern "C"
