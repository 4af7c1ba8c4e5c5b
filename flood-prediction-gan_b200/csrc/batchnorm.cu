// BatchNorm2d (training mode: batch statistics, affine, running statistics), dropout and the two-consumer activation
// backward of the Pix2Pix U-Net and its BatchNorm PatchGAN (model_architectures.py:9-85). All tensors are halo-free
// NHWC bf16; destinations may be channel slices of a wider buffer (the U-Net's concatenation buffers).
//
// A U-Net encoder activation e has two consumers because the reference's activations are in place (:33-34, :63):
// lrelu(e) feeds the next down-convolution and relu(e) feeds the up-convolution through the skip connection, so the
// apply kernel writes up to two activated copies and the backward kernels take up to two upstream gradients.
#include <stdlib.h>

#include "common.cuh"
#include "host_util.h"

namespace fpg {

struct BView {
  void* p;
  int32_t n, h, w, c, cs;
  __device__ __forceinline__ int64_t at(int64_t pix) const { return pix * cs; }
};

static BView bview_of(const fpg_act* a) {
  BView v;
  v.p = a->data;
  v.n = a->n;
  v.h = a->h;
  v.w = a->w;
  v.c = a->c;
  v.cs = a->c_stride;
  return v;
}

__device__ __forceinline__ void bn_load8(const __nv_bfloat16* p, float (&f)[8]) {
  const uint4 u = *reinterpret_cast<const uint4*>(p);
  float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
  f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
}
__device__ __forceinline__ void bn_store8(__nv_bfloat16* p, const float (&f)[8]) {
  *reinterpret_cast<uint4*>(p) = make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]),
                                            pack_bf16x2(f[6], f[7]));
}
__device__ __forceinline__ float bn_act(float v, int act) {
  return act == FPG_ACT_RELU ? fmaxf(v, 0.f) : (act == FPG_ACT_LEAKY ? (v > 0.f ? v : 0.2f * v) : v);
}
// derivative of the activation at pre-activation value v (torch: slope at v <= 0)
__device__ __forceinline__ float bn_act_grad(float v, int act) {
  return act == FPG_ACT_RELU ? (v > 0.f ? 1.f : 0.f) : (act == FPG_ACT_LEAKY ? (v > 0.f ? 1.f : 0.2f) : 1.f);
}

struct ChanParams {
  float mean[8], rstd[8], gamma[8], beta[8];
};
// per-thread channel group: {mean, rstd} pairs of stats (identity when stats == nullptr), affine weight / bias
__device__ __forceinline__ void load_chan(ChanParams& cp, const float* stats, const float* gamma, const float* beta,
                                          int c0) {
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    cp.mean[k] = stats ? stats[(c0 + k) * 2] : 0.f;
    cp.rstd[k] = stats ? stats[(c0 + k) * 2 + 1] : 1.f;
    cp.gamma[k] = gamma ? gamma[c0 + k] : 1.f;
    cp.beta[k] = beta ? beta[c0 + k] : 0.f;
  }
}

// z = (gamma * (y - mean) * rstd + beta) * mask; z1 = act1(z), z2 = act2(z). One thread per (pixel, 8 channels).
__global__ void __launch_bounds__(256)
bn_apply_kernel(BView y, const float* __restrict__ stats, const float* __restrict__ gamma,
                const float* __restrict__ beta, const uint8_t* __restrict__ mask, float mask_scale, int act1, BView z1,
                int act2, BView z2, int has_z2) {
  const int G = y.c / 8;
  const int g = threadIdx.x % G, pl = threadIdx.x / G, lanes = blockDim.x / G;
  const int64_t npix = static_cast<int64_t>(y.n) * y.h * y.w;
  ChanParams cp;
  load_chan(cp, stats, gamma, beta, g * 8);
  for (int64_t p = static_cast<int64_t>(blockIdx.x) * lanes + pl; p < npix; p += static_cast<int64_t>(gridDim.x) * lanes) {
    float f[8], o1[8], o2[8];
    bn_load8(static_cast<const __nv_bfloat16*>(y.p) + y.at(p) + g * 8, f);
    uint2 mk = make_uint2(0x01010101u, 0x01010101u);
    if (mask) mk = *reinterpret_cast<const uint2*>(mask + p * y.c + g * 8);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      float v = (f[k] - cp.mean[k]) * cp.rstd[k] * cp.gamma[k] + cp.beta[k];
      if (mask) v = ((k < 4 ? mk.x >> (8 * k) : mk.y >> (8 * (k - 4))) & 0xFFu) ? v * mask_scale : 0.f;
      o1[k] = bn_act(v, act1);
      o2[k] = bn_act(v, act2);
    }
    bn_store8(static_cast<__nv_bfloat16*>(z1.p) + z1.at(p) + g * 8, o1);
    if (has_z2) bn_store8(static_cast<__nv_bfloat16*>(z2.p) + z2.at(p) + g * 8, o2);
  }
}

// upstream gradient w.r.t. the normalised-and-affine value v: g = (dz1 * act1'(z) + dz2 * act2'(z)) * mask
__device__ __forceinline__ void bn_upstream(const BView& dz1, int act1, const BView& dz2, int has_dz2, int act2,
                                            const uint8_t* mask, float mask_scale, const BView& y, int64_t p, int g,
                                            const ChanParams& cp, float (&gv)[8], float (&zh)[8]) {
  float f[8], a[8], b[8];
  bn_load8(static_cast<const __nv_bfloat16*>(y.p) + y.at(p) + g * 8, f);
  bn_load8(static_cast<const __nv_bfloat16*>(dz1.p) + dz1.at(p) + g * 8, a);
  if (has_dz2) bn_load8(static_cast<const __nv_bfloat16*>(dz2.p) + dz2.at(p) + g * 8, b);
  uint2 mk = make_uint2(0x01010101u, 0x01010101u);
  if (mask) mk = *reinterpret_cast<const uint2*>(mask + p * y.c + g * 8);
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    zh[k] = (f[k] - cp.mean[k]) * cp.rstd[k];
    float v = zh[k] * cp.gamma[k] + cp.beta[k];
    float m = 1.f;
    if (mask) m = ((k < 4 ? mk.x >> (8 * k) : mk.y >> (8 * (k - 4))) & 0xFFu) ? mask_scale : 0.f;
    v *= m;
    float gz = a[k] * bn_act_grad(v, act1);
    if (has_dz2) gz += b[k] * bn_act_grad(v, act2);
    gv[k] = gz * m;
  }
}

constexpr int kBnBlocks = 296;

// pass 1: partial[block][c][2] = {sum g, sum g * zhat} over the block's pixels (fixed-order block reduction)
__global__ void __launch_bounds__(256)
bn_bwd_reduce_kernel(BView dz1, int act1, BView dz2, int has_dz2, int act2, const uint8_t* __restrict__ mask,
                     float mask_scale, BView y, const float* __restrict__ stats, const float* __restrict__ gamma,
                     const float* __restrict__ beta, float* __restrict__ partial) {
  const int G = y.c / 8;
  const int g = threadIdx.x % G, pl = threadIdx.x / G, lanes = blockDim.x / G;
  const int64_t npix = static_cast<int64_t>(y.n) * y.h * y.w;
  const int64_t p_begin = npix * blockIdx.x / gridDim.x, p_end = npix * (blockIdx.x + 1) / gridDim.x;
  ChanParams cp;
  load_chan(cp, stats, gamma, beta, g * 8);
  float s[8], ss[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) s[k] = ss[k] = 0.f;
  for (int64_t p = p_begin + pl; p < p_end; p += lanes) {
    float gv[8], zh[8];
    bn_upstream(dz1, act1, dz2, has_dz2, act2, mask, mask_scale, y, p, g, cp, gv, zh);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      s[k] += gv[k];
      ss[k] += gv[k] * zh[k];
    }
  }
  __shared__ float red[256][17];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    red[threadIdx.x][k] = s[k];
    red[threadIdx.x][8 + k] = ss[k];
  }
  __syncthreads();
  for (int o = threadIdx.x; o < G * 16; o += blockDim.x) {
    const int gg = o / 16, comp = o % 16;
    float acc = 0.f;
    for (int l = 0; l < lanes; ++l) acc += red[l * G + gg][comp];
    partial[(static_cast<int64_t>(blockIdx.x) * y.c + gg * 8 + (comp & 7)) * 2 + (comp >> 3)] = acc;
  }
}

// sums[c] = {sum g, sum g zhat}; dbeta = sum g, dgamma = sum g zhat (accumulated if accumulate != 0)
__global__ void bn_bwd_finalize_kernel(const float* __restrict__ partial, int nblocks, int c, float* __restrict__ sums,
                                       float* __restrict__ dgamma, float* __restrict__ dbeta, int accumulate) {
  const int ch = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (ch >= c) return;
  float a = 0.f, b = 0.f;
  for (int blk = threadIdx.x & 31; blk < nblocks; blk += 32) {
    a += partial[(static_cast<int64_t>(blk) * c + ch) * 2];
    b += partial[(static_cast<int64_t>(blk) * c + ch) * 2 + 1];
  }
  a = warp_sum(a);
  b = warp_sum(b);
  if ((threadIdx.x & 31) == 0) {
    sums[ch * 2] = a;
    sums[ch * 2 + 1] = b;
    if (dbeta) dbeta[ch] = accumulate ? dbeta[ch] + a : a;
    if (dgamma) dgamma[ch] = accumulate ? dgamma[ch] + b : b;
  }
}

// pass 2: dy = gamma * rstd * (g - mean(g) - zhat * mean(g zhat))   (normalised: stats != nullptr)
//         dy = g                                                    (no normalisation: two-consumer activation backward)
__global__ void __launch_bounds__(256)
bn_bwd_apply_kernel(BView dz1, int act1, BView dz2, int has_dz2, int act2, const uint8_t* __restrict__ mask,
                    float mask_scale, BView y, const float* __restrict__ stats, const float* __restrict__ gamma,
                    const float* __restrict__ beta, const float* __restrict__ sums, float inv_count, BView dy) {
  const int G = y.c / 8;
  const int g = threadIdx.x % G, pl = threadIdx.x / G, lanes = blockDim.x / G;
  const int64_t npix = static_cast<int64_t>(y.n) * y.h * y.w;
  ChanParams cp;
  load_chan(cp, stats, gamma, beta, g * 8);
  float m1[8], m2[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    m1[k] = sums ? sums[(g * 8 + k) * 2] * inv_count : 0.f;
    m2[k] = sums ? sums[(g * 8 + k) * 2 + 1] * inv_count : 0.f;
  }
  for (int64_t p = static_cast<int64_t>(blockIdx.x) * lanes + pl; p < npix; p += static_cast<int64_t>(gridDim.x) * lanes) {
    float gv[8], zh[8], o[8];
    bn_upstream(dz1, act1, dz2, has_dz2, act2, mask, mask_scale, y, p, g, cp, gv, zh);
#pragma unroll
    for (int k = 0; k < 8; ++k)
      o[k] = stats ? cp.gamma[k] * cp.rstd[k] * (gv[k] - m1[k] - zh[k] * m2[k]) : gv[k];
    bn_store8(static_cast<__nv_bfloat16*>(dy.p) + dy.at(p) + g * 8, o);
  }
}

// running_mean / running_var update of nn.BatchNorm2d (momentum 0.1, unbiased variance), from {mean, rstd}
__global__ void bn_running_kernel(const float* __restrict__ stats, float eps, float momentum, float unbias, int c,
                                  float* __restrict__ running_mean, float* __restrict__ running_var) {
  const int ch = blockIdx.x * blockDim.x + threadIdx.x;
  if (ch >= c) return;
  const float mean = stats[ch * 2], rstd = stats[ch * 2 + 1];
  const float var = fmaxf(1.f / (rstd * rstd) - eps, 0.f);
  running_mean[ch] = (1.f - momentum) * running_mean[ch] + momentum * mean;
  running_var[ch] = (1.f - momentum) * running_var[ch] + momentum * var * unbias;
}

// Bernoulli(keep) byte mask from a counter-based hash (splitmix64 of seed + element index): reproducible per call
// seed_dev != nullptr: the seed is read from device memory and `seed` is added to it (a captured step draws fresh
// masks on every replay: the host writes the next seeds before it launches the graph)
__global__ void dropout_mask_kernel(uint8_t* __restrict__ mask, int64_t count, uint64_t seed,
                                    const uint64_t* __restrict__ seed_dev, float keep) {
  if (seed_dev != nullptr) seed += *seed_dev;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < count; i += stride) {
    uint64_t z = seed + 0x9E3779B97F4A7C15ull * static_cast<uint64_t>(i + 1);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    const float u = static_cast<float>(z >> 40) * (1.0f / 16777216.0f);
    mask[i] = u < keep ? 1 : 0;
  }
}

// 2x2 max pooling, stride 2 (nn.MaxPool2d(2), model_architectures.py:556): one thread per (output pixel, 8 channels)
__global__ void __launch_bounds__(256)
maxpool2_kernel(BView x, BView y) {
  const int G = y.c / 8;
  const int64_t total = static_cast<int64_t>(y.n) * y.h * y.w * G;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total; idx += stride) {
    const int g = static_cast<int>(idx % G);
    int64_t r = idx / G;
    const int ox = static_cast<int>(r % y.w);
    r /= y.w;
    const int oy = static_cast<int>(r % y.h);
    const int n = static_cast<int>(r / y.h);
    const int64_t in0 = (static_cast<int64_t>(n) * x.h + 2 * oy) * x.w + 2 * ox;
    float a[8], b[8], c[8], d[8], o[8];
    const __nv_bfloat16* xp = static_cast<const __nv_bfloat16*>(x.p) + g * 8;
    bn_load8(xp + x.at(in0), a);
    bn_load8(xp + x.at(in0 + 1), b);
    bn_load8(xp + x.at(in0 + x.w), c);
    bn_load8(xp + x.at(in0 + x.w + 1), d);
#pragma unroll
    for (int k = 0; k < 8; ++k) o[k] = fmaxf(fmaxf(a[k], b[k]), fmaxf(c[k], d[k]));
    bn_store8(static_cast<__nv_bfloat16*>(y.p) + y.at((static_cast<int64_t>(n) * y.h + oy) * y.w + ox) + g * 8, o);
  }
}

static bool bn_geometry_ok(const fpg_act* a) {
  return a != nullptr && a->halo == 0 && a->c % 8 == 0 && a->c >= 8 && 256 % (a->c / 8) == 0 && !a->fp32;
}

static int bn_grid(int64_t npix, int lanes, int sms) {
  int64_t blocks = (npix + lanes - 1) / lanes;
  const int64_t cap = 8ll * (sms > 0 ? sms : 148);
  if (blocks > cap) blocks = cap;
  return static_cast<int>(blocks < 1 ? 1 : blocks);
}

}  // namespace fpg

using namespace fpg;

#define FPG_ST(stream) static_cast<cudaStream_t>(stream)

extern "C" {

int fpg_batchnorm_apply(const fpg_act* y, const float* stats, const float* gamma, const float* beta,
                        const uint8_t* mask, float mask_scale, int act1, const fpg_act* z1, int act2,
                        const fpg_act* z2, void* stream) {
  FPG_REQUIRE(bn_geometry_ok(y) && bn_geometry_ok(z1) && (!z2 || bn_geometry_ok(z2)), "unsupported geometry");
  FPG_REQUIRE(z1->c == y->c && z1->n == y->n && z1->h == y->h && z1->w == y->w, "z1 geometry");
  FPG_REQUIRE(!z2 || (z2->c == y->c && z2->n == y->n && z2->h == y->h && z2->w == y->w), "z2 geometry");
  const int lanes = 256 / (y->c / 8);
  const int64_t npix = static_cast<int64_t>(y->n) * y->h * y->w;
  bn_apply_kernel<<<bn_grid(npix, lanes, sm_count_cached()), 256, 0, FPG_ST(stream)>>>(
      bview_of(y), stats, gamma, beta, mask, mask_scale, act1, bview_of(z1), act2, z2 ? bview_of(z2) : bview_of(z1),
      z2 != nullptr);
  FPG_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int64_t fpg_batchnorm_scratch_floats(const fpg_act* y) { return static_cast<int64_t>(kBnBlocks + 1) * y->c * 2; }

int fpg_batchnorm_bwd(const fpg_act* dz1, int act1, const fpg_act* dz2, int act2, const uint8_t* mask,
                      float mask_scale, const fpg_act* y, const float* stats, const float* gamma, const float* beta,
                      const fpg_act* dy, float* dgamma, float* dbeta, int accumulate, float* scratch, void* stream) {
  FPG_REQUIRE(bn_geometry_ok(y) && bn_geometry_ok(dz1) && bn_geometry_ok(dy) && (!dz2 || bn_geometry_ok(dz2)),
              "unsupported geometry");
  FPG_REQUIRE(dz1->c == y->c && dy->c == y->c && (!dz2 || dz2->c == y->c), "channel mismatch");
  const int lanes = 256 / (y->c / 8);
  const int64_t npix = static_cast<int64_t>(y->n) * y->h * y->w;
  BView v2 = dz2 ? bview_of(dz2) : bview_of(dz1);
  float* sums = nullptr;
  if (stats != nullptr) {
    FPG_REQUIRE(scratch != nullptr, "null scratch");
    int blocks = kBnBlocks;
    if (npix / blocks < lanes) blocks = static_cast<int>(npix / lanes > 0 ? npix / lanes : 1);
    float* partial = scratch;
    sums = scratch + static_cast<int64_t>(kBnBlocks) * y->c * 2;
    bn_bwd_reduce_kernel<<<blocks, 256, 0, FPG_ST(stream)>>>(bview_of(dz1), act1, v2, dz2 != nullptr, act2, mask,
                                                            mask_scale, bview_of(y), stats, gamma, beta, partial);
    bn_bwd_finalize_kernel<<<(y->c + 7) / 8, 256, 0, FPG_ST(stream)>>>(partial, blocks, y->c, sums, dgamma, dbeta,
                                                                        accumulate);
  }
  bn_bwd_apply_kernel<<<bn_grid(npix, lanes, sm_count_cached()), 256, 0, FPG_ST(stream)>>>(
      bview_of(dz1), act1, v2, dz2 != nullptr, act2, mask, mask_scale, bview_of(y), stats, gamma, beta, sums,
      1.f / static_cast<float>(npix), bview_of(dy));
  FPG_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int fpg_batchnorm_running_update(const float* stats, int32_t c, int64_t count, float eps, float momentum,
                                 float* running_mean, float* running_var, void* stream) {
  FPG_REQUIRE(stats && running_mean && running_var && c > 0 && count > 0, "bad argument");
  const float unbias = count > 1 ? static_cast<float>(count) / static_cast<float>(count - 1) : 1.f;
  bn_running_kernel<<<(c + 127) / 128, 128, 0, FPG_ST(stream)>>>(stats, eps, momentum, unbias, c, running_mean,
                                                                 running_var);
  FPG_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int fpg_maxpool2(const fpg_act* x, const fpg_act* y, void* stream) {
  FPG_REQUIRE(x && y && x->halo == 0 && y->halo == 0 && !x->fp32 && !y->fp32 && x->c == y->c && x->c % 8 == 0 &&
                  x->n == y->n && x->h == 2 * y->h && x->w == 2 * y->w, "maxpool geometry");
  const int64_t total = static_cast<int64_t>(y->n) * y->h * y->w * (y->c / 8);
  int64_t blocks = (total + 255) / 256;
  if (blocks > 4736) blocks = 4736;
  maxpool2_kernel<<<static_cast<unsigned>(blocks), 256, 0, FPG_ST(stream)>>>(bview_of(x), bview_of(y));
  FPG_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int fpg_dropout_mask(uint8_t* mask, int64_t count, uint64_t seed, float keep, void* stream) {
  FPG_REQUIRE(mask && count > 0 && keep > 0.f && keep <= 1.f, "bad argument");
  int64_t blocks = (count + 255) / 256;
  if (blocks > 1184) blocks = 1184;
  dropout_mask_kernel<<<static_cast<unsigned>(blocks), 256, 0, FPG_ST(stream)>>>(mask, count, seed, nullptr, keep);
  FPG_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int fpg_dropout_mask_dev(uint8_t* mask, int64_t count, const uint64_t* seed_dev, uint64_t seed_add, float keep,
                         void* stream) {
  FPG_REQUIRE(mask && seed_dev && count > 0 && keep > 0.f && keep <= 1.f, "bad argument");
  int64_t blocks = (count + 255) / 256;
  if (blocks > 1184) blocks = 1184;
  dropout_mask_kernel<<<static_cast<unsigned>(blocks), 256, 0, FPG_ST(stream)>>>(mask, count, seed_add, seed_dev, keep);
  FPG_CUDA_CHECK(cudaGetLastError());
  return 0;
}

}  // extern "C"
