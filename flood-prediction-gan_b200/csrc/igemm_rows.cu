// Row-stationary implicit-GEMM convolution for stride-1 R x S filters on wide images (the 7x7 stem and heads of the
// generators, model_architectures.py:312,328 and their data gradients).
//
// The tiled kernel (igemm.cu) re-reads the 128-pixel A tile once per filter tap: 49 x for a 7x7 filter, which makes
// those layers L2-bandwidth bound at ~1/6 of the tensor peak. Here a CTA tile is TH output rows x 128 pixels, one TMEM
// accumulator per output row, and
//   * every input row of the (TH + R - 1) x (128 + S - 1) halo patch is loaded ONCE, as one TMA box, and feeds up to
//     TH x S x (CBLK/16) MMAs: accumulator a uses it as filter row r = j - a, and the S column taps are the same
//     shared-memory rows addressed through descriptors that start s pixel rows later (the swizzle is a function of the
//     shared-memory address, so a descriptor may start at any pixel row of a TMA-written box: tools/exp/desc_shift.cu);
//   * every filter row of B (S taps x N x CBLK) is loaded once per tile and used by all TH accumulators; if the whole
//     filter fits next to the A ring it is loaded once per CTA;
//   * a shared-memory-operand tcgen05.mma of K = 16 costs max(N/2, 32 + N/4) cycles (tools/exp/mma_rate.cu: the
//     128 x 16 A slice and the N x 16 B slice are read at 128 B/clk), so N = 32 or 64 MMAs run at 40 / 67 % of the
//     tensor peak. The accumulators that share an A window (input row j, column tap s) are therefore STACKED along N:
//     they sit at consecutive TMEM columns, and filter rows r = j - a are kept at consecutive (descending-r) slots of
//     a per-column-tap ring, so ONE MMA with N = (#accumulators) x block_n serves all of them (two where the ring
//     wraps). A tile starts with one MMA against an all-zero B that clears every accumulator; all others accumulate.
// Warp roles as in igemm.cu: warp 0 = TMA producer, warp 1 = MMA issuer, warps 2..5 = epilogue.
#include "common.cuh"
#include "host_util.h"

namespace fpg {

constexpr int kRowsThreads = 192;

struct RowsArgs {
  int32_t block_n, rows, cols, dy0, dx0, tile_rows;
  int32_t n_img, tiles_y, tiles_x;
  int32_t act, a_stages, b_stages, b_resident;
  uint32_t a_slot_bytes;  // slot pitch of the A ring (box bytes rounded up to 1 KB)
  uint32_t a_box_bytes;   // bytes one A box delivers
  uint32_t b_tap_bytes;   // block_n * CBLK * 2
  const float* bias;
  fpg_out_view out;
  int16_t tap_of[FPG_MAX_TAPS];
};

__device__ __forceinline__ float rows_act(float v, int act) {
  switch (act) {
    case FPG_ACT_RELU: return fmaxf(v, 0.f);
    case FPG_ACT_LEAKY: return v > 0.f ? v : 0.2f * v;
    case FPG_ACT_TANH: return tanhf(v);
    default: return v;
  }
}

template <int CBLK>
__global__ void __launch_bounds__(kRowsThreads, 1)
igemm_rows_kernel(const __grid_constant__ CUtensorMap amap, const __grid_constant__ CUtensorMap bmap,
                  const __grid_constant__ RowsArgs args) {
  constexpr uint32_t LAYOUT = swizzle_layout_type(CBLK * 2);
  constexpr uint32_t SBO = 8u * CBLK * 2u;
  constexpr uint32_t PIX_BYTES = CBLK * 2u;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));

  const int BN = args.block_n;
  const int R = args.rows, S = args.cols, TH = args.tile_rows;
  const int ASTG = args.a_stages, BSTG = args.b_stages;
  const uint32_t B_SLOT_BYTES = args.b_tap_bytes * static_cast<uint32_t>(S);     // one filter row (all column taps)
  const uint32_t B_COL_BYTES = args.b_tap_bytes * static_cast<uint32_t>(BSTG);  // ring of one column tap
  const uint32_t ACC_COLS = static_cast<uint32_t>(TH * BN);

  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + ASTG * args.a_slot_bytes;
  uint8_t* smem_z = smem_b + BSTG * B_SLOT_BYTES;  // 1 KB of zeros: B operand that clears the accumulators
  uint64_t* full_a = reinterpret_cast<uint64_t*>(smem_z + 1024);
  uint64_t* empty_a = full_a + ASTG;
  uint64_t* full_b = empty_a + ASTG;
  uint64_t* empty_b = full_b + BSTG;
  uint64_t* tfull = empty_b + BSTG;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int i = 0; i < ASTG; ++i) {
      mbar_init(&full_a[i], 1);
      mbar_init(&empty_a[i], 1);
    }
    for (int i = 0; i < BSTG; ++i) {
      mbar_init(&full_b[i], 1);
      mbar_init(&empty_b[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], 128);
    }
    fence_barrier_init();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&amap);
    tma_prefetch_desc(&bmap);
  }
  for (int i = threadIdx.x; i < 256; i += kRowsThreads) reinterpret_cast<uint32_t*>(smem_z)[i] = 0u;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy zeros visible to the tensor core
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int total_tiles = args.n_img * args.tiles_y * args.tiles_x;
  const int steps = TH + R - 1;  // input rows per tile

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (elect_one()) {
      uint32_t as = 0, aph = 0, bs = 0, bph = 0;
      bool first = true;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        int tx = tile % args.tiles_x;
        int r = tile / args.tiles_x;
        int ty = r % args.tiles_y;
        int n = r / args.tiles_y;
        const int x0 = tx * 128 + args.dx0, y0 = ty * TH + args.dy0;
        for (int j = 0; j < steps; ++j) {
          if (j < R && (first || !args.b_resident)) {
            mbar_wait(&empty_b[bs], bph ^ 1);
            mbar_arrive_expect_tx(&full_b[bs], B_SLOT_BYTES);
            // filter rows sit in DESCENDING order inside each column tap's ring: data slot = BSTG - 1 - ring index
            uint8_t* dst = smem_b + (BSTG - 1 - bs) * args.b_tap_bytes;
            for (int s = 0; s < S; ++s)
              tma_load_2d(&bmap, &full_b[bs], dst + s * B_COL_BYTES, args.tap_of[j * S + s] * CBLK, 0);
            if (++bs == static_cast<uint32_t>(BSTG)) {
              bs = 0;
              bph ^= 1;
            }
          }
          mbar_wait(&empty_a[as], aph ^ 1);
          mbar_arrive_expect_tx(&full_a[as], args.a_box_bytes);
          tma_load_5d(&amap, &full_a[as], smem_a + as * args.a_slot_bytes, 0, x0, 0, y0 + j, n);
          if (++as == static_cast<uint32_t>(ASTG)) {
            as = 0;
            aph ^= 1;
          }
        }
        first = false;
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    if (elect_one()) {
      const uint32_t idesc1 = make_idesc_bf16(128, BN, 0, 0);  // N field scaled up for stacked accumulators
      const uint64_t desc_hi = make_smem_desc(0, 0, SBO, LAYOUT);  // everything but the start address
      const uint64_t zdesc = make_smem_desc(smem_u32(smem_z), 0, 0, LAYOUT);
      const uint32_t b_lo = (smem_u32(smem_b) >> 4) & 0x3FFFu;
      uint32_t as = 0, aph = 0, it = 0;
      uint32_t bcount = 0;  // filter rows consumed before this tile (ring position of filter row 0)
      bool first = true;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
        const uint32_t acs = it & 1, acph = (it >> 1) & 1;
        mbar_wait(&tempty[acs], acph ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acs * ACC_COLS;
        for (int j = 0; j < steps; ++j) {
          mbar_wait(&full_a[as], aph);
          if (j < R && (first || !args.b_resident)) {
            const uint32_t idx = bcount + j;
            mbar_wait(&full_b[idx % BSTG], (idx / BSTG) & 1);
          }
          tc_fence_after();
          const uint32_t a_lo = (smem_u32(smem_a + as * args.a_slot_bytes) >> 4) & 0x3FFFu;
          if (j == 0) {
            // zero all TH accumulators with one MMA against the all-zero B region (8 aliased rows: SBO = 0), so that
            // every other MMA of the tile accumulates and the stacked issue loop below has no special cases
            umma_bf16(d_tmem, desc_hi | static_cast<uint64_t>(a_lo), zdesc,
                      idesc1 + (static_cast<uint32_t>((TH - 1) * BN >> 3) << 17), 0u);
          }
          // accumulators using this input row: a in [a0, a1], filter row r = j - a. Stacked along N in order of a,
          // i.e. descending r == ascending data slots starting at the slot of r = j - a0; a second MMA covers the
          // part of the stack that wraps around the ring.
          const int a1 = j < TH - 1 ? j : TH - 1;
          const int a0 = j - (R - 1) > 0 ? j - (R - 1) : 0;
          const int n_acc = a1 - a0 + 1;
          const uint32_t ring0 = args.b_resident ? static_cast<uint32_t>(j - a0) : (bcount + (j - a0)) % BSTG;
          const uint32_t slot0 = BSTG - 1 - ring0;
          const int seg1 = n_acc < static_cast<int>(BSTG - slot0) ? n_acc : static_cast<int>(BSTG - slot0);
          const int seg2 = n_acc - seg1;
          const uint32_t d1 = d_tmem + a0 * BN, d2 = d_tmem + (a0 + seg1) * BN;
          const uint32_t id1 = idesc1 + (static_cast<uint32_t>((seg1 - 1) * BN >> 3) << 17);
          const uint32_t id2 = idesc1 + (static_cast<uint32_t>((seg2 > 0 ? seg2 - 1 : 0) * BN >> 3) << 17);
          const uint32_t b_off1 = b_lo + ((slot0 * args.b_tap_bytes) >> 4);
          const uint32_t b_col16 = B_COL_BYTES >> 4;
          if (seg2 == 0) {
            for (int s = 0; s < S; ++s) {
              const uint32_t a_s = a_lo + s * (PIX_BYTES >> 4), b_s = b_off1 + s * b_col16;
#pragma unroll
              for (int k = 0; k < CBLK / 16; ++k)
                umma_bf16(d1, desc_hi | static_cast<uint64_t>(a_s + 2 * k), desc_hi | static_cast<uint64_t>(b_s + 2 * k),
                          id1, 1u);
            }
          } else {
            for (int s = 0; s < S; ++s) {
              const uint32_t a_s = a_lo + s * (PIX_BYTES >> 4), b_s = b_off1 + s * b_col16, b_w = b_lo + s * b_col16;
#pragma unroll
              for (int k = 0; k < CBLK / 16; ++k) {
                const uint64_t ad = desc_hi | static_cast<uint64_t>(a_s + 2 * k);
                umma_bf16(d1, ad, desc_hi | static_cast<uint64_t>(b_s + 2 * k), id1, 1u);
                umma_bf16(d2, ad, desc_hi | static_cast<uint64_t>(b_w + 2 * k), id2, 1u);
              }
            }
          }
          umma_commit(&empty_a[as]);
          if (!args.b_resident && j >= TH - 1) {  // filter row j - (TH - 1) has served its last accumulator
            const uint32_t idx = bcount + (j - (TH - 1));
            umma_commit(&empty_b[idx % BSTG]);
          }
          if (++as == static_cast<uint32_t>(ASTG)) {
            as = 0;
            aph ^= 1;
          }
        }
        umma_commit(&tfull[acs]);
        if (!args.b_resident) bcount += R;
        first = false;
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue (warps 2..5)
    const int q = warp & 3;
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
      const uint32_t acs = it & 1, acph = (it >> 1) & 1;
      int tx = tile % args.tiles_x;
      int r = tile / args.tiles_x;
      int ty = r % args.tiles_y;
      int n = r / args.tiles_y;
      const int px = tx * 128 + q * 32 + lane;
      mbar_wait(&tfull[acs], acph);
      tc_fence_after();
      for (int a = 0; a < TH; ++a) {
        const int py = ty * TH + a;
        const bool valid = (py < args.out.valid_h) && (px < args.out.valid_w);
        const int64_t off = static_cast<int64_t>(n) * args.out.stride_n +
                            static_cast<int64_t>(py * args.out.mul_y + args.out.off_y) * args.out.stride_y +
                            static_cast<int64_t>(px * args.out.mul_x + args.out.off_x) * args.out.stride_x;
        const uint32_t t_addr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acs * ACC_COLS + a * BN;
        for (int c = 0; c < BN; c += 16) {
          uint32_t v[16];
          tmem_ld16(t_addr + c, v);
          tmem_ld_wait();
          float f[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) f[i] = __uint_as_float(v[i]);
          if (args.bias != nullptr) {
            const float4* bp = reinterpret_cast<const float4*>(args.bias + c);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float4 b4 = __ldg(bp + i);
              f[4 * i + 0] += b4.x;
              f[4 * i + 1] += b4.y;
              f[4 * i + 2] += b4.z;
              f[4 * i + 3] += b4.w;
            }
          }
          if (args.act != FPG_ACT_NONE) {
#pragma unroll
            for (int i = 0; i < 16; ++i) f[i] = rows_act(f[i], args.act);
          }
          if (valid) {
            if (args.out.fp32 == FPG_DT_FP32) {
              store_16x32(static_cast<float*>(args.out.base) + off + c, f);
            } else {
              store_16x16(static_cast<__nv_bfloat16*>(args.out.base) + off + c, f, args.out.fp32);
            }
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&tempty[acs]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

}  // namespace fpg

using namespace fpg;

extern "C" int fpg_igemm_rows_launch(const fpg_igemm_rows_desc* d, void* stream) {
  FPG_REQUIRE(d != nullptr, "null descriptor");
  FPG_REQUIRE(d->cblk == 16 || d->cblk == 32 || d->cblk == 64, "cblk %d", d->cblk);
  FPG_REQUIRE(d->block_n >= 16 && d->block_n <= 128 && d->block_n % 16 == 0, "block_n %d", d->block_n);
  FPG_REQUIRE(d->rows >= 1 && d->cols >= 1 && d->rows * d->cols <= FPG_MAX_TAPS, "filter %dx%d", d->rows, d->cols);
  FPG_REQUIRE(d->tile_rows >= 1 && 2 * d->tile_rows * d->block_n <= 512, "tile_rows %d", d->tile_rows);
  FPG_REQUIRE(d->a_stages >= 2 && d->a_stages <= 16, "a_stages %d", d->a_stages);
  FPG_REQUIRE(d->b_stages >= d->tile_rows + 1 || d->b_stages == d->rows, "b_stages %d", d->b_stages);
  FPG_REQUIRE(static_cast<int>(d->a.box[1]) == 128 + d->cols - 1 && static_cast<int>(d->a.box[0]) == d->cblk,
              "A box %ux%u", d->a.box[0], d->a.box[1]);
  CUtensorMap amap, bmap;
  int rc = encode_tmap(&d->a, &amap);
  if (rc) return rc;
  rc = encode_tmap(&d->b, &bmap);
  if (rc) return rc;

  RowsArgs args;
  args.block_n = d->block_n;
  args.rows = d->rows;
  args.cols = d->cols;
  args.dy0 = d->dy0;
  args.dx0 = d->dx0;
  args.tile_rows = d->tile_rows;
  args.n_img = d->n_img;
  args.tiles_y = d->tiles_y;
  args.tiles_x = d->tiles_x;
  args.act = d->act;
  args.a_stages = d->a_stages;
  args.b_stages = d->b_stages;
  args.b_resident = d->b_stages == d->rows ? 1 : 0;
  args.a_box_bytes = static_cast<uint32_t>(128 + d->cols - 1) * d->cblk * 2u;
  args.a_slot_bytes = (args.a_box_bytes + 1023u) & ~1023u;
  args.b_tap_bytes = static_cast<uint32_t>(d->block_n) * d->cblk * 2u;
  args.bias = d->bias;
  args.out = d->out;
  for (int i = 0; i < FPG_MAX_TAPS; ++i) args.tap_of[i] = d->tap_of[i];

  const size_t smem = static_cast<size_t>(d->a_stages) * args.a_slot_bytes +
                      static_cast<size_t>(d->b_stages) * d->cols * args.b_tap_bytes +
                      1024 + (2 * d->a_stages + 2 * d->b_stages + 4) * 8 + 16 + 1024;
  FPG_REQUIRE(smem <= 227 * 1024, "shared memory %zu", smem);
  const int total_tiles = d->n_img * d->tiles_y * d->tiles_x;
  const int sms = sm_count_cached();
  if (sms <= 0) return fail(FPG_ENOTSUP, "no CUDA device");
  const int grid = total_tiles < sms ? total_tiles : sms;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
#define FPG_LAUNCH_ROWS(CB)                                                                                    \
  do {                                                                                                         \
    FPG_CUDA_CHECK(cudaFuncSetAttribute(igemm_rows_kernel<CB>, cudaFuncAttributeMaxDynamicSharedMemorySize,    \
                                        static_cast<int>(smem)));                                              \
    FPG_CUDA_CHECK(launch_persistent(igemm_rows_kernel<CB>, dim3(grid), dim3(kRowsThreads), smem, st, amap, bmap, \
                                     args));                                                                   \
  } while (0)
  if (d->cblk == 64) {
    FPG_LAUNCH_ROWS(64);
  } else if (d->cblk == 32) {
    FPG_LAUNCH_ROWS(32);
  } else {
    FPG_LAUNCH_ROWS(16);
  }
#undef FPG_LAUNCH_ROWS
  FPG_CUDA_CHECK(cudaGetLastError());
  return 0;
}
